"""GPU parity of the eval / feedback decoding paths: greedy token ids must match the reference bit for bit."""
import numpy as np
import pytest
import torch

from tests.golden_util import relerr
from tests.gpu_util import FixtureGlove, grads_of, load_case, to_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["s2vtatt_tiny", "s2vtatt_mid"])
def test_s2vtatt_greedy_ids_bit_exact(tag):
    from pvcr_b200.model import S2VTAttModel
    d, params, _, (B, N, V, H, E, L, Vc) = load_case(tag)
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    vid = torch.from_numpy(d["vid"]).cuda()
    logits = m(vid, None)                                   # reference API: eval forward returns logits
    assert np.array_equal(torch.argmax(logits, dim=2).cpu().numpy(), d["greedy_ids"])
    ids, _ = m.greedy(vid)
    assert np.array_equal(ids.cpu().numpy(), d["greedy_ids"])
    assert relerr(logits.cpu().numpy(), d["greedy_logits"]) < 2e-5
    assert np.abs(m.last_alphas.cpu().numpy() - d["greedy_alphas"]).max() < 1e-5


@pytest.mark.parametrize("tag", ["s2vt_tiny", "s2vt_mid", "s2vt_sched"])
def test_s2vt_greedy_ids_bit_exact(tag):
    from pvcr_b200.model import S2VTModel
    d, params, _, (B, N, V, H, E, L, Vc) = load_case(tag)
    m = to_cuda(S2VTModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    logits = m(torch.from_numpy(d["vid"]).cuda(), None)
    assert np.array_equal(torch.argmax(logits, dim=2).cpu().numpy(), d["greedy_ids"])
    assert relerr(logits.cpu().numpy(), d["greedy_logits"]) < 2e-5


def test_s2vt_scheduled_sampling(monkeypatch):
    """teacher_force_prob < 1: the per-step coin (Python RNG, as in the reference) is scripted to the golden sequence."""
    from pvcr_b200.model import S2VTModel
    import sys
    mod = sys.modules["pvcr_b200.model.S2VTModel"]
    d, params, g, (B, N, V, H, E, L, Vc) = load_case("s2vt_sched")
    m = to_cuda(S2VTModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision="bf16x3"), params).train()
    m.teacher_force_prob = 0.5
    seq = iter([0.0 if t else 1.0 for t in d["teacher"]])
    monkeypatch.setattr(mod.random, "random", lambda: next(seq))
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    s_len = torch.from_numpy(d["s_len"]).cuda()
    loss, acc, pred = m.forward_loss(vid, s, s_len)
    loss.backward()
    assert abs(loss.item() - float(d["loss"])) < 2e-6 * abs(float(d["loss"]))
    assert np.array_equal(pred.cpu().numpy(), d["pred"])
    got = grads_of(m)
    for k in g:
        assert relerr(got[k], g[k]) < 1e-4, (k, relerr(got[k], g[k]))


def test_greedy_msrvtt_shape_vs_oracle():
    """cfg5 dims at B=16 with a reduced vocabulary: ids vs the float64 oracle (bit exact up to near-ties)."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200.model import S2VTAttModel
    B, N, V, H, E, L, Vc = 16, 40, 2048, 512, 300, 30, 3000
    p = W.s2vtatt_params(V, H, E, Vc, 91)
    vid, _, _ = W.make_batch(B, N, V, L, Vc, 92)
    ids_ref, logits_ref, _ = O.s2vtatt_greedy({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64),
                                              Vc - 4, L)
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), p).eval()
    ids, logits = m.greedy(torch.from_numpy(vid).cuda())
    assert np.array_equal(ids.cpu().numpy(), ids_ref)
    assert relerr(logits.cpu().numpy(), logits_ref) < 2e-5


def test_graphed_greedy_matches_eager():
    from pvcr_b200.graphs import GraphedGreedy
    from pvcr_b200.model import S2VTAttModel
    d, params, _, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_mid")
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    vid = torch.from_numpy(d["vid"]).cuda()
    g = GraphedGreedy(m, vid)
    ids, logits = g(vid)
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy(), d["greedy_ids"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["s2vtatt_tiny", "s2vtatt_mid"])
@pytest.mark.parametrize("K", [1, 3, 5])
def test_s2vtatt_beam_search_matches_reference_driven_golden(tag, K):
    """Beam search: token ids bit-exact and scores to 1e-5 against the search driven over the reference's
    own Encoder / Decoder.forward_step (oracle/gen_golden_beam.py); beam 1 == the greedy ids of the reference."""
    import os
    from pvcr_b200.model import S2VTAttModel
    from tests.golden_util import GOLDEN
    d, params, _, (B, N, V, H, E, L, Vc) = load_case(tag)
    z = np.load(os.path.join(GOLDEN, tag.replace("s2vtatt_", "s2vtatt_beam_") + ".npz"))
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    ids, scores = m.beam_search(torch.from_numpy(d["vid"]).cuda(), beam=K)
    assert ids.shape == (B, K, L) and scores.shape == (B, K)
    assert np.array_equal(ids.cpu().numpy(), z["ids_k%d" % K])
    assert np.abs(scores.double().cpu().numpy() - z["score_k%d" % K]).max() < 1e-4
    if K == 1:
        assert np.array_equal(ids[:, 0].cpu().numpy(), d["greedy_ids"])


@pytest.mark.gpu
def test_greedy_at_full_hidden_size_matches_oracle():
    """Step-wise decoding at H = 512, N = 40 (the shape class of BASELINE's configs; the golden fixtures are small): this
    is where the vectorised per-step attention kernel runs.  Token ids equal, attention weights 1e-5, logits 1e-4 against
    the float64 oracle (vocabulary reduced so numpy finishes in seconds)."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200.model import S2VTAttModel
    B, N, V, H, E, L, Vc = 6, 40, 256, 512, 300, 8, 400
    p = W.s2vtatt_params(V, H, E, Vc, 41)
    vid, _, _ = W.make_batch(B, N, V, L, Vc, 42)
    ids_o, logits_o, alphas_o = O.s2vtatt_greedy({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64),
                                                 Vc - 4, L)
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), p).eval()
    ids, logits = m.greedy(torch.from_numpy(vid).cuda())
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert np.abs(m.last_alphas.double().cpu().numpy() - alphas_o).max() < 1e-5
    assert relerr(logits.double().cpu().numpy(), logits_o) < 1e-4
    # beam 1 runs the same step through the beam-search entry point
    ids_b, _ = m.beam_search(torch.from_numpy(vid).cuda(), beam=1)
    assert np.array_equal(ids_b[:, 0].cpu().numpy(), ids_o)


@pytest.mark.gpu
def test_graphed_beam_search_equals_eager():
    from pvcr_b200.graphs import GraphedBeam
    from pvcr_b200.model import S2VTAttModel
    d, params, _, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_mid")
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    vid = torch.from_numpy(d["vid"]).cuda()
    ids, scores = m.beam_search(vid, beam=3)
    g = GraphedBeam(m, vid, beam=3)
    ids_g, scores_g = g(vid)
    torch.cuda.synchronize()
    assert torch.equal(ids, ids_g) and torch.equal(scores, scores_g)


def test_decode_plan_reuses_and_invalidates():
    """The weights prepared for decoding (split planes + the word table W_e Emb[w] + b_ih) are kept between calls:
    a second call on the same parameters skips the preparation and must return the same bits; a parameter update that
    autograd can see (in-place op -> version counter) must be noticed; ids-only decoding returns the same ids."""
    from pvcr_b200.model import S2VTAttModel
    d, params, _, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_mid")
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    vid = torch.from_numpy(d["vid"]).cuda()
    ids0, logits0 = m.greedy(vid)
    plan = m._plan("greedy")
    key0 = plan.key
    assert key0 is not None
    ids1, logits1 = m.greedy(vid)                       # prepared weights reused (PVCR_DECODE_REUSE_PREPARED)
    assert plan.key == key0
    # (split-K partial sums are added atomically: the logits of two runs may differ in their last bits)
    assert torch.equal(ids0, ids1) and relerr(logits1.cpu().numpy(), logits0.cpu().numpy()) < 1e-6
    assert np.array_equal(ids1.cpu().numpy(), d["greedy_ids"])
    ids2, none = m.greedy(vid, return_logits=False)
    assert none is None and torch.equal(ids2, ids0)
    # another batch on the same prepared weights: the per-batch half of the workspace must not leak between calls
    vid_b = torch.flip(vid, dims=[0]).contiguous()
    ids_b, logits_b = m.greedy(vid_b)
    assert torch.equal(ids_b, torch.flip(ids0, dims=[0]))
    assert relerr(logits_b.cpu().numpy(), np.flip(d["greedy_logits"], axis=0)) < 2e-5
    # an in-place parameter update changes the fingerprint: the table and planes are rebuilt
    with torch.no_grad():
        m.decoder.embedding.weight.mul_(0.5)
    ids3, logits3 = m.greedy(vid)
    assert plan.key != key0
    fresh = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    with torch.no_grad():
        fresh.decoder.embedding.weight.mul_(0.5)
    ids4, logits4 = fresh.greedy(vid)
    assert torch.equal(ids3, ids4) and relerr(logits3.cpu().numpy(), logits4.cpu().numpy()) < 1e-6
    assert relerr(logits3.cpu().numpy(), logits0.cpu().numpy()) > 1e-3
    # train() drops the cache
    m.train()
    assert plan.key is None


def test_greedy_ex_reuse_flag_c_abi():
    """pvcr_s2vtatt_greedy_ex through ctypes: prepare + run, then run with PVCR_DECODE_REUSE_PREPARED on the same
    workspace -> identical ids and logits; decode from given encoder outputs agrees with the model's decode()."""
    import ctypes
    from pvcr_b200 import _lib, functional as F_
    from pvcr_b200.model import S2VTAttModel
    d, params, _, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_tiny")
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).eval()
    vid = torch.from_numpy(d["vid"]).cuda()
    Lb = _lib.lib()
    tensors = {f: p.detach().contiguous() for f, p in zip(F_.ATT_SEQ_FIELDS, m._seq_params())}
    lin = m.decoder.pred_linear[1]
    tensors["out_w"], tensors["out_b"] = lin.weight.detach().contiguous(), lin.bias.detach().contiguous()
    ps = F_._fill_struct(_lib.PvcrS2vtAttParams(), _lib.ATT_PARAM_FIELDS, tensors)
    dims = F_.make_dims(B, N, V, H, E, L, Vc, 3, 0.0, 0)
    ws = torch.empty(int(Lb.pvcr_s2vtatt_greedy_workspace(ctypes.byref(dims))), dtype=torch.uint8, device="cuda")
    outs = []
    for flags in (0, 1, 1):
        ids = torch.empty((B, L), dtype=torch.int64, device="cuda")
        logits = torch.empty((B, L, Vc), dtype=torch.float32, device="cuda")
        _lib.check(Lb.pvcr_s2vtatt_greedy_ex(ctypes.byref(dims), ctypes.byref(ps), _lib.ptr(vid), None, None, None,
                                             m.decoder.sos_id, _lib.ptr(ids), _lib.ptr(logits), None, _lib.ptr(ws),
                                             ws.numel(), flags, _lib.stream_ptr()), "greedy_ex")
        outs.append((ids, logits))
    torch.cuda.synchronize()
    assert np.array_equal(outs[0][0].cpu().numpy(), d["greedy_ids"])
    for ids, logits in outs[1:]:
        assert torch.equal(ids, outs[0][0]) and relerr(logits.cpu().numpy(), outs[0][1].cpu().numpy()) < 1e-6
    # both inputs at once / neither is an argument error, not a crash
    rc = Lb.pvcr_s2vtatt_greedy_ex(ctypes.byref(dims), ctypes.byref(ps), None, None, None, None, 0, _lib.ptr(outs[0][0]),
                                   None, None, _lib.ptr(ws), ws.numel(), 0, _lib.stream_ptr())
    assert rc != 0


def _decided(ids, ids_ref, logits_ref, tol=1e-6):
    """ids equal to the float64 oracle's wherever the oracle is decided: a video may leave the reference sequence only at a
    step whose float64 top-2 margin is below `tol` (everything after that step is a different, equally valid, decode)."""
    for b in range(ids.shape[0]):
        for l in range(ids.shape[1]):
            if ids[b, l] != ids_ref[b, l]:
                top2 = np.sort(logits_ref[b, l])[-2:]
                assert top2[1] - top2[0] < tol, (b, l, ids[b, l], ids_ref[b, l], top2)
                break


@pytest.mark.parametrize("B,H", [(1, 64), (3, 128), (33, 128), (70, 256), (130, 64)])
def test_greedy_batch_and_width_sweep_vs_oracle(B, H):
    """The decode path's persistent fp32 encoder kernel (gru_f32_persist.cu) serves groups of 32 videos with H/16 CTAs each:
    batches that are not multiples of 32 (tail group partly empty), a single video, and B = 130 at H = 64 (5 groups);
    ids equal to the float64 oracle's, logits 2e-5."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200.model import S2VTAttModel
    N, V, E, L, Vc = 12, 96, 40, 6, 300
    p = W.s2vtatt_params(V, H, E, Vc, 300 + B)
    vid, _, _ = W.make_batch(B, N, V, L, Vc, 400 + B)
    ids_o, logits_o, alphas_o = O.s2vtatt_greedy({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64),
                                                 Vc - 4, L)
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), p).eval()
    ids, logits = m.greedy(torch.from_numpy(vid).cuda())
    _decided(ids.cpu().numpy(), ids_o, logits_o)
    if np.array_equal(ids.cpu().numpy(), ids_o):
        assert relerr(logits.double().cpu().numpy(), logits_o) < 2e-5
        assert np.abs(m.last_alphas.double().cpu().numpy() - alphas_o).max() < 1e-5


def test_greedy_bf16x2_compact_planes():
    """nsplit = 2 through the decode entry point: the vocabulary weight is stored as two compact planes (terms 0, 1) and
    the GEMM producer maps the three virtual K planes {0,1,0} to them.  fp32-level logits (3 products: 1e-4)."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200 import functional as F_
    from pvcr_b200.model import S2VTAttModel
    B, N, V, H, E, L, Vc = 9, 10, 64, 128, 40, 5, 700
    p = W.s2vtatt_params(V, H, E, Vc, 77)
    vid, _, _ = W.make_batch(B, N, V, L, Vc, 78)
    ids_o, logits_o, _ = O.s2vtatt_greedy({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64), Vc - 4, L)
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), p).eval()
    lin = m.decoder.pred_linear[1]
    with torch.no_grad():
        ids, logits, _ = F_.s2vtatt_greedy(torch.from_numpy(vid).cuda(), None, m.decoder.sos_id, L, m._seq_params(), lin.weight,
                                           lin.bias, nsplit=2)
    _decided(ids.cpu().numpy(), ids_o, logits_o, tol=1e-3)
    if np.array_equal(ids.cpu().numpy(), ids_o):
        assert relerr(logits.double().cpu().numpy(), logits_o) < 1e-4


def test_stepwise_encoder_fallback_agrees():
    """The step-wise tensor-core encoder (bf16x3) that larger batches fall back to and the fp32 persistent kernel give the
    same token ids (separate processes: the knob is read once)."""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    outs = []
    for env in ({}, {"PVCR_NO_F32_GRU": "1"}):
        r = subprocess.run([sys.executable, os.path.join(here, "gpu_knob_f32gru.py")], env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("IDS")][0].split()
        outs.append(line)
    assert outs[0][1] == outs[1][1], outs
    assert abs(float(outs[0][2]) - float(outs[1][2])) < 1e-5 * float(outs[0][2])
