"""GPU parity at the BENCHMARKED size (cfg2: B=128, N=40, V=2048, H=512, E=300, L=30, Vc=23000) against
tests/golden/full_cfg2_s2vtatt.npz, which oracle/gen_golden_full.py produced by running the UNMODIFIED reference
S2VTAttModel (float64) on the seeded inputs / weights of oracle/workloads.py: loss, per-token loss, arg-max
predictions, attention weights, every parameter gradient (norm, sub-sample, projections) and greedy token ids.
Also here: the dropout-on path with the kernels' own Philox mask handed to the oracle (and to the reference modules
when oracle/_ref travelled), the materialised-logits backward with dropout, and the Philox range check."""
import os

import numpy as np
import pytest
import torch

from tests.golden_util import GOLDEN, relerr
from tests.gpu_util import FixtureGlove, grads_of, to_cuda

pytestmark = pytest.mark.gpu

# per arithmetic mode: loss rel, per-token loss abs, attention weights abs, gradient rel (sub-sample estimate of the
# relative Frobenius error vs the float64 reference), fraction of arg-max predictions that must agree.
# bf16 (one bf16 product per logical product): every operand carries 2^-9 relative rounding, so ONE bf16 GEMM already
# differs from exact arithmetic by ~1.6e-3 relative Frobenius; the chain of the full model measures <= 8e-3.  north_star's
# "rel 1e-3 for bf16 GEMMs" is the fp32-ACCUMULATE tolerance: it is asserted against the oracle evaluated with the same
# operand rounding in test_bf16_mode_vs_operand_rounded_oracle_full_batch below.
# Last two: fraction of arg-max predictions that must agree, and the largest reference top-2 logit margin at which a flip
# is tolerated (random-init weights give near-uniform logits, so margins at the mode's logit error are common).
TOL = {"bf16x3": (2e-6, 2e-5, 1e-5, 2e-4, 1.0, 1e-5), "bf16x2": (1e-4, 1e-3, 1e-4, 1e-3, 0.999, 2e-3),
       "bf16": (1e-3, 2e-2, 1e-4, 1e-2, 0.98, 5e-2)}


def _golden():
    z = np.load(os.path.join(GOLDEN, "full_cfg2_s2vtatt.npz"))
    return {k: z[k] for k in z.files}


def _inputs(g):
    from oracle import workloads as W
    B, N, V, H, E, L, Vc = (int(x) for x in g["dims"])
    ps, bs, _ = (int(x) for x in g["seeds"])
    p = W.s2vtatt_params(V, H, E, Vc, ps)
    vid, s, s_len = W.make_batch(B, N, V, L, Vc, bs)
    return (B, N, V, H, E, L, Vc), p, vid, s, s_len


def _sample_index(numel, n=4096):
    stride = max(1, numel // n)
    return np.arange(0, numel, stride)[:n]


def _projection_dirs(name, numel, seed, k):
    rs = np.random.RandomState((seed + 31 * k + sum(map(ord, name))) % (2 ** 31))
    return rs.standard_normal(numel).astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16x2", "bf16"])
def test_full_cfg2_vs_reference_fixture(precision):
    from pvcr_b200.model import S2VTAttModel
    g = _golden()
    (B, N, V, H, E, L, Vc), p, vid, s, s_len = _inputs(g)
    t_loss, t_tok, t_alpha, t_grad, t_pred, t_flip = TOL[precision]
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision=precision), p).train()
    tv, ts, tl = torch.from_numpy(vid).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(s_len).cuda()
    loss, acc, pred = m.forward_loss(tv, ts, tl)
    loss.backward()
    torch.cuda.synchronize()
    ref_loss = float(g["loss"])
    assert abs(loss.item() - ref_loss) <= t_loss * abs(ref_loss), (loss.item(), ref_loss)
    # per-token loss (north_star): criterion(logits, target) of train_utils.py:47-48, all B*L tokens
    tok = m.last_token_nll.cpu().numpy()
    tok_err = float(np.abs(tok - g["token_nll"]).max())
    assert tok_err < t_tok, tok_err
    # arg-max predictions; a flip is only legitimate where the reference's own top-2 margin is at rounding level
    same = pred.cpu().numpy() == g["pred"]
    assert same.mean() >= t_pred, same.mean()
    if not same.all():
        assert float(g["train_margin"][~same].max()) < t_flip, float(g["train_margin"][~same].max())
    # attention weights: first 16 videos element-wise, every video's peak weight per step
    al = m.last_alphas.cpu().numpy()
    a_err = float(np.abs(al[:, :g["alphas_head"].shape[1]] - g["alphas_head"]).max())
    a_err = max(a_err, float(np.abs(al.max(axis=2) - g["alpha_max"]).max()))
    assert a_err < t_alpha, a_err
    # gradients
    _, _, proj_seed = (int(x) for x in g["seeds"])
    got = grads_of(m)
    errs, perrs = {}, {}
    for k, gv in got.items():
        flat = gv.reshape(-1)
        ref_s = g["gsample." + k].astype(np.float64)
        errs[k] = relerr(flat[_sample_index(flat.size)], ref_s)
        nrm = float(g["gnorm." + k])
        assert abs(np.linalg.norm(flat) - nrm) <= max(t_grad, 1e-6) * nrm, (k, np.linalg.norm(flat), nrm)
        pe = 0.0
        for j, ref_p in enumerate(g["gproj." + k]):
            pe = max(pe, abs(float(np.dot(flat, _projection_dirs(k, flat.size, proj_seed, j))) - ref_p) / nrm)
        perrs[k] = pe
    worst = max(errs, key=errs.get)
    print("\n[%s, full cfg2 vs reference] loss rel %.2e  token-loss abs %.2e  alphas abs %.2e  pred agree %.5f  "
          "grads rel max %.2e (%s)  projections max %.2e" % (precision, abs(loss.item() - ref_loss) / ref_loss, tok_err, a_err,
                                                             same.mean(), errs[worst], worst, max(perrs.values())))
    for k in sorted(errs, key=errs.get, reverse=True)[:5]:
        print("   %-45s %.2e" % (k, errs[k]))
    for k, e in errs.items():
        assert e < t_grad, (k, e)
    for k, e in perrs.items():          # |<g - g_ref, r>| / |g_ref| ~ N(0, rel^2): 4 sigma
        assert e < 4 * t_grad, (k, e)


def test_bf16_mode_vs_operand_rounded_oracle_full_batch():
    """north_star's bf16-GEMM tolerance at the benchmarked batch (B = 128; vocabulary cut to 3000 words so that the
    float64 numpy oracle finishes in seconds -- the vocabulary projection is covered at 23000 words above): the bf16
    mode against the oracle evaluated with the SAME operand rounding.  rel 1e-3 on the loss, 1e-4 on attention weights;
    gradients: the measured residue (accumulation order, hardware tanh / exp2, rounding flips amplified over 70
    recurrent steps) is printed per tensor and bounded."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200.model import S2VTAttModel
    B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 3000
    p = W.s2vtatt_params(V, H, E, Vc, 177)
    vid, s, s_len = W.make_batch(B, N, V, L, Vc, 178)
    O.set_operand_rounding(O.bf16_round, O.fp16_round)
    try:
        ref = O.train_iter_s2vtatt({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64), s, s_len,
                                   Vc - 4, L)
    finally:
        O.set_operand_rounding()
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision="bf16"), p).train()
    loss, acc, pred = m.forward_loss(torch.from_numpy(vid).cuda(), torch.from_numpy(s).cuda(),
                                     torch.from_numpy(s_len).cuda())
    loss.backward()
    errs = {k: relerr(v, ref["grads"][k]) for k, v in grads_of(m).items()}
    a_err = float(np.abs(m.last_alphas.cpu().numpy() - ref["alphas"]).max())
    l_err = abs(loss.item() - ref["loss"]) / abs(ref["loss"])
    t_err = float(np.abs(m.last_token_nll.cpu().numpy() - ref["token_nll"]).max())
    print("\n[bf16 vs bf16-operand oracle, B=128] loss rel %.2e  token-loss abs %.2e  alphas abs %.2e  grads rel max %.2e (%s)"
          % (l_err, t_err, a_err, max(errs.values()), max(errs, key=errs.get)))
    for k in sorted(errs, key=errs.get, reverse=True):
        print("   %-45s %.2e" % (k, errs[k]))
    assert l_err < 1e-3
    assert t_err < 1e-3 * float(np.abs(ref["token_nll"]).max())
    assert a_err < 1e-4
    for k, e in errs.items():
        assert e < 2e-3, (k, e)


def test_full_size_greedy_ids_vs_reference():
    """Greedy token ids at B = 128, Vc = 23000 (the configuration bench.py's `greedy` times) against the reference's own
    eval branch (model/S2VTAttModel.py:172-191), float32 and float64 runs of the unmodified module."""
    from pvcr_b200.model import S2VTAttModel
    g = _golden()
    (B, N, V, H, E, L, Vc), p, vid, s, s_len = _inputs(g)
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), p).eval()
    ids, _ = m.greedy(torch.from_numpy(vid).cuda())
    ids = ids.cpu().numpy()
    same64, same32 = ids == g["greedy_ids_f64"], ids == g["greedy_ids_f32"]
    margin = g["greedy_margin_f64"]
    print("\n[greedy, full cfg2] ids == reference(f64) on %d / %d tokens, == reference(f32) on %d; smallest reference "
          "top-2 margin %.3e, %d steps with a margin below 1e-5" % (same64.sum(), ids.size, same32.sum(), float(margin.min()),
                                                                    int((margin < 1e-5).sum())))
    # Bit-exact wherever the reference itself is decided: a video may only leave the reference sequence at a step whose
    # top-2 logit margin in the float64 reference is below fp32 rounding of a K = 512 dot product (the random-init fixture
    # has 4 such steps, down to 3e-8 -- an fp32 reference with another summation order flips there too); what follows a
    # legitimate flip is a different, equally valid continuation and is not compared.
    diverged = 0
    for b in range(B):
        bad = np.nonzero(~same64[b])[0]
        if bad.size:
            diverged += 1
            assert margin[b, bad[0]] < 1e-5, (b, int(bad[0]), float(margin[b, bad[0]]))
    assert diverged <= int((margin < 1e-5).sum())


def _hs_dropout_scale(B, L, H, p, seed):
    from pvcr_b200 import _lib
    ones = torch.ones(B * L * H, dtype=torch.float32, device="cuda")
    out = torch.empty_like(ones)
    _lib.check(_lib.lib().pvcr_out_dropout_apply(_lib.ptr(ones), _lib.ptr(out), ones.numel(), p, seed, _lib.stream_ptr()),
               "pvcr_out_dropout_apply")
    return out.view(B, L, H)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_dropout_on_matches_oracle_with_the_same_mask(precision, monkeypatch):
    """The TIMED configuration has dropout 0.2 (reference default, args.py:26).  The kernels draw the mask of
    `pred_linear`'s Dropout from Philox(seed, element); the same mask (exported through pvcr_out_dropout_apply) is
    handed to the oracle, so keep-rate, the 1/(1-p) scaling and the mask agreement between the forward and the
    recomputing fused-CE backward are all checked numerically: loss, per-token loss, every gradient."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200 import functional as F_
    from pvcr_b200.model import S2VTAttModel
    B, N, V, H, E, L, Vc = 16, 40, 512, 512, 300, 30, 3000
    pdrop, seed = 0.2, 0x1234567
    p = W.s2vtatt_params(V, H, E, Vc, 41)
    vid, s, s_len = W.make_batch(B, N, V, L, Vc, 42)
    monkeypatch.setattr(F_, "next_seed", lambda: seed)
    scale = _hs_dropout_scale(B, L, H, pdrop, seed)
    keep = float((scale > 0).float().mean().item())
    assert abs(keep - (1 - pdrop)) < 0.01, keep
    assert torch.all((scale == 0) | ((scale - 1 / (1 - pdrop)).abs() < 1e-6))
    ref = O.train_iter_s2vtatt({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64), s, s_len,
                               Vc - 4, L, hs_scale=scale.double().cpu().numpy())
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), pdrop, H, V, L, precision=precision), p).train()
    tv, ts, tl = torch.from_numpy(vid).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(s_len).cuda()
    t_loss, t_grad = (2e-6, 2e-4) if precision == "bf16x3" else (1e-3, 2e-2)
    for path in ("fused", "tape_free", "logits"):
        m.zero_grad(set_to_none=True)
        if path == "fused":
            loss, _, _ = m.forward_loss(tv, ts, tl)
            loss.backward()
        elif path == "tape_free":
            loss, _, _ = m.train_step_grads(tv, ts, tl)
        else:       # the reference's own loop: logits = model(vid, s); calc_masked_loss(...); loss.backward()
            from pvcr_b200 import train_utils as TU
            loss = TU.calc_masked_loss(m(tv, ts), ts, tl, torch.nn.CrossEntropyLoss(reduction="none"))
            loss.backward()
        torch.cuda.synchronize()
        assert abs(loss.item() - ref["loss"]) <= t_loss * abs(ref["loss"]), (path, loss.item(), ref["loss"])
        if path != "logits":
            assert float(np.abs(m.last_token_nll.cpu().numpy() - ref["token_nll"]).max()) < (2e-5 if precision == "bf16x3" else 5e-2)
        errs = {k: relerr(v, ref["grads"][k]) for k, v in grads_of(m).items()}
        assert max(errs.values()) < t_grad, (path, max(errs, key=errs.get), max(errs.values()))
    # and a different seed gives a different loss (fresh masks per step)
    monkeypatch.setattr(F_, "next_seed", lambda: seed + 1)
    l2, _, _ = m.forward_loss(tv, ts, tl)
    assert abs(l2.item() - ref["loss"]) > 1e-6


def test_dropout_on_matches_reference_modules_with_the_same_mask(monkeypatch):
    """Same check against the reference MODULES (oracle/_ref bytecode, when it travelled to this box): the reference's
    nn.Dropout inside `pred_linear` is swapped for a module that applies the kernels' mask step by step."""
    from oracle import reference_runner as R
    if not R.available():
        pytest.skip("oracle/_ref not present on this box")
    from oracle import workloads as W
    from pvcr_b200 import functional as F_
    from pvcr_b200.model import S2VTAttModel
    B, N, V, H, E, L, Vc = 8, 12, 256, 128, 64, 10, 500
    pdrop, seed = 0.2, 0x7654321
    p = W.s2vtatt_params(V, H, E, Vc, 51)
    vid, s, s_len = W.make_batch(B, N, V, L, Vc, 52)
    monkeypatch.setattr(F_, "next_seed", lambda: seed)
    scale = _hs_dropout_scale(B, L, H, pdrop, seed).double().cpu()

    class GivenMask(torch.nn.Module):          # stands in for nn.Dropout(p) of decoder.pred_linear (S2VTAttModel.py:121)
        def __init__(self):
            super().__init__()
            self.step = 0

        def forward(self, x):
            y = x * scale[:, self.step]
            self.step += 1
            return y

    R.set_device("cpu")
    ref = R.build_s2vtatt((B, N, V, H, E, L, Vc), p, dropout_p=pdrop, dtype=torch.float64).train()
    ref.decoder.pred_linear[0] = GivenMask()
    loss_r, _, _, _ = R.run_iter(ref, torch.from_numpy(vid).double(), torch.from_numpy(s), torch.from_numpy(s_len))
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), pdrop, H, V, L, precision="bf16x3"), p).train()
    loss, _, _ = m.forward_loss(torch.from_numpy(vid).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(s_len).cuda())
    loss.backward()
    assert abs(loss.item() - loss_r.item()) <= 2e-6 * abs(loss_r.item())
    got = grads_of(m)
    for k, prm in ref.named_parameters():
        assert relerr(got[k], prm.grad.numpy()) < 2e-4, k


def test_philox_uniform_stays_strictly_inside_the_unit_interval():
    """the Gumbel draw is -log(-log(u)); u == 1.0 made the Exp(1) draw 0 and the probabilities NaN
    about once in 2^24 draws.  2^28 consecutive indices x 2 seeds: min > 0, max < 1, and the full 23-bit range is hit."""
    from pvcr_b200 import _lib
    mm = torch.empty(2, dtype=torch.float32, device="cuda")
    for seed in (1, 0x9E3779B97F4A7C15):
        _lib.check(_lib.lib().pvcr_debug_philox_minmax(seed, 0, 1 << 28, _lib.ptr(mm), _lib.stream_ptr()), "philox_minmax")
        lo, hi = (float(x) for x in mm.cpu())
        assert 0.0 < lo <= 2.0 ** -23 and 1.0 - 2.0 ** -23 <= hi < 1.0, (lo, hi)
        assert np.isfinite(-np.log(-np.log(np.float32(hi)))) and np.isfinite(-np.log(-np.log(np.float32(lo))))


def test_generator_in_kernel_noise_is_finite():
    """The production RationaleNet path (no injected noise): probabilities finite and in [0,1] over many draws."""
    from oracle import workloads as W
    from pvcr_b200.model import RationaleNet
    B, N, V, H, E, L, Vc = 64, 40, 128, 64, 32, 6, 100
    m = RationaleNet(FixtureGlove(Vc, E), 0.0, H, V, L, 1.0, "s2vt-att").cuda().train()
    vid = torch.randn(B, N, V, device="cuda")
    for _ in range(50):
        probs, p1, pen = m.gen.select(vid)
        assert torch.isfinite(probs).all() and torch.isfinite(pen).all()
        assert float(probs.min()) >= 0.0 and float(probs.max()) <= 1.0
