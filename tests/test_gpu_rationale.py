"""GPU parity: drop-in RationaleNet (generator + caption net) vs the reference outputs in tests/golden/."""
import numpy as np
import pytest
import torch

from tests.golden_util import CASES_RAT, relerr
from tests.gpu_util import FixtureGlove, grads_of, load_case, to_cuda

pytestmark = pytest.mark.gpu

TOL = {"bf16x3": (2e-6, 1e-5, 2e-4), "bf16": (2e-3, 5e-3, 5e-2)}      # loss rel, probs abs, grads rel


def _model(tag, precision):
    from pvcr_b200.model import RationaleNet
    d, params, grads, (B, N, V, H, E, L, Vc) = load_case(tag)
    arch = "s2vt-att" if "att" in tag else "s2vt"
    m = RationaleNet(FixtureGlove(Vc, E), 0.0, H, V, L, float(d["tau"]), arch, precision=precision)
    m = to_cuda(m, params)
    m.gen.noise = torch.from_numpy(d["noise"]).cuda()
    return m, d, grads


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("tag", CASES_RAT)
def test_joint_loss_and_grads(tag, precision):
    m, d, g = _model(tag, precision)
    t_loss, t_probs, t_grad = TOL[precision]
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    s_len = torch.from_numpy(d["s_len"]).cuda()
    m.train()
    acc, loss, loss_ce, loss_brev, loss_cont, rlen, pred, probs = m.forward_loss(
        vid, s, s_len, lambda_brev=float(d["lambda_brev"]), lambda_cont=float(d["lambda_cont"]))
    loss.backward()
    for got, key in ((loss, "loss"), (loss_ce, "loss_ce"), (loss_brev, "loss_brev"), (loss_cont, "loss_cont"),
                     (rlen, "rationale_len")):
        assert abs(got.item() - float(d[key])) <= t_loss * max(abs(float(d[key])), 1e-3), key
    assert np.abs(probs.detach().cpu().numpy() - d["probs"]).max() < t_probs
    got = grads_of(m)
    assert set(got) == set(g)
    for k in g:
        assert relerr(got[k], g[k]) < t_grad, (k, relerr(got[k], g[k]))


@pytest.mark.parametrize("tag", CASES_RAT)
def test_module_api_with_reference_loss_functions(tag):
    """model(vid, s) -> (logits, probs); the penalties are evaluated by torch on probs (train_rationale.py:34-40)."""
    m, d, g = _model(tag, "bf16x3")
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    s_len = torch.from_numpy(d["s_len"]).cuda()
    m.train()
    logits, probs = m(vid, s)
    assert relerr(logits.detach().cpu().numpy(), d["logits"]) < 2e-5
    B, L, Vc = logits.shape
    nll = torch.nn.functional.cross_entropy(logits.view(B * L, Vc), s.view(-1), reduction="none").view(B, L)
    mask = (torch.arange(L, device="cuda")[None, :] < s_len[:, None]).float()
    loss_ce = ((nll * mask).sum(1) / s_len.float()).mean()
    p1 = probs[:, :, 1]
    loss = loss_ce + float(d["lambda_brev"]) * p1.sum(1).mean() + \
        float(d["lambda_cont"]) * (p1[:, 1:] - p1[:, :-1]).abs().mean()
    loss.backward()
    assert abs(loss.item() - float(d["loss"])) < 2e-6 * abs(float(d["loss"]))
    got = grads_of(m)
    for k in g:
        assert relerr(got[k], g[k]) < 2e-4, (k, relerr(got[k], g[k]))


@pytest.mark.parametrize("tag", CASES_RAT)
def test_eval_hard_selection_and_greedy(tag):
    m, d, _ = _model(tag, "bf16x3")
    m.eval()
    ids, logits, probs = m.greedy(torch.from_numpy(d["vid"]).cuda())
    assert np.abs(probs.cpu().numpy() - d["greedy_probs"]).max() < 1e-6
    assert np.array_equal(ids.cpu().numpy(), d["greedy_ids"])
    assert relerr(logits.cpu().numpy(), d["greedy_logits"]) < 2e-5


@pytest.mark.parametrize("tag", ["rationale_att_mid"])
def test_tape_free_step_matches_autograd(tag):
    m, d, g = _model(tag, "bf16x3")
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    s_len = torch.from_numpy(d["s_len"]).cuda()
    m.train()
    loss, acc, pred = m.train_step_grads(vid, s, s_len, float(d["lambda_brev"]), float(d["lambda_cont"]))
    assert abs(loss.item() - float(d["loss"])) < 2e-6 * abs(float(d["loss"]))
    got = grads_of(m)
    for k in g:
        assert relerr(got[k], g[k]) < 2e-4, (k, relerr(got[k], g[k]))


@pytest.mark.parametrize("precision,B", [("bf16", 24), ("bf16x3", 8)])
def test_msrvtt_shape_vs_oracle(precision, B):
    """cfg3 dims (N=40, V=2048, H=512) with a reduced vocabulary: exercises the persistent LSTM / GRU / decoder kernels
    of the joint RationaleNet + S2VTAtt objective against the float64 oracle."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200.model import RationaleNet
    N, V, H, E, L, Vc, tau = 40, 2048, 512, 300, 30, 3000, 0.8
    pc = W.s2vtatt_params(V, H, E, Vc, 41)
    pg = W.generator_params(V, H, 42)
    p = {"caption_net." + k: v for k, v in pc.items()}
    p.update({"gen." + k: v for k, v in pg.items()})
    vid, s, s_len = W.make_batch(B, N, V, L, Vc, 43)
    noise = np.random.RandomState(44).exponential(size=(B * N, 2)).astype(np.float32)
    ref = O.train_iter_rationale({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64), s, s_len,
                                 Vc - 4, L, tau, noise.astype(np.float64), arch="s2vt-att", lambda_brev=0.05,
                                 lambda_cont=0.5)
    m = RationaleNet(FixtureGlove(Vc, E), 0.0, H, V, L, tau, "s2vt-att", precision=precision)
    m = to_cuda(m, p).train()
    m.gen.noise = torch.from_numpy(noise).cuda()
    acc, loss, loss_ce, loss_brev, loss_cont, rlen, pred, probs = m.forward_loss(
        torch.from_numpy(vid).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(s_len).cuda(), 0.05, 0.5)
    loss.backward()
    t_loss, t_probs, t_grad = TOL[precision]
    errs = {k: relerr(v, ref["grads"][k]) for k, v in grads_of(m).items()}
    l_err = abs(loss.item() - ref["loss"]) / abs(ref["loss"])
    p_err = float(np.abs(probs.detach().cpu().numpy() - ref["probs"]).max())
    print("\n[%s B=%d] loss rel %.2e  probs abs %.2e  grads rel max %.2e (%s)" % (
        precision, B, l_err, p_err, max(errs.values()), max(errs, key=errs.get)))
    assert l_err < t_loss
    assert p_err < t_probs
    for k, e in errs.items():
        assert e < t_grad, (k, e)
