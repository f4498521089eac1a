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
