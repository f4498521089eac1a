"""GPU parity: drop-in S2VTAttModel (C ABI -> sm_100a kernels) vs the reference outputs in tests/golden/
(generated from the unmodified reference modules) and vs the numpy oracle at a larger seeded shape."""
import numpy as np
import pytest
import torch

from tests.golden_util import CASES_ATT, relerr
from tests.gpu_util import FixtureGlove, grads_of, load_case, to_cuda

pytestmark = pytest.mark.gpu

# tolerances per arithmetic mode: (loss rel, logits rel, grads rel, alphas abs)
TOL = {"bf16x3": (2e-6, 2e-5, 1e-4, 1e-5), "bf16x2": (1e-4, 1e-3, 1e-3, 1e-4), "bf16": (1e-3, 2e-2, 3e-2, 5e-3)}


def _model(tag, precision):
    from pvcr_b200.model import S2VTAttModel
    d, params, grads, (B, N, V, H, E, L, Vc) = load_case(tag)
    m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision=precision)
    return to_cuda(m, params), d, grads


@pytest.mark.parametrize("precision", ["bf16x3", "bf16x2", "bf16"])
@pytest.mark.parametrize("tag", CASES_ATT)
def test_forward_loss_and_grads(tag, precision):
    m, d, g = _model(tag, precision)
    t_loss, t_logits, t_grad, t_alpha = TOL[precision]
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    s_len = torch.from_numpy(d["s_len"]).cuda()
    m.train()
    loss, acc, pred = m.forward_loss(vid, s, s_len)
    loss.backward()
    assert abs(loss.item() - float(d["loss"])) <= t_loss * abs(float(d["loss"]))
    assert np.abs(m.last_alphas.cpu().numpy() - d["alphas"]).max() < t_alpha
    if precision == "bf16x3":
        assert np.array_equal(pred.cpu().numpy(), d["pred"])
        assert abs(acc.item() - float(d["acc"])) < 1e-6
    got = grads_of(m)
    assert set(got) == set(g)
    for k in g:
        assert relerr(got[k], g[k]) < t_grad, (k, relerr(got[k], g[k]))


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("tag", CASES_ATT)
def test_module_forward_returns_reference_logits(tag, precision):
    """model(vid_feats, s) keeps the reference API: a [B,L,Vc] logits tensor that backpropagates."""
    m, d, g = _model(tag, precision)
    t_loss, t_logits, t_grad, _ = TOL[precision]
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    m.train()
    logits = m(vid, s)
    assert logits.shape == d["logits"].shape
    assert relerr(logits.detach().cpu().numpy(), d["logits"]) < t_logits
    # the reference loss contract evaluated by torch on our logits (train_utils.py:37-54)
    s_len = torch.from_numpy(d["s_len"]).cuda()
    B, L, Vc = logits.shape
    nll = torch.nn.functional.cross_entropy(logits.view(B * L, Vc), s.view(-1), reduction="none").view(B, L)
    mask = (torch.arange(L, device="cuda")[None, :] < s_len[:, None]).float()
    loss = ((nll * mask).sum(1) / s_len.float()).mean()
    loss.backward()
    got = grads_of(m)
    for k in g:
        assert relerr(got[k], g[k]) < t_grad, (k, relerr(got[k], g[k]))
