"""GPU parity: drop-in S2VTModel vs the reference outputs in tests/golden/ and vs the numpy oracle."""
import numpy as np
import pytest
import torch

from tests.golden_util import relerr
from tests.gpu_util import FixtureGlove, grads_of, load_case, to_cuda

pytestmark = pytest.mark.gpu

TOL = {"bf16x3": (2e-6, 2e-5, 1e-4), "bf16x2": (1e-4, 1e-3, 1e-3), "bf16": (1e-3, 2e-2, 3e-2)}


def _model(tag, precision):
    from pvcr_b200.model import S2VTModel
    d, params, grads, (B, N, V, H, E, L, Vc) = load_case(tag)
    m = S2VTModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision=precision)
    return to_cuda(m, params), d, grads


@pytest.mark.parametrize("precision", ["bf16x3", "bf16x2", "bf16"])
@pytest.mark.parametrize("tag", ["s2vt_tiny", "s2vt_mid"])
def test_forward_loss_and_grads(tag, precision):
    m, d, g = _model(tag, precision)
    t_loss, t_logits, t_grad = TOL[precision]
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    s_len = torch.from_numpy(d["s_len"]).cuda()
    m.train()
    loss, acc, pred = m.forward_loss(vid, s, s_len)
    loss.backward()
    assert abs(loss.item() - float(d["loss"])) <= t_loss * abs(float(d["loss"]))
    if precision == "bf16x3":
        assert np.array_equal(pred.cpu().numpy(), d["pred"])
    got = grads_of(m)
    assert set(got) == set(g)
    for k in g:
        assert relerr(got[k], g[k]) < t_grad, (k, relerr(got[k], g[k]))


@pytest.mark.parametrize("tag", ["s2vt_tiny", "s2vt_mid"])
def test_module_forward_logits(tag):
    m, d, g = _model(tag, "bf16x3")
    m.train()
    logits = m(torch.from_numpy(d["vid"]).cuda(), torch.from_numpy(d["s"]).cuda())
    assert relerr(logits.detach().cpu().numpy(), d["logits"]) < 2e-5


@pytest.mark.parametrize("precision,B", [("bf16", 32), ("bf16x3", 8)])
def test_msvd_shape_vs_oracle(precision, B):
    """cfg1 dims (N=80 frames of 4096-d VGG features, H=512, E=300, L=28) with a reduced vocabulary."""
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    from pvcr_b200.model import S2VTModel
    N, V, H, E, L, Vc = 80, 4096, 512, 300, 28, 2000
    p = W.s2vt_params(V, H, E, Vc, 5)
    vid, s, s_len = W.make_batch(B, N, V, L, Vc, 6)
    ref = O.train_iter_s2vt({k: v.astype(np.float64) for k, v in p.items()}, vid.astype(np.float64), s, s_len, Vc - 4, L)
    m = S2VTModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision=precision)
    m = to_cuda(m, p).train()
    loss, acc, pred = m.forward_loss(torch.from_numpy(vid).cuda(), torch.from_numpy(s).cuda(),
                                     torch.from_numpy(s_len).cuda())
    loss.backward()
    t_loss, _, t_grad = TOL[precision]
    errs = {k: relerr(v, ref["grads"][k]) for k, v in grads_of(m).items()}
    l_err = abs(loss.item() - ref["loss"]) / abs(ref["loss"])
    print("\n[%s B=%d] loss rel %.2e  grads rel max %.2e (%s)" % (precision, B, l_err, max(errs.values()),
                                                               max(errs, key=errs.get)))
    assert l_err < t_loss
    for k, e in errs.items():
        assert e < t_grad, (k, e)
