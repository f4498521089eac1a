"""SpatialNet front on the B200 kernels (csrc/conv.cu): Conv3x3 + BatchNorm2d + ReLU twice as nine row-shifted tcgen05 GEMMs per
convolution, and the per-frame attention over the K*K cells -- against torch's own Conv2d / BatchNorm2d / softmax in float64
(the ops the reference's SpatialNet is made of, model/SpatialNet.py:27-53,76-86)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.golden_util import relerr

pytestmark = pytest.mark.gpu


def _stack(F, H):
    return nn.Sequential(nn.Conv2d(F, H, 3, 1, 1), nn.BatchNorm2d(H), nn.ReLU(), nn.Conv2d(H, H, 3, 1, 1), nn.BatchNorm2d(H), nn.ReLU())


@pytest.mark.parametrize("precision,tol", [("bf16x3", 3e-4), ("bf16", 8e-2)])
# (4, 256, 256, 6): channel counts served by the fused nine-tap launch of the single-plane mode (csrc/conv.cu: fused_taps)
# (2, 1024, 512, 4): enough weight-gradient tiles per tap for the batched nine-tap d W launch
@pytest.mark.parametrize("I,F,H,K", [(6, 24, 32, 3), (10, 128, 64, 6), (7, 72, 48, 5), (4, 256, 256, 6), (2, 1024, 512, 4)])
def test_conv_bn_relu_front_matches_torch(I, F, H, K, precision, tol):
    from pvcr_b200 import functional as F_
    torch.manual_seed(I * 100 + K)
    ref = _stack(F, H).double().cuda().train()
    with torch.no_grad():
        for m in ref:
            if isinstance(m, nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
    x = torch.randn(I, F, K, K, device="cuda")
    c1, b1, c2, b2 = ref[0], ref[1], ref[3], ref[4]
    prm = [t.detach().float().clone().requires_grad_(True) for t in (c1.weight, c1.bias, b1.weight, b1.bias, c2.weight, c2.bias, b2.weight, b2.bias)]
    run = [t.detach().float().clone() for t in (b1.running_mean, b1.running_var, b2.running_mean, b2.running_var)]
    nsplit = {"bf16": 1, "bf16x3": 3}[precision]
    cfg = {"nsplit": nsplit, "training": True, "eps": b1.eps, "momentum": b1.momentum}
    conv_feats, feats_cl = F_.SpatialFront.apply(cfg, x, *run, *prm)
    y_ref = ref(x.double())                                              # [I, H, K, K]
    want = y_ref.permute(0, 2, 3, 1).reshape(I * K * K, H)
    assert relerr(conv_feats.detach().cpu().numpy(), want.detach().cpu().numpy()) < tol
    assert torch.equal(feats_cl.view(I, K * K, F), x.permute(0, 2, 3, 1).reshape(I, K * K, F))
    # running estimates updated as torch does (momentum 0.1, unbiased variance)
    for got, m, name in ((run[0], b1.running_mean, "rm1"), (run[1], b1.running_var, "rv1"), (run[2], b2.running_mean, "rm2"),
                         (run[3], b2.running_var, "rv2")):
        assert relerr(got.cpu().numpy(), m.detach().cpu().numpy()) < tol, name
    g = torch.randn(I * K * K, H, device="cuda")
    (conv_feats * g).sum().backward()
    (want * g.double()).sum().backward()
    refs = (c1.weight, c1.bias, b1.weight, b1.bias, c2.weight, c2.bias, b2.weight, b2.bias)
    names = ("conv1.w", "conv1.b", "bn1.w", "bn1.b", "conv2.w", "conv2.b", "bn2.w", "bn2.b")
    for n, p, r in zip(names, prm, refs):
        ref_g = r.grad.cpu().numpy()
        if n in ("conv1.b", "conv2.b"):      # a bias in front of a batch-statistics BatchNorm has zero gradient up to rounding
            assert np.abs(p.grad.cpu().numpy()).max() < 1e-3 * max(1.0, float(np.abs(g.cpu().numpy()).max())), n
            continue
        # single-plane bf16: ReLU-mask flips and the cancellation in the BatchNorm gradients leave ~0.1 at these tiny batches
        gtol = tol if precision != "bf16" else 0.15
        assert relerr(p.grad.cpu().numpy(), ref_g) < gtol, (n, relerr(p.grad.cpu().numpy(), ref_g))
    # eval mode: running estimates
    ref.eval()
    cfg["training"] = False
    with torch.no_grad():
        cf_e, _ = F_.SpatialFront.apply(cfg, x, *run, *[p.detach() for p in prm])
        want_e = ref(x.double()).permute(0, 2, 3, 1).reshape(I * K * K, H)
    assert relerr(cf_e.cpu().numpy(), want_e.cpu().numpy()) < tol


@pytest.mark.parametrize("B,Kc,H,Fv", [(5, 9, 32, 24), (16, 36, 512, 2048)])
def test_spatial_attention_step_matches_torch(B, Kc, H, Fv):
    from pvcr_b200 import functional as F_
    torch.manual_seed(B)
    q = torch.randn(B, H, device="cuda", requires_grad=True)
    pk = torch.randn(B, Kc, H, device="cuda", requires_grad=True)
    feats = torch.randn(B, Kc, Fv, device="cuda")
    v = (torch.randn(1, H, device="cuda") / H ** 0.5).requires_grad_(True)
    ctx, alpha = F_.SpatialAttnStep.apply(q, pk, feats, v)
    qd, pkd, vd = (t.detach().double().requires_grad_(True) for t in (q, pk, v))
    sc = torch.tanh(qd.unsqueeze(1) + pkd) @ vd.reshape(-1)
    al = torch.softmax(sc, dim=1)
    want = torch.bmm(al.unsqueeze(1), feats.double()).squeeze(1)
    assert np.abs(alpha.detach().cpu().numpy() - al.detach().cpu().numpy()).max() < 1e-6
    assert relerr(ctx.detach().cpu().numpy(), want.detach().cpu().numpy()) < 1e-5
    g = torch.randn(B, Fv, device="cuda")
    (ctx * g).sum().backward()
    (want * g.double()).sum().backward()
    for n, a, b in (("dq", q.grad, qd.grad), ("dpk", pk.grad, pkd.grad), ("dv", v.grad, vd.grad)):
        assert relerr(a.cpu().numpy(), b.cpu().numpy()) < 2e-5, n


def test_linear_function_matches_torch():
    from pvcr_b200 import functional as F_
    torch.manual_seed(3)
    x = torch.randn(50, 96, device="cuda", requires_grad=True)
    w = torch.randn(40, 96, device="cuda", requires_grad=True)
    y = F_.Linear.apply(3, x, w)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yd = xd @ wd.t()
    assert relerr(y.detach().cpu().numpy(), yd.detach().cpu().numpy()) < 1e-5
    g = torch.randn_like(y)
    (y * g).sum().backward(); (yd * g.double()).sum().backward()
    assert relerr(x.grad.cpu().numpy(), xd.grad.cpu().numpy()) < 1e-5 and relerr(w.grad.cpu().numpy(), wd.grad.cpu().numpy()) < 1e-5
