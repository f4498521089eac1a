"""Boundary of SURVEY.md section 8(b) that the reference's SpatialNet consumes: `encode_step(vid_feat, rnn_state)` and
`decode(encoder_outs, encoder_final, s)` on both caption nets (model/S2VTAttModel.py:219-243, model/S2VTModel.py:57-72,
88-177; SpatialNet.py:127,140).  Goldens: tests/golden/spatial_*_tiny.npz, produced by oracle/gen_golden_spatial.py from
the UNMODIFIED reference SpatialNet in float64 (logits, seq_alphas, loss, every parameter gradient, eval ids)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.golden_util import GOLDEN, relerr
from tests.gpu_util import FixtureGlove

pytestmark = pytest.mark.gpu

CASES = [("spatial_att_tiny", "s2vt-att"), ("spatial_s2vt_tiny", "s2vt")]


def _load(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    d = {k: z[k] for k in z.files}
    params = {k[6:]: d[k] for k in d if k.startswith("param.")}
    grads = {k[5:]: d[k] for k in d if k.startswith("grad.")}
    return d, params, grads


def _caption_net(arch, dims, params):
    from pvcr_b200.model import S2VTAttModel, S2VTModel
    B, N, Fdim, H, E, L, Vc = dims
    cls = S2VTAttModel if arch == "s2vt-att" else S2VTModel
    m = cls(FixtureGlove(Vc, E), 0.0, H, Fdim, L, precision="bf16x3")
    m.load_state_dict({k[len("caption_net."):]: torch.from_numpy(np.asarray(v, np.float32)) for k, v in params.items()
                       if k.startswith("caption_net.")})
    return m.cuda()


class SpatialFront(nn.Module):
    """The part of SpatialNet in FRONT of the caption net (conv stack + spatial attention, model/SpatialNet.py:76-86,
    27-53) in plain torch, and SpatialNet.forward's frame loop (:99-142) restated around `encode_step` / `decode`."""

    def __init__(self, Fdim, H, params):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(Fdim, H, 3, 1, 1), nn.BatchNorm2d(H), nn.ReLU(), nn.Conv2d(H, H, 3, 1, 1),
                                  nn.BatchNorm2d(H), nn.ReLU())
        self.key_layer = nn.Linear(H, H, bias=False)
        self.query_layer = nn.Linear(H, H, bias=False)
        self.energy_layer = nn.Linear(H, 1, bias=False)
        sd = {k[len("conv."):]: torch.from_numpy(np.asarray(v)) for k, v in params.items() if k.startswith("conv.")}
        self.conv.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
        for n in ("key_layer", "query_layer", "energy_layer"):
            getattr(self, n).weight.data.copy_(torch.from_numpy(np.asarray(params["attention.%s.weight" % n], np.float32)))
        self.H = H

    def forward(self, caption_net, vid, s):
        B, N, Fd, K, _ = vid.shape
        conv = self.conv(vid.view(-1, Fd, K, K)).view(B, N, -1, K * K).transpose(2, 3)       # B x N x K^2 x H
        feats = vid.view(B, N, Fd, -1).transpose(2, 3)                                        # B x N x K^2 x F
        state = torch.zeros(1, B, self.H, device=vid.device)
        outs, alphas = [], []
        for i in range(N):
            pk = self.key_layer(conv[:, i].contiguous().view(-1, self.H)).view(B, -1, self.H)
            q = self.query_layer(state.squeeze(0))
            sc = self.energy_layer(torch.tanh(q.unsqueeze(1) + pk).view(-1, self.H)).view(B, -1)
            a = F.softmax(sc, dim=1)
            ctx = torch.bmm(a.unsqueeze(1), feats[:, i]).squeeze(1)
            out, state = caption_net.encode_step(ctx, state)
            outs.append(out)
            alphas.append(a.view(-1, K, K).unsqueeze(1))
        logits = caption_net.decode(torch.cat(outs, dim=0), state, s)
        return logits, torch.cat(alphas, dim=1)


@pytest.mark.parametrize("tag,arch", CASES)
def test_spatialnet_loop_through_encode_step_and_decode(tag, arch):
    from pvcr_b200 import train_utils as TU
    d, params, grads = _load(tag)
    dims = tuple(int(x) for x in d["dims"])
    B, N, Fdim, H, E, L, Vc = dims
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cap = _caption_net(arch, dims, params).train()
    front = SpatialFront(Fdim, H, params).cuda().train()
    vid = torch.from_numpy(d["vid"]).cuda()
    s, s_len = torch.from_numpy(d["s"]).cuda(), torch.from_numpy(d["s_len"]).cuda()
    logits, seq_alphas = front(cap, vid, s)
    assert logits.shape == d["logits"].shape and seq_alphas.shape == d["seq_alphas"].shape
    assert relerr(logits.detach().cpu().numpy(), d["logits"]) < 5e-5
    assert np.abs(seq_alphas.detach().cpu().numpy() - d["seq_alphas"]).max() < 1e-5
    loss = TU.calc_masked_loss(logits, s, s_len, nn.CrossEntropyLoss(reduction="none"))
    assert abs(loss.item() - float(d["loss"])) < 5e-6 * abs(float(d["loss"]))
    loss.backward()
    checked = 0
    for k, g in grads.items():
        if k.startswith("caption_net."):
            prm = dict(cap.named_parameters())[k[len("caption_net."):]]
        elif k.startswith("conv."):
            prm = dict(front.conv.named_parameters())[k[len("conv."):]]
        else:
            prm = getattr(front, k.split(".")[1]).weight
        if prm.grad is None:       # parameters the reference never reaches on this path have zero gradient
            assert np.abs(g).max() == 0.0, k
            continue
        e = relerr(prm.grad.double().cpu().numpy(), g)
        assert e < 5e-4 or np.linalg.norm(g) < 1e-9, (k, e)       # conv / BN gradients pass through torch's fp32 kernels
        checked += 1
    assert checked >= 10
    # eval branch: greedy ids from the given encoder outputs
    cap.eval(); front.eval()
    with torch.no_grad():
        lg, al = front(cap, vid, None)
    assert np.array_equal(torch.argmax(lg, dim=2).cpu().numpy(), d["eval_ids"])
    assert relerr(lg.cpu().numpy(), d["eval_logits"]) < 5e-5


@pytest.mark.parametrize("tag,arch", CASES)
def test_reference_spatialnet_class_runs_with_the_dropin(tag, arch):
    """The reference's OWN SpatialNet class (oracle/_ref bytecode) with `caption_net` replaced by the drop-in module."""
    from oracle import reference_runner as R
    if not R.available():
        pytest.skip("oracle/_ref not present on this box")
    from pvcr_b200 import train_utils as TU
    d, params, grads = _load(tag)
    dims = tuple(int(x) for x in d["dims"])
    B, N, Fdim, H, E, L, Vc = dims
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    SpatialNet = R.modules()["model.SpatialNet"].SpatialNet
    net = SpatialNet(R.FakeGlove(Vc, E), 0.0, H, Fdim, L, arch)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)).float() if np.asarray(v).dtype.kind == "f"
                         else torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    net.caption_net = _caption_net(arch, dims, params)
    net = net.cuda().train()
    vid = torch.from_numpy(d["vid"]).cuda()
    s, s_len = torch.from_numpy(d["s"]).cuda(), torch.from_numpy(d["s_len"]).cuda()
    logits, seq_alphas = net(vid, s)
    assert relerr(logits.detach().cpu().numpy(), d["logits"]) < 5e-5
    assert np.abs(seq_alphas.detach().cpu().numpy() - d["seq_alphas"]).max() < 1e-5
    loss = TU.calc_masked_loss(logits, s, s_len, nn.CrossEntropyLoss(reduction="none"))
    loss.backward()
    for k, prm in net.named_parameters():
        if prm.grad is None:
            assert np.abs(grads[k]).max() == 0.0, k
            continue
        e = relerr(prm.grad.double().cpu().numpy(), grads[k])
        assert e < 5e-4 or np.linalg.norm(grads[k]) < 1e-9, (k, e)


@pytest.mark.parametrize("stepwise", [False, True])
@pytest.mark.parametrize("tag,arch", CASES)
def test_dropin_spatialnet_module(tag, arch, stepwise, monkeypatch):
    """pvcr_b200.model.SpatialNet: reference constructor / forward contract / state_dict keys; loads the golden's
    reference state_dict key for key and reproduces logits, seq_alphas, loss and every gradient -- with the frame loop as one
    library call per direction (pvcr_spatial_encode_fwd/bwd, the default) and as the step-wise Function chain."""
    from pvcr_b200 import train_utils as TU
    if stepwise:
        monkeypatch.setenv("PVCR_SPATIAL_STEPWISE", "1")
    else:
        monkeypatch.delenv("PVCR_SPATIAL_STEPWISE", raising=False)
    from pvcr_b200.model import SpatialNet
    d, params, grads = _load(tag)
    B, N, Fdim, H, E, L, Vc = (int(x) for x in d["dims"])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = SpatialNet(FixtureGlove(Vc, E), 0.0, H, Fdim, L, arch, precision="bf16x3")
    assert set(net.state_dict()) == set(params), set(net.state_dict()) ^ set(params)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)).float() if np.asarray(v).dtype.kind == "f"
                         else torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    net = net.cuda().train()
    vid = torch.from_numpy(d["vid"]).cuda()
    s, s_len = torch.from_numpy(d["s"]).cuda(), torch.from_numpy(d["s_len"]).cuda()
    logits, seq_alphas = net(vid, s)
    assert relerr(logits.detach().cpu().numpy(), d["logits"]) < 5e-5
    assert np.abs(seq_alphas.detach().cpu().numpy() - d["seq_alphas"]).max() < 1e-5
    loss = TU.calc_masked_loss(logits, s, s_len, nn.CrossEntropyLoss(reduction="none"))
    assert abs(loss.item() - float(d["loss"])) < 5e-6 * abs(float(d["loss"]))
    loss.backward()
    for k, prm in net.named_parameters():
        if prm.grad is None:
            assert np.abs(grads[k]).max() == 0.0, k
            continue
        e = relerr(prm.grad.double().cpu().numpy(), grads[k])
        assert e < 5e-4 or np.linalg.norm(grads[k]) < 1e-9, (k, e)
    net.eval()
    with torch.no_grad():
        lg, _ = net(vid, None)
    assert np.array_equal(torch.argmax(lg, dim=2).cpu().numpy(), d["eval_ids"])


@pytest.mark.parametrize("precision,tol", [("bf16", 3e-2), ("bf16x2", 2e-3)])
def test_spatialnet_frame_sweep_vs_stepwise(precision, tol, monkeypatch):
    """The benchmarked arithmetic modes: the fused frame loop (csrc/spatial_sweep.cu) against the step-wise chain on the same
    weights and inputs at a shape the vectorised attention kernels serve (H = 256, F = 512, 4 x 4 cells), and the same step
    captured as one CUDA graph (GraphedAutogradStep) against the eager one."""
    from pvcr_b200 import train_utils as TU
    from pvcr_b200.graphs import GraphedAutogradStep
    from pvcr_b200.model import SpatialNet
    B, N, Fd, K, H, E, L, Vc = 6, 5, 512, 4, 256, 32, 7, 300
    torch.manual_seed(11)
    net = SpatialNet(FixtureGlove(Vc, E), 0.0, H, Fd, L, "s2vt-att", precision=precision).cuda().train()
    vid = torch.randn(B, N, Fd, K, K, device="cuda")
    s = torch.randint(0, Vc - 4, (B, L), device="cuda")
    s_len = torch.randint(1, L + 1, (B,), device="cuda")
    crit = nn.CrossEntropyLoss(reduction="none")

    def run():
        net.zero_grad(set_to_none=True)
        logits, al = net(vid, s)
        loss = TU.calc_masked_loss(logits, s, s_len, crit)
        loss.backward()
        return loss.item(), al.detach().clone(), {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}

    monkeypatch.setenv("PVCR_SPATIAL_STEPWISE", "1")
    l0, a0, g0 = run()
    monkeypatch.delenv("PVCR_SPATIAL_STEPWISE")
    l1, a1, g1 = run()
    assert abs(l0 - l1) < tol * abs(l0)
    assert (a0 - a1).abs().max().item() < tol
    assert set(g0) == set(g1)
    for k in g0:
        if k in ("conv.0.bias", "conv.3.bias"):      # a bias in front of a batch-statistics BatchNorm: zero gradient up to rounding
            assert g0[k].norm().item() < 1e-6 and g1[k].norm().item() < 1e-6, k
            continue
        e = relerr(g1[k].double().cpu().numpy(), g0[k].double().cpu().numpy())
        assert e < tol or g0[k].norm().item() < 1e-9, (k, e)
    # the same step as one CUDA graph: same kernels, same order -> the eager result to rounding of the atomics
    gs = GraphedAutogradStep(net, lambda: TU.calc_masked_loss(net(vid, s)[0], s, s_len, crit))
    lg = gs.replay().item()
    assert abs(lg - l1) < 1e-5 * abs(l1)
    for k, prm in net.named_parameters():
        if k in g1 and k not in ("conv.0.bias", "conv.3.bias"):
            e = relerr(prm.grad.double().cpu().numpy(), g1[k].double().cpu().numpy())
            assert e < 1e-4 or g1[k].norm().item() < 1e-9, (k, e)


def test_run_iter_serves_spatialnet_in_training():
    """trainer.run_iter with a SpatialNet: the reference's train_spatial.py:30-39 contract ((acc, loss[, pred]), autograd loss)."""
    from pvcr_b200 import train_utils as TU
    from pvcr_b200.model import SpatialNet
    from pvcr_b200.trainer import run_iter
    B, N, Fd, K, H, E, L, Vc = 4, 3, 64, 3, 32, 16, 5, 50
    torch.manual_seed(2)
    net = SpatialNet(FixtureGlove(Vc, E), 0.0, H, Fd, L, "s2vt-att", precision="bf16x3").cuda().train()
    data = {"vid_feats": torch.randn(B, N, Fd, K, K), "sent": torch.randint(0, Vc - 4, (B, L)), "sent_len": torch.randint(1, L + 1, (B,))}
    acc, loss, pred = run_iter(None, data, net, nn.CrossEntropyLoss(reduction="none"), return_pred=True)
    assert pred.shape == (B, L) and 0.0 <= float(acc) <= 1.0
    loss.backward()
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert all(torch.isfinite(g).all() for g in got.values())
    assert {"conv.0.weight", "conv.3.weight", "attention.key_layer.weight", "attention.query_layer.weight",
            "attention.energy_layer.weight"} <= set(got)
    logits, _ = net(data["vid_feats"].cuda(), data["sent"].cuda())
    want = TU.calc_masked_loss(logits, data["sent"].cuda(), data["sent_len"].cuda(), None)
    assert abs(loss.item() - want.item()) < 1e-6 * abs(want.item())


def test_boundary_surface_matches_reference_signatures():
    """reset_parameter(s), encode_step, decode, encode and calc_sentence_mask(batch_size, max_len, s_len) exist with the
    reference's argument lists (model/S2VTAttModel.py:215-243, model/S2VTModel.py:52-88, train_utils.py:22)."""
    import inspect
    from pvcr_b200 import train_utils as TU
    from pvcr_b200.model import S2VTAttModel, S2VTModel
    assert list(inspect.signature(TU.calc_sentence_mask).parameters) == ["batch_size", "max_len", "s_len"]
    for cls, names in ((S2VTAttModel, ("reset_parameter", "encode_step", "decode")),
                       (S2VTModel, ("reset_parameters", "encode_step", "encode", "decode"))):
        for n in names:
            assert callable(getattr(cls, n)), (cls.__name__, n)
    assert list(inspect.signature(S2VTAttModel.encode_step).parameters) == ["self", "vid_feat", "rnn_state"]
    assert list(inspect.signature(S2VTAttModel.decode).parameters) == ["self", "encoder_outs", "encoder_final", "s"]
    assert list(inspect.signature(S2VTModel.decode).parameters) == ["self", "output1", "state1", "s"]
    mask = TU.calc_sentence_mask(2, 4, torch.tensor([1, 3], device="cuda"))
    assert mask.tolist() == [[1, 0, 0, 0], [1, 1, 1, 0]]
