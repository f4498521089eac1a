"""Optimizer step: the oracle restatement is pinned to the reference's own optimizer
calls (torch.nn.utils.clip_grad_norm_ + torch.optim.Adam, train.py:104-105,157-160) on CPU; the fused CUDA step is held
to the oracle and to torch on the GPU."""
import numpy as np
import pytest
import torch

from oracle import captioning_oracle as O

SHAPES = {"a.weight": (37, 19), "b.bias": (5,), "emb": (70001,), "c.weight": (128, 515), "d": (3,)}


def _make(seed, dtype):
    rng = np.random.default_rng(seed)
    p = {k: rng.standard_normal(s).astype(dtype) for k, s in SHAPES.items()}
    gs = [{k: (rng.standard_normal(s) * (10.0 if i == 0 else 0.01)).astype(dtype) for k, s in SHAPES.items()} for i in range(3)]
    return p, gs


def _torch_reference(p, gs, dtype, device, **hp):
    params = [torch.nn.Parameter(torch.from_numpy(v.copy()).to(device)) for v in p.values()]
    opt = torch.optim.Adam(params, lr=hp["lr"], weight_decay=hp["weight_decay"])
    norms = []
    for g in gs:
        for prm, gv in zip(params, g.values()):
            prm.grad = torch.from_numpy(gv.copy()).to(device)
        norms.append(float(torch.nn.utils.clip_grad_norm_(params, hp["max_norm"])))
        opt.step()
    return {k: prm.detach().cpu().numpy() for k, prm in zip(p, params)}, norms


@pytest.mark.parametrize("wd", [0.0, 1e-3])
def test_oracle_step_matches_reference_optimizer_cpu(wd):
    hp = dict(lr=2e-3, weight_decay=wd, max_norm=1.0)
    p, gs = _make(3, np.float64)
    want, norms = _torch_reference(p, gs, np.float64, "cpu", **hp)
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    for i, g in enumerate(gs):
        n = O.clip_adam_step(p, g, m, v, i + 1, **hp)
        assert abs(n - norms[i]) < 1e-9 * norms[i]
    for k in p:
        assert np.abs(p[k] - want[k]).max() < 1e-12, k


@pytest.mark.gpu
@pytest.mark.parametrize("wd,max_norm", [(0.0, 1.0), (1e-3, 1.0), (1e-3, None)])
def test_fused_clip_adam_matches_oracle_and_torch(wd, max_norm):
    import pvcr_b200  # noqa: F401
    from pvcr_b200.optim import FusedClipAdam
    p32, gs32 = _make(5, np.float32)
    params = [torch.nn.Parameter(torch.from_numpy(v.copy()).cuda()) for v in p32.values()]
    opt = FusedClipAdam(params, lr=2e-3, weight_decay=wd, max_norm=max_norm)
    p64 = {k: v.astype(np.float64) for k, v in p32.items()}
    m = {k: np.zeros_like(v) for k, v in p64.items()}
    v = {k: np.zeros_like(x) for k, x in p64.items()}
    for i, g in enumerate(gs32):
        for prm, gv in zip(params, g.values()):
            prm.grad = torch.from_numpy(gv.copy()).cuda()
        norm = opt.step()
        n = O.clip_adam_step(p64, {k: x.astype(np.float64) for k, x in g.items()}, m, v, i + 1, lr=2e-3, weight_decay=wd,
                             max_norm=max_norm)
        assert abs(norm.item() - n) < 2e-6 * n
    assert int(opt.step_count.item()) == 3
    for prm, k in zip(params, p64):
        got = prm.detach().double().cpu().numpy()
        assert np.abs(got - p64[k]).max() < 2e-6 * max(1.0, np.abs(p64[k]).max()), k       # fp32 state vs float64 oracle
    if max_norm is not None:
        want, _ = _torch_reference(p32, gs32, np.float32, "cuda", lr=2e-3, weight_decay=wd, max_norm=max_norm)
        for prm, k in zip(params, want):
            assert np.abs(prm.detach().cpu().numpy() - want[k]).max() < 2e-6, k


@pytest.mark.gpu
def test_fused_clip_adam_in_cuda_graph_advances_step():
    import pvcr_b200  # noqa: F401
    from pvcr_b200.optim import FusedClipAdam
    prm = torch.nn.Parameter(torch.ones(1000, device="cuda"))
    prm.grad = torch.full((1000,), 0.5, device="cuda")
    opt = FusedClipAdam([prm], lr=1e-2)
    opt.step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        opt.step()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    # constant gradient: every Adam step moves the parameter by exactly lr (bias-corrected m / sqrt(v) = 1)
    assert int(opt.step_count.item()) == 3            # one eager step + two replays (the capture itself does not run)
    assert torch.allclose(prm, torch.full_like(prm, 1.0 - 3 * 1e-2), atol=1e-6)


@pytest.mark.gpu
def test_training_iterations_graph_vs_eager_vs_torch_optimizer():
    """Three whole training iterations (train.py:157-160: fwd, bwd, clip_grad_norm_, Adam) of S2VTAtt: one CUDA graph per
    iteration (step + FusedClipAdam captured together) == the eager tape-free step + FusedClipAdam == the eager step +
    torch's own clip_grad_norm_ / torch.optim.Adam, compared on the parameter updates."""
    import pvcr_b200  # noqa: F401
    from pvcr_b200.graphs import GraphedTrainStep
    from pvcr_b200.model import S2VTAttModel
    from pvcr_b200.optim import FusedClipAdam
    from tests.gpu_util import FixtureGlove, load_case, to_cuda
    d, params, _, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_mid")
    vid = torch.from_numpy(d["vid"]).cuda()
    s = torch.from_numpy(d["s"]).cuda()
    s_len = torch.from_numpy(d["s_len"]).cuda()
    hp = dict(lr=2e-3, weight_decay=4e-5)

    def fresh():
        return to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision="bf16x3"), params).train()

    start = {k: v.detach().clone() for k, v in fresh().named_parameters()}
    # (a) eager step + torch optimizer (the reference's own calls)
    ma = fresh()
    oa = torch.optim.Adam(ma.parameters(), **hp)
    for _ in range(3):
        ma.train_step_grads(vid, s, s_len)
        torch.nn.utils.clip_grad_norm_(ma.parameters(), 1.0)
        oa.step()
    # (b) eager step + fused optimizer
    mb = fresh()
    ob = FusedClipAdam(mb.parameters(), max_norm=1.0, **hp)
    for _ in range(3):
        mb.train_step_grads(vid, s, s_len)
        ob.step()
    # (c) one graph per iteration; the constructor runs one eager iteration first (optimizer pointer table)
    mc = fresh()
    oc = FusedClipAdam(mc.parameters(), max_norm=1.0, **hp)
    g = GraphedTrainStep(mc, (vid, s, s_len), warmup=0, optimizer=oc)
    for _ in range(2):
        g(vid, s, s_len)
    torch.cuda.synchronize()
    assert int(oc.step_count.item()) == 3

    def upd(m):
        return {k: (v.detach() - start[k]).double() for k, v in m.named_parameters()}

    ua, ub, uc = upd(ma), upd(mb), upd(mc)
    for k in ua:
        na = ua[k].norm().item()
        assert na > 0, k
        assert (ub[k] - ua[k]).norm().item() < 2e-3 * na, (k, "fused vs torch", (ub[k] - ua[k]).norm().item() / na)
        assert (uc[k] - ub[k]).norm().item() < 2e-3 * na, (k, "graph vs eager", (uc[k] - ub[k]).norm().item() / na)
