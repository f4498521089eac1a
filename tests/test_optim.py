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
