"""GPU: the tcgen05/TMA GEMM behind pvcr_linear_fwd / pvcr_linear_bwd against torch float64 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


SHAPES = [(128, 128, 64), (128, 256, 128), (1, 8, 3), (200, 300, 100), (130, 70, 520), (512, 1536, 2048),
          (384, 2304, 512), (256, 512, 2300)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("nsplit", [1, 2, 3])
def test_linear_fwd(M, N, K, nsplit):
    from pvcr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.1
    b = torch.randn(N, device="cuda", generator=g)
    y = ops.linear_fwd(x, w, b, nsplit=nsplit)
    torch.cuda.synchronize()
    if nsplit == 1:
        ref = _bf(x).double() @ _bf(w).double().T + b.double()
        tol = 5e-6          # fp32 tensor-core accumulation over K up to 2300
    else:
        ref = x.double() @ w.double().T + b.double()
        tol = 3e-5 if nsplit == 2 else 5e-6
    assert _rel(y, ref) < tol, (_rel(y, ref), tol)
    # elementwise too: no tile may be garbage
    assert float((y.double() - ref).abs().max()) < 1e3 * tol * float(ref.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 300, 100), (640, 1536, 512), (384, 2300, 512)])
@pytest.mark.parametrize("nsplit", [1, 3])
def test_linear_bwd(M, N, K, nsplit):
    from pvcr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.1
    dy = torch.randn(M, N, device="cuda", generator=g)
    dx, dw, db = ops.linear_bwd(dy, x, w, nsplit=nsplit)
    torch.cuda.synchronize()
    if nsplit == 1:
        rdx = _bf(dy).double() @ _bf(w).double()
        rdw = _bf(dy).double().T @ _bf(x).double()
        tol = 5e-6
    else:
        rdx = dy.double() @ w.double()
        rdw = dy.double().T @ x.double()
        tol = 5e-6
    assert _rel(dx, rdx) < tol, ("dx", _rel(dx, rdx))
    assert _rel(dw, rdw) < tol, ("dw", _rel(dw, rdw))
    assert _rel(db, dy.double().sum(0)) < 1e-6


def test_linear_strided_views():
    """Row-strided inputs/outputs (weights that are column slices of a wider matrix, e.g. W_ih = [Wc | We])."""
    from pvcr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    wide = torch.randn(96, 812, device="cuda", generator=g) * 0.1
    x = torch.randn(64, 300, device="cuda", generator=g)
    y = ops.linear_fwd(x, wide[:, 512:], None, nsplit=3)
    ref = x.double() @ wide[:, 512:].double().T
    assert _rel(y, ref) < 5e-6


@pytest.mark.parametrize("R,N,K", [(64, 64, 64), (128, 256, 128), (200, 300, 100), (3840, 1536, 512), (5120, 1536, 2048),
                                   (1000, 70, 520)])
def test_wgrad_mn_major_operands(R, N, K):
    """dW = dY^T X with both tcgen05 operands MN-major (no transposed staging copies)."""
    from pvcr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(R + N + K)
    dy = torch.randn(R, N, device="cuda", generator=g)
    x = torch.randn(R, K, device="cuda", generator=g)
    dw = ops.wgrad_mn(dy, x)
    ref = _bf(dy).double().T @ _bf(x).double()
    assert _rel(dw, ref) < 1e-5, _rel(dw, ref)          # fp32 accumulation over up to 5120 rows
    assert float((dw.double() - ref).abs().max()) < 5e-3 * float(ref.abs().max())
    dw2 = ops.wgrad_mn(dy, x, accumulate_into=dw.clone())
    assert _rel(dw2, 2 * ref) < 1e-5
