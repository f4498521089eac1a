"""Thin torch-tensor wrappers over the granular C-ABI ops (device memory and streams come from torch)."""
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _req(t, dtype=torch.float32):
    assert t.is_cuda and t.dtype == dtype and t.stride(-1) == 1, (t.device, t.dtype, t.stride())
    return t


def linear_fwd(x, w, bias=None, nsplit=1):
    """y = x @ w.T + bias on the tcgen05 GEMM.  x [M,K], w [N,K] fp32 CUDA tensors (row stride arbitrary)."""
    _req(x); _req(w)
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    L = lib()
    ws = _ws(L.pvcr_linear_fwd_workspace(M, N, K, nsplit), x.device)
    check(L.pvcr_linear_fwd(ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(y), y.stride(0), M, N, K, nsplit,
                            ptr(ws), ws.numel(), stream_ptr()), "pvcr_linear_fwd")
    return y


def linear_bwd(dy, x, w, need_dx=True, need_dw=True, need_db=True, nsplit=1):
    _req(dy); _req(x); _req(w)
    M, N = dy.shape
    K = x.shape[1]
    dx = torch.empty((M, K), dtype=torch.float32, device=x.device) if need_dx else None
    dw = torch.empty((N, K), dtype=torch.float32, device=x.device) if need_dw else None
    db = torch.empty((N,), dtype=torch.float32, device=x.device) if need_db else None
    L = lib()
    ws = _ws(L.pvcr_linear_bwd_workspace(M, N, K, nsplit), x.device)
    check(L.pvcr_linear_bwd(ptr(dy), dy.stride(0), ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(dx), K, ptr(dw), K,
                            ptr(db), M, N, K, nsplit, 0, ptr(ws), ws.numel(), stream_ptr()), "pvcr_linear_bwd")
    return dx, dw, db


def wgrad_mn(dy, x, accumulate_into=None):
    """dw = dy.T @ x (bf16 operands cast in place, MN-major tcgen05 operands, fp32 accumulation)."""
    _req(dy); _req(x)
    R, N = dy.shape
    K = x.shape[1]
    dw = accumulate_into if accumulate_into is not None else torch.empty((N, K), dtype=torch.float32, device=x.device)
    L = lib()
    ws = _ws(L.pvcr_wgrad_mn_workspace(R, N, K), x.device)
    check(L.pvcr_wgrad_mn(ptr(dy), dy.stride(0), ptr(x), x.stride(0), ptr(dw), dw.stride(0), R, N, K,
                          int(accumulate_into is not None), ptr(ws), ws.numel(), stream_ptr()), "pvcr_wgrad_mn")
    return dw
