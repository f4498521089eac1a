"""Timing probe for the GEMM at the cfg2 shapes (prints TFLOP/s; not a test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200
from pvcr_b200 import _lib
from pvcr_b200._lib import lib, ptr, stream_ptr, check


def bench(M, N, K, nsplit=1, iters=20):
    L = lib()
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.05
    y = torch.empty(M, N, device="cuda")
    ws = torch.empty(L.pvcr_linear_fwd_workspace(M, N, K, nsplit), dtype=torch.uint8, device="cuda")
    def run():
        check(L.pvcr_linear_fwd(ptr(x), K, ptr(w), K, None, ptr(y), N, M, N, K, nsplit, ptr(ws), ws.numel(), stream_ptr()), "lin")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("linear_fwd(incl. staging) M=%d N=%d K=%d nsplit=%d: %.3f ms  %.1f TFLOP/s (logical)" % (M, N, K, nsplit, ms, 2.0 * M * N * K / ms / 1e9), flush=True)


if __name__ == "__main__":
    bench(5120, 1536, 2048)
    bench(3840, 23000, 512)
    bench(3840, 23000, 512, nsplit=3)
    bench(128, 2048, 512)
    bench(23000, 512, 3840)
