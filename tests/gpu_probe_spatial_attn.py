"""Micro-benchmark of SpatialNet's per-frame attention kernels at the cfg4 shape (128 videos, 36 cells, keys 512 wide, values 2048
wide, operands as the frame sweep hands them over: one frame of [B, N, Kc, .] tensors by batch stride), forward and backward timed
separately with CUDA events over all 40 frames (the 1.9 GB working set is larger than L2).
    PVCR_SPATIAL_ATTN_NT=256|512 python tests/gpu_probe_spatial_attn.py save.pt [compare.pt]  -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200  # noqa: F401
from pvcr_b200._lib import check, lib, ptr, stream_ptr

B, N, Kc, H, F = 128, 40, 36, 512, 2048
torch.manual_seed(3)
pk = torch.randn(B, N, Kc, H, device="cuda")
feats = torch.randn(B, N, Kc, F, device="cuda")
q = torch.randn(N, B, 4 * H, device="cuda")
v = torch.randn(H, device="cuda") / H ** 0.5
dctx = torch.randn(B, F, device="cuda")
alpha = torch.empty(N, B, Kc, device="cuda")
ctx = torch.empty(N, B, F, device="cuda")
d1 = torch.zeros(N, B, 4 * H, device="cuda")
dpk = torch.empty(B, N, Kc, H, device="cuda")
dvp = torch.empty(N, B, H, device="cuda")
L = lib()


def fwd():
    for t in range(N):
        check(L.pvcr_spatial_attn_fwd(B, Kc, H, F, ptr(q[t]), 4 * H, ptr(pk[:, t]), N * Kc * H, ptr(feats[:, t]), N * Kc * F, ptr(v),
                                      ptr(alpha[t]), ptr(ctx[t]), stream_ptr()), "fwd")


dq = torch.empty(N, B, H, device="cuda")
dpk_f = torch.empty(N, B, Kc, H, device="cuda")


def bwd():
    for t in range(N):
        check(L.pvcr_spatial_attn_bwd(B, Kc, H, F, ptr(dctx), ptr(q[t]), 4 * H, ptr(pk[:, t]), N * Kc * H, ptr(feats[:, t]), N * Kc * F,
                                      ptr(v), ptr(alpha[t]), ptr(dq[t]), ptr(dpk_f[t]), ptr(dvp[t]), stream_ptr()), "bwd")


def timed(fn, it=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it / N * 1e3        # microseconds per launch


us_f, us_b = timed(fwd), timed(bwd)
bytes_f = B * (Kc * H + Kc * F + F + Kc + H) * 4
bytes_b = B * (2 * Kc * H + Kc * F + 2 * F + 3 * H) * 4
out = {"nt": os.environ.get("PVCR_SPATIAL_ATTN_NT", "default"), "fwd_us": us_f, "bwd_us": us_b,
       "fwd_GBps": bytes_f / us_f / 1e3, "bwd_GBps": bytes_b / us_b / 1e3}
res = {"alpha": alpha.cpu(), "ctx": ctx.cpu(), "dq": dq.cpu(), "dpk": dpk_f[::13].cpu(), "dvp": dvp.cpu()}
if len(sys.argv) > 2:
    ref = torch.load(sys.argv[2])
    out["max_abs_diff_vs_other"] = {k: float((res[k] - ref[k]).abs().max()) for k in res}
torch.save(res, sys.argv[1])
print(json.dumps(out))
