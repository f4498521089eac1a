"""Tuning aid: per-phase clock64 breakdown of the fp32 persistent GRU kernel of a greedy decode (CTA 0), cfg2 encoder shape.
Run with PVCR_PHASE_GRU_F32=1."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pvcr_b200
from pvcr_b200 import _lib
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove

B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L).cuda().eval()
vid = torch.randn(B, N, V, device="cuda")
Lb = _lib.lib()
for _ in range(2): m.greedy(vid, return_logits=False)
torch.cuda.synchronize()
Lb.pvcr_debug_phase_timing(1)
m.greedy(vid, return_logits=False)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (N * 16))()
_lib.check(Lb.pvcr_debug_phase_read(buf, N), "read")
a = np.array(buf[:]).reshape(N, 16)
a2 = np.concatenate([a[:, :5], np.roll(a[:, 0:1], -1, axis=0)], axis=1)[:-1]
names = ["gi prefetch+wait", "load h -> smem", "matvec", "gates+stores", "arrive"]
d = np.diff(a2, axis=1)[5:]
clk = 1.965e3
for i, n in enumerate(names):
    print("  %-18s %8.0f cyc  %.2f us" % (n, np.median(d[:, i]), np.median(d[:, i]) / clk))
tot = np.median(a[6:, 0] - a[5:-1, 0])
print("  step total         %8.0f cyc  %.2f us" % (tot, tot / clk))
