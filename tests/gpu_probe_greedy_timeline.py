"""Tuning aid: in-graph per-launch timeline of the first decoding steps of GraphedGreedy at cfg2 dims (B = 128)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200
from pvcr_b200 import _lib
from pvcr_b200.graphs import GraphedGreedy
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove

B, N, V, H, E, L, Vc = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 40, 2048, 512, 300, 30, 23000
m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L).cuda().eval()
vid = torch.randn(B, N, V, device="cuda")
m.greedy(vid); torch.cuda.synchronize()
Lb = _lib.lib()
Lb.pvcr_prof_reset(); Lb.pvcr_prof_enable(2)
gg = GraphedGreedy(m, vid)
Lb.pvcr_prof_enable(0)
for _ in range(3): gg(vid)
torch.cuda.synchronize()
cap = 1024
cls = (ctypes.c_int * cap)(); t0 = (ctypes.c_float * cap)(); t1 = (ctypes.c_float * cap)()
n = Lb.pvcr_prof_timeline(cls, t0, t1, cap)
names = [Lb.pvcr_prof_class_name(i).decode() for i in range(Lb.pvcr_prof_num_classes())]
print("launches", n)
for i in range(min(n, 60)):
    print("%3d %-24s %8.1f -> %8.1f  (%6.1f us)" % (i, names[cls[i]], t0[i] * 1e3, t1[i] * 1e3, (t1[i] - t0[i]) * 1e3))
print("...")
for i in range(max(0, n - 12), n):
    print("%3d %-24s %8.1f -> %8.1f  (%6.1f us)" % (i, names[cls[i]], t0[i] * 1e3, t1[i] * 1e3, (t1[i] - t0[i]) * 1e3))
