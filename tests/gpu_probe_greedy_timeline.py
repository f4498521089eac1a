"""Tuning aid: per-launch timeline (CUDA events on the launching stream) of one greedy decode of a batch at cfg2 dims."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200
from pvcr_b200 import _lib
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, V, H, E, L, Vc = 40, 2048, 512, 300, 30, 23000
m = S2VTAttModel(FixtureGlove(Vc, E), 0.2, H, V, L).cuda().eval()
vid = torch.randn(B, N, V, device="cuda")
RL = os.environ.get("PVCR_PROBE_NOLOGITS") is None
with torch.no_grad():
    for _ in range(2): m.greedy(vid, return_logits=RL)
torch.cuda.synchronize()
Lb = _lib.lib()
Lb.pvcr_prof_reset(); Lb.pvcr_prof_enable(1)
with torch.no_grad():
    m.greedy(vid, return_logits=RL)
torch.cuda.synchronize()
Lb.pvcr_prof_enable(0)
cap = 2048
cls = (ctypes.c_int * cap)(); t0 = (ctypes.c_float * cap)(); t1 = (ctypes.c_float * cap)()
n = Lb.pvcr_prof_timeline(cls, t0, t1, cap)
names = [Lb.pvcr_prof_class_name(i).decode() for i in range(Lb.pvcr_prof_num_classes())]
for i in range(n):
    print("%3d %-24s %8.1f -> %8.1f  (%6.1f us)" % (i, names[cls[i]], t0[i] * 1e3, t1[i] * 1e3, (t1[i] - t0[i]) * 1e3))
