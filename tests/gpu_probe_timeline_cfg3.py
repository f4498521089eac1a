"""Tuning aid: per-launch timeline (CUDA events on the launching streams) of one fwd+bwd step at cfg2, eager or
(argument "graph") as recorded inside a CUDA-graph replay."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200
from pvcr_b200 import _lib
from pvcr_b200.model import RationaleNet
from tests.gpu_util import FixtureGlove

B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
m = RationaleNet(FixtureGlove(Vc, E), 0.2, H, V, L, 1.0, "s2vt-att").cuda().train()
vid = torch.randn(B, N, V, device="cuda"); s = torch.randint(0, Vc - 4, (B, L), device="cuda")
s_len = torch.randint(1, L + 1, (B,), device="cuda")
for _ in range(3): m.train_step_grads(vid, s, s_len)
torch.cuda.synchronize()
Lb = _lib.lib()
if len(sys.argv) > 1 and sys.argv[1] == "graph":
    from pvcr_b200.graphs import GraphedTrainStep
    Lb.pvcr_prof_reset(); Lb.pvcr_prof_enable(2)
    step = GraphedTrainStep(m, (vid, s, s_len), warmup=0)
    Lb.pvcr_prof_enable(0)
    for _ in range(3): step(vid, s, s_len)
    torch.cuda.synchronize()
else:
    Lb.pvcr_prof_reset(); Lb.pvcr_prof_enable(1)
    m.train_step_grads(vid, s, s_len)
    torch.cuda.synchronize()
cap = 512
cls = (ctypes.c_int * cap)(); t0 = (ctypes.c_float * cap)(); t1 = (ctypes.c_float * cap)()
n = Lb.pvcr_prof_timeline(cls, t0, t1, cap)
names = [Lb.pvcr_prof_class_name(i).decode() for i in range(Lb.pvcr_prof_num_classes())]
for i in range(n):
    print("%3d %-24s %8.1f -> %8.1f  (%6.1f us)" % (i, names[cls[i]], t0[i] * 1e3, t1[i] * 1e3, (t1[i] - t0[i]) * 1e3))
