"""Per-kernel-class event timing of one eager fwd+bwd step for a chosen config (tuning aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200
from pvcr_b200 import _lib
from pvcr_b200.model import RationaleNet, S2VTAttModel, S2VTModel
from tests.gpu_util import FixtureGlove

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
if which == "cfg1":
    B, N, V, H, E, L, Vc = 32, 80, 4096, 512, 300, 28, 10000
    m = S2VTModel(FixtureGlove(Vc, E), 0.2, H, V, L)
elif which == "cfg3":
    B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
    m = RationaleNet(FixtureGlove(Vc, E), 0.2, H, V, L, 1.0, "s2vt-att")
else:
    B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
    m = S2VTAttModel(FixtureGlove(Vc, E), 0.2, H, V, L)
m = m.cuda().train()
vid = torch.randn(B, N, V, device="cuda"); s = torch.randint(0, Vc - 4, (B, L), device="cuda")
s_len = torch.randint(1, L + 1, (B,), device="cuda")
for _ in range(3): m.train_step_grads(vid, s, s_len)
torch.cuda.synchronize()
Lb = _lib.lib()
Lb.pvcr_prof_reset(); Lb.pvcr_prof_enable(1)
for _ in range(3): m.train_step_grads(vid, s, s_len)
torch.cuda.synchronize()
tot = 0
for k, v in _lib.prof_read().items():
    if v[0]:
        print("%-26s launches/step %6.1f  ms/step %.3f" % (k, v[0] / 3, v[1] / 3)); tot += v[1] / 3
print("sum %.3f ms" % tot)
