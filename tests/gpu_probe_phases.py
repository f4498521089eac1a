"""Tuning aid: per-phase clock64 breakdown of the persistent GRU forward kernel (CTA 0), cfg2 encoder shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pvcr_b200
from pvcr_b200 import _lib
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove

B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L).cuda().train()
vid = torch.randn(B, N, V, device="cuda"); s = torch.randint(0, Vc - 4, (B, L), device="cuda")
s_len = torch.randint(1, L + 1, (B,), device="cuda")
Lb = _lib.lib()
os.environ.setdefault("X", "1")
for _ in range(3): m.train_step_grads(vid, s, s_len)
torch.cuda.synchronize()
Lb.pvcr_debug_phase_timing(1)
which = sys.argv[1] if len(sys.argv) > 1 else "enc_fwd"
# the debug buffer is overwritten by every persistent launch that stamps: run only the forward sequence
from pvcr_b200 import functional as F_
bwd = os.environ.get("PVCR_PHASE_DEC_BWD") is not None
with torch.no_grad():
    if bwd:
        m.train_step_grads(vid, s, s_len)
    else:
        F_.S2VTAttSequence.forward(F_.ManualCtx(), m._cfg(True), vid, None, m._shifted(s, B), *m._seq_params())
gru = os.environ.get("PVCR_PHASE_GRU") is not None
steps = N if gru else L
buf = (ctypes.c_longlong * (steps * 16))()
_lib.check(Lb.pvcr_debug_phase_read(buf, steps), "read")
a = np.array(buf[:]).reshape(steps, 16)
if bwd:
    a2 = np.concatenate([a[:, :11], np.roll(a[:, 0:1], -1, axis=0)], axis=1)[:-1]
    names = ["B1 gate grads", "B1 arrive+wait", "B2 4 chunks load+mma issue", "B2 wait mma A", "B2 t2s+dctx write+arrive",
             "B3 wait dctx", "B3 attention grad", "B3 arrive+wait", "B4 load+mma issue", "B4 wait mma", "B4 t2s+carry"]
elif gru:
    a2 = a[:, [0, 1, 7, 2, 3, 4, 5, 6]]
    names = ["wait", "load X (ld+st)", "fence+sync", "mma", "tmem->smem", "gates+stores", "arrive"]
else:
    a2 = np.concatenate([a[:, [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 6, 7]], np.roll(a[:, 0:1], -1, axis=0)], axis=1)[:-1]
    names = ["P1 wait h", "P1 load X", "P1 mma", "P1 t2s+q write+arrive", "P2 wait q", "P2 q load", "P2 tanh+shfl",
             "P2 sync+score+sync", "P2 softmax+ctx partial", "P2 sync", "P2 ctx reduce+write", "P2 arrive",
             "P3 wait ctx", "P3 load+mma+P4 gates+arrive"]
d = np.diff(a2, axis=1)[5:]           # skip the first steps
clk = 1.965e3  # cycles per us at max clock
print("per-step cycles (median over steps 5..):")
for i, n in enumerate(names):
    print("  %-14s %8.0f cyc  %.2f us" % (n, np.median(d[:, i]), np.median(d[:, i]) / clk))
tot = np.median(a[6:, 0] - a[5:-1, 0])
print("  step total     %8.0f cyc  %.2f us" % (tot, tot / clk))
