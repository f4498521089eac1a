import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200
from pvcr_b200 import functional as F_
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove
B, N, V, H, E, L, Vc = 16, 8, 128, 64, 32, 6, 200
m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L).cuda().train()
vid = torch.randn(B, N, V, device="cuda"); s = torch.randint(0, Vc - 4, (B, L), device="cuda"); sl = torch.randint(1, L + 1, (B,), device="cuda")
orig = F_.stream_ptr
def dbg():
    p = orig(); print("stream", p.value, flush=True); return p
F_.stream_ptr = dbg
params = [p for p in m.parameters()]
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    out = m.forward_loss(vid, s, sl); torch.autograd.grad(out[0], params)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
print("--- capture fwd only", flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        with torch.no_grad():
            hs, al = F_.S2VTAttSequence.apply(m._cfg(True), vid, None, m._shifted(s, B), *m._seq_params())
    print("fwd capture ok", flush=True)
except Exception as e:
    print("fwd capture failed", repr(e)[:300], flush=True)
torch.cuda.synchronize()
print("--- capture fwd+loss", flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        with torch.no_grad():
            out = m.forward_loss(vid, s, sl)
    print("fwd+loss capture ok", flush=True)
except Exception as e:
    print("fwd+loss capture failed", repr(e)[:300], flush=True)
print("--- capture fwd+bwd", flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        out = m.forward_loss(vid, s, sl)
        gr = torch.autograd.grad(out[0], params)
    print("full capture ok", flush=True)
except Exception as e:
    print("full capture failed", repr(e)[:300], flush=True)
