"""Timing probe: S2VTAtt fwd+bwd at cfg2 (prints ms/iter; not a test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pvcr_b200
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove

B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
torch.manual_seed(123)
m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision=prec).cuda()
m.decoder.embedding.weight.data.normal_()
vid = torch.randn(B, N, V, device="cuda")
s_len = torch.randint(1, L + 1, (B,), device="cuda")
s = torch.randint(0, Vc - 4, (B, L), device="cuda")
m.train()
def step():
    m.zero_grad(set_to_none=True)
    loss, acc, pred = m.forward_loss(vid, s, s_len)
    loss.backward()
    return loss
for _ in range(3): l = step()
torch.cuda.synchronize()
print("loss", l.item(), flush=True)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
print("%s: %.3f ms/iter (device), %.3f ms/iter (wall) -> %.0f videos/s" % (prec, e0.elapsed_time(e1) / 10, (time.time() - t0) * 100, B / (e0.elapsed_time(e1) / 10) * 1e3))
from pvcr_b200.graphs import GraphedTrainStep
gs = GraphedTrainStep(m, (vid, s, s_len))
for _ in range(3): out = gs(vid, s, s_len)
torch.cuda.synchronize()
print("graphed loss", out[0].item())
t0 = time.time(); e0.record()
for _ in range(10): gs(vid, s, s_len)
e1.record(); torch.cuda.synchronize()
print("%s graphed: %.3f ms/iter (device), %.3f ms/iter (wall) -> %.0f videos/s" % (prec, e0.elapsed_time(e1) / 10, (time.time() - t0) * 100, B / (e0.elapsed_time(e1) / 10) * 1e3))
print("max mem MB", torch.cuda.max_memory_allocated() / 1e6)
