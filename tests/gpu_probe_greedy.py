"""Tuning aid: greedy decoding alone (cfg5 dims) as a CUDA graph, ms per batch with and without materialised logits.

    python tests/gpu_probe_greedy.py [B ...]      (default 128)
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200  # noqa: F401
from pvcr_b200.graphs import GraphedGreedy
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove

Bs = [int(x) for x in sys.argv[1:]] or [128]
N, V, H, E, L, Vc = 40, 2048, 512, 300, 30, 23000
torch.manual_seed(0)
m = S2VTAttModel(FixtureGlove(Vc, E), 0.2, H, V, L).cuda().eval()
out = {}
for B in Bs:
    vid = torch.randn(B, N, V, device="cuda")
    res = {}
    for name, rl in (("logits", True), ("ids_only", False)):
        with torch.no_grad():
            g = GraphedGreedy(m, vid, return_logits=rl)
            for _ in range(3):
                g(vid)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                g(vid)
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / iters
        del g
    res["captions_per_s"] = B / (res["logits"] / 1e3)
    out[B] = res
print(json.dumps(out))
