"""Throughput probes for the other BASELINE.json configs (not the driver's bench line):
cfg1 S2VT (MSVD shape), cfg3 RationaleNet + S2VTAtt joint training, cfg5 greedy captions/sec batch sweep."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvcr_b200
from pvcr_b200.graphs import GraphedBeam, GraphedGreedy, GraphedTrainStep
from pvcr_b200.model import RationaleNet, S2VTAttModel, S2VTModel
from tests.gpu_util import FixtureGlove


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def batch(B, N, V, L, Vc):
    vid = torch.randn(B, N, V, device="cuda")
    s = torch.randint(0, Vc - 4, (B, L), device="cuda")
    s_len = torch.randint(1, L + 1, (B,), device="cuda")
    return vid, s, s_len


out = {}
torch.manual_seed(123)
# cfg1: S2VTModel, MSVD shape
B, N, V, H, E, L, Vc = 32, 80, 4096, 512, 300, 28, 10000
m = S2VTModel(FixtureGlove(Vc, E), 0.2, H, V, L).cuda().train()
m.embedding[0].weight.data.normal_(0, 0.4)
g = GraphedTrainStep(m, batch(B, N, V, L, Vc))
ms = timed(lambda: g(*g.static_in), 20)
out["cfg1_s2vt_msvd_train"] = {"B": B, "ms_per_step": ms, "videos_per_s": B / ms * 1e3}
del g, m
# cfg3: RationaleNet + S2VTAtt joint training, cfg2 dims
B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
m = RationaleNet(FixtureGlove(Vc, E), 0.2, H, V, L, 1.0, "s2vt-att").cuda().train()
m.caption_net.decoder.embedding.weight.data.normal_(0, 0.4)
g = GraphedTrainStep(m, batch(B, N, V, L, Vc))
ms = timed(lambda: g(*g.static_in), 20)
out["cfg3_rationale_s2vtatt_train"] = {"B": B, "ms_per_step": ms, "videos_per_s": B / ms * 1e3}
del g, m
torch.cuda.empty_cache()
# cfg5: greedy decoding sweep (fp32-equivalent bf16x3 arithmetic, ids bit-exact vs the fp32 reference)
m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L).cuda().eval()
m.decoder.embedding.weight.data.normal_(0, 0.4)
sweep = {}
for B in (1, 8, 64, 128, 512, 1024):
    vid = torch.randn(B, N, V, device="cuda")
    ms = timed(lambda: m.greedy(vid), 5 if B >= 512 else 10)
    gg = GraphedGreedy(m, vid)
    ms_g = timed(lambda: gg(vid), 5 if B >= 512 else 10)
    sweep[B] = {"ms_per_batch": ms, "captions_per_s": B / ms * 1e3, "graph_ms_per_batch": ms_g,
                "graph_captions_per_s": B / ms_g * 1e3}
    del gg
out["cfg5_greedy_s2vtatt"] = sweep
# cfg5: beam-5 decoding sweep (pvcr_s2vtatt_beam; eager launches)
bsweep = {}
for B in (1, 8, 64, 128, 512, 1024):
    vid = torch.randn(B, N, V, device="cuda")
    ms = timed(lambda: m.beam_search(vid, beam=5), 3 if B >= 512 else 5)
    gb = GraphedBeam(m, vid, beam=5)
    ms_g = timed(lambda: gb(vid), 3 if B >= 512 else 5)
    bsweep[B] = {"ms_per_batch": ms, "captions_per_s": B / ms * 1e3, "graph_ms_per_batch": ms_g,
                 "graph_captions_per_s": B / ms_g * 1e3}
    del gb
out["cfg5_beam5_s2vtatt"] = bsweep
print(json.dumps(out, indent=1))
