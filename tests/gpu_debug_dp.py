"""Debug aid: single-graph data-parallel step (NCCL inside the graph) at a small shape, 2 ranks, with progress prints."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import numpy as np
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_MAX_CTAS", "16")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pvcr_b200
from pvcr_b200.model import S2VTAttModel
from pvcr_b200.parallel import GradAllReducer
from pvcr_b200.graphs import GraphedTrainStep
from tests.gpu_util import FixtureGlove
def log(*a):
    print("[r%d %.1f]" % (rank, time.time() % 1000), *a, flush=True)
B, N, V, H, E, L, Vc = 128, 40, 2048, 512, 300, 30, 23000
torch.manual_seed(1)
m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L).cuda().train()
red = GradAllReducer(m, flat=True, early=m.early_grad_params())
g = torch.Generator().manual_seed(100 + rank)
vid = torch.randn(B, N, V, generator=g).cuda(); s = torch.randint(0, Vc - 4, (B, L), generator=g).cuda()
s_len = torch.randint(1, L + 1, (B,), generator=g).cuda()
for _ in range(2): m.train_step_grads(vid, s, s_len)
torch.cuda.synchronize(); log("eager ok")
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize(); log("nccl ok", t[0].item())
step = GraphedTrainStep(m, (vid, s, s_len), warmup=0, reducer=red)
torch.cuda.synchronize(); log("captured, comm_in_graph =", step.comm_in_graph)
for i in range(3):
    out = step(vid, s, s_len); torch.cuda.synchronize(); log("replay", i, out[0].item())
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): step(vid, s, s_len)
e1.record(); torch.cuda.synchronize(); log("ms/step", e0.elapsed_time(e1) / 10)
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
