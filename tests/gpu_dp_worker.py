"""Worker of tests/test_gpu_data_parallel.py (one process per GPU under torch.distributed.run): trains a drop-in module
data-parallel for a few steps through GraphedTrainStep (+ GradAllReducer + FusedClipAdam) and compares the resulting
parameters with single-process training on the concatenated batch (the correctness oracle of the
data-parallel path; the reference itself is single-process, D8).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_dp_worker.py s2vtatt|s2vt|rationale
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import pvcr_b200  # noqa: F401
from oracle import workloads as W
from pvcr_b200.graphs import GraphedTrainStep
from pvcr_b200.model import RationaleNet, S2VTAttModel, S2VTModel
from pvcr_b200.optim import FusedClipAdam
from pvcr_b200.parallel import GradAllReducer, shard_batch
from tests.gpu_util import FixtureGlove


def build(kind, dims, seed):
    B, N, V, H, E, L, Vc = dims
    torch.manual_seed(seed)
    if kind == "s2vtatt":
        m = S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in W.s2vtatt_params(V, H, E, Vc, seed).items()})
    elif kind == "s2vt":
        m = S2VTModel(FixtureGlove(Vc, E), 0.0, H, V, L)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in W.s2vt_params(V, H, E, Vc, seed).items()})
    else:
        m = RationaleNet(FixtureGlove(Vc, E), 0.0, H, V, L, 1.0, "s2vt-att")
        sd = {"caption_net." + k: torch.from_numpy(v) for k, v in W.s2vtatt_params(V, H, E, Vc, seed).items()}
        sd.update({"gen." + k: torch.from_numpy(v) for k, v in W.generator_params(V, H, seed + 1).items()})
        m.load_state_dict(sd)
    return m.cuda().train()


def main():
    kind = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    per_rank, steps = 16, 3
    dims = (per_rank * world, 40, 256, 512, 300, 30, 1200)
    B, N, V, H, E, L, Vc = dims
    vid, s, s_len = (torch.from_numpy(x).cuda() for x in W.make_batch(B, N, V, L, Vc, 900))
    noise = torch.from_numpy(np.random.RandomState(5).exponential(size=(B * N, 2)).astype(np.float32)).cuda()
    opt_args = dict(lr=1e-3, weight_decay=4e-5, max_norm=1.0)

    # ---- data parallel: every rank its shard, gradients averaged inside / behind the captured step ------------------
    m = build(kind, dims, 11)
    if kind == "rationale":
        m.gen.noise = noise[rank * per_rank * N:(rank + 1) * per_rank * N].contiguous()     # rows b*N + n of this shard
    early = m.early_grad_params() if hasattr(m, "early_grad_params") else None
    reducer = GradAllReducer(m, flat=True, early=early)
    opt = FusedClipAdam(m.parameters(), **opt_args)
    shard = tuple(t.contiguous() for t in shard_batch((vid, s, s_len), rank, world))
    step = GraphedTrainStep(m, shard, warmup=0, reducer=reducer, optimizer=opt)
    done = int(opt.step_count.item())              # the capture may have run eager steps (communicator set-up)
    for _ in range(steps - done):
        step(*shard)
    torch.cuda.synchronize()
    assert int(opt.step_count.item()) == steps, (int(opt.step_count.item()), steps)

    # ---- single process, concatenated batch ---------------------------------------------------------------------------
    ref = build(kind, dims, 11)
    if kind == "rationale":
        ref.gen.noise = noise
    ropt = FusedClipAdam(ref.parameters(), **opt_args)
    for _ in range(steps):
        ref.train_step_grads(vid, s, s_len)
        ropt.step()
    torch.cuda.synchronize()

    # Adam turns tiny gradient differences (two 16-video GEMMs + all-reduce vs one 32-video GEMM) into lr-sized ones on
    # elements whose gradient is near zero, so parameters agree to a fraction of the distance they moved, not to 1e-6;
    # what the comparison must exclude is an optimizer fed with un-reduced gradients (replicas would differ: checked
    # exactly below), a missing 1/G (every update would differ in its clip factor) or no update at all.
    init = build(kind, dims, 11)
    worst = (0.0, "")
    for (k, a), (_, b), (_, c) in zip(m.named_parameters(), ref.named_parameters(), init.named_parameters()):
        moved = float((b - c).norm().item())
        diff = float((a - b).norm().item())
        rel = diff / max(float(b.norm().item()), 1e-30)
        worst = max(worst, (rel, k))
        assert moved > 0.0, (kind, k, "parameter never updated")
        assert diff < 0.05 * moved, (kind, k, diff, moved)
        assert rel < 2e-3, (kind, k, rel)
    # replicas stay identical
    for k, a in m.named_parameters():
        lo, hi = a.detach().clone(), a.detach().clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), (kind, k)
    if rank == 0:
        print("DP_OK %s world=%d steps=%d worst rel %.2e (%s) comm_in_graph=%s graphs=%d" % (
            kind, world, steps, worst[0], worst[1], step.comm_in_graph, len(step.graphs)), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)          # see bench.py:_finish_ranks (communicator teardown under a live graph with captured NCCL kernels)


if __name__ == "__main__":
    main()
