"""On-hardware data-parallel correctness (ADVICE r1 / VERDICT r1 "weak" #6): N ranks training through
GraphedTrainStep + GradAllReducer + FusedClipAdam end with the same parameters as one process training on the
concatenated batch -- for the staged in-graph NCCL path (S2VTAtt, RationaleNet) and for a model without stages (S2VT:
all-reduce behind the replay, optimizer behind the all-reduce).  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind", ["s2vtatt", "rationale", "s2vt"])
def test_data_parallel_training_matches_single_process(kind):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29611 + ["s2vtatt", "rationale", "s2vt"].index(kind)),
           os.path.join(ROOT, "tests", "gpu_dp_worker.py"), kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DP_OK " + kind in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    print(r.stdout.strip().splitlines()[-1])
