"""CPU: host-side logic of the drop-in modules and the data-parallel plumbing (gloo, world size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden_util import load
from tests.gpu_util import FixtureGlove


@pytest.mark.parametrize("tag,cls", [("s2vtatt_tiny", "S2VTAttModel"), ("s2vt_tiny", "S2VTModel"),
                                     ("rationale_att_tiny", "RationaleNet"), ("rationale_s2vt_tiny", "RationaleNet")])
def test_state_dict_keys_match_reference(tag, cls):
    """Reference checkpoints must load: same state_dict keys and shapes as the reference modules (golden params)."""
    import pvcr_b200.model as M
    d, params, _ = load(tag)
    B, N, V, H, E, L, Vc = (int(x) for x in d["dims"])
    g = FixtureGlove(Vc, E)
    if cls == "RationaleNet":
        m = M.RationaleNet(g, 0.2, H, V, L, 1.0, "s2vt-att" if "att" in tag else "s2vt")
    else:
        m = getattr(M, cls)(g, 0.2, H, V, L)
    sd = m.state_dict()
    assert set(sd) == set(params)
    for k, v in params.items():
        assert tuple(sd[k].shape) == v.shape, k
    m.load_state_dict({k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in params.items()})


def test_unknown_arch_raises_like_reference():
    import pvcr_b200.model as M
    with pytest.raises(NotImplementedError):
        M.RationaleNet(FixtureGlove(10, 4), 0.0, 8, 8, 3, 1.0, "transformer")


def test_training_requires_sentence():
    import pvcr_b200.model as M
    m = M.S2VTAttModel(FixtureGlove(10, 4), 0.0, 8, 8, 3).train()
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 2, 8), None)


def test_s2vt_init_is_ixvr():
    """S2VTModel applies Xavier-normal / bias 0.01 in its constructor (reference S2VTModel.py:51-55); S2VTAtt does not."""
    import pvcr_b200.model as M
    m = M.S2VTModel(FixtureGlove(10, 4), 0.0, 8, 8, 3)
    assert torch.all(m.rnn1.bias_ih_l0 == 0.01) and torch.all(m.linear[1].bias == 0.01)
    a = M.S2VTAttModel(FixtureGlove(10, 4), 0.0, 8, 8, 3)
    assert not torch.all(a.encoder.rnn.bias_ih_l0 == 0.01)


def test_shard_batch():
    from pvcr_b200.parallel import shard_batch
    x, y = torch.arange(24).view(8, 3), torch.arange(8)
    parts = [shard_batch((x, y), r, 4) for r in range(4)]
    assert torch.equal(torch.cat([p[0] for p in parts]), x) and torch.equal(torch.cat([p[1] for p in parts]), y)
    with pytest.raises(AssertionError):
        shard_batch((x,), 0, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pvcr_b200.parallel import GradAllReducer, reduce_metrics, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x, y = torch.randn(8, 6), torch.randn(8, 3)
    xs, ys = shard_batch((x, y), rank, world)
    red = GradAllReducer(model, bucket_mb=0)          # one bucket per parameter: exercises the bucket loop
    loss = ((model(xs) - ys) ** 2).mean()
    loss.backward()
    red.reduce()
    gl, acc = reduce_metrics(loss, torch.tensor(3.0 + rank), torch.tensor(4.0))
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in model.parameters()], "loss": gl, "acc": acc}, out)
    dist.destroy_process_group()


def test_data_parallel_grads_equal_single_process(tmp_path):
    """N-rank reduced gradients == single-process gradients on the concatenated batch."""
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x, y = torch.randn(8, 6), torch.randn(8, 3)
    loss = ((model(x) - y) ** 2).mean()
    loss.backward()
    for g, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(g, p.grad, atol=1e-6)
    assert abs(got["loss"].item() - loss.item()) < 1e-6
    assert abs(got["acc"].item() - (3.0 + 4.0) / 8.0) < 1e-6


def _dp_flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pvcr_b200.parallel import GradAllReducer, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x, y = torch.randn(8, 6), torch.randn(8, 3)
    xs, ys = shard_batch((x, y), rank, world)
    tail = dist.new_group(backend="gloo")             # second communicator for the buckets nothing overlaps any more
    early = [[model[2].bias, model[2].weight]]        # produced first by the backward: its own bucket, begun early
    red = GradAllReducer(model, flat=True, early=early, tail_group=tail)
    loss = ((model(xs) - ys) ** 2).mean()
    grads = torch.autograd.grad(loss, list(model.parameters()))
    for p, g in zip(model.parameters(), grads):       # the tape-free steps write into the flat-bucket views in place
        p.grad.copy_(g)
    red.begin(0)                                      # overlapped bucket: default group
    for j in range(1, len(red.buckets)):
        red.begin(j, tail=True)                       # exposed buckets: tail group
    red.finish()
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in model.parameters()], "n_buckets": len(red.buckets)}, out)
    dist.destroy_process_group()


def test_flat_buckets_begin_finish_with_tail_group(tmp_path):
    """The in-place flat-bucket path the CUDA-graph data-parallel step uses (begin per bucket as it becomes final, tail
    buckets on a second communicator, finish): reduced gradients == single-process gradients on the whole batch."""
    out = str(tmp_path / "dpflat.pt")
    mp.spawn(_dp_flat_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["n_buckets"] == 2
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x, y = torch.randn(8, 6), torch.randn(8, 3)
    ((model(x) - y) ** 2).mean().backward()
    for g, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(g, p.grad, atol=1e-6)
