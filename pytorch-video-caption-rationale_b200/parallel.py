"""Data-parallel training over the videos of a batch: one process per GPU, parameters replicated, gradients
averaged with one all-reduce per bucket (NCCL over NVLink on GPUs; gloo in the CPU tests).

The reference is single-process; the contract here is that after ``reduce()`` every rank holds
the gradient of the mean loss over the concatenated global batch (equal shards: mean of per-rank means).
"""
import torch
import torch.distributed as dist


def shard_batch(tensors, rank, world):
    """Rank's contiguous slice of each [B, ...] tensor (B must divide evenly: videos are independent units)."""
    out = []
    for t in tensors:
        B = t.shape[0]
        assert B % world == 0, "global batch %d not divisible by world size %d" % (B, world)
        n = B // world
        out.append(t[rank * n:(rank + 1) * n])
    return out


class GradAllReducer:
    """Flat-bucket gradient averaging.  Buckets follow reverse parameter order (the order in which backward
    produces gradients: vocabulary projection and embedding first), so ``reduce()`` can be issued per bucket on a
    side stream while later buckets are still being computed."""

    def __init__(self, module, bucket_mb=64, group=None, flat=False, early=None, tail_group=None):
        """flat=True pre-allocates one flat fp32 buffer per bucket and installs views of it as ``param.grad``; the
        tape-free train steps write gradients straight into those views, so ``reduce()`` all-reduces in place
        without gather / scatter copies."""
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = group
        # tail_group: optional second process group (same ranks, a communicator allowed more CTAs) for the buckets whose
        # all-reduce cannot overlap the backward any more: the overlapped ones must stay within the SMs the persistent
        # sweeps leave free, the exposed ones should use the whole idle machine
        self.tail_group = tail_group
        # early: list of parameter lists whose gradients become final first (one bucket each, in that order)
        early = list(early or [])
        if early and not isinstance(early[0], (list, tuple)):
            early = [early]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # NCCL averages inside the collective (ncclAvg): no separate 1/G pass over the ~100 MB of gradients behind the
        # last all-reduce; other backends (gloo in the CPU tests) sum and divide
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self.buckets = [list(e) for e in early]
        self.n_early = len(self.buckets)
        early_ids = {id(p) for e in early for p in e}
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            if id(p) in early_ids:
                continue
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_mb * (1 << 20):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        import os
        self._skip = {int(x) for x in os.environ.get("PVCR_DP_SKIP", "").split(",") if x.strip()}
        self._flat = [None] * len(self.buckets)
        self._views = None
        self._pending = []
        if flat:
            self._views = []
            for i, bucket in enumerate(self.buckets):
                buf = torch.zeros(sum(p.numel() for p in bucket), dtype=torch.float32, device=bucket[0].device)
                self._flat[i] = buf
                off, views = 0, []
                for p in bucket:
                    v = buf[off:off + p.numel()].view_as(p)
                    p.grad = v
                    views.append(v)
                    off += p.numel()
                self._views.append(views)

    def _in_place(self, i):
        return self._views is not None and all(p.grad is not None and p.grad.data_ptr() == v.data_ptr()
                                               for p, v in zip(self.buckets[i], self._views[i]))

    def begin(self, i, tail=False):
        """Start the all-reduce of bucket i (asynchronous: later kernels on the current stream overlap with it)."""
        if self.world == 1:
            return
        assert self._in_place(i), "begin()/finish() need flat=True buckets written in place"
        if i in self._skip:              # tuning aid (PVCR_DP_SKIP="1,2"): leave a bucket un-reduced to find the exposed one
            return
        group = self.tail_group if (tail and self.tail_group is not None) else self.group
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._pending.append((i, dist.all_reduce(self._flat[i], op=op, group=group, async_op=True)))

    def finish(self):
        for i, h in self._pending:
            h.wait()
            if not self._avg:
                self._flat[i].div_(self.world)
        self._pending = []

    def reduce(self):
        if self.world == 1:
            return
        handles, in_place = [], []
        for i, bucket in enumerate(self.buckets):
            in_place.append(self._in_place(i))
            if in_place[i]:
                flat = self._flat[i]
            else:
                grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
                flat = torch.cat([g.reshape(-1).float() for g in grads])
                self._flat[i] = flat
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            handles.append(dist.all_reduce(flat, op=op, group=self.group, async_op=True))
        for i, (bucket, h) in enumerate(zip(self.buckets, handles)):
            h.wait()
            flat = self._flat[i]
            if not self._avg:
                flat.div_(self.world)
            if in_place[i]:
                continue
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n


def reduce_metrics(loss, correct, count, group=None):
    """Global loss mean and token accuracy from per-rank values (train_utils.py:37-71 semantics on the global batch)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return loss, correct / count
    v = torch.stack([loss.detach().float(), correct.float(), count.float()])
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v[0] / dist.get_world_size(group), v[1] / v[2]
