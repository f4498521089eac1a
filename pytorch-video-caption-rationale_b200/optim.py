"""Fused optimizer step of the reference training loop (train.py:104-105,157-160):

    nn.utils.clip_grad_norm_(model.parameters(), max_norm);  torch.optim.Adam(lr=..., weight_decay=...).step()

as two multi-tensor CUDA kernels (pvcr_adam_clip_step) with no host synchronisation: capturable in a CUDA graph right
behind the fwd+bwd step (and the gradient all-reduce).  State and semantics follow torch.optim.Adam (exp_avg,
exp_avg_sq, step; L2-style weight decay added to the gradient), so ``state_dict()`` round-trips with it.
"""
import ctypes

import torch

from ._lib import check, lib, ptr, stream_ptr

CHUNK = 1 << 16


class FusedClipAdam:
    def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=None):
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params), \
            "FusedClipAdam runs on contiguous fp32 CUDA parameters only (no CPU path)"
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_norm = 0.0 if max_norm is None else float(max_norm)
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)        # device-resident (graph replays advance it)
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=dev)      # gradient norm before clipping
        ct, co = [], []
        for i, p in enumerate(self.params):
            for off in range(0, p.numel(), CHUNK):
                ct.append(i); co.append(off)
        self.n_chunks = len(ct)
        self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self._chunk_off = torch.tensor(co, dtype=torch.int64, device=dev)
        self._partial = torch.zeros(self.n_chunks, dtype=torch.float32, device=dev)
        self._table = None
        self._table_key = None

    def _tensor_table(self):
        grads = []
        for p in self.params:
            assert p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32, "missing / non-contiguous grad"
            grads.append(p.grad)
        key = tuple(g.data_ptr() for g in grads) + tuple(p.data_ptr() for p in self.params)
        if key != self._table_key:         # pointers change only when param.grad is re-bound (never with flat buckets)
            rows = [[p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()]
                    for p, g, m, v in zip(self.params, grads, self.exp_avg, self.exp_avg_sq)]
            self._table = torch.tensor(rows, dtype=torch.int64).to(self.params[0].device)
            self._table_key = key
        return self._table

    @torch.no_grad()
    def step(self):
        """clip + Adam on the current ``param.grad``s; returns the (device) total gradient norm before clipping."""
        table = self._tensor_table()
        check(lib().pvcr_adam_clip_step(ptr(table), ptr(self._chunk_tensor), ptr(self._chunk_off), self.n_chunks, CHUNK,
                                        self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.max_norm,
                                        ptr(self.step_count), 0, ptr(self._partial), ptr(self.total_norm), stream_ptr()),
              "pvcr_adam_clip_step")
        return self.total_norm

    def zero_grad(self, set_to_none=False):
        """The tape-free train steps overwrite every gradient: nothing to clear (kept for optimizer-API compatibility)."""
        if set_to_none:
            for p in self.params:
                p.grad = None

    def state_dict(self):
        step = int(self.step_count.item())
        return {"state": {i: {"step": torch.tensor(float(step)), "exp_avg": m, "exp_avg_sq": v}
                          for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq))},
                "param_groups": [{"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                                  "params": list(range(len(self.params)))}]}

    def load_state_dict(self, sd):
        for i, st in sd["state"].items():
            self.exp_avg[int(i)].copy_(st["exp_avg"]); self.exp_avg_sq[int(i)].copy_(st["exp_avg_sq"])
            self.step_count.fill_(int(float(st["step"])))
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
