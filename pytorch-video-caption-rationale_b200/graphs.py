"""CUDA-graph capture of a whole training step (forward + backward) of a drop-in module.

Every kernel of the library is stream-ordered and allocation-free (include/pvcr_b200.h), so one fwd+bwd step is
captured once and replayed with a single graph launch: host launch latency and jitter leave the critical path.
The step is the module's tape-free ``train_step_grads`` (no autograd state inside the capture).  Inputs are copied
into static device buffers (directly from pinned host memory when given CPU tensors); the gradients appear in
``param.grad`` (static buffers rewritten by every replay), ready for the all-reduce / optimizer.
"""
import os

import torch

from ._lib import check, lib, ptr, stream_ptr


class _InputPipeline:
    """Input side shared by the graphed steps: the next batch is copied host -> device into staging buffers on a side stream
    while the current step computes (what a pinned-memory DataLoader with non_blocking copies gives the reference loop)."""

    def _init_pipeline(self):
        self._copy_stream = torch.cuda.Stream()
        self._staging = tuple(torch.empty_like(t) for t in self.static_in)
        self._staged = None
        self._handover = None

    def prefetch(self, *inputs):
        """Start copying the NEXT step's inputs (pinned host tensors) to the device; returns immediately."""
        cs = self._copy_stream
        if self._handover is not None:
            cs.wait_event(self._handover)                 # staging buffers must have been handed over, nothing more
        with torch.cuda.stream(cs):
            for dst, src in zip(self._staging, inputs):
                dst.copy_(src, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(cs)

    def step_prefetched(self):
        """Run one step on the inputs handed to the latest prefetch()."""
        assert self._staged is not None, "call prefetch() first"
        torch.cuda.current_stream().wait_event(self._staged)
        for dst, src in zip(self.static_in, self._staging):
            dst.copy_(src, non_blocking=True)              # device-to-device hand-over (tens of microseconds)
        self._handover = torch.cuda.Event()
        self._handover.record(torch.cuda.current_stream())
        self._staged = None
        self._replay()
        return self.static_out


class GraphedTrainStep(_InputPipeline):
    def __init__(self, model, example_inputs, warmup=3, reducer=None, optimizer=None):
        """example_inputs: tuple of CUDA tensors, the arguments of model.train_step_grads (fixes shapes/dtypes).
        reducer: a parallel.GradAllReducer(flat=True, early=model.early_grad_params()); the step is then captured as
        two graphs split where the early gradients are final, and their all-reduce overlaps the second graph."""
        self.model = model
        self.reducer = reducer
        # optimizer: an optim.FusedClipAdam; its clip + Adam kernels are captured behind the backward (and the gradient
        # all-reduce), so one replay is a whole training iteration (train.py:157-160)
        self.optimizer = optimizer
        # per-replay dropout / Gumbel seeds: a device counter incremented by the first captured graph node
        self.seed_step = torch.zeros(1, dtype=torch.int64, device=example_inputs[0].device)
        lib().pvcr_set_seed_step(ptr(self.seed_step))
        GraphedTrainStep._seed_owner = self.seed_step.data_ptr()     # the library holds ONE counter pointer per process
        # the optimizer may only be captured behind the backward when the gradients it sees are final there: with a
        # reducer whose all-reduce runs outside the graph it has to run after reducer.finish() / reduce() instead
        self._opt_after_replay = False
        self.static_in = tuple(t.clone() for t in example_inputs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                model.train_step_grads(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.graphs = [self.graph]          # one graph per stage of model.train_step_stages (a single one otherwise)
        self.comm_in_graph = False
        staged = reducer is not None and hasattr(model, "train_step_stages")
        if staged and reducer.world > 1 and os.environ.get("PVCR_DP_STAGED") is None:
            # One graph for everything, NCCL included: every bucket's all-reduce is launched from an auxiliary stream
            # that waits for the stage's gradients (capture stream + the library's side lanes) without holding the
            # capture stream back, so the side lanes keep running across the stage boundaries (deferred joins) and
            # the collectives overlap the remaining backward.
            self.comm_in_graph = True
            aux = torch.cuda.Stream()

            def staged_step():
                main = torch.cuda.current_stream()
                gen = model.train_step_stages(*self.static_in, deferred_join=True)
                i = 0
                overlapped = getattr(model, "OVERLAPPED_STAGES", 1)
                if os.environ.get("PVCR_DP_OVERLAPPED"):           # tuning knob
                    overlapped = int(os.environ["PVCR_DP_OVERLAPPED"])
                while True:
                    try:
                        stage = next(gen)
                    except StopIteration as done:
                        res = done.value
                        break
                    aux.wait_stream(main)
                    with torch.cuda.stream(aux):
                        # aux waits for the lane(s) that produce this stage's gradients: one lane when the stage names
                        # it, else all of them
                        if isinstance(stage, tuple) and stage[1] == "milestone":
                            check(lib().pvcr_side_wait_milestone(stream_ptr(), int(stage[2])), "pvcr_side_wait_milestone")
                        elif isinstance(stage, tuple):
                            check(lib().pvcr_side_join_lane(stream_ptr(), int(stage[1])), "pvcr_side_join_lane")
                        else:
                            check(lib().pvcr_side_join(stream_ptr()), "pvcr_side_join")
                        # buckets that become final while a persistent sweep is still to come go through the narrow
                        # communicator (the sweep needs its 128 SMs co-resident); the later ones through the wide one
                        reducer.begin(i, tail=i >= overlapped)
                    i += 1
                main.wait_stream(aux)
                for j in range(i, len(reducer.buckets)):
                    reducer.begin(j, tail=True)
                reducer.finish()
                if optimizer is not None:
                    optimizer.step()
                return res

            with torch.no_grad():
                with torch.cuda.stream(side):
                    staged_step()                  # eager once: communicator / lane set-up outside the capture
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                with torch.cuda.graph(self.graph):
                    self.seed_step.add_(1)
                    out = staged_step()
        elif staged:
            self._opt_after_replay = optimizer is not None
            with torch.no_grad():
                gen = model.train_step_stages(*self.static_in)
                out = None
                while out is None:
                    g = self.graphs[-1]
                    with torch.cuda.graph(g, pool=None if g is self.graph else self.graph.pool()):
                        if g is self.graph:
                            self.seed_step.add_(1)
                        try:
                            next(gen)
                            more = True
                        except StopIteration as done:
                            out, more = done.value, False
                    if more:
                        self.graphs.append(torch.cuda.CUDAGraph())
        else:
            reduce_outside = reducer is not None and reducer.world > 1       # all-reduce runs after the replay
            self._opt_after_replay = optimizer is not None and reduce_outside
            capture_opt = optimizer is not None and not reduce_outside
            if capture_opt:                    # pointer table of the optimizer is built outside the capture
                with torch.cuda.stream(side):
                    model.train_step_grads(*self.static_in)
                    optimizer.step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
            with torch.cuda.graph(self.graph):
                self.seed_step.add_(1)
                out = model.train_step_grads(*self.static_in)
                if capture_opt:
                    optimizer.step()
        self.static_out = tuple(o.detach() if torch.is_tensor(o) else o for o in out)
        self._init_pipeline()

    _seed_owner = None

    def __del__(self):
        try:
            # the counter tensor dies with this object -- unless a younger GraphedTrainStep has registered its own since
            if GraphedTrainStep._seed_owner == self.seed_step.data_ptr():
                lib().pvcr_set_seed_step(None)
                GraphedTrainStep._seed_owner = None
        except Exception:                        # noqa: BLE001  (interpreter shutdown)
            pass

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self._replay()
        return self.static_out

    def _replay(self):
        if self.comm_in_graph:
            self.graph.replay()
        elif len(self.graphs) > 1 and self.reducer is not None:
            # stage i's gradients (bucket i) are final when graph i has run: their all-reduce overlaps graph i+1
            for i, g in enumerate(self.graphs):
                g.replay()
                if i + 1 < len(self.graphs):
                    self.reducer.begin(i)
            for i in range(len(self.graphs) - 1, len(self.reducer.buckets)):
                self.reducer.begin(i)
            self.reducer.finish()
        else:
            self.graph.replay()
            if self.reducer is not None:
                self.reducer.reduce()
        if self._opt_after_replay:             # clip + Adam on the REDUCED gradients (train.py:158-160)
            self.optimizer.step()


class GraphedAutogradStep(_InputPipeline):
    """CUDA-graph capture of ``loss = loss_fn(...); backward`` WITH the autograd tape inside the capture, for the drop-in
    modules whose step is a chain of autograd Functions rather than one tape-free ``train_step_grads`` -- SpatialNet
    (model/SpatialNet.py:100-142; train_spatial.py:30-39,  ~500 launches per fwd+bwd, host-bound in eager mode).

    ``loss_fn(*inputs)`` returns the loss or a tuple whose first element is the loss (the rest, e.g. accuracy and
    predictions, must not require grad).  ``example_inputs`` (CUDA tensors) fix the shapes: they are cloned into static
    buffers that ``__call__(*inputs)`` / ``prefetch`` + ``step_prefetched`` refill; without them ``loss_fn()`` reads tensors
    the caller keeps alive.  The gradients appear in ``param.grad`` (static buffers, rewritten by every replay).  With an
    ``optimizer`` (optim.FusedClipAdam) its clip + Adam kernels are captured behind the backward: one replay = one training
    iteration (train_spatial.py: run_iter + backward + clip_grad_norm_ + optimizer.step)."""

    def __init__(self, model, loss_fn, warmup=2, example_inputs=(), optimizer=None):
        self.model = model
        self.optimizer = optimizer
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.static_in = tuple(t.clone() for t in example_inputs)
        self.seed_step = torch.zeros(1, dtype=torch.int64, device=dev)
        lib().pvcr_set_seed_step(ptr(self.seed_step))
        GraphedTrainStep._seed_owner = self.seed_step.data_ptr()
        if optimizer is not None:
            # the optimizer's pointer table wants one gradient buffer per parameter at a fixed address: allocated here, the
            # captured backward copies into them (parameters the loss does not reach keep a zero gradient)
            for p in self.params:
                p.grad = torch.zeros_like(p)

        def step():
            # torch.autograd.grad, not .backward(): no AccumulateGrad nodes (they remember the stream of an earlier eager
            # iteration, and a hand-over to the default stream invalidates the capture)
            out = loss_fn(*self.static_in)
            loss = out[0] if isinstance(out, (tuple, list)) else out
            grads = torch.autograd.grad(loss, self.params, allow_unused=True)
            if optimizer is not None:
                with torch.no_grad():
                    for p, g in zip(self.params, grads):
                        if g is not None:
                            p.grad.copy_(g)
                optimizer.step()
            rest = tuple(o.detach() for o in out[1:]) if isinstance(out, (tuple, list)) else ()
            return (loss.detach(),) + rest, grads

        keep = None
        buffers = [(b, b.clone()) for b in model.buffers()]       # BatchNorm running statistics: the warm-up steps update them
        if optimizer is not None:          # the warm-up steps must not train the model
            keep = ([p.detach().clone() for p in self.params], [m.clone() for m in optimizer.exp_avg],
                    [v.clone() for v in optimizer.exp_avg_sq], optimizer.step_count.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.seed_step.add_(1)
            self.static_out, grads = step()
        if optimizer is None:
            for p, g in zip(self.params, grads):         # static buffers of the graph's pool, rewritten by every replay
                p.grad = g
        else:
            torch.cuda.synchronize()
            with torch.no_grad():
                for p, k in zip(self.params, keep[0]):
                    p.copy_(k)
                for m, k in zip(optimizer.exp_avg, keep[1]):
                    m.copy_(k)
                for v, k in zip(optimizer.exp_avg_sq, keep[2]):
                    v.copy_(k)
                optimizer.step_count.copy_(keep[3])
        with torch.no_grad():
            for b, k in buffers:
                b.copy_(k)
        self.static_loss = self.static_out[0]
        self._init_pipeline()

    def _replay(self):
        self.graph.replay()

    def replay(self):
        self.graph.replay()
        return self.static_loss

    def __call__(self, *inputs):
        if not inputs:
            return self.replay()
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def __del__(self):
        try:
            if GraphedTrainStep._seed_owner == self.seed_step.data_ptr():
                lib().pvcr_set_seed_step(None)
                GraphedTrainStep._seed_owner = None
        except Exception:                        # noqa: BLE001  (interpreter shutdown)
            pass


class GraphedGreedy:
    """CUDA-graph capture of fixed-length greedy captioning (model.greedy) for one batch shape: at small batches the
    step-wise decode is launch-latency bound, a single graph launch removes the host from the loop."""

    def __init__(self, model, example_vid, return_logits=True):
        """The warm-up call prepares the weights (model.greedy keeps them in its DecodePlan), so the captured graph holds
        the per-batch work only: it is valid for as long as the parameters keep their values (re-create it after training)."""
        self.model = model
        self.static_vid = example_vid.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model.greedy(self.static_vid, return_logits=return_logits)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = model.greedy(self.static_vid, return_logits=return_logits)
        self.static_out = tuple(o for o in out)

    def __call__(self, vid):
        if vid is not self.static_vid:
            self.static_vid.copy_(vid, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedBeam:
    """CUDA-graph capture of fixed-length beam search (model.beam_search) for one batch shape and beam width."""

    def __init__(self, model, example_vid, beam=5):
        self.static_vid = example_vid.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model.beam_search(self.static_vid, beam=beam)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = model.beam_search(self.static_vid, beam=beam)

    def __call__(self, vid):
        if vid is not self.static_vid:
            self.static_vid.copy_(vid, non_blocking=True)
        self.graph.replay()
        return self.static_out
