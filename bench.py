#!/usr/bin/env python
"""Benchmark of the captioning hot path: train videos/sec (fwd+bwd) of S2VTAtt at the MSR-VTT shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|bf16x2|bf16x3]

One "step" = model.forward_loss(vid_feats, s, s_len) (encoder GRU, attention decoder, vocabulary projection +
masked cross entropy) followed by the backward pass producing every parameter gradient (train.py:37-40,157-158 of
the reference; optimizer excluded), on one batch of 128 synthetic videos per GPU (BASELINE.json configs[1]).
With N > 1 (torchrun, one rank per GPU) each rank runs its own 128 videos and the gradients are averaged by
NCCL all-reduce inside the timed step (weak scaling).  Rank 0 prints one JSON line.

--impl reference times the reference's own implementation on the host CPU cores: the UNMODIFIED reference modules
(oracle/_ref, the bytecode oracle/build_ref.py compiles from /root/reference) in torch fp32 on the whole 128-video
batch, thread count set explicitly; the numpy oracle port only if oracle/_ref did not travel (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "cfg2_s2vtatt_msrvtt"
DIMS = dict(B=128, N=40, V=2048, H=512, E=300, L=30, Vc=23000)
METRIC = "train videos/sec (fwd+bwd) S2VTAtt MSR-VTT shape"
CPU_SAMPLE_VIDEOS = 32          # numpy-port fallback only
NCU_TRAFFIC_FILE = os.path.join("profiles", "ncu_traffic.json")     # DRAM bytes per launch, from a committed ncu capture


def make_config(world, precision, dropout, workload=None):
    """The `config` object of the JSON line -- the same for both arms (the reference arm runs THIS workload)."""
    d = DIMS
    return dict(workload=workload or WORKLOAD, per_gpu_batch=d["B"], global_batch=d["B"] * world, parallelism="dp%d" % world,
                precision=precision, dropout_p=dropout,
                step="CUDA graph of one fwd+bwd (side lanes on; roofline pass times each kernel alone, lanes off)",
                l2="per-step working set (inputs 42 MB + fp32 weights 101 MB + activations > 1 GB) exceeds "
                   "the 126 MB L2; no explicit flush", **{k: v for k, v in d.items() if k != "B"})


def fwd_bwd_gflop(d, rationale=False):
    """Algorithmic GFLOP of one fwd+bwd step (cfg2 column, cfg3 with rationale=True)."""
    B, N, V, H, E, L, Vc = (d[k] for k in ("B", "N", "V", "H", "E", "L", "Vc"))
    enc_in = 2 * B * N * V * 3 * H
    fwd = (enc_in + 2 * B * N * H * 3 * H + 2 * B * N * H * H + 2 * B * L * H * H + 4 * B * L * N * H +
           2 * B * L * (H + E) * 3 * H + 2 * B * L * H * 3 * H + 2 * B * L * H * Vc)
    if rationale:      # + generator biLSTM in-proj / recurrent / linear; its in-proj needs no dX, the encoder's now does
        gen_in = 2 * B * N * V * 8 * H
        fwd += gen_in + 2 * 2 * B * N * H * 4 * H + 2 * B * N * 2 * H * 2
        return (3 * fwd - gen_in) / 1e9
    return (3 * fwd - enc_in) / 1e9        # the encoder input projection needs no dX (vid_feats has no grad)


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 10 ms during the timed region (a background
    thread in this process: `nvidia-smi -lms` in a child process perturbed the launch path measurably)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:            # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:             # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % getattr(self, "err", "")]}
        self._stop.set()
        self.thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def cpu_port_videos_per_sec(steps, warmup):
    """Fallback CPU arm (only when oracle/_ref is absent): oracle/captioning_oracle.py (numpy fp32, threaded BLAS),
    fwd + masked-CE loss + bwd on a bounded sample of the cfg2 workload."""
    import numpy as np
    from oracle import captioning_oracle as O
    from oracle import workloads as W
    d = DIMS
    p = W.s2vtatt_params(d["V"], d["H"], d["E"], d["Vc"], 123)
    vid, s, s_len = W.make_batch(CPU_SAMPLE_VIDEOS, d["N"], d["V"], d["L"], d["Vc"], 124)
    for _ in range(warmup):
        O.train_iter_s2vtatt(p, vid, s, s_len, d["Vc"] - 4, d["L"])
    t0 = time.perf_counter()
    for _ in range(steps):
        r = O.train_iter_s2vtatt(p, vid, s, s_len, d["Vc"] - 4, d["L"])
    dt = (time.perf_counter() - t0) / steps
    assert np.isfinite(r["loss"])
    return CPU_SAMPLE_VIDEOS / dt, dt


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count()


def reference_batch(device, dropout):
    """The reference S2VTAttModel (oracle/_ref: unmodified bytecode of /root/reference/model/S2VTAttModel.py) with the
    seeded cfg2 weights and one seeded cfg2 batch on `device`."""
    import torch
    from oracle import reference_runner as R
    from oracle import workloads as W
    d = DIMS
    dims = tuple(d[k] for k in ("B", "N", "V", "H", "E", "L", "Vc"))
    R.set_device(device)
    p = W.s2vtatt_params(d["V"], d["H"], d["E"], d["Vc"], 123)
    model = R.build_s2vtatt(dims, p, dropout_p=dropout, device=device).train()
    vid, s, s_len = W.make_batch(d["B"], d["N"], d["V"], d["L"], d["Vc"], 124)
    return model, tuple(torch.from_numpy(x).to(device) for x in (vid, s, s_len))


def cpu_reference_videos_per_sec(steps, warmup, dropout):
    """run_iter + loss.backward() of the reference (train.py:32-44,157-158) on the host cores, the WHOLE 128-video batch
    per step.  The thread count is set in this process: torchrun exports OMP_NUM_THREADS=1, which would otherwise cut
    the arm to one core.  -> (videos/s, s per step, threads)."""
    import torch
    from oracle import reference_runner as R
    threads = cpu_cores()
    torch.set_num_threads(threads)
    model, (vid, s, s_len) = reference_batch("cpu", dropout)
    for _ in range(warmup):
        R.run_iter(model, vid, s, s_len)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, _, _, _ = R.run_iter(model, vid, s, s_len)
    dt = (time.perf_counter() - t0) / steps
    assert torch.isfinite(loss).item()
    return DIMS["B"] / dt, dt, torch.get_num_threads()


def eager_b200_reference(dropout, warmup=3, iters=10):
    """The incumbent of SURVEY.md section 8(d): the unmodified reference modules in PyTorch eager on THIS GPU (cuDNN RNN,
    cuBLAS, ATen), fp32 and under torch.autocast(bfloat16); run_iter + backward, CUDA events."""
    import torch
    from oracle import reference_runner as R
    if not R.available():
        return {"unavailable": "oracle/_ref (reference bytecode) not present on this box"}
    out = {}
    model, (vid, s, s_len) = reference_batch("cuda", dropout)
    try:
        for name, ctx in (("fp32", None), ("autocast_bf16", torch.bfloat16)):
            def it():
                if ctx is None:
                    return R.run_iter(model, vid, s, s_len)[0]
                with torch.autocast("cuda", dtype=ctx):
                    return R.run_iter(model, vid, s, s_len)[0]
            for _ in range(warmup):
                it()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                loss = it()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            out[name] = {"value": DIMS["B"] / (ms / 1e3), "unit": "videos/s", "ms_per_step": ms, "iters": iters,
                         "loss": float(loss.item())}
        out["what"] = ("unmodified reference S2VTAttModel + train_utils.calc_masked_loss/accuracy + loss.backward() "
                       "(train.py:32-44,157-158) in PyTorch eager on this GPU, batch %d, dropout %.1f, tf32 off" % (DIMS["B"], dropout))
    finally:
        R.set_device("cpu")
        del model
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_runner as R
    steps, warmup = args.steps, max(args.warmup, 3)
    if R.available():
        vps, dt, threads = cpu_reference_videos_per_sec(steps, warmup, args.dropout)
        kind = "reference"
        sample = ("the whole %d-video %s batch per step: unmodified reference modules (oracle/_ref bytecode of "
                  "model/S2VTAttModel.py + train_utils.py), run_iter + loss.backward(), torch fp32 CPU, %d threads" % (
                      DIMS["B"], WORKLOAD, threads))
    else:
        steps, warmup = max(1, min(steps, 20)), max(1, min(warmup, 2))
        vps, dt = cpu_port_videos_per_sec(steps, warmup)
        kind, threads = "port", cpu_cores()
        sample = "%d of the %d videos of one %s batch per step (numpy fp32 oracle port, threaded BLAS; oracle/_ref absent)" % (
            CPU_SAMPLE_VIDEOS, DIMS["B"], WORKLOAD)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": vps, "unit": "videos/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": make_config(args.gpus, args.precision, args.dropout),
        "cpu_baseline": {"value": vps, "unit": "videos/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": vps, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def _finish_ranks(world):
    """End of a multi-rank run.  The step graph holds captured NCCL kernels; tearing the communicator down under it
    (destroy_process_group / interpreter exit) was observed to hang, so every rank synchronises and leaves directly."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def check_data_parallel_gradients(model, reducer, inputs, world, dev):
    """Correctness of the path SCALE times (NCCL all-reduce captured inside the step's CUDA graph), checked after the timed
    loop: with dropout off (deterministic step) the gradients a graph replay leaves in the buckets must equal, on every
    rank, the mean over ranks of the gradients the same rank computes locally without any reduction -- and be identical
    across ranks.  A dropped bucket, a missing 1/G or a bucket reduced before it was final all show up here."""
    import torch
    import torch.distributed as dist
    from pvcr_b200.graphs import GraphedTrainStep
    drops = [m for m in model.modules() if isinstance(m, torch.nn.Dropout)]
    p_prev = [m.p for m in drops]
    for m in drops:
        m.p = 0.0
    gen_net = getattr(model, "gen", None)          # RationaleNet: fix the Gumbel draws (Exp(1) noise) of the step as well
    if gen_net is not None:
        B_, N_ = inputs[0].shape[:2]
        gen_net.noise = torch.empty(B_ * N_, 2, device=dev).exponential_()
    try:
        step = GraphedTrainStep(model, inputs, warmup=0, reducer=reducer)
        step(*inputs)
        torch.cuda.synchronize()
        reduced = [b.clone() for b in reducer._flat]
        model.train_step_grads(*inputs)                    # local gradients, written in place into the same buckets
        torch.cuda.synchronize()
        worst, spread = 0.0, 0.0
        for red, loc in zip(reduced, reducer._flat):
            mean = loc.clone()
            dist.all_reduce(mean, op=dist.ReduceOp.SUM)
            mean /= world
            worst = max(worst, float(((red - mean).norm() / mean.norm().clamp_min(1e-30)).item()))
            lo, hi = red.clone(), red.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            spread = max(spread, float(((hi - lo).abs().max() / red.abs().max().clamp_min(1e-30)).item()))
        del step
        ok = worst < 1e-4 and spread == 0.0
        assert ok, "data-parallel gradient check failed: rel err %.3e vs mean of local gradients, rank spread %.3e" % (worst, spread)
        return {"ok": ok, "buckets": len(reduced), "rel_err_vs_mean_of_local_grads": worst, "max_spread_across_ranks": spread,
                "what": "in-graph NCCL step (dropout off) vs all-reduce-mean of per-rank eager gradients"}
    finally:
        for m, pv in zip(drops, p_prev):
            m.p = pv
        if gen_net is not None:
            gen_net.noise = None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pvcr_b200  # noqa: F401  (raises if the CUDA library is missing: there is no fallback)
    from pvcr_b200 import _lib
    from pvcr_b200.model import S2VTAttModel
    from pvcr_b200.parallel import GradAllReducer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the persistent cooperative kernels occupy 128 of the 148 SMs: keep NCCL within the remaining 20 so that the
        # overlapped gradient all-reduce can be co-resident instead of serialising in front of them
        # (per-communicator ncclConfig, not the NCCL_MAX_CTAS environment variable, which would also bind the
        # wider communicator used for the exposed tail of the gradient all-reduce)
        opts0 = dist.ProcessGroupNCCL.Options()
        opts0.config.max_ctas = args.nccl_ctas
        dist.init_process_group("nccl", device_id=dev, pg_options=opts0)
    d = DIMS
    B, N, V, H, E, L, Vc = (d[k] for k in ("B", "N", "V", "H", "E", "L", "Vc"))

    class Glove:                       # duck type of utils.GloveLoader (reference utils.py:52-66)
        word_vectors = None

        def get_id(self, w):
            return {"<sos>": Vc - 4, "<eos>": Vc - 3, "<pad>": Vc - 2, "<unk>": Vc - 1}[w]

    import numpy as np
    g = Glove()
    g.word_vectors = [np.zeros(E, np.float32)] * Vc
    torch.manual_seed(123)
    if args.workload == "cfg3":
        # BASELINE.json configs[2]: RationaleNet + S2VTAtt joint training (model/RationaleNet.py, train_rationale.py:30-44),
        # tau = 1, lambda_brev = lambda_cont = 1 (args.py:47-48); reported under profiles/, not the headline
        from pvcr_b200.model import RationaleNet
        wl_name = "cfg3_rationale_s2vtatt_msrvtt"
        model = RationaleNet(g, args.dropout, H, V, L, 1.0, "s2vt-att", precision=args.precision)
        emb = model.caption_net.decoder.embedding.weight
    else:
        wl_name = WORKLOAD
        model = S2VTAttModel(g, args.dropout, H, V, L, precision=args.precision)
        emb = model.decoder.embedding.weight
    with torch.no_grad():
        emb.normal_(0.0, 0.4)
    model = model.to(dev).train()
    tail_group = None
    if world > 1 and args.nccl_tail_ctas != args.nccl_ctas:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = args.nccl_tail_ctas       # communicator for the all-reduces nothing overlaps any more
        tail_group = dist.new_group(backend="nccl", pg_options=opts)
    reducer = GradAllReducer(model, flat=True, early=model.early_grad_params(), tail_group=tail_group)   # grads land in the buckets

    gen = torch.Generator().manual_seed(1000 + rank)
    vid_h = torch.randn(B, N, V, generator=gen)
    pad = torch.rand(B, generator=gen) < 0.25
    cut = torch.randint(N // 2, N, (B,), generator=gen)
    frame = torch.arange(N)[None, :]
    vid_h[(pad[:, None] & (frame >= cut[:, None]))] = 0.0
    s_len_h = torch.randint(1, L + 1, (B,), generator=gen)
    s_h = torch.randint(0, Vc - 4, (B, L), generator=gen)
    pos = torch.arange(L)[None, :]
    s_h[pos == (s_len_h[:, None] - 1)] = Vc - 3
    s_h[pos >= s_len_h[:, None]] = Vc - 2
    vid_h, s_h, s_len_h = vid_h.pin_memory(), s_h.pin_memory(), s_len_h.pin_memory()
    vid, s, s_len = vid_h.to(dev), s_h.to(dev), s_len_h.to(dev)

    from pvcr_b200.graphs import GraphedTrainStep
    L_ = _lib.lib()
    for _ in range(2):                      # eager warm-up (module load, attribute setup) before the capture
        model.train_step_grads(vid, s, s_len)
    L_.pvcr_prof_reset()
    # one fwd+bwd captured as CUDA graph(s); with N > 1 the vocabulary gradients' all-reduce overlaps the backward
    graphed = GraphedTrainStep(model, (vid, s, s_len), warmup=0, reducer=reducer if world > 1 else None)
    launches_per_step = sum(v[0] for v in _lib.prof_read().values())

    def step(v, t, tl):
        return graphed(v, t, tl)[0]         # copies inputs, replays the graph(s), all-reduces the gradients

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        loss = step(vid, s, s_len)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item(), "non-finite loss in warm-up"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: step(vid, s, s_len), args.steps)
    launches = launches_per_step * args.steps
    ms_step = ms_total / args.steps

    # end to end through the public API: pinned host inputs copied in, loss read back, every step
    # Every step copies ITS inputs from pinned host memory and reads ITS loss back; the copy of step k+1 is issued
    # before step k's loss is read (input prefetch on a copy stream), as a pinned-memory DataLoader would do.
    def e2e_step():
        graphed.step_prefetched()
        graphed.prefetch(vid_h, s_h, s_len_h)
        return graphed.static_out[0].item()

    graphed.prefetch(vid_h, s_h, s_len_h)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    h2d = vid_h.numel() * 4 + s_h.numel() * 8 + s_len_h.numel() * 8

    # per-kernel-class event timing (separate pass, not part of `value`): dominant class -> roofline
    # The side lanes are switched off for this pass so that every kernel is timed alone (in the measured step they
    # overlap: a GEMM sharing the SMs with a persistent sweep would be charged the sweep's duration).
    side_prev = L_.pvcr_side_mode(0)
    for _ in range(2):
        model.train_step_grads(vid, s, s_len)
    L_.pvcr_prof_reset()
    L_.pvcr_prof_enable(1)
    prof_steps = 2
    for _ in range(prof_steps):
        model.train_step_grads(vid, s, s_len)
    torch.cuda.synchronize()
    prof = _lib.prof_read()
    launch_list = _lib.prof_launch_list()
    L_.pvcr_prof_enable(0)
    L_.pvcr_side_mode(side_prev)

    dp_check = None
    if world > 1 and not os.environ.get("PVCR_DP_SKIP"):
        dp_check = check_data_parallel_gradients(model, reducer, (vid, s, s_len), world, dev)
    if rank != 0:
        _finish_ranks(world)
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
    planes = {"bf16": 1, "bf16x2": 3, "bf16x3": 6}[args.precision]
    classes = {k: {"launches_per_step": v[0] / prof_steps, "ms_per_step": v[1] / prof_steps} for k, v in prof.items()
               if v[0]}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"
    # per-kernel rooflines: algorithmic work per launch from SURVEY.md section 8(d)
    attn_fwd_step = 2 * B * N * H * 2 + B * H * 4 + B * N * 4 + B * H * 4          # proj_key + enc (bf16), q, alpha, ctx
    rec_step = B * 3 * H * 2 + B * H * 2                                           # gi in, h out (weights resident)
    alg_bytes = {
        "decoder_persistent_fwd": L * (attn_fwd_step + rec_step) + (4 * H * H + 3 * H * H) * 2,
        "decoder_persistent_bwd": L * (attn_fwd_step + B * H * 4 + 2 * B * N * H * 4 + 2 * rec_step) + 7 * H * H * 2,
        "gru_persistent_fwd": N * rec_step + 3 * H * H * 2,
        "gru_persistent_bwd": N * 2 * rec_step + 3 * H * H * 2,
    }
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a committed `ncu --set full` capture
    traffic, traffic_src = {}, None
    try:
        tj = json.load(open(os.path.join(ROOT, NCU_TRAFFIC_FILE)))
        traffic, traffic_src = tj.get("kernels", {}), tj.get("source")
    except (OSError, ValueError):
        pass
    # one entry per KERNEL LAUNCH of a step (launch i of step 0 averaged with launch i of step 1): every GEMM
    # instantiation is its own entry with its own executed FLOPs
    per_step = len(launch_list) // prof_steps if prof_steps and len(launch_list) % prof_steps == 0 else 0
    kernels, n_gemm = {}, 0
    for i in range(per_step):
        name = launch_list[i][0]
        ms = sum(launch_list[i + k * per_step][1] for k in range(prof_steps)) / prof_steps
        work = launch_list[i][2]
        if name == "gemm_tcgen05":
            ach = work / 1e12 / (ms / 1e3) if ms > 0 else 0.0
            kernels["gemm_tcgen05[%02d] %.2f GFLOP" % (n_gemm, work / 1e9)] = {
                "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                "ms_per_launch": ms, "executed_gflop": work / 1e9, "peak_source": peak_src, "traffic": None}
            n_gemm += 1
        elif name in alg_bytes:
            ach = alg_bytes[name] / 1e9 / (ms / 1e3)
            kernels[name] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                             "ms_per_launch": ms, "algorithmic_bytes_per_launch": alg_bytes[name],
                             "traffic": traffic.get(name), "traffic_source": traffic_src if name in traffic else None,
                             "peak_source": hbm_src}
    # `roofline` = the single kernel launch with the largest duration in a step
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"]) if kernels else None
    roofline = dict(kernels[dominant]) if dominant else {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s",
                                                         "frac": None, "traffic": None}
    roofline["kernel"] = dominant
    roofline["all_kernels"] = kernels
    roofline["class_ms_per_step"] = classes
    roofline["split_planes"] = planes
    gemm_ms = classes.get("gemm_tcgen05", {}).get("ms_per_step", 0.0)
    roofline["gemm_class"] = {"ms_per_step": gemm_ms, "algorithmic_gflop": fwd_bwd_gflop(d, args.workload == "cfg3"),
                              "executed_gflop": prof["gemm_tcgen05"][2] / prof_steps / 1e9 if "gemm_tcgen05" in prof else None,
                              "tflops_algorithmic": fwd_bwd_gflop(d, args.workload == "cfg3") / gemm_ms if gemm_ms else None,
                              "frac": fwd_bwd_gflop(d, args.workload == "cfg3") / gemm_ms / tensor_peak if gemm_ms else None}
    # the number that matters for the whole step: algorithmic FLOPs / measured step time / sustained tensor peak
    roofline["step_frac"] = fwd_bwd_gflop(d, args.workload == "cfg3") / ms_step / tensor_peak
    roofline["step_tflops"] = fwd_bwd_gflop(d, args.workload == "cfg3") / ms_step

    # second half of BASELINE.json's metric: greedy captions/sec (eval branch, model/S2VTAttModel.py:172-191) at the
    # same per-GPU batch, fp32-equivalent bf16x3 arithmetic (token ids bit-exact vs the fp32 reference), CUDA-graph
    # replay with the features resident; single GPU only (decoding does not communicate: N GPUs are N replicas)
    greedy = None
    if world == 1 and not args.no_greedy and args.workload == "cfg2":
        from pvcr_b200.graphs import GraphedGreedy
        model.eval()
        gg = GraphedGreedy(model, vid)
        for _ in range(3):
            gg(vid)
        g_iters = 10
        ms_g = timed(lambda: gg(vid), g_iters) / g_iters
        # one decoding step streams the bf16 weights once, whatever the batch: W_v + decoder W_ih / W_hh / W_q
        w_step = (Vc * H + 3 * H * (H + E) + 3 * H * H + H * H) * 2
        g_bytes = L * w_step + (3 * H * V + 3 * H * H + H * H) * 2          # + encoder / key weights once per batch
        g_ach = g_bytes / 1e9 / (ms_g / 1e3)
        greedy = {"value": B / (ms_g / 1e3), "unit": "captions/s", "batch": B, "ms_per_batch": ms_g, "max_len": L,
                  "arithmetic": "bf16x3 (fp32-equivalent; ids bit-exact vs the reference)", "step": "CUDA graph",
                  "roofline": {"bound": "hbm", "achieved": g_ach, "peak": hbm_peak, "unit": "GB/s", "frac": g_ach / hbm_peak,
                               "algorithmic_bytes_per_batch": g_bytes, "traffic": None,
                               "note": "whole decode of one batch (40 encoder + 30 decoder steps, many launches) vs the "
                                       "bf16 weight bytes SURVEY 8(d) counts; B-independent"}}
        # captioning needs the ids only: the same decode without materialising the [B, L, Vc] fp32 logits
        gi = GraphedGreedy(model, vid, return_logits=False)
        for _ in range(3):
            gi(vid)
        ms_i = timed(lambda: gi(vid), g_iters) / g_iters
        greedy["ids_only"] = {"value": B / (ms_i / 1e3), "unit": "captions/s", "ms_per_batch": ms_i}
        greedy["prepared_weights"] = "weight planes and the word table W_e Emb[w] + b_ih staged once, outside the timed region"
        model.train()
        del gg, gi

    # step-inclusive variant (reported separately, not the headline): the same step with the fused
    # clip_grad_norm_ + Adam kernels (train.py:157-160) captured behind the backward in the same CUDA graph
    with_opt = None
    if world == 1 and not args.no_optimizer:
        from pvcr_b200.optim import FusedClipAdam
        model.train()
        opt = FusedClipAdam(model.parameters(), lr=2e-3, weight_decay=4e-5, max_norm=1.0)     # args.py:41-45 defaults
        g2 = GraphedTrainStep(model, (vid, s, s_len), warmup=0, optimizer=opt)
        for _ in range(3):
            g2(vid, s, s_len)
        ms_o = timed(lambda: g2(vid, s, s_len), args.steps) / args.steps
        with_opt = {"value": B / (ms_o / 1e3), "unit": "videos/s", "ms_per_step": ms_o,
                    "step": "fwd + bwd + clip_grad_norm_ + Adam in one CUDA graph", "final_loss": float(g2.static_out[0].item())}
        del g2, opt

    # the mode that reproduces the reference to fp32 level on every gradient at the full size (tests/
    # test_gpu_fullsize_reference.py): each fp32 operand split into two bf16 terms, 3 tcgen05 products per logical product.
    # It runs on the step-wise kernels (the persistent sweeps hold single-plane weights).  Reported beside the headline.
    x2 = None
    if world == 1 and not args.no_bf16x2 and args.workload == "cfg2" and args.precision == "bf16":
        m2 = S2VTAttModel(g, args.dropout, H, V, L, precision="bf16x2").to(dev).train()
        m2.load_state_dict(model.state_dict())
        g3 = GraphedTrainStep(m2, (vid, s, s_len), warmup=1)
        for _ in range(3):
            g3(vid, s, s_len)
        k2 = max(3, args.steps // 4)
        ms_2 = timed(lambda: g3(vid, s, s_len), k2) / k2
        x2 = {"value": B / (ms_2 / 1e3), "unit": "videos/s", "ms_per_step": ms_2, "steps": k2,
              "what": "same step in bf16x2 arithmetic (gradients within 3e-5 of the float64 reference at this size)"}
        del g3, m2
        torch.cuda.empty_cache()

    cpu = None
    eager = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "cfg2":
        from oracle import reference_runner as R
        if R.available():
            vps, dt, threads = cpu_reference_videos_per_sec(2, 1, args.dropout)
            cpu = {"value": vps, "unit": "videos/s", "cores": threads, "kind": "reference",
                   "sample": "2 steps (after 1 warm-up) of the whole %d-video batch: unmodified reference modules "
                             "(oracle/_ref), run_iter + loss.backward(), torch fp32 CPU, %d threads" % (B, threads)}
        else:
            vps, dt = cpu_port_videos_per_sec(3, 1)
            cpu = {"value": vps, "unit": "videos/s", "cores": cpu_cores(), "kind": "port",
                   "sample": "%d of the %d videos of one batch per step, 3 steps (numpy fp32 oracle port, threaded BLAS; "
                             "oracle/_ref absent)" % (CPU_SAMPLE_VIDEOS, B)}
    if world == 1 and not args.no_eager and args.workload == "cfg2":
        del graphed
        torch.cuda.empty_cache()
        eager = eager_b200_reference(args.dropout)

    # BASELINE.json configs[3]: SpatialNet fwd + masked loss + bwd on synthetic grid features (B x 40 frames x 2048
    # channels x 6 x 6 cells), the whole step as one CUDA graph with the autograd tape inside (GraphedAutogradStep).  An extra
    # key beside the headline; its parity is tests/test_gpu_boundary.py and tests/test_gpu_spatial_front.py.
    spatial = None
    if world == 1 and not args.no_spatial and args.workload == "cfg2":
        try:
            torch.cuda.empty_cache()
            from pvcr_b200 import train_utils as TU
            from pvcr_b200.graphs import GraphedAutogradStep
            from pvcr_b200.model import SpatialNet
            Fs, Ks = 2048, 6
            net = SpatialNet(g, args.dropout, H, Fs, L, "s2vt-att", precision=args.precision).to(dev).train()
            vid4 = torch.randn(B, N, Fs, Ks, Ks, device=dev)
            crit = torch.nn.CrossEntropyLoss(reduction="none")
            gs4 = GraphedAutogradStep(net, lambda: TU.calc_masked_loss(net(vid4, s)[0], s, s_len, crit))
            for _ in range(2):
                gs4.replay()
            k4 = max(3, args.steps // 4)
            ms_4 = timed(gs4.replay, k4) / k4
            conv_tflop = 2 * B * N * Ks * Ks * 9 * (Fs * H + H * H) * 3 / 1e12 - 2 * B * N * Ks * Ks * 9 * Fs * H / 1e12
            spatial = {"value": B / (ms_4 / 1e3), "unit": "videos/s", "ms_per_step": ms_4, "steps": k4,
                       "loss": float(gs4.static_loss.item()), "conv_useful_tflop_per_step": conv_tflop,
                       "workload": "cfg4_spatialnet: B=%d x %d frames x %d channels x %dx%d cells, H=%d, L=%d, Vc=%d, dropout %.1f, %s; "
                                   "fwd + calc_masked_loss + bwd, one CUDA graph, input resident (1.5 GB: larger than L2)"
                                   % (B, N, Fs, Ks, Ks, H, L, Vc, args.dropout, args.precision)}
            del gs4, net, vid4
            torch.cuda.empty_cache()
        except Exception as e:            # noqa: BLE001  (an extra key must not take the headline line down)
            spatial = {"error": repr(e)[:300]}

    out = {
        "metric": METRIC, "value": B * world / (ms_step / 1e3), "unit": "videos/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else args.precision,
        "data": "synthetic",
        "config": make_config(world, args.precision, args.dropout, wl_name),
        "e2e": {"value": B * world / (ms_e2e / 1e3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "greedy": greedy, "with_optimizer": with_opt,
        "reference_eager_b200": eager, "dp_check": dp_check, "precision_bf16x2": x2,
        "spatialnet_cfg4": spatial,
    }
    print(json.dumps(out), flush=True)
    _finish_ranks(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x2", "bf16x3"])
    ap.add_argument("--dropout", type=float, default=0.2, help="reference default dropout_p (args.py:26)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-greedy", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true")
    ap.add_argument("--no-bf16x2", action="store_true", help="skip timing the same step in bf16x2 arithmetic")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"],
                    help="cfg2 = S2VTAtt (BASELINE.json's metric config, default); cfg3 = RationaleNet + S2VTAtt joint training")
    ap.add_argument("--no-eager", action="store_true", help="skip timing the reference modules in PyTorch eager on the GPU")
    ap.add_argument("--no-spatial", action="store_true", help="skip timing SpatialNet (BASELINE.json configs[3])")
    ap.add_argument("--nccl-ctas", type=int, default=16)
    ap.add_argument("--nccl-tail-ctas", type=int, default=64)
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the OpenMP / MKL runtimes read their environment when torch is imported
        os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = str(cpu_cores())
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
