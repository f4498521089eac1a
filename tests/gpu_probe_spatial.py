"""cfg4 (BASELINE.json configs[3]): SpatialNet fwd + masked loss + bwd on synthetic grid features (B x 40 frames x 2048 channels x
6 x 6 cells, 30 tokens, 23k vocabulary) -- ours vs the unmodified reference modules in PyTorch eager on the same GPU.
    python tests/gpu_probe_spatial.py [B]        -> one JSON line (written to profiles/ by hand)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pvcr_b200  # noqa: F401
from pvcr_b200 import _lib
from pvcr_b200 import train_utils as TU
from pvcr_b200.model import SpatialNet
from tests.gpu_util import FixtureGlove

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, F, K, H, E, L, Vc = 40, 2048, 6, 512, 300, 30, 23000
torch.manual_seed(5)
vid = torch.randn(B, N, F, K, K, device="cuda")
s_len = torch.randint(1, L + 1, (B,), device="cuda")
s = torch.randint(0, Vc - 4, (B, L), device="cuda")


def timed(fn, it=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


out = {"workload": "cfg4_spatialnet", "B": B, "N": N, "F": F, "K": K, "H": H, "L": L, "Vc": Vc}
net = SpatialNet(FixtureGlove(Vc, E), 0.2, H, F, L, "s2vt-att").cuda().train()
crit = torch.nn.CrossEntropyLoss(reduction="none")


def ours():
    net.zero_grad(set_to_none=True)
    logits, _ = net(vid, s)
    loss = TU.calc_masked_loss(logits, s, s_len, crit)
    loss.backward()
    return loss


L_ = _lib.lib()
if os.environ.get("PVCR_PROBE_NCU"):
    # exactly one steady-state eager step between cudaProfilerStart / Stop: ncu --profile-from-start off ... (profiles/r02t_cfg4_launches.csv)
    ours(); ours(); torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ours(); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
ms = timed(ours)
L_.pvcr_prof_reset(); L_.pvcr_prof_enable(1)
ours(); torch.cuda.synchronize()
prof = _lib.prof_read(); launches = _lib.prof_launch_list(); L_.pvcr_prof_enable(0)
conv_flops_useful = 2 * B * N * K * K * 9 * (F * H + H * H) * 3 - 2 * B * N * K * K * 9 * F * H      # fwd + dW + dX (no dX for layer 1)
out["ours"] = {"ms_per_step": ms, "videos_per_s": B / (ms / 1e3), "loss": float(ours().item()),
               "class_ms": {k: round(v[1], 3) for k, v in prof.items() if v[0]},
               "gemm_executed_tflop": prof["gemm_tcgen05"][2] / 1e12, "conv_useful_tflop": conv_flops_useful / 1e12}
out["ours"]["launches"] = len(launches)
out["ours"]["misc_launch_ms"] = [round(ms, 3) for cls, ms, _ in launches if cls == "misc" and ms > 0.05]
out["ours"]["staging_launch_ms"] = [round(ms, 3) for cls, ms, _ in launches if cls == "operand_staging" and ms > 0.05]
out["ours"]["gemm_launch_ms"] = [round(ms, 3) for cls, ms, _ in launches if cls == "gemm_tcgen05" and ms > 0.2]
if os.environ.get("PVCR_PROBE_GRAPH"):
    # host-side time of one eager step (launch-bound?) and the same step captured as ONE CUDA graph (autograd tape inside the capture)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ours()
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()
    out["ours"]["host_issue_ms_per_step"] = host_ms
    try:
        from pvcr_b200.graphs import GraphedAutogradStep
        gs = GraphedAutogradStep(net, lambda: TU.calc_masked_loss(net(vid, s)[0], s, s_len, crit))
        ms_g = timed(gs.replay)
        out["ours_graph"] = {"ms_per_step": ms_g, "videos_per_s": B / (ms_g / 1e3), "loss": float(gs.replay().item())}
        del gs
    except Exception as e:      # noqa: BLE001
        out["ours_graph_error"] = repr(e)[:400]
torch.cuda.empty_cache()
try:
    from oracle import reference_runner as R
    if R.available() and os.environ.get("PVCR_PROBE_NO_REF") is None:
        R.set_device("cuda")
        ref = R.modules()["model.SpatialNet"].SpatialNet(R.FakeGlove(Vc, E), 0.2, H, F, L, "s2vt-att").cuda().train()
        tu = R.modules()["train_utils"]

        def theirs():
            ref.zero_grad(set_to_none=True)
            logits, _ = ref(vid, s)
            loss = tu.calc_masked_loss(logits, s, s_len, crit)
            loss.backward()
            return loss
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        ms_r = timed(theirs)
        out["reference_eager_fp32"] = {"ms_per_step": ms_r, "videos_per_s": B / (ms_r / 1e3)}
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_t = timed(theirs)
        out["reference_eager_tf32"] = {"ms_per_step": ms_t, "videos_per_s": B / (ms_t / 1e3)}

        def theirs_bf16():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return theirs()
        ms_b = timed(theirs_bf16)
        out["reference_eager_autocast_bf16"] = {"ms_per_step": ms_b, "videos_per_s": B / (ms_b / 1e3)}
except Exception as e:          # noqa: BLE001
    out["reference_error"] = repr(e)[:300]
print(json.dumps(out))
