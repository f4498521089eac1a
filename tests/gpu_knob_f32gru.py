"""Helper of test_gpu_decode.py::test_stepwise_encoder_fallback_agrees: greedy ids with the fp32 persistent encoder switched
off (PVCR_NO_F32_GRU=1, read once per process) -- printed as a checksum for comparison with the default path."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import workloads as W
from pvcr_b200.model import S2VTAttModel
from tests.gpu_util import FixtureGlove, to_cuda

B, N, V, H, E, L, Vc = 33, 12, 96, 128, 40, 6, 300
p = W.s2vtatt_params(V, H, E, Vc, 333)
vid, _, _ = W.make_batch(B, N, V, L, Vc, 433)
m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), p).eval()
ids, logits = m.greedy(torch.from_numpy(vid).cuda())
print("IDS", hashlib.sha1(ids.cpu().numpy().tobytes()).hexdigest(), "%.9e" % float(logits.double().norm()))
