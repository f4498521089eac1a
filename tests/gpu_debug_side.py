import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tests.test_gpu_s2vtatt import *
from pvcr_b200.graphs import GraphedTrainStep
from pvcr_b200.model import S2VTAttModel
d, params, g, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_mid")
print(B,N,V,H,E,L,Vc)
vid = torch.from_numpy(d["vid"]).cuda(); s = torch.from_numpy(d["s"]).cuda(); s_len = torch.from_numpy(d["s_len"]).cuda()
m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision="bf16"), params).train()
refs=[]
for i in range(3):
    ref_loss, _, _ = m.train_step_grads(vid, s, s_len)
    refs.append({k: v.copy() for k, v in grads_of(m).items()})
for k in refs[0]:
    print("eager-eager %-45s %.2e %.2e" % (k, relerr(refs[1][k], refs[0][k]), relerr(refs[2][k], refs[0][k])))
step = GraphedTrainStep(m, (vid, s, s_len))
for i in range(2):
    loss, acc, pred = step(vid, s, s_len)
    torch.cuda.synchronize()
    got = grads_of(m)
    for k in refs[0]:
        e = relerr(got[k], refs[0][k])
        if e > 1e-6: print("graph-eager %d %-45s %.2e" % (i, k, e))
k="decoder.embedding.weight"
print("eager vs golden", relerr(refs[0][k], g[k]), "graph vs golden", relerr(got[k], g[k]))
