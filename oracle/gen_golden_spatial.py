"""Golden vectors of the reference SpatialNet (model/SpatialNet.py:55-142) driving a caption net through
`encode_step` / `decode` -- the boundary of SURVEY.md section 8(b).  TEST INFRASTRUCTURE; runs where /root/reference exists.

    python oracle/gen_golden_spatial.py      # -> tests/golden/spatial_{att,s2vt}_tiny.npz

Float64 run of the unmodified classes (dropout 0; BatchNorm in training mode, i.e. batch statistics): inputs, the whole
state_dict, logits, seq_alphas, loss and every parameter gradient.  The GPU test re-runs SpatialNet's frame loop with
the drop-in caption net behind `encode_step` / `decode` and must reproduce these numbers.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import reference_runner as R      # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def case(tag, arch, dims, K, seed):
    B, N, F, H, E, L, Vc = dims
    torch.manual_seed(seed)
    rs = np.random.RandomState(seed)
    vectors = [rs.randn(E).astype(np.float32) * 0.5 for _ in range(Vc)]
    SpatialNet = R.modules()["model.SpatialNet"].SpatialNet
    tu = R.modules()["train_utils"]
    model = SpatialNet(R.FakeGlove(Vc, E, vectors), 0.0, H, F, L, arch).double()
    model.load_state_dict({k: v.float().double() for k, v in model.state_dict().items()})     # float32-exact weights
    vid = rs.randn(B, N, F, K, K).astype(np.float32)
    s_len = rs.randint(1, L + 1, size=B).astype(np.int64)
    s_len[0] = L
    s = np.full((B, L), Vc - 2, np.int64)
    for b in range(B):
        s[b, :s_len[b] - 1] = rs.randint(0, Vc - 4, size=s_len[b] - 1)
        s[b, s_len[b] - 1] = Vc - 3
    tv, ts, tl = torch.from_numpy(vid).double(), torch.from_numpy(s), torch.from_numpy(s_len)
    model.train()
    sd0 = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}      # before BN updates its running stats
    logits, seq_alphas = model(tv, ts)                     # train_spatial.py:32
    loss = tu.calc_masked_loss(logits, ts, tl, torch.nn.CrossEntropyLoss(reduction="none"))
    acc = tu.calc_masked_accuracy(logits, ts, tl)
    loss.backward()
    out = dict(dims=np.array(dims), K=K, arch=arch, vid=vid, s=s, s_len=s_len, loss=loss.item(), acc=acc.item(),
               logits=logits.detach().numpy(), seq_alphas=seq_alphas.detach().numpy())
    for k, v in sd0.items():
        out["param." + k] = v
    for k, prm in model.named_parameters():
        out["grad." + k] = prm.grad.detach().numpy()
    model.eval()
    with torch.no_grad():
        lg, al = model(tv, None)
    out["eval_logits"] = lg.numpy()
    out["eval_ids"] = torch.argmax(lg, dim=2).numpy()
    out["eval_seq_alphas"] = al.numpy()
    path = os.path.join(OUT, tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.0f KB) loss %.6f" % (path, os.path.getsize(path) / 1024, out["loss"]))


if __name__ == "__main__":
    R.set_device("cpu")
    torch.set_default_dtype(torch.float64)     # SpatialNet.py:115 and S2VTModel.py:103,111 create default-dtype zeros
    # dims = (B, N, F = vid_feat_size, H, E, L, Vc); K = grid size
    case("spatial_att_tiny", "s2vt-att", (3, 4, 24, 32, 16, 5, 40), 3, 21)
    case("spatial_s2vt_tiny", "s2vt", (3, 4, 24, 32, 16, 5, 40), 3, 22)
