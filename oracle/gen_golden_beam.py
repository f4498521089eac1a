"""Golden vectors for beam search (test infrastructure; runs only where /root/reference is mounted).

The reference has no beam search, so there is nothing of its own to compare with; what CAN be pinned is that the
search defined in oracle/captioning_oracle.py (s2vtatt_beam_search) is carried out over the reference's own modules:
this script drives the UNMODIFIED reference Encoder / Attention.key_layer / Decoder.forward_step
(model/S2VTAttModel.py:80-96,125-148,178) in float64 with the same fixed-length search and stores ids and scores for
beams 1, 3 and 5 on the weights and videos of the existing S2VTAtt fixtures; beam 1 is additionally checked here against
the reference's own greedy eval branch.

    python oracle/gen_golden_beam.py        # writes tests/golden/s2vtatt_beam_{tiny,mid}.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
REF = os.environ.get("PVCR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from model.S2VTAttModel import S2VTAttModel          # noqa: E402  (the reference)
from oracle.gen_golden import FakeGlove              # noqa: E402
from tests.golden_util import load                   # noqa: E402


def beam_search_reference(model, vid, K, L, sos_id):
    B = vid.shape[0]
    enc, enc_final = model.encoder(vid)                                   # [B,N,H], [1,B,H]
    pk = model.decoder.attention.key_layer(enc)
    encr, pkr = enc.repeat_interleave(K, 0), pk.repeat_interleave(K, 0)
    state = enc_final.repeat_interleave(K, 1)
    w = torch.full((B * K,), sos_id, dtype=torch.long)
    score = torch.zeros(B, K, dtype=torch.float64)
    ids = torch.zeros(B, K, L, dtype=torch.long)
    for i in range(L):
        logits, state = model.decoder.forward_step(encr, pkr, state, w)
        Vc = logits.shape[1]
        cand = score[:, :, None] + torch.log_softmax(logits, 1).reshape(B, K, Vc)
        if i == 0:
            cand[:, 1:, :] = -float("inf")
        flat = cand.reshape(B, K * Vc)
        order = torch.from_numpy(np.argsort(-flat.numpy(), axis=1, kind="stable")[:, :K].copy())
        score = flat.gather(1, order)
        parent, word = order // Vc, order % Vc
        rows = (torch.arange(B)[:, None] * K + parent).reshape(-1)
        state = state[:, rows]
        ids = ids.gather(1, parent[:, :, None].expand(B, K, L))
        ids[:, :, i] = word
        w = word.reshape(-1)
    return ids.numpy(), score.numpy()


def main():
    torch.set_default_dtype(torch.float64)
    for tag in ("s2vtatt_tiny", "s2vtatt_mid"):
        d, params, _ = load(tag)
        B, N, V, H, E, L, Vc = (int(x) for x in d["dims"])
        model = S2VTAttModel(FakeGlove(Vc, E, 0), 0.0, H, V, L).double()
        model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        model.eval()
        vid = torch.from_numpy(d["vid"].astype(np.float64))
        out = {}
        with torch.no_grad():
            greedy = torch.argmax(model(vid), 2).numpy()                  # the reference's own eval branch
            for K in (1, 3, 5):
                ids, score = beam_search_reference(model, vid, K, L, int(d["sos_id"]))
                out["ids_k%d" % K], out["score_k%d" % K] = ids, score
        assert np.array_equal(out["ids_k1"][:, 0], greedy), "beam 1 must be the reference's greedy decoding"
        assert np.array_equal(greedy, d["greedy_ids"])
        path = os.path.join(HERE, "..", "tests", "golden", tag.replace("s2vtatt_", "s2vtatt_beam_") + ".npz")
        np.savez_compressed(path, **out)
        print(tag, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
