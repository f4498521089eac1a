"""Seeded synthetic workloads and reference-initialised weights for the oracle (TEST INFRASTRUCTURE).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
Input conventions follow BASELINE.md section 2 (dataset.py:65-84, utils.py:42-50): vid_feats ~ N(0,1) with
zero-padded tail frames on 25 % of the videos, s_len ~ U{1..L}, <eos> at s_len-1, <pad> after, the four
special tokens being the last four vocabulary ids.  Weights follow torch's default GRU/LSTM/Linear
initialisation U(-1/sqrt(H), 1/sqrt(H)) which S2VTAttModel and the Generator keep.
"""
import numpy as np

CONFIGS = {
    # name: (B, N, V, H, E, L, Vc)
    "cfg1_s2vt_msvd": (32, 80, 4096, 512, 300, 28, 10000),
    "cfg2_s2vtatt_msrvtt": (128, 40, 2048, 512, 300, 30, 23000),
}


def make_batch(B, N, V, L, Vc, seed, dtype=np.float32):
    rs = np.random.RandomState(seed)
    vid = rs.standard_normal((B, N, V)).astype(dtype)
    for b in range(B):
        if rs.rand() < 0.25:
            vid[b, N - rs.randint(1, max(2, N // 2)):] = 0.0
    s_len = rs.randint(1, L + 1, size=B).astype(np.int64)
    s = np.full((B, L), Vc - 2, np.int64)
    for b in range(B):
        s[b, :s_len[b] - 1] = rs.randint(0, Vc - 4, size=s_len[b] - 1)
        s[b, s_len[b] - 1] = Vc - 3
    return vid, s, s_len


def _u(rs, shape, k, dtype):
    return rs.uniform(-k, k, size=shape).astype(dtype)


def s2vtatt_params(V, H, E, Vc, seed, dtype=np.float32):
    rs = np.random.RandomState(seed)
    k = 1.0 / np.sqrt(H)
    p = {}
    for name, inp in (("encoder.rnn", V), ("decoder.rnn", H + E)):
        p[name + ".weight_ih_l0"] = _u(rs, (3 * H, inp), k, dtype)
        p[name + ".weight_hh_l0"] = _u(rs, (3 * H, H), k, dtype)
        p[name + ".bias_ih_l0"] = _u(rs, (3 * H,), k, dtype)
        p[name + ".bias_hh_l0"] = _u(rs, (3 * H,), k, dtype)
    p["decoder.embedding.weight"] = (rs.standard_normal((Vc, E)) * 0.4).astype(dtype)
    for name in ("key_layer", "query_layer"):
        p["decoder.attention.%s.weight" % name] = _u(rs, (H, H), k, dtype)
    p["decoder.attention.energy_layer.weight"] = _u(rs, (1, H), k, dtype)
    p["decoder.pred_linear.1.weight"] = _u(rs, (Vc, H), k, dtype)
    p["decoder.pred_linear.1.bias"] = _u(rs, (Vc,), k, dtype)
    return p


def s2vt_params(V, H, E, Vc, seed, dtype=np.float32):
    """S2VTModel: Xavier-normal weights, bias 0.01 (utils.py:100-118 ixvr, applied by S2VTModel.__init__)."""
    rs = np.random.RandomState(seed)

    def xav(shape):
        return (rs.standard_normal(shape) * np.sqrt(2.0 / (shape[0] + shape[1]))).astype(dtype)

    p = {"embedding.0.weight": (rs.standard_normal((Vc, E)) * 0.4).astype(dtype)}
    for name, inp in (("rnn1", V), ("rnn2", H + E)):
        p[name + ".weight_ih_l0"] = xav((3 * H, inp))
        p[name + ".weight_hh_l0"] = xav((3 * H, H))
        p[name + ".bias_ih_l0"] = np.full((3 * H,), 0.01, dtype)
        p[name + ".bias_hh_l0"] = np.full((3 * H,), 0.01, dtype)
    p["linear.1.weight"] = xav((Vc, H))
    p["linear.1.bias"] = np.full((Vc,), 0.01, dtype)
    return p


def generator_params(V, H, seed, dtype=np.float32):
    rs = np.random.RandomState(seed)
    k = 1.0 / np.sqrt(H)
    p = {}
    for sfx in ("", "_reverse"):
        p["rnn.weight_ih_l0" + sfx] = _u(rs, (4 * H, V), k, dtype)
        p["rnn.weight_hh_l0" + sfx] = _u(rs, (4 * H, H), k, dtype)
        p["rnn.bias_ih_l0" + sfx] = _u(rs, (4 * H,), k, dtype)
        p["rnn.bias_hh_l0" + sfx] = _u(rs, (4 * H,), k, dtype)
    k2 = 1.0 / np.sqrt(2 * H)
    p["linear.weight"] = _u(rs, (2, 2 * H), k2, dtype)
    p["linear.bias"] = _u(rs, (2,), k2, dtype)
    return p
