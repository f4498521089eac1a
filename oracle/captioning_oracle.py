"""CPU oracle for the captioning hot path (TEST INFRASTRUCTURE — not product code).

A numpy restatement of the reference's S2VT / S2VTAtt / RationaleNet forward pass, the
train.py / train_rationale.py loss contract, and hand-derived backward passes.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product path (``pvcr_b200``) never does.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference modules themselves, executed in the authoring
container by ``oracle/gen_golden.py`` (float64, torch autograd) and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below against them.

Reference lines restated (all relative to the reference repo root):
  model/S2VTAttModel.py:25-48   Attention.forward            -> attention_fwd / attention_bwd
  model/S2VTAttModel.py:80-96   Encoder.forward              -> gru_seq_fwd / gru_seq_bwd
  model/S2VTAttModel.py:125-196 Decoder.forward_step/forward -> s2vtatt_*
  model/S2VTModel.py:74-177     encode / decode              -> s2vt_*
  model/RationaleNet.py:32-54   Generator.forward            -> generator_fwd / generator_bwd
  train_utils.py:22-95          masked loss / acc / penalties-> masked_loss, masked_accuracy,
                                                                brevity_loss, cont_loss
Parameters are passed as dicts keyed by the reference ``state_dict`` names.
All functions are dtype-generic (float32 or float64 numpy arrays).
"""
import numpy as np


# ----------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def log_softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    y = x - m
    return y - np.log(np.exp(y).sum(axis=axis, keepdims=True))


# ----------------------------------------------------------------------------------------
# operand-rounding emulation (tests only): models "bf16 tensor-core operands, wide accumulation"
# ----------------------------------------------------------------------------------------
_Q = None          # rounding applied to every GEMM operand (None = exact arithmetic, the reference's semantics)
_QK = None         # rounding of the cached proj_key as the attention kernels hold it (fp16)


def bf16_round(x):
    """Round-to-nearest-even to bfloat16 precision (8 significand bits), returned in the input dtype."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(x.dtype).reshape(x.shape)


def fp16_round(x):
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(x.dtype)


def set_operand_rounding(q=None, qk=None):
    """q: rounding of every matrix-product operand; qk: rounding of proj_key inside the attention score.
    set_operand_rounding(bf16_round, fp16_round) emulates the CUDA path's bf16 mode; () restores exact arithmetic."""
    global _Q, _QK
    _Q, _QK = q, qk


def _mm(a, b):
    """a @ b with both operands passed through the operand rounding (if any)."""
    return a @ b if _Q is None else _Q(a) @ _Q(b)


def _q(x):
    return x if _Q is None else _Q(x)


def _sub(params, prefix):
    """View of a parameter dict under ``prefix`` (e.g. 'caption_net.')."""
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


# ----------------------------------------------------------------------------------------
# GRU (torch.nn.GRU, 1 layer, seq-first; gate order r,z,n)  S2VTAttModel.py:60-61,116-117
# ----------------------------------------------------------------------------------------
def gru_seq_fwd(gi, w_hh, b_hh, h0):
    """gi: [T,B,3H] precomputed W_ih x + b_ih.  Returns hs [T,B,H] and a cache."""
    T, B, H3 = gi.shape
    H = H3 // 3
    hs = np.empty((T, B, H), gi.dtype)
    r_ = np.empty_like(hs); z_ = np.empty_like(hs); n_ = np.empty_like(hs); ghn_ = np.empty_like(hs)
    hprev = np.empty_like(hs)
    h = h0
    for t in range(T):
        gh = _mm(h, w_hh.T) + b_hh
        r = sigmoid(gi[t, :, :H] + gh[:, :H])
        z = sigmoid(gi[t, :, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[t, :, 2 * H:] + r * gh[:, 2 * H:])
        hprev[t] = h
        h = (1.0 - z) * n + z * h
        hs[t] = h; r_[t] = r; z_[t] = z; n_[t] = n; ghn_[t] = gh[:, 2 * H:]
    return hs, dict(r=r_, z=z_, n=n_, ghn=ghn_, hprev=hprev, w_hh=w_hh)


def gru_seq_bwd(dhs, dh_last, cache):
    """dhs: [T,B,H] gradient w.r.t. every output h_t; dh_last: extra gradient on h_{T-1}
    (the returned state).  Returns dgi [T,B,3H], dw_hh, db_hh, dh0."""
    r_, z_, n_, ghn_, hprev, w_hh = (cache[k] for k in ("r", "z", "n", "ghn", "hprev", "w_hh"))
    T, B, H = r_.shape
    dgi = np.empty((T, B, 3 * H), r_.dtype)
    dgh_all = np.empty((T, B, 3 * H), r_.dtype)
    dh = np.zeros((B, H), r_.dtype) if dh_last is None else dh_last.copy()
    for t in range(T - 1, -1, -1):
        dh = dh + dhs[t]
        r, z, n, ghn, hp = r_[t], z_[t], n_[t], ghn_[t], hprev[t]
        dn = dh * (1.0 - z)
        dz = dh * (hp - n)
        dnp = dn * (1.0 - n * n)
        dzp = dz * z * (1.0 - z)
        drp = dnp * ghn * r * (1.0 - r)
        dgi[t, :, :H] = drp; dgi[t, :, H:2 * H] = dzp; dgi[t, :, 2 * H:] = dnp
        dgh_all[t, :, :H] = drp; dgh_all[t, :, H:2 * H] = dzp; dgh_all[t, :, 2 * H:] = dnp * r
        dh = dh * z + _mm(dgh_all[t], w_hh)
    dw_hh = _mm(dgh_all.reshape(T * B, 3 * H).T, hprev.reshape(T * B, H))
    db_hh = dgh_all.sum(axis=(0, 1))
    return dgi, dw_hh, db_hh, dh


# ----------------------------------------------------------------------------------------
# LSTM (torch.nn.LSTM, 1 layer; gate order i,f,g,o)  RationaleNet.py:26-27
# ----------------------------------------------------------------------------------------
def lstm_seq_fwd(gi, w_hh, b_hh):
    """gi: [T,B,4H] precomputed W_ih x + b_ih, zero initial state.  Returns hs [T,B,H]."""
    T, B, H4 = gi.shape
    H = H4 // 4
    hs = np.empty((T, B, H), gi.dtype)
    i_ = np.empty_like(hs); f_ = np.empty_like(hs); g_ = np.empty_like(hs); o_ = np.empty_like(hs)
    c_ = np.empty_like(hs); cprev = np.empty_like(hs); hprev = np.empty_like(hs)
    h = np.zeros((B, H), gi.dtype); c = np.zeros((B, H), gi.dtype)
    for t in range(T):
        a = gi[t] + h @ w_hh.T + b_hh
        i = sigmoid(a[:, :H]); f = sigmoid(a[:, H:2 * H]); g = np.tanh(a[:, 2 * H:3 * H]); o = sigmoid(a[:, 3 * H:])
        hprev[t] = h; cprev[t] = c
        c = f * c + i * g
        h = o * np.tanh(c)
        hs[t] = h; i_[t] = i; f_[t] = f; g_[t] = g; o_[t] = o; c_[t] = c
    return hs, dict(i=i_, f=f_, g=g_, o=o_, c=c_, cprev=cprev, hprev=hprev, w_hh=w_hh)


def lstm_seq_bwd(dhs, cache):
    i_, f_, g_, o_, c_, cprev, hprev, w_hh = (cache[k] for k in ("i", "f", "g", "o", "c", "cprev", "hprev", "w_hh"))
    T, B, H = i_.shape
    da_all = np.empty((T, B, 4 * H), i_.dtype)
    dh = np.zeros((B, H), i_.dtype); dc = np.zeros((B, H), i_.dtype)
    for t in range(T - 1, -1, -1):
        dh = dh + dhs[t]
        i, f, g, o = i_[t], f_[t], g_[t], o_[t]
        tc = np.tanh(c_[t])
        do = dh * tc
        dc = dc + dh * o * (1.0 - tc * tc)
        da_all[t, :, :H] = dc * g * i * (1.0 - i)
        da_all[t, :, H:2 * H] = dc * cprev[t] * f * (1.0 - f)
        da_all[t, :, 2 * H:3 * H] = dc * i * (1.0 - g * g)
        da_all[t, :, 3 * H:] = do * o * (1.0 - o)
        dc = dc * f
        dh = da_all[t] @ w_hh
    dw_hh = da_all.reshape(T * B, 4 * H).T @ hprev.reshape(T * B, H)
    db_hh = da_all.sum(axis=(0, 1))
    return da_all, dw_hh, db_hh      # da_all is also dgi


# ----------------------------------------------------------------------------------------
# Bahdanau attention step  S2VTAttModel.py:25-48
# ----------------------------------------------------------------------------------------
def attention_fwd(q, proj_key, enc, v):
    """q: [B,H] (= W_q h), proj_key/enc: [B,N,H], v: [H] -> ctx [B,H], alphas [B,N], tanh e."""
    e = np.tanh(q[:, None, :] + (proj_key if _QK is None else _QK(proj_key)))
    scores = e @ v
    scores = scores - scores.max(axis=1, keepdims=True)
    a = np.exp(scores)
    a = a / a.sum(axis=1, keepdims=True)
    ctx = np.einsum("bn,bnh->bh", a, _q(enc))
    return ctx, a, e


def attention_bwd(dctx, a, e, enc, v):
    """Returns dq [B,H], dproj_key [B,N,H], denc [B,N,H], dv [H]."""
    da = np.einsum("bh,bnh->bn", dctx, _q(enc))
    denc = a[:, :, None] * dctx[:, None, :]
    ds = a * (da - (a * da).sum(axis=1, keepdims=True))
    dv = np.einsum("bn,bnh->h", ds, e)
    de = ds[:, :, None] * v[None, None, :] * (1.0 - e * e)
    return de.sum(axis=1), de, denc, dv


# ----------------------------------------------------------------------------------------
# loss contract  train_utils.py:22-95
# ----------------------------------------------------------------------------------------
def sentence_mask(B, L, s_len, dtype):
    return (np.arange(L)[None, :] < np.asarray(s_len)[:, None]).astype(dtype)


def masked_loss(logits, target, s_len):
    """train_utils.py:37-54.  Returns (loss, dlogits, per-token nll [B,L])."""
    B, L, V = logits.shape
    lsm = log_softmax(logits, axis=2)
    nll = -np.take_along_axis(lsm, target[:, :, None], axis=2)[:, :, 0]
    mask = sentence_mask(B, L, s_len, logits.dtype)
    cnt = mask.sum(axis=1)
    loss = ((nll * mask).sum(axis=1) / cnt).mean()
    w = mask / (cnt[:, None] * B)
    dlogits = np.exp(lsm) * w[:, :, None]
    np.put_along_axis(dlogits, target[:, :, None],
                      np.take_along_axis(dlogits, target[:, :, None], axis=2) - w[:, :, None], axis=2)
    return loss, dlogits, nll


def masked_accuracy(logits, target, s_len):
    """train_utils.py:56-71 (+ train.py:38 argmax).  Returns (acc, pred [B,L] int64)."""
    B, L, _ = logits.shape
    pred = np.argmax(logits, axis=2)           # first max index, as torch.argmax
    mask = sentence_mask(B, L, s_len, logits.dtype)
    acc = ((pred == target).astype(logits.dtype) * mask).sum() / mask.sum()
    return acc, pred


def brevity_loss(probs):
    """train_utils.py:85-95.  Returns (loss, dprobs)."""
    p1 = probs[:, :, 1]
    d = np.zeros_like(probs)
    d[:, :, 1] = 1.0 / p1.shape[0]
    return p1.sum(axis=1).mean(), d


def cont_loss(probs):
    """train_utils.py:73-83.  Returns (loss, dprobs)."""
    p1 = probs[:, :, 1]
    diff = p1[:, 1:] - p1[:, :-1]
    d = np.zeros_like(probs)
    if diff.size:
        sgn = np.sign(diff) / diff.size
        d[:, 1:, 1] += sgn
        d[:, :-1, 1] -= sgn
        return np.abs(diff).mean(), d
    return np.asarray(np.nan, probs.dtype), d


# ----------------------------------------------------------------------------------------
# S2VTAttModel  S2VTAttModel.py:50-264
# ----------------------------------------------------------------------------------------
def s2vtatt_encode(p, vid):
    """Encoder.forward: hoisted input projection + GRU.  vid [B,N,V] -> enc [B,N,H] + cache."""
    B, N, V = vid.shape
    gi = (_mm(vid.reshape(B * N, V), p["encoder.rnn.weight_ih_l0"].T) + p["encoder.rnn.bias_ih_l0"])
    gi = gi.reshape(B, N, -1).transpose(1, 0, 2)
    H = gi.shape[2] // 3
    hs, c = gru_seq_fwd(gi, p["encoder.rnn.weight_hh_l0"], p["encoder.rnn.bias_hh_l0"],
                        np.zeros((B, H), vid.dtype))
    return hs.transpose(1, 0, 2), c            # [B,N,H]


def _dec_in_words(s, sos_id):
    B, L = s.shape
    return np.concatenate([np.full((B, 1), sos_id, s.dtype), s[:, :L - 1]], axis=1)


def s2vtatt_decode_train(p, enc, h0, s, sos_id, max_len, hs_scale=None):
    """Decoder.forward in training mode (always teacher-forced, S2VTAttModel.py:188-189).
    enc [B,N,H], h0 [B,H], s [B,L].  Returns logits [B,L,Vc] and a cache.
    hs_scale [B,L,H] (optional): the Dropout of `pred_linear` (S2VTAttModel.py:121-122,145) with a GIVEN mask, as
    mask / (1 - p) -- what nn.Dropout multiplies the GRU output of step i by before the Linear."""
    B, N, H = enc.shape
    L = max_len
    Wq = p["decoder.attention.query_layer.weight"]; Wk = p["decoder.attention.key_layer.weight"]
    v = p["decoder.attention.energy_layer.weight"][0]
    W_ih = p["decoder.rnn.weight_ih_l0"]; W_hh = p["decoder.rnn.weight_hh_l0"]
    b_ih = p["decoder.rnn.bias_ih_l0"]; b_hh = p["decoder.rnn.bias_hh_l0"]
    Wc, We = W_ih[:, :H], W_ih[:, H:]
    emb = p["decoder.embedding.weight"]
    s_in = _dec_in_words(s, sos_id)[:, :L]
    erow = emb[s_in]                                           # [B,L,E]
    ep = _mm(erow.reshape(B * L, -1), We.T) + b_ih             # hoisted embedding half of in-proj
    ep = ep.reshape(B, L, 3 * H)
    pk = _mm(enc.reshape(B * N, H), Wk.T).reshape(B, N, H)
    h = h0
    steps = []
    hs = np.empty((B, L, H), enc.dtype)
    for i in range(L):
        q = _mm(h, Wq.T)
        ctx, a, e = attention_fwd(q, pk, enc, v)
        gi = _mm(ctx, Wc.T) + ep[:, i]
        gh = _mm(h, W_hh.T) + b_hh
        r = sigmoid(gi[:, :H] + gh[:, :H]); z = sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        steps.append(dict(hprev=h, ctx=ctx, a=a, e=e, r=r, z=z, n=n, ghn=gh[:, 2 * H:]))
        h = (1.0 - z) * n + z * h
        hs[:, i] = h
    hs_out = hs if hs_scale is None else hs * hs_scale
    logits = _mm(hs_out.reshape(B * L, H), p["decoder.pred_linear.1.weight"].T) + p["decoder.pred_linear.1.bias"]
    cache = dict(steps=steps, hs=hs, hs_out=hs_out, hs_scale=hs_scale, pk=pk, enc=enc, erow=erow, s_in=s_in, h0=h0)
    return logits.reshape(B, L, -1), cache


def s2vtatt_decode_bwd(p, cache, dlogits):
    """Backward of s2vtatt_decode_train.  Returns (grads dict, denc [B,N,H], dh0 [B,H])."""
    steps, hs, pk, enc, erow, s_in = (cache[k] for k in ("steps", "hs", "pk", "enc", "erow", "s_in"))
    B, L, H = hs.shape
    N = enc.shape[1]
    Wq = p["decoder.attention.query_layer.weight"]; Wk = p["decoder.attention.key_layer.weight"]
    v = p["decoder.attention.energy_layer.weight"][0]
    W_ih = p["decoder.rnn.weight_ih_l0"]; W_hh = p["decoder.rnn.weight_hh_l0"]
    Wc, We = W_ih[:, :H], W_ih[:, H:]
    Wv = p["decoder.pred_linear.1.weight"]
    Vc = Wv.shape[0]
    dl = dlogits.reshape(B * L, Vc)
    g = {}
    g["decoder.pred_linear.1.weight"] = _mm(dl.T, cache.get("hs_out", hs).reshape(B * L, H))
    g["decoder.pred_linear.1.bias"] = _q(dl).sum(axis=0)
    dhs = _mm(dl, Wv).reshape(B, L, H)
    if cache.get("hs_scale") is not None:
        dhs = dhs * cache["hs_scale"]
    dWq = np.zeros_like(Wq); dv = np.zeros_like(v); dWc = np.zeros_like(Wc); dW_hh = np.zeros_like(W_hh)
    db_hh = np.zeros(3 * H, enc.dtype)
    dgi_all = np.empty((B, L, 3 * H), enc.dtype)
    dpk = np.zeros_like(pk); denc = np.zeros_like(enc)
    dh = np.zeros((B, H), enc.dtype)
    for i in range(L - 1, -1, -1):
        st = steps[i]
        dh = dh + dhs[:, i]
        r, z, n, ghn, hp = st["r"], st["z"], st["n"], st["ghn"], st["hprev"]
        dn = dh * (1.0 - z); dz = dh * (hp - n)
        dnp = dn * (1.0 - n * n); dzp = dz * z * (1.0 - z); drp = dnp * ghn * r * (1.0 - r)
        dgi = np.concatenate([drp, dzp, dnp], axis=1)
        dgh = np.concatenate([drp, dzp, dnp * r], axis=1)
        dgi_all[:, i] = dgi
        dW_hh += _mm(dgh.T, hp); db_hh += dgh.sum(axis=0)
        dWc += _mm(dgi.T, st["ctx"])
        dctx = _mm(dgi, Wc)
        dq, de, denc_i, dv_i = attention_bwd(dctx, st["a"], st["e"], enc, v)
        dpk += de; denc += denc_i; dv += dv_i
        dWq += _mm(dq.T, hp)
        dh = dh * z + _mm(dgh, W_hh) + _mm(dq, Wq)
    dgi_flat = dgi_all.reshape(B * L, 3 * H)
    dWe = _mm(dgi_flat.T, erow.reshape(B * L, -1))
    derow = _mm(dgi_flat, We)
    demb = np.zeros_like(p["decoder.embedding.weight"])
    np.add.at(demb, s_in.reshape(-1), derow)
    g["decoder.embedding.weight"] = demb
    g["decoder.rnn.weight_ih_l0"] = np.concatenate([dWc, dWe], axis=1)
    g["decoder.rnn.weight_hh_l0"] = dW_hh
    g["decoder.rnn.bias_ih_l0"] = dgi_flat.sum(axis=0)
    g["decoder.rnn.bias_hh_l0"] = db_hh
    g["decoder.attention.query_layer.weight"] = dWq
    g["decoder.attention.energy_layer.weight"] = dv[None, :]
    g["decoder.attention.key_layer.weight"] = _mm(dpk.reshape(B * N, H).T, enc.reshape(B * N, H))
    denc += _mm(dpk.reshape(B * N, H), Wk).reshape(B, N, H)
    return g, denc, dh


def s2vtatt_encode_bwd(p, vid, enc_cache, denc, dh_final, need_dvid=False):
    """Backward of s2vtatt_encode.  denc [B,N,H] gradient on every encoder output, dh_final [B,H]
    gradient on the returned final state."""
    B, N, V = vid.shape
    dgi, dw_hh, db_hh, _ = gru_seq_bwd(denc.transpose(1, 0, 2), dh_final, enc_cache)
    dgi_bn = dgi.transpose(1, 0, 2).reshape(B * N, -1)
    g = {"encoder.rnn.weight_ih_l0": _mm(dgi_bn.T, vid.reshape(B * N, V)),
         "encoder.rnn.bias_ih_l0": dgi_bn.sum(axis=0),
         "encoder.rnn.weight_hh_l0": dw_hh, "encoder.rnn.bias_hh_l0": db_hh}
    dvid = (dgi_bn @ p["encoder.rnn.weight_ih_l0"]).reshape(B, N, V) if need_dvid else None
    return g, dvid


def s2vtatt_forward_train(p, vid, s, sos_id, max_len, hs_scale=None):
    """S2VTAttModel.forward (training).  Returns logits [B,L,Vc], cache (cache['alphas'] [L,B,N])."""
    enc, ec = s2vtatt_encode(p, vid)
    logits, dc = s2vtatt_decode_train(p, enc, enc[:, -1], s, sos_id, max_len, hs_scale=hs_scale)
    dc["alphas"] = np.stack([st["a"] for st in dc["steps"]])
    return logits, dict(vid=vid, enc_cache=ec, dec_cache=dc, alphas=dc["alphas"])


def s2vtatt_backward(p, cache, dlogits, need_dvid=False):
    g, denc, dh0 = s2vtatt_decode_bwd(p, cache["dec_cache"], dlogits)
    ge, dvid = s2vtatt_encode_bwd(p, cache["vid"], cache["enc_cache"], denc, dh0, need_dvid)
    g.update(ge)
    return g, dvid


def s2vtatt_greedy(p, vid, sos_id, max_len):
    """S2VTAttModel.forward in eval mode (S2VTAttModel.py:172-173,190-191): fixed max_len steps,
    argmax feedback, no early stop.  Returns (ids [B,L] int64, logits [B,L,Vc], alphas [L,B,N])."""
    enc, _ = s2vtatt_encode(p, vid)
    B, N, H = enc.shape
    Wq = p["decoder.attention.query_layer.weight"]; Wk = p["decoder.attention.key_layer.weight"]
    v = p["decoder.attention.energy_layer.weight"][0]
    W_ih = p["decoder.rnn.weight_ih_l0"]; W_hh = p["decoder.rnn.weight_hh_l0"]
    b_ih = p["decoder.rnn.bias_ih_l0"]; b_hh = p["decoder.rnn.bias_hh_l0"]
    Wv = p["decoder.pred_linear.1.weight"]; bv = p["decoder.pred_linear.1.bias"]
    emb = p["decoder.embedding.weight"]
    pk = (enc.reshape(B * N, H) @ Wk.T).reshape(B, N, H)
    h = enc[:, -1]
    w = np.full((B,), sos_id, np.int64)
    ids, outs, alphas = [], [], []
    for _ in range(max_len):
        ctx, a, _e = attention_fwd(h @ Wq.T, pk, enc, v)
        gi = np.concatenate([ctx, emb[w]], axis=1) @ W_ih.T + b_ih
        gh = h @ W_hh.T + b_hh
        r = sigmoid(gi[:, :H] + gh[:, :H]); z = sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        o = h @ Wv.T + bv
        w = np.argmax(o, axis=1)
        ids.append(w); outs.append(o); alphas.append(a)
    return np.stack(ids, 1), np.stack(outs, 1), np.stack(alphas)


# ----------------------------------------------------------------------------------------
# S2VTModel  S2VTModel.py:74-202
# ----------------------------------------------------------------------------------------
def _s2vt_params(p):
    return (p["rnn1.weight_ih_l0"], p["rnn1.weight_hh_l0"], p["rnn1.bias_ih_l0"], p["rnn1.bias_hh_l0"],
            p["rnn2.weight_ih_l0"], p["rnn2.weight_hh_l0"], p["rnn2.bias_ih_l0"], p["rnn2.bias_hh_l0"],
            p["embedding.0.weight"], p["linear.1.weight"], p["linear.1.bias"])


def s2vt_forward(p, vid, s, sos_id, max_len, teacher=None, train=True):
    """S2VTModel.forward.  ``teacher``: list of L bools — the outcome of the per-step coin
    ``random.random() < teacher_force_prob`` (S2VTModel.py:134); None = all True (prob 1.0).
    In eval (train=False) the argmax is always fed back (S2VTModel.py:147-177).
    Returns logits [B,L,Vc] and a cache."""
    W1i, W1h, b1i, b1h, W2i, W2h, b2i, b2h, emb, Wv, bv = _s2vt_params(p)
    B, N, V = vid.shape
    H = W1h.shape[1]
    L = max_len
    dt = vid.dtype
    # rnn1 runs N frames then L steps on an all-zero input (in-proj == b_ih): one sequence of N+L
    gi1 = (vid.reshape(B * N, V) @ W1i.T + b1i).reshape(B, N, 3 * H).transpose(1, 0, 2)
    gi1 = np.concatenate([gi1, np.broadcast_to(b1i, (L, B, 3 * H)).astype(dt)], axis=0)
    out1, c1 = gru_seq_fwd(gi1, W1h, b1h, np.zeros((B, H), dt))           # [N+L,B,H]
    W2o, W2e = W2i[:, :H], W2i[:, H:]
    gi2h = (out1.reshape((N + L) * B, H) @ W2o.T + b2i).reshape(N + L, B, 3 * H)
    if train:
        s_in = np.concatenate([np.full((B, 1), sos_id, s.dtype), s], axis=1)     # [B,L+1]
        teacher = [True] * L if teacher is None else list(teacher)
    # layer 2: N encode steps with zero word padding, then L decode steps
    h2 = np.zeros((B, H), dt)
    w = np.full((B,), sos_id, np.int64)
    T = N + L
    gi2 = np.empty((T, B, 3 * H), dt)
    hs2 = np.empty((T, B, H), dt)
    r_ = np.empty((T, B, H), dt); z_ = np.empty_like(r_); n_ = np.empty_like(r_); ghn_ = np.empty_like(r_)
    hprev = np.empty_like(r_)
    words = np.empty((L, B), np.int64)
    logits = np.empty((B, L, Wv.shape[0]), dt)
    for t in range(T):
        g = gi2h[t]
        if t >= N:
            words[t - N] = w
            g = g + emb[w] @ W2e.T
        gi2[t] = g
        gh = h2 @ W2h.T + b2h
        r = sigmoid(g[:, :H] + gh[:, :H]); z = sigmoid(g[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(g[:, 2 * H:] + r * gh[:, 2 * H:])
        hprev[t] = h2
        h2 = (1.0 - z) * n + z * h2
        hs2[t] = h2; r_[t] = r; z_[t] = z; n_[t] = n; ghn_[t] = gh[:, 2 * H:]
        if t >= N:
            i = t - N
            o = h2 @ Wv.T + bv
            logits[:, i] = o
            if train and teacher[i]:
                w = s_in[:, i + 1]
            else:
                w = np.argmax(o, axis=1)
    c2 = dict(r=r_, z=z_, n=n_, ghn=ghn_, hprev=hprev, w_hh=W2h)
    return logits, dict(vid=vid, c1=c1, c2=c2, out1=out1, hs2=hs2, words=words, N=N, L=L)


def s2vt_backward(p, cache, dlogits):
    """Backward of s2vt_forward (argmax feedback carries no gradient)."""
    W1i, W1h, b1i, b1h, W2i, W2h, b2i, b2h, emb, Wv, bv = _s2vt_params(p)
    vid, c1, c2, out1, hs2, words, N, L = (cache[k] for k in ("vid", "c1", "c2", "out1", "hs2", "words", "N", "L"))
    B, _, V = vid.shape
    H = W1h.shape[1]
    T = N + L
    Vc = Wv.shape[0]
    W2o, W2e = W2i[:, :H], W2i[:, H:]
    dl = dlogits.transpose(1, 0, 2).reshape(L * B, Vc)                  # [L,B,Vc]
    hdec = hs2[N:].reshape(L * B, H)
    g = {"linear.1.weight": dl.T @ hdec, "linear.1.bias": dl.sum(axis=0)}
    dhs2 = np.zeros((T, B, H), vid.dtype)
    dhs2[N:] = (dl @ Wv).reshape(L, B, H)
    dgi2, dW2h, db2h, _ = gru_seq_bwd(dhs2, None, c2)
    dgi2f = dgi2.reshape(T * B, 3 * H)
    erow = emb[words.reshape(-1)]                                       # [L*B,E]
    dgi2dec = dgi2[N:].reshape(L * B, 3 * H)
    dW2e = dgi2dec.T @ erow
    dW2o = dgi2f.T @ out1.reshape(T * B, H)
    demb = np.zeros_like(emb)
    np.add.at(demb, words.reshape(-1), dgi2dec @ W2e)
    g["embedding.0.weight"] = demb
    g["rnn2.weight_ih_l0"] = np.concatenate([dW2o, dW2e], axis=1)
    g["rnn2.weight_hh_l0"] = dW2h; g["rnn2.bias_ih_l0"] = dgi2f.sum(axis=0); g["rnn2.bias_hh_l0"] = db2h
    dout1 = (dgi2f @ W2o).reshape(T, B, H)
    dgi1, dW1h, db1h, _ = gru_seq_bwd(dout1, None, c1)
    dgi1_bn = dgi1[:N].transpose(1, 0, 2).reshape(B * N, 3 * H)
    g["rnn1.weight_ih_l0"] = dgi1_bn.T @ vid.reshape(B * N, V)
    g["rnn1.bias_ih_l0"] = dgi1.reshape(T * B, 3 * H).sum(axis=0)
    g["rnn1.weight_hh_l0"] = dW1h; g["rnn1.bias_hh_l0"] = db1h
    return g, dgi1_bn @ W1i


# ----------------------------------------------------------------------------------------
# RationaleNet generator  RationaleNet.py:14-54
# ----------------------------------------------------------------------------------------
def generator_fwd(p, vid, tau, noise, hard=False):
    """Generator.forward with the Gumbel noise injected: ``noise`` [B*N,2] are the Exp(1) draws
    of ``torch.empty_like(logits).exponential_()`` (F.gumbel_softmax), g = -log(noise).
    Returns sel_vid_feats [B,N,V], probs [B,N,2], cache."""
    B, N, V = vid.shape
    x = vid.reshape(B * N, V)
    outs, caches = [], []
    for sfx, rev in (("", False), ("_reverse", True)):
        gi = (x @ p["rnn.weight_ih_l0" + sfx].T + p["rnn.bias_ih_l0" + sfx]).reshape(B, N, -1).transpose(1, 0, 2)
        if rev:
            gi = gi[::-1]
        hs, c = lstm_seq_fwd(gi, p["rnn.weight_hh_l0" + sfx], p["rnn.bias_hh_l0" + sfx])
        outs.append(hs[::-1] if rev else hs); caches.append(c)
    out = np.concatenate(outs, axis=2).transpose(1, 0, 2)                 # [B,N,2H]
    logits = out.reshape(B * N, -1) @ p["linear.weight"].T + p["linear.bias"]
    gum = -np.log(noise.astype(vid.dtype))
    y = (logits + gum) / tau
    y = y - y.max(axis=1, keepdims=True)
    y = np.exp(y); y = y / y.sum(axis=1, keepdims=True)                  # soft sample
    if hard:
        idx = np.argmax(y, axis=1)
        yh = np.zeros_like(y); yh[np.arange(y.shape[0]), idx] = 1.0
        probs = (yh - y) + y                                             # straight-through value
    else:
        probs = y
    probs = probs.reshape(B, N, 2)
    sel = vid * probs[:, :, 1:2]
    return sel, probs, dict(vid=vid, out=out, y=y.reshape(B, N, 2), caches=caches, tau=tau)


def generator_bwd(p, cache, dsel, dprobs):
    """Backward of generator_fwd (soft or straight-through: gradient flows through y)."""
    vid, out, y, caches, tau = (cache[k] for k in ("vid", "out", "y", "caches", "tau"))
    B, N, V = vid.shape
    H2 = out.shape[2]; H = H2 // 2
    dprobs = dprobs.copy()
    dprobs[:, :, 1] += (dsel * vid).sum(axis=2)
    dy = dprobs.reshape(B * N, 2)
    yf = y.reshape(B * N, 2)
    dlog = yf * (dy - (yf * dy).sum(axis=1, keepdims=True)) / tau
    g = {"linear.weight": dlog.T @ out.reshape(B * N, H2), "linear.bias": dlog.sum(axis=0)}
    dout = (dlog @ p["linear.weight"]).reshape(B, N, H2).transpose(1, 0, 2)       # [N,B,2H]
    x = vid.reshape(B * N, V)
    for d, (sfx, rev) in enumerate((("", False), ("_reverse", True))):
        dhs = dout[:, :, d * H:(d + 1) * H]
        if rev:
            dhs = dhs[::-1]
        dgi, dw_hh, db_hh = lstm_seq_bwd(dhs, caches[d])
        if rev:
            dgi = dgi[::-1]
        dgi_bn = dgi.transpose(1, 0, 2).reshape(B * N, -1)
        g["rnn.weight_ih_l0" + sfx] = dgi_bn.T @ x
        g["rnn.bias_ih_l0" + sfx] = dgi_bn.sum(axis=0)
        g["rnn.weight_hh_l0" + sfx] = dw_hh
        g["rnn.bias_hh_l0" + sfx] = db_hh
    return g


# ----------------------------------------------------------------------------------------
# whole training iterations (train.py:32-44, train_rationale.py:30-44 run_iter + backward)
# ----------------------------------------------------------------------------------------
def train_iter_s2vtatt(p, vid, s, s_len, sos_id, max_len, hs_scale=None):
    logits, cache = s2vtatt_forward_train(p, vid, s, sos_id, max_len, hs_scale=hs_scale)
    loss, dlogits, nll = masked_loss(logits, s, s_len)
    acc, pred = masked_accuracy(logits, s, s_len)
    grads, _ = s2vtatt_backward(p, cache, dlogits)
    return dict(loss=loss, acc=acc, pred=pred, logits=logits, alphas=cache["alphas"], grads=grads, token_nll=nll)


def train_iter_s2vt(p, vid, s, s_len, sos_id, max_len, teacher=None):
    logits, cache = s2vt_forward(p, vid, s, sos_id, max_len, teacher=teacher, train=True)
    loss, dlogits, _ = masked_loss(logits, s, s_len)
    acc, pred = masked_accuracy(logits, s, s_len)
    grads, _ = s2vt_backward(p, cache, dlogits)
    return dict(loss=loss, acc=acc, pred=pred, logits=logits, grads=grads)


def train_iter_rationale(p, vid, s, s_len, sos_id, max_len, tau, noise, arch="s2vt-att",
                         lambda_brev=1.0, lambda_cont=1.0, teacher=None):
    """RationaleNet.forward + train_rationale.py:30-44 loss + backward."""
    pg, pc = _sub(p, "gen."), _sub(p, "caption_net.")
    sel, probs, gc = generator_fwd(pg, vid, tau, noise, hard=False)
    if arch == "s2vt-att":
        logits, cache = s2vtatt_forward_train(pc, sel, s, sos_id, max_len)
    else:
        logits, cache = s2vt_forward(pc, sel, s, sos_id, max_len, teacher=teacher, train=True)
    loss_ce, dlogits, _ = masked_loss(logits, s, s_len)
    lb, db = brevity_loss(probs)
    lc, dc = cont_loss(probs)
    acc, pred = masked_accuracy(logits, s, s_len)
    if arch == "s2vt-att":
        gcap, dsel = s2vtatt_backward(pc, cache, dlogits, need_dvid=True)
    else:
        gcap, dsel = s2vt_backward(pc, cache, dlogits)
        dsel = dsel.reshape(vid.shape)
    ggen = generator_bwd(pg, gc, dsel, lambda_brev * db + lambda_cont * dc)
    grads = {"caption_net." + k: v for k, v in gcap.items()}
    grads.update({"gen." + k: v for k, v in ggen.items()})
    return dict(loss=loss_ce + lambda_brev * lb + lambda_cont * lc, loss_ce=loss_ce, loss_brev=lambda_brev * lb,
                loss_cont=lambda_cont * lc, rationale_len=probs[:, :, 1].sum(axis=1).mean(), acc=acc, pred=pred,
                logits=logits, probs=probs, grads=grads)


# ----------------------------------------------------------------------------------------
# optimizer step (train.py:104-105,157-160): clip_grad_norm_(params, max_norm) + torch.optim.Adam(lr, weight_decay)
# ----------------------------------------------------------------------------------------
def clip_adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                   max_norm=None):
    """One optimizer step on dicts of arrays (updated in place); ``step`` is the 1-based count of this step.
    Returns the total gradient norm before clipping.  Follows torch.nn.utils.clip_grad_norm_ (coefficient
    max_norm / (norm + 1e-6), clamped to 1) and torch.optim.Adam with L2-style decay (grad += weight_decay * param)."""
    total = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))
    coef = 1.0
    if max_norm is not None and max_norm > 0:
        coef = min(max_norm / (total + 1e-6), 1.0)
    b1, b2 = betas
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    for k in params:
        g = grads[k] * coef + weight_decay * params[k]
        exp_avg[k] += (1.0 - b1) * (g - exp_avg[k])
        exp_avg_sq[k] *= b2
        exp_avg_sq[k] += (1.0 - b2) * g * g
        denom = np.sqrt(exp_avg_sq[k]) / np.sqrt(bc2) + eps
        params[k] -= (lr / bc1) * exp_avg[k] / denom
    return total


# ----------------------------------------------------------------------------------------
# beam search over Decoder.forward_step.  The reference ships no beam search: this DEFINES it as
# the plain fixed-length search over the reference step (S2VTAttModel.py:125-148) -- every hypothesis is extended for
# exactly max_len steps like the reference's greedy eval branch (:172-191, no early stop), score = sum of
# log-softmax, the K best of the K x Vc candidates of a video survive (ties: lower flat index beam*Vc + word first) --
# so that beam 1 IS the reference's greedy decoding.  Pinned by tests/golden/s2vtatt_beam_*.npz, which
# oracle/gen_golden_beam.py produces by driving the unmodified reference modules with the same search.
# ----------------------------------------------------------------------------------------
def beam_select(score, logp, K, first):
    """score [B,K], logp [B,K,Vc] -> (new score [B,K], parent [B,K], word [B,K]).  At the first step only beam 0 is live."""
    B, _, Vc = logp.shape
    cand = score[:, :, None] + logp
    if first:
        cand[:, 1:, :] = -np.inf
    flat = cand.reshape(B, K * Vc)
    order = np.argsort(-flat, axis=1, kind="stable")[:, :K]          # stable: lower flat index wins ties
    return np.take_along_axis(flat, order, 1), order // Vc, order % Vc


def s2vtatt_beam_search(p, vid, sos_id, max_len, K):
    """Returns (ids [B,K,L] int64, scores [B,K]) sorted best first."""
    enc, _ = s2vtatt_encode(p, vid)
    B, N, H = enc.shape
    Wq = p["decoder.attention.query_layer.weight"]; Wk = p["decoder.attention.key_layer.weight"]
    v = p["decoder.attention.energy_layer.weight"][0]
    W_ih = p["decoder.rnn.weight_ih_l0"]; W_hh = p["decoder.rnn.weight_hh_l0"]
    b_ih = p["decoder.rnn.bias_ih_l0"]; b_hh = p["decoder.rnn.bias_hh_l0"]
    Wv = p["decoder.pred_linear.1.weight"]; bv = p["decoder.pred_linear.1.bias"]
    emb = p["decoder.embedding.weight"]
    pk = (enc.reshape(B * N, H) @ Wk.T).reshape(B, N, H)
    encr, pkr = np.repeat(enc, K, axis=0), np.repeat(pk, K, axis=0)         # rows b*K + k
    h = np.repeat(enc[:, -1], K, axis=0)
    w = np.full((B * K,), sos_id, np.int64)
    score = np.zeros((B, K), enc.dtype)
    ids = np.zeros((B, K, max_len), np.int64)
    for i in range(max_len):
        ctx, _a, _e = attention_fwd(h @ Wq.T, pkr, encr, v)
        gi = np.concatenate([ctx, emb[w]], axis=1) @ W_ih.T + b_ih
        gh = h @ W_hh.T + b_hh
        r = sigmoid(gi[:, :H] + gh[:, :H]); z = sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        logp = log_softmax(h @ Wv.T + bv).reshape(B, K, -1)
        score, parent, word = beam_select(score, logp, K, first=(i == 0))
        rows = (np.arange(B)[:, None] * K + parent).reshape(-1)
        h = h[rows]
        ids = np.take_along_axis(ids, parent[:, :, None], 1)
        ids[:, :, i] = word
        w = word.reshape(-1)
    return ids, score
