"""CPU restatement of the reference's SpatialNet (model/SpatialNet.py:12-142) for both caption nets ('s2vt-att', 's2vt'), forward and
hand-derived backward, in numpy.  TEST INFRASTRUCTURE: only tests/ may import this; the product path never does.

Pinned (tests/test_oracle_golden.py::test_spatialnet_train) at 1e-10 against tests/golden/spatial_att_tiny.npz / spatial_s2vt_tiny.npz, which
oracle/gen_golden_spatial.py produced by running the UNMODIFIED reference SpatialNet (float64, torch autograd) in the authoring
container: logits, seq_alphas, loss and every parameter gradient.

What it restates, with the reference lines each function follows:
  * conv_bn_relu_fwd / _bwd   -- nn.Conv2d(C_in, C_out, 3, 1, 1) + nn.BatchNorm2d (training: batch statistics, biased variance)
                                 + nn.ReLU, SpatialNet.py:76-86 (the `self.conv` Sequential applied at :106)
  * spatialnet_forward_train  -- SpatialNet.forward :99-142: conv features -> [B,N,K^2,H] keys, input features -> [B,N,K^2,F]
                                 values (:106-112), zero initial state (:114), per frame Attention.forward (:27-53: query_layer,
                                 key_layer, energy_layer, softmax over the cells, weighted sum of the frame's input features) and
                                 caption_net.encode_step (:127 = one encoder GRU step, S2VTAttModel.py:63-78), then
                                 caption_net.decode (:140 = S2VTAttModel.py:231-243, restated by captioning_oracle.s2vtatt_decode_train;
                                 or S2VTModel.decode, S2VTModel.py:88-177, restated by s2vt_decode_fwd below)
  * spatialnet_backward       -- autograd of the above
The key projection is hoisted out of the frame loop (the reference re-applies key_layer per frame, :39; it does not depend on the
state), exactly as the CUDA path does; the numbers are the same.
"""
import numpy as np

from oracle import captioning_oracle as O


# ----------------------------------------------------------------------------------------
# Conv3x3 (stride 1, pad 1) + BatchNorm2d (batch statistics) + ReLU      SpatialNet.py:76-86
# ----------------------------------------------------------------------------------------
def conv3x3_fwd(x, w, b):
    """x [I,C,K,K], w [O,C,3,3], b [O] -> y [I,O,K,K] (cross-correlation, zero padding 1, as torch.nn.Conv2d)."""
    I, C, K, _ = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    y = np.zeros((I, w.shape[0], K, K), x.dtype)
    for dy in range(3):
        for dx in range(3):
            y += np.einsum("icyx,oc->ioyx", xp[:, :, dy:dy + K, dx:dx + K], w[:, :, dy, dx])
    return y + b[None, :, None, None]


def conv3x3_bwd(x, w, dy, need_dx=True):
    """Returns (dw [O,C,3,3], db [O], dx [I,C,K,K] | None)."""
    I, C, K, _ = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    dw = np.empty_like(w)
    dxp = np.zeros_like(xp) if need_dx else None
    for a in range(3):
        for c in range(3):
            dw[:, :, a, c] = np.einsum("ioyx,icyx->oc", dy, xp[:, :, a:a + K, c:c + K])
            if need_dx:
                dxp[:, :, a:a + K, c:c + K] += np.einsum("ioyx,oc->icyx", dy, w[:, :, a, c])
    return dw, dy.sum(axis=(0, 2, 3)), (dxp[:, :, 1:K + 1, 1:K + 1] if need_dx else None)


def bn_relu_fwd(y, gamma, beta, eps=1e-5):
    """BatchNorm2d in training mode (statistics over images and cells per channel, biased variance) + ReLU."""
    mean = y.mean(axis=(0, 2, 3))
    var = y.var(axis=(0, 2, 3))
    invstd = 1.0 / np.sqrt(var + eps)
    xhat = (y - mean[None, :, None, None]) * invstd[None, :, None, None]
    pre = xhat * gamma[None, :, None, None] + beta[None, :, None, None]
    return np.maximum(pre, 0.0), dict(xhat=xhat, invstd=invstd, pre=pre, mean=mean, var=var)


def bn_relu_bwd(dz, cache, gamma):
    """Returns (dy, dgamma, dbeta)."""
    xhat, invstd, pre = cache["xhat"], cache["invstd"], cache["pre"]
    d = dz * (pre > 0.0)
    M = d.shape[0] * d.shape[2] * d.shape[3]
    dbeta = d.sum(axis=(0, 2, 3))
    dgamma = (d * xhat).sum(axis=(0, 2, 3))
    dy = (gamma * invstd)[None, :, None, None] * (d - dbeta[None, :, None, None] / M - xhat * dgamma[None, :, None, None] / M)
    return dy, dgamma, dbeta


def bn_running_update(running_mean, running_var, cache, count, momentum=0.1):
    """torch.nn.BatchNorm2d's in-place update of the running estimates in training (unbiased variance)."""
    unbiased = cache["var"] * count / (count - 1.0) if count > 1 else cache["var"]
    return (1.0 - momentum) * running_mean + momentum * cache["mean"], (1.0 - momentum) * running_var + momentum * unbiased


# ----------------------------------------------------------------------------------------
# S2VTModel.decode(output1, state1, s)      S2VTModel.py:88-177 (training: teacher_force_prob = 1)
# ----------------------------------------------------------------------------------------
def s2vt_decode_fwd(cp, out1_enc, h1, s, sos_id, max_len):
    """out1_enc [N,B,H] = rnn1 outputs of the encoding stage, h1 [B,H] its final state.  rnn2 runs over [out1 ; 0_E] for the N
    frames (:98-107), then L steps: rnn1 on a zero input (its input projection is b_ih, :120-122), rnn2 on [out1 ; Emb[w]] (:123-129),
    Linear (:130); the next word is the teacher's (:134-136).  Returns logits [B,L,Vc] and a cache."""
    W1i, W1h, b1i, b1h, W2i, W2h, b2i, b2h, emb, Wv, bv = O._s2vt_params(cp)
    N, B, H = out1_enc.shape
    L, dt = max_len, out1_enc.dtype
    gi1d = np.broadcast_to(b1i, (L, B, 3 * H)).astype(dt)
    out1_dec, c1d = O.gru_seq_fwd(gi1d, W1h, b1h, h1)
    out1 = np.concatenate([out1_enc, out1_dec], axis=0)                     # [N+L,B,H]
    W2o, W2e = W2i[:, :H], W2i[:, H:]
    s_in = np.concatenate([np.full((B, 1), sos_id, s.dtype), s], axis=1)    # [B,L+1]
    T = N + L
    gi2 = (out1.reshape(T * B, H) @ W2o.T + b2i).reshape(T, B, 3 * H).copy()
    words = s_in[:, :L].T.copy()                                            # [L,B]: <sos>, s_0, ..., s_{L-2}
    gi2[N:] += (emb[words.reshape(-1)] @ W2e.T).reshape(L, B, 3 * H)
    hs2, c2 = O.gru_seq_fwd(gi2, W2h, b2h, np.zeros((B, H), dt))
    logits = (hs2[N:].reshape(L * B, H) @ Wv.T + bv).reshape(L, B, -1).transpose(1, 0, 2)
    return logits, dict(c1d=c1d, c2=c2, out1=out1, hs2=hs2, words=words, N=N, L=L)


def s2vt_decode_bwd(cp, cache, dlogits):
    """Returns (grads of the S2VT parameters from the decode stage, d out1_enc [N,B,H], d h1 [B,H])."""
    W1i, W1h, b1i, b1h, W2i, W2h, b2i, b2h, emb, Wv, bv = O._s2vt_params(cp)
    c1d, c2, out1, hs2, words, N, L = (cache[k] for k in ("c1d", "c2", "out1", "hs2", "words", "N", "L"))
    T, B, H = out1.shape
    Vc = Wv.shape[0]
    W2o, W2e = W2i[:, :H], W2i[:, H:]
    dl = dlogits.transpose(1, 0, 2).reshape(L * B, Vc)
    g = {"linear.1.weight": dl.T @ hs2[N:].reshape(L * B, H), "linear.1.bias": dl.sum(axis=0)}
    dhs2 = np.zeros((T, B, H), out1.dtype)
    dhs2[N:] = (dl @ Wv).reshape(L, B, H)
    dgi2, dW2h, db2h, _ = O.gru_seq_bwd(dhs2, None, c2)
    dgi2f = dgi2.reshape(T * B, 3 * H)
    dgi2dec = dgi2[N:].reshape(L * B, 3 * H)
    demb = np.zeros_like(emb)
    np.add.at(demb, words.reshape(-1), dgi2dec @ W2e)
    g["embedding.0.weight"] = demb
    g["rnn2.weight_ih_l0"] = np.concatenate([dgi2f.T @ out1.reshape(T * B, H), dgi2dec.T @ emb[words.reshape(-1)]], axis=1)
    g["rnn2.weight_hh_l0"] = dW2h; g["rnn2.bias_ih_l0"] = dgi2f.sum(axis=0); g["rnn2.bias_hh_l0"] = db2h
    dout1 = (dgi2f @ W2o).reshape(T, B, H)
    dgi1d, dW1h, db1h, dh1 = O.gru_seq_bwd(dout1[N:], None, c1d)            # rnn1's L zero-input steps
    g["rnn1.weight_hh_l0"] = dW1h; g["rnn1.bias_hh_l0"] = db1h
    g["rnn1.bias_ih_l0"] = dgi1d.reshape(L * B, 3 * H).sum(axis=0)
    g["rnn1.weight_ih_l0"] = np.zeros_like(W1i)                             # a zero input carries no weight gradient
    return g, dout1[:N], dh1


# ----------------------------------------------------------------------------------------
# SpatialNet.forward / backward      SpatialNet.py:99-142
# ----------------------------------------------------------------------------------------
def _enc_gru(cp):
    """The GRU behind caption_net.encode_step: S2VTAttModel's encoder.rnn (S2VTAttModel.py:63-78) or S2VTModel's rnn1 (S2VTModel.py:57-72)."""
    pre = "encoder.rnn." if "encoder.rnn.weight_ih_l0" in cp else "rnn1."
    return pre, cp[pre + "weight_ih_l0"], cp[pre + "weight_hh_l0"], cp[pre + "bias_ih_l0"], cp[pre + "bias_hh_l0"]
def _caption(p):
    return {k[len("caption_net."):]: v for k, v in p.items() if k.startswith("caption_net.")}


def spatialnet_forward_train(p, vid, s, sos_id, max_len):
    """p: the reference state_dict as numpy arrays (keys conv.*, attention.*, caption_net.*); vid [B,N,F,K,K]; s [B,L].
    Returns logits [B,L,Vc], seq_alphas [B,N,K,K] and a cache."""
    B, N, F, K, _ = vid.shape
    cells = K * K
    x = vid.reshape(B * N, F, K, K)
    y1 = conv3x3_fwd(x, p["conv.0.weight"], p["conv.0.bias"])
    z1, c1 = bn_relu_fwd(y1, p["conv.1.weight"], p["conv.1.bias"])
    y2 = conv3x3_fwd(z1, p["conv.3.weight"], p["conv.3.bias"])
    z2, c2 = bn_relu_fwd(y2, p["conv.4.weight"], p["conv.4.bias"])
    H = z2.shape[1]
    conv_feats = z2.reshape(B, N, H, cells).transpose(0, 1, 3, 2)            # [B,N,K^2,H]   (:107-110)
    feats = vid.reshape(B, N, F, cells).transpose(0, 1, 3, 2)               # [B,N,K^2,F]   (:111-112)
    Wk, Wq = p["attention.key_layer.weight"], p["attention.query_layer.weight"]
    v = p["attention.energy_layer.weight"][0]
    cp = _caption(p)
    pre, W_ih, W_hh, b_ih, b_hh = _enc_gru(cp)
    pk = conv_feats @ Wk.T                                                  # hoisted key_layer (:39)
    h = np.zeros((B, H), vid.dtype)                                         # :114
    steps, outs, alphas = [], [], []
    for t in range(N):
        q = h @ Wq.T
        ctx, a, e = O.attention_fwd(q, pk[:, t], feats[:, t], v)            # keys H wide, values F wide (:27-53)
        gi = ctx @ W_ih.T + b_ih                                            # encode_step (:127), one GRU step
        gh = h @ W_hh.T + b_hh
        r = O.sigmoid(gi[:, :H] + gh[:, :H]); z = O.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        steps.append(dict(hprev=h, ctx=ctx, a=a, e=e, r=r, z=z, n=n, ghn=gh[:, 2 * H:]))
        h = (1.0 - z) * n + z * h
        outs.append(h); alphas.append(a)
    if pre == "encoder.rnn.":
        enc = np.stack(outs, axis=1)                                        # [B,N,H] = output1 [N,B,H] transposed (:231-243)
        logits, dc = O.s2vtatt_decode_train(cp, enc, h, s, sos_id, max_len)
    else:
        logits, dc = s2vt_decode_fwd(cp, np.stack(outs, axis=0), h, s, sos_id, max_len)
    seq_alphas = np.stack(alphas, axis=1).reshape(B, N, K, K)
    cache = dict(vid=vid, x=x, z1=z1, c1=c1, c2=c2, conv_feats=conv_feats, feats=feats, pk=pk, steps=steps, dec=dc, H=H)
    return logits, seq_alphas, cache


def spatialnet_backward(p, cache, dlogits):
    """Returns the gradient of every parameter, keyed like the reference's named_parameters()."""
    vid, x, z1, c1, c2 = (cache[k] for k in ("vid", "x", "z1", "c1", "c2"))
    conv_feats, feats, pk, steps, H = (cache[k] for k in ("conv_feats", "feats", "pk", "steps", "H"))
    B, N, F, K, _ = vid.shape
    cells = K * K
    Wk, Wq = p["attention.key_layer.weight"], p["attention.query_layer.weight"]
    v = p["attention.energy_layer.weight"][0]
    cp = _caption(p)
    pre, W_ih, W_hh, _, _ = _enc_gru(cp)
    if pre == "encoder.rnn.":
        gdec, denc, dh = O.s2vtatt_decode_bwd(cp, cache["dec"], dlogits)    # d enc [B,N,H], d (final state) [B,H]
    else:
        gdec, dout1, dh = s2vt_decode_bwd(cp, cache["dec"], dlogits)
        denc = dout1.transpose(1, 0, 2)
    g = {"caption_net." + k: val for k, val in gdec.items()}
    dW_ih = np.zeros_like(W_ih); dW_hh = np.zeros_like(W_hh)
    db_ih = np.zeros(3 * H, vid.dtype); db_hh = np.zeros(3 * H, vid.dtype)
    dWq = np.zeros_like(Wq); dv = np.zeros_like(v)
    dpk = np.zeros_like(pk)
    for t in range(N - 1, -1, -1):
        st = steps[t]
        dh = dh + denc[:, t]
        r, z, n, ghn, hp = st["r"], st["z"], st["n"], st["ghn"], st["hprev"]
        dn = dh * (1.0 - z); dz = dh * (hp - n)
        dnp = dn * (1.0 - n * n); dzp = dz * z * (1.0 - z); drp = dnp * ghn * r * (1.0 - r)
        dgi = np.concatenate([drp, dzp, dnp], axis=1)
        dgh = np.concatenate([drp, dzp, dnp * r], axis=1)
        dW_ih += dgi.T @ st["ctx"]; db_ih += dgi.sum(axis=0)
        dW_hh += dgh.T @ hp; db_hh += dgh.sum(axis=0)
        dctx = dgi @ W_ih
        dq, de, _, dv_t = O.attention_bwd(dctx, st["a"], st["e"], feats[:, t], v)      # the input features need no gradient
        dpk[:, t] = de; dv += dv_t
        dWq += dq.T @ hp
        dh = dh * z + dgh @ W_hh + dq @ Wq
    for name, val in (("weight_ih_l0", dW_ih), ("weight_hh_l0", dW_hh), ("bias_ih_l0", db_ih), ("bias_hh_l0", db_hh)):
        k = "caption_net." + pre + name                    # S2VT: rnn1 also ran the decode stage (added to its gradients there)
        g[k] = g[k] + val if k in g else val
    g["attention.query_layer.weight"] = dWq
    g["attention.energy_layer.weight"] = dv[None, :]
    g["attention.key_layer.weight"] = dpk.reshape(-1, H).T @ conv_feats.reshape(-1, H)
    dconv = (dpk.reshape(-1, H) @ Wk).reshape(B, N, cells, H)
    dz2 = dconv.transpose(0, 1, 3, 2).reshape(B * N, H, K, K)
    dy2, g["conv.4.weight"], g["conv.4.bias"] = bn_relu_bwd(dz2, c2, p["conv.4.weight"])
    g["conv.3.weight"], g["conv.3.bias"], dz1 = conv3x3_bwd(z1, p["conv.3.weight"], dy2)
    dy1, g["conv.1.weight"], g["conv.1.bias"] = bn_relu_bwd(dz1, c1, p["conv.1.weight"])
    g["conv.0.weight"], g["conv.0.bias"], _ = conv3x3_bwd(x, p["conv.0.weight"], dy1, need_dx=False)
    return g


def train_iter_spatialnet(p, vid, s, s_len, sos_id, max_len):
    """train_spatial.py:30-39 + loss.backward()."""
    logits, seq_alphas, cache = spatialnet_forward_train(p, vid, s, sos_id, max_len)
    loss, dlogits, nll = O.masked_loss(logits, s, s_len)
    acc, pred = O.masked_accuracy(logits, s, s_len)
    grads = spatialnet_backward(p, cache, dlogits)
    return dict(loss=loss, acc=acc, pred=pred, logits=logits, seq_alphas=seq_alphas, grads=grads, token_nll=nll, cache=cache)
