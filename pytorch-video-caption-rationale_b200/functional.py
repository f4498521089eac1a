"""torch.autograd bindings of the model-level C-ABI entry points (include/pvcr_b200.h).

PyTorch supplies device memory, the current stream and the autograd tape; every FLOP runs in
libpvcr_b200.so.  Nothing here falls back to torch ops: without the library the calls raise.
"""
import ctypes
import itertools

import torch

from . import _lib
from ._lib import (ATT_PARAM_FIELDS, FRONT_PARAM_FIELDS, GEN_PARAM_FIELDS, S2VT_PARAM_FIELDS, PvcrDims, PvcrGenGrads,
                   PvcrGenParams, PvcrS2vtAttGrads, PvcrS2vtAttParams, PvcrS2vtGrads, PvcrS2vtParams, PvcrSpatialFrontParams,
                   check, lib, ptr, stream_ptr)

_seed_counter = itertools.count(1)


def next_seed():
    """Per-call Philox seed derived from torch's global seed (dropout / Gumbel draws inside the kernels)."""
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_seed_counter) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


def _f32c(t):
    assert t.is_cuda, "pvcr_b200 runs on CUDA tensors only (no CPU path)"
    return t.detach().to(torch.float32).contiguous()


def _i64c(t):
    assert t.is_cuda
    return t.detach().to(torch.int64).contiguous()


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def make_dims(B, N, V, H, E, L, Vc, nsplit=1, dropout_p=0.0, seed=0):
    return PvcrDims(B, N, V, H, E, L, Vc, nsplit, float(dropout_p), seed)


def _fill_struct(struct, fields, tensors):
    for f in fields:
        t = tensors.get(f)
        setattr(struct, f, None if t is None else t.data_ptr())
    return struct


def _grad_buffers(cfg, key, fields, tensors):
    """Gradient output buffers: caller-provided (cfg[key][field], e.g. views into a flat all-reduce bucket) or fresh."""
    given = cfg.get(key) or {}
    out = {}
    for f in fields:
        t = tensors[f]
        g = given.get(f)
        ok = g is not None and g.shape == t.shape and g.dtype == torch.float32 and g.is_contiguous() and g.device == t.device
        out[f] = g if ok else torch.empty_like(t)
    return out


class ManualCtx:
    """Stand-in for the autograd context when the Functions below are chained by hand (tape-free train step)."""

    def mark_non_differentiable(self, *a):
        pass


ATT_SEQ_FIELDS = [f for f in ATT_PARAM_FIELDS if f not in ("out_w", "out_b")]


class S2VTAttSequence(torch.autograd.Function):
    """Encoder GRU + attention decoder, teacher forced: (vid_feats, frame_scale, s_in, params) -> hs, alphas.

    Replaces Encoder.forward + the Decoder.forward loop of the reference (model/S2VTAttModel.py:80-96,150-196)
    up to the vocabulary projection."""

    @staticmethod
    def forward(ctx, cfg, vid, frame_scale, s_in, *params):
        B, N, V = vid.shape
        L = s_in.shape[1]
        tensors = {f: _f32c(p) for f, p in zip(ATT_SEQ_FIELDS, params)}
        H = tensors["enc_w_hh"].shape[1]
        Vc, E = tensors["emb"].shape
        dims = make_dims(B, N, V, H, E, L, Vc, cfg["nsplit"], 0.0, 0)
        vid_c = _f32c(vid)
        fs_c = None if frame_scale is None else _f32c(frame_scale)
        s_c = _i64c(s_in)
        need_fg = int(frame_scale is not None and frame_scale.requires_grad)
        Lb = lib()
        ws = _ws(Lb.pvcr_s2vtatt_workspace(ctypes.byref(dims), need_fg), vid.device)
        hs = torch.empty((B, L, H), dtype=torch.float32, device=vid.device)
        alphas = torch.empty((L, B, N), dtype=torch.float32, device=vid.device)
        ps = _fill_struct(PvcrS2vtAttParams(), ATT_SEQ_FIELDS, tensors)
        check(Lb.pvcr_s2vtatt_fwd(ctypes.byref(dims), ctypes.byref(ps), ptr(vid_c), ptr(fs_c), ptr(s_c), ptr(hs),
                                  ptr(alphas), ptr(ws), ws.numel(), stream_ptr()), "pvcr_s2vtatt_fwd")
        ctx.dims = dims
        ctx.need_fg = need_fg
        ctx.cfg = cfg
        ctx.keep = (vid_c, fs_c, s_c, hs.detach(), ws, tensors)     # detached alias: the returned hs gets grad_fn = this node (no cycle)
        ctx.mark_non_differentiable(alphas)
        return hs, alphas

    @staticmethod
    def backward(ctx, d_hs, _d_alphas):
        gen = S2VTAttSequence.backward_in_parts(ctx, d_hs, parts=(0,))
        try:
            while True:
                next(gen)
        except StopIteration as done:
            return done.value

    @staticmethod
    def backward_in_parts(ctx, d_hs, parts=(1, 2)):
        """Generator: runs the C backward part by part, yielding after every part but the last (parts=(1, 2): the
        decoder-half gradients are final at the yield); returns the autograd gradient tuple."""
        vid_c, fs_c, s_c, hs, ws, tensors = ctx.keep
        d_hs = _f32c(d_hs)
        grads = _grad_buffers(ctx.cfg, "grad_out", ATT_SEQ_FIELDS, tensors)
        d_fs = torch.empty_like(fs_c) if ctx.need_fg else None
        ps = _fill_struct(PvcrS2vtAttParams(), ATT_SEQ_FIELDS, tensors)
        gs = _fill_struct(PvcrS2vtAttGrads(), ATT_SEQ_FIELDS, grads)
        Lb = lib()
        for i, part in enumerate(parts):
            check(Lb.pvcr_s2vtatt_bwd_part(ctypes.byref(ctx.dims), ctypes.byref(ps), ptr(vid_c), ptr(fs_c), ptr(s_c),
                                           ptr(hs), ptr(d_hs), ctypes.byref(gs), ptr(d_fs), ptr(ws), ws.numel(),
                                           stream_ptr(), part), "pvcr_s2vtatt_bwd_part")
            if i + 1 < len(parts):
                yield grads
        return (None, None, d_fs, None) + tuple(grads[f] for f in ATT_SEQ_FIELDS)


class S2VTAttDecode(torch.autograd.Function):
    """Attention decoder on caller-given encoder outputs: (enc_outs [B,N,H], enc_final [B,H], s_in, params) -> hs,
    alphas.  `S2VTAttModel.decode` of the reference (model/S2VTAttModel.py:231-243), the entry SpatialNet uses
    (model/SpatialNet.py:140); differentiable in enc_outs, enc_final and every decoder parameter."""

    @staticmethod
    def forward(ctx, cfg, enc_outs, enc_final, s_in, *params):
        B, N, H = enc_outs.shape
        L = s_in.shape[1]
        tensors = {f: _f32c(p) for f, p in zip(ATT_SEQ_FIELDS, params)}
        Vc, E = tensors["emb"].shape
        dims = make_dims(B, N, 1, H, E, L, Vc, cfg["nsplit"], 0.0, 0)
        enc_c, fin_c, s_c = _f32c(enc_outs), _f32c(enc_final), _i64c(s_in)
        Lb = lib()
        ws = _ws(Lb.pvcr_s2vtatt_workspace(ctypes.byref(dims), 0), enc_c.device)
        hs = torch.empty((B, L, H), dtype=torch.float32, device=enc_c.device)
        alphas = torch.empty((L, B, N), dtype=torch.float32, device=enc_c.device)
        ps = _fill_struct(PvcrS2vtAttParams(), ATT_SEQ_FIELDS, tensors)
        check(Lb.pvcr_s2vtatt_decode_fwd(ctypes.byref(dims), ctypes.byref(ps), ptr(enc_c), ptr(fin_c), ptr(s_c), ptr(hs),
                                         ptr(alphas), ptr(ws), ws.numel(), stream_ptr()), "pvcr_s2vtatt_decode_fwd")
        ctx.dims = dims
        ctx.keep = (s_c, hs.detach(), ws, tensors)     # detached alias: the returned hs gets grad_fn = this node (no cycle)
        ctx.mark_non_differentiable(alphas)
        return hs, alphas

    @staticmethod
    def backward(ctx, d_hs, _d_alphas):
        s_c, hs, ws, tensors = ctx.keep
        B, N, H = ctx.dims.B, ctx.dims.N, ctx.dims.H
        dec_fields = [f for f in ATT_SEQ_FIELDS if not f.startswith("enc_")]
        grads = {f: torch.empty_like(tensors[f]) for f in dec_fields}
        d_enc = torch.empty((B, N, H), dtype=torch.float32, device=hs.device)
        d_fin = torch.empty((B, H), dtype=torch.float32, device=hs.device)
        ps = _fill_struct(PvcrS2vtAttParams(), ATT_SEQ_FIELDS, tensors)
        gs = _fill_struct(PvcrS2vtAttGrads(), ATT_SEQ_FIELDS, grads)
        check(lib().pvcr_s2vtatt_decode_bwd(ctypes.byref(ctx.dims), ctypes.byref(ps), ptr(s_c), ptr(hs), ptr(_f32c(d_hs)),
                                            ctypes.byref(gs), ptr(d_enc), ptr(d_fin), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_s2vtatt_decode_bwd")
        return (None, d_enc, d_fin, None) + tuple(grads.get(f) for f in ATT_SEQ_FIELDS)


class GruStep(torch.autograd.Function):
    """h' = GRU(x, h_prev), one step: `encode_step` of both caption nets (model/S2VTAttModel.py:63-78,
    model/S2VTModel.py:57-72), called once per frame by SpatialNet (model/SpatialNet.py:127)."""

    @staticmethod
    def forward(ctx, nsplit, x, h_prev, w_ih, w_hh, b_ih, b_hh):
        B, V = x.shape
        H = w_hh.shape[1]
        x_c = _f32c(x)
        h_c = None if h_prev is None else _f32c(h_prev)
        w = tuple(_f32c(t) for t in (w_ih, w_hh, b_ih, b_hh))
        Lb = lib()
        ws = _ws(Lb.pvcr_gru_step_workspace(B, V, H, nsplit), x_c.device)
        h_out = torch.empty((B, H), dtype=torch.float32, device=x_c.device)
        saved = torch.empty((4, B, H), dtype=torch.float32, device=x_c.device)
        check(Lb.pvcr_gru_step_fwd(ptr(x_c), ptr(h_c), ptr(w[0]), ptr(w[1]), ptr(w[2]), ptr(w[3]), B, V, H, nsplit,
                                   ptr(h_out), ptr(saved), ptr(ws), ws.numel(), stream_ptr()), "pvcr_gru_step_fwd")
        ctx.meta = (B, V, H, nsplit)
        ctx.keep = (x_c, h_c, w, saved)
        return h_out

    @staticmethod
    def backward(ctx, d_h):
        B, V, H, nsplit = ctx.meta
        x_c, h_c, w, saved = ctx.keep
        dev = x_c.device
        d_x = torch.empty_like(x_c)
        d_hp = None if h_c is None else torch.empty_like(h_c)
        d_w = tuple(torch.empty_like(t) for t in w)
        Lb = lib()
        ws = _ws(Lb.pvcr_gru_step_workspace(B, V, H, nsplit), dev)
        check(Lb.pvcr_gru_step_bwd(ptr(_f32c(d_h)), ptr(x_c), ptr(h_c), ptr(w[0]), ptr(w[1]), ptr(saved), B, V, H, nsplit,
                                   ptr(d_x), ptr(d_hp), ptr(d_w[0]), ptr(d_w[1]), ptr(d_w[2]), ptr(d_w[3]), 0, ptr(ws),
                                   ws.numel(), stream_ptr()), "pvcr_gru_step_bwd")
        return (None, d_x, d_hp) + d_w


S2VT_SEQ_FIELDS = [f for f in S2VT_PARAM_FIELDS if f not in ("out_w", "out_b")]


class S2VTSequence(torch.autograd.Function):
    """S2VT encode + decode with given input words: (vid_feats, frame_scale, s_in, params) -> hs [B,L,H]
    (model/S2VTModel.py:74-145 up to, but excluding, the vocabulary projection)."""

    @staticmethod
    def forward(ctx, cfg, vid, frame_scale, s_in, *params):
        B, N, V = vid.shape
        L = s_in.shape[1]
        tensors = {f: _f32c(p) for f, p in zip(S2VT_SEQ_FIELDS, params)}
        H = tensors["rnn1_w_hh"].shape[1]
        Vc, E = tensors["emb"].shape
        dims = make_dims(B, N, V, H, E, L, Vc, cfg["nsplit"], cfg.get("emb_dropout_p", 0.0), cfg.get("seed", 0))
        vid_c = _f32c(vid)
        fs_c = None if frame_scale is None else _f32c(frame_scale)
        s_c = _i64c(s_in)
        need_fg = int(frame_scale is not None and frame_scale.requires_grad)
        Lb = lib()
        ws = _ws(Lb.pvcr_s2vt_workspace(ctypes.byref(dims), need_fg), vid.device)
        hs = torch.empty((B, L, H), dtype=torch.float32, device=vid.device)
        ps = _fill_struct(PvcrS2vtParams(), S2VT_SEQ_FIELDS, tensors)
        check(Lb.pvcr_s2vt_fwd(ctypes.byref(dims), ctypes.byref(ps), ptr(vid_c), ptr(fs_c), ptr(s_c), ptr(hs), ptr(ws),
                               ws.numel(), stream_ptr()), "pvcr_s2vt_fwd")
        ctx.dims = dims
        ctx.need_fg = need_fg
        ctx.cfg = cfg
        ctx.keep = (vid_c, fs_c, s_c, hs.detach(), ws, tensors)     # detached alias: the returned hs gets grad_fn = this node (no cycle)
        return hs

    @staticmethod
    def backward(ctx, d_hs):
        vid_c, fs_c, s_c, hs, ws, tensors = ctx.keep
        d_hs = _f32c(d_hs)
        grads = _grad_buffers(ctx.cfg, "grad_out", S2VT_SEQ_FIELDS, tensors)
        d_fs = torch.empty_like(fs_c) if ctx.need_fg else None
        ps = _fill_struct(PvcrS2vtParams(), S2VT_SEQ_FIELDS, tensors)
        gs = _fill_struct(PvcrS2vtGrads(), S2VT_SEQ_FIELDS, grads)
        Lb = lib()
        check(Lb.pvcr_s2vt_bwd(ctypes.byref(ctx.dims), ctypes.byref(ps), ptr(vid_c), ptr(fs_c), ptr(s_c), ptr(hs),
                               ptr(d_hs), ctypes.byref(gs), ptr(d_fs), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_s2vt_bwd")
        return (None, None, d_fs, None) + tuple(grads[f] for f in S2VT_SEQ_FIELDS)


class S2VTDecode(torch.autograd.Function):
    """S2VT decode on caller-given rnn1 outputs: (out1 [B,N,H], state1 [B,H], s_in, params) -> hs [B,L,H]
    (model/S2VTModel.py:88-145 up to the vocabulary projection; SpatialNet.py:140)."""

    @staticmethod
    def forward(ctx, cfg, out1, state1, s_in, *params):
        B, N, H = out1.shape
        L = s_in.shape[1]
        tensors = {f: _f32c(p) for f, p in zip(S2VT_SEQ_FIELDS, params)}
        Vc, E = tensors["emb"].shape
        dims = make_dims(B, N, 1, H, E, L, Vc, cfg["nsplit"], cfg.get("emb_dropout_p", 0.0), cfg.get("seed", 0))
        o_c, st_c, s_c = _f32c(out1), _f32c(state1), _i64c(s_in)
        Lb = lib()
        ws = _ws(Lb.pvcr_s2vt_workspace(ctypes.byref(dims), 0), o_c.device)
        hs = torch.empty((B, L, H), dtype=torch.float32, device=o_c.device)
        ps = _fill_struct(PvcrS2vtParams(), S2VT_SEQ_FIELDS, tensors)
        check(Lb.pvcr_s2vt_decode_fwd(ctypes.byref(dims), ctypes.byref(ps), ptr(o_c), ptr(st_c), ptr(s_c), ptr(hs), ptr(ws),
                                      ws.numel(), stream_ptr()), "pvcr_s2vt_decode_fwd")
        ctx.dims = dims
        ctx.keep = (s_c, hs.detach(), ws, tensors)     # detached alias: the returned hs gets grad_fn = this node (no cycle)
        return hs

    @staticmethod
    def backward(ctx, d_hs):
        s_c, hs, ws, tensors = ctx.keep
        B, N, H = ctx.dims.B, ctx.dims.N, ctx.dims.H
        fields = [f for f in S2VT_SEQ_FIELDS if f != "rnn1_w_ih"]
        grads = {f: torch.empty_like(tensors[f]) for f in fields}
        d_out1 = torch.empty((B, N, H), dtype=torch.float32, device=hs.device)
        d_state1 = torch.empty((B, H), dtype=torch.float32, device=hs.device)
        ps = _fill_struct(PvcrS2vtParams(), S2VT_SEQ_FIELDS, tensors)
        gs = _fill_struct(PvcrS2vtGrads(), S2VT_SEQ_FIELDS, grads)
        check(lib().pvcr_s2vt_decode_bwd(ctypes.byref(ctx.dims), ctypes.byref(ps), ptr(s_c), ptr(hs), ptr(_f32c(d_hs)),
                                         ctypes.byref(gs), ptr(d_out1), ptr(d_state1), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_s2vt_decode_bwd")
        return (None, d_out1, d_state1, None) + tuple(grads.get(f) for f in S2VT_SEQ_FIELDS)


class GeneratorSelect(torch.autograd.Function):
    """RationaleNet Generator (model/RationaleNet.py:32-54): (vid_feats, noise?, params) -> probs [B,N,2],
    p1 [B,N] (the frame scale handed to the caption network instead of materialising vid_feats * p1) and
    pen [2] = (calc_brevity_loss(probs), calc_cont_loss(probs))."""

    @staticmethod
    def forward(ctx, cfg, vid, noise, *params):
        B, N, V = vid.shape
        tensors = {f: _f32c(p) for f, p in zip(GEN_PARAM_FIELDS, params)}
        H = tensors["w_hh"].shape[1]
        dims = make_dims(B, N, V, H, 1, 1, 1, cfg["nsplit"], cfg.get("dropout_p", 0.0), cfg.get("seed", 0))
        vid_c = _f32c(vid)
        noise_c = None if noise is None else _f32c(noise)
        Lb = lib()
        ws = _ws(Lb.pvcr_generator_workspace(ctypes.byref(dims)), vid.device)
        probs = torch.empty((B, N, 2), dtype=torch.float32, device=vid.device)
        p1 = torch.empty((B, N), dtype=torch.float32, device=vid.device)
        pen = torch.empty((2,), dtype=torch.float32, device=vid.device)
        ps = _fill_struct(PvcrGenParams(), GEN_PARAM_FIELDS, tensors)
        check(Lb.pvcr_generator_fwd(ctypes.byref(dims), ctypes.byref(ps), ptr(vid_c), ptr(noise_c), float(cfg["tau"]),
                                    int(cfg["hard"]), ptr(probs), ptr(p1), ptr(pen), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_generator_fwd")
        ctx.dims, ctx.tau = dims, float(cfg["tau"])
        ctx.keep = (vid_c, ws, tensors)
        return probs, p1, pen

    @staticmethod
    def backward(ctx, d_probs, d_p1, d_pen, grad_out=None):
        """grad_out (tape-free callers only): {field: existing gradient buffer} written in place where shape / dtype
        match (views into a flat all-reduce bucket)."""
        vid_c, ws, tensors = ctx.keep
        dev = vid_c.device
        H4, V = tensors["w_ih"].shape
        grad_out = grad_out or {}

        def usable(f):
            g = grad_out.get(f)
            t = tensors[f]
            return g is not None and g.shape == t.shape and g.dtype == torch.float32 and g.is_contiguous() and g.device == t.device

        grads = {f: (grad_out[f] if usable(f) else torch.empty_like(t)) for f, t in tensors.items()
                 if f not in ("w_ih", "w_ih_r")}
        if usable("w_ih") and usable("w_ih_r"):       # adjacent or not: the C side checks and issues one or two GEMMs
            grads["w_ih"], grads["w_ih_r"] = grad_out["w_ih"], grad_out["w_ih_r"]
        else:
            wih_cat = torch.empty((2 * H4, V), dtype=torch.float32, device=dev)       # both directions: one GEMM
            grads["w_ih"], grads["w_ih_r"] = wih_cat[:H4], wih_cat[H4:]
        ps = _fill_struct(PvcrGenParams(), GEN_PARAM_FIELDS, tensors)
        gs = _fill_struct(PvcrGenGrads(), GEN_PARAM_FIELDS, grads)
        d_probs = None if d_probs is None else _f32c(d_probs)
        d_p1 = None if d_p1 is None else _f32c(d_p1)
        d_pen = None if d_pen is None else _f32c(d_pen)
        Lb = lib()
        check(Lb.pvcr_generator_bwd(ctypes.byref(ctx.dims), ctypes.byref(ps), ptr(vid_c), ctx.tau, ptr(d_p1),
                                    ptr(d_probs), ptr(d_pen), ctypes.byref(gs), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_generator_bwd")
        return (None, None, None) + tuple(grads[f] for f in GEN_PARAM_FIELDS)


class VocabCrossEntropy(torch.autograd.Function):
    """Dropout + Linear(H -> Vc) fused with calc_masked_loss / calc_masked_accuracy / argmax
    (model/S2VTAttModel.py:145, train_utils.py:37-71, train.py:38): (hs, W, b, target, s_len) -> loss, stats, pred."""

    @staticmethod
    def forward(ctx, cfg, hs, out_w, out_b, target, s_len):
        B, L, H = hs.shape
        Vc = out_w.shape[0]
        hs_c, w_c, b_c = _f32c(hs), _f32c(out_w), _f32c(out_b)
        t_c, l_c = _i64c(target), _i64c(s_len)
        Lb = lib()
        nsplit, p, seed = cfg["nsplit"], float(cfg.get("dropout_p", 0.0)), int(cfg.get("seed", 0))
        ws = cfg.get("vocab_ws")              # pre-allocated (and possibly pre-staged: vocab_prepare) by the caller
        if ws is None:
            ws = _ws(Lb.pvcr_vocab_ce_workspace(B, L, H, Vc, nsplit, p), hs.device)
        loss3 = torch.empty(3, dtype=torch.float32, device=hs.device)
        pred = torch.empty((B, L), dtype=torch.int64, device=hs.device)
        lse = torch.empty((B, L), dtype=torch.float32, device=hs.device)
        nll = torch.empty((B, L), dtype=torch.float32, device=hs.device)
        check(Lb.pvcr_vocab_ce_fwd(ptr(hs_c), ptr(w_c), ptr(b_c), ptr(t_c), ptr(l_c), B, L, H, Vc, nsplit, p, seed,
                                   ptr(loss3), ptr(pred), ptr(lse), ptr(nll), None, 0, ptr(ws), ws.numel(),
                                   stream_ptr()), "pvcr_vocab_ce_fwd")
        cfg["token_nll"] = nll             # unmasked per-token loss [B,L] (criterion(logits, target), train_utils.py:47-48)
        ctx.cfg = (B, L, H, Vc, nsplit, p, seed)
        ctx.grad_out = cfg.get("vocab_grad_out") or {}
        ctx.keep = (hs_c, w_c, b_c, t_c, l_c, ws, lse, pred)
        ctx.mark_non_differentiable(pred)
        stats = loss3[1:].clone()
        ctx.mark_non_differentiable(stats)
        return loss3[0].clone(), stats, pred

    @staticmethod
    def backward(ctx, d_loss, _d_stats, _d_pred):
        B, L, H, Vc, nsplit, p, seed = ctx.cfg
        hs_c, w_c, b_c, t_c, l_c, ws, lse, pred = ctx.keep
        gscale = _f32c(d_loss).reshape(1)
        d_hs = torch.empty_like(hs_c)
        gb = _grad_buffers({"g": ctx.grad_out}, "g", ("out_w", "out_b"), {"out_w": w_c, "out_b": b_c})
        d_w, d_b = gb["out_w"], gb["out_b"]
        Lb = lib()
        check(Lb.pvcr_vocab_ce_bwd(ptr(hs_c), ptr(w_c), ptr(b_c), ptr(t_c), ptr(l_c), B, L, H, Vc, nsplit, p, seed, ptr(gscale),
                                   ptr(d_hs), ptr(d_w), ptr(d_b), ptr(lse), ptr(pred), ptr(ws), ws.numel(),
                                   stream_ptr()), "pvcr_vocab_ce_bwd")
        return None, d_hs, d_w, d_b, None, None


def vocab_prepare(cfg, B, L, out_w):
    """Allocate the workspace of the coming VocabCrossEntropy.forward and start staging out_w on a side lane of the
    library (pvcr_vocab_ce_prepare); the workspace travels in cfg["vocab_ws"]."""
    Vc, H = out_w.shape
    Lb = lib()
    nsplit, p = cfg["nsplit"], float(cfg.get("dropout_p", 0.0))
    ws = _ws(Lb.pvcr_vocab_ce_workspace(B, L, H, Vc, nsplit, p), out_w.device)
    check(Lb.pvcr_vocab_ce_prepare(ptr(_f32c(out_w)), B, L, H, Vc, nsplit, ptr(ws), ws.numel(), stream_ptr()),
          "pvcr_vocab_ce_prepare")
    cfg["vocab_ws"] = ws


class VocabLogits(torch.autograd.Function):
    """logits = Dropout(hs) W^T + b, materialised (the reference module API returns the [B,L,Vc] tensor)."""

    @staticmethod
    def forward(ctx, cfg, hs, out_w, out_b):
        B, L, H = hs.shape
        Vc = out_w.shape[0]
        hs_c, w_c, b_c = _f32c(hs), _f32c(out_w), _f32c(out_b)
        nsplit, p, seed = cfg["nsplit"], float(cfg.get("dropout_p", 0.0)), int(cfg.get("seed", 0))
        Lb = lib()
        ws = _ws(Lb.pvcr_vocab_ce_workspace(B, L, H, Vc, nsplit, p), hs.device)
        logits = torch.empty((B, L, Vc), dtype=torch.float32, device=hs.device)
        check(Lb.pvcr_vocab_ce_fwd(ptr(hs_c), ptr(w_c), ptr(b_c), None, None, B, L, H, Vc, nsplit, p, seed, None, None,
                                   None, None, ptr(logits), Vc, ptr(ws), ws.numel(), stream_ptr()), "pvcr_vocab_ce_fwd")
        ctx.cfg = (B, L, H, Vc, nsplit, p, seed)
        ctx.keep = (hs_c, w_c)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        B, L, H, Vc, nsplit, p, seed = ctx.cfg
        hs_c, w_c = ctx.keep
        d_logits = _f32c(d_logits).view(B * L, Vc)
        Lb = lib()
        d_hs = torch.empty_like(hs_c)
        d_w = torch.empty_like(w_c)
        d_b = torch.empty((Vc,), dtype=torch.float32, device=hs_c.device)
        ws = _ws(Lb.pvcr_linear_bwd_workspace(B * L, Vc, H, nsplit), hs_c.device)
        x = hs_c
        if p > 0.0:
            # the Linear saw Dropout(hs): regenerate the forward's mask from (p, seed) for d W = d logits^T Dropout(hs) ...
            x = torch.empty_like(hs_c)
            check(Lb.pvcr_out_dropout_apply(ptr(hs_c), ptr(x), hs_c.numel(), p, seed, stream_ptr()), "pvcr_out_dropout_apply")
        check(Lb.pvcr_linear_bwd(ptr(d_logits), Vc, ptr(x), H, ptr(w_c), H, ptr(d_hs), H, ptr(d_w), H, ptr(d_b),
                                 B * L, Vc, H, nsplit, 0, ptr(ws), ws.numel(), stream_ptr()), "pvcr_linear_bwd")
        if p > 0.0:        # ... and d hs = Dropout'(d logits W)
            check(Lb.pvcr_out_dropout_apply(ptr(d_hs), ptr(d_hs), d_hs.numel(), p, seed, stream_ptr()), "pvcr_out_dropout_apply")
        return None, d_hs, d_w, d_b


class DecodePlan:
    """Workspace of pvcr_s2vtatt_greedy_ex kept between calls, with the fingerprint of what it was prepared from: the staged
    weight planes and the word table T[w] = W_e Emb[w] + b_ih are parameter-only work, so an eval loop that decodes batch
    after batch pays for them once.  The fingerprint is (shape, storage address and autograd version of every parameter
    tensor): in-place updates (optimizer steps, load_state_dict) change the version, re-allocated parameters the address.
    Writes through ``.data`` are invisible to it -- call ``invalidate()`` after those."""

    def __init__(self):
        self.ws = None
        self.key = None

    def invalidate(self):
        self.key = None

    def lookup(self, key, nbytes, device):
        """-> (workspace, reuse flag)."""
        if self.ws is None or self.ws.numel() != int(nbytes) or self.ws.device != device:
            self.ws, self.key = _ws(nbytes, device), None
        reuse = self.key == key
        self.key = key
        return self.ws, reuse


def _fingerprint(dims_tuple, srcs):
    return dims_tuple + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in srcs)


def s2vtatt_greedy(vid, frame_scale, sos_id, max_len, seq_params, out_w, out_b, nsplit=3, plan=None, return_logits=True,
                   enc_outs=None, enc_final=None):
    """-> (ids [B,L] int64, logits [B,L,Vc] or None, alphas [L,B,N]); eval branch of S2VTAttModel (from the frames, or from
    caller-given encoder outputs [B,N,H] / final state [B,H]: decode() in eval mode).  ``plan``: a DecodePlan that keeps the
    prepared weights between calls."""
    given = enc_outs is not None
    srcs = list(seq_params) + [out_w, out_b]
    tensors = {f: _f32c(p) for f, p in zip(ATT_SEQ_FIELDS, seq_params)}
    tensors["out_w"], tensors["out_b"] = _f32c(out_w), _f32c(out_b)
    H = tensors["enc_w_hh"].shape[1]
    Vc, E = tensors["emb"].shape
    if given:
        B, N, _ = enc_outs.shape
        V = 1
        a_c, b_c = _f32c(enc_outs), _f32c(enc_final)
        dev = a_c.device
    else:
        B, N, V = vid.shape
        a_c = _f32c(vid)
        b_c = None if frame_scale is None else _f32c(frame_scale)
        dev = a_c.device
    dims = make_dims(B, N, V, H, E, max_len, Vc, nsplit, 0.0, 0)
    Lb = lib()
    nbytes = Lb.pvcr_s2vtatt_greedy_workspace(ctypes.byref(dims))
    flags = 0
    zero_copy = all(tensors[f].data_ptr() == p.data_ptr() for f, p in zip(ATT_SEQ_FIELDS, seq_params)) and \
        tensors["out_w"].data_ptr() == out_w.data_ptr() and tensors["out_b"].data_ptr() == out_b.data_ptr()
    if plan is not None and zero_copy:                 # (a converted copy has no stable identity: prepare every call)
        ws, reuse = plan.lookup(_fingerprint((B, N, V, H, E, max_len, Vc, nsplit, given), srcs), nbytes, dev)
        flags = 1 if reuse else 0                      # PVCR_DECODE_REUSE_PREPARED
    else:
        ws = _ws(nbytes, dev)
    ids = torch.empty((B, max_len), dtype=torch.int64, device=dev)
    logits = torch.empty((B, max_len, Vc), dtype=torch.float32, device=dev) if return_logits else None
    alphas = torch.empty((max_len, B, N), dtype=torch.float32, device=dev)
    ps = _fill_struct(PvcrS2vtAttParams(), ATT_PARAM_FIELDS, tensors)
    check(Lb.pvcr_s2vtatt_greedy_ex(ctypes.byref(dims), ctypes.byref(ps), None if given else ptr(a_c),
                                    None if given else ptr(b_c), ptr(a_c) if given else None, ptr(b_c) if given else None,
                                    int(sos_id), ptr(ids), ptr(logits), ptr(alphas), ptr(ws), ws.numel(), flags,
                                    stream_ptr()), "pvcr_s2vtatt_greedy_ex")
    return ids, logits, alphas


def s2vtatt_decode_greedy(enc_outs, enc_final, sos_id, max_len, seq_params, out_w, out_b, nsplit=3, plan=None,
                          return_logits=True):
    """Greedy decoding from caller-given encoder outputs [B,N,H] / final state [B,H] -> (ids [B,L], logits [B,L,Vc],
    alphas [L,B,N]); eval branch of S2VTAttModel.decode (model/S2VTAttModel.py:231-243)."""
    return s2vtatt_greedy(None, None, sos_id, max_len, seq_params, out_w, out_b, nsplit, plan, return_logits,
                          enc_outs=enc_outs, enc_final=enc_final)


def s2vt_decode_greedy(out1, state1, sos_id, max_len, seq_params, out_w, out_b, nsplit=3):
    """Greedy decoding from caller-given rnn1 outputs [B,N,H] / state [B,H] -> (ids [B,L], logits [B,L,Vc]); eval branch
    of S2VTModel.decode (model/S2VTModel.py:147-177)."""
    B, N, H = out1.shape
    tensors = {f: _f32c(p) for f, p in zip(S2VT_SEQ_FIELDS, seq_params)}
    tensors["out_w"], tensors["out_b"] = _f32c(out_w), _f32c(out_b)
    Vc, E = tensors["emb"].shape
    dims = make_dims(B, N, 1, H, E, max_len, Vc, nsplit, 0.0, 0)
    o_c, st_c = _f32c(out1), _f32c(state1)
    Lb = lib()
    ws = _ws(Lb.pvcr_s2vt_decode_steps_workspace(ctypes.byref(dims)), o_c.device)
    ids = torch.empty((B, max_len), dtype=torch.int64, device=o_c.device)
    logits = torch.empty((B, max_len, Vc), dtype=torch.float32, device=o_c.device)
    ps = _fill_struct(PvcrS2vtParams(), S2VT_PARAM_FIELDS, tensors)
    check(Lb.pvcr_s2vt_decode_greedy(ctypes.byref(dims), ctypes.byref(ps), ptr(o_c), ptr(st_c), int(sos_id), ptr(ids),
                                     ptr(logits), ptr(ws), ws.numel(), stream_ptr()), "pvcr_s2vt_decode_greedy")
    return ids, logits


def s2vtatt_beam(vid, frame_scale, sos_id, max_len, beam, seq_params, out_w, out_b, nsplit=3):
    """Fixed-length beam search (pvcr_s2vtatt_beam) -> (ids [B,beam,L] int64, best first; scores [B,beam])."""
    B, N, V = vid.shape
    tensors = {f: _f32c(p) for f, p in zip(ATT_SEQ_FIELDS, seq_params)}
    tensors["out_w"], tensors["out_b"] = _f32c(out_w), _f32c(out_b)
    H = tensors["enc_w_hh"].shape[1]
    Vc, E = tensors["emb"].shape
    dims = make_dims(B, N, V, H, E, max_len, Vc, nsplit, 0.0, 0)
    vid_c = _f32c(vid)
    fs_c = None if frame_scale is None else _f32c(frame_scale)
    Lb = lib()
    ws = _ws(Lb.pvcr_s2vtatt_beam_workspace(ctypes.byref(dims), int(beam)), vid.device)
    ids = torch.empty((B, beam, max_len), dtype=torch.int64, device=vid.device)
    scores = torch.empty((B, beam), dtype=torch.float32, device=vid.device)
    ps = _fill_struct(PvcrS2vtAttParams(), ATT_PARAM_FIELDS, tensors)
    check(Lb.pvcr_s2vtatt_beam(ctypes.byref(dims), ctypes.byref(ps), ptr(vid_c), ptr(fs_c), int(sos_id), int(beam),
                               ptr(ids), ptr(scores), ptr(ws), ws.numel(), stream_ptr()), "pvcr_s2vtatt_beam")
    return ids, scores


def s2vt_decode_steps(vid, frame_scale, sos_id, max_len, seq_params, out_w, out_b, cfg, teacher_words=None,
                      teacher_mask=None, want_logits=True):
    """-> (ids [B,L] arg-max of every step, logits [B,L,Vc] or None, fed [B,L] words fed to every step)."""
    B, N, V = vid.shape
    tensors = {f: _f32c(p) for f, p in zip(S2VT_SEQ_FIELDS, seq_params)}
    tensors["out_w"], tensors["out_b"] = _f32c(out_w), _f32c(out_b)
    H = tensors["rnn1_w_hh"].shape[1]
    Vc, E = tensors["emb"].shape
    dims = make_dims(B, N, V, H, E, max_len, Vc, cfg["nsplit"], cfg.get("emb_dropout_p", 0.0), cfg.get("seed", 0))
    vid_c = _f32c(vid)
    fs_c = None if frame_scale is None else _f32c(frame_scale)
    tw = None if teacher_words is None else _i64c(teacher_words)
    mask = None if teacher_mask is None else (ctypes.c_int32 * max_len)(*[int(bool(t)) for t in teacher_mask])
    Lb = lib()
    ws = _ws(Lb.pvcr_s2vt_decode_steps_workspace(ctypes.byref(dims)), vid.device)
    ids = torch.empty((B, max_len), dtype=torch.int64, device=vid.device)
    fed = torch.empty((B, max_len), dtype=torch.int64, device=vid.device)
    logits = torch.empty((B, max_len, Vc), dtype=torch.float32, device=vid.device) if want_logits else None
    ps = _fill_struct(PvcrS2vtParams(), S2VT_PARAM_FIELDS, tensors)
    check(Lb.pvcr_s2vt_decode_steps(ctypes.byref(dims), ctypes.byref(ps), ptr(vid_c), ptr(fs_c), int(sos_id), ptr(tw),
                                    mask, float(cfg.get("dropout_p", 0.0)), ptr(ids), ptr(fed), ptr(logits), ptr(ws),
                                    ws.numel(), stream_ptr()), "pvcr_s2vt_decode_steps")
    return ids, logits, fed


class Linear(torch.autograd.Function):
    """y = x W^T (+ b) on the tcgen05 GEMM with hand-written gradients (pvcr_linear_fwd / pvcr_linear_bwd): the nn.Linear
    layers of SpatialNet's attention (key_layer / query_layer, model/SpatialNet.py:23-25,39-41)."""

    @staticmethod
    def forward(ctx, nsplit, x, w, b=None):
        x_c, w_c = _f32c(x), _f32c(w)
        b_c = None if b is None else _f32c(b)
        M, K = x_c.shape
        N = w_c.shape[0]
        Lb = lib()
        ws = _ws(Lb.pvcr_linear_fwd_workspace(M, N, K, nsplit), x_c.device)
        y = torch.empty((M, N), dtype=torch.float32, device=x_c.device)
        check(Lb.pvcr_linear_fwd(ptr(x_c), K, ptr(w_c), K, ptr(b_c), ptr(y), N, M, N, K, nsplit, ptr(ws), ws.numel(),
                                 stream_ptr()), "pvcr_linear_fwd")
        ctx.meta = (M, N, K, nsplit, b is not None)
        ctx.keep = (x_c, w_c)
        return y

    @staticmethod
    def backward(ctx, dy):
        M, N, K, nsplit, has_b = ctx.meta
        x_c, w_c = ctx.keep
        dy_c = _f32c(dy)
        need_dx = ctx.needs_input_grad[1]
        dx = torch.empty_like(x_c) if need_dx else None
        dw = torch.empty_like(w_c)
        db = torch.empty((N,), dtype=torch.float32, device=x_c.device) if has_b else None
        Lb = lib()
        ws = _ws(Lb.pvcr_linear_bwd_workspace(M, N, K, nsplit), x_c.device)
        check(Lb.pvcr_linear_bwd(ptr(dy_c), N, ptr(x_c), K, ptr(w_c), K, ptr(dx), K, ptr(dw), K, ptr(db), M, N, K, nsplit, 0,
                                 ptr(ws), ws.numel(), stream_ptr()), "pvcr_linear_bwd")
        return None, dx, dw, db


class SpatialFront(torch.autograd.Function):
    """Two Conv3x3 + BatchNorm2d + ReLU blocks on the grid features of every frame (model/SpatialNet.py:76-86,106):
    (vid [I,F,K,K], 8 parameter tensors, 4 running-stat buffers) -> conv_feats [I*K*K, H] (differentiable in the
    parameters), feats_cl [I*K*K, F] (the input in channels-last row order).  Running estimates are updated in place in
    training, as torch.nn.BatchNorm2d does."""

    @staticmethod
    def forward(ctx, cfg, vid, rm1, rv1, rm2, rv2, *params):
        I, F, K, _ = vid.shape
        tensors = {f: _f32c(p) for f, p in zip(FRONT_PARAM_FIELDS, params)}
        H = tensors["conv1_w"].shape[0]
        vid_c = _f32c(vid)
        nsplit, training = int(cfg["nsplit"]), int(cfg["training"])
        Lb = lib()
        ws = _ws(Lb.pvcr_spatial_front_workspace(I, K, F, H, nsplit), vid.device)
        conv_feats = torch.empty((I * K * K, H), dtype=torch.float32, device=vid.device)
        feats_cl = torch.empty((I * K * K, F), dtype=torch.float32, device=vid.device)
        ps = _fill_struct(PvcrSpatialFrontParams(), FRONT_PARAM_FIELDS, tensors)
        check(Lb.pvcr_spatial_front_fwd(I, K, F, H, nsplit, ptr(vid_c), ctypes.byref(ps), ptr(rm1), ptr(rv1), ptr(rm2), ptr(rv2),
                                        training, float(cfg.get("eps", 1e-5)), float(cfg.get("momentum", 0.1)),
                                        ptr(conv_feats), ptr(feats_cl), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_spatial_front_fwd")
        ctx.meta = (I, K, F, H, nsplit, training)
        ctx.keep = (ws, tensors)
        ctx.mark_non_differentiable(feats_cl)
        return conv_feats, feats_cl

    @staticmethod
    def backward(ctx, d_conv_feats, _d_feats):
        I, K, F, H, nsplit, training = ctx.meta
        ws, tensors = ctx.keep
        grads = {f: torch.empty_like(t) for f, t in tensors.items()}
        ps = _fill_struct(PvcrSpatialFrontParams(), FRONT_PARAM_FIELDS, tensors)
        gs = _fill_struct(PvcrSpatialFrontParams(), FRONT_PARAM_FIELDS, grads)
        check(lib().pvcr_spatial_front_bwd(I, K, F, H, nsplit, ctypes.byref(ps), training, ptr(_f32c(d_conv_feats)),
                                           ctypes.byref(gs), ptr(ws), ws.numel(), stream_ptr()), "pvcr_spatial_front_bwd")
        return (None, None, None, None, None, None) + tuple(grads[f] for f in FRONT_PARAM_FIELDS)


class SpatialAttnStep(torch.autograd.Function):
    """One frame's attention over the K*K cells (model/SpatialNet.py:27-53): (q [B,H], proj_key [B,Kc,H], feats [B,Kc,F],
    v [H]) -> context [B,F], alphas [B,Kc].  Differentiable in q, proj_key and v (the features are inputs)."""

    @staticmethod
    def forward(ctx, q, pk, feats, v):
        B, Kc, H = pk.shape
        Fv = feats.shape[2]
        q_c, pk_c, v_c = _f32c(q), _f32c(pk), _f32c(v).reshape(-1)
        # the features (inputs, no gradient) may be a batch-strided view -- one frame of a [B, N, Kc, F] tensor: no frame-major copy
        if feats.dtype == torch.float32 and feats.stride(2) == 1 and feats.stride(1) == Fv and feats.stride(0) >= Kc * Fv:
            f_c, f_bs = feats.detach(), feats.stride(0)
        else:
            f_c, f_bs = _f32c(feats), Kc * Fv
        alpha = torch.empty((B, Kc), dtype=torch.float32, device=q_c.device)
        out = torch.empty((B, Fv), dtype=torch.float32, device=q_c.device)
        check(lib().pvcr_spatial_attn_fwd(B, Kc, H, Fv, ptr(q_c), H, ptr(pk_c), Kc * H, ptr(f_c), f_bs, ptr(v_c), ptr(alpha),
                                          ptr(out), stream_ptr()), "pvcr_spatial_attn_fwd")
        ctx.meta = (B, Kc, H, Fv, tuple(v.shape), f_bs)
        ctx.keep = (q_c, pk_c, f_c, v_c, alpha)
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, dctx, _dalpha):
        B, Kc, H, Fv, vshape, f_bs = ctx.meta
        q_c, pk_c, f_c, v_c, alpha = ctx.keep
        dq = torch.empty_like(q_c)
        dpk = torch.empty_like(pk_c)
        dv_part = torch.empty((B, H), dtype=torch.float32, device=q_c.device)
        check(lib().pvcr_spatial_attn_bwd(B, Kc, H, Fv, ptr(_f32c(dctx)), ptr(q_c), H, ptr(pk_c), Kc * H, ptr(f_c), f_bs,
                                          ptr(v_c), ptr(alpha), ptr(dq), ptr(dpk), ptr(dv_part), stream_ptr()),
              "pvcr_spatial_attn_bwd")
        return dq, dpk, None, dv_part.sum(dim=0).reshape(vshape)


class SpatialEncode(torch.autograd.Function):
    """SpatialNet's whole frame loop (model/SpatialNet.py:114-138) in one library call per direction
    (pvcr_spatial_encode_fwd / _bwd, csrc/spatial_sweep.cu): (proj_key [B,N,Kc,H], feats [B,N,Kc,F], query_layer.weight,
    energy_layer.weight, the encoder GRU's four parameters) -> outs [N,B,H], alphas [N,B,Kc].  Differentiable in proj_key and the
    parameters (the features are inputs)."""

    @staticmethod
    def forward(ctx, nsplit, pk, feats, w_q, v, w_ih, w_hh, b_ih, b_hh):
        B, N, Kc, H = pk.shape
        Fv = feats.shape[3]
        pk_c, f_c = _f32c(pk), _f32c(feats)
        w = tuple(_f32c(t) for t in (w_q, v, w_ih, w_hh, b_ih, b_hh))
        Lb = lib()
        ws = _ws(Lb.pvcr_spatial_encode_workspace(B, N, Kc, H, Fv, nsplit), pk_c.device)
        outs = torch.empty((N, B, H), dtype=torch.float32, device=pk_c.device)
        alphas = torch.empty((N, B, Kc), dtype=torch.float32, device=pk_c.device)
        check(Lb.pvcr_spatial_encode_fwd(B, N, Kc, H, Fv, nsplit, ptr(pk_c), ptr(f_c), ptr(w[0]), ptr(w[1]), ptr(w[2]), ptr(w[3]),
                                         ptr(w[4]), ptr(w[5]), ptr(outs), ptr(alphas), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_spatial_encode_fwd")
        ctx.meta = (B, N, Kc, H, Fv, nsplit, tuple(v.shape))
        ctx.keep = (pk_c, f_c, w, outs.detach(), alphas, ws)     # detached alias: the returned outs gets grad_fn = this node
        ctx.mark_non_differentiable(alphas)
        return outs, alphas

    @staticmethod
    def backward(ctx, d_outs, _d_alphas):
        B, N, Kc, H, Fv, nsplit, vshape = ctx.meta
        pk_c, f_c, w, outs, alphas, ws = ctx.keep
        d_pk = torch.empty_like(pk_c)
        g = tuple(torch.empty_like(t) for t in w)
        check(lib().pvcr_spatial_encode_bwd(B, N, Kc, H, Fv, nsplit, ptr(pk_c), ptr(f_c), ptr(w[0]), ptr(w[1]), ptr(w[2]), ptr(w[3]),
                                            ptr(outs), ptr(alphas), ptr(_f32c(d_outs)), ptr(d_pk), ptr(g[0]), ptr(g[1]), ptr(g[2]),
                                            ptr(g[3]), ptr(g[4]), ptr(g[5]), ptr(ws), ws.numel(), stream_ptr()),
              "pvcr_spatial_encode_bwd")
        return None, d_pk, None, g[0], g[1].reshape(vshape), g[2], g[3], g[4], g[5]
