"""S2VTAttModel on the B200 kernels.

Same constructor, ``forward(vid_feats, s)`` contract, helper methods and ``state_dict`` keys as the reference
class (model/S2VTAttModel.py:199-264); the torch.nn layers below are parameter containers only (so reference
checkpoints load and default initialisation under a seed is identical) — their ``forward`` is never called.
All arithmetic goes through the C ABI (include/pvcr_b200.h).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from .. import functional as F_
from ..utils import ixvr


class Attention(nn.Module):
    """Parameter container of the additive attention (model/S2VTAttModel.py:12-23)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.key_layer = nn.Linear(hidden_size, hidden_size, bias=False)
        self.query_layer = nn.Linear(hidden_size, hidden_size, bias=False)
        self.energy_layer = nn.Linear(hidden_size, 1, bias=False)


class Encoder(nn.Module):
    """model/S2VTAttModel.py:50-61."""

    def __init__(self, vid_feat_size, hidden_size):
        super().__init__()
        self.rnn = nn.GRU(vid_feat_size, hidden_size, num_layers=1)


class Decoder(nn.Module):
    """model/S2VTAttModel.py:98-123."""

    def __init__(self, glove_loader, hidden_size, dropout_p, max_len):
        super().__init__()
        word_vectors = np.vstack(glove_loader.word_vectors)
        self.vocab_size, self.embed_size = word_vectors.shape
        self.max_len = max_len
        self.sos_id = glove_loader.get_id('<sos>')
        self.embedding = nn.Embedding(self.vocab_size, self.embed_size)
        self.embedding.load_state_dict({'weight': torch.Tensor(word_vectors)})
        self.rnn = nn.GRU(hidden_size + self.embed_size, hidden_size, num_layers=1)
        self.attention = Attention(hidden_size=hidden_size)
        self.pred_linear = nn.Sequential(nn.Dropout(p=dropout_p), nn.Linear(hidden_size, self.vocab_size))


class S2VTAttModel(nn.Module):
    """S2VT with Bahdanau temporal attention.

    ``precision``: 'bf16' (one bf16 tcgen05 product per logical product, fp32 accumulation; training default),
    'bf16x2' or 'bf16x3' (split-bf16 operands; 'bf16x3' reproduces fp32 arithmetic and is always used for greedy
    decoding so that token ids match the fp32 reference)."""

    NSPLIT = {'bf16': 1, 'bf16x2': 2, 'bf16x3': 3}

    def __init__(self, glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, precision='bf16'):
        super().__init__()
        self.encoder = Encoder(vid_feat_size, hidden_size)
        self.decoder = Decoder(glove_loader, hidden_size, dropout_p, max_len)
        self.hidden_size = hidden_size
        self.precision = precision
        self.teacher_force_prob = 1.0      # set by train.py:145 on every arch; unused here as in the reference
        self.last_alphas = None            # [L,B,N] attention weights of the latest forward (reference never exposes them)
        self.last_token_nll = None         # [B,L] per-token loss of the latest fused forward_loss / train step

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _seq_params(self):
        e, d = self.encoder.rnn, self.decoder
        return (e.weight_ih_l0, e.weight_hh_l0, e.bias_ih_l0, e.bias_hh_l0, d.embedding.weight, d.rnn.weight_ih_l0,
                d.rnn.weight_hh_l0, d.rnn.bias_ih_l0, d.rnn.bias_hh_l0, d.attention.key_layer.weight,
                d.attention.query_layer.weight, d.attention.energy_layer.weight)

    def _cfg(self, train):
        p = float(self.decoder.pred_linear[0].p) if train else 0.0
        return {"nsplit": self.NSPLIT[self.precision] if train else 3, "dropout_p": p,
                "seed": F_.next_seed() if p > 0 else 0}

    def _shifted(self, s, B):
        sos = torch.full((B, 1), self.decoder.sos_id, dtype=torch.long, device=s.device)
        return torch.cat((sos, s[:, :self.decoder.max_len - 1]), dim=1)

    def _hidden_states(self, vid_feats, s, frame_scale=None):
        cfg = self._cfg(True)
        hs, alphas = F_.S2VTAttSequence.apply(cfg, vid_feats, frame_scale, self._shifted(s, vid_feats.shape[0]),
                                              *self._seq_params())
        self.last_alphas = alphas
        return hs, cfg

    # ---- reference API --------------------------------------------------------------------------------------
    def reset_parameter(self):
        """Xavier-normal weights, bias 0.01 (model/S2VTAttModel.py:215-217; never called by the reference constructor)."""
        self.apply(ixvr)

    def encode_step(self, vid_feat, rnn_state=None):
        """vid_feat [B,V], rnn_state [1,B,H] | None -> (output [1,B,H], rnn_state [1,B,H]): one encoder GRU step
        (model/S2VTAttModel.py:63-78,219-229), the call SpatialNet makes per frame (model/SpatialNet.py:127)."""
        r = self.encoder.rnn
        h_prev = None if rnn_state is None else rnn_state.reshape(-1, self.hidden_size)
        nsplit = self.NSPLIT[self.precision] if self.training else 3
        h = F_.GruStep.apply(nsplit, vid_feat, h_prev, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0)
        out = h.unsqueeze(0)
        return out, out

    def decode(self, encoder_outs, encoder_final, s):
        """encoder_outs [N,B,H], encoder_final [1,B,H], s [B,L] | None -> logits [B,L,Vc]
        (model/S2VTAttModel.py:231-243; SpatialNet.py:140)."""
        lin = self.decoder.pred_linear[1]
        enc = encoder_outs.transpose(0, 1)
        fin = encoder_final.reshape(-1, self.hidden_size)
        if self.training:
            assert s is not None
            cfg = self._cfg(True)
            hs, alphas = F_.S2VTAttDecode.apply(cfg, enc, fin, self._shifted(s, enc.shape[0]), *self._seq_params())
            self.last_alphas = alphas
            return F_.VocabLogits.apply(cfg, hs, lin.weight, lin.bias)
        with torch.no_grad():
            d = self.decoder
            _, logits, alphas = F_.s2vtatt_decode_greedy(enc, fin, d.sos_id, d.max_len, self._seq_params(), lin.weight,
                                                         lin.bias, plan=self._plan("decode"))
        self.last_alphas = alphas
        return logits

    def forward(self, vid_feats, s=None, frame_scale=None):
        """vid_feats [B,N,V], s [B,L] (required in training) -> logits [B,L,Vc] (model/S2VTAttModel.py:245-264)."""
        lin = self.decoder.pred_linear[1]
        if self.training:
            assert s is not None
            hs, cfg = self._hidden_states(vid_feats, s, frame_scale)
            return F_.VocabLogits.apply(cfg, hs, lin.weight, lin.bias)
        return self.greedy(vid_feats, frame_scale)[1]

    def forward_loss(self, vid_feats, s, s_len, frame_scale=None):
        """Fused run_iter (train.py:37-40): returns (loss, acc, pred) without materialising logits for the caller.
        loss == calc_masked_loss(model(vid_feats, s), s, s_len, CrossEntropyLoss(reduction='none'))."""
        assert self.training and s is not None
        lin = self.decoder.pred_linear[1]
        hs, cfg = self._hidden_states(vid_feats, s, frame_scale)
        loss, stats, pred = F_.VocabCrossEntropy.apply(cfg, hs, lin.weight, lin.bias, s, s_len)
        self.last_token_nll = cfg.get("token_nll")      # [B,L] unmasked per-token loss (criterion(...), train_utils.py:47-48)
        return loss, stats[0] / stats[1], pred

    def train_step_stages(self, vid_feats, s, s_len, frame_scale=None, deferred_join=False):
        """Generator form of the tape-free fwd+bwd: yields once when the vocabulary-projection gradients (out_w, out_b:
        ~47 % of all gradient bytes) are final, so a data-parallel caller can start their all-reduce while the rest of
        the backward runs; returns (loss, acc, pred)."""
        assert self.training and s is not None
        # deferred_join: the library's side lane (weight-gradient GEMMs etc.) is joined once, at the end of the step,
        # instead of at the end of every C call; a caller that consumes gradients at the yields must not set it
        Lb = F_.lib()
        deferred_join = deferred_join and Lb.pvcr_side_mode(-1) != 0
        prev_mode = Lb.pvcr_side_mode(2) if deferred_join else None
        ok = False
        try:
            out = yield from self._train_step_stages(vid_feats, s, s_len, frame_scale)
            ok = True
        finally:
            # deferred joins end here; after an exception (e.g. an allocation failure between two C calls) the lanes are
            # joined and the one-shot staging notes dropped in any mode, so that nothing of the aborted step is trusted
            # (or still being written) when a later step reuses the workspace addresses
            if deferred_join or not ok:
                F_.check(Lb.pvcr_side_join(F_.stream_ptr()), "pvcr_side_join")
            if deferred_join:
                Lb.pvcr_side_mode(prev_mode)
        return out

    def _train_step_stages(self, vid_feats, s, s_len, frame_scale):
        cfg = self._cfg(True)
        params = self._seq_params()
        lin = self.decoder.pred_linear[1]
        # existing .grad tensors (e.g. views into the all-reduce buckets of parallel.GradAllReducer) are written in place
        cfg["grad_out"] = {f: p.grad for f, p in zip(F_.ATT_SEQ_FIELDS, params) if p.grad is not None}
        cfg["vocab_grad_out"] = {"out_w": lin.weight.grad, "out_b": lin.bias.grad}
        c1, c2 = F_.ManualCtx(), F_.ManualCtx()
        F_.vocab_prepare(cfg, vid_feats.shape[0], 1 + min(s.shape[1], self.decoder.max_len - 1), lin.weight)      # W_v cast overlaps the sweeps
        hs, alphas = F_.S2VTAttSequence.forward(c1, cfg, vid_feats, frame_scale, self._shifted(s, vid_feats.shape[0]),
                                                *params)
        self.last_alphas = alphas
        loss, stats, pred = F_.VocabCrossEntropy.forward(c2, cfg, hs, lin.weight, lin.bias, s, s_len)
        self.last_token_nll = cfg.get("token_nll")
        one = torch.ones((), dtype=torch.float32, device=hs.device)
        _, d_hs, d_w, d_b, _, _ = F_.VocabCrossEntropy.backward(c2, one, None, None)
        lin.weight.grad, lin.bias.grad = d_w, d_b
        yield "vocab_grads"
        seq = F_.S2VTAttSequence.backward_in_parts(c1, d_hs, parts=(1, 2))
        try:
            part_grads = next(seq)                       # decoder half done
            for f, p in zip(F_.ATT_SEQ_FIELDS, params):
                if not f.startswith("enc_"):
                    p.grad = part_grads[f]
            # the embedding gradient (the bulk of the decoder half: Vc x E) is produced by side lane 1 alone, early in
            # the shadow of the encoder sweep: a data-parallel caller can start its all-reduce after joining THAT lane
            yield ("embedding_grad", "milestone", 0)
            if not self.MERGE_TAIL:
                yield "decoder_grads"
            next(seq)
            raise RuntimeError("backward_in_parts yielded more than once")
        except StopIteration as done:
            grads = done.value
        for p, g in zip(params, grads[4:]):
            p.grad = g
        self.last_frame_scale_grad = grads[2]
        return loss, stats[0] / stats[1], pred

    def early_grad_params(self):
        """Per yield of train_step_stages, the parameters whose gradients are final at that point (a yield is a name, or
        a (name, lane) pair when joining that one side lane of the library is enough)."""
        lin = self.decoder.pred_linear[1]
        d = self.decoder
        dec = [d.rnn.weight_ih_l0, d.rnn.weight_hh_l0, d.rnn.bias_ih_l0, d.rnn.bias_hh_l0,
               d.attention.key_layer.weight, d.attention.query_layer.weight, d.attention.energy_layer.weight]
        early = [[lin.bias, lin.weight], [d.embedding.weight]]
        if not self.MERGE_TAIL:
            early.append(dec)
        return early

    # buckets whose all-reduce overlaps a persistent sweep must stay within the SMs the sweep leaves free
    OVERLAPPED_STAGES = 2
    # PVCR_DP_MERGE_TAIL=1 reduces the small decoder-side gradients (6.6 MB, final while the encoder sweep runs) together with
    # the encoder's in ONE collective at the end of the step instead of two; measured slower (2 GPUs: +25 us, 8 GPUs: +15 us)
    MERGE_TAIL = os.environ.get("PVCR_DP_MERGE_TAIL", "0") != "0"

    @torch.no_grad()
    def train_step_grads(self, vid_feats, s, s_len, frame_scale=None):
        """Tape-free fwd+bwd of run_iter (train.py:37-40 + loss.backward()): chains the C-ABI forward and backward
        entry points by hand and (over)writes ``param.grad`` of every parameter.  Stream-ordered and free of autograd
        state, hence capturable in a CUDA graph (pvcr_b200.graphs.GraphedTrainStep).  Returns (loss, acc, pred)."""
        gen = self.train_step_stages(vid_feats, s, s_len, frame_scale, deferred_join=True)
        try:
            while True:
                next(gen)
        except StopIteration as done:
            return done.value

    def _plan(self, which):
        """Prepared-weights cache of the decoding entry points (functional.DecodePlan), one per entry point."""
        plans = self.__dict__.setdefault("_decode_plans", {})
        if which not in plans:
            plans[which] = F_.DecodePlan()
        return plans[which]

    def __getstate__(self):
        # the prepared decode workspaces (hundreds of MB of derived data) are neither pickled nor deep-copied
        state = self.__dict__.copy()
        state.pop("_decode_plans", None)
        return state

    def invalidate_decode_cache(self):
        """Forget the weights prepared for decoding (needed only after parameter writes autograd cannot see, e.g. through
        ``.data``; optimizer steps, ``load_state_dict`` and ``train()`` are noticed)."""
        for pl in self.__dict__.get("_decode_plans", {}).values():
            pl.invalidate()

    def train(self, mode=True):
        if mode:
            self.invalidate_decode_cache()
        return super().train(mode)

    @torch.no_grad()
    def greedy(self, vid_feats, frame_scale=None, return_logits=True):
        """Fixed-length greedy decoding (eval branch, model/S2VTAttModel.py:172-191): -> (ids [B,L], logits [B,L,Vc]);
        ``return_logits=False`` skips materialising the logits (captioning needs the ids only) and returns None for them."""
        d = self.decoder
        lin = d.pred_linear[1]
        ids, logits, alphas = F_.s2vtatt_greedy(vid_feats, frame_scale, d.sos_id, d.max_len, self._seq_params(),
                                                lin.weight, lin.bias, plan=self._plan("greedy"),
                                                return_logits=return_logits)
        self.last_alphas = alphas
        return ids, logits

    @torch.no_grad()
    def beam_search(self, vid_feats, beam=5, frame_scale=None):
        """Fixed-length beam search over the decoder step (BASELINE config 5; the reference itself only decodes
        greedily): -> (ids [B,beam,L], best hypothesis first; scores [B,beam] = sum of log-probabilities).
        beam=1 reproduces ``greedy``.  fp32-equivalent bf16x3 arithmetic, as for greedy decoding."""
        d = self.decoder
        lin = d.pred_linear[1]
        return F_.s2vtatt_beam(vid_feats, frame_scale, d.sos_id, d.max_len, beam, self._seq_params(), lin.weight, lin.bias)
