"""S2VTModel on the B200 kernels.

Same constructor, ``forward(vid_feats, s)`` contract, ``teacher_force_prob`` attribute and ``state_dict`` keys as the
reference class (model/S2VTModel.py:12-202).  The torch.nn layers are parameter containers (identical default
initialisation under a seed: Xavier-normal / bias 0.01 through ``ixvr`` exactly as the reference constructor does);
their ``forward`` is never called — all arithmetic goes through the C ABI (include/pvcr_b200.h).
"""
import random

import numpy as np
import torch
import torch.nn as nn

from .. import functional as F_
from ..utils import ixvr


class S2VTModel(nn.Module):
    NSPLIT = {'bf16': 1, 'bf16x2': 2, 'bf16x3': 3}

    def __init__(self, glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, precision='bf16'):
        super().__init__()
        word_vectors = np.vstack(glove_loader.word_vectors)
        self.vocab_size, self.embed_size = word_vectors.shape
        self.vid_feat_size = vid_feat_size
        self.hidden_size = hidden_size
        self.max_len = max_len
        self.sos_id = glove_loader.get_id('<sos>')
        self.teacher_force_prob = 1.0
        self.precision = precision
        self.embedding = nn.Sequential(nn.Embedding(self.vocab_size, self.embed_size), nn.Dropout(p=dropout_p))
        self.embedding[0].load_state_dict({'weight': torch.Tensor(word_vectors)})
        self.rnn1 = nn.GRU(input_size=vid_feat_size, hidden_size=hidden_size, num_layers=1)
        self.rnn2 = nn.GRU(input_size=hidden_size + self.embed_size, hidden_size=hidden_size, num_layers=1)
        self.linear = nn.Sequential(nn.Dropout(p=dropout_p), nn.Linear(hidden_size, self.vocab_size))
        self.reset_parameters()

    def reset_parameters(self):
        self.apply(ixvr)

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _seq_params(self):
        r1, r2 = self.rnn1, self.rnn2
        return (self.embedding[0].weight, r1.weight_ih_l0, r1.weight_hh_l0, r1.bias_ih_l0, r1.bias_hh_l0,
                r2.weight_ih_l0, r2.weight_hh_l0, r2.bias_ih_l0, r2.bias_hh_l0)

    def _cfg(self, train):
        p_out = float(self.linear[0].p) if train else 0.0
        p_emb = float(self.embedding[1].p) if train else 0.0
        return {"nsplit": self.NSPLIT[self.precision] if train else 3, "dropout_p": p_out, "emb_dropout_p": p_emb,
                "seed": F_.next_seed() if (p_out > 0 or p_emb > 0) else 0}

    def _fed_words(self, vid_feats, s, frame_scale, cfg):
        """Input word of every decoding step.  Teacher forcing: [<sos>, s[:, :L-1]].  With scheduled sampling
        (teacher_force_prob < 1) one coin per step for the whole batch, drawn from Python's RNG exactly as the
        reference does (model/S2VTModel.py:134); steps that lose the coin are fed the arg-max of the previous step,
        obtained from a gradient-free step-wise decode with the same dropout masks."""
        B, L = vid_feats.shape[0], self.max_len
        sos = torch.full((B, 1), self.sos_id, dtype=torch.long, device=s.device)
        teacher = torch.cat((sos, s[:, :L - 1]), dim=1)
        coins = [random.random() < self.teacher_force_prob for _ in range(L)]
        if all(coins):
            return teacher
        lin = self.linear[1]
        _, _, fed = F_.s2vt_decode_steps(vid_feats, frame_scale, self.sos_id, L, self._seq_params(), lin.weight,
                                         lin.bias, cfg, teacher_words=teacher, teacher_mask=coins, want_logits=False)
        return fed

    def _hidden_states(self, vid_feats, s, frame_scale=None):
        cfg = self._cfg(True)
        s_in = self._fed_words(vid_feats, s, frame_scale, cfg)
        hs = F_.S2VTSequence.apply(cfg, vid_feats, frame_scale, s_in, *self._seq_params())
        return hs, cfg

    # ---- reference API --------------------------------------------------------------------------------------
    def encode_step(self, vid_feat, rnn_state=None):
        """vid_feat [B,V], rnn_state [1,B,H] | None -> (output [1,B,H], rnn_state [1,B,H]): one rnn1 step
        (model/S2VTModel.py:57-72; SpatialNet.py:127)."""
        r = self.rnn1
        h_prev = None if rnn_state is None else rnn_state.reshape(-1, self.hidden_size)
        nsplit = self.NSPLIT[self.precision] if self.training else 3
        h = F_.GruStep.apply(nsplit, vid_feat, h_prev, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0)
        out = h.unsqueeze(0)
        return out, out

    def encode(self, vid_feats):
        """vid_feats [B,N,V] -> (output [N,B,H], rnn_state [1,B,H]): rnn1 over the frames (model/S2VTModel.py:74-86),
        step by step through ``encode_step`` (forward() itself hoists the input projection and never calls this)."""
        state, outs = None, []
        for n in range(vid_feats.shape[1]):
            out, state = self.encode_step(vid_feats[:, n], state)
            outs.append(out)
        return torch.cat(outs, dim=0), state

    def decode(self, output1, state1, s):
        """output1 [N,B,H], state1 [1,B,H], s [B,L] | None -> logits [B,L,Vc] (model/S2VTModel.py:88-177; SpatialNet.py:140).
        Teacher forcing only on this entry (scheduled sampling needs the fused forward())."""
        lin = self.linear[1]
        out1 = output1.transpose(0, 1)
        st1 = state1.reshape(-1, self.hidden_size)
        if self.training:
            assert s is not None
            assert self.teacher_force_prob >= 1.0, "decode(): scheduled sampling is only implemented on forward()"
            cfg = self._cfg(True)
            B, L = out1.shape[0], self.max_len
            sos = torch.full((B, 1), self.sos_id, dtype=torch.long, device=s.device)
            s_in = torch.cat((sos, s[:, :L - 1]), dim=1)
            hs = F_.S2VTDecode.apply(cfg, out1, st1, s_in, *self._seq_params())
            return F_.VocabLogits.apply(cfg, hs, lin.weight, lin.bias)
        with torch.no_grad():
            _, logits = F_.s2vt_decode_greedy(out1, st1, self.sos_id, self.max_len, self._seq_params(), lin.weight,
                                              lin.bias)
        return logits

    def forward(self, vid_feats, s=None, frame_scale=None):
        """vid_feats [B,N,V], s [B,L] (required in training) -> logits [B,L,Vc] (model/S2VTModel.py:179-202)."""
        lin = self.linear[1]
        if self.training:
            assert s is not None
            hs, cfg = self._hidden_states(vid_feats, s, frame_scale)
            return F_.VocabLogits.apply(cfg, hs, lin.weight, lin.bias)
        return self.greedy(vid_feats, frame_scale)[1]

    def forward_loss(self, vid_feats, s, s_len, frame_scale=None):
        """Fused run_iter (train.py:37-40): (loss, acc, pred) with the loss contract of calc_masked_loss /
        calc_masked_accuracy evaluated inside the vocabulary-projection kernels."""
        assert self.training and s is not None
        lin = self.linear[1]
        hs, cfg = self._hidden_states(vid_feats, s, frame_scale)
        loss, stats, pred = F_.VocabCrossEntropy.apply(cfg, hs, lin.weight, lin.bias, s, s_len)
        self.last_token_nll = cfg.get("token_nll")
        return loss, stats[0] / stats[1], pred

    @torch.no_grad()
    def train_step_grads(self, vid_feats, s, s_len, frame_scale=None):
        """Tape-free fwd+bwd (see S2VTAttModel.train_step_grads); capturable in a CUDA graph when
        teacher_force_prob == 1."""
        assert self.training and s is not None
        cfg = self._cfg(True)
        params = self._seq_params()
        lin = self.linear[1]
        cfg["grad_out"] = {f: p.grad for f, p in zip(F_.S2VT_SEQ_FIELDS, params) if p.grad is not None}
        cfg["vocab_grad_out"] = {"out_w": lin.weight.grad, "out_b": lin.bias.grad}
        s_in = self._fed_words(vid_feats, s, frame_scale, cfg)
        c1, c2 = F_.ManualCtx(), F_.ManualCtx()
        hs = F_.S2VTSequence.forward(c1, cfg, vid_feats, frame_scale, s_in, *params)
        loss, stats, pred = F_.VocabCrossEntropy.forward(c2, cfg, hs, lin.weight, lin.bias, s, s_len)
        one = torch.ones((), dtype=torch.float32, device=hs.device)
        _, d_hs, d_w, d_b, _, _ = F_.VocabCrossEntropy.backward(c2, one, None, None)
        grads = F_.S2VTSequence.backward(c1, d_hs)
        for p, g in zip(params, grads[4:]):
            p.grad = g
        lin.weight.grad, lin.bias.grad = d_w, d_b
        self.last_frame_scale_grad = grads[2]
        return loss, stats[0] / stats[1], pred

    @torch.no_grad()
    def greedy(self, vid_feats, frame_scale=None):
        """Eval branch (model/S2VTModel.py:147-177): fixed max_len steps with arg-max feedback.
        -> (ids [B,L], logits [B,L,Vc])."""
        lin = self.linear[1]
        ids, logits, _ = F_.s2vt_decode_steps(vid_feats, frame_scale, self.sos_id, self.max_len, self._seq_params(),
                                              lin.weight, lin.bias, self._cfg(False))
        return ids, logits
