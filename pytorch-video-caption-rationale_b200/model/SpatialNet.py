"""SpatialNet (BASELINE config 4) on the B200 kernels.

Same constructor, ``forward(vid_feats [B,N,F,K,K], s) -> (logits [B,L,Vc], seq_alphas [B,N,K,K])`` contract and
``state_dict`` keys as the reference (model/SpatialNet.py:55-142).  The torch.nn layers below are parameter / buffer
containers only (so reference checkpoints load key for key, BatchNorm running statistics included); the arithmetic runs in
libpvcr_b200.so:

* the two Conv3x3 + BatchNorm2d + ReLU blocks (:76-86,106) -- nine tcgen05 GEMMs per convolution on row-shifted views of a flat
  zero-padded channels-last matrix, BatchNorm statistics / apply kernels, hand-written backward (csrc/conv.cu,
  ``functional.SpatialFront``);
* ``key_layer`` hoisted out of the frame loop (the reference re-applies it per frame, :39; it does not depend on the
  recurrent state) and ``query_layer`` per frame -- ``functional.Linear`` (pvcr_linear_fwd/bwd);
* the per-frame attention over the K*K cells (:27-53) -- ``functional.SpatialAttnStep`` (pvcr_spatial_attn_fwd/bwd);
* the caption network behind ``encode_step`` / ``decode`` (per-frame encoder GRU step, persistent decoder sweeps,
  vocabulary projection).

Still a per-frame launch chain (query GEMM, attention, GRU-step GEMMs and gates: the encoder input depends on the attention
of the same step, so nothing of it can be hoisted); a persistent fused spatial-attention + GRU sweep is the open part
(DESIGN.md section 7).  torch supplies memory, views / transposes between layouts and the autograd tape.

Parity: tests/test_gpu_boundary.py against goldens of the unmodified reference SpatialNet (oracle/gen_golden_spatial.py).
"""
import os

import torch
import torch.nn as nn

from .. import functional as F_
from .S2VTAttModel import S2VTAttModel
from .S2VTModel import S2VTModel


class Attention(nn.Module):
    """Parameter container of the spatial attention (model/SpatialNet.py:14-25)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.key_layer = nn.Linear(hidden_size, hidden_size, bias=False)
        self.query_layer = nn.Linear(hidden_size, hidden_size, bias=False)
        self.energy_layer = nn.Linear(hidden_size, 1, bias=False)


class SpatialNet(nn.Module):
    NSPLIT = {'bf16': 1, 'bf16x2': 2, 'bf16x3': 3}

    def __init__(self, glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, arch, precision='bf16'):
        super().__init__()
        if arch == 's2vt':
            self.caption_net = S2VTModel(glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, precision)
        elif arch == 's2vt-att':
            self.caption_net = S2VTAttModel(glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, precision)
        else:
            raise NotImplementedError('unknown video captioning arch')
        self.conv = nn.Sequential(
            nn.Conv2d(vid_feat_size, hidden_size, 3, 1, 1), nn.BatchNorm2d(hidden_size), nn.ReLU(),
            nn.Conv2d(hidden_size, hidden_size, 3, 1, 1), nn.BatchNorm2d(hidden_size), nn.ReLU())
        self.attention = Attention(hidden_size)
        self.hidden_size = hidden_size
        self.precision = precision

    def _front(self, x):
        """x [I,F,K,K] -> conv_feats [I*K*K, H], feats_cl [I*K*K, F]."""
        c1, b1, c2, b2 = self.conv[0], self.conv[1], self.conv[3], self.conv[4]
        cfg = {"nsplit": self.NSPLIT[self.precision] if self.training else 3, "training": self.training, "eps": b1.eps,
               "momentum": b1.momentum}
        out = F_.SpatialFront.apply(cfg, x, b1.running_mean, b1.running_var, b2.running_mean, b2.running_var, c1.weight,
                                    c1.bias, b1.weight, b1.bias, c2.weight, c2.bias, b2.weight, b2.bias)
        if self.training:
            with torch.no_grad():
                b1.num_batches_tracked += 1
                b2.num_batches_tracked += 1
        return out

    def forward(self, vid_feats, s=None):
        B, N, Fd, K, _ = vid_feats.shape
        cells, H = K * K, self.hidden_size
        ns = self.NSPLIT[self.precision] if self.training else 3
        conv_feats, feats_cl = self._front(vid_feats.reshape(B * N, Fd, K, K))
        proj_key = F_.Linear.apply(ns, conv_feats, self.attention.key_layer.weight)
        v = self.attention.energy_layer.weight
        if os.environ.get("PVCR_SPATIAL_STEPWISE") is None:
            # the whole frame loop in one library call per direction (csrc/spatial_sweep.cu): weights staged once, q and W_hh h from
            # one stacked product, parameter gradients as products over all frames
            rnn = self.caption_net.encoder.rnn if isinstance(self.caption_net, S2VTAttModel) else self.caption_net.rnn1
            outs, alphas = F_.SpatialEncode.apply(ns, proj_key.view(B, N, cells, H), feats_cl.detach().view(B, N, cells, Fd),
                                                  self.attention.query_layer.weight, v, rnn.weight_ih_l0, rnn.weight_hh_l0,
                                                  rnn.bias_ih_l0, rnn.bias_hh_l0)
            logits = self.caption_net.decode(outs, outs[N - 1:], s)
            return logits, alphas.view(N, B, K, K).transpose(0, 1).contiguous()
        # step-wise variant (A/B knob, and the shape of the reference's loop): one Linear + attention + encode_step per frame
        # frame-major copies so that every frame's [B, K*K, .] slice is contiguous (unbind: its backward is one stack)
        pk_frames = proj_key.view(B, N, cells, H).transpose(0, 1).contiguous().unbind(0)
        # (the features need no gradient: frame i is read in place as a batch-strided view, no frame-major copy of 1.5 GB at cfg4)
        feat_view = feats_cl.detach().view(B, N, cells, Fd)
        state = torch.zeros(1, B, H, device=vid_feats.device, dtype=torch.float32)
        outs, seq_alphas = [], []
        for i in range(N):
            q = F_.Linear.apply(ns, state.squeeze(0), self.attention.query_layer.weight)
            context, alphas = F_.SpatialAttnStep.apply(q, pk_frames[i], feat_view[:, i], v)
            out, state = self.caption_net.encode_step(context, state)
            outs.append(out)
            seq_alphas.append(alphas.view(-1, K, K).unsqueeze(1))
        logits = self.caption_net.decode(torch.cat(outs, dim=0), state, s)
        return logits, torch.cat(seq_alphas, dim=1)
