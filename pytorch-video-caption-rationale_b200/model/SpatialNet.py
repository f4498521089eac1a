"""SpatialNet (BASELINE config 4) around the B200 caption networks -- FIRST STAGE.

Same constructor, ``forward(vid_feats [B,N,F,K,K], s) -> (logits [B,L,Vc], seq_alphas [B,N,K,K])`` contract and
``state_dict`` keys as the reference (model/SpatialNet.py:55-142).  What runs where today:

* the caption network behind ``encode_step`` / ``decode`` (per-frame encoder GRU step, whole decoder with attention,
  vocabulary projection, their hand-written backward passes) runs on the sm_100a kernels of libpvcr_b200.so;
* the front of the network -- two Conv3x3 + BatchNorm + ReLU blocks and the per-frame spatial attention over the K*K
  cells (model/SpatialNet.py:76-86, 27-53, 120-138) -- is still expressed with torch.nn ops (cuDNN / cuBLAS library
  calls), exactly the reference's arithmetic; hand-written kernels for it (implicit-GEMM convolution on the tcgen05 GEMM,
  spatial attention fused into the encoder step) are the open part of row f1.

Parity: tests/test_gpu_boundary.py against goldens of the unmodified reference SpatialNet (oracle/gen_golden_spatial.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .S2VTAttModel import S2VTAttModel
from .S2VTModel import S2VTModel


class Attention(nn.Module):
    """Bahdanau attention over the K*K cells of one frame (model/SpatialNet.py:14-53): returns (context, alphas)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.key_layer = nn.Linear(hidden_size, hidden_size, bias=False)
        self.query_layer = nn.Linear(hidden_size, hidden_size, bias=False)
        self.energy_layer = nn.Linear(hidden_size, 1, bias=False)

    def project_keys(self, conv_feats):
        """key_layer applied to every frame's cells at once (the reference re-applies it inside its frame loop,
        model/SpatialNet.py:39; it does not depend on the recurrent state, so it is hoisted)."""
        return self.key_layer(conv_feats)

    def forward(self, query, proj_key, feats):
        """query [B,H] (encoder state), proj_key [B,K^2,H], feats [B,K^2,F] -> context [B,F], alphas [B,K^2]."""
        q = self.query_layer(query)
        scores = self.energy_layer(torch.tanh(q.unsqueeze(1) + proj_key)).squeeze(-1)
        alphas = F.softmax(scores, dim=1)
        return torch.bmm(alphas.unsqueeze(1), feats).squeeze(1), alphas


class SpatialNet(nn.Module):
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

    def forward(self, vid_feats, s=None):
        B, N, Fd, K, _ = vid_feats.shape
        cells = K * K
        conv = self.conv(vid_feats.reshape(-1, Fd, K, K)).view(B, N, -1, cells).transpose(2, 3)      # B x N x K^2 x H
        feats = vid_feats.view(B, N, Fd, cells).transpose(2, 3)                                      # B x N x K^2 x F
        proj_key = self.attention.project_keys(conv)
        state = torch.zeros(1, B, self.hidden_size, device=vid_feats.device, dtype=vid_feats.dtype)
        outs, seq_alphas = [], []
        for i in range(N):
            context, alphas = self.attention(state.squeeze(0), proj_key[:, i], feats[:, i])
            out, state = self.caption_net.encode_step(context, state)        # sm_100a GRU step (pvcr_gru_step_fwd/bwd)
            outs.append(out)
            seq_alphas.append(alphas.view(-1, K, K).unsqueeze(1))
        logits = self.caption_net.decode(torch.cat(outs, dim=0), state, s)    # sm_100a decoder + vocabulary projection
        return logits, torch.cat(seq_alphas, dim=1)
