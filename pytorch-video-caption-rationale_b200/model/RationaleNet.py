"""RationaleNet on the B200 kernels: the frame-selection Generator in front of a caption network.

Same constructors, ``forward(vid_feats, s) -> (logits, probs)`` contract and ``state_dict`` keys as the reference
(model/RationaleNet.py:14-106).  The selected features ``vid_feats * probs[:, :, 1]`` are never materialised: the
generator hands p1 to the caption network as a per-frame scale applied while the encoder input is staged, and
receives d p1 from the caption network's backward (include/pvcr_b200.h).
"""
import torch
import torch.nn as nn

from .. import functional as F_
from .S2VTAttModel import S2VTAttModel
from .S2VTModel import S2VTModel


class Generator(nn.Module):
    """Parameter container + launcher of pvcr_generator_* (model/RationaleNet.py:14-54)."""

    def __init__(self, dropout_p, hidden_size, vid_feat_size, tau, precision='bf16'):
        super().__init__()
        self.rnn = nn.LSTM(input_size=vid_feat_size, hidden_size=hidden_size, bidirectional=True, num_layers=1)
        self.drop = nn.Dropout(p=dropout_p)
        self.linear = nn.Linear(hidden_size * 2, 2)
        self.tau = tau
        self.precision = precision
        self.noise = None        # optional [B*N,2] Exp(1) draws injected by tests (F.gumbel_softmax's exponential_())

    def _params(self):
        r = self.rnn
        return (r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, r.weight_ih_l0_reverse,
                r.weight_hh_l0_reverse, r.bias_ih_l0_reverse, r.bias_hh_l0_reverse, self.linear.weight, self.linear.bias)

    def _cfg(self):
        train = self.training
        p = float(self.drop.p) if train else 0.0
        return {"nsplit": S2VTAttModel.NSPLIT[self.precision] if train else 3, "dropout_p": p, "tau": float(self.tau),
                "hard": 0 if train else 1, "seed": F_.next_seed()}

    def select(self, vid_feats):
        """-> probs [B,N,2], p1 [B,N], pen [2] (brevity, continuity)."""
        return F_.GeneratorSelect.apply(self._cfg(), vid_feats, self.noise, *self._params())

    def forward(self, vid_feats):
        """Reference signature: (sel_vid_feats, probs).  Materialises the product; the RationaleNet paths below do not."""
        probs, p1, _ = self.select(vid_feats)
        return vid_feats * p1.unsqueeze(-1), probs


class RationaleNet(nn.Module):
    def __init__(self, glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, tau, arch, pretrained_base=None,
                 precision='bf16'):
        super().__init__()
        if arch == 's2vt':
            self.caption_net = S2VTModel(glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, precision)
        elif arch == 's2vt-att':
            self.caption_net = S2VTAttModel(glove_loader, dropout_p, hidden_size, vid_feat_size, max_len, precision)
        else:
            raise NotImplementedError('unknown video captioning arch')
        if pretrained_base is not None:
            pretrained_dict = torch.load(pretrained_base, map_location='cpu')['state_dict']
            self.caption_net.load_state_dict(pretrained_dict)
        self.gen = Generator(dropout_p, hidden_size, vid_feat_size, tau, precision)

    def forward(self, vid_feats, s=None):
        """-> (logits [B,L,Vc], probs [B,N,2])  (model/RationaleNet.py:86-106)."""
        probs, p1, _ = self.gen.select(vid_feats)
        logits = self.caption_net(vid_feats, s, frame_scale=p1)
        return logits, probs

    def forward_loss(self, vid_feats, s, s_len, lambda_brev=1.0, lambda_cont=1.0):
        """Fused run_iter of train_rationale.py:30-44.
        -> (acc, loss, loss_ce, loss_brev, loss_cont, rationale_len, pred, probs)."""
        assert self.training and s is not None
        probs, p1, pen = self.gen.select(vid_feats)
        loss_ce, acc, pred = self.caption_net.forward_loss(vid_feats, s, s_len, frame_scale=p1)
        loss_brev, loss_cont = pen[0] * lambda_brev, pen[1] * lambda_cont
        return acc, loss_ce + loss_brev + loss_cont, loss_ce, loss_brev, loss_cont, pen[0].detach(), pred, probs

    def train_step_stages(self, vid_feats, s, s_len, lambda_brev=1.0, lambda_cont=1.0, deferred_join=False):
        """Generator form of the tape-free joint step (train_rationale.py:30-44 + backward): passes the caption network's
        stages through, yields once more when ALL caption-network gradients are final (the generator's backward -- two
        LSTM sweeps and the 86-GFLOP input-projection gradient -- still to come), and returns (loss, acc, pred).  A
        data-parallel caller overlaps each stage's all-reduce with what follows (parallel.GradAllReducer)."""
        assert self.training and s is not None
        gen, cap = self.gen, self.caption_net
        cg = F_.ManualCtx()
        gparams = gen._params()
        probs, p1, pen = F_.GeneratorSelect.forward(cg, gen._cfg(), vid_feats, gen.noise, *gparams)
        p1.requires_grad_(True)          # makes the caption network emit d loss / d frame_scale
        if hasattr(cap, "train_step_stages"):
            loss_ce, acc, pred = yield from cap.train_step_stages(vid_feats, s, s_len, frame_scale=p1,
                                                                  deferred_join=deferred_join)
        else:
            loss_ce, acc, pred = cap.train_step_grads(vid_feats, s, s_len, frame_scale=p1)
        yield "caption_net_grads"
        dev = vid_feats.device         # fill kernels (no host-to-device copy): stays CUDA-graph capturable
        d_pen = torch.cat([torch.full((1,), float(lambda_brev), device=dev), torch.full((1,), float(lambda_cont), device=dev)])
        given = {f: p.grad for f, p in zip(F_.GEN_PARAM_FIELDS, gparams) if p.grad is not None}
        grads = F_.GeneratorSelect.backward(cg, None, cap.last_frame_scale_grad, d_pen, grad_out=given)
        for p, g in zip(gparams, grads[3:]):
            p.grad = g
        return loss_ce + lambda_brev * pen[0] + lambda_cont * pen[1], acc, pred

    def early_grad_params(self):
        """Per yield of train_step_stages, the parameters whose gradients are final at that point."""
        cap = self.caption_net
        early = [list(e) for e in cap.early_grad_params()] if hasattr(cap, "early_grad_params") else []
        seen = {id(p) for e in early for p in e}
        early.append([p for p in cap.parameters() if id(p) not in seen])
        return early

    @property
    def OVERLAPPED_STAGES(self):
        # every caption-network bucket is followed by a persistent sweep (the generator's LSTM backward at the latest)
        return len(self.early_grad_params())

    @torch.no_grad()
    def train_step_grads(self, vid_feats, s, s_len, lambda_brev=1.0, lambda_cont=1.0):
        """Tape-free fwd+bwd of the joint objective (CUDA-graph capturable); (over)writes every ``param.grad``.
        -> (loss, acc, pred)."""
        gen = self.train_step_stages(vid_feats, s, s_len, lambda_brev, lambda_cont, deferred_join=True)
        try:
            while True:
                next(gen)
        except StopIteration as done:
            return done.value

    @torch.no_grad()
    def greedy(self, vid_feats):
        """Eval: straight-through hard selection (still stochastic, as in the reference) + greedy captioning."""
        probs, p1, _ = self.gen.select(vid_feats)
        ids, logits = self.caption_net.greedy(vid_feats, frame_scale=p1)
        return ids, logits, probs
