"""The reference's loss / metric contract (train_utils.py:22-95) on the B200 kernels, same function signatures.

These operate on a materialised ``logits`` tensor, for callers that keep the reference's ``model(vid_feats, s)`` +
``calc_masked_loss(...)`` structure.  The fused path (``model.forward_loss``) never materialises logits.
"""
import torch

from ._lib import check, lib, ptr, stream_ptr


def calc_sentence_mask(batch_size, max_len, s_len):
    """mask[b, l] = l < s_len[b], float [batch_size, max_len]  (train_utils.py:22-35, same signature)."""
    mask = torch.arange(0, max_len, device=s_len.device).expand(batch_size, -1)
    return (mask < s_len.unsqueeze(-1)).float()


class _MaskedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, s_len):
        B, L, Vc = logits.shape
        lg = logits.detach().float().contiguous()
        t, sl = target.detach().long().contiguous(), s_len.detach().long().contiguous()
        dev = lg.device
        loss3 = torch.empty(3, dtype=torch.float32, device=dev)
        pred = torch.empty((B, L), dtype=torch.int64, device=dev)
        lse = torch.empty((B, L), dtype=torch.float32, device=dev)
        nll = torch.empty((B, L), dtype=torch.float32, device=dev)
        check(lib().pvcr_masked_ce(ptr(lg), Vc, B, L, Vc, ptr(t), ptr(sl), None, ptr(loss3), ptr(pred), ptr(lse), ptr(nll),
                                   None, 0, stream_ptr()), "pvcr_masked_ce")
        ctx.keep = (lg, t, sl, pred, lse, nll)
        ctx.mark_non_differentiable(pred)
        stats = loss3[1:].clone()
        ctx.mark_non_differentiable(stats)
        return loss3[0].clone(), stats, pred

    @staticmethod
    def backward(ctx, d_loss, _s, _p):
        lg, t, sl, pred, lse, nll = ctx.keep
        B, L, Vc = lg.shape
        d = torch.empty_like(lg)
        gs = d_loss.detach().float().reshape(1).contiguous()
        check(lib().pvcr_masked_ce(ptr(lg), Vc, B, L, Vc, ptr(t), ptr(sl), ptr(gs), None, ptr(pred), ptr(lse), ptr(nll),
                                   ptr(d), Vc, stream_ptr()), "pvcr_masked_ce")
        return d, None, None


def calc_masked_loss(logits, target, s_len, criterion=None):
    """mean_b( sum_l nll[b,l] * mask[b,l] / s_len[b] )  (train_utils.py:37-54).  ``criterion`` is accepted for
    signature compatibility (the reference passes CrossEntropyLoss(reduction='none')) and ignored."""
    return _MaskedCE.apply(logits, target, s_len)[0]


def calc_masked_accuracy(logits, target, s_len):
    """Token accuracy under the sentence mask (train_utils.py:56-71)."""
    with torch.no_grad():
        _, stats, _ = _MaskedCE.apply(logits, target, s_len)
    return stats[0] / stats[1]


class _Penalties(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs):
        B, N, _ = probs.shape
        p = probs.detach().float().contiguous()
        pen = torch.empty(2, dtype=torch.float32, device=p.device)
        check(lib().pvcr_rationale_penalties(ptr(p), B, N, ptr(pen), stream_ptr()), "pvcr_rationale_penalties")
        ctx.keep = p
        return pen

    @staticmethod
    def backward(ctx, d_pen):
        p = ctx.keep
        B, N, _ = p.shape
        d = torch.empty_like(p)
        g = d_pen.detach().float().contiguous()
        check(lib().pvcr_rationale_penalties_bwd(ptr(p), B, N, ptr(g), ptr(d), stream_ptr()),
              "pvcr_rationale_penalties_bwd")
        return d


def calc_cont_loss(probs):
    """mean |p1[b,n] - p1[b,n-1]|  (train_utils.py:73-83)."""
    return _Penalties.apply(probs)[1]


def calc_brevity_loss(probs):
    """mean_b sum_n p1[b,n]  (train_utils.py:85-95)."""
    return _Penalties.apply(probs)[0]
