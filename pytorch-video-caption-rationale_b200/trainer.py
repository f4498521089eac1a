"""The caller side of the hot path: `run_iter` of the reference's train scripts and the training
iteration around it, on the B200 kernels.

* ``run_iter(opts, data, model, criterion, return_pred=False)`` -- same signature and return values as train.py:32-44 /
  train_spatial.py:30-39 (``run_iter_rationale``: train_rationale.py:30-44).  ``data`` is the reference's collated batch
  (dataset.py:118-139: 'vid_feats' [B,N,V] float, 'sent' [B,L] long, 'sent_len' [B] long).  In training mode it goes
  through the fused ``model.forward_loss`` (logits never materialised) and returns an autograd loss, so the reference
  loop ``optimizer.zero_grad(); loss.backward(); clip_grad_norm_(); optimizer.step()`` (train.py:157-160) works unchanged.
* ``Trainer`` -- the same iteration without the host in the loop (S2VT / S2VTAtt / RationaleNet through their tape-free
  ``train_step_grads``; SpatialNet, train_spatial.py, through ``graphs.GraphedAutogradStep``): forward, backward, ``clip_grad_norm_`` and Adam
  captured as ONE CUDA graph (graphs.GraphedTrainStep + optim.FusedClipAdam, gradient all-reduce included when
  data-parallel), the NEXT batch copied host -> device from pinned staging buffers on a copy stream while the current
  step computes, and no device -> host synchronisation per iteration (the reference syncs twice: train.py:151
  ``pred.data.cpu()`` and logger.py:34 ``.item()``): metrics stay on the device until ``metrics()`` is called.
  Checkpoints keep the reference's dictionary layout (train.py:183-193) so `--resume` / `pretrained_base` files interchange.
"""
import torch

from .graphs import GraphedTrainStep
from .optim import FusedClipAdam


def _device_of(model):
    return next(model.parameters()).device


def run_iter(opts, data, model, criterion=None, return_pred=False):
    """train.py:32-44.  -> (acc, loss) or (acc, loss, pred).  ``criterion`` (CrossEntropyLoss(reduction='none') in the
    reference) is accepted for signature compatibility; the masked loss is evaluated inside the kernels."""
    dev = _device_of(model)
    vid_feats, s, s_len = data['vid_feats'].to(dev), data['sent'].to(dev), data['sent_len'].to(dev)
    if model.training and hasattr(model, 'forward_loss'):
        out = model.forward_loss(vid_feats, s, s_len)
        loss, acc, pred = out[0], out[1], out[2]
    else:                                        # eval, and SpatialNet in training (train_spatial.py:30-39: logits materialised)
        from . import train_utils as TU
        logits = model(vid_feats, s)
        if isinstance(logits, tuple):            # SpatialNet / RationaleNet return (logits, extra)
            logits = logits[0]
        pred = torch.argmax(logits, dim=2)
        loss = TU.calc_masked_loss(logits, s, s_len, criterion)
        acc = TU.calc_masked_accuracy(logits, s, s_len)
    if not return_pred:
        return acc, loss
    return acc, loss, pred


def run_iter_rationale(opts, data, model, criterion=None, return_pred=False):
    """train_rationale.py:30-44.  -> (acc, loss, loss_ce, loss_brev, loss_cont, rationale_len[, pred])."""
    dev = _device_of(model)
    vid_feats, s, s_len = data['vid_feats'].to(dev), data['sent'].to(dev), data['sent_len'].to(dev)
    acc, loss, loss_ce, loss_brev, loss_cont, rlen, pred, _ = model.forward_loss(
        vid_feats, s, s_len, lambda_brev=getattr(opts, 'lambda_brev', 1.0), lambda_cont=getattr(opts, 'lambda_cont', 1.0))
    out = (acc, loss, loss_ce, loss_brev, loss_cont, rlen)
    return out + (pred,) if return_pred else out


class Trainer:
    """One object per process (per GPU).  ``example`` = a batch dict with the shapes of every later batch."""

    def __init__(self, model, example, lr=2e-3, weight_decay=4e-5, max_norm=1.0, reducer=None):
        self.model = model.train()
        dev = _device_of(model)
        self.optimizer = FusedClipAdam(model.parameters(), lr=lr, weight_decay=weight_decay, max_norm=max_norm)
        # pinned host staging (what a DataLoader(pin_memory=True) would hand over): the H2D copies are asynchronous
        self._pinned = tuple(torch.empty(example[k].shape, dtype=example[k].dtype, pin_memory=True)
                             for k in ('vid_feats', 'sent', 'sent_len'))
        dev_example = tuple(example[k].to(dev) for k in ('vid_feats', 'sent', 'sent_len'))
        # capturing the step executes it (pointer tables, communicator set-up) -- the example batch must not train the
        # model: parameters and optimizer state are put back afterwards
        keep = [p.detach().clone() for p in model.parameters()]
        if hasattr(model, 'train_step_grads'):
            self.step = GraphedTrainStep(self.model, dev_example, warmup=1, reducer=reducer, optimizer=self.optimizer)
        else:
            # SpatialNet (train_spatial.py:30-39,  the same iteration around a module whose step is a chain of autograd
            # Functions): forward + masked loss / accuracy / predictions + backward + clip + Adam as one graph, tape inside
            assert reducer is None, "data-parallel training is implemented for the modules with train_step_stages"
            from . import train_utils as TU
            from .graphs import GraphedAutogradStep

            def loss_fn(vid_feats, s, s_len):
                logits = self.model(vid_feats, s)
                logits = logits[0] if isinstance(logits, tuple) else logits
                loss, stats, pred = TU._MaskedCE.apply(logits, s, s_len)
                return loss, stats[0] / stats[1], pred

            self.step = GraphedAutogradStep(self.model, loss_fn, warmup=1, example_inputs=dev_example, optimizer=self.optimizer)
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, k in zip(model.parameters(), keep):
                p.copy_(k)
            for m, v in zip(self.optimizer.exp_avg, self.optimizer.exp_avg_sq):
                m.zero_(); v.zero_()
            self.optimizer.step_count.zero_()
        self.n_iter = 0
        self.epoch = 0
        self._sums = torch.zeros(2, dtype=torch.float32, device=dev)     # running (loss, acc) since the last metrics()
        self._count = 0
        self._primed = False
        self._last_h2d = None

    def _stage(self, data):
        if self._last_h2d is not None:       # the previous copy out of the pinned buffers must have finished reading them
            self._last_h2d.synchronize()
        for dst, k in zip(self._pinned, ('vid_feats', 'sent', 'sent_len')):
            src = data[k]
            if src.is_cuda:
                return tuple(data[k] for k in ('vid_feats', 'sent', 'sent_len'))      # already on the device
            dst.copy_(src)
        return self._pinned

    def train_iter(self, data, next_data=None):
        """One training iteration (train.py:150-160) on ``data``; if ``next_data`` is given its host -> device copy is
        started before returning, overlapping this step's compute.  Returns the device tensors (loss, acc, pred) of this
        step WITHOUT synchronising (they are overwritten by the next call)."""
        if not self._primed:
            self.step.prefetch(*self._stage(data))
            self._last_h2d = self.step._staged
        loss, acc, pred = self.step.step_prefetched()
        self._primed = False
        if next_data is not None:
            self.step.prefetch(*self._stage(next_data))
            self._last_h2d = self.step._staged
            self._primed = True
        self._sums += torch.stack((loss.detach(), acc.detach()))
        self._count += 1
        self.n_iter += 1
        return loss, acc, pred

    def metrics(self):
        """(mean loss, mean acc) since the last call -- the only host synchronisation (logger.py:32-43 every log_iter)."""
        if not self._count:
            return 0.0, 0.0
        l, a = (self._sums / self._count).tolist()
        self._sums.zero_()
        self._count = 0
        return l, a

    # ---- checkpoints: the reference's save_state dictionary (train.py:183-193) -------------------------------------------
    def save_state(self, opts=None, val_meteor_score=0.0, best_val_meteor_score=0.0):
        return {'epoch': self.epoch, 'state_dict': self.model.state_dict(), 'optimizer': self.optimizer.state_dict(),
                'n_iter': self.n_iter, 'opts': opts, 'val_meteor_score': val_meteor_score,
                'best_val_meteor_score': best_val_meteor_score}

    def load_state(self, save_state, load_optimizer=True):
        """The reference restores the model and counters only (train.py:125-134: the saved optimizer state is never
        loaded); ``load_optimizer`` also restores Adam's moments when the checkpoint came from this Trainer."""
        self.model.load_state_dict(save_state['state_dict'])
        self.n_iter = save_state.get('n_iter', 0)
        self.epoch = save_state.get('epoch', -1) + 1
        opt = save_state.get('optimizer')
        if load_optimizer and opt and 'state' in opt and len(opt['state']) == len(self.optimizer.params):
            self.optimizer.load_state_dict(opt)
