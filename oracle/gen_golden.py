"""Generate golden vectors by executing the REFERENCE modules (test infrastructure).

Runs only in the authoring container, where the reference is mounted at /root/reference; the
GPU box never sees it.  Imports model/S2VTModel.py, model/S2VTAttModel.py, model/RationaleNet.py
and train_utils.py unmodified, runs them in float64 on seeded synthetic inputs, and stores
inputs, weights (reference state_dict under the seed) and outputs (loss, accuracy, logits,
hooked attention weights, every parameter gradient from torch autograd, greedy ids) in
``tests/golden/*.npz``.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

RNG neutralisation: dropout_p = 0; the Gumbel Exp(1) draws of F.gumbel_softmax are
injected by patching ``Tensor.exponential_``; the scheduled-sampling coin ``random.random()``
(model/S2VTModel.py:134) is replaced by a scripted sequence.
"""
import contextlib
import os
import random
import sys

import numpy as np
import torch

REF = os.environ.get("PVCR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class FakeGlove:
    """Duck type of utils.GloveLoader used by the model ctors (utils.py:52-66)."""

    def __init__(self, vocab, embed, seed):
        rs = np.random.RandomState(seed)
        self.word_vectors = [rs.randn(embed).astype(np.float32).astype(np.float64) * 0.5 for _ in range(vocab)]
        self.vocab = vocab

    def get_id(self, w):
        return {"<sos>": self.vocab - 4, "<eos>": self.vocab - 3, "<pad>": self.vocab - 2, "<unk>": self.vocab - 1}[w]


def make_batch(B, N, V, L, Vc, seed):
    rs = np.random.RandomState(seed)
    vid = rs.randn(B, N, V).astype(np.float32).astype(np.float64)
    if B > 1:
        vid[1, N - 2:] = 0.0                      # zero-padded tail frames (dataset.py:77-78)
    s_len = rs.randint(1, L + 1, size=B)
    s_len[0] = L
    s = np.full((B, L), Vc - 2, np.int64)         # <pad>
    for b in range(B):
        s[b, :s_len[b] - 1] = rs.randint(0, Vc - 4, size=s_len[b] - 1)
        s[b, s_len[b] - 1] = Vc - 3               # <eos>
    return vid, s, s_len.astype(np.int64)


def round_f32(sd):
    """Weights are stored float32-exact so that fp32 product code sees identical values."""
    return {k: v.detach().to(torch.float32).to(torch.float64) for k, v in sd.items()}


@contextlib.contextmanager
def inject_exponential(noise):
    orig = torch.Tensor.exponential_

    def fake(self, *a, **k):
        assert self.shape == noise.shape, (self.shape, noise.shape)
        return self.copy_(noise)

    torch.Tensor.exponential_ = fake
    try:
        yield
    finally:
        torch.Tensor.exponential_ = orig


@contextlib.contextmanager
def scripted_coin(module, outcomes, prob):
    """Make ``random.random() < prob`` evaluate to the scripted outcomes inside ``module``."""
    seq = iter(outcomes)
    orig = module.random.random
    module.random.random = lambda: (prob - 1.0) if next(seq) else (prob + 1.0)
    try:
        yield
    finally:
        module.random.random = orig


def grads_of(model):
    return {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}


def run_iter_plain(model, tu, vid, s, s_len):
    """train.py:32-44 run_iter restated (train.py itself needs tensorboardX/nlgeval)."""
    crit = torch.nn.CrossEntropyLoss(reduction="none")
    logits = model(vid, s)
    pred = torch.argmax(logits, dim=2)
    loss = tu.calc_masked_loss(logits, s, s_len, crit)
    acc = tu.calc_masked_accuracy(logits, s, s_len)
    return logits, pred, loss, acc


def case_s2vtatt(tag, dims, seed, S2VTAttModel, tu):
    B, N, V, H, E, L, Vc = dims
    torch.manual_seed(seed)
    glove = FakeGlove(Vc, E, seed)
    model = S2VTAttModel(glove, 0.0, H, V, L).double()
    model.load_state_dict(round_f32(model.state_dict()))
    vid, s, s_len = make_batch(B, N, V, L, Vc, seed + 1)
    tv, ts, tl = torch.from_numpy(vid), torch.from_numpy(s), torch.from_numpy(s_len)
    alphas = []
    hook = model.decoder.attention.energy_layer.register_forward_hook(
        lambda m, i, o: alphas.append(torch.softmax(o.view(B, -1), dim=1).detach().numpy().copy()))
    model.train()
    logits, pred, loss, acc = run_iter_plain(model, tu, tv, ts, tl)
    loss.backward()
    out = dict(dims=np.array(dims), sos_id=glove.get_id("<sos>"), vid=vid.astype(np.float32), s=s, s_len=s_len,
               loss=loss.item(), acc=acc.item(), pred=pred.numpy(), logits=logits.detach().numpy(),
               alphas=np.stack(alphas))
    alphas.clear()
    for k, g in grads_of(model).items():
        out["grad." + k] = g
    for k, v in model.state_dict().items():
        out["param." + k] = v.numpy().astype(np.float32)
    model.eval()
    with torch.no_grad():
        glog = model(tv, None)
    out["greedy_logits"] = glog.numpy()
    out["greedy_ids"] = torch.argmax(glog, dim=2).numpy()
    out["greedy_alphas"] = np.stack(alphas)
    hook.remove()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    print(tag, "loss", out["loss"], "acc", out["acc"])


def case_s2vt(tag, dims, seed, S2VTModel, mod, tu, teacher):
    B, N, V, H, E, L, Vc = dims
    torch.manual_seed(seed)
    glove = FakeGlove(Vc, E, seed)
    model = S2VTModel(glove, 0.0, H, V, L).double()
    model.load_state_dict(round_f32(model.state_dict()))
    vid, s, s_len = make_batch(B, N, V, L, Vc, seed + 1)
    tv, ts, tl = torch.from_numpy(vid), torch.from_numpy(s), torch.from_numpy(s_len)
    model.train()
    model.teacher_force_prob = 1.0 if all(teacher) else 0.5
    with scripted_coin(mod, teacher, model.teacher_force_prob):
        logits, pred, loss, acc = run_iter_plain(model, tu, tv, ts, tl)
    loss.backward()
    out = dict(dims=np.array(dims), sos_id=glove.get_id("<sos>"), vid=vid.astype(np.float32), s=s, s_len=s_len,
               teacher=np.array(teacher), loss=loss.item(), acc=acc.item(), pred=pred.numpy(),
               logits=logits.detach().numpy())
    for k, g in grads_of(model).items():
        out["grad." + k] = g
    for k, v in model.state_dict().items():
        out["param." + k] = v.numpy().astype(np.float32)
    model.eval()
    with torch.no_grad():
        glog = model(tv, None)
    out["greedy_logits"] = glog.numpy()
    out["greedy_ids"] = torch.argmax(glog, dim=2).numpy()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    print(tag, "loss", out["loss"], "acc", out["acc"])


def case_rationale(tag, dims, seed, arch, tau, RationaleNet, tu):
    B, N, V, H, E, L, Vc = dims
    torch.manual_seed(seed)
    glove = FakeGlove(Vc, E, seed)
    model = RationaleNet(glove, 0.0, H, V, L, tau, arch).double()
    model.load_state_dict(round_f32(model.state_dict()))
    vid, s, s_len = make_batch(B, N, V, L, Vc, seed + 1)
    tv, ts, tl = torch.from_numpy(vid), torch.from_numpy(s), torch.from_numpy(s_len)
    noise = torch.from_numpy(np.random.RandomState(seed + 2).exponential(size=(B * N, 2)).astype(np.float32)
                             .astype(np.float64))
    crit = torch.nn.CrossEntropyLoss(reduction="none")
    lam_b, lam_c = 0.7, 1.3
    model.train()
    with inject_exponential(noise):              # train_rationale.py:30-44 run_iter restated
        logits, probs = model(tv, ts)
    pred = torch.argmax(logits, dim=2)
    loss_ce = tu.calc_masked_loss(logits, ts, tl, crit)
    loss_brev = tu.calc_brevity_loss(probs) * lam_b
    loss_cont = tu.calc_cont_loss(probs) * lam_c
    rlen = torch.sum(probs[:, :, 1], dim=1).mean()
    acc = tu.calc_masked_accuracy(logits, ts, tl)
    loss = loss_ce + loss_brev + loss_cont
    loss.backward()
    out = dict(dims=np.array(dims), sos_id=glove.get_id("<sos>"), vid=vid.astype(np.float32), s=s, s_len=s_len,
               noise=noise.numpy().astype(np.float32), tau=tau, lambda_brev=lam_b, lambda_cont=lam_c,
               loss=loss.item(), loss_ce=loss_ce.item(), loss_brev=loss_brev.item(), loss_cont=loss_cont.item(),
               rationale_len=rlen.item(), acc=acc.item(), pred=pred.numpy(), logits=logits.detach().numpy(),
               probs=probs.detach().numpy())
    for k, g in grads_of(model).items():
        out["grad." + k] = g
    for k, v in model.state_dict().items():
        out["param." + k] = v.numpy().astype(np.float32)
    model.eval()
    with torch.no_grad(), inject_exponential(noise):
        glog, gprobs = model(tv, None)
    out["greedy_logits"] = glog.numpy()
    out["greedy_ids"] = torch.argmax(glog, dim=2).numpy()
    out["greedy_probs"] = gprobs.numpy()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    print(tag, "loss", out["loss"], "acc", out["acc"], "rlen", out["rationale_len"])


def main():
    sys.path.insert(0, REF)
    torch.set_default_dtype(torch.float64)       # S2VTModel.decode hard-codes default-dtype zeros
    import train_utils as tu
    from model import S2VTModel as s2vt_mod
    from model.S2VTModel import S2VTModel
    from model.S2VTAttModel import S2VTAttModel
    from model.RationaleNet import RationaleNet
    tu.device = torch.device("cpu")
    os.makedirs(OUT, exist_ok=True)
    random.seed(0)
    #            B  N  V   H   E   L  Vc
    tiny = (3, 5, 24, 16, 12, 6, 30)
    mid = (4, 6, 96, 64, 20, 5, 70)
    case_s2vtatt("s2vtatt_tiny", tiny, 11, S2VTAttModel, tu)
    case_s2vtatt("s2vtatt_mid", mid, 12, S2VTAttModel, tu)
    case_s2vt("s2vt_tiny", tiny, 21, S2VTModel, s2vt_mod, tu, [True] * tiny[5])
    case_s2vt("s2vt_mid", mid, 22, S2VTModel, s2vt_mod, tu, [True] * mid[5])
    case_s2vt("s2vt_sched", tiny, 23, S2VTModel, s2vt_mod, tu, [True, False, True, True, False, True])
    case_rationale("rationale_att_tiny", tiny, 31, "s2vt-att", 1.0, RationaleNet, tu)
    case_rationale("rationale_att_mid", mid, 32, "s2vt-att", 0.7, RationaleNet, tu)
    case_rationale("rationale_s2vt_tiny", tiny, 33, "s2vt", 1.0, RationaleNet, tu)


if __name__ == "__main__":
    main()
