"""The caller side: `run_iter` with the reference's signature (train.py:32-44) and the graph-captured
`Trainer` iteration (train.py:150-160: run_iter + zero_grad + backward + clip_grad_norm_ + Adam)."""
import numpy as np
import pytest
import torch

from tests.golden_util import relerr
from tests.gpu_util import FixtureGlove, load_case, to_cuda

pytestmark = pytest.mark.gpu


def _batch(d):
    return {"vid_feats": torch.from_numpy(d["vid"]), "sent": torch.from_numpy(d["s"]), "sent_len": torch.from_numpy(d["s_len"])}


def test_run_iter_matches_reference_loop_contract():
    from pvcr_b200.model import S2VTAttModel
    from pvcr_b200.trainer import run_iter
    d, params, grads, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_mid")
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L, precision="bf16x3"), params).train()
    crit = torch.nn.CrossEntropyLoss(reduction="none")
    acc, loss, pred = run_iter(None, _batch(d), m, crit, return_pred=True)          # CPU batch, as a DataLoader yields it
    assert abs(loss.item() - float(d["loss"])) < 2e-6 * abs(float(d["loss"]))
    assert abs(acc.item() - float(d["acc"])) < 1e-6 and np.array_equal(pred.cpu().numpy(), d["pred"])
    # the reference's optimizer block (train.py:157-160) on top of it
    opt = torch.optim.Adam(m.parameters(), lr=2e-3, weight_decay=4e-5)
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    for k, p in m.named_parameters():
        pass
    opt.step()
    m.eval()
    with torch.no_grad():
        acc_e, loss_e = run_iter(None, _batch(d), m, crit)
    assert torch.isfinite(loss_e) and 0.0 <= float(acc_e) <= 1.0


def test_trainer_iterations_match_the_manual_loop():
    from pvcr_b200.model import S2VTAttModel
    from pvcr_b200.optim import FusedClipAdam
    from pvcr_b200.trainer import Trainer
    d, params, grads, (B, N, V, H, E, L, Vc) = load_case("s2vtatt_mid")
    data = _batch(d)
    m = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).train()
    tr = Trainer(m, data, lr=2e-3, weight_decay=4e-5, max_norm=1.0)
    for k, p in m.named_parameters():            # building the trainer did not train the model
        assert torch.equal(p.detach().cpu(), torch.from_numpy(np.asarray(params[k], np.float32))), k
    losses = []
    for it in range(3):
        loss, acc, pred = tr.train_iter(data, next_data=data if it < 2 else None)
        losses.append(float(loss.item()))
    assert tr.n_iter == 3 and losses[2] < losses[0]
    mean_loss, mean_acc = tr.metrics()
    assert abs(mean_loss - sum(losses) / 3) < 1e-5
    ref = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).train()
    opt = FusedClipAdam(ref.parameters(), lr=2e-3, weight_decay=4e-5, max_norm=1.0)
    dv = tuple(data[k].cuda() for k in ("vid_feats", "sent", "sent_len"))
    for it in range(3):
        l, _, _ = ref.train_step_grads(*dv)
        assert abs(float(l.item()) - losses[it]) < 1e-5 * abs(losses[it]), (it, float(l.item()), losses[it])
        opt.step()
    for (k, a), (_, b) in zip(m.named_parameters(), ref.named_parameters()):
        assert relerr(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < 1e-5, k
    st = tr.save_state(opts={"arch": "s2vt-att"})
    assert set(st) == {"epoch", "state_dict", "optimizer", "n_iter", "opts", "val_meteor_score", "best_val_meteor_score"}
    m2 = to_cuda(S2VTAttModel(FixtureGlove(Vc, E), 0.0, H, V, L), params).train()
    tr2 = Trainer(m2, data)
    tr2.load_state(st)
    assert tr2.n_iter == 3 and tr2.epoch == 1
    for (k, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), k


def test_trainer_drives_spatialnet():
    """Trainer on a SpatialNet (train_spatial.py's loop): the graph-captured iteration (autograd tape + clip + Adam inside one CUDA
    graph) against the manual loop run_iter + backward + FusedClipAdam on a second copy of the model."""
    import copy
    from pvcr_b200.model import SpatialNet
    from pvcr_b200.optim import FusedClipAdam
    from pvcr_b200.trainer import Trainer, run_iter
    B, N, Fd, K, H, E, L, Vc = 6, 4, 64, 3, 64, 16, 6, 60
    torch.manual_seed(4)
    m = SpatialNet(FixtureGlove(Vc, E), 0.0, H, Fd, L, "s2vt-att", precision="bf16x3").cuda().train()
    ref = copy.deepcopy(m)
    data = {"vid_feats": torch.randn(B, N, Fd, K, K), "sent": torch.randint(0, Vc - 4, (B, L)), "sent_len": torch.randint(1, L + 1, (B,))}
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    tr = Trainer(m, data, lr=2e-3, weight_decay=4e-5, max_norm=1.0)
    for k, v in m.state_dict().items():          # building the trainer trained nothing and left the BatchNorm statistics alone
        assert torch.equal(v, before[k]), k
    losses = []
    for it in range(3):
        loss, acc, pred = tr.train_iter(data, next_data=data if it < 2 else None)
        losses.append(float(loss.item()))
        assert pred.shape == (B, L) and 0.0 <= float(acc) <= 1.0
    assert losses[2] < losses[0]
    opt = FusedClipAdam(ref.parameters(), lr=2e-3, weight_decay=4e-5, max_norm=1.0)
    for p in ref.parameters():
        p.grad = torch.zeros_like(p)
    for it in range(3):
        for p in ref.parameters():
            p.grad.zero_()
        acc, l = run_iter(None, data, ref, None)
        assert abs(float(l.item()) - losses[it]) < 1e-4 * abs(losses[it]), (it, float(l.item()), losses[it])
        l.backward()
        opt.step()
    for (k, a), (_, b) in zip(m.state_dict().items(), ref.state_dict().items()):
        if a.dtype.is_floating_point:
            assert relerr(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < 1e-3, k      # (a dropped step or gradient shows as >= 1e-2)
        else:
            assert torch.equal(a, b), k
