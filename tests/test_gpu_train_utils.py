"""GPU parity of the loss-contract functions (reference train_utils.py:22-95) incl. the edge cases."""
import numpy as np
import pytest
import torch

from tests.golden_util import CASES_ATT, CASES_RAT, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", CASES_ATT + ["s2vt_mid"])
def test_masked_loss_accuracy_on_reference_logits(tag):
    from pvcr_b200 import train_utils as TU
    d, _, _ = load(tag)
    logits = torch.from_numpy(d["logits"]).float().cuda().requires_grad_(True)
    s, s_len = torch.from_numpy(d["s"]).cuda(), torch.from_numpy(d["s_len"]).cuda()
    loss = TU.calc_masked_loss(logits, s, s_len, torch.nn.CrossEntropyLoss(reduction="none"))
    acc = TU.calc_masked_accuracy(logits, s, s_len)
    assert abs(loss.item() - float(d["loss"])) < 2e-6 * abs(float(d["loss"]))
    assert abs(acc.item() - float(d["acc"])) < 1e-6
    loss.backward()
    ref = logits.detach().clone().requires_grad_(True)
    B, L, Vc = ref.shape
    nll = torch.nn.functional.cross_entropy(ref.view(B * L, Vc), s.view(-1), reduction="none").view(B, L)
    mask = TU.calc_sentence_mask(B, L, s_len)
    ((nll * mask).sum(1) / s_len.float()).mean().backward()
    assert torch.allclose(logits.grad, ref.grad, atol=1e-7, rtol=1e-4)


def test_masked_loss_edge_cases():
    """s_len == 1 (only <eos>), s_len == L, exact ties in the arg-max (first index wins, as torch.argmax)."""
    from pvcr_b200 import train_utils as TU
    from oracle import captioning_oracle as O
    rs = np.random.RandomState(3)
    logits = rs.randn(3, 5, 11).astype(np.float32)
    logits[0, 0, 4] = logits[0, 0, 7] = 9.0
    s = rs.randint(0, 11, size=(3, 5))
    s_len = np.array([1, 5, 3])
    lt = torch.from_numpy(logits).cuda()
    loss = TU.calc_masked_loss(lt, torch.from_numpy(s).cuda(), torch.from_numpy(s_len).cuda())
    ref, _, _ = O.masked_loss(logits.astype(np.float64), s, s_len)
    assert abs(loss.item() - ref) < 1e-6 * abs(ref)
    from pvcr_b200.train_utils import _MaskedCE
    _, _, pred = _MaskedCE.apply(lt, torch.from_numpy(s).cuda(), torch.from_numpy(s_len).cuda())
    assert pred[0, 0].item() == 4
    assert np.array_equal(pred.cpu().numpy(), logits.argmax(2))


@pytest.mark.parametrize("tag", CASES_RAT)
def test_penalties(tag):
    from pvcr_b200 import train_utils as TU
    d, _, _ = load(tag)
    probs = torch.from_numpy(d["probs"]).float().cuda().requires_grad_(True)
    brev, cont = TU.calc_brevity_loss(probs), TU.calc_cont_loss(probs)
    assert abs(brev.item() * float(d["lambda_brev"]) - float(d["loss_brev"])) < 1e-5
    assert abs(cont.item() * float(d["lambda_cont"]) - float(d["loss_cont"])) < 1e-5
    (2.0 * brev + 3.0 * cont).backward()
    ref = probs.detach().clone().requires_grad_(True)
    p1 = ref[:, :, 1]
    (2.0 * p1.sum(1).mean() + 3.0 * (p1[:, 1:] - p1[:, :-1]).abs().mean()).backward()
    assert torch.allclose(probs.grad, ref.grad, atol=1e-6)
