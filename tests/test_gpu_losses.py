"""models.losses.nll_bernoulli / nll_categorical (CUDA kernels behind autograd Functions) against
the reference formulas (models/losses.py:23-66) on the same inputs, Weizmann-shaped included."""
import pytest
import torch

from test_emu_losses import bernoulli_case, ref_bernoulli, ref_categorical
from multimodal_dmm_b200 import _lib
from multimodal_dmm_b200.models import losses

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('T,B,shape,with_mask,edge', [
    (5, 3, (8,), True, False), (4, 2, (3, 5, 5), True, False), (3, 2, (5,), False, False),
    (2, 3, (1, 8, 8), True, True), (25, 25, (3, 64, 64), True, False)])
def test_nll_bernoulli_matches_reference(T, B, shape, with_mask, edge):
    theta, x, mask = bernoulli_case(T, B, shape, 11, with_mask, edge)
    th_ref = theta.clone().double().requires_grad_(True)
    ref = ref_bernoulli(th_ref, x.double(), mask)
    (ref * 0.37).backward()
    th = theta.cuda().requires_grad_(True)
    m = None if mask is None else mask.cuda().unsqueeze(-1)        # (T, B, 1) as the trainers pass it
    ours = losses.nll_bernoulli(th, x.cuda(), m)
    (ours * 0.37).backward()
    assert abs(ours.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    if edge:     # saturated pixels: compare with the fp32 clamp (1e-12 is below fp64's product there)
        th32 = theta.clone().requires_grad_(True)
        (ref_bernoulli(th32, x, mask) * 0.37).backward()
        want = th32.grad
    else:
        want = th_ref.grad.float()
    assert torch.allclose(th.grad.cpu(), want, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize('T,B,K', [(6, 4, 10), (25, 25, 10), (1, 1, 2)])
def test_nll_categorical_matches_reference(T, B, K):
    g = torch.Generator().manual_seed(5)
    probs = torch.softmax(torch.randn(T, B, K, generator=g), dim=2)
    x = torch.randint(0, K, (T, B, 1), generator=g).float()
    x[torch.rand(T, B, 1, generator=g) < 0.3] = float('nan')
    mask = torch.rand(T, B, generator=g) < 0.8
    pr_ref = probs.clone().requires_grad_(True)
    ref = ref_categorical(pr_ref, x, mask)
    (ref * 2.0).backward()
    pr = probs.cuda().requires_grad_(True)
    ours = losses.nll_categorical(pr, x.cuda(), mask.cuda().unsqueeze(-1))
    (ours * 2.0).backward()
    assert abs(ours.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    assert torch.equal(pr.grad.cpu(), pr_ref.grad)


def test_losses_refuse_cpu_tensors():
    with pytest.raises(_lib.BfviError):
        losses.nll_bernoulli(torch.rand(2, 2, 3), torch.rand(2, 2, 3))
    with pytest.raises(_lib.BfviError):
        losses.nll_categorical(torch.rand(2, 2, 3), torch.zeros(2, 2, 1))
