"""bfvi_nll_bernoulli_* / bfvi_nll_categorical_* against the reference formulas
(models/losses.py:23-66, restated with today's bool masks) on CPU through the emulated
kernels; the GPU twin (through models.losses + autograd) is tests/test_gpu_losses.py."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import helpers
from multimodal_dmm_b200 import _lib


def ref_bernoulli(theta, x, mask):
    keep = ~torch.isnan(x)
    if mask is not None:
        keep = keep & mask.view(list(mask.shape) + [1] * (x.dim() - mask.dim()))
    return F.binary_cross_entropy(theta.masked_select(keep), x.masked_select(keep), reduction='sum')


def ref_categorical(probs, x, mask):
    keep = ~torch.isnan(x)
    if mask is not None:
        keep = keep & mask.view(list(mask.shape) + [1] * (x.dim() - mask.dim()))
    cols = torch.stack([probs[:, :, k:k + 1].masked_select(keep) for k in range(probs.shape[2])], dim=-1)
    return F.nll_loss(cols, x.masked_select(keep).long(), reduction='sum')


def bernoulli_case(T, B, shape, seed, with_mask=True, edge=False):
    g = torch.Generator().manual_seed(seed)
    theta = torch.rand(T, B, *shape, generator=g) * 0.98 + 0.01
    x = (torch.rand(T, B, *shape, generator=g) < 0.4).float()
    if edge:                       # saturated probabilities: the -100 / 1e-12 clamps
        theta.view(-1)[::7] = 0.0
        theta.view(-1)[3::11] = 1.0
    x[torch.rand(T, B, generator=g) < 0.3] = float('nan')          # dropped frames
    x.view(-1)[::13] = float('nan')                                # stray missing pixels
    mask = (torch.rand(T, B, generator=g) < 0.8) if with_mask else None
    return theta, x, mask


@pytest.mark.parametrize('T,B,shape,with_mask,edge', [
    (5, 3, (8,), True, False), (4, 2, (3, 4, 4), True, False), (3, 2, (5,), False, False),
    (2, 3, (1, 8, 8), True, True), (1, 1, (1,), True, False)])
def test_nll_bernoulli_emulated(T, B, shape, with_mask, edge):
    lib = helpers.emu_library()
    theta, x, mask = bernoulli_case(T, B, shape, 3, with_mask, edge)
    rows, d = T * B, x[0, 0].numel()
    rmask = None if mask is None else mask.to(torch.uint8).reshape(-1).contiguous()
    out = torch.zeros(1, dtype=torch.float64)
    lib.call('bfvi_nll_bernoulli_fwd', _lib.ptr(theta), _lib.ptr(x), _lib.ptr(rmask), rows, d, _lib.ptr(out), None)
    th = theta.clone().requires_grad_(True)
    ref = ref_bernoulli(th, x, mask)
    assert abs(out.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    ref.backward()
    d_th = torch.full_like(theta, 7.0)
    lib.call('bfvi_nll_bernoulli_bwd', _lib.ptr(theta), _lib.ptr(x), _lib.ptr(rmask), rows, d, C.c_float(1.0),
             _lib.ptr(d_th), None)
    assert torch.allclose(d_th, th.grad, rtol=1e-5, atol=1e-6), (d_th - th.grad).abs().max()


@pytest.mark.parametrize('T,B,K,with_mask', [(6, 4, 10, True), (3, 5, 3, False), (1, 1, 2, True)])
def test_nll_categorical_emulated(T, B, K, with_mask):
    lib = helpers.emu_library()
    g = torch.Generator().manual_seed(5)
    probs = torch.softmax(torch.randn(T, B, K, generator=g), dim=2)
    x = torch.randint(0, K, (T, B, 1), generator=g).float()
    x[torch.rand(T, B, 1, generator=g) < 0.3] = float('nan')
    mask = (torch.rand(T, B, generator=g) < 0.8) if with_mask else None
    rmask = None if mask is None else mask.to(torch.uint8).reshape(-1).contiguous()
    out = torch.zeros(1, dtype=torch.float64)
    lib.call('bfvi_nll_categorical_fwd', _lib.ptr(probs), _lib.ptr(x), _lib.ptr(rmask), T * B, K, _lib.ptr(out), None)
    pr = probs.clone().requires_grad_(True)
    ref = ref_categorical(pr, x, mask)
    assert abs(out.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    ref.backward()
    d_pr = torch.full_like(probs, 7.0)
    lib.call('bfvi_nll_categorical_bwd', _lib.ptr(probs), _lib.ptr(x), _lib.ptr(rmask), T * B, K, C.c_float(1.0),
             _lib.ptr(d_pr), None)
    assert torch.equal(d_pr, pr.grad)


def test_loss_entry_points_reject_bad_arguments():
    lib = helpers.emu_library()
    out = torch.zeros(1, dtype=torch.float64)
    t = torch.rand(4)
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_nll_bernoulli_fwd', None, _lib.ptr(t), None, 1, 4, _lib.ptr(out), None)
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_nll_categorical_fwd', _lib.ptr(t), _lib.ptr(t), None, 0, 4, _lib.ptr(out), None)
