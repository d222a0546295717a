"""BASELINE config 4 (Weizmann-shaped): conv image encoders / decoders passed as custom modules (their layers run
on this library's bfvi_conv_* / bfvi_bn2d_* / bfvi_dense_* kernels, models/common.py), Bernoulli + Categorical
modalities, a dropped modality — the composed, differentiable path encode -> z_filter (fused temporal core,
large-dim family) -> decode against the REFERENCE's golden loss, posterior and gradients
(oracle/make_golden_weizmann.py)."""
import os
import sys

import pytest
import torch

from conftest import ROOT, rel_err
import multimodal_dmm_b200.models as models

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import make_golden_weizmann as gw  # noqa: E402

pytestmark = pytest.mark.gpu


def test_weizmann_shaped_forward_loss_backward_matches_reference():
    torch.backends.cudnn.allow_tf32 = False          # the reference ran fp32 convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    fx = torch.load(os.path.join(ROOT, 'tests', 'golden', 'weizmann', 'forward_fsmooth.pt'), weights_only=False)
    cfg = fx['cfg']
    m = gw.build(models, cfg, device='cuda:0').train()
    assert m._custom_enc == {'video', 'mask'} and m.dists['action'] == 'Categorical'
    inputs, targets, mask, eps_flt, eps_smt = gw.make_data(cfg)
    cu = lambda d: {k: v.cuda() for k, v in d.items()}
    infer, prior, recon = m(cu(inputs), lengths=cfg['lengths'], mode='fsmooth', flt_particles=cfg['k_flt'],
                            noise=(eps_flt.cuda(), eps_smt.cuda()))
    assert recon['video'][0].shape == targets['video'].shape and recon['action'][0].shape[-1] == 10
    loss = m.loss(cu(targets), infer, prior, recon, mask.cuda(), cfg['kld_mult'], cfg['rec_mults'])
    loss.backward()
    assert abs(loss.item() - fx['ref_loss']) / abs(fx['ref_loss']) < 1e-4, (loss.item(), fx['ref_loss'])
    assert torch.allclose(infer[0].cpu(), fx['ref_infer_mean'], rtol=2e-3, atol=2e-4)
    assert torch.allclose(infer[1].cpu(), fx['ref_infer_std'], rtol=2e-3, atol=2e-4)
    named = dict(m.named_parameters())
    checked = 0
    # noise floor: a convolution bias in front of a BatchNorm has a mathematically ZERO gradient
    # (both implementations return rounding noise there); errors are judged against the largest
    # gradient of the model as well as against the tensor itself
    floor = 1e-5 * max(r['full'].norm().item() if 'full' in r else r['norm'] for r in fx['ref_grads'].values())
    for k, ref in fx['ref_grads'].items():
        g = named[k].grad
        assert g is not None, k
        g = g.detach().float().cpu()
        if 'full' in ref:
            err = (g - ref['full']).norm().item()
            assert err < max(1e-3 * ref["full"].norm().item(), floor), (k, err, ref['full'].norm().item())
        else:
            w = torch.cos(torch.arange(g.numel(), dtype=torch.float32) * 0.37).reshape(g.shape)
            assert abs(g.norm().item() - ref['norm']) <= 1e-3 * ref["norm"], k
            assert abs((g * w).sum().item() - ref['proj']) <= 1e-3 * ref["norm"], k
        checked += 1
    assert checked == len(fx['ref_grads']) and checked > 50
