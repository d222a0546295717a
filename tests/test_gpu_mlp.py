"""bfvi_mlp_fwd / _bwd on B200 (tcgen05 3xTF32 GEMMs + fused elementwise kernels) against the fp64 torch restatement, and
through the Python composed path of a model with a categorical modality."""
import pytest
import torch

import mlp_cases
from multimodal_dmm_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', sorted(mlp_cases.CASES))
def test_mlp_matches_torch(name):
    lib = _lib.load()
    mlp_cases.check(name, lib, 'cuda', 1e-4)
    assert 'mlp_bwd' in ';'.join(lib.last_dispatch())


def test_categorical_modality_runs_on_our_kernels():
    """encode / decode of a categorical modality (models/dmm.py:78-82, 96-98; models/common.py:9-23) no longer touch
    nn.Embedding / nn.Linear forward: same values and gradients as the torch modules that own the weights."""
    import multimodal_dmm_b200.models as models
    torch.manual_seed(3)
    m = models.MultiDMM(['x', 'act'], [4, 10], dists=['Normal', 'Categorical'], h_dim=32, z_dim=16,
                        device=torch.device('cuda:0')).train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(6, 5, 4, generator=g).cuda()
    act = torch.randint(0, 10, (6, 5, 1), generator=g).float().cuda()
    act[2, 1] = float('nan')
    mean, std, mask = m.encode({'x': x, 'act': act})
    assert not mask[1, 2, 1] and mask[1].sum() == 29
    ref_mu, ref_sd = m.enc['act'](torch.nan_to_num(act, nan=0.0).long().flatten(0, 1).squeeze(-1))
    assert torch.allclose(mean[1].flatten(0, 1), ref_mu, rtol=1e-4, atol=1e-5)
    assert torch.allclose(std[1].flatten(0, 1), ref_sd, rtol=1e-4, atol=1e-5)
    z = torch.randn(6, 5, 16, generator=g).cuda().requires_grad_(True)
    probs = m.decode(z)['act'][0]
    assert 'mlp_fwd softmax' in ';'.join(_lib.load().last_dispatch())   # (the log is per host thread: forward calls)
    ref_p = m.dec['act'](z.reshape(-1, 16))[0].reshape(6, 5, 10)
    assert torch.allclose(probs, ref_p, rtol=1e-4, atol=1e-6)
    w = torch.randn(6, 5, 10, generator=g).cuda()
    gz, gw = torch.autograd.grad((probs * w).sum(), [z, m.dec['act'].h_to_out[0].weight], retain_graph=True)
    rz, rw = torch.autograd.grad((ref_p * w).sum(), [z, m.dec['act'].h_to_out[0].weight])
    assert torch.allclose(gz, rz, rtol=1e-3, atol=1e-6) and torch.allclose(gw, rw, rtol=1e-3, atol=1e-6)
    ge, = torch.autograd.grad((mean[1] * std[1]).sum(), [m.enc['act'][0].weight])
    re, = torch.autograd.grad((ref_mu * ref_sd).sum(), [m.enc['act'][0].weight])
    assert torch.allclose(ge, re, rtol=1e-3, atol=1e-6)
