"""Large-dim (tcgen05) family on B200: bfvi_forward and MultiDMM.forward against the oracle on
identical injected noise.  Default precision is error-compensated 3xTF32 (FP32-class); the
single-pass TF32 mode is checked at TF32 tolerance on well-conditioned (filtering) modes."""
import pytest
import torch

import helpers
from multimodal_dmm_b200 import _lib
import multimodal_dmm_b200.models as models

pytestmark = pytest.mark.gpu

CASES = {
    'default32': dict(z_dim=32, h_dim=32, dims=[3, 5], t_max=12, lengths=[12, 12, 10, 7, 4, 2], seed=1),
    'c3_dims': dict(z_dim=64, h_dim=512, dims=[16] * 8, t_max=10, lengths=[10] * 9 + [6, 3], seed=2),
    'weizmann_core': dict(z_dim=256, h_dim=256, dims=[24, 10], t_max=8, lengths=[8, 8, 5], seed=3),
    'odd': dict(z_dim=20, h_dim=72, dims=[1, 7, 2], t_max=9, lengths=[9, 8, 8, 1], seed=4),
}


def noise(fx, k_flt, seed):
    g = torch.Generator().manual_seed(seed)
    t_max, b_dim, z = max(fx['lengths']), len(fx['lengths']), fx['z_dim']
    return torch.randn(t_max, b_dim, k_flt, z, generator=g), torch.randn(t_max, b_dim, 1, z, generator=g)


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('mode,sample,k_flt', [('bfilter', False, 1), ('ffilter', True, 1), ('fsmooth', True, 5),
                                               ('bsmooth', True, 1), ('fsmooth', False, 25)])
def test_forward_3xtf32_matches_oracle(name, mode, sample, k_flt):
    lib = _lib.load()
    fx = helpers.large_case(**CASES[name])
    eps_flt, eps_smt = noise(fx, k_flt, 7)
    ours = helpers.run_forward_large(lib, fx, 'cuda', mode, sample, k_flt, eps_flt, eps_smt, precision=0)
    ref = helpers.oracle_forward(fx, mode, sample, k_flt, eps_flt, eps_smt, dtype=torch.float64)
    # smoothing cancels precisions (inverse-prior expert) and amplifies fp32 rounding: same
    # tolerance as the small-dim family's forward tests, against an fp64 oracle
    bad = helpers.compare_forward(ours, ref, rtol=5e-4, atol=5e-5)
    assert not bad, bad


@pytest.mark.parametrize('name', ['default32', 'c3_dims'])
def test_forward_tf32_fast_mode(name):
    lib = _lib.load()
    fx = helpers.large_case(**CASES[name])
    eps_flt, eps_smt = noise(fx, 5, 9)
    ours = helpers.run_forward_large(lib, fx, 'cuda', 'bfilter', True, 5, eps_flt, eps_smt, precision=1)
    ref = helpers.oracle_forward(fx, 'bfilter', True, 5, eps_flt, eps_smt, dtype=torch.float64)
    bad = helpers.compare_forward(ours, ref, rtol=2e-2, atol=2e-2)
    assert not bad, bad


def test_python_api_forward_large_and_no_training():
    fx = helpers.large_case(**CASES['default32'])
    m = models.MultiDMM(fx['modalities'], fx['dims'], h_dim=fx['h_dim'], z_dim=fx['z_dim'],
                        device=torch.device('cuda:0'))
    m.load_state_dict(fx['state_dict'])
    m.eval()
    eps_flt, eps_smt = noise(fx, 5, 7)
    inputs = {k: v.cuda() for k, v in fx['inputs'].items()}
    with torch.no_grad():
        infer, prior, recon = m(inputs, lengths=fx['lengths'], mode='fsmooth', flt_particles=5,
                                noise=(eps_flt.cuda(), eps_smt.cuda()))
    ref = helpers.oracle_forward(fx, 'fsmooth', True, 5, eps_flt, eps_smt, dtype=torch.float64)
    bad = helpers.compare_forward((infer, prior, recon), ref, rtol=5e-4, atol=5e-5)
    assert not bad, bad
    with pytest.raises(_lib.BfviError):                      # training kernels: small-dim family only
        m(inputs, lengths=fx['lengths'])
    with pytest.raises(_lib.BfviError):
        mask = torch.ones(max(fx['lengths']), len(fx['lengths']), 1, dtype=torch.bool, device='cuda')
        m.step(inputs, mask, 1.0, {}, lengths=fx['lengths'])
