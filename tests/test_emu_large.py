"""Large-dim family on CPU: the launch sequence and the elementwise kernels of bfvi_forward
(SIMT emulator build; its GEMM stand-in is an exact fp32 loop) against the oracle on the same
injected noise, all four forward() modes, NaN-deleted / padded / dropped inputs."""
import pytest
import torch

import helpers


@pytest.fixture(scope='module')
def lib():
    return helpers.emu_library()


@pytest.mark.parametrize('mode', ['bfilter', 'ffilter', 'fsmooth', 'bsmooth'])
@pytest.mark.parametrize('sample,k_flt', [(False, 1), (True, 1), (True, 4)])
def test_forward_matches_oracle(lib, mode, sample, k_flt):
    fx = helpers.large_case(z_dim=7, h_dim=10, dims=[3, 2], t_max=6, lengths=[6, 6, 5, 3, 2], seed=11)
    g = torch.Generator().manual_seed(5)
    t_max, b_dim, z = 6, 5, 7
    eps_flt = torch.randn(t_max, b_dim, k_flt, z, generator=g)
    eps_smt = torch.randn(t_max, b_dim, 1, z, generator=g)
    ours = helpers.run_forward_large(lib, fx, 'cpu', mode, sample, k_flt, eps_flt, eps_smt)
    ref = helpers.oracle_forward(fx, mode, sample, k_flt, eps_flt, eps_smt, dtype=torch.float32)
    bad = helpers.compare_forward(ours, ref, rtol=2e-4, atol=2e-5)
    assert not bad, bad


def test_forward_with_a_dropped_modality(lib):
    fx = helpers.large_case(z_dim=9, h_dim=12, dims=[2, 4, 1], t_max=5, lengths=[5, 4, 4], seed=3, drop=('m1',))
    g = torch.Generator().manual_seed(6)
    eps_flt = torch.randn(5, 3, 3, 9, generator=g)
    eps_smt = torch.randn(5, 3, 1, 9, generator=g)
    ours = helpers.run_forward_large(lib, fx, 'cpu', 'fsmooth', True, 3, eps_flt, eps_smt)
    ref = helpers.oracle_forward(fx, 'fsmooth', True, 3, eps_flt, eps_smt, dtype=torch.float32)
    assert set(ours[2]) == {'m0', 'm1', 'm2'}                 # every decoder runs, models/dmm.py:207
    bad = helpers.compare_forward(ours, ref, rtol=2e-4, atol=2e-5)
    assert not bad, bad


from conftest import golden_names, load_golden, rel_err

SMALL = [n for n in golden_names() if n != 'medium_dims']


@pytest.mark.parametrize('name', SMALL)
def test_training_step_of_the_large_family_on_golden_fixtures(lib, name, monkeypatch):
    """BFVI_FAMILY=2 routes the small golden models through the large-dim launch sequence
    (GEMMs + elementwise kernels): loss and every parameter gradient must match the reference."""
    monkeypatch.setenv('BFVI_FAMILY', '2')
    fx = load_golden(name)
    loss, grads, launches = helpers.run_step(lib, fx, 'cpu')
    assert launches > 0
    ref = fx['ref_loss_fp64']
    assert abs(loss - ref) / abs(ref) < 1e-4, (loss, ref)
    n = float(sum(fx['lengths']))
    for k, g_ref in fx['ref_grads_fp64'].items():
        g = grads[k] / n
        if g_ref.norm() == 0:
            assert g.norm() == 0, k
        else:
            assert rel_err(g, g_ref) < 1e-3, (k, rel_err(g, g_ref))
