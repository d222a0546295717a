"""Parity of the CUDA path (through the C ABI) against the reference's golden
outputs and against the oracle.  Tolerances are BASELINE.json's: ELBO 1e-4
relative, parameter gradients 1e-3 relative (per tensor, L2)."""
import pytest
import torch

import bfvi_oracle as bo
from conftest import golden_names, load_golden, rel_err
import helpers
from multimodal_dmm_b200 import _lib

SMALL = golden_names()        # 'medium_dims' (Z=16, H=48) is served by the large-dim family
ELBO_TOL, GRAD_TOL = 1e-4, 1e-3


@pytest.fixture(scope='module')
def lib():
    return _lib.load()


@pytest.mark.gpu
@pytest.mark.parametrize('name', SMALL)
def test_step_matches_reference_golden(lib, name):
    fx = load_golden(name)
    loss, grads, launches = helpers.run_step(lib, fx, 'cuda')
    assert launches > 0
    ref = fx['ref_loss_fp64']
    assert abs(loss - ref) / abs(ref) < ELBO_TOL, (loss, ref)
    n = float(sum(fx['lengths']))
    for k, g_ref in fx['ref_grads_fp64'].items():
        g = grads[k] / n
        if g_ref.norm() == 0:
            assert g.norm() == 0, k
        else:
            assert rel_err(g, g_ref) < GRAD_TOL, (k, rel_err(g, g_ref))


@pytest.mark.gpu
def test_forward_only_loss_equals_fwd_bwd_loss(lib):
    fx = load_golden('spirals_ragged')
    l1, _, _ = helpers.run_step(lib, fx, 'cuda', with_grad=True)
    l2, g, _ = helpers.run_step(lib, fx, 'cuda', with_grad=False)
    assert g is None and abs(l1 - l2) <= 1e-5 * abs(l1)


@pytest.mark.gpu
def test_philox_stream_matches_oracle_on_dumped_noise(lib):
    """Throughput mode: in-kernel Philox noise.  bfvi_dump_noise materialises the
    same stream; the oracle run on it must agree with the kernels."""
    import ctypes as C
    fx = load_golden('spirals_half_missing')
    kw = dict(fx['step_kwargs'])
    seed = 2024
    loss, grads, _ = helpers.run_step(lib, fx, 'cuda', noise=None, seed=seed)
    t_max, b_dim = fx['mask'].shape[:2]
    z, k_tr, k_m = fx['z_dim'], kw.get('train_particles', 25), kw.get('match_particles', 50)
    n_sets = len(bo.step_sets(len(fx['modalities']), kw.get('uni_loss', True)))

    def dump(stream_id, S, T, B, K):
        out = torch.empty(S, T, B, K, z, device='cuda')
        lib.call('bfvi_dump_noise', C.c_uint64(seed), stream_id, 0, S, T, B, K, z, _lib.ptr(out),
                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return out.cpu()
    noise = {'match': torch.stack([dump(100, 1, 1, 1, k_m).reshape(k_m, z),
                                   dump(101, 1, 1, 1, k_m).reshape(k_m, z)]),
             'filt': dump(1, n_sets, t_max, b_dim, 1),
             'sflt': dump(2, n_sets, t_max, b_dim, k_tr),
             'ssmt': dump(3, n_sets, t_max, b_dim, 1)}
    # the stream must look standard normal
    flat = noise['sflt'].flatten()
    assert abs(flat.mean()) < 0.05 and abs(flat.std() - 1) < 0.05
    params = {k: v.clone().double().requires_grad_(True) for k, v in fx['state_dict'].items()}
    orc = bo.OracleDMM(fx['modalities'], fx['dims'], params, h_dim=fx['h_dim'], z_dim=z,
                       min_std=fx['min_std'], draw=bo.step_noise_tape(noise))
    cast = lambda d: {k: v.double() for k, v in d.items()}
    ref = orc.step(cast(fx['inputs']), fx['mask'], fx['kld_mult'], fx['rec_mults'],
                   targets=cast(fx['targets']), lengths=fx['lengths'], **kw)
    ref.backward()
    assert abs(loss - ref.item()) / abs(ref.item()) < ELBO_TOL
    for k, p in params.items():
        assert rel_err(grads[k], p.grad) < GRAD_TOL, k
