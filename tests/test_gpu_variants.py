"""The kernel VARIANTS bench.py times, under GPU parity.

Small batches (every golden fixture, every random case) dispatch the latency mappings
(`chain_fwd<Z,H,1>`, the 2-warp `chain_bwd<Z,H,0>`, the z-split single-particle kernels).  The
throughput mappings — 5 particles per lane, the time-segmented COOPERATIVE backward with its
spin-wait hand-over flags, batch chunks on side streams, the one-chain-per-lane K=1 backward —
only dispatch above 148*16 chains.  Here they run on B200

 (i)  forced through the tuning knobs on the reference's golden fixtures, and
 (ii) by natural dispatch at BASELINE shapes: C1 exactly (spirals.py:31-50 defaults: T=100, B=100,
      K=25) and the C2 shape (B=2400 and the full B=4096, 50 % missing) in Philox mode against the
      fp64 oracle on the dumped stream,

and every test asserts through bfvi_last_dispatch() WHICH kernel ran.
"""
import ctypes as C
import os
import sys

import pytest
import torch

import bfvi_oracle as bo
from conftest import ROOT, golden_names, load_golden, rel_err
import helpers
from multimodal_dmm_b200 import _lib

pytestmark = pytest.mark.gpu
ELBO_TOL, GRAD_TOL = 1e-4, 1e-3
SMALL = [n for n in golden_names() if n != 'medium_dims']


@pytest.fixture(scope='module')
def lib():
    return _lib.load()


def check_golden(fx, loss, grads):
    ref = fx['ref_loss_fp64']
    assert abs(loss - ref) / abs(ref) < ELBO_TOL, (loss, ref)
    n = float(sum(fx['lengths']))
    for k, g_ref in fx['ref_grads_fp64'].items():
        if g_ref.norm() > 0:
            assert rel_err(grads[k] / n, g_ref) < GRAD_TOL, (k, rel_err(grads[k] / n, g_ref))


@pytest.mark.parametrize('name', SMALL)
@pytest.mark.parametrize('segments,chunks', [('2', '1'), ('4', '1'), ('2', '2'), ('3', '2')])
def test_throughput_mapping_cooperative_segments_on_golden(lib, name, segments, chunks, monkeypatch):
    """chain_fwd<Z,H,5> + time-segmented cooperative chain_bwd<Z,H,1> (+ batch chunks) against the
    REFERENCE's golden loss / gradients."""
    fx = load_golden(name)
    if int(chunks) > fx['mask'].shape[1] or int(segments) > fx['mask'].shape[0]:
        pytest.skip('fixture smaller than the knob')
    z, h = fx['z_dim'], fx['h_dim']
    monkeypatch.setenv('BFVI_LANES', '5')
    monkeypatch.setenv('BFVI_BWD_SEGMENTS', segments)
    monkeypatch.setenv('BFVI_CHUNKS', chunks)
    loss, grads, _ = helpers.run_step(lib, fx, 'cuda')
    ran = lib.last_dispatch()
    k_train = fx['step_kwargs'].get('train_particles', 25)
    if k_train > 1:
        assert any(d.startswith('chain_fwd<%d,%d,5> lanes=5' % (z, h)) for d in ran), ran
        assert any(d.startswith('chain_bwd<%d,%d,1> K=%d lanes=5' % (z, h, k_train)) for d in ran), ran
        assert 'segmented:cooperative seg=%s' % segments in ran, ran
    assert 'step:chunks=%s' % chunks in ran, ran
    check_golden(fx, loss, grads)


def test_per_launch_segments_equal_cooperative(lib, monkeypatch):
    """BFVI_COOPERATIVE=0 (one stream-ordered launch per segment) and the cooperative launch run the
    same (task, segment) work items: bit-identical loss, gradients to rounding."""
    fx = load_golden('spirals_ragged')
    monkeypatch.setenv('BFVI_LANES', '5')
    monkeypatch.setenv('BFVI_BWD_SEGMENTS', '3')
    l1, g1, _ = helpers.run_step(lib, fx, 'cuda')
    assert 'segmented:cooperative seg=3' in lib.last_dispatch()
    monkeypatch.setenv('BFVI_COOPERATIVE', '0')
    l0, g0, _ = helpers.run_step(lib, fx, 'cuda')
    assert 'segmented:per-launch seg=3' in lib.last_dispatch()
    assert l0 == l1
    for k in g0:
        if g0[k].norm() > 0:
            assert rel_err(g1[k], g0[k]) < 2e-6, k


# ---------------------------------------------------------------------------------------
# natural dispatch at BASELINE shapes, Philox noise, fp64 oracle on the dumped stream
# ---------------------------------------------------------------------------------------
def spirals_fixture(b_dim, t_max, corrupt, seed):
    """C1 / C2 data exactly as bench.py builds it (datasets/spirals.py + multiseq.py semantics)."""
    sys.path.insert(0, ROOT)
    import bench
    if corrupt:
        inputs, targets, mask, lengths = bench.make_c2_batch(b_dim, t_max=t_max, seed=seed)
        rec = {m: 1.0 for m in bench.C2.mods}
    else:
        inputs, targets, mask, lengths = bench.make_c1_batch(b_dim, t_max=t_max, seed=seed)
        rec = {m: 0.5 for m in bench.C2.mods}
    mods, dims = list(bench.C2.mods), list(bench.C2.dims)
    return dict(modalities=mods, dims=dims, z_dim=5, h_dim=20, min_std=1e-3, inputs=inputs, targets=targets,
                mask=mask, lengths=lengths, kld_mult=1.0, rec_mults=rec,
                step_kwargs={'train_particles': 25, 'match_particles': 50},
                state_dict=bo.init_params(mods, dims, h_dim=20, z_dim=5, seed=1))


def dump_noise(lib, fx, seed):
    kw = fx['step_kwargs']
    t_max, b_dim = fx['mask'].shape[:2]
    z, k_tr, k_m = fx['z_dim'], kw['train_particles'], kw['match_particles']
    n_sets = len(bo.step_sets(len(fx['modalities']), kw.get('uni_loss', True)))

    def dump(stream_id, S, T, B, K):
        out = torch.empty(S, T, B, K, z, device='cuda')
        lib.call('bfvi_dump_noise', C.c_uint64(seed), stream_id, 0, S, T, B, K, z, _lib.ptr(out),
                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return out.cpu()
    return {'match': torch.stack([dump(100, 1, 1, 1, k_m).reshape(k_m, z), dump(101, 1, 1, 1, k_m).reshape(k_m, z)]),
            'filt': dump(1, n_sets, t_max, b_dim, 1), 'sflt': dump(2, n_sets, t_max, b_dim, k_tr),
            'ssmt': dump(3, n_sets, t_max, b_dim, 1)}


def oracle_on(fx, noise):
    torch.set_num_threads(os.cpu_count() or 1)
    params = {k: v.clone().double().requires_grad_(True) for k, v in fx['state_dict'].items()}
    orc = bo.OracleDMM(fx['modalities'], fx['dims'], params, h_dim=fx['h_dim'], z_dim=fx['z_dim'],
                       min_std=fx['min_std'], draw=bo.step_noise_tape({k: v.double() for k, v in noise.items()}))
    cast = lambda d: {k: v.double() for k, v in d.items()}
    ref = orc.step(cast(fx['inputs']), fx['mask'], fx['kld_mult'], fx['rec_mults'], targets=cast(fx['targets']),
                   lengths=fx['lengths'], **fx['step_kwargs'])
    ref.backward()
    return ref.item(), {k: p.grad for k, p in params.items()}


def test_c1_exact_shape_against_oracle(lib):
    """BASELINE configs[0]: M=2, Z=5, H=20, T=100, B=100, K=25, burst_delete(0.1) on the inputs — a
    100-step chain through the inverse-prior cancellation, at the reference's default batch."""
    fx = spirals_fixture(100, 100, corrupt=False, seed=1)
    loss, grads, _ = helpers.run_step(lib, fx, 'cuda', noise=None, seed=77)
    ran = lib.last_dispatch()
    assert any(d.startswith('chain_fwd<5,20,1>') for d in ran), ran          # 300 chains: latency mapping
    ref, ref_grads = oracle_on(fx, dump_noise(lib, fx, 77))
    assert abs(loss - ref) / abs(ref) < ELBO_TOL, (loss, ref)
    for k, g in ref_grads.items():
        assert rel_err(grads[k], g) < GRAD_TOL, (k, rel_err(grads[k], g))


@pytest.mark.parametrize('b_dim', [2400, 4096])
def test_c2_shape_natural_dispatch_against_oracle(lib, b_dim):
    """BASELINE configs[1] shape (T=100, 50 % uniform missing + burst, K=25) above the 148*16-chain
    threshold: the dispatcher picks by itself what BENCH times — chain_fwd<5,20,5>, the packed
    particle backward, and (past 8 warps/SM of z-split work) the one-chain-per-lane K=1 backward.
    B=4096 is bench.py's C2 line itself."""
    fx = spirals_fixture(b_dim, 100, corrupt=True, seed=1)
    loss, grads, _ = helpers.run_step(lib, fx, 'cuda', noise=None, seed=2024)
    ran = lib.last_dispatch()
    assert any(d.startswith('chain_fwd<5,20,5>') for d in ran), ran
    assert any(d.startswith('chain_bwd<5,20,') and ' K=25 ' in d and 'warps=4' in d for d in ran), ran
    assert any(d.startswith('chain_bwd<5,20,0> K=1 ') for d in ran), ran      # one chain per lane, K = 1
    if b_dim == 4096:
        assert any(d.startswith('segmented:cooperative') for d in ran), ran   # 46.6 % of the BENCH step
    ref, ref_grads = oracle_on(fx, dump_noise(lib, fx, 2024))
    assert abs(loss - ref) / abs(ref) < ELBO_TOL, (loss, ref)
    for k, g in ref_grads.items():
        assert rel_err(grads[k], g) < GRAD_TOL, (k, rel_err(grads[k], g))
