"""Kernel-logic check on CPU: the CUDA sources compiled against the SIMT emulator
(tests/emu) must reproduce the reference's golden step outputs.  Development
tool; the parity tests proper are the `-m gpu` ones."""
import pytest

from conftest import golden_names, load_golden, rel_err
import helpers

SMALL = golden_names()        # 'medium_dims' (Z=16, H=48) is served by the large-dim family


@pytest.fixture(scope='module')
def lib():
    return helpers.emu_library()


@pytest.mark.parametrize('name', SMALL)
def test_step_loss_and_grads(lib, name):
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


@pytest.mark.parametrize('lanes', ['1', '3', '8', '32'])
def test_lane_group_geometries_agree(lib, lanes, monkeypatch):
    """Every lanes-per-chain mapping of the particle kernels (bfvi_chain.cuh) must give
    the same step: exercised through the BFVI_LANES tuning knob."""
    fx = load_golden('spirals_ragged')
    monkeypatch.setenv('BFVI_LANES', lanes)
    loss, grads, _ = helpers.run_step(lib, fx, 'cpu')
    ref = fx['ref_loss_fp64']
    assert abs(loss - ref) / abs(ref) < 1e-4, (loss, ref)
    n = float(sum(fx['lengths']))
    for k, g_ref in fx['ref_grads_fp64'].items():
        if g_ref.norm() > 0:
            assert rel_err(grads[k] / n, g_ref) < 1e-3, (k, rel_err(grads[k] / n, g_ref))


@pytest.mark.parametrize('chunks', ['2', '3'])
def test_batch_chunked_step_agrees(lib, chunks, monkeypatch):
    """bfvi_step_fwd_bwd splits the batch into chunks on side streams (BFVI_CHUNKS knob):
    the chunked step must equal the whole-batch step."""
    fx = load_golden('spirals_half_missing')
    monkeypatch.setenv('BFVI_CHUNKS', chunks)
    loss, grads, _ = helpers.run_step(lib, fx, 'cpu')
    ref = fx['ref_loss_fp64']
    assert abs(loss - ref) / abs(ref) < 1e-4, (loss, ref)
    n = float(sum(fx['lengths']))
    for k, g_ref in fx['ref_grads_fp64'].items():
        if g_ref.norm() > 0:
            assert rel_err(grads[k] / n, g_ref) < 1e-3, (k, rel_err(grads[k] / n, g_ref))


@pytest.mark.parametrize('segments,chunks', [('2', '1'), ('3', '1'), ('7', '1'), ('4', '2')])
def test_time_segmented_backward_agrees(lib, segments, chunks, monkeypatch):
    """The particle backward kernel cuts its time loop into segments that hand the gradient carry
    over through scratch memory (BFVI_BWD_SEGMENTS knob; on the CPU one launch per segment): the
    step must equal the unsegmented one to rounding, also combined with batch chunks."""
    fx = load_golden('spirals_ragged')
    monkeypatch.setenv('BFVI_LANES', '5')            # throughput mapping (the one that is segmented)
    monkeypatch.setenv('BFVI_CHUNKS', chunks)
    loss0, grads0, _ = helpers.run_step(lib, fx, 'cpu')
    monkeypatch.setenv('BFVI_BWD_SEGMENTS', segments)
    loss, grads, _ = helpers.run_step(lib, fx, 'cpu')
    assert loss == loss0
    for k, g0 in grads0.items():
        if g0.norm() > 0:
            assert rel_err(grads[k], g0) < 2e-6, (k, rel_err(grads[k], g0))
    ref = fx['ref_loss_fp64']
    assert abs(loss - ref) / abs(ref) < 1e-4
    n = float(sum(fx['lengths']))
    for k, g_ref in fx['ref_grads_fp64'].items():
        if g_ref.norm() > 0:
            assert rel_err(grads[k] / n, g_ref) < 1e-3, (k, rel_err(grads[k] / n, g_ref))
