"""Differential test on seeded random problems (tests/random_cases.py): the kernels compiled
against the SIMT emulator vs the fp64 oracle.  The GPU twin is tests/test_gpu_random.py."""
import math

import pytest

import helpers
import random_cases as rc


@pytest.fixture(scope='module')
def lib():
    return helpers.emu_library()


@pytest.mark.parametrize('seed', range(12))
def test_random_step_matches_oracle(lib, seed):
    fx = rc.make_case(100 + seed)
    ref_loss, ref_grads = rc.oracle_step(fx)
    loss, grads, _ = helpers.run_step(lib, fx, 'cpu', kwargs=fx['step_kwargs'])
    if not math.isfinite(ref_loss):            # ill-posed draw: the reference returns NaN, and so must we
        assert not math.isfinite(loss)
        return
    bad = rc.check(loss, grads, ref_loss, ref_grads)
    assert not bad, (fx['step_kwargs'], fx['lengths'], bad[:4])
