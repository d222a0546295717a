"""Differential test on seeded random problems (tests/random_cases.py) on B200: both kernel
families vs the fp64 oracle — random modality counts, feature dims, ragged lengths, NaN patterns,
dropped modalities, all four mode combinations, particle counts that hit every lane geometry."""
import math

import pytest

import helpers
import random_cases as rc
from multimodal_dmm_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed', range(40))
def test_random_step_small_family(seed):
    fx = rc.make_case(1000 + seed)
    ref_loss, ref_grads = rc.oracle_step(fx)
    loss, grads, _ = helpers.run_step(_lib.load(), fx, 'cuda', kwargs=fx['step_kwargs'])
    if not math.isfinite(ref_loss):            # ill-posed draw: the reference returns NaN, and so must we
        assert not math.isfinite(loss)
        return
    bad = rc.check(loss, grads, ref_loss, ref_grads)
    assert not bad, (fx['step_kwargs'], fx['lengths'], bad[:4])


@pytest.mark.parametrize('seed', range(12))
def test_random_step_large_family(seed):
    fx = rc.make_case(2000 + seed, dims_zh=[(16, 24), (8, 40), (32, 32)][seed % 3])
    ref_loss, ref_grads = rc.oracle_step(fx)
    loss, grads, _ = helpers.run_step(_lib.load(), fx, 'cuda', kwargs=fx['step_kwargs'])
    if not math.isfinite(ref_loss):            # ill-posed draw: the reference returns NaN, and so must we
        assert not math.isfinite(loss)
        return
    bad = rc.check(loss, grads, ref_loss, ref_grads, grad_tol=1e-3)
    assert not bad, (fx['step_kwargs'], fx['lengths'], bad[:4])
