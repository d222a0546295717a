"""Evaluation metrics on CPU: the numpy restatement oracle/metrics_oracle.py against the
golden outputs of the unmodified reference `eval_ssim` (oracle/make_golden_ssim.py), and the library's fused
SSIM kernel (emulated) behind multimodal_dmm_b200.metrics.eval_ssim against the same fixtures."""
import os

import numpy as np
import pytest
import torch

import metrics_oracle as orc
from multimodal_dmm_b200 import _lib, metrics, multiseq
from test_emu_multiseq import EmuRuntime

GOLD = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'metrics', 'ssim.pt'), weights_only=False)
TOL = 2e-5          # fp32 blur in another summation order; sigma = E[x^2] - mu^2 cancels


@pytest.fixture
def emu(monkeypatch):
    monkeypatch.setattr(multiseq, '_Runtime', EmuRuntime)
    monkeypatch.setattr(metrics, '_Runtime', EmuRuntime)


@pytest.mark.parametrize('case', GOLD, ids=lambda c: c['name'])
def test_ssim_oracle_matches_reference_golden(case):
    ssim, cs = orc.eval_ssim(case['x'].numpy(), case['y'].numpy(), case['win_size'], case['win_sigma'],
                             data_range=case['data_range'])
    assert np.allclose(ssim, case['ssim'].numpy(), rtol=0, atol=1e-5)
    assert np.allclose(cs, case['cs'].numpy(), rtol=0, atol=1e-5)


@pytest.mark.parametrize('case', GOLD, ids=lambda c: c['name'])
def test_ssim_kernel_matches_reference_golden(emu, case):
    ssim, cs = metrics.eval_ssim(case['x'], case['y'], win_size=case['win_size'], win_sigma=case['win_sigma'],
                                 data_range=case['data_range'], full=True)
    assert torch.allclose(ssim, case['ssim'], rtol=0, atol=TOL), (ssim, case['ssim'])
    assert torch.allclose(cs, case['cs'], rtol=0, atol=TOL)
    only = metrics.eval_ssim(case['x'], case['y'], win_size=case['win_size'], win_sigma=case['win_sigma'],
                             data_range=case['data_range'], size_average=True)
    assert abs(only.item() - case['ssim'].mean().item()) < TOL


def test_ssim_argument_errors(emu):
    x = torch.zeros(2, 1, 16, 16)
    with pytest.raises(ValueError):
        metrics.eval_ssim(x[0], x[0])
    with pytest.raises(ValueError):
        metrics.eval_ssim(x, x, win_size=10)
    with pytest.raises(ValueError):
        metrics.eval_ssim(x, x[:, :, :8])
    with pytest.raises(_lib.BfviError):
        metrics.eval_ssim(x[:, :, :8, :8], x[:, :, :8, :8])          # smaller than the window
    same = metrics.eval_ssim(x + 0.5, x + 0.5)
    assert torch.allclose(same, torch.ones(2), atol=1e-6)            # identical images: SSIM = 1
