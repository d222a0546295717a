"""multimodal_dmm_b200.metrics on the GPU: the fused SSIM kernel against the reference's golden outputs and
against the reference formulation (grouped conv2d) run with torch on the same device at the Weizmann
evaluation shape."""
import os

import pytest
import torch
import torch.nn.functional as F

from multimodal_dmm_b200 import metrics

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'metrics', 'ssim.pt'), weights_only=False)
TOL = 2e-5


@pytest.mark.parametrize('case', GOLD, ids=lambda c: c['name'])
def test_ssim_matches_reference_golden(case):
    ssim, cs = metrics.eval_ssim(case['x'].cuda(), case['y'].cuda(), win_size=case['win_size'],
                                 win_sigma=case['win_sigma'], data_range=case['data_range'], full=True)
    assert torch.allclose(ssim.cpu(), case['ssim'], rtol=0, atol=TOL), (ssim, case['ssim'])
    assert torch.allclose(cs.cpu(), case['cs'], rtol=0, atol=TOL)


def torch_ssim(x, y, win):
    """utils.py:93-163 restated with torch ops in float64 (the checker, not the product)."""
    c = x.shape[1]
    w = win.to(x).double().view(1, 1, 1, -1).repeat(5 * c, 1, 1, 1)
    z = torch.cat([x, y, x * x, y * y, x * y], 1).double()
    z = F.conv2d(z, w, groups=5 * c)
    z = F.conv2d(z.transpose(2, 3), w, groups=5 * c).transpose(2, 3)
    mu1, mu2, s1, s2, s12 = (z[:, i * c:(i + 1) * c] for i in range(5))
    s1, s2, s12 = s1 - mu1 ** 2, s2 - mu2 ** 2, s12 - mu1 * mu2
    cs = (2 * s12 + 0.03 ** 2) / (s1 + s2 + 0.03 ** 2)
    return (((2 * mu1 * mu2 + 0.01 ** 2) / (mu1 ** 2 + mu2 ** 2 + 0.01 ** 2)) * cs).mean((1, 2, 3))


def test_ssim_at_weizmann_eval_shape():
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.rand(625, 3, 64, 64, device='cuda', generator=g)               # T * B = 25 * 25 frames
    y = (x + 0.05 * torch.randn(x.shape, device='cuda', generator=g)).clamp(0, 1)
    got = metrics.eval_ssim(x, y)
    ref = torch_ssim(x, y, metrics._fspecial_gauss_1d(11, 1.5).view(-1))
    assert torch.allclose(got.double(), ref, rtol=0, atol=TOL)
    assert torch.allclose(metrics.eval_ssim(x, x), torch.ones(625, device='cuda'), atol=1e-6)
