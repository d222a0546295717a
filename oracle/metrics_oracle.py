"""CPU restatement (numpy, float64 accumulation) of the reference's SSIM metric, utils.py:76-212 — TEST
INFRASTRUCTURE ONLY (imported by tests/, never by the product path multimodal-dmm_b200/metrics.py).

Parity status: PINNED — oracle/make_golden_ssim.py runs the unmodified reference `eval_ssim` (imported from
/root/reference with matplotlib stubbed) and this restatement on the same images, asserts
agreement (1e-5) and stores the reference's outputs in tests/golden/metrics/ssim.pt."""
import numpy as np


def fspecial_gauss_1d(size, sigma):
    """utils.py:76-91."""
    coords = np.arange(size, dtype=np.float32) - size // 2
    g = np.exp(-(coords ** 2) / np.float32(2 * sigma ** 2)).astype(np.float32)
    return g / g.sum()


def _blur(a, w):
    """Valid-padding separable blur along W then H (utils.py:93-108), float64."""
    k = len(w)
    n, c, h, wd = a.shape
    out = np.zeros((n, c, h, wd - k + 1))
    for i in range(k):
        out += w[i] * a[:, :, :, i:i + wd - k + 1]
    out2 = np.zeros((n, c, h - k + 1, wd - k + 1))
    for i in range(k):
        out2 += w[i] * out[:, :, i:i + h - k + 1, :]
    return out2


def eval_ssim(x, y, win_size=11, win_sigma=1.5, win=None, data_range=1.0):
    """utils.py:110-212 with size_average=False, full=True: per-image (ssim, cs)."""
    w = fspecial_gauss_1d(win_size, win_sigma).astype(np.float64) if win is None else np.asarray(win, dtype=np.float64)
    x, y = x.astype(np.float64), y.astype(np.float64)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _blur(x, w), _blur(y, w)
    s1, s2, s12 = _blur(x * x, w) - mu1 ** 2, _blur(y * y, w) - mu2 ** 2, _blur(x * y, w) - mu1 * mu2
    cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim_map = ((2 * mu1 * mu2 + c1) / (mu1 ** 2 + mu2 ** 2 + c1)) * cs_map
    return ssim_map.mean(axis=(1, 2, 3)), cs_map.mean(axis=(1, 2, 3))
