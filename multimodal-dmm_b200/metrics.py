"""Evaluation metrics of the reference's trainers on the device, same names and
signatures: `eval_ssim` (utils.py:165-212, used at weizmann.py:133,141 / vidTIMIT.py:122) and the
per-sequence MSE of `SpiralsTrainer.compute_metrics` (spirals.py:105-111, `multiseq.seq_mse`).

`eval_ssim` is ONE fused kernel (`bfvi_ssim`, csrc/bfvi_data.cuh) per call: the reference concatenates
X, Y, X², Y², XY into a (N, 5C, H, W) tensor, runs two grouped convolutions with transposes in between and a
dozen elementwise kernels over five blurred maps.  CUDA tensors only; no CPU fallback."""
import ctypes as C

import torch

from . import _lib
from .multiseq import _Runtime, seq_mse  # noqa: F401  (re-exported)


def _fspecial_gauss_1d(size, sigma):
    """1-D Gaussian window (utils.py:76-91), shape (1, 1, size)."""
    coords = torch.arange(size).to(dtype=torch.float)
    coords -= size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g.unsqueeze(0).unsqueeze(0)


def eval_ssim(X, Y, win_size=11, win_sigma=1.5, win=None, data_range=1.0, size_average=False, full=False):
    """SSIM of two (N, C, H, W) image batches, per image (utils.py:165-212).  `win`: optional 1-D window
    (any shape whose last axis holds the taps; one window for all channels, as `_fspecial_gauss_1d` gives)."""
    if len(X.shape) != 4:
        raise ValueError('Input images must 4-d tensor.')
    if not X.type() == Y.type():
        raise ValueError('Input images must have the same dtype.')
    if not X.shape == Y.shape:
        raise ValueError('Input images must have the same dimensions.')
    if not (win_size % 2 == 1):
        raise ValueError('Window size must be odd.')
    if win is None:
        win = _fspecial_gauss_1d(win_size, win_sigma)
    taps = win.detach().reshape(-1, win.shape[-1])[0].float().cpu()
    win_size = int(taps.numel())
    N, Cc, H, W = (int(v) for v in X.shape)
    with _Runtime(X.device) as rt:
        x = X.detach().contiguous().float()
        y = Y.detach().contiguous().float()
        ssim = torch.empty(N, dtype=torch.float32, device=x.device)
        cs = torch.empty(N, dtype=torch.float32, device=x.device)
        nbytes = int(rt.lib.dll.bfvi_ssim_scratch(N, Cc, H, W, win_size))
        if nbytes == 0:
            raise _lib.BfviError('eval_ssim: images (%d x %d) smaller than the %d-tap window' % (H, W, win_size))
        scratch = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device)
        w = (C.c_float * win_size)(*[float(v) for v in taps])
        rt.call('bfvi_ssim', _lib.ptr(x), _lib.ptr(y), N, Cc, H, W, w, win_size, C.c_float(float(data_range)),
                _lib.ptr(ssim), _lib.ptr(cs), _lib.ptr(scratch))
    if size_average:
        ssim, cs = ssim.mean(), cs.mean()
    return (ssim, cs) if full else ssim
