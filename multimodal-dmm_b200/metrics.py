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
    rows = win.detach().reshape(-1, win.shape[-1]).float().cpu()
    taps = rows[0]
    if rows.shape[0] > 1 and not bool((rows == taps[None]).all()):
        # the reference applies row c of `win` to channel c (grouped conv, utils.py:119-124); the kernel takes one
        # window for all channels, which is what every caller in the reference passes
        raise _lib.BfviError('eval_ssim: per-channel windows that differ are not supported by the fused kernel')
    win_size = int(taps.numel())
    N, Cc, H, W = (int(v) for v in X.shape)
    with _Runtime(X.device) as rt:
        x = X.detach().contiguous().float()
        y = Y.detach().contiguous().float()
        ssim = torch.empty(N, dtype=torch.float32, device=x.device)
        cs = torch.empty(N, dtype=torch.float32, device=x.device)
        w = (C.c_float * win_size)(*[float(v) for v in taps])
        # the kernel maps (image, channel) to grid dimensions capped at 65 535: the trainers call this on
        # flattened T*B frame batches (weizmann.py:133), so walk N in chunks of that size
        n_max = 65535
        for n0 in range(0, max(N, 1), n_max):
            n = min(n_max, N - n0)
            if n <= 0:
                break
            nbytes = int(rt.lib.dll.bfvi_ssim_scratch(n, Cc, H, W, win_size))
            if nbytes == 0:
                raise _lib.BfviError('eval_ssim: images (%d x %d) smaller than the %d-tap window, or more than '
                                     '65535 channels' % (H, W, win_size))
            scratch = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device)
            rt.call('bfvi_ssim', _lib.ptr(x[n0:n0 + n]), _lib.ptr(y[n0:n0 + n]), n, Cc, H, W, w, win_size,
                    C.c_float(float(data_range)), _lib.ptr(ssim[n0:n0 + n]), _lib.ptr(cs[n0:n0 + n]),
                    _lib.ptr(scratch))
    if size_average:
        ssim, cs = ssim.mean(), cs.mean()
    return (ssim, cs) if full else ssim
