"""Loss functions with the reference signatures (models/losses.py:14-89).

kld_gauss and nll_gauss run as CUDA kernels (bfvi_kld_*, bfvi_nll_gauss_*) wrapped
in autograd Functions; the Bernoulli / categorical likelihoods (Weizmann-shaped
models only) are device-side tensor ops with the reference's exact semantics,
including its quirk of feeding probabilities to nll_loss."""
import ctypes as C

import torch
import torch.nn.functional as F

from .. import _lib


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.BfviError('BFVI losses run on CUDA tensors only (no CPU fallback)')


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _row_mask(mask, lead_shape):
    """(T,B[,1...]) mask -> flat uint8 row mask, or None."""
    if mask is None:
        return None
    m = mask
    while m.dim() > len(lead_shape) and m.shape[-1] == 1:
        m = m.squeeze(-1)
    if tuple(m.shape) != tuple(lead_shape):
        return False
    return m.to(torch.uint8).contiguous().reshape(-1)


class _Kld(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m1, s1, m2, s2, rmask, z_dim):
        lib = _lib.load()
        t = [x.detach().contiguous().float() for x in (m1, s1, m2, s2)]
        rows = t[0].numel() // z_dim
        out = torch.zeros(1, dtype=torch.float64, device=t[0].device)
        lib.call('bfvi_kld_fwd', *[_lib.ptr(x) for x in t], _lib.ptr(rmask), rows, z_dim,
                 _lib.ptr(out), _stream())
        ctx.save_for_backward(*t)
        ctx.rmask, ctx.z_dim, ctx.shapes = rmask, z_dim, [x.shape for x in (m1, s1, m2, s2)]
        return out.sum().to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        t = ctx.saved_tensors
        rows = t[0].numel() // ctx.z_dim
        outs = [torch.empty_like(x) for x in t]
        lib.call('bfvi_kld_bwd', *[_lib.ptr(x) for x in t], _lib.ptr(ctx.rmask), rows, ctx.z_dim,
                 C.c_float(1.0), *[_lib.ptr(x) for x in outs], _stream())
        return tuple(o.reshape(s) * g for o, s in zip(outs, ctx.shapes)) + (None, None)


def kld_gauss(mean_1, std_1, mean_2, std_2, mask=None):
    """0.5 * sum over (masked) elements of the Gaussian KL, models/losses.py:14-21."""
    _require_cuda(mean_1, std_1, mean_2, std_2, mask)
    shape = torch.broadcast_shapes(mean_1.shape, std_1.shape, mean_2.shape, std_2.shape)
    ex = [x.expand(shape) for x in (mean_1, std_1, mean_2, std_2)]
    rmask = _row_mask(mask, shape[:-1])
    if rmask is False:      # unusual mask shape: elementwise mask, device ops
        el = (2 * torch.log(ex[3]) - 2 * torch.log(ex[1])
              + (ex[1].pow(2) + (ex[0] - ex[2]).pow(2)) / ex[3].pow(2) - 1)
        return 0.5 * el.masked_select(mask.bool()).sum()
    return _Kld.apply(ex[0], ex[1], ex[2], ex[3], rmask, shape[-1])


class _NllGauss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, std, x, rmask, rows):
        lib = _lib.load()
        t = [v.detach().contiguous().float() for v in (mean, std, x)]
        d = t[0].numel() // rows
        out = torch.zeros(1, dtype=torch.float64, device=t[0].device)
        lib.call('bfvi_nll_gauss_fwd', *[_lib.ptr(v) for v in t], _lib.ptr(rmask), rows, d,
                 _lib.ptr(out), _stream())
        ctx.save_for_backward(*t)
        ctx.rmask, ctx.rows, ctx.shapes = rmask, rows, (mean.shape, std.shape)
        return out.sum().to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        mean, std, x = ctx.saved_tensors
        d = mean.numel() // ctx.rows
        dm, ds = torch.empty_like(mean), torch.empty_like(std)
        lib.call('bfvi_nll_gauss_bwd', _lib.ptr(mean), _lib.ptr(std), _lib.ptr(x), _lib.ptr(ctx.rmask),
                 ctx.rows, d, C.c_float(1.0), _lib.ptr(dm), _lib.ptr(ds), _stream())
        return dm.reshape(ctx.shapes[0]) * g, ds.reshape(ctx.shapes[1]) * g, None, None, None


def nll_gauss(mean, std, x, mask=None):
    """Gaussian NLL summed over observed (non-NaN) in-sequence elements,
    models/losses.py:68-89.  `mask` broadcasts from the leading dims of x."""
    _require_cuda(mean, std, x, mask)
    if mask is None:
        rmask, rows = None, x.numel()
    else:
        m = mask
        while m.dim() > 2 and m.shape[-1] == 1:
            m = m.squeeze(-1)
        rmask = _row_mask(m, x.shape[:m.dim()])
        if rmask is False:
            raise _lib.BfviError('nll_gauss: mask %s does not broadcast over x %s'
                                 % (tuple(mask.shape), tuple(x.shape)))
        rows = rmask.numel()
    return _NllGauss.apply(mean.expand(x.shape), std.expand(x.shape), x, rmask, rows)


def _elem_mask(x, mask):
    obs = ~torch.isnan(x)
    if mask is None:
        return obs
    shape = list(mask.shape) + [1] * (x.dim() - mask.dim())
    return obs & mask.bool().view(*shape)


def nll_bernoulli(theta, x, mask=None):
    """models/losses.py:23-42."""
    _require_cuda(theta, x, mask)
    keep = _elem_mask(x, mask)
    return F.binary_cross_entropy(theta.masked_select(keep), x.masked_select(keep), reduction='sum')


def nll_categorical(probs, x, mask=None):
    """models/losses.py:44-66 (value is -sum p[label]; quirk preserved)."""
    _require_cuda(probs, x, mask)
    keep = _elem_mask(x, mask)
    cols = [probs[:, :, k:k + 1].masked_select(keep) for k in range(probs.shape[2])]
    return F.nll_loss(torch.stack(cols, dim=-1), x.masked_select(keep).long(), reduction='sum')
