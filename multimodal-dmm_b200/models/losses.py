"""Loss functions with the reference signatures (models/losses.py:14-89).

Every loss runs as a CUDA kernel of libbfvi_b200 (bfvi_kld_*, bfvi_nll_gauss_*,
bfvi_nll_bernoulli_*, bfvi_nll_categorical_*) wrapped in an autograd Function, with
the reference's exact semantics, including its quirk of feeding probabilities to
nll_loss in the categorical likelihood."""
import ctypes as C

import torch

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


def _lead_mask(mask, x, what):
    """Row mask over the leading (T, B) dims of x as flat uint8 (None = every row)."""
    if mask is None:
        return None, x.shape[:2]
    m = mask
    while m.dim() > 2 and m.shape[-1] == 1:
        m = m.squeeze(-1)
    rmask = _row_mask(m, x.shape[:m.dim()])
    if rmask is False or m.dim() != 2:
        raise _lib.BfviError('%s: mask %s does not broadcast over x %s' % (what, tuple(mask.shape), tuple(x.shape)))
    return rmask, x.shape[:2]


class _NllBernoulli(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, x, rmask, rows):
        lib = _lib.load()
        th, xv = theta.detach().contiguous().float(), x.detach().contiguous().float()
        d = xv.numel() // rows
        out = torch.zeros(1, dtype=torch.float64, device=th.device)
        lib.call('bfvi_nll_bernoulli_fwd', _lib.ptr(th), _lib.ptr(xv), _lib.ptr(rmask), rows, d,
                 _lib.ptr(out), _stream())
        ctx.save_for_backward(th, xv)
        ctx.rmask, ctx.rows, ctx.shape = rmask, rows, theta.shape
        return out.sum().to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        th, xv = ctx.saved_tensors
        d_th = torch.empty_like(th)
        lib.call('bfvi_nll_bernoulli_bwd', _lib.ptr(th), _lib.ptr(xv), _lib.ptr(ctx.rmask), ctx.rows,
                 xv.numel() // ctx.rows, C.c_float(1.0), _lib.ptr(d_th), _stream())
        return d_th.reshape(ctx.shape) * g, None, None, None


def nll_bernoulli(theta, x, mask=None):
    """Bernoulli NLL (binary cross-entropy, summed) over observed in-sequence elements,
    models/losses.py:23-42; theta, x: (T, B, D, ...), mask: (T, B[, 1...])."""
    _require_cuda(theta, x, mask)
    rmask, lead = _lead_mask(mask, x, 'nll_bernoulli')
    rows = int(lead[0]) * int(lead[1])
    if rows == 0 or x.numel() == 0:
        return theta.sum() * 0.0
    return _NllBernoulli.apply(theta.expand(x.shape), x, rmask, rows)


class _NllCategorical(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, x, rmask, rows):
        lib = _lib.load()
        pr, xv = probs.detach().contiguous().float(), x.detach().contiguous().float()
        n_cat = pr.numel() // rows
        out = torch.zeros(1, dtype=torch.float64, device=pr.device)
        lib.call('bfvi_nll_categorical_fwd', _lib.ptr(pr), _lib.ptr(xv), _lib.ptr(rmask), rows, n_cat,
                 _lib.ptr(out), _stream())
        ctx.save_for_backward(pr, xv)
        ctx.rmask, ctx.rows, ctx.n_cat, ctx.shape = rmask, rows, n_cat, probs.shape
        return out.sum().to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        pr, xv = ctx.saved_tensors
        d_pr = torch.empty_like(pr)
        lib.call('bfvi_nll_categorical_bwd', _lib.ptr(pr), _lib.ptr(xv), _lib.ptr(ctx.rmask), ctx.rows,
                 ctx.n_cat, C.c_float(1.0), _lib.ptr(d_pr), _stream())
        return d_pr.reshape(ctx.shape) * g, None, None, None


def nll_categorical(probs, x, mask=None):
    """Categorical NLL as the reference computes it, models/losses.py:44-66: F.nll_loss fed with
    probabilities, i.e. -sum p[label] over observed in-sequence rows (quirk preserved).
    probs: (T, B, K[, 1]), x: (T, B[, 1]) float labels (NaN = missing).  Like the reference, only a
    scalar label per (t, b) is meaningful (its mask/probs broadcast fails for wider label tensors)."""
    _require_cuda(probs, x, mask)
    rmask, lead = _lead_mask(mask, x, 'nll_categorical')
    rows = int(lead[0]) * int(lead[1])
    if rows == 0:
        return probs.sum() * 0.0
    if x.numel() != rows or probs.numel() % rows != 0 or probs.shape[:2] != x.shape[:2]:
        raise _lib.BfviError('nll_categorical: one label per (t, b) expected; probs %s, x %s'
                             % (tuple(probs.shape), tuple(x.shape)))
    return _NllCategorical.apply(probs, x, rmask, rows)
