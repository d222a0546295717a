"""Parameter containers with the reference's module tree and state_dict keys
(models/common.py:9-68).  In the fused path these modules only OWN the weights
(aliased onto the flat CUDA parameter buffer); their `forward` methods exist for
composition with custom user modules and run as ordinary device ops."""
import ctypes as C
import functools
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


class CategoricalMLP(nn.Module):
    """in -> hidden -> softmax probabilities (keys: in_to_h.0.*, h_to_out.0.*)."""

    def __init__(self, in_dim, out_dim, h_dim):
        super().__init__()
        self.in_to_h = nn.Sequential(nn.Linear(in_dim, h_dim), nn.ReLU())
        self.h_to_out = nn.Sequential(nn.Linear(h_dim, out_dim), nn.Softmax(dim=1))

    def forward(self, x):
        return (self.h_to_out(self.in_to_h(x)),)


class GaussianMLP(nn.Module):
    """in -> hidden -> (mean, softplus std + min_std)
    (keys: in_to_h.0.*, h_to_mean.*, h_to_std.0.*)."""

    def __init__(self, in_dim, out_dim, h_dim, min_std=1e-3):
        super().__init__()
        self.min_std = min_std
        self.in_to_h = nn.Sequential(nn.Linear(in_dim, h_dim), nn.ReLU())
        self.h_to_mean = nn.Linear(h_dim, out_dim)
        self.h_to_std = nn.Sequential(nn.Linear(h_dim, out_dim), nn.Softplus())

    def forward(self, x):
        h = self.in_to_h(x)
        return self.h_to_mean(h), self.h_to_std(h) + self.min_std


class GaussianGTF(nn.Module):
    """Gated transition function (keys: z_to_gate.{0,2}.*, z_lin.*,
    z_nonlin.{0,2}.*, z_to_std.0.*)."""

    def __init__(self, z_dim, h_dim, min_std=0):
        super().__init__()
        self.min_std = min_std
        self.z_to_gate = nn.Sequential(nn.Linear(z_dim, h_dim), nn.ReLU(),
                                       nn.Linear(h_dim, z_dim), nn.Sigmoid())
        self.z_lin = nn.Linear(z_dim, z_dim)
        self.z_nonlin = nn.Sequential(nn.Linear(z_dim, h_dim), nn.ReLU(),
                                      nn.Linear(h_dim, z_dim))
        self.z_to_std = nn.Sequential(nn.Linear(z_dim, z_dim), nn.Softplus())

    def forward(self, z):
        gate = self.z_to_gate(z)
        nonlin = self.z_nonlin(z)
        mean = torch.lerp(self.z_lin(z), nonlin, gate)
        return mean, self.z_to_std(nonlin) + self.min_std


# ---------------------------------------------------------------------------------------
# Image modules of the Weizmann / vidTIMIT models (models/common.py:70-175).  The caller injects
# them as custom `encoders=` / `decoders=` (weizmann.py:64-76 looks them up as
# `models.common.ImageEncoder / ImageDecoder`).  The nn.Conv2d / nn.ConvTranspose2d /
# nn.BatchNorm2d / nn.Linear sub-modules OWN the weights and buffers under the reference's
# state_dict keys — including the doubly registered `conv` / `net.0` entry — and, on a CUDA
# device, every layer runs through this library's kernels (include/bfvi.h: bfvi_conv_*,
# bfvi_bn2d_*, bfvi_sigmoid_bwd, bfvi_dense_*): no cuDNN, no cuBLAS.
# On CPU tensors (constructing a model, state_dict round trips, the reference-side tests)
# the sub-modules run as the plain torch modules they are.
# ---------------------------------------------------------------------------------------
# layer kinds served by libbfvi_b200 on CUDA tensors; BFVI_IMAGE_KERNELS="conv" / "" is a measurement aid (the torch /
# cuDNN modules beside the kernels, tools/time_conv.py)
IMAGE_KERNELS = {k: k in os.environ.get('BFVI_IMAGE_KERNELS', 'conv,dense').split(',') for k in ('conv', 'dense')}


def _library():
    return _lib.load()


def _use_kernels(x, kind):
    return x.is_cuda and IMAGE_KERNELS[kind]


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0)


def _scratch(lib, channels, like):
    n = lib.dll.bfvi_chan_scratch(channels)
    return torch.empty(n // 8 + 1, dtype=torch.float64, device=like.device), C.c_size_t(n)


def _on_tensor_device(fn):
    """Kernels launch on the CURRENT device: make that the device of the call's first tensor (a model on cuda:1
    while cuda:0 is current must not launch on cuda:0 with cuda:1 pointers)."""
    @functools.wraps(fn)
    def wrapped(ctx, *args):
        t = next((a for a in args if torch.is_tensor(a)), None)
        if t is not None and t.is_cuda:
            with torch.cuda.device(t.device):
                return fn(ctx, *args)
        return fn(ctx, *args)
    return wrapped


def _check_determinism(what):
    """torch.use_deterministic_algorithms(True) and float atomics do not go together: the library keeps one writer per
    output element under BFVI_DETERMINISTIC=1 (include/bfvi.h); without it follow torch's convention (raise / warn)."""
    if torch.are_deterministic_algorithms_enabled() and os.environ.get('BFVI_DETERMINISTIC', '0') in ('', '0'):
        msg = ('%s accumulates partial sums with float atomics; set BFVI_DETERMINISTIC=1 for run-to-run identical '
               'results under torch.use_deterministic_algorithms(True)' % what)
        if torch.is_deterministic_algorithms_warn_only_enabled():
            import warnings
            warnings.warn(msg)
        else:
            raise RuntimeError(msg)


def _square(v, what):
    a, b = (v, v) if isinstance(v, int) else tuple(v)
    if a != b:
        raise _lib.BfviError('image kernels: %s must be square, got %r' % (what, v))
    return int(a)


def _layer_geometry(layer):
    """(kernel, stride, padding, transposed) of an nn.Conv2d / nn.ConvTranspose2d the kernels serve."""
    transposed = isinstance(layer, nn.ConvTranspose2d)
    if _square(layer.dilation, 'dilation') != 1 or layer.groups != 1 or layer.padding_mode != 'zeros' or \
            isinstance(layer.padding, str) or (transposed and _square(layer.output_padding, 'output_padding') != 0):
        raise _lib.BfviError('image kernels: dilation 1, groups 1, zero padding, output_padding 0 only')
    return (_square(layer.kernel_size, 'kernel_size'), _square(layer.stride, 'stride'),
            _square(layer.padding, 'padding'), transposed)


class _ConvFn(torch.autograd.Function):
    """nn.Conv2d / nn.ConvTranspose2d forward + backward (models/common.py:75-78, 96-99) through
    bfvi_conv_gather / _scatter / _wgrad and bfvi_chan_bias_grad; `sigmoid` fuses the decoder's
    final nn.Sigmoid (models/common.py:148) into the transposed convolution."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, w, b, k, s, p, transposed, sigmoid):
        lib = _library()
        x, w = x.detach().contiguous().float(), w.detach().contiguous().float()
        b = None if b is None else b.detach().contiguous().float()
        g = _lib.ConvGeom()
        g.n, g.kernel, g.stride, g.padding = x.shape[0], k, s, p
        if transposed:
            g.c_small, g.h_small, g.w_small = x.shape[1:]
            g.c_big, g.h_big, g.w_big = w.shape[1], (x.shape[2] - 1) * s - 2 * p + k, (x.shape[3] - 1) * s - 2 * p + k
            y = torch.empty(g.n, g.c_big, g.h_big, g.w_big, device=x.device)
            lib.call('bfvi_conv_scatter', C.byref(g), _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y),
                     _lib.ACT_SIGMOID if sigmoid else _lib.ACT_NONE, _stream(x))
        else:
            g.c_big, g.h_big, g.w_big = x.shape[1:]
            g.c_small, g.h_small, g.w_small = w.shape[0], (x.shape[2] + 2 * p - k) // s + 1, (x.shape[3] + 2 * p - k) // s + 1
            y = torch.empty(g.n, g.c_small, g.h_small, g.w_small, device=x.device)
            lib.call('bfvi_conv_gather', C.byref(g), _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y),
                     _lib.ACT_SIGMOID if sigmoid else _lib.ACT_NONE, _stream(x))
        ctx.geom, ctx.transposed, ctx.sigmoid, ctx.has_bias = g, transposed, sigmoid, b is not None
        ctx.save_for_backward(x, w, y if sigmoid else None)
        return y

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dy):
        _check_determinism('bfvi_conv_wgrad')
        lib = _library()
        x, w, y = ctx.saved_tensors
        g, st = ctx.geom, _stream(x)
        dy = dy.contiguous().float()
        if ctx.sigmoid:
            pre = torch.empty_like(dy)
            lib.call('bfvi_sigmoid_bwd', _lib.ptr(y), _lib.ptr(dy), dy.numel(), _lib.ptr(pre), st)
            dy = pre
        small, big = (x, dy) if ctx.transposed else (dy, x)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            lib.call('bfvi_conv_gather' if ctx.transposed else 'bfvi_conv_scatter', C.byref(g), _lib.ptr(dy), _lib.ptr(w),
                     None, _lib.ptr(dx), _lib.ACT_NONE, st)
        dw = torch.zeros_like(w)
        lib.call('bfvi_conv_wgrad', C.byref(g), _lib.ptr(small), _lib.ptr(big), _lib.ptr(dw), st)
        db = None
        if ctx.has_bias:
            db = torch.zeros(dy.shape[1], device=dy.device)
            sc, n = _scratch(lib, dy.shape[1], dy)
            lib.call('bfvi_chan_bias_grad', _lib.ptr(dy), dy.shape[0], dy.shape[1], dy.shape[2] * dy.shape[3],
                     _lib.ptr(db), _lib.ptr(sc), n, st)
        return dx, dw, db, None, None, None, None, None


class _BatchNormFn(torch.autograd.Function):
    """nn.BatchNorm2d [-> nn.ReLU] (models/common.py:79-84) through bfvi_bn2d_fwd / _bwd; the running
    statistics of `bn` are updated in place by the kernel like torch does."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, gamma, beta, bn, relu):
        lib = _library()
        x = x.detach().contiguous().float()
        N, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
        training = bn.training or bn.running_mean is None
        momentum = 0.0 if bn.momentum is None else float(bn.momentum)
        track = training and bn.track_running_stats and bn.running_mean is not None
        if track:
            bn.num_batches_tracked.add_(1)
            if bn.momentum is None:                              # cumulative moving average
                momentum = 1.0 / float(bn.num_batches_tracked.item())
        gamma = None if gamma is None else gamma.detach().contiguous().float()
        beta = None if beta is None else beta.detach().contiguous().float()
        y = torch.empty_like(x)
        save = torch.empty(Cc, 2, device=x.device)
        sc, n = _scratch(lib, Cc, x)
        lib.call('bfvi_bn2d_fwd', _lib.ptr(x), N, Cc, HW, _lib.ptr(gamma), _lib.ptr(beta),
                 _lib.ptr(bn.running_mean if (track or not training) else None),
                 _lib.ptr(bn.running_var if (track or not training) else None), int(training), momentum, float(bn.eps),
                 int(relu), _lib.ptr(y), _lib.ptr(save), _lib.ptr(sc), n, _stream(x))
        ctx.training, ctx.relu = training, relu
        ctx.save_for_backward(x, y, gamma, save)
        return y

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dy):
        lib = _library()
        x, y, gamma, save = ctx.saved_tensors
        N, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dg = torch.zeros(Cc, device=x.device)
        db = torch.zeros(Cc, device=x.device)
        sc, n = _scratch(lib, Cc, x)
        lib.call('bfvi_bn2d_bwd', _lib.ptr(dy), _lib.ptr(x), _lib.ptr(y), _lib.ptr(save), _lib.ptr(gamma), N, Cc, HW,
                 int(ctx.training), int(ctx.relu), _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(sc), n, _stream(x))
        return dx, (dg if gamma is not None else None), (db if ctx.needs_input_grad[2] else None), None, None


class _DenseFn(torch.autograd.Function):
    """nn.Linear [-> nn.ReLU] (models/common.py:127-133, 146-149: feat_to_z_mean / feat_to_z_std.0 /
    z_to_feat.0) through bfvi_dense_fwd / _bwd: FP32 like the reference (the 4096-long contraction
    on the TF32 tensor cores, even error-compensated, flips ReLU masks downstream)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, w, b, relu):
        _check_determinism('bfvi_dense_fwd')
        lib = _library()
        x, w = x.detach().contiguous().float(), w.detach().contiguous().float()
        b = None if b is None else b.detach().contiguous().float()
        rows, n_in, n_out = x.shape[0], x.shape[1], w.shape[0]
        y = torch.empty(rows, n_out, device=x.device)
        lib.call('bfvi_dense_fwd', _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), rows, n_in, n_out, int(relu),
                 _stream(x))
        ctx.relu, ctx.has_bias = relu, b is not None
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dy):
        lib = _library()
        x, w, y = ctx.saved_tensors
        rows, n_in, n_out = x.shape[0], x.shape[1], w.shape[0]
        dy = dy.contiguous().float()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w)
        db = torch.zeros(n_out, device=x.device) if ctx.has_bias else None
        masked = torch.empty_like(dy) if ctx.relu else None
        sc, n = _scratch(lib, n_out, x)
        lib.call('bfvi_dense_bwd', _lib.ptr(x), _lib.ptr(w), _lib.ptr(y), _lib.ptr(dy), _lib.ptr(masked), rows, n_in, n_out,
                 int(ctx.relu), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(sc), n, _stream(x))
        return dx, dw, db, None


def _dense(layer, x, relu=False):
    if _use_kernels(x, 'dense'):
        return _DenseFn.apply(x, layer.weight, layer.bias, relu)
    y = layer(x)
    return torch.relu(y) if relu else y


def _conv_block(module, layer, n_out, last):
    """layer [-> BatchNorm2d -> ReLU]; the bare layer when it is the last of a stack."""
    module.net = layer if last else nn.Sequential(layer, nn.BatchNorm2d(n_out), nn.ReLU())
    nn.init.xavier_uniform_(layer.weight)


def _run_block(module, layer, x, sigmoid=False):
    """common.Conv / common.Deconv forward: the layer, then BatchNorm2d -> ReLU unless it is the last one."""
    if not _use_kernels(x, 'conv'):
        y = module.net(x)
        return torch.sigmoid(y) if sigmoid else y
    k, s, p, transposed = _layer_geometry(layer)
    y = _ConvFn.apply(x, layer.weight, layer.bias, k, s, p, transposed, sigmoid)
    if module.net is not layer:
        bn = module.net[1]
        y = _BatchNormFn.apply(y, bn.weight, bn.bias, bn, True)
    return y


class Conv(nn.Module):
    """Strided 3x3 convolution block (keys: conv.*, net.0.* / net.*, net.1.*)."""

    def __init__(self, n_channels, n_kernels, kernel_size=3, stride=2, padding=1, last=False):
        super().__init__()
        self.conv = nn.Conv2d(n_channels, n_kernels, kernel_size, stride, padding)
        _conv_block(self, self.conv, n_kernels, last)

    def forward(self, x):
        return _run_block(self, self.conv, x)


class Deconv(nn.Module):
    """Strided 4x4 transposed convolution block (keys: deconv.*, net.*)."""

    def __init__(self, n_channels, n_kernels, kernel_size=4, stride=2, padding=1, last=False):
        super().__init__()
        self.deconv = nn.ConvTranspose2d(n_channels, n_kernels, kernel_size, stride, padding)
        _conv_block(self, self.deconv, n_kernels, last)

    def forward(self, x, sigmoid=False):
        return _run_block(self, self.deconv, x, sigmoid)


def _stack_widths(n_kernels, n_layers):
    """Channel widths of the hidden feature maps, narrowest first: n_kernels / 2^(L-1) .. n_kernels."""
    return [n_kernels // 2 ** (n_layers - 1 - i) for i in range(n_layers)]


class ImageEncoder(nn.Module):
    """img -> 3 stride-2 conv blocks -> (z mean, softplus z std); `gauss_out=False` returns the
    feature maps (keys: conv_stack.<i>.*, feat_to_z_mean.*, feat_to_z_std.0.*)."""

    def __init__(self, z_dim, gauss_out=True, img_size=64, n_channels=3, n_kernels=64, n_layers=3):
        super().__init__()
        self.feat_size = img_size // 2 ** n_layers
        self.feat_dim = self.feat_size ** 2 * n_kernels
        widths = [n_channels] + _stack_widths(n_kernels, n_layers)
        self.conv_stack = nn.Sequential(*[Conv(widths[i], widths[i + 1], last=(i == n_layers - 1))
                                          for i in range(n_layers)])
        self.gauss_out = gauss_out
        if gauss_out:
            self.feat_to_z_mean = nn.Linear(self.feat_dim, z_dim)
            self.feat_to_z_std = nn.Sequential(nn.Linear(self.feat_dim, z_dim), nn.Softplus())
            nn.init.xavier_uniform_(self.feat_to_z_mean.weight)
            nn.init.xavier_uniform_(self.feat_to_z_std[0].weight)

    def forward(self, x):
        feats = self.conv_stack(x)
        if not self.gauss_out:
            return feats
        flat = feats.reshape(-1, self.feat_dim)
        return _dense(self.feat_to_z_mean, flat), F.softplus(_dense(self.feat_to_z_std[0], flat))


class ImageDecoder(nn.Module):
    """z -> Linear + ReLU -> 3 stride-2 transposed conv blocks -> sigmoid pixel probabilities
    (keys: z_to_feat.0.*, deconv_stack.<i>.*)."""

    def __init__(self, z_dim, img_size=64, n_channels=3, n_kernels=64, n_layers=3):
        super().__init__()
        self.feat_size = img_size // 2 ** n_layers
        self.feat_dim = self.feat_size ** 2 * n_kernels
        self.feat_shape = (n_kernels, self.feat_size, self.feat_size)
        self.z_to_feat = nn.Sequential(nn.Linear(z_dim, self.feat_dim), nn.ReLU())
        widths = _stack_widths(n_kernels, n_layers)[::-1] + [n_channels]
        self.deconv_stack = nn.Sequential(*([Deconv(widths[i], widths[i + 1], last=(i == n_layers - 1))
                                             for i in range(n_layers)] + [nn.Sigmoid()]))
        nn.init.xavier_uniform_(self.z_to_feat[0].weight)

    def forward(self, z):
        x = _dense(self.z_to_feat[0], z, relu=True).reshape(-1, *self.feat_shape)
        blocks = list(self.deconv_stack)[:-1]                    # the trailing nn.Sigmoid is fused into the last block
        for i, blk in enumerate(blocks):
            x = blk(x, sigmoid=(i == len(blocks) - 1))
        return (x,)
