"""MultiDMM with the reference's constructor / forward / sample / step API
(models/dmm.py:28-554), computed by the sm_100a kernels behind include/bfvi.h.

Host-side design
  * the sub-module tree (and therefore state_dict keys and seeded initial values)
    is the reference's; after construction every default parameter is re-pointed
    at a slice of ONE flat fp32 CUDA buffer whose layout the C library defines, so
    kernels take a single pointer and data-parallel training all-reduces a single
    flat gradient tensor;
  * `step` on an all-default Gaussian model is one C call (bfvi_step_fwd_bwd) that
    computes the loss AND every parameter gradient in a fixed sequence of kernel
    launches; autograd only scales the stored gradient in `.backward()`;
  * models with custom (e.g. convolutional) encoders/decoders or non-Gaussian
    modalities compose the differentiable ops `encode -> z_filter -> decode`
    (each an autograd.Function over bfvi_*_fwd / bfvi_*_bwd) with their torch
    modules, exactly like the reference's forward();
  * no CPU fallback: compute entry points raise unless the model lives on CUDA.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from . import common, losses
from .dgts import MultiDGTS


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def args_need_grad(keep):
    """`keep` carries the caller's grad mode as its first element."""
    return keep[0]


def _aligned_empty(nbytes, device, align=256):
    buf = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]


class _StepFn(torch.autograd.Function):
    """loss = bfvi_step_fwd_bwd(...); the gradient of every flat-buffer parameter is
    computed in the same call and handed to autograd scaled by grad_output."""

    @staticmethod
    def forward(ctx, model, args, keep, *params):
        lib = _lib.load()
        # grad mode is always off inside Function.forward: the caller decides
        need_grad = bool(args_need_grad(keep)) and any(p.requires_grad for p in params)
        if (need_grad and getattr(model, 'cuda_graph', False) and model._family == 2
                and not (args.eps_match or args.eps_filt or args.eps_sflt or args.eps_ssmt)):
            out = _graph_step(model, lib, args, keep[-1])
            if out is not None:
                ctx.model, ctx.flat_grad = model, out[1]
                return out[0]
        flat_grad = torch.empty_like(model._flat) if need_grad else None
        loss = torch.empty((), dtype=torch.float32, device=model._flat.device)
        nbytes = C.c_size_t(0)
        lib.call('bfvi_step_workspace', C.byref(model._cmodel), C.byref(args), C.byref(nbytes))
        ws = model._workspace(nbytes.value)
        launches = C.c_int32(0)
        lib.call('bfvi_step_fwd_bwd', C.byref(model._cmodel), _lib.ptr(model._flat),
                 _lib.ptr(flat_grad), C.byref(args), _lib.ptr(ws), C.c_size_t(nbytes.value),
                 _lib.ptr(loss), C.byref(launches), _stream())
        model.last_launches = launches.value
        ctx.model, ctx.flat_grad = model, flat_grad
        return loss

    @staticmethod
    def backward(ctx, g):
        model, flat_grad = ctx.model, ctx.flat_grad
        if flat_grad is None:
            raise _lib.BfviError('step() ran without gradients enabled')
        flat_grad = flat_grad * g
        if model.grad_sync is not None:        # data parallel: one all-reduce of the flat buffer
            model.grad_sync(flat_grad)
        model.last_flat_grad = flat_grad
        return (None, None, None) + model._views(flat_grad)


def _graph_step(model, lib, args, tens):
    """Large-dim family: the ~4 500 launches of one step are captured ONCE per (shape, multipliers)
    in a CUDA graph over static input / output buffers and replayed per step; the Philox seed is read
    from device memory (bfvi_step_args.seed_dev), so every replay draws fresh noise.  Same numbers as the
    eager call (tests/test_gpu_large.py::test_cuda_graph_step_matches_eager)."""
    n_mods = model._cmodel.n_mods
    key = (args.T, args.B, tuple(args.rec_mults[i] for i in range(n_mods)), args.kld_mult, args.uni_loss,
           args.f_mode, args.s_mode, args.f_mult, args.s_mult, args.match_mult, args.train_particles,
           args.match_particles, args.sample, args.sample_init, args.b_offset, model._flat.data_ptr())
    cache = model.__dict__.setdefault('_graphs', {})
    ent = cache.get(key)
    dev = model._flat.device
    tb = args.T * args.B

    if ent is not None:
        cache[key] = cache.pop(key)                 # most recently used last
        model._graph_misses = 0
    if ent is None:
        # The multipliers (kld_mult anneals every batch, trainer.py:228-230) and the shape are baked into the
        # captured launches, so a key that changes every step would capture and keep a graph + static buffers
        # per step.  The cache is a small LRU whose evicted entries free their buffers, and after
        # `graph_max_misses` consecutive misses graph mode switches itself off for this model (eager launches).
        model._graph_misses = getattr(model, '_graph_misses', 0) + 1
        if model._graph_misses > getattr(model, 'graph_max_misses', 3):
            model.cuda_graph = False
            cache.clear()
            return None
        while len(cache) >= getattr(model, 'graph_cache_size', 2):
            old = cache.pop(next(iter(cache)))
            old.clear()
        dims = [int(model._cmodel.dims[i]) for i in range(n_mods)]
        ent = {'inputs': [torch.empty(tb * d, device=dev) for d in dims],
               'targets': [torch.empty(tb * d, device=dev) for d in dims],
               'mask': torch.empty(tb, dtype=torch.uint8, device=dev),
               'seed': torch.zeros(1, dtype=torch.int64, device=dev),
               'loss': torch.empty((), dtype=torch.float32, device=dev),
               'grad': torch.empty_like(model._flat)}
        a2 = _lib.StepArgs()
        C.memmove(C.byref(a2), C.byref(args), C.sizeof(_lib.StepArgs))
        for i in range(n_mods):
            a2.inputs[i] = ent['inputs'][i].data_ptr()
            a2.targets[i] = ent['targets'][i].data_ptr()
        a2.seq_mask = ent['mask'].data_ptr()
        a2.seed, a2.seed_dev = 0, ent['seed'].data_ptr()
        nbytes = C.c_size_t(0)
        lib.call('bfvi_step_workspace', C.byref(model._cmodel), C.byref(a2), C.byref(nbytes))
        ent['ws'] = _aligned_empty(max(nbytes.value, 256), dev)
        ent['args'], ent['launches'] = a2, C.c_int32(0)
        ent['graph'] = None
        cache[key] = ent
    # stage this step's batch and seed into the static buffers (device-to-device, stream-ordered)
    for i in range(n_mods):
        ent['inputs'][i].copy_(tens['inputs'][i].reshape(-1))
        ent['targets'][i].copy_(tens['targets'][i].reshape(-1))
    ent['mask'].copy_(tens['mask'].reshape(-1))
    ent['seed'].fill_(int(args.seed))

    def call(stream):
        lib.call('bfvi_step_fwd_bwd', C.byref(model._cmodel), _lib.ptr(model._flat), _lib.ptr(ent['grad']),
                 C.byref(ent['args']), _lib.ptr(ent['ws']), C.c_size_t(ent['ws'].numel()), _lib.ptr(ent['loss']),
                 C.byref(ent['launches']), stream)
    if ent['graph'] is None:
        call(_stream())                 # eager once: lazy initialisation (side streams, attributes) outside capture
        model.last_launches = ent['launches'].value
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            with torch.cuda.graph(graph, stream=cap):
                call(_stream())
        torch.cuda.current_stream(dev).wait_stream(cap)
        ent['graph'] = graph
    else:
        ent['graph'].replay()
    # the static gradient buffer is overwritten by the next replay: hand out a copy (two step() calls before
    # one backward(), e.g. gradient accumulation, must not alias)
    return ent['loss'].clone(), ent['grad'].clone()


class _OpFn(torch.autograd.Function):
    """Differentiable encode / decode over the flat parameter buffer."""

    @staticmethod
    def forward(ctx, model, kind, mod, x, *params):
        lib = _lib.load()
        cm = model._cmodel
        x = x.detach().contiguous().float()
        rows = x.shape[0]
        if kind == 'enc':
            n_out = model.z_dim
            mean = torch.empty(rows, n_out, device=x.device)
            std = torch.empty_like(mean)
            mask = torch.empty(rows, dtype=torch.uint8, device=x.device)
            lib.call('bfvi_encode_fwd', C.byref(cm), _lib.ptr(model._flat), mod, _lib.ptr(x), rows,
                     _lib.ptr(mean), _lib.ptr(std), _lib.ptr(mask), _stream())
            ctx.mark_non_differentiable(mask)
            out = (mean, std, mask)
        else:
            n_out = int(cm.dims[mod])
            mean = torch.empty(rows, n_out, device=x.device)
            std = torch.empty_like(mean)
            lib.call('bfvi_decode_fwd', C.byref(cm), _lib.ptr(model._flat), mod, _lib.ptr(x), rows,
                     _lib.ptr(mean), _lib.ptr(std), _stream())
            out = (mean, std)
        ctx.model, ctx.kind, ctx.mod = model, kind, mod
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, d_mean, d_std, *unused):
        lib = _lib.load()
        model, (x,) = ctx.model, ctx.saved_tensors
        cm = model._cmodel
        rows = x.shape[0]
        flat_grad = torch.zeros_like(model._flat)
        d_mean, d_std = d_mean.contiguous().float(), d_std.contiguous().float()
        d_x = None
        if ctx.kind == 'enc':
            lib.call('bfvi_encode_bwd', C.byref(cm), _lib.ptr(model._flat), _lib.ptr(flat_grad), ctx.mod,
                     _lib.ptr(x), rows, _lib.ptr(d_mean), _lib.ptr(d_std), _stream())
        else:
            d_x = torch.zeros_like(x)
            lib.call('bfvi_decode_bwd', C.byref(cm), _lib.ptr(model._flat), _lib.ptr(flat_grad), ctx.mod,
                     _lib.ptr(x), rows, _lib.ptr(d_mean), _lib.ptr(d_std), _lib.ptr(d_x), _stream())
        return (None, None, None, d_x) + model._views(flat_grad)


class _MlpFn(torch.autograd.Function):
    """One default encoder / decoder module of the composed path through bfvi_mlp_fwd / _bwd: GaussianMLP,
    CategoricalMLP and the categorical encoder Embedding -> ReLU -> GaussianMLP (models/common.py:9-41,
    models/dmm.py:78-82) at any size — tcgen05 GEMMs + fused elementwise kernels, weights by pointer (the modules
    only own them).  `weights` = (emb | None, w1, b1, wa, ba, wb | None, bb | None)."""

    @staticmethod
    def forward(ctx, cfg, x, *weights):
        lib = _lib.load()
        head, nan_mask, min_std = cfg
        emb, w1, b1, wa, ba, wb, bb = [None if w is None else w.detach().contiguous().float() for w in weights]
        x = x.detach().contiguous().float()
        rows = x.shape[0]
        d = _lib.MlpDesc()
        d.emb = 0 if emb is None else emb.data_ptr()
        d.w1, d.b1, d.wa, d.ba = w1.data_ptr(), b1.data_ptr(), wa.data_ptr(), ba.data_ptr()
        d.wb = 0 if wb is None else wb.data_ptr()
        d.bb = 0 if bb is None else bb.data_ptr()
        d.n_in, d.h_dim, d.n_out = w1.shape[1], w1.shape[0], wa.shape[0]
        d.n_classes = 0 if emb is None else emb.shape[0]
        d.head, d.nan_mask, d.min_std = head, int(nan_mask), float(min_std)
        out_a = torch.empty(rows, d.n_out, device=x.device)
        out_b = torch.empty_like(out_a) if head == _lib.HEAD_GAUSSIAN else None
        mask = torch.empty(rows, dtype=torch.uint8, device=x.device) if nan_mask else None
        nbytes = C.c_size_t(0)
        lib.call('bfvi_mlp_workspace', C.byref(d), rows, 0, C.byref(nbytes))
        ws = _aligned_bytes(nbytes.value, x.device)
        lib.call('bfvi_mlp_fwd', C.byref(d), _lib.ptr(x), rows, _lib.ptr(out_a), _lib.ptr(out_b), _lib.ptr(mask),
                 _lib.ptr(ws), C.c_size_t(nbytes.value), _stream())
        ctx.desc, ctx.keep = d, (emb, w1, b1, wa, ba, wb, bb)
        ctx.save_for_backward(x, out_a, *(() if out_b is None else (out_b,)))
        outs = (out_a,) + (() if out_b is None else (out_b,)) + (() if mask is None else (mask,))
        if mask is not None:
            ctx.mark_non_differentiable(mask)
        ctx.n_out = len(outs)
        return outs

    @staticmethod
    def backward(ctx, *d_outs):
        lib = _lib.load()
        d, (emb, w1, b1, wa, ba, wb, bb) = ctx.desc, ctx.keep
        saved = ctx.saved_tensors
        x, out_a = saved[0], saved[1]
        out_b = saved[2] if len(saved) > 2 else None
        rows = x.shape[0]
        gauss = d.head == _lib.HEAD_GAUSSIAN
        cont = lambda t: None if t is None else t.contiguous().float()
        d_a = cont(d_outs[0])
        d_b = cont(d_outs[1]) if gauss else None
        if d_a is None and not gauss:
            d_a = torch.zeros_like(out_a)
        grads = [None if w is None else torch.zeros_like(w) for w in (emb, w1, b1, wa, ba, wb, bb)]
        g = _lib.MlpGrads()
        for name, t in zip(('emb', 'w1', 'b1', 'wa', 'ba', 'wb', 'bb'), grads):
            setattr(g, name, 0 if t is None else t.data_ptr())
        want_dx = ctx.needs_input_grad[1] and emb is None
        d_x = torch.empty(rows, d.n_in, device=x.device) if want_dx else None
        nbytes = C.c_size_t(0)
        lib.call('bfvi_mlp_workspace', C.byref(d), rows, 1, C.byref(nbytes))
        ws = _aligned_bytes(nbytes.value, x.device)
        lib.call('bfvi_mlp_bwd', C.byref(d), C.byref(g), _lib.ptr(x), rows, _lib.ptr(out_a), _lib.ptr(out_b),
                 _lib.ptr(d_a), _lib.ptr(d_b), _lib.ptr(d_x), _lib.ptr(ws), C.c_size_t(nbytes.value), _stream())
        return (None, d_x) + tuple(grads)


def _aligned_bytes(nbytes, device):
    buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    return buf[(-buf.data_ptr()) % 256:]


class _FilterFn(torch.autograd.Function):
    """MultiDMM.z_filter (models/dmm.py:319-412) on arbitrary expert tensors."""

    @staticmethod
    def forward(ctx, model, cfg, z_mean, z_std, z_masks, eps, *params):
        lib = _lib.load()
        z_mean = z_mean.detach().contiguous().float()
        z_std = z_std.detach().contiguous().float()
        masks = z_masks.to(torch.uint8).contiguous()
        n_exp, t_max, b_dim, z = z_mean.shape
        outs = [torch.empty(t_max, b_dim, z, device=z_mean.device) for _ in range(5)]
        args = model._filter_args(cfg, z_mean, z_std, masks, eps, outs)
        keep_ws = model._filter_workspace(lib, args)
        lib.call('bfvi_filter_fwd', C.byref(model._cmodel), _lib.ptr(model._flat), C.byref(args),
                 _stream())
        del keep_ws
        ctx.model, ctx.cfg, ctx.eps = model, cfg, eps
        ctx.save_for_backward(z_mean, z_std, masks, *outs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *d_outs):
        lib = _lib.load()
        model = ctx.model
        z_mean, z_std, masks = ctx.saved_tensors[:3]
        outs = list(ctx.saved_tensors[3:])
        flat_grad = torch.zeros_like(model._flat)
        d_mean, d_std = torch.zeros_like(z_mean), torch.zeros_like(z_std)
        d_outs = [None if d is None else d.contiguous().float() for d in d_outs]
        args = model._filter_args(ctx.cfg, z_mean, z_std, masks, ctx.eps, outs, d_mean, d_std, d_outs)
        keep_ws = model._filter_workspace(lib, args)
        lib.call('bfvi_filter_bwd', C.byref(model._cmodel), _lib.ptr(model._flat), _lib.ptr(flat_grad),
                 C.byref(args), _stream())
        del keep_ws
        return (None, None, d_mean, d_std, None, None) + model._views(flat_grad)


class MultiDMM(MultiDGTS):

    def __init__(self, modalities, dims, dists=None, encoders=None, decoders=None, h_dim=32,
                 z_dim=32, z0_mean=0.0, z0_std=1.0, min_std=1e-3, device=torch.device('cuda:0')):
        super().__init__()
        self.modalities = list(modalities)
        self.n_mods = len(self.modalities)
        self.dims = dict(zip(self.modalities, dims))      # `dims` may be a generator
        self.h_dim, self.z_dim = h_dim, z_dim
        if dists is None:
            dists = ['Normal'] * self.n_mods
        self.dists = dict(zip(self.modalities, dists))

        # same construction order as the reference (models/dmm.py:75-116) so that a
        # seeded construction yields identical initial weights
        self.enc = nn.ModuleDict()
        for m in self.modalities:
            n_in = int(np.prod(self.dims[m]))
            if self.dists[m] == 'Categorical':
                self.enc[m] = nn.Sequential(nn.Embedding(n_in, h_dim), nn.ReLU(),
                                            common.GaussianMLP(h_dim, z_dim, h_dim))
            else:
                self.enc[m] = common.GaussianMLP(n_in, z_dim, h_dim)
        self._custom_enc = set()
        if encoders is not None:
            if isinstance(encoders, list):
                encoders = dict(zip(self.modalities, encoders))
            self.enc.update(encoders)
            self._custom_enc = set(encoders.keys())
        self.dec = nn.ModuleDict()
        for m in self.modalities:
            n_out = int(np.prod(self.dims[m]))
            if self.dists[m] == 'Categorical':
                self.dec[m] = common.CategoricalMLP(z_dim, n_out, h_dim)
            else:
                self.dec[m] = common.GaussianMLP(z_dim, n_out, h_dim)
        self._custom_dec = set()
        if decoders is not None:
            if isinstance(decoders, list):
                decoders = dict(zip(self.modalities, decoders))
            self.dec.update(decoders)
            self._custom_dec = set(decoders.keys())
        self.trans = nn.ModuleDict()
        self.trans['fwd'] = common.GaussianGTF(z_dim, h_dim, min_std=min_std)
        self.trans['bwd'] = common.GaussianGTF(z_dim, h_dim, min_std=min_std)
        self.z0_mean = nn.Parameter(z0_mean * torch.ones(1, z_dim))
        self.z0_log_std = nn.Parameter((z0_std * torch.ones(1, z_dim)).log())
        self.min_std = min_std

        # no silent CPU fallback (the reference falls back at models/dmm.py:120-121)
        self.device = torch.device(device)
        self._flat = None
        self._slots = []
        self._ws = None
        self.grad_sync = None          # optional callable(flat_grad) -> None (data parallel)
        self.noise_seed = None         # fixed Philox seed (None: drawn from torch's RNG per call)
        self.precision = 'fused'       # large-dim family step(): 'fused' (on-chip GTF kernels) | 'tf32x3' | 'tf32'
        self.batch_tile = 0            # large-dim family step(): sequences per batch tile (0 = library default)
        self.last_launches = 0
        self.last_flat_grad = None
        self.to(self.device)

    # ------------------------------------------------------------------ plumbing
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat = None              # parameters were re-created: re-alias lazily
        self._ws = None
        return out

    def _mods_flat(self):
        return [int(np.prod(self.dims[m])) for m in self.modalities]

    def _default_enc(self, m):
        return m not in self._custom_enc and self.dists[m] != 'Categorical'

    def _default_dec(self, m):
        return m not in self._custom_dec and self.dists[m] == 'Normal'

    def _require_cuda(self):
        p = self.z0_mean
        if not p.is_cuda:
            raise _lib.BfviError('MultiDMM computes on CUDA only; model is on %s '
                                 '(there is no CPU fallback)' % p.device)

    def _ensure_flat(self):
        """Alias every default parameter onto one flat CUDA buffer (C-defined layout)."""
        self._require_cuda()
        if self._flat is not None:
            base = self._flat.data_ptr()
            if all(p.data_ptr() == base + 4 * off for _, off, p in self._slots):
                return
        lib = _lib.load()
        dist_list = [self.dists[m] for m in self.modalities]
        self._cmodel = _lib.make_model(self._mods_flat(), dist_list, self.z_dim, self.h_dim,
                                       self.min_std)
        lay = lib.layout(self._cmodel)
        self._family = lib.dll.bfvi_kernel_family(C.byref(self._cmodel))
        named = dict(self.named_parameters())
        dev = self.z0_mean.device
        flat = torch.zeros(lay.total, dtype=torch.float32, device=dev)
        slots = []
        for key, off in _lib.param_slots(self.modalities, dist_list, lay):
            mod_name = key.split('.')[1] if key[:4] in ('enc.', 'dec.') else None
            if key.startswith('enc.') and not self._default_enc(mod_name):
                continue
            if key.startswith('dec.') and not self._default_dec(mod_name):
                continue
            p = named[key]
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            slots.append((key, off, p))
        self._flat, self._slots = flat, slots

    def _slot_params(self):
        return [p for _, _, p in self._slots]

    def _views(self, flat_grad):
        return tuple(flat_grad[off:off + p.numel()].view(p.shape) for _, off, p in self._slots)

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = _aligned_empty(nbytes, self._flat.device)
        return self._ws

    def _next_seed(self):
        if self.noise_seed is not None:
            return int(self.noise_seed)
        if getattr(self, 'seed_source', None) is not None:        # data parallel: same seed on every rank
            return int(self.seed_source())
        return int(torch.randint(0, 2 ** 62, (1,)).item())       # host RNG, no device sync

    @property
    def fused_step_available(self):
        self._ensure_flat()
        return (self._family in (1, 2) and all(self._default_enc(m) and self._default_dec(m)
                                               for m in self.modalities))

    # ------------------------------------------------------------- reference API
    def prior(self, shape):
        """Global prior parameters broadcast to `shape` (models/dmm.py:124-129)."""
        mean = self.z0_mean.repeat(*shape)
        std = (self.z0_log_std.exp() + self.min_std).repeat(*shape)
        mask = torch.ones(shape[:-1], dtype=torch.bool, device=mean.device)
        return mean, std, mask

    def encode(self, inputs, combine=False):
        """Per-modality q'(z|x_m) with NaN -> mask (models/dmm.py:131-190)."""
        self._ensure_flat()
        first = inputs[list(inputs.keys())[0]]
        t_max, b_dim = first.shape[:2]
        means, stds, masks = [], [], []
        for i, m in enumerate(self.modalities):
            if m not in inputs:
                continue
            x = inputs[m]
            if self._default_enc(m) and self._family == 1:
                mu, sd, mk = _OpFn.apply(self, 'enc', i, x.reshape(t_max * b_dim, -1),
                                         *self._slot_params())
                mk = mk.bool().reshape(t_max, b_dim)
            elif m not in self._custom_enc:
                # default modules outside the register-resident family: tcgen05 GEMMs + fused elementwise kernels
                # (bfvi_mlp_fwd / _bwd), the modules only own the weights
                mu, sd, mk = self._mlp_encode(m, x.reshape(t_max * b_dim, -1))
                mk = mk.bool().reshape(t_max, b_dim)
            else:
                mk = ~torch.isnan(x).flatten(2, -1).any(dim=-1)
                xz = torch.nan_to_num(x.detach(), nan=0.0)
                mu, sd = self.enc[m](xz.flatten(0, 1))
            means.append(mu.reshape(t_max, b_dim, -1))
            stds.append(sd.reshape(t_max, b_dim, -1))
            masks.append(mk)
        z_mean, z_std, z_masks = torch.stack(means), torch.stack(stds), torch.stack(masks)
        if combine:
            z_mean, z_std = self.product_of_experts(z_mean, z_std, z_masks)
            z_masks = z_masks.any(dim=0)
        return z_mean, z_std, z_masks

    def decode(self, z):
        """Every decoder on (T, B, Z) latents (models/dmm.py:192-212)."""
        self._ensure_flat()
        t_max, b_dim = z.shape[:2]
        recon = {}
        for i, m in enumerate(self.modalities):
            flat = z.reshape(-1, self.z_dim)
            if self._default_dec(m) and self._family == 1:
                out = _OpFn.apply(self, 'dec', i, flat, *self._slot_params())
            elif m not in self._custom_dec:            # GaussianMLP / CategoricalMLP through bfvi_mlp_fwd / _bwd
                out = self._mlp_decode(m, flat)
            else:
                out = self.dec[m](flat)
            recon[m] = tuple(r.reshape(t_max, b_dim, *r.shape[1:]) for r in out)
        return recon

    def _mlp_encode(self, m, x):
        """Default encoder of modality m on (rows, D) inputs -> (mean, std, mask) through bfvi_mlp_fwd."""
        enc = self.enc[m]
        if self.dists[m] == 'Categorical':            # Embedding -> ReLU -> GaussianMLP (models/dmm.py:78-82)
            emb, mlp = enc[0].weight, enc[2]
            x = x.reshape(-1)
        else:
            emb, mlp = None, enc
        cfg = (_lib.HEAD_GAUSSIAN, True, mlp.min_std)
        return _MlpFn.apply(cfg, x, emb, mlp.in_to_h[0].weight, mlp.in_to_h[0].bias, mlp.h_to_mean.weight,
                            mlp.h_to_mean.bias, mlp.h_to_std[0].weight, mlp.h_to_std[0].bias)

    def _mlp_decode(self, m, z):
        """Default decoder of modality m on (rows, Z) latents through bfvi_mlp_fwd: (mean, std) or (probs,)."""
        dec = self.dec[m]
        if self.dists[m] == 'Categorical':            # CategoricalMLP (models/common.py:9-23)
            cfg = (_lib.HEAD_SOFTMAX, False, 0.0)
            return _MlpFn.apply(cfg, z, None, dec.in_to_h[0].weight, dec.in_to_h[0].bias, dec.h_to_out[0].weight,
                                dec.h_to_out[0].bias, None, None)
        cfg = (_lib.HEAD_GAUSSIAN, False, dec.min_std)
        return _MlpFn.apply(cfg, z, None, dec.in_to_h[0].weight, dec.in_to_h[0].bias, dec.h_to_mean.weight,
                            dec.h_to_mean.bias, dec.h_to_std[0].weight, dec.h_to_std[0].bias)

    def z_next(self, z, direction='fwd', glb_prior=None):
        """p(z_next | particles z) (models/dmm.py:214-258); tensor-level helper."""
        if glb_prior is None:
            glb_mean, glb_std, _ = self.prior(z.shape[1:])
        else:
            glb_mean, glb_std = glb_prior
        k = z.shape[0]
        q_mean, q_std = self.trans[direction](z.reshape(-1, self.z_dim))
        mean, std = self.product_of_experts(torch.stack([glb_mean.repeat(k, 1), q_mean]),
                                            torch.stack([glb_std.repeat(k, 1), q_std]))
        if k == 1:
            return mean, std
        return self.mean_of_experts(mean.view(*z.shape), std.view(*z.shape))

    def z_sample(self, t_max, b_dim, direction='fwd', sample=True, n_particles=1, z_init=None,
                 inclusive=False):
        """Ancestral sampling of the latent chain (models/dmm.py:260-317)."""
        glb_mean, glb_std, _ = self.prior((b_dim, 1))
        mean_t, std_t = (glb_mean, glb_std) if z_init is None else z_init
        means, stds = [], []
        if inclusive:
            means.append(mean_t)
            stds.append(std_t)
        for _ in range(t_max - int(inclusive)):
            if sample or n_particles > 1:
                z_t = self._sample_gauss(mean_t.expand(n_particles, -1, -1),
                                         std_t.expand(n_particles, -1, -1))
            else:
                z_t = mean_t.unsqueeze(0)
            mean_t, std_t = self.z_next(z_t, direction, (glb_mean, glb_std))
            means.append(mean_t)
            stds.append(std_t)
        if direction == 'bwd':
            means.reverse()
            stds.reverse()
        return torch.stack(means), torch.stack(stds)

    def _filter_args(self, cfg, z_mean, z_std, masks, eps, outs, d_mean=None, d_std=None,
                     d_outs=None):
        n_exp, t_max, b_dim, z = z_mean.shape
        if n_exp > _lib.MAX_EXPERTS:
            raise _lib.BfviError('too many experts')
        a = _lib.FilterArgs()
        a.T, a.B, a.S, a.n_experts = t_max, b_dim, 1, n_exp
        tbz, tb = t_max * b_dim * z, t_max * b_dim
        for e in range(n_exp):
            ex = a.experts[e]
            ex.mean = z_mean.data_ptr() + 4 * e * tbz
            ex.std = z_std.data_ptr() + 4 * e * tbz
            ex.mask = masks.data_ptr() + e * tb
            ex.stride_s, ex.stride_t, ex.stride_b = 0, b_dim * z, z
            ex.mstride_s, ex.mstride_t, ex.mstride_b = 0, b_dim, 1
            if d_mean is not None:
                ex.d_mean = d_mean.data_ptr() + 4 * e * tbz
                ex.d_std = d_std.data_ptr() + 4 * e * tbz
            ex.kind = _lib.EXPERT_TENSOR
        a.set_expert_bits[0] = (1 << n_exp) - 1
        a.direction = _lib.DIR_BWD if cfg['direction'] == 'bwd' else _lib.DIR_FWD
        a.n_particles = int(cfg['n_particles'])
        a.sample, a.sample_init = int(bool(cfg['sample'])), int(bool(cfg['sample_init']))
        a.noise.eps = None if eps is None else eps.data_ptr()
        a.noise.seed, a.noise.stream_id, a.noise.b_offset = cfg['seed'], 7, 0
        (a.infer_mean, a.infer_std, a.prior_mean, a.prior_std, a.samples) = [t.data_ptr() for t in outs]
        if d_outs is not None:
            names = ('d_infer_mean', 'd_infer_std', 'd_prior_mean', 'd_prior_std', 'd_samples')
            self._keep = d_outs
            for n, d in zip(names, d_outs):
                setattr(a, n, None if d is None else d.data_ptr())
        return a

    def _filter_workspace(self, lib, args):
        """Scratch of the large-dim family for one z_filter call (0 bytes for the small-dim family)."""
        nbytes = C.c_size_t(0)
        lib.call('bfvi_filter_workspace', C.byref(self._cmodel), C.byref(args), C.byref(nbytes))
        if nbytes.value == 0:
            return None
        ws = self._workspace(nbytes.value)
        args.workspace, args.workspace_bytes = ws.data_ptr(), nbytes.value
        return ws

    def z_filter(self, z_mean, z_std, z_masks, direction='fwd', sample=True, n_particles=1,
                 sample_init=False, eps=None):
        """Product-of-experts filtering along time (models/dmm.py:319-412).
        z_mean/z_std (E, T, B, Z), z_masks (E, T, B).  `eps` optionally injects the
        N(0,1) draws as a (T, B, K, Z) tensor; default is the in-kernel generator."""
        self._ensure_flat()
        cfg = dict(direction=direction, sample=sample, n_particles=n_particles,
                   sample_init=sample_init, seed=self._next_seed())
        if eps is not None:
            eps = eps.detach().contiguous().float()
        im, isd, pm, ps, smp = _FilterFn.apply(self, cfg, z_mean, z_std, z_masks, eps,
                                               *self._slot_params())
        return (im, isd), (pm, ps), smp

    def sample(self, t_max, b_dim, direction='fwd'):
        """Unconditional generation (models/dmm.py:414-418)."""
        self._ensure_flat()
        z_mean, _ = self.z_sample(t_max, b_dim, direction, sample=True)
        return self.decode(z_mean)

    def forward(self, inputs, **kwargs):
        """Reconstruct (optionally missing) inputs; returns (infer, prior, recon)
        exactly like models/dmm.py:420-494.  Extra keyword `noise=(eps_flt, eps_smt)`
        injects the reparameterisation draws of the two passes."""
        self._ensure_flat()
        lengths = kwargs.get('lengths')
        mode = kwargs.get('mode', 'fsmooth')
        sample = kwargs.get('sample', True)
        sample_init = kwargs.get('sample_init', False)
        flt_particles = kwargs.get('flt_particles', 1)
        smt_particles = kwargs.get('smt_particles', 1)
        eps_flt, eps_smt = kwargs.get('noise', (None, None))
        t_max, b_dim = max(lengths), len(lengths)
        all_default = all(self._default_enc(m) and self._default_dec(m) for m in self.modalities)
        needs_graph = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self._family == 2 and all_default and not needs_graph:
            # inference fast path: ONE C call, every Linear layer a tcgen05 GEMM
            return self._forward_large(inputs, t_max, b_dim, mode, sample, sample_init, flt_particles,
                                       smt_particles, eps_flt, eps_smt, kwargs.get('precision', self.precision))
        # composed, differentiable path: encode -> z_filter (fused temporal core) -> decode

        obs_mean, obs_std, obs_mask = self.encode(inputs)
        direction = 'fwd' if mode in ('ffilter', 'bsmooth') else 'bwd'
        flt_init = sample_init if mode in ('ffilter', 'bfilter') else False
        infer, prior, z_samples = self.z_filter(obs_mean, obs_std, obs_mask, direction=direction,
                                                sample=sample, n_particles=flt_particles,
                                                sample_init=flt_init, eps=eps_flt)
        if mode in ('fsmooth', 'bsmooth'):
            direction = 'fwd' if mode == 'fsmooth' else 'bwd'
            inv_mean, inv_std, inv_mask = self.prior((t_max, b_dim, 1))
            flt_mask = torch.ones((t_max, b_dim), dtype=torch.bool, device=obs_mask.device)
            flt_mask[-1] = False                                   # models/dmm.py:482
            infer, prior, z_samples = self.z_filter(
                torch.cat([obs_mean, prior[0][None], inv_mean[None]], 0),
                torch.cat([obs_std, prior[1][None], -inv_std[None]], 0),
                torch.cat([obs_mask.bool(), flt_mask[None], inv_mask[None]], 0),
                direction=direction, sample=sample, n_particles=smt_particles,
                sample_init=sample_init, eps=eps_smt)
        recon = self.decode(z_samples)
        return infer, prior, recon

    def _forward_large(self, inputs, t_max, b_dim, mode, sample, sample_init, flt_particles,
                       smt_particles, eps_flt, eps_smt, precision):
        """Inference forward() of the large-dim kernel family for all-default Gaussian models: ONE C
        call (bfvi_forward) that runs every Linear layer as a tcgen05 GEMM.  No autograd graph:
        training goes through step(), and forward() under grad mode takes the composed path."""
        lib = _lib.load()
        dev = self._flat.device
        a = _lib.ForwardArgs()
        keep = []
        a.T, a.B = t_max, b_dim
        for i, m in enumerate(self.modalities):
            if m in inputs:
                x = inputs[m]
                if not x.is_cuda:
                    raise _lib.BfviError('forward() inputs must be CUDA tensors')
                x = x.detach().reshape(t_max, b_dim, -1).contiguous().float()
                keep.append(x)
                a.inputs[i] = x.data_ptr()
        a.mode = _lib.MODE_CODES[mode]
        a.sample, a.sample_init = int(bool(sample)), int(bool(sample_init))
        a.flt_particles, a.smt_particles = int(flt_particles), int(smt_particles)
        for t, field in ((eps_flt, 'eps_flt'), (eps_smt, 'eps_smt')):
            if t is not None:
                t = t.detach().to(dev).contiguous().float()
                keep.append(t)
                setattr(a, field, t.data_ptr())
        a.seed, a.b_offset = self._next_seed(), int(getattr(self, 'b_offset', 0))
        a.precision = _lib.PRECISION_CODES[precision]
        outs = [torch.empty(t_max, b_dim, self.z_dim, device=dev) for _ in range(4)]
        a.infer_mean, a.infer_std, a.prior_mean, a.prior_std = [t.data_ptr() for t in outs]
        recon = {}
        for i, m in enumerate(self.modalities):
            d = self._mods_flat()[i]
            mean, std = torch.empty(t_max, b_dim, d, device=dev), torch.empty(t_max, b_dim, d, device=dev)
            a.recon_mean[i], a.recon_std[i] = mean.data_ptr(), std.data_ptr()
            shape = (t_max, b_dim) + tuple(np.atleast_1d(self.dims[m]).tolist())
            recon[m] = (mean.view(shape), std.view(shape))
        nbytes = C.c_size_t(0)
        lib.call('bfvi_forward_workspace', C.byref(self._cmodel), C.byref(a), C.byref(nbytes))
        ws = self._workspace(nbytes.value)
        lib.call('bfvi_forward', C.byref(self._cmodel), _lib.ptr(self._flat), C.byref(a), _lib.ptr(ws),
                 C.c_size_t(nbytes.value), _stream())
        return (outs[0], outs[1]), (outs[2], outs[3]), recon

    def kld_prior(self, n_particles, direction='fwd'):
        """KL(p(z) || E_k p(z_next | z_k)) (models/dmm.py:496-501)."""
        glb_mean, glb_std, _ = self.prior((1, 1, 1))
        nxt_mean, nxt_std = self.z_sample(1, 1, direction, True, n_particles)
        return losses.kld_gauss(glb_mean, glb_std, nxt_mean, nxt_std)

    def step(self, inputs, mask, kld_mult, rec_mults, targets=None, uni_loss=True, **kwargs):
        """Bidirectional training loss (models/dmm.py:503-554): prior matching +
        f_mult * ELBO(f_mode) + s_mult * ELBO(s_mode with train_particles).
        Returns the un-normalised summed loss as a 0-dim tensor; call
        `(loss / sum(lengths)).backward()` as trainer.py:242-243 does.
        Extra keyword `noise=` (dict with keys match/filt/sflt/ssmt in the layouts of
        include/bfvi.h) injects the reparameterisation draws."""
        self._ensure_flat()
        fused_kw = {'f_mode', 's_mode', 'f_mult', 's_mult', 'match_mult', 'train_particles',
                    'match_particles', 'lengths', 'sample', 'sample_init', 'noise', 'precision', 'batch_tile'}
        # the reference forwards ANY f_mode / s_mode to forward() (models/dmm.py:547-553): the fused C call
        # covers the default pairing (filtering ELBO + smoothing ELBO), everything else composes the ops
        modes_ok = (kwargs.get('f_mode', 'bfilter') in ('bfilter', 'ffilter') and
                    kwargs.get('s_mode', 'fsmooth') in ('fsmooth', 'bsmooth'))
        if self.fused_step_available and set(kwargs) <= fused_kw and modes_ok and \
                all(m in inputs for m in self.modalities):
            return self._fused_step(inputs, mask, kld_mult, rec_mults, targets, uni_loss, kwargs)
        # composed path (custom modules / non-Gaussian modalities / unusual kwargs)
        kwargs = dict(kwargs)
        kwargs.pop('noise', None)
        kwargs.pop('precision', None)
        kwargs.pop('batch_tile', None)
        f_mode, s_mode = kwargs.pop('f_mode', 'bfilter'), kwargs.pop('s_mode', 'fsmooth')
        f_mult, s_mult = kwargs.pop('f_mult', 0.5), kwargs.pop('s_mult', 0.5)
        match_mult = kwargs.pop('match_mult', 0.01)
        train_particles = kwargs.pop('train_particles', 25)
        match_particles = kwargs.pop('match_particles', 50)
        total = 0
        if match_mult > 0:
            n_obs = mask.sum().float()
            total = total + match_mult * kld_mult * n_obs * self.kld_prior(match_particles, 'fwd')
            total = total + match_mult * kld_mult * n_obs * self.kld_prior(match_particles, 'bwd')
        total = total + f_mult * super().step(inputs, mask, kld_mult, rec_mults, targets, uni_loss,
                                              mode=f_mode, **kwargs)
        total = total + s_mult * super().step(inputs, mask, kld_mult, rec_mults, targets, uni_loss,
                                              mode=s_mode, flt_particles=train_particles, **kwargs)
        return total

    def _fused_step(self, inputs, mask, kld_mult, rec_mults, targets, uni_loss, kw):
        if targets is None:
            targets = inputs
        first = inputs[self.modalities[0]]
        t_max, b_dim = first.shape[:2]
        dev = self._flat.device
        a = _lib.StepArgs()
        keep = [torch.is_grad_enabled()]

        def dev_f32(t):
            if not t.is_cuda:
                raise _lib.BfviError('step() inputs must be CUDA tensors')
            t = t.detach().reshape(t_max, b_dim, -1).contiguous().float()
            keep.append(t)
            return t.data_ptr()
        a.T, a.B = t_max, b_dim
        tens = {'inputs': [], 'targets': []}       # the tensors behind the pointers (CUDA-graph staging)
        for i, m in enumerate(self.modalities):
            a.inputs[i] = dev_f32(inputs[m])
            tens['inputs'].append(keep[-1])
            a.targets[i] = dev_f32(targets[m]) if m in targets else a.inputs[i]
            tens['targets'].append(keep[-1])
            mult = float(rec_mults.get(m, 1.0)) if m in targets else 0.0
            a.rec_mults[i] = mult
        mk = mask.reshape(t_max, b_dim).to(device=dev, dtype=torch.uint8).contiguous()
        keep.append(mk)
        tens['mask'] = mk
        a.seq_mask = mk.data_ptr()
        a.kld_mult, a.uni_loss = float(kld_mult), int(bool(uni_loss))
        a.f_mode = _lib.MODE_CODES[kw.get('f_mode', 'bfilter')]
        a.s_mode = _lib.MODE_CODES[kw.get('s_mode', 'fsmooth')]
        a.f_mult, a.s_mult = float(kw.get('f_mult', 0.5)), float(kw.get('s_mult', 0.5))
        a.match_mult = float(kw.get('match_mult', 0.01))
        a.train_particles = int(kw.get('train_particles', 25))
        a.match_particles = int(kw.get('match_particles', 50))
        a.sample, a.sample_init = int(bool(kw.get('sample', True))), int(bool(kw.get('sample_init', False)))
        a.seed, a.b_offset, a.match_count = self._next_seed(), int(getattr(self, 'b_offset', 0)), -1.0
        # large-dim family: 'fused' (on-chip transition kernels, FP32-class forward) / 'tf32x3' / 'tf32'
        # (launch-sequence path) and the batch tile the step walks (0 = chosen by the library)
        a.precision = _lib.PRECISION_CODES[kw.get('precision', self.precision)]
        a.batch_tile = int(kw.get('batch_tile', self.batch_tile))
        noise = kw.get('noise')
        if noise is not None:
            for name, field in (('match', 'eps_match'), ('filt', 'eps_filt'), ('sflt', 'eps_sflt'),
                                ('ssmt', 'eps_ssmt')):
                t = noise[name].to(dev).contiguous().float()
                keep.append(t)
                setattr(a, field, t.data_ptr())
        keep.append(tens)
        return _StepFn.apply(self, a, keep, *self._slot_params())
