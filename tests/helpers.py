"""Shared test plumbing: call the C ABI directly on torch tensors.

The same code drives (a) the emulated kernels on CPU tensors (tests/emu, a
development check of kernel logic) and (b) the real CUDA library on cuda tensors
(`-m gpu`, the parity tests proper)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))

from multimodal_dmm_b200 import _lib  # noqa: E402


def emu_library():
    import build_emu
    return _lib.Library(build_emu.build())


def aligned_empty(nbytes, device, align=256):
    buf = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]


def fixture_model(fx):
    dists = ['Normal'] * len(fx['modalities'])
    return _lib.make_model(fx['dims'], dists, fx['z_dim'], fx['h_dim'], fx['min_std']), dists


def pack_params(lib, model, modalities, dists, state_dict, device):
    lay = lib.layout(model)
    flat = torch.zeros(lay.total, dtype=torch.float32)
    for key, off in _lib.param_slots(modalities, dists, lay):
        v = state_dict[key].reshape(-1).float()
        flat[off:off + v.numel()] = v
    return flat.to(device), lay


def unpack(lib, model, modalities, dists, flat, like):
    lay = lib.layout(model)
    out = {}
    flat = flat.cpu()
    for key, off in _lib.param_slots(modalities, dists, lay):
        n = like[key].numel()
        out[key] = flat[off:off + n].reshape(like[key].shape).clone()
    return out


def step_args(fx, device, noise=None, seed=0, kwargs=None, b_offset=0):
    """Builds bfvi_step_args for a golden fixture; returns (args, keepalive)."""
    kw = dict(fx['step_kwargs'])
    kw.update(kwargs or {})
    mods = fx['modalities']
    t_max, b_dim = fx['mask'].shape[:2]
    a = _lib.StepArgs()
    keep = []
    a.T, a.B = t_max, b_dim
    for i, m in enumerate(mods):
        x = fx['inputs'][m].to(device).contiguous()
        y = fx['targets'][m].to(device).contiguous()
        keep += [x, y]
        a.inputs[i], a.targets[i] = x.data_ptr(), y.data_ptr()
        a.rec_mults[i] = float(fx['rec_mults'].get(m, 1.0))
    mask = fx['mask'].reshape(t_max, b_dim).to(torch.uint8).to(device).contiguous()
    keep.append(mask)
    a.seq_mask = mask.data_ptr()
    a.kld_mult = float(fx['kld_mult'])
    a.uni_loss = int(kw.get('uni_loss', True))
    a.f_mode = _lib.MODE_CODES[kw.get('f_mode', 'bfilter')]
    a.s_mode = _lib.MODE_CODES[kw.get('s_mode', 'fsmooth')]
    a.f_mult, a.s_mult = float(kw.get('f_mult', 0.5)), float(kw.get('s_mult', 0.5))
    a.match_mult = float(kw.get('match_mult', 0.01))
    a.train_particles = int(kw.get('train_particles', 25))
    a.match_particles = int(kw.get('match_particles', 50))
    a.sample, a.sample_init = int(kw.get('sample', True)), int(kw.get('sample_init', False))
    a.seed, a.b_offset, a.match_count = seed, b_offset, -1.0
    a.precision, a.batch_tile = int(kw.get('precision', 0)), int(kw.get('batch_tile', 0))   # large-dim family knobs
    if noise is not None:
        for name, field in (('match', 'eps_match'), ('filt', 'eps_filt'), ('sflt', 'eps_sflt'),
                            ('ssmt', 'eps_ssmt')):
            t = noise[name].to(device).contiguous()
            keep.append(t)
            setattr(a, field, t.data_ptr())
    return a, keep


def run_step(lib, fx, device, with_grad=True, noise='fixture', seed=0, kwargs=None, b_offset=0,
             return_flat=False):
    """Calls bfvi_step_fwd_bwd; returns (loss float, {param: grad of the summed loss})."""
    model, dists = fixture_model(fx)
    mods = fx['modalities']
    flat, lay = pack_params(lib, model, mods, dists, fx['state_dict'], device)
    nz = fx['noise'] if noise == 'fixture' else noise
    a, keep = step_args(fx, device, nz, seed, kwargs, b_offset)
    nbytes = C.c_size_t(0)
    lib.call('bfvi_step_workspace', C.byref(model), C.byref(a), C.byref(nbytes))
    ws = aligned_empty(nbytes.value, device)
    grads = torch.full((lay.total,), float('nan'), dtype=torch.float32, device=device) if with_grad else None
    loss = torch.zeros(1, dtype=torch.float32, device=device)
    launches = C.c_int32(0)
    stream = None
    if device != 'cpu':
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.call('bfvi_step_fwd_bwd', C.byref(model), _lib.ptr(flat), _lib.ptr(grads), C.byref(a),
             _lib.ptr(ws), C.c_size_t(nbytes.value), _lib.ptr(loss), C.byref(launches), stream)
    if device != 'cpu':
        torch.cuda.synchronize()
    if return_flat:
        return loss.item(), grads, launches.value
    g = unpack(lib, model, mods, dists, grads, fx['state_dict']) if with_grad else None
    return loss.item(), g, launches.value


# ---------------------------------------------------------------------------------------
# large-dim family: bfvi_forward through the C ABI + the oracle on the same injected noise
# ---------------------------------------------------------------------------------------
def large_case(z_dim, h_dim, dims, t_max, lengths, seed, nan_frac=0.3, drop=()):
    """Random model + NaN-padded / NaN-deleted inputs for a (z_dim, h_dim) outside the
    small-dim family.  Returns a fixture-like dict."""
    import bfvi_oracle as bo
    mods = ['m%d' % i for i in range(len(dims))]
    g = torch.Generator().manual_seed(seed)
    b_dim = len(lengths)
    inputs = {}
    for m, d in zip(mods, dims):
        if m in drop:
            continue
        x = torch.randn(t_max, b_dim, d, generator=g)
        miss = torch.rand(t_max, b_dim, generator=g) < nan_frac
        x[miss] = float('nan')
        for b, n in enumerate(lengths):
            x[n:, b] = float('nan')                      # padding, datasets/multiseq.py:340-353
        inputs[m] = x
    return dict(modalities=mods, dims=list(dims), z_dim=z_dim, h_dim=h_dim, min_std=1e-3, inputs=inputs,
                lengths=list(lengths),
                state_dict=bo.init_params(mods, dims, h_dim=h_dim, z_dim=z_dim, seed=seed, scale=1.5))


def oracle_forward(fx, mode, sample, k_flt, eps_flt, eps_smt, dtype=torch.float64):
    import bfvi_oracle as bo
    t_max = max(fx['lengths'])
    flt_dir = 'fwd' if mode in ('ffilter', 'bsmooth') else 'bwd'
    smt_dir = 'fwd' if mode == 'fsmooth' else 'bwd'
    order = lambda d: range(t_max - 1, -1, -1) if d == 'bwd' else range(t_max)
    tape = []
    if sample or k_flt > 1:
        tape += [eps_flt[t].permute(1, 0, 2).contiguous().to(dtype) for t in order(flt_dir)]
    if mode in ('fsmooth', 'bsmooth') and sample:
        tape += [eps_smt[t].permute(1, 0, 2).contiguous().to(dtype) for t in order(smt_dir)]
    params = {k: v.to(dtype) for k, v in fx['state_dict'].items()}
    orc = bo.OracleDMM(fx['modalities'], fx['dims'], params, h_dim=fx['h_dim'], z_dim=fx['z_dim'],
                       min_std=fx['min_std'], draw=bo.NoiseTape(tape))
    with torch.no_grad():
        return orc.forward({k: v.to(dtype) for k, v in fx['inputs'].items()}, fx['lengths'], mode=mode,
                           sample=sample, flt_particles=k_flt)


def run_forward_large(lib, fx, device, mode, sample, k_flt, eps_flt, eps_smt, precision=0):
    model, dists = fixture_model(fx)
    mods = fx['modalities']
    flat, lay = pack_params(lib, model, mods, dists, fx['state_dict'], device)
    t_max, b_dim, z = max(fx['lengths']), len(fx['lengths']), fx['z_dim']
    a = _lib.ForwardArgs()
    keep = []
    a.T, a.B = t_max, b_dim
    for i, m in enumerate(mods):
        if m in fx['inputs']:
            x = fx['inputs'][m].to(device).contiguous()
            keep.append(x)
            a.inputs[i] = x.data_ptr()
    a.mode, a.sample, a.sample_init = _lib.MODE_CODES[mode], int(sample), 0
    a.flt_particles, a.smt_particles, a.precision = k_flt, 1, precision
    for t, f in ((eps_flt, 'eps_flt'), (eps_smt, 'eps_smt')):
        if t is not None:
            t = t.to(device).contiguous()
            keep.append(t)
            setattr(a, f, t.data_ptr())
    outs = [torch.full((t_max, b_dim, z), float('nan'), device=device) for _ in range(4)]
    a.infer_mean, a.infer_std, a.prior_mean, a.prior_std = [t.data_ptr() for t in outs]
    recon = {}
    for i, (m, d) in enumerate(zip(mods, fx['dims'])):
        recon[m] = (torch.full((t_max, b_dim, d), float('nan'), device=device),
                    torch.full((t_max, b_dim, d), float('nan'), device=device))
        a.recon_mean[i], a.recon_std[i] = recon[m][0].data_ptr(), recon[m][1].data_ptr()
    nbytes = C.c_size_t(0)
    lib.call('bfvi_forward_workspace', C.byref(model), C.byref(a), C.byref(nbytes))
    ws = aligned_empty(nbytes.value, device)
    stream = None if device == 'cpu' else C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.call('bfvi_forward', C.byref(model), _lib.ptr(flat), C.byref(a), _lib.ptr(ws), C.c_size_t(nbytes.value),
             stream)
    if device != 'cpu':
        torch.cuda.synchronize()
    return (outs[0], outs[1]), (outs[2], outs[3]), recon


def compare_forward(ours, ref, rtol, atol):
    """Max violation of |a-b| <= atol + rtol|b| over infer / prior / recon; NaNs must coincide."""
    (im, isd), (pm, ps), recon = ours
    (rim, risd), (rpm, rps), rrecon = ref
    pairs = [('infer_mean', im, rim), ('infer_std', isd, risd), ('prior_mean', pm, rpm), ('prior_std', ps, rps)]
    for m in recon:
        pairs += [('recon_mean_' + m, recon[m][0], rrecon[m][0]), ('recon_std_' + m, recon[m][1], rrecon[m][1])]
    bad = []
    for name, a, b in pairs:
        a, b = a.detach().cpu().double(), b.detach().cpu().double().reshape(a.shape)
        if not torch.equal(torch.isnan(a), torch.isnan(b)):
            bad.append((name, 'NaN pattern differs'))
            continue
        ok = ~torch.isnan(b)
        viol = ((a - b).abs() - (atol + rtol * b.abs()))[ok]
        if viol.numel() and viol.max() > 0:
            bad.append((name, float(((a - b).abs()[ok]).max()), float(b.abs()[ok].max())))
    return bad
