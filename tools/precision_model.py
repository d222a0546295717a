"""Numerical model (CPU, fp64 + simulated operand rounding) of the fused GTF backward's precision.

Runs the c3_dims step case of tests/test_gpu_large.py through the fp64 oracle with the GaussianGTF backward
replaced by a model of csrc/bfvi_fused.cuh: which operand is rounded to what before each contraction.  Prints the
worst parameter-gradient errors per variant, so that a precision change can be chosen WITHOUT GPU time.

    python tools/precision_model.py [variant ...]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bfvi_oracle as bo  # noqa: E402
import helpers  # noqa: E402

CASE = dict(z_dim=64, h_dim=512, dims=[16] * 8, t_max=10, lengths=[10] * 9 + [6, 3], seed=2)


def tf32(x):
    """round to nearest even at 10 mantissa bits (cvt.rn.tf32.f32)"""
    i = x.float().contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32).double()


def f16(x, scale=1.0):
    return (x * scale).float().half().double() / scale


def f16_split(x, scale=1.0):
    xs = (x * scale).float()
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return (hi.double() + lo.double()) / scale


class Variant(object):
    def __init__(self, name, **kw):
        self.name = name
        self.dgrad = kw.get('dgrad', tf32)            # operands of the input-gradient contractions
        self.wg_x = kw.get('wg_x', f16)               # hidden-sized operand of the weight gradients (h, dh)
        self.wg_y = kw.get('wg_y', f16)               # latent-sized operand (z, d_g, d_nl)
        self.zz = kw.get('zz', tf32)                  # Z x Z level (lin / std layers) on the launch-sequence GEMMs
        self.mlp = kw.get('mlp', tf32)                # encoder / decoder GEMMs (launch sequence, forward and backward)
        self.pre_err = kw.get('pre_err', 0.0)         # relative error of the hidden pre-activations (ReLU sign flips only)


def f16_ftz(x, scale=1.0):
    h = (x * scale).float().half().double()
    return torch.where(h.abs() < 2.0 ** -14, torch.zeros_like(h), h) / scale


ident = lambda x: x
VARIANTS = {
    'exact': Variant('exact', dgrad=ident, wg_x=ident, wg_y=ident, zz=ident, mlp=ident),
    'mlp_only': Variant('mlp_only', dgrad=ident, wg_x=ident, wg_y=ident, zz=ident),
    'mlp_exact': Variant('mlp_exact', mlp=ident),
    'fused': Variant('fused', mlp=ident, zz=ident),
    'flip1e-7': Variant('flip1e-7', dgrad=ident, wg_x=ident, wg_y=ident, zz=ident, mlp=ident, pre_err=1e-7),
    'flip3e-7': Variant('flip3e-7', dgrad=ident, wg_x=ident, wg_y=ident, zz=ident, mlp=ident, pre_err=3e-7),
    'flip1e-6': Variant('flip1e-6', dgrad=ident, wg_x=ident, wg_y=ident, zz=ident, mlp=ident, pre_err=1e-6),
    'flip3e-6': Variant('flip3e-6', dgrad=ident, wg_x=ident, wg_y=ident, zz=ident, mlp=ident, pre_err=3e-6),
    'flip1e-5': Variant('flip1e-5', dgrad=ident, wg_x=ident, wg_y=ident, zz=ident, mlp=ident, pre_err=1e-5),
    'fused_ftz': Variant('fused_ftz', mlp=ident, zz=ident, wg_x=f16_ftz, wg_y=f16_ftz),
    'fused_scaled12': Variant('fused_scaled12', mlp=ident, zz=ident, wg_x=lambda x: f16(x, 2.0 ** 12) if x.abs().max() < 4 else f16(x), wg_y=lambda x: f16(x, 2.0 ** 12) if x.abs().max() < 4 else f16(x)),
    'fused_dgrad_exact': Variant('fused_dgrad_exact', mlp=ident, zz=ident, dgrad=ident),
    'fused_wg_exact': Variant('fused_wg_exact', mlp=ident, zz=ident, wg_x=ident, wg_y=ident),
    'shipped': Variant('shipped'),
    'dgrad_exact': Variant('dgrad_exact', dgrad=ident),
    'wgx_exact': Variant('wgx_exact', wg_x=ident),
    'wgy_exact': Variant('wgy_exact', wg_y=ident),
    'zz_exact': Variant('zz_exact', zz=ident),
    'wgx_scaled': Variant('wgx_scaled', wg_x=lambda x: f16(x, 2.0 ** 12)),
    'wg_scaled': Variant('wg_scaled', wg_x=lambda x: f16(x, 2.0 ** 12), wg_y=lambda x: f16(x, 2.0 ** 12) if x.abs().max() < 1 else f16(x)),
    'wgy_split': Variant('wgy_split', wg_y=f16_split),
    'wgy_split_dgrad_exact': Variant('wgy_split_dgrad_exact', wg_y=f16_split, dgrad=ident),
    'wgy_split_dgrad_exact_zz': Variant('wgy_split_dgrad_exact_zz', wg_y=f16_split, dgrad=ident, zz=ident),
    'only_wgx': Variant('only_wgx', dgrad=ident, wg_y=ident, zz=ident),
    'only_wgx_scaled': Variant('only_wgx_scaled', dgrad=ident, wg_y=ident, zz=ident, wg_x=lambda x: f16(x, 2.0 ** 12)),
}
STATS = {}


def make_gtf_fn(v):
    class GTF(torch.autograd.Function):
        @staticmethod
        def forward(ctx, z, w0g, b0g, w2g, b2g, wl, bl, w0n, b0n, w2n, b2n, ws, bs):
            pg, pn = F.linear(z, w0g, b0g), F.linear(z, w0n, b0n)
            if v.pre_err > 0:                         # the kernel's pre-activations: sign decided on a perturbed value
                gen = torch.Generator().manual_seed(z.shape[0])
                pg = pg + v.pre_err * pg.abs().mean() * torch.randn(pg.shape, generator=gen, dtype=pg.dtype)
                pn = pn + v.pre_err * pn.abs().mean() * torch.randn(pn.shape, generator=gen, dtype=pn.dtype)
            hg, hn = torch.relu(pg), torch.relu(pn)
            g = F.linear(hg, w2g, b2g)
            nl = F.linear(hn, w2n, b2n)
            lin = F.linear(z, wl, bl)
            a_s = F.linear(nl, ws, bs)
            ctx.save_for_backward(z, hg, hn, nl, w0g, w2g, wl, w0n, w2n, ws)
            return g, nl, lin, a_s

        @staticmethod
        def backward(ctx, d_g, d_nl, d_lin, d_as):
            z, hg, hn, nl, w0g, w2g, wl, w0n, w2n, ws = ctx.saved_tensors
            r, zz = v.dgrad, v.zz
            d_nl = d_nl + zz(d_as) @ zz(ws)
            dws = zz(d_as).t() @ zz(nl)
            dwl = zz(d_lin).t() @ zz(z)
            dhg = (r(d_g) @ r(w2g)) * (hg > 0)
            dhn = (r(d_nl) @ r(w2n)) * (hn > 0)
            dz = r(d_lin) @ r(wl) + r(dhg) @ r(w0g) + r(dhn) @ r(w0n)
            st = STATS.setdefault(v.name, dict(dh=[], dhead=[]))
            st['dh'].append(torch.cat([dhg, dhn], 1).abs().flatten())
            st['dhead'].append(torch.cat([d_g, d_nl], 1).abs().flatten())
            xg, xn, yz = v.wg_x(dhg), v.wg_x(dhn), v.wg_y(z)
            dw0g, dw0n = xg.t() @ yz, xn.t() @ yz
            db0g, db0n = xg.sum(0), xn.sum(0)
            dw2g = v.wg_y(d_g).t() @ v.wg_x(hg)
            dw2n = v.wg_y(d_nl).t() @ v.wg_x(hn)
            return (dz, dw0g, db0g, dw2g, d_g.sum(0), dwl, d_lin.sum(0), dw0n, db0n, dw2n, d_nl.sum(0), dws, d_as.sum(0))
    return GTF


def make_linear_fn(r):
    class Lin(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b):
            ctx.save_for_backward(x, w)
            return r(x) @ r(w).t() + b

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            return r(dy) @ r(w), r(dy).t() @ r(x), dy.sum(0)
    return Lin


class ModelDMM(bo.OracleDMM):
    variant = None

    def _linear(self, key, x):
        if key.startswith('enc.') or key.startswith('dec.'):
            return make_linear_fn(self.variant.mlp).apply(x, self.p[key + '.weight'], self.p[key + '.bias'])
        return bo.OracleDMM._linear(self, key, x)

    def gtf(self, direction, z):
        pre = 'trans.%s' % direction
        p = self.p
        keys = ['z_to_gate.0', 'z_to_gate.2', 'z_lin', 'z_nonlin.0', 'z_nonlin.2', 'z_to_std.0']
        args = []
        for k in keys:
            args += [p['%s.%s.weight' % (pre, k)], p['%s.%s.bias' % (pre, k)]]
        shape = z.shape
        g, nl, lin, a_s = make_gtf_fn(self.variant).apply(z.reshape(-1, shape[-1]), *args)
        gate = torch.sigmoid(g)
        z_std = F.softplus(a_s) + self.min_std
        z_mean = (1 - gate) * lin + gate * nl
        return z_mean.reshape(shape), z_std.reshape(shape)


def run(variant, fx):
    params = {k: v.clone().double().requires_grad_(True) for k, v in fx['state_dict'].items()}
    cls = bo.OracleDMM if variant is None else ModelDMM
    orc = cls(fx['modalities'], fx['dims'], params, h_dim=fx['h_dim'], z_dim=fx['z_dim'], min_std=fx['min_std'],
              draw=bo.step_noise_tape(fx['noise']))
    orc.variant = variant
    cast = lambda d: {k: v.double() for k, v in d.items()}
    loss = orc.step(cast(fx['inputs']), fx['mask'], fx['kld_mult'], fx['rec_mults'], targets=cast(fx['targets']),
                    lengths=fx['lengths'], **fx['step_kwargs'])
    loss.backward()
    return loss.item(), {k: p.grad.clone() for k, p in params.items()}


def step_case(k_train=5, k_match=7, seed=21):
    fx = helpers.large_case(**CASE)
    t_max, b_dim, z = max(fx['lengths']), len(fx['lengths']), fx['z_dim']
    n_sets = len(bo.step_sets(len(fx['modalities'])))
    g = torch.Generator().manual_seed(seed)
    fx['noise'] = {'match': torch.randn(2, k_match, z, generator=g),
                   'filt': torch.randn(n_sets, t_max, b_dim, 1, z, generator=g),
                   'sflt': torch.randn(n_sets, t_max, b_dim, k_train, z, generator=g),
                   'ssmt': torch.randn(n_sets, t_max, b_dim, 1, z, generator=g)}
    fx['targets'] = fx['inputs']
    mask = torch.zeros(t_max, b_dim, 1, dtype=torch.bool)
    for b, n in enumerate(fx['lengths']):
        mask[:n, b] = True
    fx['mask'] = mask
    fx['kld_mult'] = 0.8
    fx['rec_mults'] = {m: 1.0 / (d * len(fx['dims'])) for m, d in zip(fx['modalities'], fx['dims'])}
    fx['step_kwargs'] = {'train_particles': k_train, 'match_particles': k_match}
    return fx


def main():
    names = sys.argv[1:] or ['exact', 'shipped']
    fx = step_case()
    _, ref = run(None, fx)
    for n in names:
        v = VARIANTS[n]
        _, g = run(v, fx)
        errs = sorted(((g[k] - ref[k]).norm().item() / ref[k].norm().item(), k) for k in ref if ref[k].norm() > 0)
        print('%-28s worst: %s' % (n, '  '.join('%s %.2e' % (k.replace('trans.', ''), e) for e, k in [e for e in errs[::-1] if "trans" in e[1]][:14])))
        if n in STATS:
            for what in ('dhead', 'dh'):
                x = torch.cat(STATS[n][what])
                x = x[x > 0]
                q = torch.quantile(x[torch.randperm(x.numel())[:2000000]], torch.tensor([0.01, 0.1, 0.5, 0.9, 0.99], dtype=x.dtype))
                print('   |%s| quantiles 1/10/50/90/99 %%: %s  max %.2e  frac < 6.1e-5 (fp16 subnormal): %.3f' %
                      (what, ' '.join('%.1e' % t for t in q), x.max().item(), (x < 6.1e-5).double().mean().item()))


if __name__ == '__main__':
    main()
