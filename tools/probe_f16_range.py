"""Dynamic-range probe of the fused GTF kernels (csrc/bfvi_fused.cuh) on B200.

The forward contracts FP16 hi / lo splits, the weight gradients FP16 operand tiles: both lose precision when operands
approach the FP16 subnormal range (|x| < 6.1e-5).  This probe scales the inputs (z for the forward, the head gradients for
the backward; biases zeroed so that every output is homogeneous in the scaled input) and prints relative errors against
the fp64 restatement, next to what a CPU model of the operand roundings predicts (--model: CPU only).

    python tools/probe_f16_range.py            (GPU, through bfvi_gtf_fwd / bfvi_gtf_bwd)
    python tools/probe_f16_range.py --model    (CPU model of the roundings, gradual underflow and flush-to-zero)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_fused as tf  # noqa: E402

H, ROWS, Z = 512, 2304, 64


def tf32(x):
    i = x.float().contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32).double()


def f16(x, ftz=False):
    h = x.float().half().double()
    return torch.where(h.abs() < 2.0 ** -14, torch.zeros_like(h), h) if ftz else h


def model_bwd(sd, d, z, d_heads, ftz):
    p = 'trans.%s.' % d
    w = {k[len(p):]: v.double() for k, v in sd.items() if k.startswith(p)}
    z = z.double()
    d_g, d_nl, d_lin, d_as = [t.double() for t in d_heads]
    pre_g = z @ w['z_to_gate.0.weight'].T + w['z_to_gate.0.bias']
    pre_n = z @ w['z_nonlin.0.weight'].T + w['z_nonlin.0.bias']
    hg, hn = torch.relu(pre_g), torch.relu(pre_n)
    d_nl = d_nl + d_as @ w['z_to_std.0.weight']
    dhg = (tf32(d_g) @ tf32(w['z_to_gate.2.weight'])) * (pre_g > 0)
    dhn = (tf32(d_nl) @ tf32(w['z_nonlin.2.weight'])) * (pre_n > 0)
    dz = tf32(d_lin) @ tf32(w['z_lin.weight']) + tf32(dhg) @ tf32(w['z_to_gate.0.weight']) + tf32(dhn) @ tf32(w['z_nonlin.0.weight'])
    q = lambda x: f16(x, ftz)
    g = {p + 'z_to_gate.0.weight': q(dhg).T @ q(z), p + 'z_nonlin.0.weight': q(dhn).T @ q(z),
         p + 'z_to_gate.0.bias': q(dhg).sum(0), p + 'z_nonlin.0.bias': q(dhn).sum(0),
         p + 'z_to_gate.2.weight': q(d_g).T @ q(hg), p + 'z_nonlin.2.weight': q(d_nl).T @ q(hn)}
    return dz, g


def main():
    model = '--model' in sys.argv
    mods, dims, sd = tf.make_params(H, 7, lattice=False)
    for k in list(sd):
        if k.startswith('trans.') and k.endswith('.bias'):
            sd[k] = torch.zeros_like(sd[k])
    g = torch.Generator().manual_seed(11)
    z0 = torch.randn(ROWS, Z, generator=g)
    dh0 = [torch.randn(ROWS, Z, generator=g) * s for s in (0.3, 1.0, 1.0, 0.5)]
    lib = None
    if not model:
        from multimodal_dmm_b200 import _lib
        lib = _lib.load()
    print('== forward: z scaled by 2^-k (biases zero: heads are homogeneous in z)')
    for k in (0, 4, 8, 12, 16):
        z = z0 * 2.0 ** -k
        ref, _, _ = tf.reference(sd, 'bwd', z)
        if model:
            continue
        ours, _, _ = tf.run(lib, mods, dims, sd, H, 'bwd', z)
        print('  k=%2d  ' % k + '  '.join('%s %.2e' % (n, tf.rel(a, b)) for n, a, b in zip(('gate', 'nonlin', 'lin', 'std'), ours, ref)))
    print('== backward: head gradients scaled by 2^-k')
    keys = ['z_to_gate.0.weight', 'z_nonlin.0.weight', 'z_to_gate.0.bias', 'z_to_gate.2.weight', 'z_nonlin.2.weight', 'z_lin.weight']
    for k in (0, 6, 10, 14, 18, 22):
        d_heads = [t * 2.0 ** -k for t in dh0]
        ref, dz_ref, g_ref = tf.reference(sd, 'bwd', z0, d_heads)
        if model:
            for ftz in (False, True):
                dz, gr = model_bwd(sd, 'bwd', z0, d_heads, ftz)
                print('  k=%2d model%s dz %.2e  ' % (k, ' ftz' if ftz else '    ', tf.rel(dz, dz_ref)) +
                      '  '.join('%s %.2e' % (n, tf.rel(gr['trans.bwd.' + n], g_ref['trans.bwd.' + n])) for n in keys if 'trans.bwd.' + n in gr))
            continue
        ours, dz, grads = tf.run(lib, mods, dims, sd, H, 'bwd', z0, d_heads)
        print('  k=%2d  dz %.2e  ' % (k, tf.rel(dz, dz_ref)) +
              '  '.join('%s %.2e' % (n, tf.rel(grads['trans.bwd.' + n], g_ref['trans.bwd.' + n])) for n in keys))


if __name__ == '__main__':
    main()
