"""Fused on-chip GaussianGTF kernels (bfvi_gtf_fwd / bfvi_gtf_bwd, csrc/bfvi_fused.cuh) on B200 against a
plain fp64 restatement of models/common.py:62-68 and its autograd.

Two kinds of inputs:
 * "lattice" inputs whose every intermediate is exactly representable in TF32 and FP16 (small multiples of
   powers of two): the tensor-core result must then equal the fp64 reference to fp32 rounding — any indexing,
   swizzle, descriptor or pipeline-ordering error shows up as an O(1) difference;
 * random inputs: forward heads at fp32-class accuracy (error-compensated FP16 hi / lo products — the ReLU signs
   must match the fp64 reference's, or whole gradient terms flip), gradients at the tolerance single-pass
   round-to-nearest TF32 operands give (FP16 operands for the H-wide weight gradients).
"""
import ctypes as C

import pytest
import torch

import bfvi_oracle as bo
import helpers
from multimodal_dmm_b200 import _lib

pytestmark = pytest.mark.gpu
Z = 64


def make_params(h_dim, seed, lattice):
    mods, dims = ['m0'], [3]
    sd = bo.init_params(mods, dims, h_dim=h_dim, z_dim=Z, seed=seed, scale=1.5)
    if lattice:
        g = torch.Generator().manual_seed(seed)
        for k, v in sd.items():
            if k.startswith('trans.'):
                # weights in {0, +-1/4} (sparse), biases multiples of 1/8
                if v.dim() == 2:
                    w = torch.randint(-1, 2, v.shape, generator=g).float() * 0.25
                    w *= (torch.rand(v.shape, generator=g) < 0.25).float()
                    sd[k] = w
                else:
                    sd[k] = torch.randint(-4, 5, v.shape, generator=g).float() / 8
    return mods, dims, sd


def reference(sd, d, z, d_heads=None):
    """fp64 GTF heads (pre-activations) and, given head gradients, d_z and parameter gradients."""
    p = 'trans.%s.' % d
    names = ['z_to_gate.0.weight', 'z_to_gate.0.bias', 'z_to_gate.2.weight', 'z_to_gate.2.bias', 'z_lin.weight',
             'z_lin.bias', 'z_nonlin.0.weight', 'z_nonlin.0.bias', 'z_nonlin.2.weight', 'z_nonlin.2.bias',
             'z_to_std.0.weight', 'z_to_std.0.bias']
    w = {n: sd[p + n].double().clone().requires_grad_(True) for n in names}
    z = z.double().clone().requires_grad_(True)
    h1 = torch.relu(z @ w['z_to_gate.0.weight'].T + w['z_to_gate.0.bias'])
    g = h1 @ w['z_to_gate.2.weight'].T + w['z_to_gate.2.bias']
    h3 = torch.relu(z @ w['z_nonlin.0.weight'].T + w['z_nonlin.0.bias'])
    nl = h3 @ w['z_nonlin.2.weight'].T + w['z_nonlin.2.bias']
    lin = z @ w['z_lin.weight'].T + w['z_lin.bias']
    a_s = nl @ w['z_to_std.0.weight'].T + w['z_to_std.0.bias']
    heads = (g, nl, lin, a_s)
    if d_heads is None:
        return [t.detach() for t in heads], None, None
    loss = sum((t * dt.double()).sum() for t, dt in zip(heads, d_heads))
    loss.backward()
    return [t.detach() for t in heads], z.grad, {p + n: w[n].grad for n in names}


def run(lib, mods, dims, sd, h_dim, d, z, d_heads=None):
    dev = 'cuda'
    model = _lib.make_model(dims, ['Normal'], Z, h_dim, 1e-3)
    flat, lay = helpers.pack_params(lib, model, mods, ['Normal'], sd, dev)
    rows = z.shape[0]
    nbytes = int(lib.dll.bfvi_gtf_workspace(C.byref(model), rows))
    assert nbytes > 0
    ws = helpers.aligned_empty(nbytes, dev)
    zc = z.to(dev).contiguous()
    outs = [torch.full((rows, Z), float('nan'), device=dev) for _ in range(4)]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    direction = 1 if d == 'bwd' else 0
    lib.call('bfvi_gtf_fwd', C.byref(model), _lib.ptr(flat), direction, _lib.ptr(zc), rows, *[_lib.ptr(t) for t in outs],
             1 if d_heads is not None else 0, _lib.ptr(ws), C.c_size_t(nbytes), st)
    torch.cuda.synchronize()
    if d_heads is None:
        return [t.cpu() for t in outs], None, None
    grads = torch.zeros_like(flat)
    dz = torch.full((rows, Z), float('nan'), device=dev)
    dh = [t.to(dev).contiguous() for t in d_heads]
    lib.call('bfvi_gtf_bwd', C.byref(model), _lib.ptr(flat), _lib.ptr(grads), direction, _lib.ptr(zc), _lib.ptr(outs[1]), rows,
             _lib.ptr(dh[0]), _lib.ptr(dh[1]), _lib.ptr(dh[2]), _lib.ptr(dh[3]), _lib.ptr(dz), _lib.ptr(ws),
             C.c_size_t(nbytes), st)
    torch.cuda.synchronize()
    g = helpers.unpack(lib, model, mods, ['Normal'], grads, sd)
    return [t.cpu() for t in outs], dz.cpu(), g


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize('h_dim', [128, 512])
@pytest.mark.parametrize('rows', [1, 100, 128, 129, 1000, 148 * 128 + 77])
def test_lattice_forward_is_exact(h_dim, rows):
    lib = _lib.load()
    mods, dims, sd = make_params(h_dim, 3, lattice=True)
    g = torch.Generator().manual_seed(rows)
    z = torch.randint(-2, 3, (rows, Z), generator=g).float() * 0.5
    ours, _, _ = run(lib, mods, dims, sd, h_dim, 'fwd', z)
    ref, _, _ = reference(sd, 'fwd', z)
    assert 'gtf_fwd_fused' in ';'.join(lib.last_dispatch())
    for name, a, b in zip(('gate', 'nonlin', 'lin'), ours, ref):        # exact: every operand is a lattice point
        assert torch.isfinite(a).all(), name
        assert (a.double() - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item()), (name, (a.double() - b).abs().max())
    # the std head contracts the hi / lo split of the nonlinear head: fp32-class
    assert rel(ours[3], ref[3]) < 1e-5


@pytest.mark.parametrize('h_dim,rows,d', [(128, 300, 'fwd'), (512, 128, 'bwd'), (512, 1000, 'fwd'), (256, 64 * 37 + 5, 'bwd')])
def test_lattice_backward_is_exact(h_dim, rows, d):
    lib = _lib.load()
    mods, dims, sd = make_params(h_dim, 5, lattice=True)
    g = torch.Generator().manual_seed(rows)
    z = torch.randint(-2, 3, (rows, Z), generator=g).float() * 0.5
    # head gradients on a lattice too; the std head's gradient is kept zero so that d_nl stays on the lattice
    d_heads = [torch.randint(-2, 3, (rows, Z), generator=g).float() * 0.5 for _ in range(3)] + [torch.zeros(rows, Z)]
    ours, dz, grads = run(lib, mods, dims, sd, h_dim, d, z, d_heads)
    ref, dz_ref, g_ref = reference(sd, d, z, d_heads)
    ran = ';'.join(lib.last_dispatch())
    assert 'gtf_bwd_fused' in ran and 'wgrad16' in ran, ran
    assert (dz.double() - dz_ref).abs().max().item() <= 1e-5 * max(1.0, dz_ref.abs().max().item())
    for k, gr in g_ref.items():
        if 'z_to_std' in k:
            continue
        err = (grads[k].double() - gr).abs().max().item()
        assert err <= 1e-5 * max(1.0, gr.abs().max().item()), (k, err, gr.abs().max().item())


@pytest.mark.parametrize('h_dim,rows,d', [(512, 2304, 'bwd'), (512, 5 * 1152, 'fwd'), (128, 700, 'fwd'), (1024, 640, 'bwd')])
def test_random_inputs_at_tf32_tolerance(h_dim, rows, d):
    lib = _lib.load()
    mods, dims, sd = make_params(h_dim, 7, lattice=False)
    g = torch.Generator().manual_seed(11)
    z = torch.randn(rows, Z, generator=g)
    d_heads = [torch.randn(rows, Z, generator=g) * s for s in (0.3, 1.0, 1.0, 0.5)]
    ours, dz, grads = run(lib, mods, dims, sd, h_dim, d, z, d_heads)
    ref, dz_ref, g_ref = reference(sd, d, z, d_heads)
    errs = {n: rel(a, b) for n, a, b in zip(('gate', 'nonlin', 'lin', 'std'), ours, ref)}
    assert max(errs.values()) < 2e-5, errs              # forward: error-compensated FP16 hi/lo products (fp32-class)
    assert rel(dz, dz_ref) < 1.5e-3, rel(dz, dz_ref)    # input gradient: single-pass TF32 operands (2^-11 each)
    gerrs = {k: rel(grads[k], gr) for k, gr in g_ref.items()}
    assert max(gerrs.values()) < 2e-3, sorted(gerrs.items(), key=lambda kv: -kv[1])[:3]
