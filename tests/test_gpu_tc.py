"""tcgen05 TF32 linear layer (bfvi_linear_tf32) against a float64 reference on the
TF32-rounded operands (tight) and against plain fp32 torch (TF32 tolerance)."""
import ctypes as C

import pytest
import torch

from multimodal_dmm_b200 import _lib

pytestmark = pytest.mark.gpu


def round_tf32(x):
    """Round-to-nearest-even to 10 explicit mantissa bits, like cvt.rn.tf32.f32 (the kernel's F2FP)."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def run(lib, x, w, b, act, ldx=None, ldy=None):
    n_rows, n_in = x.shape
    n_out = w.shape[0]
    ldx = ldx or n_in
    ldy = ldy or n_out
    xs = torch.zeros(n_rows, ldx, device='cuda')
    xs[:, :n_in] = x
    y = torch.full((n_rows, ldy), float('nan'), device='cuda')
    lib.call('bfvi_linear_tf32', _lib.ptr(xs), ldx, _lib.ptr(w), w.stride(0), _lib.ptr(b), _lib.ptr(y), ldy,
             n_rows, n_in, n_out, act, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize('shape', [(128, 64, 64), (300, 16, 512), (1000, 512, 64), (257, 20, 40),
                                   (5000, 256, 256), (33, 7, 3), (4096, 64, 512), (129, 512, 16)])
@pytest.mark.parametrize('act', [0, 1])
def test_linear_tf32_matches_reference(shape, act):
    lib = _lib.load()
    m, k, n = shape
    g = torch.Generator(device='cuda').manual_seed(m * 31 + k * 7 + n)
    x = torch.randn(m, k, device='cuda', generator=g)
    w = torch.randn(n, k, device='cuda', generator=g) / k ** 0.5
    b = torch.randn(n, device='cuda', generator=g)
    pad = 4 if k % 4 else 0
    y = run(lib, x, w, b, act | 16, ldx=k + pad, ldy=n + 3)          # single-pass TF32
    assert torch.isnan(y[:, n:]).all()                       # nothing written outside (n_rows, n_out)
    y = y[:, :n]
    ref = round_tf32(x).double() @ round_tf32(w).double().t() + b.double()
    exact = x.double() @ w.double().t() + b.double()
    if act:
        ref, exact = ref.clamp_min(0), exact.clamp_min(0)
    scale = ref.abs().max().item()
    assert (y.double() - ref).abs().max().item() < 2e-5 * scale, (y.double() - ref).abs().max().item()
    assert (y.double() - exact).abs().max().item() < 5e-3 * scale
    y3 = run(lib, x, w, b, act, ldx=k + pad, ldy=n + 3)[:, :n]        # error-compensated 3xTF32
    assert (y3.double() - exact).abs().max().item() < 3e-6 * scale, (y3.double() - exact).abs().max().item()


def test_linear_tf32_argument_errors():
    lib = _lib.load()
    x = torch.zeros(4, 4, device='cuda')
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_linear_tf32', _lib.ptr(x), 2, _lib.ptr(x), 4, None, _lib.ptr(x), 4, 4, 4, 4, 0, None)
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_linear_tf32', _lib.ptr(x), 4, _lib.ptr(x), 4, None, _lib.ptr(x), 4, 4, 4, 4, 7, None)


@pytest.mark.parametrize('shape', [(256, 64, 64), (1000, 512, 64), (77, 20, 40), (4096, 64, 512), (333, 16, 512),
                                   (50, 5, 3)])
def test_wgrad_tf32_matches_reference(shape):
    """dW (+)= dY^T X from the transposed copies (contraction over the rows)."""
    lib = _lib.load()
    rows, n_out, n_in = shape
    g = torch.Generator(device='cuda').manual_seed(rows + n_out)
    dy = torch.randn(rows, n_out, device='cuda', generator=g)
    x = torch.randn(rows, n_in, device='cuda', generator=g)
    dw0 = torch.randn(n_out, n_in, device='cuda', generator=g)
    dy_t, x_t = dy.t().contiguous(), x.t().contiguous()
    exact = dy.double().t() @ x.double()
    scale = exact.abs().max().item()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for flags, tol in ((0, 3e-6 + 2e-8 * rows), (16, 5e-3)):      # tensor-core accumulation error grows with K
        dw = dw0.clone()
        lib.call('bfvi_wgrad_tf32', _lib.ptr(dy_t), rows, _lib.ptr(x_t), rows, _lib.ptr(dw), n_in, rows, n_out,
                 n_in, 1, flags, st)
        torch.cuda.synchronize()
        err = (dw.double() - dw0.double() - exact).abs().max().item()
        assert err < tol * scale, (flags, err, scale)
        dw2 = torch.full_like(dw0, float('nan'))
        lib.call('bfvi_wgrad_tf32', _lib.ptr(dy_t), rows, _lib.ptr(x_t), rows, _lib.ptr(dw2), n_in, rows, n_out,
                 n_in, 0, flags, st)
        torch.cuda.synchronize()
        assert (dw2.double() - exact).abs().max().item() < tol * scale
