"""bfvi_adam_step (fused clip_grad_norm_ + Adam on the flat buffers) against torch on CPU
through the emulated kernels; the GPU twin is in tests/test_gpu_model.py."""
import ctypes as C

import pytest
import torch

import helpers
from multimodal_dmm_b200 import _lib


def torch_reference(p0, grads, lr, wd, max_norm):
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=lr, weight_decay=wd)
    for g in grads:
        p.grad = g.clone()
        if max_norm:
            torch.nn.utils.clip_grad_norm_([p], max_norm)
        opt.step()
    return p.detach()


def run_ours(lib, device, p0, grads, lr, wd, max_norm):
    p = p0.clone().to(device)
    m, v, norm = torch.zeros_like(p), torch.zeros_like(p), torch.zeros(1, device=device)
    st = None if device == 'cpu' else C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for t, g in enumerate(grads, 1):
        g = g.to(device)
        lib.call('bfvi_adam_step', _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), p.numel(), lr, 0.9, 0.999, 1e-8,
                 wd, t, 1.0, max_norm or 0.0, _lib.ptr(norm), st)
    return p.cpu()


@pytest.mark.parametrize('wd,max_norm', [(0.0, None), (1e-2, None), (0.0, 0.5), (1e-3, 10.0)])
def test_adam_step_matches_torch(wd, max_norm):
    lib = helpers.emu_library()
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1937, generator=g)
    grads = [torch.randn(1937, generator=g) * s for s in (1.0, 0.1, 3.0, 0.5)]
    ours = run_ours(lib, 'cpu', p0, grads, 3e-3, wd, max_norm)
    ref = torch_reference(p0, grads, 3e-3, wd, max_norm)
    assert torch.allclose(ours, ref, rtol=1e-5, atol=1e-6), (ours - ref).abs().max()
