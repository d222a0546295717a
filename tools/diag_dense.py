"""Accuracy of bfvi_linear_tf32 / bfvi_wgrad_tf32 (error-compensated 3xTF32) at the dense-layer shapes of the image modules
(feat_dim 4096 <-> z_dim 256, few to many rows) against fp64, per 512-column block of the output."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_dmm_b200 import _lib  # noqa: E402

lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = []
g = torch.Generator(device='cuda').manual_seed(0)
for rows, n_in, n_out, act in [(6, 256, 4096, 1), (625, 256, 4096, 1), (625, 256, 4096, 0), (625, 256, 512, 0), (625, 256, 1024, 0),
                               (6, 4096, 256, 0), (625, 4096, 256, 0), (625, 4096, 4096, 0), (6, 4096, 256, 16)]:
    x = torch.randn(rows, n_in, device='cuda', generator=g)
    w = torch.randn(n_out, n_in, device='cuda', generator=g) / n_in ** 0.5
    b = torch.randn(n_out, device='cuda', generator=g)
    y = torch.full((rows, n_out), float('nan'), device='cuda')
    lib.call('bfvi_linear_tf32', _lib.ptr(x), n_in, _lib.ptr(w), n_in, _lib.ptr(b), _lib.ptr(y), n_out, rows, n_in, n_out, act, st())
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t() + b.double()
    if act & 1:
        ref = ref.clamp_min(0)
    err = (y.double() - ref)
    blocks = [round((err[:, i:i + 512].norm() / ref[:, i:i + 512].norm()).item(), 9) for i in range(0, n_out, 512)]
    out.append({'linear': [rows, n_in, n_out, act], 'rel': (err.norm() / ref.norm()).item(), 'max': err.abs().max().item(),
                'blocks': blocks[:8]})
for rows, n_out, n_in in [(6, 256, 4096), (625, 256, 4096), (6, 4096, 256), (625, 4096, 256), (20, 256, 4096)]:
    dy = torch.randn(rows, n_out, device='cuda', generator=g)
    x = torch.randn(rows, n_in, device='cuda', generator=g)
    dy_t, x_t = dy.t().contiguous(), x.t().contiguous()
    for acc in (0, 1):
        dw = torch.zeros(n_out, n_in, device='cuda')
        lib.call('bfvi_wgrad_tf32', _lib.ptr(dy_t), rows, _lib.ptr(x_t), rows, _lib.ptr(dw), n_in, rows, n_out, n_in, acc, 0, st())
        torch.cuda.synchronize()
        ref = dy.double().t() @ x.double()
        err = dw.double() - ref
        out.append({'wgrad': [rows, n_out, n_in, acc], 'rel': (err.norm() / ref.norm()).item(), 'max': err.abs().max().item()})
print(json.dumps(out, indent=0))
