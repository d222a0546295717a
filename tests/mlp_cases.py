"""bfvi_mlp_fwd / bfvi_mlp_bwd (per-modality MLPs of the composed path, weights by pointer) against plain torch fp64
restatements of common.GaussianMLP / common.CategoricalMLP / Embedding -> ReLU -> GaussianMLP
(models/common.py:9-41, models/dmm.py:78-82).  Shared by the emulator (CPU) and B200 tests."""
import ctypes as C

import torch
import torch.nn.functional as F

from multimodal_dmm_b200 import _lib

CASES = {
    # name: (kind, n_in, h_dim, n_out, n_classes, rows)
    'gauss_enc_small': ('gauss_enc', 3, 10, 7, 0, 37),
    'gauss_dec_small': ('gauss_dec', 7, 12, 5, 0, 40),
    'gauss_enc_c4': ('gauss_enc', 24, 256, 256, 0, 300),
    'softmax_dec': ('softmax_dec', 9, 20, 10, 0, 33),
    'softmax_dec_c4': ('softmax_dec', 256, 256, 10, 0, 200),
    'embed_enc': ('embed_enc', 16, 16, 6, 10, 45),
    'embed_enc_c4': ('embed_enc', 256, 256, 256, 10, 130),
}


def make(name, seed=0):
    kind, n_in, h, n_out, n_cls, rows = CASES[name]
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    w = {'w1': r(h, n_in) / n_in ** 0.5, 'b1': 0.1 * r(h), 'wa': r(n_out, h) / h ** 0.5, 'ba': 0.1 * r(n_out)}
    if kind != 'softmax_dec':
        w['wb'], w['bb'] = r(n_out, h) / h ** 0.5, 0.1 * r(n_out)
    if kind == 'embed_enc':
        w['emb'] = r(n_cls, h)
        x = torch.randint(0, n_cls, (rows,), generator=g).float()
        x[3] = float('nan')
    else:
        x = r(rows, n_in)
        if kind == 'gauss_enc':
            x[2, 0] = float('nan')
            x[5] = float('nan')
    d_a, d_b = r(rows, n_out), r(rows, n_out)
    return kind, w, x, d_a, d_b


def reference(kind, w, x, d_a, d_b):
    """fp64 torch restatement + autograd."""
    p = {k: v.double().clone().requires_grad_(True) for k, v in w.items()}
    xd = x.double()
    mask = None
    if kind in ('gauss_enc', 'embed_enc'):
        mask = ~torch.isnan(xd.reshape(xd.shape[0], -1)).any(dim=1)
        xd = torch.nan_to_num(xd, nan=0.0)
    if kind == 'embed_enc':
        inp = torch.relu(F.embedding(xd.long(), p['emb']))
    else:
        inp = xd.clone().requires_grad_(kind != 'gauss_enc')
    h = torch.relu(F.linear(inp, p['w1'], p['b1']))
    a = F.linear(h, p['wa'], p['ba'])
    if kind == 'softmax_dec':
        out = (torch.softmax(a, dim=1),)
        loss = (out[0] * d_a.double()).sum()
    else:
        out = (a, F.softplus(F.linear(h, p['wb'], p['bb'])) + 1e-3)
        loss = (out[0] * d_a.double()).sum() + (out[1] * d_b.double()).sum()
    loss.backward()
    d_x = inp.grad if (kind in ('gauss_dec', 'softmax_dec')) else None
    return [o.detach() for o in out], mask, {k: v.grad for k, v in p.items()}, d_x


def run(lib, device, kind, w, x, d_a, d_b):
    """Through the C ABI on `device` ('cpu' = emulator library)."""
    dev = torch.device(device)
    wd = {k: v.to(dev).contiguous() for k, v in w.items()}
    xd = x.to(dev).contiguous()
    rows = x.shape[0]
    d = _lib.MlpDesc()
    d.emb = wd['emb'].data_ptr() if 'emb' in wd else 0
    d.w1, d.b1, d.wa, d.ba = (wd[k].data_ptr() for k in ('w1', 'b1', 'wa', 'ba'))
    d.wb = wd['wb'].data_ptr() if 'wb' in wd else 0
    d.bb = wd['bb'].data_ptr() if 'bb' in wd else 0
    d.n_in, d.h_dim, d.n_out = w['w1'].shape[1], w['w1'].shape[0], w['wa'].shape[0]
    d.n_classes = w['emb'].shape[0] if 'emb' in w else 0
    d.head = _lib.HEAD_SOFTMAX if kind == 'softmax_dec' else _lib.HEAD_GAUSSIAN
    d.nan_mask = 1 if kind in ('gauss_enc', 'embed_enc') else 0
    d.min_std = 1e-3
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream if dev.type == 'cuda' else 0)

    def ws_for(backward):
        n = C.c_size_t(0)
        lib.call('bfvi_mlp_workspace', C.byref(d), rows, backward, C.byref(n))
        buf = torch.empty(n.value + 256, dtype=torch.uint8, device=dev)
        return buf[(-buf.data_ptr()) % 256:], n.value

    out_a = torch.full((rows, d.n_out), float('nan'), device=dev)
    out_b = torch.full((rows, d.n_out), float('nan'), device=dev) if d.head == _lib.HEAD_GAUSSIAN else None
    mask = torch.zeros(rows, dtype=torch.uint8, device=dev) if d.nan_mask else None
    ws, n = ws_for(0)
    lib.call('bfvi_mlp_fwd', C.byref(d), _lib.ptr(xd), rows, _lib.ptr(out_a), _lib.ptr(out_b), _lib.ptr(mask),
             _lib.ptr(ws), C.c_size_t(n), st)
    grads = {k: torch.zeros_like(v) for k, v in wd.items()}
    g = _lib.MlpGrads()
    for k in ('emb', 'w1', 'b1', 'wa', 'ba', 'wb', 'bb'):
        setattr(g, k, grads[k].data_ptr() if k in grads else 0)
    want_dx = kind in ('gauss_dec', 'softmax_dec')
    d_x = torch.full((rows, d.n_in), float('nan'), device=dev) if want_dx else None
    da, db = d_a.to(dev).contiguous(), d_b.to(dev).contiguous()
    ws, n = ws_for(1)
    lib.call('bfvi_mlp_bwd', C.byref(d), C.byref(g), _lib.ptr(xd), rows, _lib.ptr(out_a), _lib.ptr(out_b), _lib.ptr(da),
             _lib.ptr(db if d.head == _lib.HEAD_GAUSSIAN else None), _lib.ptr(d_x), _lib.ptr(ws), C.c_size_t(n), st)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    outs = [out_a.cpu()] + ([out_b.cpu()] if out_b is not None else [])
    return outs, None if mask is None else mask.cpu().bool(), {k: v.cpu() for k, v in grads.items()}, \
        None if d_x is None else d_x.cpu()


def check(name, lib, device, tol):
    kind, w, x, d_a, d_b = make(name)
    outs, mask, grads, d_x = run(lib, device, kind, w, x, d_a, d_b)
    r_outs, r_mask, r_grads, r_dx = reference(kind, w, x, d_a, d_b)
    rel = lambda a, b: ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
    if r_mask is not None:
        assert torch.equal(mask, r_mask)                       # masks bit-exact
    for a, b in zip(outs, r_outs):
        assert rel(a, b) < tol, (name, 'out', rel(a, b))
    for k, gr in r_grads.items():
        assert rel(grads[k], gr) < tol, (name, k, rel(grads[k], gr))
    if r_dx is not None:
        assert rel(d_x, r_dx) < tol, (name, 'd_x', rel(d_x, r_dx))
