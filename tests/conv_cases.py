"""The image-module kernels (bfvi_conv_gather / _scatter / _wgrad, bfvi_chan_bias_grad, bfvi_bn2d_fwd / _bwd,
bfvi_sigmoid_bwd) through the C ABI against torch fp64 restatements of nn.Conv2d, nn.ConvTranspose2d and
nn.BatchNorm2d -> ReLU (models/common.py:70-107, 148).  Shared by the emulator (CPU) and B200 tests."""
import ctypes as C

import torch
import torch.nn.functional as F

from multimodal_dmm_b200 import _lib

# name: (N, c_small, c_big, h_big, w_big, kernel, stride, padding, transposed)
#   conv:   x (N, c_big, h_big, w_big) -> y (N, c_small, h_small, w_small)
#   deconv: x (N, c_small, h_small, w_small) -> y (N, c_big, h_big, w_big)
CONV_CASES = {
    'conv3s2_odd': (3, 5, 3, 10, 9, 3, 2, 1, False),
    'conv3s2_w1': (2, 16, 3, 64, 64, 3, 2, 1, False),         # Weizmann conv_stack.0
    'conv3s2_w3': (2, 64, 32, 16, 16, 3, 2, 1, False),        # Weizmann conv_stack.2
    'conv5s1': (2, 6, 4, 9, 11, 5, 1, 2, False),
    'conv3s3p0': (2, 3, 7, 12, 9, 3, 3, 0, False),
    'conv4s2_chunked': (1, 8, 36, 6, 8, 4, 2, 1, False),   # 36 x 16 taps x 16 > one weight stage
    'deconv4s2_odd': (3, 5, 3, 10, 14, 4, 2, 1, True),
    'deconv4s2_w1': (2, 64, 32, 16, 16, 4, 2, 1, True),       # Weizmann deconv_stack.0
    'deconv4s2_w3': (2, 16, 3, 64, 64, 4, 2, 1, True),        # Weizmann deconv_stack.2 (+ sigmoid)
    'deconv3s1': (2, 4, 6, 7, 8, 3, 1, 1, True),
    'deconv5s3': (2, 3, 2, 18, 15, 5, 3, 1, True),
}

# name: (N, C, H, W, training, relu)
BN_CASES = {
    'bn_train_relu': (5, 7, 6, 5, True, True),
    'bn_train_plain': (3, 4, 8, 8, True, False),
    'bn_eval_relu': (4, 6, 5, 7, False, True),
    'bn_big': (6, 16, 32, 32, True, True),
}


# name: (rows, n_in, n_out, relu)
DENSE_CASES = {
    'dense_odd_relu': (37, 19, 70, True),
    'dense_odd': (5, 130, 3, False),
    'dense_splitk': (3, 1100, 5, False),            # one output tile, long contraction: K slices + atomicAdd
    'dense_feat_to_z': (6, 4096, 256, False),       # Weizmann feat_to_z_mean / feat_to_z_std.0
    'dense_z_to_feat': (20, 256, 4096, True),       # Weizmann z_to_feat.0 -> ReLU
}


def small_size(big, k, s, p, transposed):
    if not transposed:
        return (big + 2 * p - k) // s + 1
    assert (big + 2 * p - k) % s == 0
    return (big + 2 * p - k) // s + 1


def geom(case):
    N, cs, cb, hb, wb, k, s, p, tr = CONV_CASES[case]
    g = _lib.ConvGeom()
    g.n, g.c_small, g.c_big, g.h_big, g.w_big, g.kernel, g.stride, g.padding = N, cs, cb, hb, wb, k, s, p
    g.h_small, g.w_small = small_size(hb, k, s, p, tr), small_size(wb, k, s, p, tr)
    return g, tr


def make_conv(case, seed=0):
    g, tr = geom(case)
    gen = torch.Generator().manual_seed(seed)
    r = lambda *sh: torch.randn(*sh, generator=gen)
    w = r(g.c_small, g.c_big, g.kernel, g.kernel) / (g.kernel * (g.c_big if not tr else g.c_small) ** 0.5)
    small = r(g.n, g.c_small, g.h_small, g.w_small)
    big = r(g.n, g.c_big, g.h_big, g.w_big)
    bias = 0.3 * r(g.c_big if tr else g.c_small)
    return g, tr, w, bias, small, big


def stream_of(dev):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream if dev.type == 'cuda' else 0)


def scratch_for(lib, channels, dev):
    n = lib.dll.bfvi_chan_scratch(channels)
    buf = torch.empty(n // 8 + 2, dtype=torch.float64, device=dev)
    return buf, n


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


def check_conv(case, lib, device, tol, sigmoid=False):
    """forward, input gradient, weight gradient (accumulating) and bias gradient of one layer."""
    dev = torch.device(device)
    g, tr, w, bias, small, big = make_conv(case)
    st = stream_of(dev)
    wd, bd = w.to(dev), bias.to(dev)
    w64 = w.double().requires_grad_(True)
    b64 = bias.double().requires_grad_(True)
    if not tr:
        x64 = big.double().requires_grad_(True)
        y64 = F.conv2d(x64, w64, b64, stride=g.stride, padding=g.padding)
        dy = small                                       # reuse the random small map as the output gradient
        y64.backward(dy.double())
        xd, dyd = big.to(dev), dy.to(dev)
        y = torch.full(tuple(y64.shape), float('nan'), device=dev)
        lib.call('bfvi_conv_gather', C.byref(g), _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(y), _lib.ACT_NONE, st)
        dx = torch.full_like(xd, float('nan'))
        lib.call('bfvi_conv_scatter', C.byref(g), _lib.ptr(dyd), _lib.ptr(wd), None, _lib.ptr(dx), _lib.ACT_NONE, st)
        dw = torch.ones_like(wd)                         # += semantics: starts at one
        lib.call('bfvi_conv_wgrad', C.byref(g), _lib.ptr(dyd), _lib.ptr(xd), _lib.ptr(dw), st)
    else:
        x64 = small.double().requires_grad_(True)
        y64 = F.conv_transpose2d(x64, w64, b64, stride=g.stride, padding=g.padding)
        if sigmoid:
            y64 = torch.sigmoid(y64)
        dy = big
        y64.backward(dy.double())
        xd, dyd = small.to(dev), dy.to(dev)
        y = torch.full(tuple(y64.shape), float('nan'), device=dev)
        lib.call('bfvi_conv_scatter', C.byref(g), _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(y),
                 _lib.ACT_SIGMOID if sigmoid else _lib.ACT_NONE, st)
        if sigmoid:
            pre = torch.full_like(dyd, float('nan'))
            lib.call('bfvi_sigmoid_bwd', _lib.ptr(y), _lib.ptr(dyd), dyd.numel(), _lib.ptr(pre), st)
            dyd = pre
        dx = torch.full_like(xd, float('nan'))
        lib.call('bfvi_conv_gather', C.byref(g), _lib.ptr(dyd), _lib.ptr(wd), None, _lib.ptr(dx), _lib.ACT_NONE, st)
        dw = torch.ones_like(wd)
        lib.call('bfvi_conv_wgrad', C.byref(g), _lib.ptr(xd), _lib.ptr(dyd), _lib.ptr(dw), st)
    db = torch.ones_like(bd)
    sc, n = scratch_for(lib, dyd.shape[1], dev)
    lib.call('bfvi_chan_bias_grad', _lib.ptr(dyd), dyd.shape[0], dyd.shape[1], dyd.shape[2] * dyd.shape[3],
             _lib.ptr(db), _lib.ptr(sc), C.c_size_t(n), st)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    errs = {'y': rel(y, y64.detach()), 'dx': rel(dx, x64.grad), 'dw': rel(dw - 1, w64.grad), 'db': rel(db - 1, b64.grad)}
    assert all(e < tol for e in errs.values()), (case, errs)
    return errs


def check_bn(case, lib, device, tol):
    N, Cc, H, W, training, relu = BN_CASES[case]
    dev = torch.device(device)
    gen = torch.Generator().manual_seed(1)
    r = lambda *sh: torch.randn(*sh, generator=gen)
    x = 1.5 * r(N, Cc, H, W) + 0.7 * r(1, Cc, 1, 1)
    gamma, beta = 1 + 0.2 * r(Cc), 0.3 * r(Cc)
    rm, rv = 0.1 * r(Cc), 1 + 0.3 * torch.rand(Cc, generator=gen)
    dy = r(N, Cc, H, W)
    momentum, eps = 0.1, 1e-5
    # fp64 restatement
    x64 = x.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm64, rv64 = rm.double().clone(), rv.double().clone()
    y64 = F.batch_norm(x64, rm64, rv64, g64, b64, training, momentum, eps)
    if relu:
        y64 = torch.relu(y64)
    y64.backward(dy.double())
    # kernels
    st = stream_of(dev)
    xd, gd, bd, rmd, rvd, dyd = (t.to(dev).contiguous() for t in (x, gamma, beta, rm, rv, dy))
    y = torch.full_like(xd, float('nan'))
    save = torch.full((Cc, 2), float('nan'), device=dev)
    sc, n = scratch_for(lib, Cc, dev)
    lib.call('bfvi_bn2d_fwd', _lib.ptr(xd), N, Cc, H * W, _lib.ptr(gd), _lib.ptr(bd), _lib.ptr(rmd), _lib.ptr(rvd),
             int(training), momentum, eps, int(relu), _lib.ptr(y), _lib.ptr(save), _lib.ptr(sc), C.c_size_t(n), st)
    dx = torch.full_like(xd, float('nan'))
    dg, dbeta = torch.ones_like(gd), torch.ones_like(bd)
    lib.call('bfvi_bn2d_bwd', _lib.ptr(dyd), _lib.ptr(xd), _lib.ptr(y), _lib.ptr(save), _lib.ptr(gd), N, Cc, H * W,
             int(training), int(relu), _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(dbeta), _lib.ptr(sc), C.c_size_t(n), st)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    errs = {'y': rel(y, y64.detach()), 'dx': rel(dx, x64.grad), 'dgamma': rel(dg - 1, g64.grad),
            'dbeta': rel(dbeta - 1, b64.grad), 'running_mean': rel(rmd, rm64), 'running_var': rel(rvd, rv64)}
    assert all(e < tol for e in errs.values()), (case, errs)
    return errs


def check_modules(common, device, tol, img_size=16, n_kernels=8, z_dim=12, frames=5, steps=2):
    """common.ImageEncoder -> common.ImageDecoder on `device` (this library's kernels) against a float64 copy of the
    same modules run by torch on the CPU: outputs, every parameter gradient, BatchNorm running statistics over
    `steps` training passes, then an evaluation-mode forward."""
    import copy
    torch.manual_seed(3)
    enc = common.ImageEncoder(z_dim, img_size=img_size, n_kernels=n_kernels)
    dec = common.ImageDecoder(z_dim, img_size=img_size, n_kernels=n_kernels)
    enc64, dec64 = copy.deepcopy(enc).double(), copy.deepcopy(dec).double()
    enc, dec = enc.to(device), dec.to(device)
    gen = torch.Generator().manual_seed(5)
    worst = {}

    def note(k, a, b):
        worst[k] = max(worst.get(k, 0.0), rel(a.detach(), b.detach()))

    for step in range(steps):
        x = torch.rand(frames, 3, img_size, img_size, generator=gen)
        tgt = (torch.rand(frames, 3, img_size, img_size, generator=gen) > 0.5).float()
        wz = torch.randn(frames, z_dim, generator=gen)
        outs = []
        for e, d, cast in ((enc, dec, lambda t: t.to(device)), (enc64, dec64, lambda t: t.double())):
            for mod in (e, d):
                mod.train()
                mod.zero_grad()
            mean, std = e(cast(x))
            (probs,) = d(mean + 0.1 * std)
            loss = F.binary_cross_entropy(probs, cast(tgt), reduction='sum') + (mean * cast(wz)).sum() + (std ** 2).sum()
            loss.backward()
            outs.append((mean, std, probs, loss))
        for name, a, b in zip(('mean', 'std', 'probs', 'loss'), outs[0], outs[1]):
            note(name, a, b)
        for (k, p), (_, p64) in zip(list(enc.named_parameters()) + list(dec.named_parameters()),
                                    list(enc64.named_parameters()) + list(dec64.named_parameters())):
            assert p.grad is not None, k
            if p64.grad.norm().item() < 1e-9 * max(q.grad.norm().item() for q in enc64.parameters()):
                continue                      # a convolution bias in front of a BatchNorm: mathematically zero gradient
            note('grad ' + k, p.grad, p64.grad)
        for (k, b), (_, b64) in zip(list(enc.named_buffers()) + list(dec.named_buffers()),
                                    list(enc64.named_buffers()) + list(dec64.named_buffers())):
            note('buffer ' + k, b.double(), b64.double())
    for mod in (enc, dec, enc64, dec64):
        mod.eval()
    with torch.no_grad():
        x = torch.rand(frames, 3, img_size, img_size, generator=gen)
        mean, std = enc(x.to(device))
        m64, s64 = enc64(x.double())
        note('eval mean', mean, m64)
        note('eval probs', dec(mean)[0], dec64(m64)[0])
    bad = {k: v for k, v in worst.items() if not v < tol}
    assert not bad, bad
    return worst


def check_dense(case, lib, device, tol):
    rows, n_in, n_out, relu = DENSE_CASES[case]
    dev = torch.device(device)
    gen = torch.Generator().manual_seed(2)
    r = lambda *sh: torch.randn(*sh, generator=gen)
    x, w, b, dy = r(rows, n_in), r(n_out, n_in) / n_in ** 0.5, 0.2 * r(n_out), r(rows, n_out)
    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x, w, b))
    y64 = F.linear(x64, w64, b64)
    if relu:
        y64 = torch.relu(y64)
    y64.backward(dy.double())
    st = stream_of(dev)
    xd, wd, bd, dyd = (t.to(dev).contiguous() for t in (x, w, b, dy))
    y = torch.full((rows, n_out), float('nan'), device=dev)
    lib.call('bfvi_dense_fwd', _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(y), rows, n_in, n_out, int(relu), st)
    dx = torch.full_like(xd, float('nan'))
    dw, db = torch.ones_like(wd), torch.ones_like(bd)          # += semantics
    masked = torch.empty_like(dyd) if relu else None
    sc, n = scratch_for(lib, n_out, dev)
    lib.call('bfvi_dense_bwd', _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(y), _lib.ptr(dyd), _lib.ptr(masked), rows, n_in, n_out,
             int(relu), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(sc), C.c_size_t(n), st)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    errs = {'y': rel(y, y64.detach()), 'dx': rel(dx, x64.grad), 'dw': rel(dw - 1, w64.grad), 'db': rel(db - 1, b64.grad)}
    assert all(e < tol for e in errs.values()), (case, errs)
    return errs


def check_golden_image(common, device, tol, check_init=False):
    """common.ImageEncoder / ImageDecoder on `device` against the fixture oracle/make_golden_image.py wrote from the
    UNMODIFIED reference's modules in float64 (tests/golden/image/modules.pt): two training passes (outputs, parameter
    gradients, BatchNorm buffers) and the evaluation-mode pass, from the reference's own initial weights."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'oracle'))
    import make_golden_image as gi
    fx = torch.load(os.path.join(root, 'tests', 'golden', 'image', 'modules.pt'), weights_only=False)
    cfg = fx['cfg']
    enc, dec = gi.build(common, cfg)
    if check_init:                           # same constructor order -> the reference's seeded initial values, bit for bit
        for mod, key in ((enc, 'enc'), (dec, 'dec')):
            sd = mod.state_dict()
            assert list(sd.keys()) == list(fx['init'][key].keys())
            assert all(torch.equal(sd[k], v) for k, v in fx['init'][key].items()), key
    enc.load_state_dict(fx['init']['enc'])
    dec.load_state_dict(fx['init']['dec'])
    enc, dec = enc.to(device), dec.to(device)
    got = gi.run(enc, dec, cfg, lambda t: t.to(device))
    worst, floor = {}, None
    for step, (g, r) in enumerate(zip(got['steps'], fx['ref']['steps'])):
        assert set(g) == set(r)
        floor = 1e-9 * max(v.norm().item() for k, v in r.items() if k.startswith('grad '))
        for k in r:
            if k.startswith('grad ') and r[k].norm().item() < floor:
                continue                      # convolution bias in front of a BatchNorm: zero gradient up to rounding
            worst['%d %s' % (step, k)] = rel(g[k], r[k])
    for k, v in fx['ref']['eval'].items():
        worst['eval ' + k] = rel(got['eval'][k], v)
    bad = {k: v for k, v in worst.items() if not v < tol}
    assert not bad, bad
    return worst
