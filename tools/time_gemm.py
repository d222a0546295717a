"""Device-side timing of the tcgen05 tile kernel on the GEMM shapes of one C3-dims time step
(rows = 9 chain sets x 256 sequences x 25 particles = 57 600, and the 2 304-row single-particle
passes), through the public bfvi_linear_tf32 / bfvi_wgrad_tf32 entry points (3xTF32).
BFVI_GEMM_V1=1 selects the round-1 kernel for comparison."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_dmm_b200 import _lib  # noqa: E402


ONCE = False


def timed(fn, iters=20):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return float('nan')
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def main():
    global ONCE
    ap = argparse.ArgumentParser()
    ap.add_argument('--single', action='store_true', help='single-pass TF32 instead of 3xTF32')
    ap.add_argument('--once', action='store_true', help='one launch per shape (for ncu)')
    ap.add_argument('--rows', type=int, nargs='*', default=[57600, 2304])
    ap.add_argument('--no-wgrad', action='store_true')
    a = ap.parse_args()
    ONCE = a.once
    fl = 16 if a.single else 0
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    tag = ('v1' if os.environ.get('BFVI_GEMM_V1') else 'v2') + ('-1xtf32' if a.single else '')
    for rows in a.rows:
        for n_in, n_out, act in ((64, 512, 1), (512, 64, 0), (64, 64, 0)):
            x = torch.randn(rows, n_in, device='cuda')
            w = torch.randn(n_out, n_in, device='cuda')
            b = torch.randn(n_out, device='cuda')
            y = torch.empty(rows, n_out, device='cuda')
            us = timed(lambda: lib.call('bfvi_linear_tf32', _lib.ptr(x), n_in, _lib.ptr(w), n_in, _lib.ptr(b),
                                        _lib.ptr(y), n_out, rows, n_in, n_out, act | fl, st))
            ref = x.double() @ w.double().t() + b.double()
            if act:
                ref = torch.relu(ref)
            err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
            byts = 4.0 * rows * (n_in + n_out)
            print(json.dumps({'kernel': tag, 'op': 'linear', 'rows': rows, 'in': n_in, 'out': n_out, 'us': round(us, 1),
                              'GBps': round(byts / us / 1e3, 1), 'TFLOPs_alg': round(2.0 * rows * n_in * n_out / us / 1e6, 1),
                              'rel_err': err}))
        for n_out, n_in in (() if a.no_wgrad else ((512, 64), (64, 512), (64, 64))):
            dyt = torch.randn(n_out, rows, device='cuda')
            xt = torch.randn(n_in, rows, device='cuda')
            dw = torch.zeros(n_out, n_in, device='cuda')

            def wg():
                lib.call('bfvi_wgrad_tf32', _lib.ptr(dyt), rows, _lib.ptr(xt), rows, _lib.ptr(dw), n_in, rows, n_out,
                         n_in, 1, fl, st)
            us = timed(wg)
            dw.zero_()
            wg()
            ref = dyt.double() @ xt.double().t()
            err = (dw.double() - ref).abs().max().item() / ref.abs().max().item()
            byts = 4.0 * rows * (n_in + n_out)
            print(json.dumps({'kernel': tag, 'op': 'wgrad', 'rows': rows, 'in': n_in, 'out': n_out, 'us': round(us, 1),
                              'GBps': round(byts / us / 1e3, 1), 'TFLOPs_alg': round(2.0 * rows * n_in * n_out / us / 1e6, 1),
                              'rel_err': err}))


if __name__ == '__main__':
    main()
