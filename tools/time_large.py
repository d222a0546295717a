"""Device-side timing of the large-dim (tcgen05) family: bfvi_step_fwd_bwd at C3 dims
(M=8, D=16, Z=64, H=512, K=25) on synthetic data, in-kernel Philox noise."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import bfvi_oracle as bo      # noqa: E402
import helpers                # noqa: E402
from multimodal_dmm_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=64)
    ap.add_argument('--T', type=int, default=100)
    ap.add_argument('--K', type=int, default=25)
    ap.add_argument('--M', type=int, default=8)
    ap.add_argument('--Z', type=int, default=64)
    ap.add_argument('--H', type=int, default=512)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--fwd-only', action='store_true')
    ap.add_argument('--graph', action='store_true', help='capture the step in a CUDA graph and time replays')
    ap.add_argument('--precision', type=int, default=2, help='0 tf32x3 launch sequence, 1 tf32, 2 fused on-chip kernels')
    ap.add_argument('--batch-tile', type=int, default=0)
    ap.add_argument('--f-mult', type=float, default=0.5, help='0 skips pass A (the f_mode filter)')
    ap.add_argument('--s-mult', type=float, default=0.5, help='0 skips passes B and C (the s_mode smoother)')
    ap.add_argument('--match-mult', type=float, default=0.01)
    a = ap.parse_args()
    lib = _lib.load()
    mods, dims = ['m%d' % i for i in range(a.M)], [16] * a.M
    g = torch.Generator().manual_seed(1)
    x = {m: torch.randn(a.T, a.B, 16, generator=g) for m in mods}
    fx = dict(modalities=mods, dims=dims, z_dim=a.Z, h_dim=a.H, min_std=1e-3, inputs=x, targets=x,
              mask=torch.ones(a.T, a.B, 1, dtype=torch.bool), lengths=[a.T] * a.B, kld_mult=1.0,
              rec_mults={m: 1.0 / (16 * a.M) for m in mods}, step_kwargs={'train_particles': a.K},
              state_dict=bo.init_params(mods, dims, h_dim=a.H, z_dim=a.Z, seed=1))
    model, dists = helpers.fixture_model(fx)
    flat, lay = helpers.pack_params(lib, model, mods, dists, fx['state_dict'], 'cuda')
    args, keep = helpers.step_args(fx, 'cuda', None, seed=2024, kwargs={'precision': a.precision, 'batch_tile': a.batch_tile, 'f_mult': a.f_mult,
                                                                    's_mult': a.s_mult, 'match_mult': a.match_mult})
    nbytes = C.c_size_t(0)
    lib.call('bfvi_step_workspace', C.byref(model), C.byref(args), C.byref(nbytes))
    ws = helpers.aligned_empty(nbytes.value, 'cuda')
    grads = None if a.fwd_only else torch.zeros(lay.total, device='cuda')
    loss = torch.zeros(1, device='cuda')
    launches = C.c_int32(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        nonlocal st
        lib.call('bfvi_step_fwd_bwd', C.byref(model), _lib.ptr(flat), _lib.ptr(grads), C.byref(args),
                 _lib.ptr(ws), C.c_size_t(nbytes.value), _lib.ptr(loss), C.byref(launches), st)
    step()
    torch.cuda.synchronize()
    if a.graph:                                   # the step is stream-ordered (side stream forked / joined by events)
        cap = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            st = C.c_void_p(cap.cuda_stream)
            step()                                # warm-up on the capture stream
            cap.synchronize()
            with torch.cuda.graph(graph, stream=cap):
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                step()
        eager_loss = loss.item()
        run = graph.replay
        run()
        torch.cuda.synchronize()
        print('graph replay loss %.3f (eager %.3f)' % (loss.item(), eager_loss))
    else:
        run = step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    flop = 216023040.0 * a.B * a.T if (a.M, a.Z, a.H, a.K) == (8, 64, 512, 25) else float('nan')
    print('dispatch:', ';'.join(lib.last_dispatch())[:400])
    print('C3 dims B=%d T=%d K=%d: %.1f ms/step  %.3e seq-ts/s  %.1f TFLOP/s algorithmic  loss=%.3f  launches=%d  ws=%.0f MB'
          % (a.B, a.T, a.K, ms, a.B * a.T / ms * 1e3, flop / ms / 1e9, loss.item(), launches.value, nbytes.value / 1e6))


if __name__ == '__main__':
    main()
