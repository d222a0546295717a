"""Quick device-side timing of bfvi_step_fwd_bwd on synthetic spirals-shaped data with
the per-phase breakdown of bfvi_step_profile (development aid; bench.py is the
contract).  BFVI_LIB_PATH selects a tuning variant of the library."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import bfvi_oracle as bo      # noqa: E402
import helpers                # noqa: E402
from multimodal_dmm_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=4096)
    ap.add_argument('--T', type=int, default=100)
    ap.add_argument('--K', type=int, default=25)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--fwd-only', action='store_true')
    ap.add_argument('--tag', default='')
    a = ap.parse_args()
    lib = _lib.load()
    mods, dims, Z, H = ['spiral-x', 'spiral-y'], [1, 1], 5, 20
    g = torch.Generator().manual_seed(1)
    x = {m: torch.randn(a.T, a.B, 1, generator=g) for m in mods}
    inp = {m: v.clone() for m, v in x.items()}
    for m in mods:
        drop = torch.rand(a.T, a.B, generator=g) < 0.5
        inp[m][drop] = float('nan')
    fx = dict(modalities=mods, dims=dims, z_dim=Z, h_dim=H, min_std=1e-3,
              inputs=inp, targets=inp, mask=torch.ones(a.T, a.B, 1, dtype=torch.bool),
              lengths=[a.T] * a.B, kld_mult=1.0, rec_mults={m: 1.0 for m in mods},
              step_kwargs={'train_particles': a.K}, state_dict=bo.init_params(mods, dims, h_dim=H, z_dim=Z, seed=1))
    model, dists = helpers.fixture_model(fx)
    flat, lay = helpers.pack_params(lib, model, mods, dists, fx['state_dict'], 'cuda')
    args, keep = helpers.step_args(fx, 'cuda', None, seed=2024)
    nbytes = C.c_size_t(0)
    lib.call('bfvi_step_workspace', C.byref(model), C.byref(args), C.byref(nbytes))
    ws = helpers.aligned_empty(nbytes.value, 'cuda')
    grads = None if a.fwd_only else torch.zeros(lay.total, device='cuda')
    loss = torch.zeros(1, device='cuda')
    launches = C.c_int32(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        lib.call('bfvi_step_fwd_bwd', C.byref(model), _lib.ptr(flat), _lib.ptr(grads), C.byref(args),
                 _lib.ptr(ws), C.c_size_t(nbytes.value), _lib.ptr(loss), C.byref(launches), st)
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    phase_ms = (C.c_float * len(_lib.PHASES))()
    acc = [0.0] * len(_lib.PHASES)
    for _ in range(3):
        lib.call('bfvi_step_profile', C.byref(model), _lib.ptr(flat), _lib.ptr(grads), C.byref(args),
                 _lib.ptr(ws), C.c_size_t(nbytes.value), _lib.ptr(loss), phase_ms, st)
        acc = [x + y / 3 for x, y in zip(acc, phase_ms)]
    print('%s B=%d T=%d K=%d  %.3f ms/step  %.3e seq-ts/s  loss=%.4f  launches=%d  ws=%.0f MB'
          % (a.tag, a.B, a.T, a.K, ms, a.B * a.T / ms * 1e3, loss.item(), launches.value, nbytes.value / 1e6))
    print('   ' + '  '.join('%s=%.3f' % (n, v) for n, v in zip(_lib.PHASES, acc) if v > 0.0005))


if __name__ == '__main__':
    main()
