"""Device-side timing of bfvi_forward (inference, config C5 at C3 dims): launch sequence vs fused transitions."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import multimodal_dmm_b200.models as models  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=1024)
    ap.add_argument('--T', type=int, default=100)
    ap.add_argument('--K', type=int, default=25)
    ap.add_argument('--steps', type=int, default=3)
    a = ap.parse_args()
    mods, dims = ['m%d' % i for i in range(8)], [16] * 8
    torch.manual_seed(1)
    m = models.MultiDMM(mods, dims, h_dim=512, z_dim=64, device=torch.device('cuda:0')).eval()
    g = torch.Generator().manual_seed(1)
    x = {k: torch.randn(a.T, a.B, 16, generator=g).cuda() for k in mods}
    lengths = [a.T] * a.B
    for prec in ('tf32x3', 'fused'):
        with torch.no_grad():
            out = m(x, lengths=lengths, mode='fsmooth', sample=False, flt_particles=a.K, precision=prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                out = m(x, lengths=lengths, mode='fsmooth', sample=False, flt_particles=a.K, precision=prec)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print('forward fsmooth C3 dims B=%d T=%d K=%d precision=%s: %.1f ms  %.3e seq-ts/s  infer_mean[0,0,:2]=%s' %
              (a.B, a.T, a.K, prec, ms, a.B * a.T / ms * 1e3, out[0][0][0, 0, :2].tolist()))


if __name__ == '__main__':
    main()
