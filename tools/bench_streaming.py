"""HBM roofline of the streaming kernels either side of the step (SURVEY §8f-2 batch
preparation, §8a17 Bernoulli likelihood): algorithmic bytes / CUDA-event time against the
measured copy bandwidth of MEASURED_PEAKS.json, on buffers larger than the 126 MB L2; plus the
batch-preparation call at the C2 shape next to the reference's per-sequence loop restated with
torch device ops (what `mseq.burst_delete(targets, ...)` does on a CUDA batch, trainer.py:235).

    python tools/bench_streaming.py            # one JSON line per kernel
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_dmm_b200 import _lib, multiseq            # noqa: E402
from multimodal_dmm_b200.models import losses              # noqa: E402


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    lib = _lib.load()
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        src = 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        peak, src = 6550.0, 'fallback (B200_PROFILING.md)'
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    T, B, D = 64, 512, 3 * 64 * 64                       # 403 M elements = 1.6 GB per buffer (>> L2)
    n = T * B * D
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.rand(T, B, D, device='cuda', generator=g)
    th = torch.rand(T, B, D, device='cuda', generator=g) * 0.98 + 0.01
    out = torch.empty_like(x)
    flags = (torch.rand(T, B, device='cuda', generator=g) < 0.3).to(torch.uint8)
    rmask = (torch.rand(T * B, device='cuda', generator=g) < 0.9).to(torch.uint8)
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    starts = torch.arange(0, (B + 1) * T, T, dtype=torch.int64, device='cuda')
    rows = []

    def report(name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({'kernel': name, 'ms': ms, 'elements': n, 'algorithmic_bytes': nbytes,
                     'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
                                  'peak_source': src}, 'note': note})
        print(json.dumps(rows[-1]))

    ms = timed(lambda: lib.call('bfvi_delete_rows', _lib.ptr(x), _lib.ptr(flags), T, B, D, _lib.ptr(out), st))
    report('delete_rows_kernel<vec>', ms, 8 * n, 'func_delete: read 4 B + write 4 B per element')
    ms = timed(lambda: lib.call('bfvi_pad_merge', _lib.ptr(x), _lib.ptr(starts), T, B, D, _lib.ptr(out), st))
    report('pad_merge_kernel<vec>', ms, 8 * n, 'pad_and_merge: gather rows, read 4 B + write 4 B per element')
    # seq_decoll: de-pad (ragged lengths: 3/4 of the rows kept on average) + reorder, one gather pass
    lens = torch.randint(T // 2, T + 1, (B,), generator=torch.Generator().manual_seed(1))
    order = torch.randperm(B, generator=torch.Generator().manual_seed(2))
    ustarts = torch.zeros(B + 1, dtype=torch.int64)
    ustarts[1:] = torch.cumsum(lens[order], 0)
    total = int(ustarts[-1])
    ustarts_d, order_d = ustarts.cuda(), order.to(torch.int32).cuda()
    ms = timed(lambda: lib.call('bfvi_unpad', _lib.ptr(x), _lib.ptr(ustarts_d), _lib.ptr(order_d), B, B, total, D,
                                _lib.ptr(out), st))
    rows_kept = total
    gbs_n = 8 * rows_kept * D
    report('unpad_kernel<vec>', ms, gbs_n, 'seq_decoll: read 4 B + write 4 B per KEPT element (%d of %d rows)' % (rows_kept, T * B))
    # per-sequence MSE metric: read reconstruction + target
    lens_f = lens.float().cuda()
    mk = (torch.arange(T).view(T, 1) < lens.view(1, B)).to(torch.uint8).cuda()
    mse_out = torch.empty(B, device='cuda')
    rp, tp, dd = (C.c_void_p * 1)(x.data_ptr()), (C.c_void_p * 1)(th.data_ptr()), (C.c_int64 * 1)(D)
    n_split = int(lib.dll.bfvi_seq_mse_splits(T, B))
    scratch = torch.empty(B * n_split, device='cuda')
    ms = timed(lambda: lib.call('bfvi_seq_mse', rp, tp, dd, 1, _lib.ptr(mk), _lib.ptr(lens_f), T, B, _lib.ptr(mse_out),
                                _lib.ptr(scratch), n_split, st))
    report('seq_mse_kernel', ms, 8 * rows_kept * D, 'per-sequence MSE: read reconstruction + target on the unmasked rows')
    ms = timed(lambda: lib.call('bfvi_nll_bernoulli_fwd', _lib.ptr(th), _lib.ptr(x), _lib.ptr(rmask), T * B, D,
                                _lib.ptr(acc), st))
    report('nll_bernoulli_kernel<vec> fwd', ms, 8 * n, 'BCE sum: read theta + x')
    ms = timed(lambda: lib.call('bfvi_nll_bernoulli_bwd', _lib.ptr(th), _lib.ptr(x), _lib.ptr(rmask), T * B, D,
                                C.c_float(1.0), _lib.ptr(out), st))
    report('nll_bernoulli_kernel<vec> bwd', ms, 12 * n, 'BCE gradient: read theta + x, write d_theta')
    # torch's own masked-select formulation of the same loss (what the reference runs on a GPU)
    keep = (~torch.isnan(x)) & rmask.bool().view(T, B, 1)
    ms_t = timed(lambda: torch.nn.functional.binary_cross_entropy(th.masked_select(keep), x.masked_select(keep),
                                                                    reduction='sum'), steps=5, warmup=2)
    print(json.dumps({'kernel': 'torch masked_select + binary_cross_entropy (reference formulation, same GPU)',
                      'ms': ms_t, 'elements': n}))
    del x, th, out, keep
    torch.cuda.empty_cache()

    # fused SSIM (utils.py:110-212) on 8 192 RGB 64 x 64 frames vs the reference formulation with torch ops
    from multimodal_dmm_b200 import metrics
    import torch.nn.functional as F
    xi = torch.rand(8192, 3, 64, 64, device='cuda', generator=g)
    yi = (xi + 0.05 * torch.randn(xi.shape, device='cuda', generator=g)).clamp(0, 1)
    ms = timed(lambda: metrics.eval_ssim(xi, yi), steps=10, warmup=2)
    npx = xi.numel()
    win = metrics._fspecial_gauss_1d(11, 1.5).cuda()

    def ref_ssim():                                        # utils.py:93-163 as the reference runs it on a GPU
        c = xi.shape[1]
        w = win.repeat(5 * c, 1, 1, 1)
        z = torch.cat([xi, yi, xi * xi, yi * yi, xi * yi], 1)
        z = F.conv2d(z, w, groups=5 * c).transpose(2, 3).contiguous()
        z = F.conv2d(z, w, groups=5 * c).transpose(2, 3).contiguous()
        mu1, mu2, s1, s2, s12 = (z[:, i * c:(i + 1) * c] for i in range(5))
        s1, s2, s12 = s1 - mu1.pow(2), s2 - mu2.pow(2), s12 - mu1 * mu2
        cs = (2 * s12 + 9e-4) / (s1 + s2 + 9e-4)
        return (((2 * mu1 * mu2 + 1e-4) / (mu1.pow(2) + mu2.pow(2) + 1e-4)) * cs).mean(-1).mean(-1).mean(-1)
    ms_ref = timed(ref_ssim, steps=5, warmup=2)
    gbs = 8 * npx / (ms * 1e-3) / 1e9
    print(json.dumps({'kernel': 'ssim_kernel (fused eval_ssim)', 'ms': ms, 'elements': npx, 'algorithmic_bytes': 8 * npx,
                      'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
                                   'peak_source': src},
                      'note': 'reads X + Y once (8 B per pixel); 550 FMA per output pixel from shared memory: compute / '
                              'shared-memory bound, not HBM bound', 'torch_reference_formulation_ms_same_gpu': ms_ref}))
    del xi, yi
    torch.cuda.empty_cache()

    # batch preparation at the C2 shape: ours (device draw / numpy replay) vs the reference's loop
    T, B = 100, 4096
    batch = {m: torch.randn(T, B, 1, device='cuda', generator=g) for m in ('spiral-x', 'spiral-y')}
    lengths = [T] * B

    def ref_loop():                                        # datasets/multiseq.py:405-434 on a CUDA batch
        outb = {}
        for m in batch:
            outb[m] = batch[m].clone().detach()
            for b in range(B):
                t0 = np.random.randint(lengths[b])
                outb[m][list(range(t0, min(t0 + int(0.1 * lengths[b]), lengths[b]))), b] = float('nan')
        return outb

    def wall(fn, reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps * 1e3
    wall(lambda: multiseq.burst_delete(batch, 0.1, lengths, seed=1), 3)
    prep = {'workload': 'burst_delete(0.1) on the C2 batch: 2 modalities x (100, 4096, 1)', 'unit': 'ms per batch (wall)',
            'ours_device_draw': wall(lambda: multiseq.burst_delete(batch, 0.1, lengths, seed=1), 20),
            'ours_numpy_replay': wall(lambda: multiseq.burst_delete(batch, 0.1, lengths), 20),
            'ours_rand_delete_device_draw': wall(lambda: multiseq.rand_delete(batch, 0.5, lengths, seed=1), 20),
            'ours_rand_delete_numpy_replay': wall(lambda: multiseq.rand_delete(batch, 0.5, lengths), 3),
            'reference_loop_same_gpu': wall(ref_loop, 2)}
    print(json.dumps(prep))


if __name__ == '__main__':
    main()
