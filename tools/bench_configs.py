"""Throughput of the BASELINE.json configurations that are NOT the bench.py headline
(bench.py times configs[1] "C2"): the scaled model "C3" (M=8, D=16, Z=64, H=512) through
MultiDMM.step + backward, and the inference sweep "C5" through MultiDMM.forward.  Prints one
JSON line per measurement with bench.py's keys (device-timed with CUDA events, warm-up first).

    python tools/bench_configs.py c3 [--batch 256] [--T 100]
    python tools/bench_configs.py c5 [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import multimodal_dmm_b200.models as models   # noqa: E402

METRIC, UNIT = 'bfvi_elbo_fwd_bwd_seq_timesteps_per_sec', 'seq-timesteps/s'


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def timed(fn, steps, warmup):
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


def c3(args):
    """Weak-scaling data parallelism under torchrun (one rank per GPU, B sequences per GPU, one NCCL
    all-reduce of the flat gradient per step): python -m torch.distributed.run --nproc-per-node N
    tools/bench_configs.py c3"""
    import bfvi_oracle as bo
    import torch.distributed as dist
    M, D, Z, H, K, KM = 8, 16, 64, 512, 25, 50
    mods, dims = ['m%d' % i for i in range(M)], [D] * M
    T, B = args.T, args.batch
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda:%d' % local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(1)
    model = models.MultiDMM(mods, dims, h_dim=H, z_dim=Z, device=dev).train()
    model.cuda_graph = not args.eager     # replay the captured step (seed read from device memory)
    if world > 1:
        model.b_offset = rank * B
        model.grad_sync = lambda flat: dist.all_reduce(flat)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = {m: torch.randn(T, B, D, device=dev, generator=g) for m in mods}
    mask = torch.ones(T, B, 1, dtype=torch.bool, device=dev)
    lengths = [T] * B
    rec = {m: 1.0 / (D * M) for m in mods}
    model.noise_seed = 2024

    def step():
        loss = model.step(x, mask, 1.0, rec, targets=x, lengths=lengths, train_particles=K, match_particles=KM)
        (loss / (T * B)).backward()
        for p in model.parameters():
            p.grad = None
    if world > 1:
        dist.barrier()
    ms = timed(step, args.steps, 1)
    if world > 1:                                    # max over ranks, whole-job throughput
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
        if rank != 0:
            dist.destroy_process_group()
            return
    B_total = B * world
    flop = 216023040.0 * T * B                       # SURVEY.md §8d algorithmic FLOP per seq-timestep (per GPU)
    pk = peaks()
    peak = float(pk.get('bf16_tflops_sustained', 1400.0))
    # CPU baseline: oracle port on a bounded sample of the same model
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bc, tc = 8, 20
    params = bo.init_params(mods, dims, h_dim=H, z_dim=Z, seed=1)
    for p in params.values():
        p.requires_grad_(True)
    orc = bo.OracleDMM(mods, dims, params, h_dim=H, z_dim=Z, draw=bo.RandomDraw(seed=3))
    xc = {m: torch.randn(tc, bc, D) for m in mods}
    t0 = time.perf_counter()
    loss = orc.step(xc, torch.ones(tc, bc, 1, dtype=torch.bool), 1.0, rec, targets=xc, lengths=[tc] * bc,
                    train_particles=K, match_particles=KM)
    loss.backward()
    cpu_s = time.perf_counter() - t0
    print(json.dumps({
        'metric': METRIC, 'value': T * B_total / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': 1,
        'scaling': 'weak',
        'ms_per_step': ms, 'higher_is_better': True, 'dtype': 'tf32x3 (fp32 accumulate)', 'data': 'synthetic',
        'config': {'workload': 'C3 dims: M=8 D=16 Z=64 H=512 K=25 K_match=50, REDUCED to T=%d, B=%d per GPU '
                               '(BASELINE: T=1000, B=65536): large-dim tcgen05 launch-sequence family, %s' % (
                                   T, B, 'eager launches' if args.eager else 'CUDA-graph replay of the step')},
        'gpu_launches': model.last_launches * args.steps,
        'roofline': {'bound': 'tensor', 'achieved': flop / (ms * 1e-3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': flop / (ms * 1e-3) / 1e12 / peak, 'traffic': None,
                     'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if pk else 'fallback'},
        'cpu_baseline': {'value': tc * bc / cpu_s, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'oracle port, one step + backward at B=%d, T=%d' % (bc, tc)}}))
    if world > 1:
        dist.destroy_process_group()


def c5(args):
    """Inference: forward(mode=fsmooth, sample=False, flt_particles=K) under no_grad on sequences with
    50 % random deletion and only the middle half kept (forecast + backcast, trainer.py:284-287)."""
    dev = torch.device('cuda:0')
    rows = []
    grid = [(100, 10000), (1000, 1000), (10000, 100)]
    ks = [1, 25, 200]
    if args.quick:
        grid, ks = [(100, 2000)], [1, 25]
    torch.manual_seed(1)
    spir = models.MultiDMM(['spiral-x', 'spiral-y'], [1, 1], h_dim=20, z_dim=5, device=dev).eval()
    big = models.MultiDMM(['m%d' % i for i in range(8)], [16] * 8, h_dim=512, z_dim=64, device=dev).eval()
    rng = np.random.RandomState(1)

    def make(model, T, B):
        out = {}
        for m in model.modalities:
            d = int(np.prod(model.dims[m]))
            x = torch.randn(T, B, d)
            x[torch.from_numpy(rng.rand(T, B) < 0.5)] = float('nan')
            x[:T // 4] = float('nan')
            x[3 * T // 4:] = float('nan')
            out[m] = x.to(dev)
        return out
    for name, model, combos in (('spirals dims (Z=5,H=20), small-dim family', spir, [(t, b, k) for t, b in grid for k in ks]),
                                ('C3 dims (Z=64,H=512), tcgen05 family', big,
                                 [(100, 256, 25)] if args.quick else [(100, 1024, 25), (1000, 128, 25), (100, 1024, 1)])):
        for T, B, K in combos:
            x = make(model, T, B)
            lengths = [T] * B

            def fwd():
                with torch.no_grad():
                    model(x, lengths=lengths, mode='fsmooth', sample=False, flt_particles=K)
            ms = timed(fwd, 3, 1)
            rows.append({'metric': 'bfvi_forward_seq_timesteps_per_sec', 'value': T * B / (ms * 1e-3), 'unit': UNIT,
                         'ms_per_call': ms, 'config': {'workload': 'C5 inference, %s: T=%d, B=%d, flt_particles=%d' % (name, T, B, K)}})
            print(json.dumps(rows[-1]))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('which', choices=['c3', 'c5'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--T', type=int, default=100)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--eager', action='store_true', help='c3: launch the step eagerly instead of replaying its CUDA graph')
    a = ap.parse_args()
    (c3 if a.which == 'c3' else c5)(a)
