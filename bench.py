"""bench.py — BFVI ELBO fwd+bwd sequence-timesteps/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2]

One "step" = MultiDMM.step(...) + (loss / sum(lengths)).backward() exactly as
trainer.py:237-243 drives it, on one batch of synthetic spirals-shaped data.

Workload at N GPUs (weak scaling: per-GPU batch fixed): BASELINE.json configs[1]
"C2": spirals model (M=2, D=1, Z=5, H=20; spirals.py:44-51), T=100, B=4096 per GPU,
50 % uniformly missing timesteps per sequence and modality in inputs AND targets
(corrupt_(0.5,'uniform'), datasets/multiseq.py:252-267), + burst_delete(0.1) on the
inputs, rec_mults=1.0, kld_mult=1.0, train_particles=25, match_particles=50,
in-kernel Philox noise.  One NCCL all-reduce of the flat gradient per step for N>1.

Printed JSON (one line, rank 0): value = device-timed throughput with inputs
resident in HBM; e2e = the same through the public API from PINNED HOST buffers
(H2D of every step's inputs and a D2H read of every step's loss inside the timed region,
the loss of step i is read while step i+1 runs, so the GPU never waits for the host);
roofline / roofline_fp32 for the dominant kernel from CUDA-event phase timing;
cpu_baseline = the oracle port of the reference timed on this box's host cores.

--impl reference: the reference algorithm (oracle/bfvi_oracle.py, a PyTorch-CPU
port; the Python reference itself cannot travel to the GPU box) on all host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODS, DIMS, Z_DIM, H_DIM = ['spiral-x', 'spiral-y'], [1, 1], 5, 20
T_MAX, B_PER_GPU = 100, 4096
K_TRAIN, K_MATCH = 25, 50
METRIC = 'bfvi_elbo_fwd_bwd_seq_timesteps_per_sec'
UNIT = 'seq-timesteps/s'

# algorithmic work per sequence-timestep
F_GTF = 8 * Z_DIM * H_DIM + 4 * Z_DIM * Z_DIM                       # 900
N_SETS = 3                                                          # {x,y}, {x}, {y}


# ----------------------------------------------------------------------------
# synthetic C2 batch (shape and missingness of datasets/spirals.py + multiseq.py)
# ----------------------------------------------------------------------------
def spirals_batch(b_dim, t_max, seed):
    """Noisy 2-D spirals like datasets/spirals.py:47-84 (vectorised restatement)."""
    rng = np.random.RandomState(seed)
    direction = np.where(np.arange(b_dim) >= b_dim / 2, 1.0, -1.0)
    start_r = 0.25 + rng.rand(b_dim) * 0.5
    stop_r = 2.25 + rng.rand(b_dim) * 0.5
    start_th = direction * (rng.rand(b_dim) * np.pi)
    stop_th = direction * (rng.rand(b_dim) * np.pi + 4 * np.pi)
    ratio = 2.0 ** (2 * rng.rand(b_dim) - 1)
    lin = np.linspace(0.0, 1.0, t_max)[:, None]
    r = start_r + (stop_r - start_r) * lin
    th = start_th + (stop_th - start_th) * lin
    x = np.sqrt(ratio) * r * np.cos(th) + 0.1 * rng.randn(t_max, b_dim)
    y = r * np.sin(th) / np.sqrt(ratio) + 0.1 * rng.randn(t_max, b_dim)
    return {'spiral-x': x[:, :, None].astype(np.float32),
            'spiral-y': y[:, :, None].astype(np.float32)}, rng


def make_c2_batch(b_dim, t_max=T_MAX, seed=1):
    data, rng = spirals_batch(b_dim, t_max, seed)
    targets, inputs = {}, {}
    n_del, burst = int(0.5 * t_max), int(0.1 * t_max)
    for m in MODS:
        tgt = data[m].copy()
        # corrupt_(0.5, 'uniform'): exactly n_del timesteps per sequence, no replacement
        order = np.argsort(rng.rand(t_max, b_dim), axis=0)[:n_del]
        tgt[order, np.arange(b_dim)[None, :], 0] = np.nan
        inp = tgt.copy()
        # burst_delete(0.1): one burst per sequence per modality (multiseq.py:428-434)
        start = rng.randint(t_max, size=b_dim)
        tt = np.arange(t_max)[:, None]
        inp[(tt >= start[None, :]) & (tt < np.minimum(start + burst, t_max)[None, :]), 0] = np.nan
        targets[m], inputs[m] = torch.from_numpy(tgt), torch.from_numpy(inp)
    lengths = [t_max] * b_dim
    mask = torch.ones(t_max, b_dim, 1, dtype=torch.bool)
    return inputs, targets, mask, lengths


def workload_name(b_dim):
    return ('C2: spirals BFVI step, M=2 D=1 Z=5 H=20, T=%d, B=%d per GPU, 50%% uniform missing + 10%% burst, '
            'K=%d, K_match=%d' % (T_MAX, b_dim, K_TRAIN, K_MATCH))


REC_MULTS = {m: 1.0 for m in MODS}          # (1/D)/M * 1/(1-0.5), spirals.py:64-73
KLD_MULT = 1.0


# ----------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6
                          for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------
def oracle_step_time(b_dim, steps, warmup, threads):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import bfvi_oracle as bo
    torch.set_num_threads(threads)
    inputs, targets, mask, lengths = make_c2_batch(b_dim, seed=1)
    params = bo.init_params(MODS, DIMS, h_dim=H_DIM, z_dim=Z_DIM, seed=1)
    for p in params.values():
        p.requires_grad_(True)
    orc = bo.OracleDMM(MODS, DIMS, params, h_dim=H_DIM, z_dim=Z_DIM, draw=bo.RandomDraw(seed=3))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = orc.step(inputs, mask, KLD_MULT, REC_MULTS, targets=targets, lengths=lengths,
                        train_particles=K_TRAIN, match_particles=K_MATCH)
        (loss / sum(lengths)).backward()
        for p in params.values():
            p.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return float(np.mean(times)), b_dim * T_MAX


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    # bounded sample: ~12 s/step at B=4096 on 8 cores; keep the whole run to a few minutes
    b_dim = B_PER_GPU if total <= 6 else max(128, int(B_PER_GPU * 6 / total) // 128 * 128)
    sec, seq_ts = oracle_step_time(b_dim, args.steps, args.warmup, cores)
    value = seq_ts / sec
    sample = 'C2 workload at B=%d, T=%d (K=%d particles), %d timed steps' % (b_dim, T_MAX, K_TRAIN, args.steps)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32',
        'data': 'synthetic',
        'config': {'workload': workload_name(B_PER_GPU), 'global_batch': B_PER_GPU * args.gpus, 'seq_len': T_MAX,
                   'parallelism': 'cpu', 'reference_sample_batch': b_dim},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=B_PER_GPU, help='sequences per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import multimodal_dmm_b200.models as models
    from multimodal_dmm_b200 import _lib

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the BFVI path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    # ---- model + data (identical weights on every rank; per-rank data shard) ------
    torch.manual_seed(1)
    model = models.MultiDMM(MODS, DIMS, h_dim=H_DIM, z_dim=Z_DIM, device=dev).train()
    b_dim = args.batch
    inputs_h, targets_h, mask_h, lengths = make_c2_batch(b_dim, seed=1 + rank)
    pin = lambda d: {k: v.pin_memory() for k, v in d.items()}
    inputs_h, targets_h = pin(inputs_h), pin(targets_h)
    mask_d = mask_h.to(dev)
    inputs_d = {k: v.to(dev) for k, v in inputs_h.items()}
    targets_d = {k: v.to(dev) for k, v in targets_h.items()}
    n_global = float(sum(lengths) * world)          # normalise by the GLOBAL sum(lengths)
    model.b_offset = rank * b_dim
    if world > 1:
        model.grad_sync = lambda g: dist.all_reduce(g)      # one NCCL all-reduce of the flat grad
    model.noise_seed = None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def one_step(inp, tgt):
        loss = model.step(inp, mask_d, KLD_MULT, REC_MULTS, targets=tgt, lengths=lengths)
        (loss / n_global).backward()
        for p in model.parameters():
            p.grad = None
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------
    for _ in range(args.warmup):
        one_step(inputs_d, targets_d)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    launches = 0
    for i in range(args.steps):
        flush.fill_(float(i))                         # L2 flush between timed iterations
        ev[i][0].record()
        one_step(inputs_d, targets_d)
        ev[i][1].record()
        launches += model.last_launches
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps

    # ---- end-to-end timing: pinned host inputs -> H2D -> step -> D2H loss ----------
    h2d = sum(v.numel() * 4 for v in inputs_h.values()) + sum(v.numel() * 4 for v in targets_h.values())
    for _ in range(2):
        one_step({k: v.to(dev, non_blocking=True) for k, v in inputs_h.items()},
                 {k: v.to(dev, non_blocking=True) for k, v in targets_h.items()}).item()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # A training loop that does not stall the GPU on the loss: every step copies its batch from pinned host
    # memory (same stream, ahead of the step) and copies its loss to pinned host memory asynchronously; the
    # host reads step i's loss while step i+1 runs.  Every step still pays its H2D copy and its D2H loss read
    # inside the timed region.  (A per-step .item() leaves the GPU idle for ~0.65 ms of launch latency per
    # step; staging the next batch on a second stream was measured and gained nothing over this.)
    main_stream = torch.cuda.current_stream(dev)
    loss_pinned = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    loss_host = None
    for i in range(args.steps):
        j = i % 2
        inp = {k: v.to(dev, non_blocking=True) for k, v in inputs_h.items()}
        tgt = {k: v.to(dev, non_blocking=True) for k, v in targets_h.items()}
        loss = one_step(inp, tgt)
        loss_pinned[j].copy_(loss.detach(), non_blocking=True)
        loss_ready[j].record(main_stream)
        if i > 0:                                      # D2H read of the PREVIOUS step's result
            loss_ready[1 - j].synchronize()
            loss_host = float(loss_pinned[1 - j])
    loss_ready[(args.steps - 1) % 2].synchronize()
    loss_host = float(loss_pinned[(args.steps - 1) % 2])
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps

    if os.environ.get('BFVI_BENCH_PROBE'):             # development aid: where does the e2e loop lose time?
        def loop(fn, n):
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n):
                fn()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / n, (time.perf_counter() - w0) * 1e3 / n
        print('probe device inputs, no flush, no sync   (gpu ms, wall ms):', loop(lambda: one_step(inputs_d, targets_d), args.steps), file=sys.stderr)
        print('probe device inputs + .item() each step  (gpu ms, wall ms):', loop(lambda: one_step(inputs_d, targets_d).item(), args.steps), file=sys.stderr)
        hs = time.perf_counter()
        for _ in range(args.steps):
            model.step(inputs_d, mask_d, KLD_MULT, REC_MULTS, targets=targets_d, lengths=lengths)
        print('probe host time of step() enqueue only (ms):', (time.perf_counter() - hs) * 1e3 / args.steps, file=sys.stderr)
        torch.cuda.synchronize()

    # ---- max over ranks ------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t[0].item(), t[1].item()
    seq_ts_global = b_dim * T_MAX * world

    # ---- dominant-kernel roofline from CUDA-event phase timing (rank 0) -------------
    roofline = roofline_fp32 = phases = None
    if rank == 0:
        lib = _lib.load()
        model._ensure_flat()
        fx_args, keep = build_step_args(model, inputs_d, targets_d, mask_d)
        nbytes = C.c_size_t(0)
        lib.call('bfvi_step_workspace', C.byref(model._cmodel), C.byref(fx_args), C.byref(nbytes))
        ws = model._workspace(nbytes.value)
        grads = torch.empty_like(model._flat)
        loss = torch.empty((), device=dev)
        acc = np.zeros(len(_lib.PHASES))
        reps = max(3, min(args.steps, 10))
        phase_ms = (C.c_float * len(_lib.PHASES))()
        for i in range(reps + 1):
            flush.fill_(float(i))
            lib.call('bfvi_step_profile', C.byref(model._cmodel), _lib.ptr(model._flat), _lib.ptr(grads),
                     C.byref(fx_args), _lib.ptr(ws), C.c_size_t(nbytes.value), _lib.ptr(loss), phase_ms,
                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
            if i > 0:
                acc += np.array(list(phase_ms))
        acc /= reps
        phases = {n: round(float(v), 4) for n, v in zip(_lib.PHASES, acc)}
        dom = max(phases, key=phases.get)
        chain_steps = N_SETS * b_dim * (T_MAX - 1)
        # algorithmic work of the dominant kernel per launch (recomputation NOT counted)
        flops = {'filter_s_flt_bwd': 2 * K_TRAIN * F_GTF, 'filter_s_flt_fwd': K_TRAIN * F_GTF}.get(dom, F_GTF)
        flops *= chain_steps
        # algorithmic bytes per launch: saved infer/prior (+ d_prior in backward) and the
        # observation experts of each chain set (2 sets of 1, 1 set of 2 modalities)
        per_chain = 4 * Z_DIM * 4 + (2 * Z_DIM * 4 if dom.endswith('bwd') else 0)
        expert_bytes = (2 * Z_DIM * 4 + 1) * 4 * b_dim * T_MAX * (2 if dom.endswith('bwd') else 1)
        bytes_alg = per_chain * N_SETS * b_dim * T_MAX + expert_bytes
        dur = phases[dom] * 1e-3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        traffic = None
        try:       # dram__bytes_read.sum + dram__bytes_write.sum of the same launch (ncu --set full)
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'r1_dominant_kernel.json')))
            if dom == 'filter_s_flt_bwd' and b_dim == B_PER_GPU:
                traffic = prof['dram_bytes_read'] + prof['dram_bytes_write']
        except Exception:
            pass
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': bytes_alg / dur / 1e9, 'peak': hbm_peak,
                    'unit': 'GB/s', 'frac': bytes_alg / dur / 1e9 / hbm_peak, 'traffic': traffic,
                    'algorithmic_bytes': bytes_alg,
                    'peak_source': 'measured' if peaks else 'fallback',
                    'note': 'kernel is FP32-FFMA bound (Z=5,H=20 cannot feed tcgen05); see roofline_fp32'}
        # FP32 FFMA peak measured live (register-only FMA chains on every SM)
        probe_out = torch.zeros(1, device=dev)
        iters, blocks = 20000, 148 * 16
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib.call('bfvi_ffma_probe', _lib.ptr(probe_out), iters, blocks, st)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        lib.call('bfvi_ffma_probe', _lib.ptr(probe_out), iters, blocks, st)
        p1.record()
        torch.cuda.synchronize()
        ffma_peak = blocks * 256 * iters * 16 * 2 / (p0.elapsed_time(p1) * 1e-3) / 1e12
        roofline_fp32 = {'bound': 'fp32_ffma', 'kernel': dom, 'achieved': flops / dur / 1e12,
                         'peak': ffma_peak, 'unit': 'TFLOP/s', 'frac': flops / dur / 1e12 / ffma_peak,
                         'peak_source': 'measured live (bfvi_ffma_probe)',
                         'kernel_share_of_step': phases[dom] / max(sum(phases.values()), 1e-9)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sec, seq_ts = oracle_step_time(B_PER_GPU, 1, 1, cores)
        cpu_baseline = {'value': seq_ts / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': 'oracle port of the reference, 1 warm-up + 1 timed step of the full '
                                  'C2 batch (B=%d, T=%d) on %d torch threads' % (B_PER_GPU, T_MAX, cores)}

    if rank == 0:
        out = {
            'metric': METRIC, 'value': seq_ts_global / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': workload_name(b_dim),
                       'global_batch': b_dim * world, 'seq_len': T_MAX,
                       'parallelism': 'dp%d' % world,
                       'l2': '256 MiB flush write between timed steps; the step itself streams a '
                             '>500 MB workspace (> 126 MB L2)'},
            'clocks': clocks,
            'e2e': {'value': seq_ts_global / (ms_e2e * 1e-3), 'unit': UNIT, 'loop': 'loss read one step late (no per-step GPU stall)',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches,
            'roofline': roofline, 'roofline_fp32': roofline_fp32, 'phase_ms': phases,
            'cpu_baseline': cpu_baseline,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def build_step_args(model, inputs, targets, mask):
    """bfvi_step_args for the profiling call (same values MultiDMM.step passes)."""
    from multimodal_dmm_b200 import _lib
    a = _lib.StepArgs()
    keep = []
    t_max, b_dim = mask.shape[:2]
    a.T, a.B = t_max, b_dim
    for i, m in enumerate(MODS):
        x, y = inputs[m].contiguous(), targets[m].contiguous()
        keep += [x, y]
        a.inputs[i], a.targets[i], a.rec_mults[i] = x.data_ptr(), y.data_ptr(), REC_MULTS[m]
    mk = mask.reshape(t_max, b_dim).to(torch.uint8).contiguous()
    keep.append(mk)
    a.seq_mask, a.kld_mult, a.uni_loss = mk.data_ptr(), KLD_MULT, 1
    a.f_mode, a.s_mode = _lib.MODE_CODES['bfilter'], _lib.MODE_CODES['fsmooth']
    a.f_mult, a.s_mult, a.match_mult = 0.5, 0.5, 0.01
    a.train_particles, a.match_particles, a.sample, a.sample_init = K_TRAIN, K_MATCH, 1, 0
    a.seed, a.b_offset, a.match_count = 2024, 0, -1.0
    return a, keep


if __name__ == '__main__':
    main()
