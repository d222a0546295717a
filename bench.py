"""bench.py — BFVI ELBO fwd+bwd sequence-timesteps/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c1|c4|c5]
                    [--batch B] [--scaling weak|strong] [--precision fused|tf32|tf32x3]

One "step" = MultiDMM.step(...) + (loss / sum(lengths)).backward() exactly as
trainer.py:237-243 drives it, on one batch of synthetic data of the named shape.

Workloads (BASELINE.json `configs`, made concrete in SURVEY.md §8d):
  c3 (default) the configuration the metric's "1/2/4/8 B200" is quoted on: M=8 Gaussian modalities, D=16,
               Z=64, H=512, T=1000, K=25, K_match=50, x ~ N(0,1), rec_mults = 1/(D*M), in-kernel Philox
               noise; the named global batch is 65 536 = 8 192 per GPU on the 8-GPU box.  At N GPUs the run is
               the N-GPU SHARD of that job (weak scaling, per-GPU batch fixed, see C3.batch); the per-GPU
               batch actually timed is stated in config.workload.
  c2           spirals model (M=2, D=1, Z=5, H=20), T=100, B=4096 per GPU, 50 % uniformly missing + burst.
  c1           spirals defaults (spirals.py:31-50): T=100, B=100, burst_delete(0.1).
  c4           Weizmann-shaped video model (BASELINE configs[3]): conv image encoders / decoders as custom torch modules,
               Bernoulli + Categorical likelihoods, dropped modalities, Z=H=256, T=25, B=25 — an auxiliary line through
               the composed path (conv / BatchNorm / dense layers of the image modules on this library's FP32 kernels,
               bfvi_conv_* / bfvi_bn2d_* / bfvi_dense_*; no roofline object).
  c5           inference only (BASELINE configs[4]): MultiDMM.forward as Trainer.evaluate calls it (fsmooth, MAP
               estimate, 25 particles in the filtering pass) on the C3-dims model, T=1000, B=1024 per GPU;
               metric bfvi_forward_seq_timesteps_per_sec (forward only), ranks run independent shards.
One NCCL all-reduce of the flat gradient per step for N>1 (no collective inside the step).

Printed JSON (one line, rank 0): value = device-timed throughput with inputs resident in HBM; e2e = the same
through the public API from PINNED HOST buffers (H2D of every step's inputs and targets and a D2H read of
every step's loss inside the timed region; the loss of step i is read while step i+1 runs); roofline =
algorithmic FLOPs / measured time against the measured tensor peak (c3) or the FP32-FFMA
bound of the dominant kernel (c2/c1); cpu_baseline = the reference itself (byte-compiled from the unmodified
sources into oracle/_ref by oracle/build_ref.py; kind "reference") timed on this box's host cores on a bounded
sample — the oracle port (kind "port") only where oracle/_ref is absent.

--impl reference: the reference's own MultiDMM.step + backward (oracle/_ref: sourceless bytecode of the unmodified
reference, which travels to the GPU box; falls back to the pinned PyTorch-CPU port oracle/bfvi_oracle.py where that
directory is absent) on all host cores, on a bounded sample of the SAME workload; it prints the same `config`
object as our arm.
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

METRIC = 'bfvi_elbo_fwd_bwd_seq_timesteps_per_sec'
UNIT = 'seq-timesteps/s'
KLD_MULT = 1.0


# ----------------------------------------------------------------------------
# workloads
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


def _burst(inp, rng, t_max, b_dim, frac=0.1):
    """burst_delete(frac): one burst per sequence per modality (datasets/multiseq.py:428-434)."""
    burst = int(frac * t_max)
    start = rng.randint(t_max, size=b_dim)
    tt = np.arange(t_max)[:, None]
    inp[(tt >= start[None, :]) & (tt < np.minimum(start + burst, t_max)[None, :]), 0] = np.nan


def make_c2_batch(b_dim, t_max=100, seed=1):
    data, rng = spirals_batch(b_dim, t_max, seed)
    targets, inputs = {}, {}
    n_del = int(0.5 * t_max)
    for m in C2.mods:
        tgt = data[m].copy()
        # corrupt_(0.5, 'uniform'): exactly n_del timesteps per sequence, no replacement
        order = np.argsort(rng.rand(t_max, b_dim), axis=0)[:n_del]
        tgt[order, np.arange(b_dim)[None, :], 0] = np.nan
        inp = tgt.copy()
        _burst(inp, rng, t_max, b_dim)
        targets[m], inputs[m] = torch.from_numpy(tgt), torch.from_numpy(inp)
    return inputs, targets, torch.ones(t_max, b_dim, 1, dtype=torch.bool), [t_max] * b_dim


def make_c1_batch(b_dim, t_max=100, seed=1):
    data, rng = spirals_batch(b_dim, t_max, seed)
    targets, inputs = {}, {}
    for m in C2.mods:
        inp = data[m].copy()
        _burst(inp, rng, t_max, b_dim)
        targets[m], inputs[m] = torch.from_numpy(data[m].copy()), torch.from_numpy(inp)
    return inputs, targets, torch.ones(t_max, b_dim, 1, dtype=torch.bool), [t_max] * b_dim


def make_c3_batch(b_dim, t_max=1000, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = {m: torch.randn(t_max, b_dim, C3.dims[0], generator=g) for m in C3.mods}
    # targets are the same values in their own buffers (the trainer passes two dicts, trainer.py:231-237)
    return x, {m: v.clone() for m, v in x.items()}, torch.ones(t_max, b_dim, 1, dtype=torch.bool), [t_max] * b_dim


class Workload(object):
    def __init__(self, key, mods, dims, z, h, t_max, batch, make, rec, desc, ref_sample, k_train=25, k_match=50):
        self.key, self.mods, self.dims, self.z, self.h, self.t_max, self.batch = key, mods, dims, z, h, t_max, batch
        self.make, self.rec, self.desc, self.ref_sample = make, rec, desc, ref_sample
        self.k_train, self.k_match = k_train, k_match
        m = len(mods)
        f_gtf = 8 * z * h + 4 * z * z
        f_enc = sum(2 * h * (d + 2 * z) for d in dims)
        f_dec = sum(2 * h * (z + 2 * d) for d in dims)
        # F_fwd = (1+M)(K+2) F_gtf + sum F_enc + 4 sum F_dec ; F_step = 3 F_fwd
        self.flops_per_seq_ts = 3 * ((1 + m) * (k_train + 2) * f_gtf + f_enc + 4 * f_dec)
        self.bytes_per_seq_ts = 2 * 4 * sum(dims)
        self.f_gtf, self.n_sets = f_gtf, (1 + m if m > 1 else 1)

    def name(self, b_dim):
        return self.desc % {'B': b_dim}


C2 = Workload('c2', ['spiral-x', 'spiral-y'], [1, 1], 5, 20, 100, 4096, make_c2_batch, 1.0,
              'C2: spirals BFVI step, M=2 D=1 Z=5 H=20, T=100, B=%(B)d per GPU, 50%% uniform missing + 10%% burst, '
              'K=25, K_match=50', (896, 100))
C1 = Workload('c1', ['spiral-x', 'spiral-y'], [1, 1], 5, 20, 100, 100, make_c1_batch, 0.5,
              'C1: spirals.py defaults, M=2 D=1 Z=5 H=20, T=100, B=%(B)d, burst_delete(0.1), K=25, K_match=50',
              (100, 100))
# per-GPU batch of the c3 line: the 8-GPU shard of the named job is 8 192 sequences; bfvi_step_fwd_bwd walks the batch
# in tiles (workspace bounded), so any batch runs as ONE step() call; the default keeps a driver run within minutes
C3 = Workload('c3', ['m%d' % i for i in range(8)], [16] * 8, 64, 512, 1000, 2048, make_c3_batch, 1.0 / (16 * 8),
              'C3: scaled MDMM BFVI step, M=8 D=16 Z=64 H=512, T=1000, K=25, K_match=50, B=%(B)d per GPU, N(0,1) data, '
              'Philox noise (BASELINE names global batch 65536 = 8192 per GPU on 8 B200; the per-GPU batch is cut to '
              'keep a 25-step run within minutes: a step of 8192 x 1000 takes ~20 s; --batch 8192 runs the full shard, '
              'same code path: the step walks the batch in tiles)', (24, 100))
# C5 (BASELINE configs[4]): inference-only reconstruction sweep point — forward(), fsmooth, MAP estimate, K particles in the
# filtering pass — on the C3-dims model; metric = seq-timesteps/s of forward (NOT the fwd+bwd training metric)
C5 = Workload('c5', ['m%d' % i for i in range(8)], [16] * 8, 64, 512, 1000, 1024, make_c3_batch, 1.0 / (16 * 8),
              'C5: inference-only forward (fsmooth, sample=False, flt_particles=25) of the scaled MDMM, M=8 D=16 Z=64 H=512, '
              'T=1000, B=%(B)d per GPU, N(0,1) data, Philox noise', (24, 100))
# C4 (BASELINE configs[3]): Weizmann-shaped video model — conv image encoders / decoders (common.ImageEncoder / ImageDecoder passed as
# custom encoders= / decoders=, weizmann.py:53-77; their layers run on bfvi_conv_* / bfvi_bn2d_* / bfvi_dense_*), Bernoulli + Categorical likelihoods, dropped modalities — through the
# composed path: our temporal core (fused z_filter), likelihood kernels and categorical encoder / decoder kernels around it
C4_MODS = ['video', 'mask', 'action']
C4_DIMS = {'video': (3, 64, 64), 'mask': (1, 64, 64), 'action': 10}
C4_DISTS = {'video': 'Bernoulli', 'mask': 'Bernoulli', 'action': 'Categorical'}
C4_REC = {'video': 1.0, 'mask': 1.0, 'action': 10.0}


def make_c4_batch(b_dim, t_max=25, seed=1):
    g = torch.Generator().manual_seed(seed)
    targets = {'video': torch.rand(t_max, b_dim, 3, 64, 64, generator=g),
               'mask': (torch.rand(t_max, b_dim, 1, 64, 64, generator=g) > 0.5).float(),
               'action': torch.randint(0, 10, (t_max, b_dim, 1), generator=g).float()}
    inputs = {'video': targets['video'].clone(),                       # mask and action are dropped (all NaN),
              'mask': torch.full_like(targets['mask'], float('nan')),  # trainer.py:289-290
              'action': torch.full_like(targets['action'], float('nan'))}
    burst = int(0.2 * t_max)                                           # burst_delete(0.2) on the video
    start = torch.randint(0, t_max, (b_dim,), generator=g)
    for b in range(b_dim):
        inputs['video'][start[b]:start[b] + burst, b] = float('nan')
    return inputs, targets, torch.ones(t_max, b_dim, 1, dtype=torch.bool), [t_max] * b_dim


def build_c4_model(models_pkg, device, z_dim=256, h_dim=256):
    """Same construction as weizmann.py:53-77."""
    c = models_pkg.common
    enc = {'video': c.ImageEncoder(z_dim, True), 'mask': c.ImageEncoder(z_dim, True, n_channels=1)}
    dec = {'video': c.ImageDecoder(z_dim), 'mask': c.ImageDecoder(z_dim, n_channels=1)}
    return models_pkg.MultiDMM(C4_MODS, dims=[C4_DIMS[m] for m in C4_MODS], dists=[C4_DISTS[m] for m in C4_MODS],
                               encoders=enc, decoders=dec, z_dim=z_dim, h_dim=h_dim, device=device)


C4 = Workload('c4', C4_MODS, [3 * 64 * 64, 64 * 64, 10], 256, 256, 25, 25, make_c4_batch, 1.0,
              'C4: Weizmann-shaped video BFVI step (video 3x64x64 + silhouette mask Bernoulli, 10-way action Categorical; conv '
              'encoders / decoders as custom torch modules; mask and action dropped from the inputs, burst_delete(0.2) on the '
              'video), Z=H=256, T=25, B=%(B)d per GPU, K=25, K_match=50', (8, 25))
WORKLOADS = {'c1': C1, 'c2': C2, 'c3': C3, 'c4': C4, 'c5': C5}
METRIC_FORWARD = 'bfvi_forward_seq_timesteps_per_sec'


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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def config_of(wl, b_dim, world, scaling):
    """The `config` object: IDENTICAL in our arm and in --impl reference (which times a bounded sample
    of the same workload and says so in its cpu_baseline.sample)."""
    return {'workload': wl.name(b_dim), 'global_batch': b_dim * world, 'seq_len': wl.t_max,
            'parallelism': 'dp%d' % world, 'scaling': scaling,
            'reference_arm_sample': 'CPU arm (--impl reference / cpu_baseline) times B=%d, T=%d of this model '
                                    'and data generator on all host cores' % wl.ref_sample,
            'l2': '256 MiB flush write between timed steps; the step streams a workspace far larger than the '
                  '126 MB L2'}


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------
def oracle_step_time(wl, b_dim, t_max, steps, warmup, threads):
    """Mean seconds per step + seq-timesteps per step of the oracle port at (b_dim, t_max)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import bfvi_oracle as bo
    torch.set_num_threads(threads)
    inputs, targets, mask, lengths = wl.make(b_dim, t_max, 1)
    rec = {m: wl.rec for m in wl.mods}
    params = bo.init_params(wl.mods, wl.dims, h_dim=wl.h, z_dim=wl.z, seed=1)
    for p in params.values():
        p.requires_grad_(True)
    orc = bo.OracleDMM(wl.mods, wl.dims, params, h_dim=wl.h, z_dim=wl.z, draw=bo.RandomDraw(seed=3))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = orc.step(inputs, mask, KLD_MULT, rec, targets=targets, lengths=lengths,
                        train_particles=wl.k_train, match_particles=wl.k_match)
        (loss / sum(lengths)).backward()
        for p in params.values():
            p.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return float(np.mean(times)), b_dim * t_max


def reference_step_time(wl, b_dim, t_max, steps, warmup, threads):
    """Mean seconds per step + seq-timesteps per step of the UNMODIFIED reference's own MultiDMM.step + backward
    (models/dmm.py:503-554, trainer.py:237-243) on the host cores, imported from the byte-compiled copy under
    oracle/_ref/ (oracle/build_ref.py; /root/reference itself does not exist on the GPU box).  None when absent."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_shim
    if ref_shim.reference_root() is None:
        return None
    ref_models = ref_shim.import_reference_models()
    torch.set_num_threads(threads)
    inputs, targets, mask, lengths = wl.make(b_dim, t_max, 1)
    rec = dict(C4_REC) if wl.key == 'c4' else {m: wl.rec for m in wl.mods}
    torch.manual_seed(1)
    if wl.key == 'c4':
        model = build_c4_model(ref_models, torch.device('cpu'))
    else:
        model = ref_models.MultiDMM(wl.mods, wl.dims, h_dim=wl.h, z_dim=wl.z, device=torch.device('cpu'))
    model.train()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if wl.key == 'c5':                           # Trainer.evaluate's call (trainer.py:294-296)
            model.eval()
            with torch.no_grad():
                model(inputs, lengths=lengths, sample=False, flt_particles=wl.k_train)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            continue
        loss = model.step(inputs, mask, KLD_MULT, rec, targets=targets, lengths=lengths,
                          train_particles=wl.k_train, match_particles=wl.k_match)
        (loss / sum(lengths)).backward()
        model.zero_grad()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return float(np.mean(times)), b_dim * t_max


def cpu_step_time(wl, b_dim, t_max, steps, warmup, threads):
    """(seconds per step, seq-timesteps per step, kind): the reference itself when oracle/_ref/ holds it, else the
    oracle port of the same algorithm."""
    r = reference_step_time(wl, b_dim, t_max, steps, warmup, threads)
    if r is not None:
        return r[0], r[1], 'reference'
    sec, seq_ts = oracle_step_time(wl, b_dim, t_max, steps, warmup, threads)
    return sec, seq_ts, 'port'


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cores = os.cpu_count() or 1
    b_ref, t_ref = wl.ref_sample
    sec, seq_ts, kind = cpu_step_time(wl, b_ref, t_ref, args.steps, args.warmup, cores)
    value = seq_ts / sec
    b_dim = per_gpu_batch(args, wl, world)
    sample = ('%s on %d torch threads: %s workload at B=%d, T=%d (K=%d particles), %d warm-up + %d timed steps' %
              ("the reference's own MultiDMM.step + backward (byte-compiled from the unmodified sources, oracle/_ref)"
               if kind == 'reference' else 'oracle port of the reference', cores, wl.key.upper(), b_ref, t_ref,
               wl.k_train, args.warmup, args.steps))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC_FORWARD if wl.key == 'c5' else METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'fp32',
        'data': 'synthetic', 'config': config_of(wl, b_dim, world, args.scaling),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def per_gpu_batch(args, wl, world):
    b = args.batch if args.batch else wl.batch
    if args.scaling == 'strong':
        b = max(1, b // world)
    return b


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='sequences per GPU (weak) / in total (strong); 0 = workload default')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--precision', default='fused', choices=['fused', 'tf32', 'tf32x3'],
                    help='GEMM operand precision of the large-dim family (c3)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=0, help='steps of the end-to-end loop (0 = min(steps, 5))')
    ap.add_argument('--seq-len', type=int, default=0, help='override T (profiling runs under ncu only; 0 = workload T)')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.seq_len > 0:
        wl.t_max = args.seq_len
        wl.desc = wl.desc.replace('T=1000', 'T=%d (PROFILING OVERRIDE of T=1000)' % args.seq_len)
    if args.impl == 'reference':
        return run_reference(args, wl)
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
    if wl.key == 'c4':
        return run_weizmann(args, wl, models, dev, rank, local_rank, world, dist)
    model = models.MultiDMM(wl.mods, wl.dims, h_dim=wl.h, z_dim=wl.z, device=dev).train()
    large = wl.key == 'c3'
    if large:
        model.precision = args.precision
    if wl.key == 'c5':
        return run_inference(args, wl, model, dev, rank, local_rank, world, dist)
    b_dim = per_gpu_batch(args, wl, world)
    t_max = wl.t_max
    inputs_h, targets_h, mask_h, lengths = wl.make(b_dim, t_max, (1234 if large else 1) + rank)
    pin = lambda d: {k: v.pin_memory() for k, v in d.items()}
    inputs_h, targets_h = pin(inputs_h), pin(targets_h)
    mask_d = mask_h.to(dev)
    inputs_d = {k: v.to(dev) for k, v in inputs_h.items()}
    targets_d = {k: v.to(dev) for k, v in targets_h.items()}
    rec = {m: wl.rec for m in wl.mods}
    n_global = float(sum(lengths) * world)          # normalise by the GLOBAL sum(lengths)
    model.b_offset = rank * b_dim
    if world > 1:
        model.grad_sync = lambda g: dist.all_reduce(g)      # one NCCL all-reduce of the flat grad
    model.noise_seed = None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def one_step(inp, tgt):
        loss = model.step(inp, mask_d, KLD_MULT, rec, targets=tgt, lengths=lengths,
                          train_particles=wl.k_train, match_particles=wl.k_match)
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
    dispatch = _lib.load().last_dispatch()

    # ---- end-to-end timing: pinned host inputs -> H2D -> step -> D2H loss ----------
    # A training loop that does not stall the GPU on the loss: every step copies its batch from pinned host
    # memory (same stream, ahead of the step) and copies its loss to pinned host memory asynchronously; the
    # host reads step i's loss while step i+1 runs.  Every step still pays its H2D copy and its D2H loss read
    # inside the timed region.
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else min(args.steps, 5)
    h2d = sum(v.numel() * 4 for v in inputs_h.values()) + sum(v.numel() * 4 for v in targets_h.values())
    del inputs_d, targets_d                           # the e2e loop owns its device copies
    one_step({k: v.to(dev, non_blocking=True) for k, v in inputs_h.items()},
             {k: v.to(dev, non_blocking=True) for k, v in targets_h.items()}).item()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    main_stream = torch.cuda.current_stream(dev)
    loss_pinned = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    loss_host = None
    for i in range(e2e_steps):
        j = i % 2
        inp = {k: v.to(dev, non_blocking=True) for k, v in inputs_h.items()}
        tgt = {k: v.to(dev, non_blocking=True) for k, v in targets_h.items()}
        loss = one_step(inp, tgt)
        loss_pinned[j].copy_(loss.detach(), non_blocking=True)
        loss_ready[j].record(main_stream)
        del inp, tgt
        if i > 0:                                      # D2H read of the PREVIOUS step's result
            loss_ready[1 - j].synchronize()
            loss_host = float(loss_pinned[1 - j])
    loss_ready[(e2e_steps - 1) % 2].synchronize()
    loss_host = float(loss_pinned[(e2e_steps - 1) % 2])
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    inputs_d = {k: v.to(dev) for k, v in inputs_h.items()}
    targets_d = {k: v.to(dev) for k, v in targets_h.items()}

    # ---- max over ranks ------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t[0].item(), t[1].item()
    seq_ts_global = b_dim * t_max * world
    peaks = load_peaks()

    roofline = roofline_fp32 = phases = kernel_probe = None
    if rank == 0 and large:
        # whole step: algorithmic FLOPs (recomputation NOT counted) over the device-timed step
        # against the measured sustained dense BF16 tensor peak (TF32 operands run at half that rate, so a
        # perfect TF32 step reads 0.5)
        peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
        ach = wl.flops_per_seq_ts * b_dim * t_max / (ms * 1e-3) / 1e12
        traffic = traffic_src = None
        try:                                         # committed ncu --set full capture of the dominant kernel
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'r2_dominant_kernel.json')))
            traffic = (prof['dram_bytes_read'] + prof['dram_bytes_write']) / float(prof['rows'])   # per latent row
            traffic_src = prof['source']
        except Exception:
            pass
        step_roofline = {'bound': 'tensor', 'kernel': 'whole step', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': ach / peak, 'algorithmic_flops_per_seq_ts': wl.flops_per_seq_ts,
                         'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback (B200_PROFILING.md)'}
        # dominant kernels, each timed alone (CUDA events around back-to-back launches on the launching stream) on the
        # rows ONE particle-pass time step launches: (1 + M) chain sets x batch tile x K particles
        tile = b_dim
        for d in dispatch:
            if d.startswith('step:batch_tiles='):          # "step:batch_tiles=<n> x <sequences per tile> lanes=<l>"
                tile = int(d.split('x')[-1].split()[0])
        rows = (1 + len(wl.mods)) * tile * wl.k_train
        kernel_probe = gtf_kernel_rooflines(model, wl, rows, peaks)
        dom = kernel_probe['gtf_fwd_kernel<keep>']
        roofline = {'bound': 'tensor', 'kernel': 'gtf_fwd_kernel<keep> (fused transition forward, the largest share of the '
                                                 'step; one launch = one particle-pass time step of a batch tile)',
                    'achieved': dom['achieved'], 'peak': dom['peak'], 'unit': 'TFLOP/s', 'frac': dom['frac'],
                    'traffic': None if traffic is None else traffic * rows, 'traffic_source': traffic_src,
                    'algorithmic_flops_per_launch': dom['flops'], 'rows_per_launch': rows,
                    'ms_per_launch': dom['ms'], 'peak_source': dom['peak_source'], 'whole_step': step_roofline}
    if rank == 0 and not large:
        roofline, roofline_fp32, phases = small_roofline(model, wl, inputs_d, targets_d, mask_d, rec, b_dim, t_max,
                                                         flush, peaks, args)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        b_ref, t_ref = wl.ref_sample
        sec, seq_ts, kind = cpu_step_time(wl, b_ref, t_ref, 3, 1, cores)
        cpu_baseline = {'value': seq_ts / sec, 'unit': UNIT, 'cores': cores, 'kind': kind,
                        'sample': '%s, 1 warm-up + 3 timed steps of the %s workload at B=%d, T=%d on %d torch threads' %
                                  ("the reference's own MultiDMM.step + backward (oracle/_ref)" if kind == 'reference'
                                   else 'oracle port of the reference', wl.key.upper(), b_ref, t_ref, cores)}

    if rank == 0:
        out = {
            'metric': METRIC, 'value': seq_ts_global / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None,
            'dtype': {'fused': 'fp16x3 split forward (fp32-class) / tf32 input gradients / fp16 weight-gradient '
                               'operands, fp32 accumulate', 'tf32': 'tf32 operands, fp32 accumulate',
                      'tf32x3': 'tf32x3 (fp32-class), fp32 accumulate'}[args.precision] if large else 'fp32',
            'data': 'synthetic',
            'config': config_of(wl, b_dim, world, args.scaling),
            'clocks': clocks,
            'e2e': {'value': seq_ts_global / (ms_e2e * 1e-3), 'unit': UNIT, 'steps': e2e_steps,
                    'loop': 'loss read one step late (no per-step GPU stall)',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches,
            'roofline': roofline, 'roofline_fp32': roofline_fp32, 'phase_ms': phases,
            'kernel_probe': kernel_probe, 'dispatch': dispatch[:24],
            'cpu_baseline': cpu_baseline,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def run_weizmann(args, wl, models, dev, rank, local_rank, world, dist):
    """--workload c4: MultiDMM.step + backward of the Weizmann-shaped model through the composed path (the image modules'
    convolutions, BatchNorm -> ReLU and dense layers, the temporal core, the Bernoulli / categorical likelihoods and the
    categorical encoder / decoder are all this library's kernels).  Data parallel like the other training workloads: one all-reduce of the gradients per step."""
    from multimodal_dmm_b200 import _lib
    model = build_c4_model(models, dev).train()
    b_dim, t_max = per_gpu_batch(args, wl, world), wl.t_max
    inputs_h, targets_h, mask_h, lengths = wl.make(b_dim, t_max, 1 + rank)
    pin = lambda d: {k: v.pin_memory() for k, v in d.items()}
    inputs_h, targets_h = pin(inputs_h), pin(targets_h)
    mask_d = mask_h.to(dev)
    inputs_d = {k: v.to(dev) for k, v in inputs_h.items()}
    targets_d = {k: v.to(dev) for k, v in targets_h.items()}
    n_global = float(sum(lengths) * world)
    params = [p for p in model.parameters() if p.requires_grad]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def one_step(inp, tgt):
        loss = model.step(inp, mask_d, KLD_MULT, C4_REC, targets=tgt, lengths=lengths,
                          train_particles=wl.k_train, match_particles=wl.k_match)
        (loss / n_global).backward()
        if dist is not None:                           # (custom modules keep their own .grad tensors: flatten, one all-reduce)
            flat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
            dist.all_reduce(flat)
        for p in params:
            p.grad = None
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(inputs_d, targets_d)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(float(i))
        ev[i][0].record()
        one_step(inputs_d, targets_d)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    dispatch = _lib.load().last_dispatch()
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else min(args.steps, 5)
    h2d = sum(v.numel() * 4 for v in inputs_h.values()) + sum(v.numel() * 4 for v in targets_h.values())
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        inp = {k: v.to(dev, non_blocking=True) for k, v in inputs_h.items()}
        tgt = {k: v.to(dev, non_blocking=True) for k, v in targets_h.items()}
        float(one_step(inp, tgt).detach())
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t[0].item(), t[1].item()
    if rank != 0:
        return
    seq_ts = b_dim * t_max * world
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        b_ref, t_ref = wl.ref_sample
        r = reference_step_time(wl, b_ref, t_ref, 2, 1, cores)
        if r is not None:
            cpu_baseline = {'value': r[1] / r[0], 'unit': UNIT, 'cores': cores, 'kind': 'reference',
                            'sample': "the reference's own MultiDMM.step + backward of the C4 model at B=%d, T=%d on %d torch "
                                      'threads, 1 warm-up + 2 timed steps' % (b_ref, t_ref, cores)}
    print(json.dumps({
        'metric': METRIC, 'value': seq_ts / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'fp32 image modules (direct FFMA convolutions, bfvi_conv_*) around the tcgen05 temporal core (3xTF32, fp32-class)', 'data': 'synthetic',
        'config': config_of(wl, b_dim, world, args.scaling), 'clocks': clocks,
        'e2e': {'value': seq_ts / (ms_e2e * 1e-3), 'unit': UNIT, 'steps': e2e_steps,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
        'gpu_launches': None, 'roofline': None,
        'note': 'auxiliary line: image encoders / decoders (SURVEY 8f-3), temporal core, likelihood kernels and the categorical '
                'encoder / decoder all run on this library (no cuDNN / cuBLAS on the path)',
        'dispatch': dispatch, 'cpu_baseline': cpu_baseline}))


def run_inference(args, wl, model, dev, rank, local_rank, world, dist):
    """--workload c5: MultiDMM.forward (models/dmm.py:420-494 as Trainer.evaluate calls it) on one batch per step;
    ranks run independent shards (no collective on this path)."""
    from multimodal_dmm_b200 import _lib
    model.eval()
    model.precision = args.precision
    b_dim, t_max = per_gpu_batch(args, wl, world), wl.t_max
    inputs_h, _, _, lengths = wl.make(b_dim, t_max, 1234 + rank)
    inputs_h = {k: v.pin_memory() for k, v in inputs_h.items()}
    inputs_d = {k: v.to(dev) for k, v in inputs_h.items()}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def one_step(inp):
        with torch.no_grad():
            infer, prior, recon = model(inp, lengths=lengths, sample=False, flt_particles=wl.k_train)
        return recon[wl.mods[0]][0].sum() + infer[0].sum()          # a scalar that depends on the whole result

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(inputs_d)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(float(i))
        ev[i][0].record()
        one_step(inputs_d)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    dispatch = _lib.load().last_dispatch()
    # end to end: pinned host inputs -> H2D -> forward -> D2H of the result scalar, every step
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else min(args.steps, 5)
    h2d = sum(v.numel() * 4 for v in inputs_h.values())
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        inp = {k: v.to(dev, non_blocking=True) for k, v in inputs_h.items()}
        float(one_step(inp))
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t[0].item(), t[1].item()
    if rank != 0:
        return
    peaks = load_peaks()
    m = len(wl.mods)
    f_enc = sum(2 * wl.h * (d + 2 * wl.z) for d in wl.dims)
    f_dec = sum(2 * wl.h * (wl.z + 2 * d) for d in wl.dims)
    flops_per_seq_ts = f_enc + (wl.k_train + 1) * wl.f_gtf + f_dec
    seq_ts = b_dim * t_max * world
    peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
    ach = flops_per_seq_ts * b_dim * t_max / (ms * 1e-3) / 1e12
    probe = gtf_kernel_rooflines(model, wl, b_dim * wl.k_train, peaks)
    dom = probe['gtf_fwd_kernel'] if probe else None
    # launches: per present modality prep + 2 GEMM launches + softplus; per pass T step kernels + (T - 1) transitions
    # (+ 2 pack launches per direction); per decoder 2 GEMM launches + softplus
    launches = (m * 4 + 2 * (2 * t_max - 1) + 4 + m * 3) * args.steps
    roofline = None if dom is None else {
        'bound': 'tensor', 'kernel': 'gtf_fwd_kernel (fused transition forward of the K-particle filtering pass)',
        'achieved': dom['achieved'], 'peak': dom['peak'], 'unit': 'TFLOP/s', 'frac': dom['frac'], 'traffic': None,
        'ms_per_launch': dom['ms'], 'rows_per_launch': b_dim * wl.k_train, 'peak_source': dom['peak_source'],
        'whole_step': {'achieved': ach, 'peak': peak, 'frac': ach / peak, 'algorithmic_flops_per_seq_ts': flops_per_seq_ts}}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        b_ref, t_ref = wl.ref_sample
        sec, n_ts, kind = cpu_step_time(wl, b_ref, t_ref, 3, 1, cores)
        cpu_baseline = {'value': n_ts / sec, 'unit': UNIT, 'cores': cores, 'kind': kind,
                        'sample': 'forward() of the C5 workload at B=%d, T=%d on %d torch threads, 1 warm-up + 3 timed' %
                                  (b_ref, t_ref, cores)}
    print(json.dumps({
        'metric': METRIC_FORWARD, 'value': seq_ts / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'fp16x3 split forward (fp32-class), fp32 accumulate' if args.precision == 'fused' else args.precision,
        'data': 'synthetic', 'config': config_of(wl, b_dim, world, args.scaling), 'clocks': clocks,
        'e2e': {'value': seq_ts / (ms_e2e * 1e-3), 'unit': UNIT, 'steps': e2e_steps,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
        'gpu_launches': launches, 'roofline': roofline, 'kernel_probe': probe, 'dispatch': dispatch,
        'cpu_baseline': cpu_baseline}))


def gtf_kernel_rooflines(model, wl, rows, peaks):
    """Per-kernel roofline of the fused transition kernels: bfvi_gtf_probe times each kernel alone on `rows` latent rows.
    Algorithmic FLOPs per row (F_gtf = 8ZH + 4Z^2): forward F_gtf; input gradient F_gtf; weight gradients F_gtf
    (the KEEP forward is the backward's recompute: its F_gtf is NOT algorithmic work of the step, but it is what the
    launch computes, and it is reported as such).  Peak: measured burst dense BF16 (a kernel timed alone)."""
    from multimodal_dmm_b200 import _lib
    lib = _lib.load()
    dev = model._flat.device
    peak = float(peaks.get('bf16_tflops', 1590.0))
    nbytes = int(lib.dll.bfvi_gtf_workspace(C.byref(model._cmodel), rows))
    if nbytes == 0:
        return None
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    ws = ws[(-ws.data_ptr()) % 256:]
    z = torch.randn(rows, wl.z, device=dev)
    scratch = torch.empty(5 * rows * 64, device=dev)
    out = {}
    names = ['gtf_fwd_kernel', 'gtf_fwd_kernel<keep>', 'gtf_bwd_kernel', 'wgrad16_kernel']
    ms = C.c_float(0.0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for which, name in enumerate(names):
        lib.call('bfvi_gtf_probe', C.byref(model._cmodel), _lib.ptr(model._flat), 1, which, _lib.ptr(z), rows, 3,
                 _lib.ptr(scratch), _lib.ptr(ws), C.c_size_t(nbytes), C.byref(ms), st)      # warm-up
        lib.call('bfvi_gtf_probe', C.byref(model._cmodel), _lib.ptr(model._flat), 1, which, _lib.ptr(z), rows, 10,
                 _lib.ptr(scratch), _lib.ptr(ws), C.c_size_t(nbytes), C.byref(ms), st)
        # the Z x Z corners outside a kernel are not counted for it: input gradient without the std layer, weight
        # gradients of the four H-wide layers only
        per_row = [wl.f_gtf, wl.f_gtf, wl.f_gtf - 2 * wl.z * wl.z, 8 * wl.z * wl.h][which]
        flops = float(rows) * per_row
        out[name] = {'ms': ms.value, 'flops': flops, 'achieved': flops / (ms.value * 1e-3) / 1e12, 'peak': peak,
                     'frac': flops / (ms.value * 1e-3) / 1e12 / peak,
                     'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst)' if peaks else 'fallback'}
    return out


def small_roofline(model, wl, inputs_d, targets_d, mask_d, rec, b_dim, t_max, flush, peaks, args):
    """Dominant-kernel roofline of the small-dim family from CUDA-event phase timing (bfvi_step_profile)."""
    from multimodal_dmm_b200 import _lib
    dev = mask_d.device
    lib = _lib.load()
    model._ensure_flat()
    fx_args, keep = build_step_args(model, wl, inputs_d, targets_d, mask_d, rec)
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
    chain_steps = wl.n_sets * b_dim * (t_max - 1)
    # algorithmic work of the dominant kernel per launch (recomputation NOT counted)
    flops = {'filter_s_flt_bwd': 2 * wl.k_train * wl.f_gtf, 'filter_s_flt_fwd': wl.k_train * wl.f_gtf}.get(dom, wl.f_gtf)
    flops *= chain_steps
    dur = phases[dom] * 1e-3
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
    traffic = None
    try:       # dram__bytes_read.sum + dram__bytes_write.sum of the same launch (ncu --set full)
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'r1_dominant_kernel.json')))
        if dom == 'filter_s_flt_bwd' and b_dim == C2.batch and wl.key == 'c2':
            traffic = prof['dram_bytes_read'] + prof['dram_bytes_write']
    except Exception:
        pass
    # the binding roofline of this family is FP32 issue (5x20 matrices cannot feed tcgen05; HBM traffic is
    # 0.6 % of DRAM throughput): the contract object carries that bound
    roofline = {'bound': 'fp32_ffma', 'kernel': dom, 'achieved': flops / dur / 1e12, 'peak': ffma_peak,
                'unit': 'TFLOP/s', 'frac': flops / dur / 1e12 / ffma_peak, 'traffic': traffic,
                'peak_source': 'measured live (bfvi_ffma_probe: register-only FFMA chains on every SM)',
                'kernel_share_of_step': phases[dom] / max(sum(phases.values()), 1e-9)}
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    bytes_alg = wl.bytes_per_seq_ts * b_dim * t_max
    step_s = sum(phases.values()) * 1e-3
    roofline_hbm = {'bound': 'hbm', 'kernel': 'whole step', 'achieved': bytes_alg / step_s / 1e9, 'peak': hbm_peak,
                    'unit': 'GB/s', 'frac': bytes_alg / step_s / 1e9 / hbm_peak,
                    'note': 'SURVEY 8d algorithmic bytes (inputs + targets): not the bound of this family'}
    return roofline, roofline_hbm, phases


def build_step_args(model, wl, inputs, targets, mask, rec):
    """bfvi_step_args for the profiling call (same values MultiDMM.step passes)."""
    from multimodal_dmm_b200 import _lib
    a = _lib.StepArgs()
    keep = []
    t_max, b_dim = mask.shape[:2]
    a.T, a.B = t_max, b_dim
    for i, m in enumerate(wl.mods):
        x, y = inputs[m].contiguous(), targets[m].contiguous()
        keep += [x, y]
        a.inputs[i], a.targets[i], a.rec_mults[i] = x.data_ptr(), y.data_ptr(), rec[m]
    mk = mask.reshape(t_max, b_dim).to(torch.uint8).contiguous()
    keep.append(mk)
    a.seq_mask, a.kld_mult, a.uni_loss = mk.data_ptr(), KLD_MULT, 1
    a.f_mode, a.s_mode = _lib.MODE_CODES['bfilter'], _lib.MODE_CODES['fsmooth']
    a.f_mult, a.s_mult, a.match_mult = 0.5, 0.5, 0.01
    a.train_particles, a.match_particles, a.sample, a.sample_init = wl.k_train, wl.k_match, 1, 0
    a.seed, a.b_offset, a.match_count = 2024, 0, -1.0
    return a, keep


if __name__ == '__main__':
    main()
