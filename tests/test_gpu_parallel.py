"""N-rank NCCL data-parallel BFVI step on B200s against the single-GPU step on the concatenated
batch: MultiDMM + parallel.attach — parameters broadcast from rank 0, the same
Philox seed on every rank, noise indexed by the global sequence index, one all-reduce of the flat
gradient.  Both kernel families.  Skips below two devices (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CASES = {
    'small': dict(mods=['x', 'y'], dims=[1, 1], z=5, h=20, t=12, b=9),
    'large': dict(mods=['a', 'b', 'c'], dims=[4, 6, 3], z=32, h=64, t=7, b=6),
}


def _batch(case, seed=5):
    g = torch.Generator().manual_seed(seed)
    x = {m: torch.randn(case['t'], case['b'], d, generator=g) for m, d in zip(case['mods'], case['dims'])}
    x[case['mods'][1]][2:4, 1] = float('nan')
    lengths = sorted([case['t']] * (case['b'] - 2) + [case['t'] - 2, 3], reverse=True)
    mask = torch.zeros(case['t'], case['b'], 1, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[:n, b] = True
        for m in x:
            x[m][n:, b] = float('nan')
    return x, mask, lengths


def _worker(rank, world, port, name, out_path):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import multimodal_dmm_b200.models as models
    from multimodal_dmm_b200 import parallel
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    case = CASES[name]
    x, mask, lengths = _batch(case)
    rec = {m: 1.0 for m in case['mods']}
    torch.manual_seed(100 + rank)                    # DIFFERENT initial weights per rank: attach() must fix that
    model = models.MultiDMM(case['mods'], case['dims'], h_dim=case['h'], z_dim=case['z'], device=dev).train()
    model.precision = 'tf32x3'
    sh = parallel.shard_batch(x, x, mask, lengths, rank, world)
    sh['inputs'] = {k: v.to(dev) for k, v in sh['inputs'].items()}
    sh['targets'] = {k: v.to(dev) for k, v in sh['targets'].items()}
    sh['mask'] = sh['mask'].to(dev)
    parallel.attach(model, sh, base_seed=77 + 1000 * rank)      # rank 0's base seed wins
    loss = parallel.step_or_zero(model, sh, 0.9, rec, train_particles=5, match_particles=7)
    tot = loss.double().clone()
    dist.all_reduce(tot)
    flat = model.last_flat_grad.clone()
    if rank == 0:
        # single-GPU reference on the concatenated batch with rank 0's weights and the same first seed
        ref = models.MultiDMM(case['mods'], case['dims'], h_dim=case['h'], z_dim=case['z'], device=dev).train()
        ref.precision = 'tf32x3'
        ref.load_state_dict(model.state_dict())
        ref.noise_seed = parallel._SharedSeeds(77)()
        cu = {k: v.to(dev) for k, v in x.items()}
        l1 = ref.step(cu, mask.to(dev), 0.9, rec, targets=cu, lengths=lengths, train_particles=5, match_particles=7)
        (l1 / float(sum(lengths))).backward()
        g1 = ref.last_flat_grad
        torch.save({'loss': tot.item(), 'ref_loss': l1.item(),
                    'err': ((flat - g1).norm() / g1.norm()).item(), 'finite': bool(torch.isfinite(flat).all())},
                   out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize('name', sorted(CASES))
def test_nccl_ranks_equal_single_gpu(name, tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs >= 2 CUDA devices (gpurun --gpus 2)')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(world, port, name, out), nprocs=world, join=True)
    r = torch.load(out)
    assert r['finite'], r
    assert abs(r['loss'] - r['ref_loss']) <= 2e-5 * abs(r['ref_loss']), r
    assert r['err'] < 2e-4, r            # float summation order across ranks / atomics only
