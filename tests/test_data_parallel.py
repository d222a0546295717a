"""The N > 1 path on CPU: two gloo ranks each run the BFVI step on a contiguous batch
shard (through the SIMT-emulated kernels, tests only), all-reduce the flat gradient, and
must reproduce the single-process step on the whole batch: global t_max, noise indexed by
the global sequence index (b_offset), linear prior-matching term, SUM of gradients."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 4242


def _worker(rank, world, port, out_path):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import helpers
    from conftest import load_golden
    from multimodal_dmm_b200 import parallel
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    lib = helpers.emu_library()
    fx = load_golden('spirals_ragged')
    sh = parallel.shard_batch(fx['inputs'], fx['targets'], fx['mask'], fx['lengths'], rank, world)
    assert sh['t_max'] == fx['mask'].shape[0]                  # the time axis is never trimmed
    local = dict(fx, inputs=sh['inputs'], targets=sh['targets'], mask=sh['mask'], lengths=sh['lengths'])
    seed = parallel.sync_seed(SEED + 1000 * rank)              # rank 0's seed wins
    assert seed == SEED
    loss, flat, _ = helpers.run_step(lib, local, 'cpu', noise=None, seed=seed, b_offset=sh['b_offset'],
                                     return_flat=True)
    flat = parallel.all_reduce_flat(flat)                      # the step's one collective
    tot = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(tot)
    if rank == 0:
        ref_loss, ref_flat, _ = helpers.run_step(lib, fx, 'cpu', noise=None, seed=SEED, return_flat=True)
        torch.save({'loss': tot.item(), 'ref_loss': ref_loss, 'n_global': sh['n_global'],
                    'err': ((flat - ref_flat).norm() / ref_flat.norm()).item(),
                    'finite': bool(torch.isfinite(flat).all())}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    from multimodal_dmm_b200 import parallel
    for n in (1, 5, 8, 13):
        for world in (1, 2, 3, 8):
            cuts = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(600)
def test_two_rank_gloo_step_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'dp.pt')
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
    import build_emu
    build_emu.build()                      # once, here: two ranks must not rebuild the emulator library concurrently
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r['finite']
    assert abs(r['loss'] - r['ref_loss']) <= 1e-5 * abs(r['ref_loss']), r
    assert r['err'] < 1e-4, r            # float summation order differs across ranks only
