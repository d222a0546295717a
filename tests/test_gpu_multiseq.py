"""multimodal_dmm_b200.multiseq on the GPU: bit-exact against the reference's golden outputs
(tests/golden/multiseq, numpy draws replayed), seeded device draws against the integer
restatement, and size-independent properties at BASELINE.json's C2 batch shape."""
import os

import numpy as np
import pytest
import torch

import multiseq_oracle as orc
from test_emu_multiseq import GOLD, case_id, same
from multimodal_dmm_b200 import multiseq

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', GOLD['delete'], ids=case_id)
def test_kernels_match_reference_golden(case):
    np.random.seed(case['np_seed'])
    inputs = {m: v.cuda() for m, v in case['inputs'].items()}
    got = getattr(multiseq, case['op'])(inputs, *case['args'], lengths=case['lengths'], modalities=case['modalities'])
    for m in inputs:
        assert got[m].is_cuda and same(got[m].cpu().numpy(), case['outputs'][m].numpy()), m
        assert same(inputs[m].cpu().numpy(), case['inputs'][m].numpy())


@pytest.mark.parametrize('i', range(len(GOLD['collate'])))
def test_collation_matches_reference_golden(i):
    c = GOLD['collate'][i]
    seqs = [s.numpy() for s in c['sequences']]
    assert same(multiseq.pad_and_merge(seqs, c['max_len']).cpu().numpy(), c['output'].numpy())
    if 'mask' in c:
        assert torch.equal(multiseq.len_to_mask(c['lengths']).cpu(), c['mask'])


@pytest.mark.parametrize('mode', [0, 1])
def test_seeded_draws_match_integer_restatement(mode):
    T, lengths = 40, [40, 33, 33, 17, 8, 2, 1]
    x = {'a': torch.zeros(T, len(lengths), 4, device='cuda'), 'b': torch.zeros(T, len(lengths), 1, device='cuda')}
    fn = multiseq.rand_delete if mode == 0 else multiseq.burst_delete
    got = fn(x, 0.3, lengths, seed=99, b_offset=5)
    for idx, m in enumerate(x):
        want = orc.draw_deletions(lengths, T, len(lengths), 0.3, mode, 99, stream_id=idx, b_offset=5)
        flags = torch.isnan(got[m]).all(dim=-1).cpu().numpy().astype(np.uint8)
        assert np.array_equal(flags, want)
        assert torch.isnan(got[m]).any(dim=-1).cpu().numpy().astype(np.uint8).tolist() == want.tolist()


def test_c2_shape_properties():
    """B = 4096, T = 100 (BASELINE configs[1]): exact deletion counts, untouched survivors, bursts
    contiguous, keep_segment o rand_delete composition as Trainer.evaluate applies it."""
    T, B = 100, 4096
    g = torch.Generator(device='cuda').manual_seed(0)
    x = {m: torch.randn(T, B, 1, device='cuda', generator=g) for m in ('spiral-x', 'spiral-y')}
    lengths = [T] * B
    r = multiseq.rand_delete(x, 0.5, lengths, seed=1)
    for m in x:
        nan = torch.isnan(r[m][..., 0])
        assert torch.equal(nan.sum(0), torch.full((B,), 50, device='cuda'))
        assert torch.equal(r[m][~nan.unsqueeze(-1)], x[m][~nan.unsqueeze(-1)])
    assert not torch.equal(torch.isnan(r['spiral-x']), torch.isnan(r['spiral-y']))
    b = multiseq.burst_delete(x, 0.1, lengths, seed=2)
    nan = torch.isnan(b['spiral-x'][..., 0]).int()
    assert int(nan.sum(0).max()) == 10 and int(nan.sum(0).min()) >= 1
    assert int((nan[1:] - nan[:-1]).abs().sum(0).max()) <= 2                  # one contiguous run per sequence
    k = multiseq.keep_segment(r, 0.25, 0.75, lengths)
    nan_k = torch.isnan(k['spiral-x'][..., 0])
    assert bool(nan_k[:25].all()) and bool(nan_k[75:].all())
    assert torch.equal(nan_k[25:75], torch.isnan(r['spiral-x'][25:75, :, 0]))
    # numpy-replay mode at the same size equals the restatement bit for bit
    np.random.seed(3)
    got = multiseq.burst_delete(x, 0.1, lengths)
    np.random.seed(3)
    want = orc.burst_delete({m: v.cpu().numpy() for m, v in x.items()}, 0.1, lengths)
    assert all(same(got[m].cpu().numpy(), want[m]) for m in x)


def test_image_rows_use_the_vector_path():
    T, B = 6, 5
    x = {'video': torch.rand(T, B, 3, 8, 8, device='cuda')}
    np.random.seed(4)
    got = multiseq.rand_delete(x, 0.5, [6, 6, 5, 3, 2])
    np.random.seed(4)
    want = orc.rand_delete({'video': x['video'].cpu().numpy()}, 0.5, [6, 6, 5, 3, 2])
    assert same(got['video'].cpu().numpy(), want['video'])


EVAL = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'multiseq', 'eval.pt'), weights_only=False)


@pytest.mark.parametrize('case', EVAL, ids=lambda c: c['name'])
def test_eval_outputs_match_reference_golden(case):
    """seq_decoll_dict (bit for bit) and the per-sequence MSE metric against the reference's outputs."""
    dev = torch.device('cuda:0')
    batch = {m: v.to(dev) for m, v in case['batch'].items()}
    got = multiseq.seq_decoll_dict(batch, case['lengths'], case['order'])
    for m, want in case['decoll'].items():
        assert len(got[m]) == len(want)
        assert all(same(a, b.numpy()) for a, b in zip(got[m], want)), m
    recon = {m: (v.to(dev), None) for m, v in case['recon'].items()}
    targets = {m: v.to(dev) for m, v in case['targets'].items()}
    mse = multiseq.seq_mse(recon, targets, case['mask'].to(dev), case['lengths'], case['order'])
    assert torch.allclose(mse.cpu(), case['mse'], rtol=2e-6, atol=0)


def test_eval_outputs_at_bench_size():
    """C2-sized batch: decollation round trip (collate -> decollate returns the sequences) and the MSE
    metric against torch ops on the device."""
    dev = torch.device('cuda:0')
    rng = np.random.RandomState(3)
    lengths = sorted(rng.randint(1, 101, size=4096).tolist(), reverse=True)
    seqs = [rng.standard_normal((n, 3)).astype(np.float32) for n in lengths]
    x = multiseq.pad_and_merge(seqs, device=dev)
    order = rng.permutation(len(lengths)).tolist()
    out = multiseq.seq_decoll(x, lengths, order)
    assert all(np.array_equal(out[j], seqs[i]) for j, i in enumerate(order))
    mask = multiseq.len_to_mask(lengths, device=dev)
    tgt = torch.nan_to_num(x, nan=0.0)
    rec = tgt + 0.1 * torch.randn_like(tgt)
    mse = multiseq.seq_mse({'a': (rec, None)}, {'a': tgt}, mask, lengths)
    ref = ((rec - tgt) ** 2).sum(-1) * mask.squeeze(-1)
    ref = ref.sum(0) / torch.tensor(lengths, dtype=torch.float32, device=dev)
    assert torch.allclose(mse, ref, rtol=1e-5, atol=0)
