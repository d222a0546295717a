"""Batch preparation on CPU: (1) the numpy restatement oracle/multiseq_oracle.py
against the golden fixtures made from the unmodified reference (oracle/make_golden_multiseq.py);
(2) the library's kernels (emulated) behind multimodal_dmm_b200.multiseq against the same
fixtures, bit for bit, replaying the reference's numpy draws; (3) the seeded device draw rule
against its integer-exact restatement and its distributional contract."""
import os

import numpy as np
import pytest
import torch

import helpers
import multiseq_oracle as orc
from multimodal_dmm_b200 import _lib, multiseq

GOLD = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'multiseq', 'cases.pt'), weights_only=False)


class EmuRuntime:
    """multiseq._Runtime bound to the emulated kernels and CPU tensors."""

    def __init__(self, device):
        self.lib, self.device = helpers.emu_library(), torch.device('cpu')

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def call(self, name, *args):
        self.lib.call(name, *args, None)


@pytest.fixture
def emu(monkeypatch):
    monkeypatch.setattr(multiseq, '_Runtime', EmuRuntime)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


def case_id(c):
    return '%s-%s%s-%s' % (c['name'], c['op'], c['args'], 'all' if c['modalities'] is None else 'one')


@pytest.mark.parametrize('case', GOLD['delete'], ids=case_id)
def test_oracle_restatement_matches_reference_golden(case):
    np.random.seed(case['np_seed'])
    got = getattr(orc, case['op'])({m: v.numpy() for m, v in case['inputs'].items()}, *case['args'],
                                   lengths=case['lengths'], modalities=case['modalities'])
    assert all(same(got[m], case['outputs'][m].numpy()) for m in got)


@pytest.mark.parametrize('case', GOLD['delete'], ids=case_id)
def test_kernels_match_reference_golden(emu, case):
    np.random.seed(case['np_seed'])
    inputs = {m: v.clone() for m, v in case['inputs'].items()}
    got = getattr(multiseq, case['op'])(inputs, *case['args'], lengths=case['lengths'], modalities=case['modalities'])
    for m in inputs:
        assert same(got[m].numpy(), case['outputs'][m].numpy()), m
        assert same(inputs[m].numpy(), case['inputs'][m].numpy())          # inputs are not mutated
        assert got[m].data_ptr() != inputs[m].data_ptr()
    # the numpy stream was consumed exactly like the reference consumes it
    after = np.random.randint(1 << 30)
    np.random.seed(case['np_seed'])
    getattr(orc, case['op'])({m: v.numpy() for m, v in case['inputs'].items()}, *case['args'],
                             lengths=case['lengths'], modalities=case['modalities'])
    assert after == np.random.randint(1 << 30)


@pytest.mark.parametrize('i', range(len(GOLD['collate'])))
def test_collation_matches_reference_golden(emu, i):
    c = GOLD['collate'][i]
    seqs = [s.numpy() for s in c['sequences']]
    assert same(orc.pad_and_merge(seqs, c['max_len']), c['output'].numpy())
    assert same(multiseq.pad_and_merge(seqs, c['max_len']).numpy(), c['output'].numpy())
    if 'mask' in c:
        got = multiseq.len_to_mask(c['lengths'])
        assert got.dtype == torch.bool and torch.equal(got, c['mask'])
        assert torch.equal(multiseq.len_to_mask(c['lengths'], time_first=False), c['mask'].transpose(0, 1))


def test_seq_collate_dict_matches_reference_layout(emu):
    rng = np.random.RandomState(0)
    lens = [3, 7, 5, 7]
    data = [{'a': rng.randn(n, 2).astype(np.float32), 'b': rng.randn(n, 1).astype(np.float32), 'length': n, 'id': 10 + i}
            for i, n in enumerate(lens)]
    want_order = sorted(range(4), key=lambda i: lens[i], reverse=True)
    raw = [dict(d) for d in data]
    batch, mask, lengths, order, ids = multiseq.seq_collate_dict(data)
    assert lengths == [7, 7, 5, 3] and order == want_order and ids == [10 + i for i in want_order]
    for m in ('a', 'b'):
        assert same(batch[m].numpy(), orc.pad_and_merge([raw[i][m] for i in want_order], 7))
    assert torch.equal(mask, torch.from_numpy(orc.len_to_mask(lengths)))


def test_func_delete_accepts_any_del_func(emu):
    x = {'a': torch.arange(24, dtype=torch.float32).reshape(6, 2, 2)}
    got = multiseq.func_delete(x, lambda n: [0, n - 1], lengths=[6, 4])
    want = orc.func_delete({'a': x['a'].numpy()}, lambda n: [0, n - 1], lengths=[6, 4])
    assert same(got['a'].numpy(), want['a'])


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('T,lengths,frac', [(20, [20, 17, 9, 4, 1, 20, 13], 0.5), (50, None, 0.1), (7, [7, 7], 1.0),
                                            (9, [9, 3], 0.0)])
def test_seeded_draws_match_integer_restatement(mode, T, lengths, frac):
    lib = helpers.emu_library()
    B = 5 if lengths is None else len(lengths)
    flags = torch.full((T, B), 9, dtype=torch.uint8)
    len_t = None if lengths is None else torch.tensor(lengths, dtype=torch.int32)
    lib.call('bfvi_draw_deletions', _lib.ptr(len_t), T, B, float(frac), mode, 0x1234567890, 3, 11, _lib.ptr(flags), None)
    want = orc.draw_deletions(lengths, T, B, frac, mode, 0x1234567890, stream_id=3, b_offset=11)
    assert np.array_equal(flags.numpy(), want)
    # contract: exactly int(frac * length) deleted steps inside the sequence (burst: clipped, contiguous)
    for b in range(B):
        n = T if lengths is None else lengths[b]
        col = flags[:, b].numpy()
        assert col[n:].sum() == 0
        if mode == 0:
            assert col.sum() == int(frac * n)
        else:
            idx = np.flatnonzero(col)
            assert len(idx) <= int(frac * n) and (len(idx) == 0 or idx[-1] - idx[0] + 1 == len(idx))
            assert len(idx) == int(frac * n) or (len(idx) > 0 and idx[-1] == n - 1) or int(frac * n) == 0


def test_seeded_uniform_deletion_is_unbiased():
    """every time step is deleted with probability k / length (selection sampling)"""
    lib = helpers.emu_library()
    T, B = 10, 4000
    flags = torch.empty(T, B, dtype=torch.uint8)
    lib.call('bfvi_draw_deletions', None, T, B, 0.3, 0, 77, 0, 0, _lib.ptr(flags), None)
    freq = flags.float().mean(dim=1).numpy()
    assert np.all(np.abs(freq - 0.3) < 4 * np.sqrt(0.3 * 0.7 / B))
    starts = torch.empty(T, B, dtype=torch.uint8)
    lib.call('bfvi_draw_deletions', None, T, B, 0.1, 1, 77, 0, 0, _lib.ptr(starts), None)   # 1-step bursts = t_start
    hist = starts.float().mean(dim=1).numpy()
    assert np.all(np.abs(hist - 0.1) < 4 * np.sqrt(0.1 * 0.9 / B))


def test_seeded_api_shards_like_the_whole_batch(emu):
    """data parallel: rank-local batches with b_offset reproduce the global batch's deletions"""
    g = torch.Generator().manual_seed(0)
    x = {'a': torch.randn(12, 6, 3, generator=g), 'b': torch.randn(12, 6, 1, generator=g)}
    lengths = [12, 12, 10, 8, 5, 2]
    whole = multiseq.rand_delete(x, 0.5, lengths, seed=5)
    for lo, hi in ((0, 3), (3, 6)):
        part = multiseq.rand_delete({m: v[:, lo:hi].contiguous() for m, v in x.items()}, 0.5, lengths[lo:hi], seed=5,
                                    b_offset=lo)
        for m in x:
            assert same(part[m].numpy(), whole[m][:, lo:hi].numpy())
    da, db = torch.isnan(whole['a'][..., 0]), torch.isnan(whole['b'][..., 0])
    assert not torch.equal(da, db)                       # modalities draw independent streams


def test_entry_points_reject_bad_arguments():
    lib = helpers.emu_library()
    t = torch.zeros(4)
    flags = torch.zeros(4, dtype=torch.uint8)
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_delete_rows', None, _lib.ptr(flags), 2, 2, 1, _lib.ptr(t), None)
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_draw_deletions', None, 2, 2, 1.5, 0, 1, 0, 0, _lib.ptr(flags), None)
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_draw_deletions', None, 2, 2, 0.5, 7, 1, 0, 0, _lib.ptr(flags), None)
    with pytest.raises(_lib.BfviError):
        multiseq.burst_delete({'a': torch.zeros(3, 2, 1)}, 0.1)            # CPU tensor, real runtime


# ---- evaluation outputs: seq_decoll / seq_decoll_dict and the per-sequence MSE metric ----
EVAL = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'multiseq', 'eval.pt'), weights_only=False)


@pytest.mark.parametrize('case', EVAL, ids=lambda c: c['name'])
def test_eval_oracle_matches_reference_golden(case):
    for m, want in case['decoll'].items():
        got = orc.seq_decoll(case['batch'][m].numpy(), case['lengths'], case['order'])
        assert len(got) == len(want) and all(same(a, b.numpy()) for a, b in zip(got, want))
    mse = orc.seq_mse({m: v.numpy() for m, v in case['recon'].items()}, {m: v.numpy() for m, v in case['targets'].items()},
                      case['mask'].numpy(), case['lengths'], case['order'])
    assert np.allclose(mse, case['mse'].numpy(), rtol=2e-6, atol=0)


@pytest.mark.parametrize('case', EVAL, ids=lambda c: c['name'])
def test_eval_kernels_match_reference_golden(emu, case):
    got = multiseq.seq_decoll_dict(case['batch'], case['lengths'], case['order'])
    for m, want in case['decoll'].items():
        assert len(got[m]) == len(want)
        assert all(same(a, b.numpy()) for a, b in zip(got[m], want)), m           # bit for bit
    if case['decoll_tuple'] is not None:
        tup = multiseq.seq_decoll(tuple(case['batch'].values()), case['lengths'], case['order'])
        assert all(same(a, b.numpy()) for a, b in zip(tup, case['decoll_tuple']))
    # batch-first layout gives the same sequences
    bf = multiseq.seq_decoll(case['batch'][list(case['batch'])[0]].transpose(0, 1), case['lengths'], case['order'],
                             time_first=False)
    assert all(same(a, b.numpy()) for a, b in zip(bf, case['decoll'][list(case['batch'])[0]]))
    recon = {m: (v, None) for m, v in case['recon'].items()}
    mse = multiseq.seq_mse(recon, case['targets'], case['mask'], case['lengths'], case['order'])
    assert torch.allclose(mse, case['mse'], rtol=2e-6, atol=0)                    # fp32 sums in another order
    with pytest.raises(_lib.BfviError):
        multiseq.seq_mse({'a': torch.zeros(3, 2, 1)}, {'a': torch.zeros(3, 2, 2)}, torch.ones(3, 2, 1, dtype=torch.bool), [3, 3])
