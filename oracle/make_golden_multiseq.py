"""Pins oracle/multiseq_oracle.py against the UNMODIFIED reference functions
(datasets/multiseq.py:321-353, 405-448, imported from /root/reference) and stores the
reference's outputs as fixtures: tests/golden/multiseq/cases.pt.  Run in the build container
(the GPU box has no /root/reference):  python oracle/make_golden_multiseq.py"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')
import ref_shim  # noqa: E402,F401  (1 - bool_tensor shim, see oracle/ref_shim.py)
warnings.filterwarnings('ignore', category=SyntaxWarning)
import datasets.multiseq as ref  # noqa: E402
import multiseq_oracle as orc    # noqa: E402


def make_batch(rng, T, lengths, dims):
    batch = {}
    for m, d in dims.items():
        x = rng.standard_normal((T, len(lengths)) + d).astype(np.float32)
        for b, n in enumerate(lengths):
            x[n:, b] = np.nan
        batch[m] = x
    return batch


def same(a, b):
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


def main():
    rng = np.random.RandomState(7)
    cases = []
    shapes = [
        ('spirals_ragged', 20, [20, 17, 17, 9, 4, 1], {'spiral-x': (1,), 'spiral-y': (1,)}),
        ('full_lengths', 12, None, {'a': (4,), 'b': (3,)}),
        ('image', 8, [8, 8, 5], {'video': (3, 4, 4), 'action': (1,)}),
    ]
    ops = [('burst_delete', (0.1,)), ('burst_delete', (0.5,)), ('rand_delete', (0.5,)), ('rand_delete', (0.0,)),
           ('keep_segment', (0.25, 0.75)), ('del_segment', (0.2, 0.6)), ('keep_segment', (0.0, 1.0))]
    for name, T, lengths, dims in shapes:
        B = 5 if lengths is None else len(lengths)
        batch = make_batch(rng, T, lengths if lengths is not None else [T] * B, dims)
        for op, args in ops:
            for mods in (None, [list(dims)[0]]):
                seed = int(rng.randint(1 << 30))
                np.random.seed(seed)
                want = getattr(ref, op)({m: torch.from_numpy(v.copy()) for m, v in batch.items()}, *args,
                                        lengths=lengths, modalities=mods)
                want = {m: v.numpy() for m, v in want.items()}
                np.random.seed(seed)
                got = getattr(orc, op)(batch, *args, lengths=lengths, modalities=mods)
                assert all(same(want[m], got[m]) for m in batch), (name, op, args)
                cases.append({'name': name, 'op': op, 'args': args, 'lengths': lengths, 'modalities': mods,
                              'np_seed': seed, 'inputs': {m: torch.from_numpy(v) for m, v in batch.items()},
                              'outputs': {m: torch.from_numpy(v) for m, v in want.items()}})
    # collation
    coll = []
    for lengths, dims in (([9, 7, 7, 2], (2,)), ([5], (3,)), ([6, 6, 3], (2, 2, 2))):
        seqs = [rng.standard_normal((n,) + dims).astype(np.float32) for n in lengths]
        for max_len in (None, max(lengths) + 2):
            want = ref.pad_and_merge(seqs, max_len).numpy().astype(np.float32)
            assert same(want, orc.pad_and_merge(seqs, max_len))
            coll.append({'sequences': [torch.from_numpy(s) for s in seqs], 'max_len': max_len,
                         'output': torch.from_numpy(want)})
        m_ref = ref.len_to_mask(lengths).numpy()
        assert np.array_equal(m_ref.astype(bool), orc.len_to_mask(lengths))
        coll[-1]['mask'] = torch.from_numpy(m_ref.astype(bool))
        coll[-1]['lengths'] = lengths
    out = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'multiseq', 'cases.pt')
    torch.save({'delete': cases, 'collate': coll}, out)
    print('wrote', out, len(cases), 'delete cases,', len(coll), 'collate cases')


if __name__ == '__main__':
    main()
