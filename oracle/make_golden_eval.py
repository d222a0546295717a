"""Pins the evaluation-output restatements of oracle/multiseq_oracle.py (seq_decoll, seq_mse) against the
UNMODIFIED reference: datasets/multiseq.py:388-403 (imported from /root/reference) and the 'mse' lines of
SpiralsTrainer.compute_metrics (spirals.py:105-111, executed verbatim on torch tensors under the documented
`1 - bool` shim — spirals.py itself cannot be imported here because matplotlib is absent).
Stores the reference's outputs: tests/golden/multiseq/eval.pt.   python oracle/make_golden_eval.py"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')
import ref_shim  # noqa: E402
ref_shim.install()
warnings.filterwarnings('ignore', category=SyntaxWarning)
import datasets.multiseq as ref  # noqa: E402
import multiseq_oracle as orc    # noqa: E402


def reference_mse(recon, targets, mask, lengths, order):
    """spirals.py:98-111, verbatim apart from the device argument."""
    if type(lengths) != torch.Tensor:
        lengths = torch.FloatTensor(lengths)
    mse = sum([(recon[m][0] - targets[m]).pow(2) for m in list(recon.keys())])
    mse = mse.sum(dim=list(range(2, mse.dim())))

    def time_avg(val):
        val[1 - mask.squeeze(-1)] = 0.0
        return val.sum(dim=0) / lengths
    return time_avg(mse)[order].tolist()


def main():
    rng = np.random.RandomState(11)
    cases = []
    for name, T, lengths, dims in (('spirals', 20, [20, 17, 17, 9, 4, 1], {'spiral-x': (1,), 'spiral-y': (1,)}),
                                   ('wide', 9, [9, 9, 6, 2], {'a': (5,), 'b': (5,)}),
                                   ('image', 6, [6, 4, 4], {'video': (3, 4, 4)}),
                                   ('single', 7, [7], {'a': (2,)})):
        B = len(lengths)
        order = list(rng.permutation(B))
        batch = {m: rng.standard_normal((T, B) + d).astype(np.float32) for m, d in dims.items()}
        for m in batch:
            for b, n in enumerate(lengths):
                batch[m][n:, b] = np.nan
        tb = {m: torch.from_numpy(v.copy()) for m, v in batch.items()}
        want = ref.seq_decoll_dict(tb, lengths, order)
        got = {m: orc.seq_decoll(v, lengths, order) for m, v in batch.items()}
        for m in batch:
            assert all(np.array_equal(a, b) for a, b in zip(want[m], got[m]))
        tup = tuple(tb.values())
        want_t = ref.seq_decoll(tup, lengths, order) if len({v.shape for v in tup}) == 1 else None
        if want_t is not None:
            got_t = orc.seq_decoll(tuple(batch.values()), lengths, order)
            assert all(np.array_equal(a, b) for a, b in zip(want_t, got_t))
        # metrics: reconstruction = target + noise on the valid steps, arbitrary finite values on the padding
        targets = {m: np.nan_to_num(v, nan=0.0).astype(np.float32) for m, v in batch.items()}
        recon = {m: (v + 0.3 * rng.standard_normal(v.shape)).astype(np.float32) for m, v in targets.items()}
        mask = ref.len_to_mask(lengths)
        want_mse = reference_mse({m: (torch.from_numpy(v.copy()), None) for m, v in recon.items()},
                                 {m: torch.from_numpy(v.copy()) for m, v in targets.items()}, mask, lengths, order)
        got_mse = orc.seq_mse(recon, targets, mask.numpy().astype(bool), lengths, order)
        assert np.allclose(want_mse, got_mse, rtol=2e-6, atol=0), (want_mse, got_mse)
        cases.append({'name': name, 'lengths': lengths, 'order': [int(i) for i in order],
                      'batch': tb, 'decoll': {m: [torch.from_numpy(np.ascontiguousarray(a)) for a in want[m]] for m in want},
                      'decoll_tuple': None if want_t is None else [torch.from_numpy(np.ascontiguousarray(a)) for a in want_t],
                      'recon': {m: torch.from_numpy(v) for m, v in recon.items()},
                      'targets': {m: torch.from_numpy(v) for m, v in targets.items()},
                      'mask': mask.bool(), 'mse': torch.tensor(want_mse, dtype=torch.float32)})
    out = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'multiseq', 'eval.pt')
    torch.save(cases, out)
    print('wrote', out, len(cases), 'cases')


if __name__ == '__main__':
    main()
