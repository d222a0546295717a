"""CPU restatement (numpy) of the reference's batch preparation functions,
datasets/multiseq.py:321-353 and 405-448 — TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product path (multimodal-dmm_b200/multiseq.py calls the CUDA kernels).

Parity status: PINNED — oracle/make_golden_multiseq.py runs the unmodified reference functions
and this restatement on the same seeded inputs, asserts bit equality and stores the
reference's outputs under tests/golden/multiseq/.

The second half restates the library's OWN seeded draw rule (bfvi_draw_deletions,
csrc/bfvi_data.cuh): Philox4x32-10 words and integer-exact selection sampling; it has no
reference counterpart (the reference draws from numpy's global generator)."""
import numpy as np

NAN = np.float32('nan')


def len_to_mask(lengths):
    """datasets/multiseq.py:321-327, time first: (T, B, 1) bool."""
    t = np.arange(max(lengths))[:, None]
    return (t < np.asarray(lengths)[None, :])[..., None]


def pad_and_merge(sequences, max_len=None):
    """datasets/multiseq.py:342-353."""
    dims = sequences[0].shape[1:]
    lengths = [len(s) for s in sequences]
    if max_len is None:
        max_len = max(lengths)
    out = np.full((max_len, len(sequences)) + tuple(dims), NAN, dtype=np.float32)
    for i, s in enumerate(sequences):
        out[:lengths[i], i] = s[:lengths[i]]
    return out


def seq_decoll(batch, lengths, order, time_first=True):
    """datasets/multiseq.py:388-398 (numpy in, list of numpy out)."""
    if isinstance(batch, tuple):
        return [np.stack([b[:lengths[idx], idx] for b in batch], axis=1) for idx in order]
    if time_first:
        return [batch[:lengths[idx], idx] for idx in order]
    return [batch[idx, :lengths[idx]] for idx in order]


def seq_mse(recon_means, targets, mask, lengths, order):
    """The 'mse' metric of spirals.py:105-111: float32 like the reference tensors."""
    mse = sum([(recon_means[m] - targets[m]) ** 2 for m in recon_means])
    mse = mse.reshape(mse.shape[0], mse.shape[1], -1).sum(axis=2, dtype=np.float32)
    mse[~mask.reshape(mask.shape[0], mask.shape[1])] = 0.0
    return (mse.sum(axis=0, dtype=np.float32) / np.asarray(lengths, dtype=np.float32))[list(order)]


def func_delete(batch_in, del_func, lengths=None, modalities=None):
    """datasets/multiseq.py:405-420 (draw order: modality-major, then b)."""
    if modalities is None:
        modalities = list(batch_in.keys())
    out = {}
    for m in batch_in.keys():
        out[m] = np.array(batch_in[m], dtype=np.float32, copy=True)
        if m not in modalities:
            continue
        t_max, b_dim = out[m].shape[:2]
        if lengths is None:
            lengths = [t_max] * b_dim
        for b in range(b_dim):
            out[m][del_func(lengths[b]), b] = NAN
    return out


def rand_delete(batch_in, del_frac, lengths=None, modalities=None):
    """datasets/multiseq.py:422-426."""
    return func_delete(batch_in, lambda n: np.random.choice(n, int(del_frac * n), False), lengths, modalities)


def burst_delete(batch_in, burst_frac, lengths=None, modalities=None):
    """datasets/multiseq.py:428-434."""
    def del_func(n):
        t0 = np.random.randint(n)
        return list(range(t0, min(t0 + int(burst_frac * n), n)))
    return func_delete(batch_in, del_func, lengths, modalities)


def keep_segment(batch_in, f_start, f_stop, lengths=None, modalities=None):
    """datasets/multiseq.py:436-441."""
    return func_delete(batch_in, lambda n: list(range(0, int(f_start * n))) + list(range(int(f_stop * n), n)),
                       lengths, modalities)


def del_segment(batch_in, f_start, f_stop, lengths=None, modalities=None):
    """datasets/multiseq.py:443-448."""
    return func_delete(batch_in, lambda n: list(range(int(f_start * n), int(f_stop * n))), lengths, modalities)


# ---- the library's seeded draw rule (csrc/bfvi_data.cuh: draw_deletions_kernel) -----------------
def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (2,) uint32 -> (..., 4) uint32 (Salmon et al. 2011)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]), int(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & mask,
             (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & mask]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def draw_deletions(lengths, T, B, frac, mode, seed, stream_id=0, b_offset=0):
    """(T, B) uint8 deletion flags: mode 0 = exactly int(frac * length) steps by selection
    sampling, mode 1 = one burst; same integer rules as the kernel."""
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    flags = np.zeros((T, B), dtype=np.uint8)
    for b in range(B):
        n = min(T if lengths is None else int(lengths[b]), T)
        k = int(frac * n)
        if mode == 1:
            r = philox4x32_10(np.array([b + b_offset, 0, 1, stream_id], dtype=np.uint32), key)
            t0 = (int(r[0]) * n) >> 32
            flags[t0:min(t0 + k, n), b] = 1
            continue
        need = k
        ctr = np.zeros(((T + 3) // 4, 4), dtype=np.uint32)
        ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3] = b + b_offset, np.arange((T + 3) // 4), 0, stream_id
        words = philox4x32_10(ctr, key).reshape(-1)
        for t in range(n):
            if need > 0 and int(words[t]) * (n - t) < (need << 32):
                flags[t, b] = 1
                need -= 1
    return flags
