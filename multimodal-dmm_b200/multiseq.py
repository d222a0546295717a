"""Device-side batch preparation with the names and signatures of the reference's
`datasets/multiseq.py` batch functions (:321-353, 405-448) — what `Trainer.train` /
`Trainer.evaluate` call on every batch right before `model.step` / `model(...)`
(trainer.py:235, 284-287).

The reference loops over the batch in Python (one numpy draw and one indexed device write per
sequence and modality: ~8 000 tiny ops per C2 batch).  Here each call is ONE kernel of
libbfvi_b200 per modality (`bfvi_delete_rows / bfvi_delete_spans`, csrc/bfvi_data.cuh) and the
random index sets come from either

* `seed=None` (default): the reference's own numpy draws, made on the host in the reference's
  order from the global `np.random` state — results are bit-identical to the reference for a
  seeded run (tests/test_*_multiseq.py); only the draws stay on the host, or
* `seed=<int>`: the library's Philox stream (`bfvi_draw_deletions`): no host work, no H2D; a
  different (documented, integer-exact) stream than numpy's, same distribution.

CUDA tensors only; there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

DELETE_UNIFORM, DELETE_BURST = 0, 1


class _Runtime:
    """Library handle + launch stream + device guard for one CUDA device.  The CPU tests swap
    this class for one bound to the emulated kernels (tests/test_emu_multiseq.py); the product
    has no other implementation and refuses CPU tensors."""

    def __init__(self, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise _lib.BfviError('multiseq batch functions run on CUDA tensors only (no CPU fallback)')
        self.lib = _lib.load()
        self.device = device
        self.guard = torch.cuda.device(device)

    def __enter__(self):
        self.guard.__enter__()
        return self

    def __exit__(self, *exc):
        return self.guard.__exit__(*exc)

    def call(self, name, *args):
        self.lib.call(name, *args, C.c_void_p(torch.cuda.current_stream().cuda_stream))


def _i32(values, device):
    return torch.as_tensor(np.asarray(values, dtype=np.int32)).to(device, non_blocking=True)


def len_to_mask(lengths, time_first=True, device='cuda'):
    """List of sequence lengths -> (T, B, 1) bool mask (datasets/multiseq.py:321-327)."""
    T, B = int(max(lengths)), len(lengths)
    with _Runtime(device) as rt:
        mask = torch.empty(T, B, dtype=torch.uint8, device=rt.device)
        len_d = _i32(lengths, rt.device)
        rt.call('bfvi_len_to_mask', _lib.ptr(len_d), T, B, _lib.ptr(mask))
    mask = mask.bool()
    return (mask if time_first else mask.transpose(0, 1)).unsqueeze(-1)


def pad_and_merge(sequences, max_len=None, device='cuda'):
    """Unequal-length numpy sequences -> (T, B, D...) fp32 batch padded with NaN
    (datasets/multiseq.py:342-353): one host concatenation, one H2D copy, one kernel."""
    dims = tuple(sequences[0].shape[1:])
    lengths = [len(s) for s in sequences]
    T = int(max(lengths) if max_len is None else max_len)
    B, D = len(sequences), int(np.prod(dims)) if dims else 1
    kept = [min(n, T) for n in lengths]
    packed = np.concatenate([np.asarray(s[:n], dtype=np.float32).reshape(n, D) for s, n in zip(sequences, kept)], 0)
    starts = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(kept, out=starts[1:])
    with _Runtime(device) as rt:
        out = torch.empty((T, B) + dims, dtype=torch.float32, device=rt.device)
        if packed.shape[0] == 0:
            return out.fill_(float('nan'))
        packed_d = torch.from_numpy(packed).to(rt.device, non_blocking=True)
        starts_d = torch.from_numpy(starts).to(rt.device, non_blocking=True)
        rt.call('bfvi_pad_merge', _lib.ptr(packed_d), _lib.ptr(starts_d), T, B, D, _lib.ptr(out))
    return out


def seq_collate_dict(data, time_first=True, device='cuda'):
    """Collate a list of {modality: array, 'length', 'id'} into padded device batches
    (datasets/multiseq.py:373-386): returns (batch, mask, lengths, order, seq_ids)."""
    modalities = [k for k in data[0] if k not in ['length', 'id']]
    order = sorted(range(len(data)), key=lambda i: data[i]['length'], reverse=True)
    data.sort(key=lambda d: d['length'], reverse=True)
    lengths = [d['length'] for d in data]
    seq_ids = [d['id'] for d in data]
    batch = {}
    for m in modalities:
        padded = pad_and_merge([d[m] for d in data], max(lengths), device=device)
        batch[m] = padded if time_first else padded.transpose(0, 1)
    return batch, len_to_mask(lengths, time_first, device=device), lengths, order, seq_ids


def seq_decoll(batch, lengths, order, time_first=True):
    """De-pad and reorder a device batch into a list of numpy arrays, one per sequence
    (datasets/multiseq.py:388-398): sequence j of the result is batch column `order[j]`, its first
    `lengths[order[j]]` steps.  One gather kernel into a packed buffer and ONE device-to-host copy
    (the reference issues one indexed read and one D2H copy per sequence).  A tuple of batches is
    stacked along a new axis 1 per sequence, as in the reference."""
    if isinstance(batch, tuple):
        parts = [seq_decoll(b, lengths, order, time_first) for b in batch]
        return [np.stack([p[j] for p in parts], axis=1) for j in range(len(order))]
    x = batch if time_first else batch.transpose(0, 1)
    T, B = x.shape[:2]
    dims = tuple(x.shape[2:])
    D = int(np.prod(dims)) if dims else 1
    order = [int(i) for i in order]
    kept = [min(int(lengths[i]), T) for i in order]
    starts = np.zeros(len(order) + 1, dtype=np.int64)
    np.cumsum(kept, out=starts[1:])
    total = int(starts[-1])
    with _Runtime(x.device) as rt:
        xc = x.detach().contiguous().float()
        packed = torch.empty((total, D), dtype=torch.float32, device=xc.device)
        if total > 0:
            starts_d = torch.from_numpy(starts).to(xc.device, non_blocking=True)
            order_d = _i32(order, xc.device)
            rt.call('bfvi_unpad', _lib.ptr(xc), _lib.ptr(starts_d), _lib.ptr(order_d), B, len(order), total, D,
                    _lib.ptr(packed))
        host = packed.cpu().numpy()
    return [host[starts[j]:starts[j + 1]].reshape((kept[j],) + dims) for j in range(len(order))]


def seq_decoll_dict(batch_dict, lengths, order, time_first=True):
    """Dictionary of batch tensors -> dictionary of per-sequence lists (datasets/multiseq.py:400-403)."""
    return {k: seq_decoll(batch, lengths, order, time_first) for k, batch in batch_dict.items()}


def seq_mse(recon, targets, mask, lengths, order=None):
    """The per-sequence MSE of the evaluation metrics (spirals.py:105-111): squared error of the
    reconstructed means summed over modalities and features, zero outside the sequence mask, averaged over
    each sequence's length; `recon[m]` is the decoder's tuple (mean, std) or a mean tensor.  Returns a (B,)
    device tensor, indexed by `order` when given — `.tolist()` of it is `metrics['mse']`."""
    mods = list(recon.keys())
    means = [(recon[m][0] if isinstance(recon[m], (tuple, list)) else recon[m]).detach().contiguous().float() for m in mods]
    tgts = [targets[m].detach().contiguous().float() for m in mods]
    T, B = means[0].shape[:2]
    with _Runtime(means[0].device) as rt:
        dev = means[0].device
        mk = mask.reshape(T, B).to(device=dev, dtype=torch.uint8).contiguous()
        if not torch.is_tensor(lengths):
            lengths = torch.as_tensor(np.asarray(lengths, dtype=np.float32))
        len_d = lengths.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(B, dtype=torch.float32, device=dev)
        n = len(mods)
        rp = (C.c_void_p * n)(*[t.data_ptr() for t in means])
        tp = (C.c_void_p * n)(*[t.data_ptr() for t in tgts])
        dims = (C.c_int64 * n)(*[int(t[0, 0].numel()) for t in means])
        for a, b in zip(means, tgts):
            if a.shape != b.shape:
                raise _lib.BfviError('seq_mse: reconstruction %s and target %s differ in shape' % (tuple(a.shape), tuple(b.shape)))
        n_split = int(rt.lib.dll.bfvi_seq_mse_splits(T, B))
        scratch = torch.empty(B * n_split, dtype=torch.float32, device=dev) if n_split > 1 else None
        rt.call('bfvi_seq_mse', rp, tp, dims, n, _lib.ptr(mk), _lib.ptr(len_d), T, B, _lib.ptr(out), _lib.ptr(scratch),
                n_split)
    return out if order is None else out[torch.as_tensor(list(order), device=out.device)]


def _rows_view(x):
    T, B = x.shape[:2]
    D = x[0, 0].numel() if x.dim() > 2 else 1
    return T, B, int(D)


def _apply(batch_in, modalities, launch):
    """Shared skeleton of func_delete (datasets/multiseq.py:405-420): every modality is copied,
    the listed ones through `launch(runtime, modality index, x, out)`."""
    if modalities is None:
        modalities = list(batch_in.keys())
    batch_out = {}
    for idx, m in enumerate(batch_in.keys()):
        x = batch_in[m]
        with _Runtime(x.device) as rt:
            if m not in modalities or x.numel() == 0:
                batch_out[m] = x.clone().detach()
                continue
            xc = x.detach().contiguous().float()
            out = torch.empty_like(xc)
            launch(rt, idx, xc, out)
            batch_out[m] = out
    return batch_out


def func_delete(batch_in, del_func, lengths=None, modalities=None):
    """`del_func(length)` -> time indices to delete, per sequence and modality, evaluated on the
    host in the reference's order; applied by one kernel per modality."""
    state = {'lengths': lengths}

    def launch(rt, idx, x, out):
        T, B, D = _rows_view(x)
        if state['lengths'] is None:
            state['lengths'] = [T] * B
        flags = np.zeros((T, B), dtype=np.uint8)
        for b in range(B):
            flags[del_func(state['lengths'][b]), b] = 1
        flags_d = torch.from_numpy(flags).to(x.device, non_blocking=True)
        rt.call('bfvi_delete_rows', _lib.ptr(x), _lib.ptr(flags_d), T, B, D, _lib.ptr(out))
    return _apply(batch_in, modalities, launch)


def _seeded(batch_in, frac, mode, lengths, modalities, seed, b_offset):
    def launch(rt, idx, x, out):
        T, B, D = _rows_view(x)
        len_d = None if lengths is None else _i32(lengths, x.device)
        flags = torch.empty(T, B, dtype=torch.uint8, device=x.device)
        rt.call('bfvi_draw_deletions', _lib.ptr(len_d), T, B, float(frac), mode, int(seed), idx, int(b_offset),
                _lib.ptr(flags))
        rt.call('bfvi_delete_rows', _lib.ptr(x), _lib.ptr(flags), T, B, D, _lib.ptr(out))
    return _apply(batch_in, modalities, launch)


def _spans(batch_in, span_func, invert, lengths, modalities):
    state = {'lengths': lengths}

    def launch(rt, idx, x, out):
        T, B, D = _rows_view(x)
        if state['lengths'] is None:
            state['lengths'] = [T] * B
        lens = np.asarray(state['lengths'], dtype=np.int64)
        lo, hi = span_func(lens)
        lo_d, hi_d, len_d = _i32(lo, x.device), _i32(hi, x.device), _i32(lens, x.device)   # alive across the call
        rt.call('bfvi_delete_spans', _lib.ptr(x), _lib.ptr(lo_d), _lib.ptr(hi_d), _lib.ptr(len_d), int(invert),
                T, B, D, _lib.ptr(out))
    return _apply(batch_in, modalities, launch)


def rand_delete(batch_in, del_frac, lengths=None, modalities=None, seed=None, b_offset=0):
    """Random memoryless deletions: int(del_frac * length) steps of every sequence
    (datasets/multiseq.py:422-426)."""
    if seed is not None:
        return _seeded(batch_in, del_frac, DELETE_UNIFORM, lengths, modalities, seed, b_offset)
    return func_delete(batch_in, lambda n: np.random.choice(n, int(del_frac * n), False), lengths, modalities)


def burst_delete(batch_in, burst_frac, lengths=None, modalities=None, seed=None, b_offset=0):
    """One random burst of int(burst_frac * length) deleted steps per sequence
    (datasets/multiseq.py:428-434)."""
    if seed is not None:
        return _seeded(batch_in, burst_frac, DELETE_BURST, lengths, modalities, seed, b_offset)

    def spans(lens):
        # np.random.randint(array) consumes the legacy stream exactly like the reference's
        # one-call-per-sequence loop (checked in tests/test_emu_multiseq.py)
        t_start = np.random.randint(lens)
        return t_start, np.minimum(t_start + (burst_frac * lens).astype(np.int64), lens)
    return _spans(batch_in, spans, False, lengths, modalities)


def _frac_span(f_start, f_stop):
    return lambda lens: ((f_start * lens).astype(np.int64), (f_stop * lens).astype(np.int64))


def keep_segment(batch_in, f_start, f_stop, lengths=None, modalities=None):
    """Delete everything outside the time fraction [f_start, f_stop) (datasets/multiseq.py:436-441)."""
    return _spans(batch_in, _frac_span(f_start, f_stop), True, lengths, modalities)


def del_segment(batch_in, f_start, f_stop, lengths=None, modalities=None):
    """Delete the time fraction [f_start, f_stop) (datasets/multiseq.py:443-448)."""
    return _spans(batch_in, _frac_span(f_start, f_stop), False, lengths, modalities)
