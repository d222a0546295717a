"""Batch-sharded data parallelism for the BFVI step (one process per GPU).

The reference has no distributed code; sequences are independent
units of the step, so the batch axis shards with no exchange inside the step and
exactly ONE collective per step: a SUM all-reduce of the flat fp32 gradient buffer
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  What has to be global for
the N-rank result to equal the single-process reference:

  * `t_max`: padded steps still advance the chain (models/dmm.py:373-378), so every
    rank keeps the GLOBAL maximum length — `shard_batch` never trims the time axis;
  * the normaliser `sum(lengths)` of trainer.py:242 — returned as `n_global`;
  * reparameterisation noise is indexed by the GLOBAL sequence index: shards are
    contiguous slices and `b_offset` is the global index of the shard's first
    sequence (bfvi_noise.b_offset in include/bfvi.h);
  * the prior-matching term (models/dmm.py:541-545) is linear in `mask.sum()`: each
    rank uses its local count and the same match noise (same seed), the sum over
    ranks is the global term;
  * all ranks draw the same Philox seed per step (`sync_seed`).

Every chain runs all `t_max` steps whatever its length, so contiguous shards are
load-balanced even though the collated batch is sorted by length.
"""
import torch


def shard_bounds(n_seq, rank, world):
    """[b0, b1) of the contiguous shard of `rank`; sizes differ by at most one."""
    base, rem = divmod(n_seq, world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def shard_batch(inputs, targets, mask, lengths, rank, world):
    """Slices a collated batch (time-first tensors, datasets/multiseq.py:372-386) for one
    rank.  Returns a dict with the local `inputs`, `targets`, `mask`, `lengths`, the noise
    offset `b_offset` and the global normaliser `n_global` = sum of ALL lengths."""
    n_seq = len(lengths)
    b0, b1 = shard_bounds(n_seq, rank, world)
    cut = lambda d: None if d is None else {k: v[:, b0:b1].contiguous() for k, v in d.items()}
    return {'inputs': cut(inputs), 'targets': cut(targets), 'mask': mask[:, b0:b1].contiguous(),
            'lengths': list(lengths[b0:b1]), 'b_offset': b0, 'n_global': float(sum(lengths)),
            't_max': int(mask.shape[0])}


def all_reduce_flat(flat_grad, group=None):
    """The step's only collective: in-place SUM of the flat gradient buffer."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def sync_seed(seed, group=None, device=None):
    """Broadcast rank 0's Philox seed so every rank regenerates the same match noise."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(seed)
    t = torch.tensor([int(seed)], dtype=torch.int64, device=device)
    dist.broadcast(t, src=0, group=group)
    return int(t.item())


class _SharedSeeds(object):
    """Per-step Philox seeds that are the same on every rank without a collective per step: rank 0
    draws ONE base seed, broadcasts it once, and every rank steps the same generator."""

    def __init__(self, base):
        self.gen = torch.Generator().manual_seed(int(base))

    def __call__(self):
        return int(torch.randint(0, 2 ** 62, (1,), generator=self.gen).item())


def attach(model, shard, group=None, base_seed=None):
    """Wires a MultiDMM replica for data-parallel training on `shard` (from shard_batch):
      * rank 0's parameters are broadcast (replicas must start identical);
      * every rank draws the same Philox seed per step (`model.seed_source`, one broadcast here);
      * noise is indexed by the global sequence index (`b_offset`);
      * `loss.backward()` all-reduces the flat gradient (the step's only collective).
    The caller divides the loss by shard['n_global'] (trainer.py:242 with the GLOBAL sum of lengths).
    A rank whose shard is EMPTY (world > sequences) must not call model.step — use `step_or_zero`,
    which still joins the all-reduce with a zero gradient so the other ranks do not hang."""
    import torch.distributed as dist
    model._ensure_flat()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(model._flat, src=0, group=group)
    if base_seed is None:
        base_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    model.seed_source = _SharedSeeds(sync_seed(base_seed, group, device=model._flat.device
                                               if dist.is_initialized() and dist.get_backend(group) == 'nccl' else None))
    model.b_offset = int(shard['b_offset'])
    model.grad_sync = lambda flat_grad: all_reduce_flat(flat_grad, group)
    return model


def step_or_zero(model, shard, kld_mult, rec_mults, group=None, **kwargs):
    """model.step on the shard + backward of loss / n_global; an EMPTY shard contributes a zero gradient to
    the same all-reduce (and consumes the step's shared seed) instead of failing argument checks while the
    other ranks wait in the collective.  Returns the local un-normalised loss (0-dim tensor)."""
    if len(shard['lengths']) == 0:
        model._ensure_flat()
        if getattr(model, 'seed_source', None) is not None:
            model.seed_source()
        zero = torch.zeros_like(model._flat)
        all_reduce_flat(zero, group)
        model.last_flat_grad = zero
        for _, off, p in model._slots:
            p.grad = zero[off:off + p.numel()].view(p.shape)
        return torch.zeros((), device=model._flat.device)
    loss = model.step(shard['inputs'], shard['mask'], kld_mult, rec_mults, targets=shard['targets'],
                      lengths=shard['lengths'], **kwargs)
    (loss / shard['n_global']).backward()
    return loss.detach()
