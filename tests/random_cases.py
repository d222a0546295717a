"""Seeded random BFVI step problems for differential testing against the oracle: random
modality counts / feature dims / lengths / NaN patterns / particle counts / modes, including the
edge cases the reference's collation produces (length-1 sequences, fully missing modalities at a
step, a batch of one, T = 1) and particle counts that exercise every lane-group geometry."""
import numpy as np
import torch

import bfvi_oracle as bo

SMALL_DIMS = [(5, 20), (4, 8), (3, 6), (6, 12)]


def make_case(seed, dims_zh=None):
    rng = np.random.RandomState(seed)
    z, h = dims_zh or SMALL_DIMS[rng.randint(len(SMALL_DIMS))]
    n_mods = int(rng.choice([1, 2, 2, 3]))
    dims = [int(rng.randint(1, 5)) for _ in range(n_mods)]
    mods = ['m%d' % i for i in range(n_mods)]
    t_max = int(rng.choice([1, 2, 3, 5, 8, 12]))
    b_dim = int(rng.choice([1, 2, 3, 5, 9]))
    lengths = sorted([int(rng.randint(1, t_max + 1)) for _ in range(b_dim)], reverse=True)
    lengths[0] = t_max
    s_mode = str(rng.choice(['fsmooth', 'fsmooth', 'fsmooth', 'bsmooth']))
    if s_mode == 'bsmooth':
        # the reference's backward smoothing starts at t = T-1, where the filtering-prior expert is
        # masked (models/dmm.py:482): a sequence that is padded or unobserved there has NO expert
        # left and the reference itself returns NaN.  Keep bsmooth cases well-posed.
        lengths = [t_max] * b_dim
    k_train = int(rng.choice([1, 2, 5, 7, 25, 32, 33, 40]))
    k_match = int(rng.choice([1, 3, 50, 70]))
    g = torch.Generator().manual_seed(seed)
    targets, inputs = {}, {}
    for m, d in zip(mods, dims):
        x = torch.randn(t_max, b_dim, d, generator=g)
        for b, n in enumerate(lengths):
            x[n:, b] = float('nan')
        tgt = x.clone()
        tgt[torch.rand(t_max, b_dim, generator=g) < 0.1] = float('nan')         # corrupted dataset
        inp = tgt.clone()
        inp[torch.rand(t_max, b_dim, generator=g) < rng.choice([0.0, 0.3, 0.6])] = float('nan')
        if rng.rand() < 0.15 and m != mods[0]:
            inp[:] = float('nan')                                                # dropped modality
        if m == mods[0]:
            inp[t_max - 1] = x[t_max - 1]          # the smoothing passes need an observation at the ends
            inp[0] = x[0]
        targets[m], inputs[m] = tgt, inp
    mask = torch.zeros(t_max, b_dim, 1, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[:n, b] = True
    uni_loss = bool(rng.rand() < 0.8) or n_mods == 1
    n_sets = len(bo.step_sets(n_mods, uni_loss))
    kw = {'train_particles': k_train, 'match_particles': k_match, 'uni_loss': uni_loss,
          'f_mode': str(rng.choice(['bfilter', 'ffilter'])), 's_mode': s_mode,
          'f_mult': float(rng.choice([0.5, 0.3])), 's_mult': float(rng.choice([0.5, 0.7])),
          'match_mult': float(rng.choice([0.01, 0.0, 0.1]))}
    noise = {'match': torch.randn(2, k_match, z, generator=g),
             'filt': torch.randn(n_sets, t_max, b_dim, 1, z, generator=g),
             'sflt': torch.randn(n_sets, t_max, b_dim, k_train, z, generator=g),
             'ssmt': torch.randn(n_sets, t_max, b_dim, 1, z, generator=g)}
    return dict(modalities=mods, dims=dims, z_dim=z, h_dim=h, min_std=1e-3, inputs=inputs, targets=targets,
                mask=mask, lengths=lengths, kld_mult=float(rng.choice([1.0, 0.4])),
                rec_mults={m: float(rng.choice([1.0, 0.5])) for m in mods}, step_kwargs=kw, noise=noise,
                state_dict=bo.init_params(mods, dims, h_dim=h, z_dim=z, seed=seed, scale=1.5))


def oracle_step(fx, dtype=torch.float64):
    kw = dict(fx['step_kwargs'])
    uni = kw.pop('uni_loss')
    params = {k: v.clone().to(dtype).requires_grad_(True) for k, v in fx['state_dict'].items()}
    tape = bo.step_noise_tape({k: v.to(dtype) for k, v in fx['noise'].items()}, f_mode=kw['f_mode'],
                              s_mode=kw['s_mode'], with_match=kw['match_mult'] > 0)
    orc = bo.OracleDMM(fx['modalities'], fx['dims'], params, h_dim=fx['h_dim'], z_dim=fx['z_dim'],
                       min_std=fx['min_std'], draw=tape)
    cast = lambda d: {k: v.to(dtype) for k, v in d.items()}
    loss = orc.step(cast(fx['inputs']), fx['mask'], fx['kld_mult'], fx['rec_mults'], targets=cast(fx['targets']),
                    uni_loss=uni, lengths=fx['lengths'], **kw)
    loss.backward()
    return loss.item(), {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}


def check(loss, grads, ref_loss, ref_grads, elbo_tol=1e-4, grad_tol=1e-3):
    """Returns a list of violations (empty = parity)."""
    bad = []
    if not (abs(loss - ref_loss) <= elbo_tol * max(abs(ref_loss), 1e-6)):
        bad.append(('loss', loss, ref_loss))
    scale = max(g.norm().item() for g in ref_grads.values())
    for k, g_ref in ref_grads.items():
        err = (grads[k].double() - g_ref.double()).norm().item()
        if not (err <= grad_tol * g_ref.norm().item() or err <= 1e-6 * scale):
            bad.append((k, err, g_ref.norm().item()))
    return bad
