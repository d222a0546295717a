"""Pin the oracle against the real reference and write tests/golden/*.pt.

TEST INFRASTRUCTURE ONLY.  Run in the build container (it needs
/root/reference):      python oracle/make_golden.py

For every case below it
  1. builds the reference `models.MultiDMM` (models/dmm.py:29-122) on CPU,
  2. copies its state_dict into the oracle restatement,
  3. runs `step` + backward (and `forward` in all four modes) in BOTH with the
     same injected noise, in fp32 and fp64,
  4. asserts the two agree (fp64: 1e-7 relative; fp32: 2e-5), and
  5. stores inputs, weights, noise and the REFERENCE's outputs as a fixture.
The committed fixtures are what the `-m "not gpu"` oracle tests and the
`-m gpu` parity tests read; /root/reference is never needed at test time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import bfvi_oracle as bo      # noqa: E402
import ref_shim               # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def synth_inputs(mods, dims, t_max, b_dim, lengths, seed, nan_frac=0.0,
                 burst=0, drop_mods=(), corrupt_targets=False):
    """Smooth synthetic multimodal sequences, NaN padded like
    datasets/multiseq.py:340-353 and deleted like :405-434."""
    rng = np.random.RandomState(seed)
    targets, inputs = {}, {}
    tt = np.linspace(0, 3.0, t_max)[:, None, None]
    for m, d in zip(mods, dims):
        phase = rng.uniform(0, 6.28, size=(1, b_dim, d))
        x = np.sin(tt * rng.uniform(0.5, 2.0, size=(1, b_dim, d)) + phase) * (1 + tt / 3)
        x = x + 0.1 * rng.randn(t_max, b_dim, d)
        x = x.astype(np.float32)
        for b, n in enumerate(lengths):
            x[n:, b] = np.nan
        tgt = x.copy()
        inp = x.copy()
        for b, n in enumerate(lengths):
            if burst > 0:
                t0 = rng.randint(n)
                inp[t0:min(t0 + burst, n), b] = np.nan
            if nan_frac > 0:
                idx = rng.choice(n, int(nan_frac * n), False)
                inp[idx, b] = np.nan
                if corrupt_targets:
                    tgt[idx, b] = np.nan
        if m in drop_mods:
            inp[:] = np.nan
        targets[m] = torch.from_numpy(tgt)
        inputs[m] = torch.from_numpy(inp)
    return inputs, targets


def len_to_mask(lengths):
    """datasets/multiseq.py:321-327 (time first, trailing singleton)."""
    t_max = max(lengths)
    ar = torch.arange(t_max).unsqueeze(1)
    return (ar < torch.tensor(lengths).unsqueeze(0)).unsqueeze(-1)


def run_pair(case, dtype):
    """Runs reference and oracle; returns the reference's results."""
    models = ref_shim.import_reference_models()
    mods, dims = case['modalities'], case['dims']
    torch.manual_seed(case['seed'])
    ref = models.MultiDMM(mods, dims, h_dim=case['h_dim'], z_dim=case['z_dim'],
                          min_std=case.get('min_std', 1e-3),
                          device=torch.device('cpu'))
    # perturb so that biases / prior are not at their symmetric initial values
    g = torch.Generator().manual_seed(case['seed'] + 1)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(case.get('perturb', 0.3) * torch.randn(p.shape, generator=g))
    state32 = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dtype)
    params = {k: v.detach().clone().to(dtype).requires_grad_(True)
              for k, v in state32.items()}
    cast = lambda d: {k: v.to(dtype) for k, v in d.items()}
    inputs, targets = cast(case['inputs']), cast(case['targets'])
    mask, lengths = case['mask'], case['lengths']
    kw = dict(case['step_kwargs'])
    n_sets = len(bo.step_sets(len(mods), kw.get('uni_loss', True)))
    noise = case['noise']
    f_mode, s_mode = kw.get('f_mode', 'bfilter'), kw.get('s_mode', 'fsmooth')

    # --- step + backward --------------------------------------------------
    ref_shim.inject_noise(ref, bo.step_noise_tape(noise, f_mode, s_mode, kw.get('match_mult', 0.01) > 0))
    ref.zero_grad()
    loss_ref = ref.step(inputs, mask, case['kld_mult'], case['rec_mults'],
                        targets=targets, lengths=lengths, **kw)
    (loss_ref / sum(lengths)).backward()
    # parameters the loss does not reach (rec_mult == 0) have grad None: store 0
    grads_ref = {k: (p.grad.detach().clone() if p.grad is not None
                     else torch.zeros_like(p)) for k, p in ref.named_parameters()}

    orc = bo.OracleDMM(mods, dims, params, h_dim=case['h_dim'],
                       z_dim=case['z_dim'], min_std=case.get('min_std', 1e-3),
                       draw=bo.step_noise_tape(noise, f_mode, s_mode, kw.get('match_mult', 0.01) > 0))
    loss_orc = orc.step(inputs, mask, case['kld_mult'], case['rec_mults'],
                        targets=targets, lengths=lengths, **kw)
    (loss_orc / sum(lengths)).backward()
    tol = 1e-7 if dtype == torch.float64 else 2e-5
    rel = abs(loss_ref.item() - loss_orc.item()) / abs(loss_ref.item())
    assert rel < tol, ('loss', case['name'], dtype, rel)
    for k, gref in grads_ref.items():
        gorc = params[k].grad if params[k].grad is not None else torch.zeros_like(gref)
        err = (gref - gorc).norm() / (gref.norm() + 1e-30)
        assert err < max(tol, 1e-4 if dtype == torch.float32 else 0), \
            ('grad', case['name'], k, dtype, err.item())

    # --- forward in every mode (no grad) -----------------------------------
    fwd = {}
    with torch.no_grad():
        for mode in ('bfilter', 'ffilter', 'fsmooth', 'bsmooth'):
            for sample, kf in ((False, 1), (True, 1), (False, 6), (True, 6)):
                draws = []
                gen = torch.Generator().manual_seed(1000 + case['seed'])
                t_max, b_dim, z = max(lengths), len(lengths), case['z_dim']
                # enough draws for both passes, in whatever order they are asked
                class Rec(object):
                    def __call__(self, shape):
                        e = torch.randn(*shape, generator=gen)
                        draws.append(e)
                        return e.to(dtype)
                ref._sample_gauss = (lambda mean, std, _r=Rec():
                                     _r(tuple(std.shape)).mul(std).add(mean))
                fkw = dict(lengths=lengths, mode=mode, sample=sample,
                           sample_init=case.get('sample_init', False),
                           flt_particles=kf, smt_particles=1)
                inf_r, pri_r, rec_r = ref(inputs, **fkw)
                orc.draw = bo.NoiseTape([d.clone() for d in draws])
                fkw2 = dict(fkw)
                fkw2.pop('lengths')
                inf_o, pri_o, rec_o = orc.forward(inputs, lengths, **fkw2)
                for a, b in ((inf_r[0], inf_o[0]), (inf_r[1], inf_o[1]),
                             (pri_r[0], pri_o[0]), (pri_r[1], pri_o[1])):
                    assert torch.allclose(a, b, rtol=tol * 10, atol=tol * 10, equal_nan=True), \
                        ('forward', case['name'], mode, sample, kf)
                for m in mods:
                    for a, b in zip(rec_r[m], rec_o[m]):
                        assert torch.allclose(a, b, rtol=tol * 10, atol=tol * 10, equal_nan=True)
                if dtype == torch.float32:
                    fwd['%s/%d/%d' % (mode, int(sample), kf)] = {
                        'draws': [d.clone() for d in draws],
                        'infer': [t.clone() for t in inf_r],
                        'prior': [t.clone() for t in pri_r],
                        'recon': {m: [t.clone() for t in rec_r[m]] for m in mods}}
    return state32, loss_ref.item(), grads_ref, fwd


def build_cases():
    cases = []

    def add(name, mods, dims, z, h, t_max, lengths, seed, kld_mult, rec_mults,
            step_kwargs=None, **synth):
        b_dim = len(lengths)
        perturb = synth.pop('perturb', 0.3)
        inputs, targets = synth_inputs(mods, dims, t_max, b_dim, lengths, seed, **synth)
        kw = dict(step_kwargs or {})
        n_sets = len(bo.step_sets(len(mods), kw.get('uni_loss', True)))
        noise = bo.make_step_noise(n_sets, t_max, b_dim, z,
                                   kw.get('train_particles', 25),
                                   kw.get('match_particles', 50), seed=seed + 7)
        cases.append(dict(name=name, modalities=mods, dims=dims, z_dim=z, h_dim=h,
                          lengths=lengths, mask=len_to_mask(lengths), seed=seed,
                          inputs=inputs, targets=targets, kld_mult=kld_mult,
                          rec_mults=rec_mults, step_kwargs=kw, noise=noise,
                          perturb=perturb))

    # C1-shaped (spirals dims, defaults of spirals.py:44-51,64-73), ragged + burst
    add('spirals_ragged', ['spiral-x', 'spiral-y'], [1, 1], 5, 20, 12,
        [12, 12, 10, 9, 7, 5], seed=11, kld_mult=0.7,
        rec_mults={'spiral-x': 0.5, 'spiral-y': 0.5}, burst=2)
    # C2-shaped: 50 % uniform missing in inputs AND targets, rec_mults 1.0
    add('spirals_half_missing', ['spiral-x', 'spiral-y'], [1, 1], 5, 20, 10,
        [10] * 8, seed=12, kld_mult=1.0,
        rec_mults={'spiral-x': 1.0, 'spiral-y': 1.0}, nan_frac=0.5, burst=1,
        corrupt_targets=True)
    # three wider modalities, forward-filter term, odd particle counts, one
    # modality dropped entirely, one rec_mult == 0
    add('gauss3_ffilter', ['a', 'b', 'c'], [3, 2, 4], 4, 8, 7,
        [7, 6, 6, 3], seed=13, kld_mult=0.5,
        rec_mults={'a': 0.3, 'b': 0.0},
        step_kwargs=dict(f_mode='ffilter', s_mode='fsmooth', train_particles=7,
                         match_particles=9, f_mult=0.3, s_mult=0.7, match_mult=0.05),
        nan_frac=0.3, drop_mods=('b',))
    # backward smoothing (forward filter first).  NOTE the reference zeroes the
    # filter-prior mask at index T-1 in BOTH smoothing directions
    # (models/dmm.py:482), so bsmooth is only finite when something is observed
    # at T-1: complete data here.
    add('bsmooth_full', ['a', 'b'], [2, 3], 4, 8, 6, [6, 6, 6], seed=17,
        kld_mult=0.9, rec_mults={'a': 0.5, 'b': 0.25},
        step_kwargs=dict(s_mode='bsmooth', train_particles=5, match_particles=6))
    # single modality (no multimodal term, models/dgts.py:119), no unimodal skip
    add('single_mod', ['only'], [2], 3, 6, 5, [5, 4, 2], seed=14, kld_mult=1.0,
        rec_mults={}, step_kwargs=dict(train_particles=3, match_particles=4),
        burst=1)
    # uni_loss=False, match_mult=0
    add('no_uni_no_match', ['u', 'v'], [2, 2], 6, 12, 6, [6, 6, 5], seed=15,
        kld_mult=1.3, rec_mults={'u': 1.0, 'v': 2.0},
        step_kwargs=dict(uni_loss=False, match_mult=0.0, train_particles=4),
        nan_frac=0.2)
    # medium dims for the generic-dimension kernel path
    add('medium_dims', ['p', 'q'], [16, 8], 16, 48, 5, [5, 5, 4, 2], seed=16,
        kld_mult=1.0, rec_mults={'p': 1.0 / 32, 'q': 1.0 / 16},
        step_kwargs=dict(train_particles=5, match_particles=6), nan_frac=0.25,
        perturb=0.1)
    return cases


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for case in build_cases():
        state32, loss64, grads64, _ = run_pair(case, torch.float64)
        state32b, loss32, grads32, fwd = run_pair(case, torch.float32)
        for k in state32:
            assert torch.equal(state32[k], state32b[k])
        fixture = {k: case[k] for k in
                   ('name', 'modalities', 'dims', 'z_dim', 'h_dim', 'lengths', 'mask',
                    'inputs', 'targets', 'kld_mult', 'rec_mults', 'step_kwargs',
                    'noise')}
        fixture['min_std'] = case.get('min_std', 1e-3)
        fixture['state_dict'] = state32
        fixture['ref_loss_fp32'] = loss32
        fixture['ref_loss_fp64'] = loss64
        fixture['ref_grads_fp32'] = grads32
        fixture['ref_grads_fp64'] = {k: v.to(torch.float64) for k, v in grads64.items()}
        fixture['ref_forward_fp32'] = fwd
        if case['name'] == 'spirals_ragged':
            # seeded construction of the reference (spirals.py:44-51 shapes): lets a
            # CPU test check that our module tree initialises bit-identically
            models = ref_shim.import_reference_models()
            torch.manual_seed(1)
            ref0 = models.MultiDMM(case['modalities'], case['dims'], h_dim=case['h_dim'],
                                   z_dim=case['z_dim'], device=torch.device('cpu'))
            fixture['seeded_init_seed1'] = {k: v.clone() for k, v in ref0.state_dict().items()}
        fixture['provenance'] = (
            'generated by oracle/make_golden.py from the unmodified reference '
            '/root/reference/models/dmm.py (torch %s, CPU); grads are of '
            'loss/sum(lengths) as in trainer.py:242-243' % torch.__version__)
        path = os.path.join(OUT_DIR, case['name'] + '.pt')
        torch.save(fixture, path)
        print('%-28s loss32=%.6f loss64=%.6f  -> %s (%.1f kB)' % (
            case['name'], loss32, loss64, os.path.relpath(path),
            os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
