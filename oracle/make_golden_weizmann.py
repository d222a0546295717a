"""Golden outputs of the UNMODIFIED reference for a Weizmann-shaped model (BASELINE config 4):
conv image encoders / decoders injected as custom modules, Bernoulli + Categorical modalities,
a dropped modality, NaN bursts (weizmann.py:53-77, trainer.py:289-296).

TEST INFRASTRUCTURE ONLY; run in the build container:   python oracle/make_golden_weizmann.py

The fixture stores no weights: our modules reproduce the reference's seeded initialisation
bit-for-bit (tests/test_host_api.py), so `seed` regenerates the same model on the GPU box.
Large gradient tensors are stored as (norm, projection on a seeded random vector).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim               # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'weizmann', 'forward_fsmooth.pt')

CFG = dict(seed=7, z_dim=32, h_dim=32, t_max=4, lengths=[4, 4, 3], k_flt=3, kld_mult=0.7,
           rec_mults={'video': 1.0, 'mask': 1.0, 'action': 10.0})
MODS = ['video', 'mask', 'action']
DIMS = {'video': (3, 64, 64), 'mask': (1, 64, 64), 'action': 10}
DISTS = {'video': 'Bernoulli', 'mask': 'Bernoulli', 'action': 'Categorical'}


def build(models_pkg, cfg, device='cpu'):
    """Same construction order as weizmann.py:64-76 (it fixes the RNG consumption order)."""
    torch.manual_seed(cfg['seed'])
    c = models_pkg.common
    enc = {'video': c.ImageEncoder(cfg['z_dim'], True), 'mask': c.ImageEncoder(cfg['z_dim'], True, n_channels=1)}
    dec = {'video': c.ImageDecoder(cfg['z_dim']), 'mask': c.ImageDecoder(cfg['z_dim'], n_channels=1)}
    return models_pkg.MultiDMM(MODS, dims=[DIMS[m] for m in MODS], dists=[DISTS[m] for m in MODS],
                               encoders=enc, decoders=dec, z_dim=cfg['z_dim'], h_dim=cfg['h_dim'],
                               device=torch.device(device))


def make_data(cfg):
    g = torch.Generator().manual_seed(cfg['seed'] + 1)
    t_max, b_dim = cfg['t_max'], len(cfg['lengths'])
    targets = {'video': torch.rand(t_max, b_dim, 3, 64, 64, generator=g),
               'mask': (torch.rand(t_max, b_dim, 1, 64, 64, generator=g) > 0.5).float(),
               'action': torch.randint(0, 10, (t_max, b_dim, 1), generator=g).float()}
    for b, n in enumerate(cfg['lengths']):
        for v in targets.values():
            v[n:, b] = float('nan')                         # padding
    inputs = {'video': targets['video'].clone(), 'action': targets['action'].clone()}   # `mask` dropped
    inputs['video'][1, 0] = float('nan')                    # deleted frames
    inputs['action'][2, 1] = float('nan')
    mask = torch.zeros(t_max, b_dim, 1, dtype=torch.bool)
    for b, n in enumerate(cfg['lengths']):
        mask[:n, b] = True
    eps_flt = torch.randn(t_max, b_dim, cfg['k_flt'], cfg['z_dim'], generator=g)
    eps_smt = torch.randn(t_max, b_dim, 1, cfg['z_dim'], generator=g)
    return inputs, targets, mask, eps_flt, eps_smt


def projector(shape, key):
    g = torch.Generator().manual_seed(abs(hash(key)) % (2 ** 31))
    return torch.randn(shape, generator=g)


def summarise(grads):
    out = {}
    for k, g in grads.items():
        g = g.detach().float().cpu()
        out[k] = {'full': g.clone()} if g.numel() <= 4096 else {'norm': g.norm().item(), 'numel': g.numel()}
        if g.numel() > 4096:
            # deterministic projection: cosine-weighted sum with a fixed pattern (no hash seeds)
            w = torch.cos(torch.arange(g.numel(), dtype=torch.float32) * 0.37).reshape(g.shape)
            out[k]['proj'] = (g * w).sum().item()
    return out


def main():
    ref_shim.install()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    import models as ref_models
    cfg = CFG
    model = build(ref_models, cfg)
    model.train()
    inputs, targets, mask, eps_flt, eps_smt = make_data(cfg)
    t_max = cfg['t_max']
    tape = [eps_flt[t].permute(1, 0, 2).contiguous() for t in range(t_max - 1, -1, -1)] + \
           [eps_smt[t].permute(1, 0, 2).contiguous() for t in range(t_max)]
    it = iter(tape)
    model._sample_gauss = lambda mean, std: next(it).to(mean.dtype) * std + mean
    infer, prior, recon = model(inputs, lengths=cfg['lengths'], mode='fsmooth', flt_particles=cfg['k_flt'])
    loss = model.loss(targets, infer, prior, recon, mask, cfg['kld_mult'], cfg['rec_mults'])
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    torch.save({'cfg': cfg, 'ref_loss': loss.item(), 'ref_infer_mean': infer[0].detach(), 'ref_infer_std': infer[1].detach(),
                'ref_grads': summarise(grads), 'provenance': 'reference models/dmm.py forward(fsmooth)+loss+backward, torch %s, fp32 CPU'
                % torch.__version__}, OUT)
    print('wrote', OUT, 'loss', loss.item(), 'n grads', len(grads))


if __name__ == '__main__':
    main()
