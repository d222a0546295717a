"""Golden outputs of the UNMODIFIED reference's image modules (models/common.py:70-175: Conv, Deconv, ImageEncoder,
ImageDecoder) run in float64: two training passes (outputs, every parameter gradient, BatchNorm running statistics)
and an evaluation-mode pass, at a small size (16 x 16 images, n_kernels 8) and with the reference's own seeded
initial weights stored in the fixture.

TEST INFRASTRUCTURE ONLY; run in the build container:   python oracle/make_golden_image.py
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim               # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'image', 'modules.pt')
CFG = dict(seed=21, z_dim=12, img_size=16, n_channels=3, n_kernels=8, frames=6, steps=2)


def make_batch(cfg, step):
    g = torch.Generator().manual_seed(1000 * cfg['seed'] + step)
    shape = (cfg['frames'], cfg['n_channels'], cfg['img_size'], cfg['img_size'])
    return (torch.rand(*shape, generator=g), (torch.rand(*shape, generator=g) > 0.5).float(),
            torch.randn(cfg['frames'], cfg['z_dim'], generator=g))


def build(common, cfg):
    torch.manual_seed(cfg['seed'])
    kw = dict(img_size=cfg['img_size'], n_channels=cfg['n_channels'], n_kernels=cfg['n_kernels'])
    return common.ImageEncoder(cfg['z_dim'], True, **kw), common.ImageDecoder(cfg['z_dim'], **kw)


def run(enc, dec, cfg, cast, on_step=None):
    """The passes both sides run: loss = BCE(decoder(mean + 0.1 std)) + <mean, w> + |std|^2 per training step."""
    record = {'steps': []}
    for step in range(cfg['steps']):
        x, tgt, wz = (cast(t) for t in make_batch(cfg, step))
        for m in (enc, dec):
            m.train()
            m.zero_grad()
        mean, std = enc(x)
        (probs,) = dec(mean + 0.1 * std)
        loss = F.binary_cross_entropy(probs, tgt, reduction='sum') + (mean * wz).sum() + (std ** 2).sum()
        loss.backward()
        rec = {'mean': mean, 'std': std, 'probs': probs, 'loss': loss}
        rec.update({'grad enc.' + k: p.grad for k, p in enc.named_parameters()})
        rec.update({'grad dec.' + k: p.grad for k, p in dec.named_parameters()})
        rec.update({'buffer enc.' + k: b for k, b in enc.named_buffers()})
        rec.update({'buffer dec.' + k: b for k, b in dec.named_buffers()})
        record['steps'].append({k: v.detach().double().cpu().clone() for k, v in rec.items()})
    for m in (enc, dec):
        m.eval()
    with torch.no_grad():
        x = cast(make_batch(cfg, 99)[0])
        mean, std = enc(x)
        record['eval'] = {'mean': mean.double().cpu(), 'std': std.double().cpu(), 'probs': dec(mean)[0].double().cpu()}
    return record


def main():
    ref_models = ref_shim.import_reference_models()
    enc, dec = build(ref_models.common, CFG)
    init = {'enc': {k: v.clone() for k, v in enc.state_dict().items()},
            'dec': {k: v.clone() for k, v in dec.state_dict().items()}}
    record = run(enc.double(), dec.double(), CFG, lambda t: t.double())
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save({'cfg': CFG, 'init': init, 'ref': record}, OUT)
    print(OUT, os.path.getsize(OUT), 'bytes; loss', [s['loss'].item() for s in record['steps']])


if __name__ == '__main__':
    main()
