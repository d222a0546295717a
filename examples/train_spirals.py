"""Train the spirals MDMM with BFVI on synthetic noisy spirals — the loop of the reference's
trainer.py:218-262 (KL anneal, burst deletion, model.step, /sum(lengths), backward, optimiser
step) on the B200-native step, without needing the reference repository.

    python examples/train_spirals.py [--epochs 20] [--batch 100] [--fused-adam]

With the reference checked out, the same model drops in under its own trainer instead: see
INTEGRATION.md (swap `import models` for `import multimodal_dmm_b200.models as models`).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_dmm_b200.models as models   # noqa: E402
from multimodal_dmm_b200 import optim          # noqa: E402
from bench import spirals_batch                # noqa: E402  (datasets/spirals.py:47-84 restated)


def burst_delete(x, frac, rng):
    """One NaN burst per sequence (datasets/multiseq.py:428-434)."""
    t_max, b_dim = x.shape[:2]
    n = int(frac * t_max)
    out = x.clone()
    for b in range(b_dim):
        s = rng.randint(t_max)
        out[s:s + n, b] = float('nan')
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=20)
    ap.add_argument('--batch', type=int, default=100)
    ap.add_argument('--n-train', type=int, default=600)
    ap.add_argument('--lr', type=float, default=5e-3)
    ap.add_argument('--fused-adam', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(1)
    rng = np.random.RandomState(1)
    data, _ = spirals_batch(a.n_train, 100, seed=1)
    data = {k: torch.from_numpy(v) for k, v in data.items()}
    mods = ['spiral-x', 'spiral-y']
    model = models.MultiDMM(mods, [1, 1], h_dim=20, z_dim=5, device=dev).train()
    opt = optim.FlatAdam(model, lr=a.lr, weight_decay=1e-4) if a.fused_adam else \
        torch.optim.Adam(model.parameters(), lr=a.lr, weight_decay=1e-4)
    rec_mults = {m: 0.5 for m in mods}
    for epoch in range(1, a.epochs + 1):
        kld_mult = min(1.0, epoch / 10.0)                               # utils.anneal
        perm = rng.permutation(a.n_train)
        t0, total, seen = time.perf_counter(), 0.0, 0
        for i in range(0, a.n_train, a.batch):
            idx = torch.from_numpy(perm[i:i + a.batch])
            targets = {m: data[m][:, idx] for m in mods}
            inputs = {m: burst_delete(targets[m], 0.1, rng) for m in mods}
            lengths = [100] * len(idx)
            mask = torch.ones(100, len(idx), 1, dtype=torch.bool, device=dev)
            loss = model.step({m: v.to(dev) for m, v in inputs.items()}, mask, kld_mult, rec_mults,
                              targets={m: v.to(dev) for m, v in targets.items()}, lengths=lengths)
            (loss / sum(lengths)).backward()
            opt.step()
            opt.zero_grad()
            total += loss.item()
            seen += sum(lengths)
        dt = time.perf_counter() - t0
        print('epoch %3d  loss/timestep %9.4f  kld_mult %.2f  %.0f seq-timesteps/s (host loop included)'
              % (epoch, total / seen, kld_mult, seen / dt))


if __name__ == '__main__':
    main()
