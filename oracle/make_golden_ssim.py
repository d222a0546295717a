"""Pins oracle/metrics_oracle.py against the UNMODIFIED reference `eval_ssim` (utils.py:165-212, imported from
/root/reference with empty matplotlib stubs — utils.py:8-9 import it for plotting only) and stores
the reference's outputs: tests/golden/metrics/ssim.pt.   python oracle/make_golden_ssim.py"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')
for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.lines', 'matplotlib.collections'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['matplotlib.lines'].Line2D = object
sys.modules['matplotlib.collections'].EllipseCollection = object
warnings.filterwarnings('ignore', category=SyntaxWarning)
import utils as ref          # noqa: E402  (the reference's utils.py)
import metrics_oracle as orc  # noqa: E402


def main():
    rng = np.random.RandomState(5)
    cases = []
    for name, shape, win_size, sigma, dr in (('weizmann_video', (3, 3, 64, 64), 11, 1.5, 1.0),
                                             ('weizmann_mask', (3, 1, 64, 64), 11, 1.5, 1.0),
                                             ('odd_size', (4, 2, 23, 37), 7, 1.0, 1.0),
                                             ('wide', (2, 1, 40, 90), 11, 1.5, 255.0),
                                             ('exact_window', (3, 1, 11, 11), 11, 1.5, 1.0)):
        x = rng.rand(*shape).astype(np.float32) * dr
        # reconstruction-like pair: smooth target, noisy / blurred estimate
        y = np.clip(x + 0.1 * dr * rng.standard_normal(shape), 0, dr).astype(np.float32)
        if name == 'weizmann_mask':
            x = (x > 0.5 * dr).astype(np.float32)
        ssim, cs = ref.eval_ssim(torch.from_numpy(x), torch.from_numpy(y), win_size=win_size, win_sigma=sigma,
                                 data_range=dr, full=True)
        o_ssim, o_cs = orc.eval_ssim(x, y, win_size, sigma, data_range=dr)
        assert np.allclose(ssim.numpy(), o_ssim, rtol=0, atol=1e-5), (name, ssim, o_ssim)
        assert np.allclose(cs.numpy(), o_cs, rtol=0, atol=1e-5)
        cases.append({'name': name, 'x': torch.from_numpy(x), 'y': torch.from_numpy(y), 'win_size': win_size,
                      'win_sigma': sigma, 'data_range': dr, 'ssim': ssim.clone(), 'cs': cs.clone()})
    out_dir = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'metrics')
    os.makedirs(out_dir, exist_ok=True)
    torch.save(cases, os.path.join(out_dir, 'ssim.pt'))
    print('wrote', len(cases), 'cases')


if __name__ == '__main__':
    main()
