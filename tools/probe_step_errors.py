"""Per-tensor gradient errors of the C3-dims step case (tests/test_gpu_large.py) against the fp64 oracle, on B200.
    python tools/probe_step_errors.py [k_train ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bfvi_oracle as bo  # noqa: E402
import helpers  # noqa: E402
import test_gpu_large as tl  # noqa: E402
from multimodal_dmm_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    for k_train in [int(a) for a in sys.argv[1:]] or [5, 1]:
        fx = tl.step_case('c3_dims', k_train=k_train, k_match=7, seed=21)
        params = {k: v.clone().double().requires_grad_(True) for k, v in fx['state_dict'].items()}
        orc = bo.OracleDMM(fx['modalities'], fx['dims'], params, h_dim=fx['h_dim'], z_dim=fx['z_dim'],
                           min_std=fx['min_std'], draw=bo.step_noise_tape(fx['noise']))
        cast = lambda d: {k: v.double() for k, v in d.items()}
        ref = orc.step(cast(fx['inputs']), fx['mask'], fx['kld_mult'], fx['rec_mults'], targets=cast(fx['targets']),
                       lengths=fx['lengths'], **fx['step_kwargs'])
        ref.backward()
        for prec in (2, 0):
            loss, grads, _ = helpers.run_step(lib, fx, 'cuda', kwargs={'precision': prec})
            errs = sorted(((tl.rel_err(grads[k], p.grad), k) for k, p in params.items() if p.grad.norm() > 0), reverse=True)
            print('K=%d precision=%d loss rel %.2e | %s' % (k_train, prec, abs(loss - ref.item()) / abs(ref.item()),
                  '  '.join('%s %.2e' % (k.replace('trans.', ''), e) for e, k in errs if 'trans' in k)))
            print('    non-transition worst: %s' % '  '.join('%s %.2e' % (k, e) for e, k in [x for x in errs if 'trans' not in x[1]][:4]))


if __name__ == '__main__':
    main()
