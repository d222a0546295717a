"""Where do the fused GTF kernels spend their time?  bfvi_gtf_probe (each kernel alone on ROWS latent rows, the
particle-pass launch of a C3 batch tile) under the development ablation masks of csrc/bfvi_fused.cuh (BFVI_FUSED_ABL:
1 no FP16 tile stores, 2 no fp32 row stores, 4 no ReLU bits, 8 no second-level MMAs (heads / dz), 16 no first-level
MMAs (hidden), 32 no row-warp arithmetic, 64 row results through st.global, 128 no ring copies).  Ablated results are
garbage: timing only.  The masks exist only in the ablation build:
    python tools/variants.py ablate:BFVI_FUSED_ABLATE
    BFVI_LIB_PATH=$PWD/tools/_variants/libbfvi_ablate.so python tools/probe_fused_ablate.py 0 1 2 4 8 16 32 63 191
(the product library ignores BFVI_FUSED_ABL: mask 0 there times the shipped kernels)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import bfvi_oracle as bo      # noqa: E402
import helpers                # noqa: E402
from multimodal_dmm_b200 import _lib  # noqa: E402

ROWS = int(os.environ.get('ROWS', 158400))
H = 512


def main():
    lib = _lib.load()
    mods, dims = ['m0'], [16]
    sd = bo.init_params(mods, dims, h_dim=H, z_dim=64, seed=1)
    model = _lib.make_model(dims, ['Normal'], 64, H, 1e-3)
    flat, lay = helpers.pack_params(lib, model, mods, ['Normal'], sd, 'cuda')
    nbytes = int(lib.dll.bfvi_gtf_workspace(C.byref(model), ROWS))
    ws = helpers.aligned_empty(nbytes, 'cuda')
    z = torch.randn(ROWS, 64, device='cuda')
    scratch = torch.empty(5 * ROWS * 64, device='cuda')
    ms = C.c_float(0.0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    names = ['fwd', 'fwd<keep>', 'bwd', 'wgrad16']
    masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 32, 3, 7, 24, 63]
    print('rows %d, H %d; ms per launch' % (ROWS, H))
    print('%-10s' % 'abl' + ''.join('%12s' % n for n in names))
    for m in masks:
        os.environ['BFVI_FUSED_ABL'] = str(m)
        line = '%-10d' % m
        for which in range(4):
            for iters in (2, 10):
                lib.call('bfvi_gtf_probe', C.byref(model), _lib.ptr(flat), 1, which, _lib.ptr(z), ROWS, iters, _lib.ptr(scratch),
                         _lib.ptr(ws), C.c_size_t(nbytes), C.byref(ms), st)
            line += '%12.4f' % ms.value
        print(line, flush=True)
    os.environ['BFVI_FUSED_ABL'] = '0'


if __name__ == '__main__':
    main()
