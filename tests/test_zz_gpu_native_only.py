"""Runs LAST in the GPU suite (file name): it attaches the profiler (CUPTI) to the process, which no other test needs."""
import pytest

import multimodal_dmm_b200.models.common as common

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(180)
def test_no_library_convolution_or_gemm_on_the_image_path():
    """A forward + backward of ImageEncoder -> ImageDecoder launches this library's kernels (bfvi::conv::*) for every
    convolution, BatchNorm and dense layer: no cuDNN / cuBLAS / CUTLASS kernel name shows up in the profile.
    (Skipped when the profiler cannot collect kernel names on this box.)"""
    import re
    import torch
    import torch.nn.functional as F
    enc = common.ImageEncoder(32, img_size=32, n_kernels=16).cuda().train()
    dec = common.ImageDecoder(32, img_size=32, n_kernels=16).cuda().train()
    x = torch.rand(4, 3, 32, 32, device='cuda')

    def step():
        mean, std = enc(x)
        F.binary_cross_entropy(dec(mean + 0.1 * std)[0], x, reduction='sum').backward()
    step()
    torch.cuda.synchronize()
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if getattr(e, 'device_type', None) is not None
                 and 'cuda' in str(e.device_type).lower()]
    except Exception as exc:                                    # no CUPTI on the box: nothing to assert on
        pytest.skip('profiler unavailable: %r' % (exc,))
    ours = [n for n in names if 'bfvi::' in n]
    if not names or not ours:
        pytest.skip('the profiler collected no kernel names')
    for kind in ('conv_gather_kernel', 'conv_scatter_kernel', 'conv_wgrad_kernel', 'chan_reduce_kernel', 'bn_apply_kernel',
                 'bn_bwd_apply_kernel', 'dense_gemm_kernel'):
        assert any(kind in n for n in ours), kind
    library = re.compile(r'cudnn|cublas|cutlass|xmma|sgemm|gemm_|gemv|convolve|wgrad|dgrad|implicit', re.I)
    foreign = [n for n in names if 'bfvi::' not in n and library.search(n)]
    assert not foreign, foreign
