"""Image-module kernels on B200 (bfvi_conv_*, bfvi_bn2d_*, bfvi_sigmoid_bwd, bfvi_chan_bias_grad; the dense layers on
bfvi_dense_fwd / _bwd) against torch fp64 on the CPU: every layer kind of models/common.py:70-175 at the
Weizmann sizes and at odd sizes, then the ImageEncoder / ImageDecoder modules end to end."""
import os

import pytest

import conv_cases
import multimodal_dmm_b200.models.common as common
from multimodal_dmm_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', sorted(conv_cases.CONV_CASES))
def test_conv_layer_matches_torch(case):
    conv_cases.check_conv(case, _lib.load(), 'cuda:0', 1e-5)


@pytest.mark.parametrize('case', ['deconv4s2_odd', 'deconv4s2_w3'])
def test_deconv_sigmoid_epilogue(case):
    conv_cases.check_conv(case, _lib.load(), 'cuda:0', 1e-5, sigmoid=True)


@pytest.mark.parametrize('case', sorted(conv_cases.BN_CASES))
def test_batchnorm_relu_matches_torch(case):
    conv_cases.check_bn(case, _lib.load(), 'cuda:0', 1e-5)


@pytest.mark.parametrize('case', sorted(conv_cases.DENSE_CASES))
def test_dense_layer_matches_torch(case):
    conv_cases.check_dense(case, _lib.load(), 'cuda:0', 2e-6)


@pytest.mark.parametrize('size', ['small', 'weizmann'])
def test_image_encoder_decoder_modules(size):
    """every layer through this library (no cuDNN / cuBLAS): outputs, gradients, running statistics, evaluation mode"""
    if 'BFVI_IMAGE_KERNELS' not in os.environ:                  # the measurement aid of tools/time_conv.py
        assert common.IMAGE_KERNELS == {'conv': True, 'dense': True}
    kw = dict(img_size=16, n_kernels=8, z_dim=12, frames=5) if size == 'small' else \
        dict(img_size=64, n_kernels=64, z_dim=256, frames=6)
    worst = conv_cases.check_modules(common, 'cuda:0', 1e-4, **kw)
    assert len(worst) > 30


def test_image_modules_match_the_reference_golden():
    """tests/golden/image/modules.pt: the UNMODIFIED reference's ImageEncoder / ImageDecoder in float64"""
    worst = conv_cases.check_golden_image(common, 'cuda:0', 2e-4)
    assert len(worst) > 60


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
