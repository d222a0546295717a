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
