"""Image-module kernels on CPU (emulator library: the same kernel sources as host C++) against torch fp64."""
import ctypes as C

import pytest
import torch

import conv_cases
import helpers
from multimodal_dmm_b200 import _lib


@pytest.fixture(scope='module')
def lib():
    return helpers.emu_library()


# the two largest Weizmann layers take ~50 s of fibre switching in the emulator; the B200 test runs them
EMU_CASES = sorted(set(conv_cases.CONV_CASES) - {'deconv4s2_w1', 'conv3s2_w3'})


@pytest.mark.parametrize('case', EMU_CASES)
def test_conv_layer_matches_torch(lib, case):
    conv_cases.check_conv(case, lib, 'cpu', 2e-6)


def test_deconv_sigmoid_epilogue(lib):
    conv_cases.check_conv('deconv4s2_odd', lib, 'cpu', 2e-6, sigmoid=True)


@pytest.mark.parametrize('case', sorted(conv_cases.BN_CASES))
def test_batchnorm_relu_matches_torch(lib, case):
    conv_cases.check_bn(case, lib, 'cpu', 2e-6)


@pytest.mark.parametrize('case', ['dense_odd_relu', 'dense_odd', 'dense_splitk'])
def test_dense_layer_matches_torch(lib, case):
    conv_cases.check_dense(case, lib, 'cpu', 2e-6)


def test_argument_errors(lib):
    g, _ = conv_cases.geom('conv3s2_odd')
    x = torch.zeros(8)
    with pytest.raises(_lib.BfviError, match='null tensor'):
        lib.call('bfvi_conv_gather', C.byref(g), None, _lib.ptr(x), None, _lib.ptr(x), 0, None)
    g.h_small += 1
    with pytest.raises(_lib.BfviError, match='do not match'):
        lib.call('bfvi_conv_gather', C.byref(g), _lib.ptr(x), _lib.ptr(x), None, _lib.ptr(x), 0, None)
    g, _ = conv_cases.geom('conv3s2_odd')
    g.kernel = 9
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_conv_wgrad', C.byref(g), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None)
    with pytest.raises(_lib.BfviError, match='scratch too small'):
        lib.call('bfvi_chan_bias_grad', _lib.ptr(x), 1, 2, 4, _lib.ptr(x), _lib.ptr(x), C.c_size_t(8), None)


def test_image_encoder_decoder_modules_through_the_kernels(lib, monkeypatch):
    """models.common.ImageEncoder / ImageDecoder with every layer routed through the (emulated) kernels — convolutions,
    BatchNorm2d -> ReLU with running statistics, the dense layers, the fused final sigmoid — against float64 torch."""
    import multimodal_dmm_b200.models.common as common
    monkeypatch.setattr(common, '_library', lambda: lib)
    monkeypatch.setattr(common, '_use_kernels', lambda x, kind: x.dtype == torch.float32)
    worst = conv_cases.check_modules(common, 'cpu', 2e-5)
    assert len(worst) > 30
