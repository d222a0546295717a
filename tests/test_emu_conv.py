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


def test_layer_options_of_the_autograd_functions(lib, monkeypatch):
    """bias-free convolutions, a BatchNorm2d without affine parameters and with a cumulative moving average
    (momentum=None), a BatchNorm2d that tracks no running statistics: the module options _ConvFn / _BatchNormFn accept."""
    import torch.nn as nn
    import multimodal_dmm_b200.models.common as common
    monkeypatch.setattr(common, '_library', lambda: lib)
    torch.manual_seed(11)
    conv = nn.Conv2d(3, 5, 3, 2, 1, bias=False)
    deconv = nn.ConvTranspose2d(5, 2, 4, 2, 1, bias=False)
    bn_cma = nn.BatchNorm2d(5, affine=False, momentum=None)
    bn_free = nn.BatchNorm2d(2, track_running_stats=False)
    ref = [__import__('copy').deepcopy(m).double() for m in (conv, deconv, bn_cma, bn_free)]
    for step in range(2):
        x = torch.randn(4, 3, 8, 8)
        r = torch.randn(4, 2, 8, 8)              # a random projection: sum(y^2) of a normalised y has a vanishing gradient
        # kernels
        k, s, p, tr = common._layer_geometry(conv)
        y = common._ConvFn.apply(x.clone().requires_grad_(True), conv.weight, None, k, s, p, tr, False)
        y = common._BatchNormFn.apply(y, None, None, bn_cma, True)
        k, s, p, tr = common._layer_geometry(deconv)
        y = common._ConvFn.apply(y, deconv.weight, None, k, s, p, tr, False)
        y = common._BatchNormFn.apply(y, bn_free.weight, bn_free.bias, bn_free, False)
        for m in (conv, deconv, bn_free):
            m.zero_grad()
        (y * r).sum().backward()
        # torch fp64
        y64 = ref[3](ref[1](torch.relu(ref[2](ref[0](x.double())))))
        for m in ref:
            m.zero_grad()
        (y64 * r.double()).sum().backward()
        assert conv_cases.rel(y, y64) < 2e-6
        assert conv_cases.rel(conv.weight.grad, ref[0].weight.grad) < 2e-5
        assert conv_cases.rel(deconv.weight.grad, ref[1].weight.grad) < 2e-5
        assert conv_cases.rel(bn_free.weight.grad, ref[3].weight.grad) < 2e-5
        assert conv_cases.rel(bn_cma.running_mean, ref[2].running_mean) < 2e-6
        assert conv_cases.rel(bn_cma.running_var, ref[2].running_var) < 2e-6
        assert int(bn_cma.num_batches_tracked) == step + 1


def test_unsupported_layer_options_fail_loudly():
    import torch.nn as nn
    import multimodal_dmm_b200.models.common as common
    for layer in (nn.Conv2d(3, 4, 3, dilation=2), nn.Conv2d(4, 4, 3, groups=2), nn.Conv2d(3, 4, (3, 5)),
                  nn.ConvTranspose2d(3, 4, 4, 2, 1, output_padding=1), nn.Conv2d(3, 4, 3, padding='same')):
        with pytest.raises(_lib.BfviError):
            common._layer_geometry(layer)


def test_image_modules_match_the_reference_golden(lib, monkeypatch):
    """the reference's own ImageEncoder / ImageDecoder (float64, oracle/make_golden_image.py) against ours through the
    emulated kernels: identical seeded initial values, then outputs / gradients / running statistics of two training
    passes and the evaluation pass"""
    import multimodal_dmm_b200.models.common as common
    monkeypatch.setattr(common, '_library', lambda: lib)
    monkeypatch.setattr(common, '_use_kernels', lambda x, kind: True)
    worst = conv_cases.check_golden_image(common, 'cpu', 2e-5, check_init=True)
    assert len(worst) > 60


def test_deterministic_mode(lib, monkeypatch):
    """BFVI_DETERMINISTIC=1 (one writer per output element: no pixel splits, no K slices) computes the same results; under
    torch.use_deterministic_algorithms(True) the autograd functions insist on it"""
    import multimodal_dmm_b200.models.common as common
    monkeypatch.setenv('BFVI_DETERMINISTIC', '1')
    conv_cases.check_conv('conv3s2_odd', lib, 'cpu', 2e-6)
    conv_cases.check_dense('dense_splitk', lib, 'cpu', 2e-6)
    monkeypatch.delenv('BFVI_DETERMINISTIC')
    monkeypatch.setattr(common, '_library', lambda: lib)
    torch.use_deterministic_algorithms(True)
    try:
        with pytest.raises(RuntimeError, match='BFVI_DETERMINISTIC'):
            common._DenseFn.apply(torch.zeros(2, 3), torch.zeros(4, 3), None, False)
        monkeypatch.setenv('BFVI_DETERMINISTIC', '1')
        assert common._DenseFn.apply(torch.ones(2, 3), torch.ones(4, 3), None, False).sum().item() == 24.0
    finally:
        torch.use_deterministic_algorithms(False)
