"""bfvi_mlp_fwd / _bwd on CPU (emulator library: exact fp32 GEMM stand-in, the same elementwise kernels and launch
sequence) against the fp64 torch restatement."""
import pytest

import helpers
import mlp_cases


@pytest.fixture(scope='module')
def lib():
    return helpers.emu_library()


@pytest.mark.parametrize('name', ['gauss_enc_small', 'gauss_dec_small', 'softmax_dec', 'embed_enc'])
def test_mlp_matches_torch(lib, name):
    mlp_cases.check(name, lib, 'cpu', 2e-5)
