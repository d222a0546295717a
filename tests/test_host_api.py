"""Host-side logic that needs no GPU: module tree / state_dict compatibility,
seeded-init equality with the reference, C-ABI symbol export, argument errors."""
import ctypes as C
import os

import pytest
import torch

from conftest import load_golden, ROOT
from multimodal_dmm_b200 import _lib
import multimodal_dmm_b200.models as models


def test_registry_matches_reference():
    assert models.names['dmm'] == 'MultiDMM' and hasattr(models, models.names['dmm'])


def test_state_dict_keys_and_seeded_init_match_reference():
    fx = load_golden('spirals_ragged')
    ref_state = fx['seeded_init_seed1']
    torch.manual_seed(1)
    ours = models.MultiDMM(fx['modalities'], (d for d in fx['dims']), h_dim=fx['h_dim'],
                           z_dim=fx['z_dim'], device=torch.device('cpu'))
    state = ours.state_dict()
    assert list(state.keys()) == list(ref_state.keys())
    for k in ref_state:
        assert state[k].shape == ref_state[k].shape, k
        assert torch.equal(state[k], ref_state[k]), k          # bit-identical initial weights
    assert ours.dims == dict(zip(fx['modalities'], fx['dims']))
    assert ours.h_dim == fx['h_dim'] and ours.z_dim == fx['z_dim']
    ours.load_state_dict(fx['state_dict'])                       # reference checkpoints load


def test_no_cpu_fallback():
    fx = load_golden('single_mod')
    m = models.MultiDMM(fx['modalities'], fx['dims'], h_dim=fx['h_dim'], z_dim=fx['z_dim'],
                        device=torch.device('cpu'))
    with pytest.raises(_lib.BfviError):
        m.step(fx['inputs'], fx['mask'], 1.0, {}, targets=fx['targets'], lengths=fx['lengths'])
    with pytest.raises(_lib.BfviError):
        models.losses.kld_gauss(torch.zeros(2, 3), torch.ones(2, 3), torch.zeros(2, 3), torch.ones(2, 3))


def test_cabi_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    path = g.build()
    lib = _lib.Library(path)                # resolves every entry of _lib.SYMBOLS
    header = open(os.path.join(ROOT, 'include', 'bfvi.h')).read()
    import re
    declared = set(re.findall(r'\b(bfvi_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SYMBOLS.keys()), declared ^ set(_lib.SYMBOLS.keys())
    for name in declared:
        assert hasattr(lib.dll, name)
    assert lib.dll.bfvi_version() == int(re.search(r'#define BFVI_VERSION (\d+)', header).group(1)) == 140
    # the shipped binary is git-ignored: its stamp must match the sources in the tree
    assert lib.build_id() == _lib.source_id() and len(lib.build_id()) == 16


def test_layout_and_argument_errors():
    import __graft_entry__ as g
    lib = _lib.Library(g.build())
    m = _lib.make_model([1, 1], ['Normal', 'Normal'], 5, 20, 1e-3)
    lay = lib.layout(m)
    slots = _lib.param_slots(['a', 'b'], ['Normal', 'Normal'], lay)
    offs = [o for _, o in slots]
    assert offs == sorted(offs) and all(o % 4 == 0 for o in offs)    # 16-byte aligned blocks
    assert lay.total == 1920 and lib.dll.bfvi_kernel_family(C.byref(m)) == 1
    bad = _lib.make_model([1], ['Normal'], 5, 20, 1e-3)
    bad.n_mods = 0
    with pytest.raises(_lib.BfviError):
        lib.layout(bad)
    big = _lib.make_model([4], ['Normal'], 999, 999, 1e-3)
    assert lib.dll.bfvi_kernel_family(C.byref(big)) == 2            # tcgen05 family: step, forward, z_filter
    a = _lib.StepArgs()
    a.T, a.B = 4, 4
    n = C.c_size_t(0)
    with pytest.raises(_lib.BfviError):
        lib.call('bfvi_step_workspace', C.byref(big), C.byref(a), C.byref(n))
