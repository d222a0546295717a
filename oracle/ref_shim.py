"""Import the UNMODIFIED reference (ztangent/multimodal-dmm) in this container.

TEST INFRASTRUCTURE ONLY (see oracle/bfvi_oracle.py).  Used by
oracle/make_golden.py to pin the oracle and to generate tests/golden/*.pt.
`/root/reference` does not exist on the GPU box, so nothing that runs there
imports this module.

The reference targets torch 1.1; under torch >= 1.2 the idiom `1 - <bool tensor>`
(models/dmm.py:165, models/dgts.py:45,76, models/losses.py:35-82) raises.  The
shim below maps exactly that expression to logical-not and leaves every other
subtraction alone.  No reference file is modified or copied.
"""
import sys

import torch

import os
import types

REFERENCE_ROOT = '/root/reference'
# sourceless bytecode of the same modules, built by oracle/build_ref.py (travels to the GPU box)
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
_installed = False


def reference_root():
    """Where the reference's modules are imported from: the read-only checkout when it exists (this container),
    else the byte-compiled copy under oracle/_ref/ (GPU box); None when neither exists."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, 'models')):
        return REFERENCE_ROOT
    if os.path.exists(os.path.join(COMPILED_ROOT, 'models', 'dmm' + COMPILED_EXT)):
        _install_compiled_finder()
        return COMPILED_ROOT
    return None


COMPILED_EXT = '.pybc'
_finder = None


def _install_compiled_finder():
    """Importer for the sourceless bytecode under oracle/_ref/ (`<module>.pybc` = the .pyc format): resolves the
    reference's top-level modules (models, datasets, trainer, spirals, utils) and their submodules."""
    global _finder
    if _finder is not None:
        return
    import importlib.abc
    import importlib.machinery
    import importlib.util

    class Finder(importlib.abc.MetaPathFinder):
        def find_spec(self, fullname, path=None, target=None):
            rel = os.path.join(COMPILED_ROOT, *fullname.split('.'))
            pkg = os.path.join(rel, '__init__' + COMPILED_EXT)
            if os.path.exists(pkg):
                return importlib.util.spec_from_file_location(
                    fullname, pkg, loader=importlib.machinery.SourcelessFileLoader(fullname, pkg),
                    submodule_search_locations=[rel])
            if os.path.exists(rel + COMPILED_EXT):
                return importlib.util.spec_from_file_location(
                    fullname, rel + COMPILED_EXT,
                    loader=importlib.machinery.SourcelessFileLoader(fullname, rel + COMPILED_EXT))
            return None

    _finder = Finder()
    sys.meta_path.insert(0, _finder)


def install():
    global _installed
    if _installed:
        return
    orig_rsub = torch.Tensor.__rsub__

    def rsub(self, other):
        if self.dtype == torch.bool and isinstance(other, int) and other == 1:
            return ~self
        return orig_rsub(self, other)

    torch.Tensor.__rsub__ = rsub
    _installed = True


def import_reference_models():
    """Returns the reference's `models` package (models/__init__.py)."""
    install()
    root = reference_root()
    if root is None:
        raise RuntimeError('reference not available: neither %s nor %s exists' % (REFERENCE_ROOT, COMPILED_ROOT))
    if root not in sys.path:
        sys.path.insert(0, root)
    # our own package also has a sub-package called `models`; make sure the
    # name resolves to the reference here
    mod = sys.modules.get('models')
    if mod is not None and not (getattr(mod, '__file__', None) or '').startswith(root):
        raise RuntimeError("a different `models` package is already imported")
    import models  # noqa: E402  (the reference's)
    return models


def stub_matplotlib():
    """utils.py:8-9 and spirals.py:11-12 import matplotlib (absent from this image) for plotting only: empty
    stand-in modules let trainer.py / spirals.py import; nothing on the training / evaluation path calls them."""
    if 'matplotlib' in sys.modules:
        return
    names = {'matplotlib': [], 'matplotlib.pyplot': [], 'matplotlib.lines': ['Line2D'],
             'matplotlib.collections': ['EllipseCollection']}
    for name, attrs in names.items():
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, type(a, (object,), {}))
        sys.modules[name] = mod
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


def import_reference_trainer():
    """Returns the reference's (spirals, trainer) modules (spirals.py, trainer.py) with `models` = the reference's."""
    import_reference_models()
    stub_matplotlib()
    import spirals  # noqa: E402  (the reference's)
    import trainer  # noqa: E402
    return spirals, trainer


def inject_noise(ref_model, tape):
    """Replace MultiDGTS._sample_gauss (models/dgts.py:177-180) by tape draws."""
    def sample(mean, std):
        return tape(tuple(std.shape)).to(std.dtype).mul(std).add(mean)
    ref_model._sample_gauss = sample
