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

REFERENCE_ROOT = '/root/reference'
_installed = False


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
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own package also has a sub-package called `models`; make sure the
    # name resolves to the reference here
    mod = sys.modules.get('models')
    if mod is not None and not getattr(mod, '__file__', '').startswith(REFERENCE_ROOT):
        raise RuntimeError("a different `models` package is already imported")
    import models  # noqa: E402  (the reference's)
    return models


def inject_noise(ref_model, tape):
    """Replace MultiDGTS._sample_gauss (models/dgts.py:177-180) by tape draws."""
    def sample(mean, std):
        return tape(tuple(std.shape)).to(std.dtype).mul(std).add(mean)
    ref_model._sample_gauss = sample
