"""Import alias: `import multimodal_dmm_b200` loads the package that lives in the
directory `multimodal-dmm_b200/` (a hyphen is not importable by name)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'multimodal-dmm_b200')]
__package__ = __name__
with open(_os.path.join(__path__[0], '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], '__init__.py'), 'exec'))
del _f
