"""Drop-in replacement for the reference `models` package (models/__init__.py:1-6)
restricted to the BFVI hot path: `names['dmm'] -> 'MultiDMM'` is what
trainer.py:193-199 resolves."""
from .dgts import MultiDGTS
from .dmm import MultiDMM
from . import common, losses

names = {'dmm': 'MultiDMM'}

__all__ = ['MultiDGTS', 'MultiDMM', 'common', 'losses', 'names']
