"""Fused Adam (+ gradient-norm clipping) over the flat parameter buffer of a MultiDMM.

The step either side of the hot path (trainer.py:212-213, 248-252): `torch.optim.Adam` walks ~50
parameter tensors with several launches each, which is a visible fraction of a step once the
BFVI step itself takes milliseconds.  Every default parameter of a MultiDMM aliases ONE flat
buffer and `loss.backward()` leaves ONE flat gradient (already all-reduced in data-parallel
runs), so the optimiser is two kernels (`bfvi_adam_step`).  Parameters of custom user modules
(conv encoders / decoders) are not in the flat buffer: give them to a torch optimiser.
"""
import ctypes as C

import torch

from . import _lib


class FlatAdam(object):
    """Drop-in for `torch.optim.Adam(model.parameters(), lr, weight_decay=wd)` +
    `clip_grad_norm_(model.parameters(), max_norm)` on an all-default MultiDMM."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=None):
        model._ensure_flat()
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.exp_avg = torch.zeros_like(model._flat)
        self.exp_avg_sq = torch.zeros_like(model._flat)
        self._norm = torch.zeros(1, device=model._flat.device)
        self.t = 0

    def step(self, flat_grad=None):
        """Consumes the flat gradient of the last `loss.backward()` (model.last_flat_grad)."""
        g = self.model.last_flat_grad if flat_grad is None else flat_grad
        if g is None:
            raise _lib.BfviError('FlatAdam.step(): no flat gradient (call loss.backward() on a step() loss first)')
        self.t += 1
        flat = self.model._flat
        _lib.load().call('bfvi_adam_step', _lib.ptr(flat), _lib.ptr(g), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                         flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.t, 1.0,
                         float(self.max_norm) if self.max_norm else 0.0, _lib.ptr(self._norm),
                         C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None
        self.model.last_flat_grad = None
