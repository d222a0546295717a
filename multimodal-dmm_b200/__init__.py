"""B200-native BFVI training step for the Multimodal Deep Markov Model.

Drop-in for the `models` package of ztangent/multimodal-dmm on ONE hot path: the
MultiDMM BFVI ELBO forward+backward step (`MultiDMM.step` + `.backward()`), its
`forward` / `sample` inference API and the `models/losses.py` functions.  The
compute runs in hand-written CUDA for sm_100a behind the C ABI of
`include/bfvi.h` (built in-tree as `libbfvi_b200.so`); this package is the thin
Python host side.  There is no CPU fallback.
"""
__version__ = '0.1.0'
