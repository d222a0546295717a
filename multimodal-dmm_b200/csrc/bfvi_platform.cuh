// bfvi_platform.cuh — build-mode switch.
//   nvcc (product):  real CUDA for sm_100a.
//   -DBFVI_EMU (tests/emu only): the same kernel sources compiled as host C++
//   against tests/emu/cuda_emu.h, so kernel logic can be checked on a CPU-only
//   box.  The emulated library is never shipped or loaded by the package.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef BFVI_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define BFVI_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BFVI_DYN_SMEM(type, name) \
  extern __shared__ __align__(16) unsigned char bfvi_dyn_smem_raw[]; \
  type* name = reinterpret_cast<type*>(bfvi_dyn_smem_raw)
#endif
