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

namespace bfvi {
// An integer zero the compiler cannot see through.  Adding it to the base of the
// shared-memory weight block inside the time loop stops loop-invariant hoisting of
// the ~130 weight loads of one GTF evaluation (which would otherwise be "kept in
// registers" across iterations, i.e. spilled to local memory).
__device__ __forceinline__ int opaque_zero() {
#ifdef BFVI_EMU
  return 0;
#else
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return z;
#endif
}
}  // namespace bfvi
