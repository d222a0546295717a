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
// Inter-CTA hand-over inside ONE (cooperative, fully resident) launch: a producer warp publishes
// "segments 0..v-1 of this task are done" with a release store, the consumer spins on an acquire
// load; payload written before seg_post is read after seg_wait with ld_cg (L1 is not coherent).
// The emulator never runs several segments in one launch (stream-ordered launches instead).
__device__ __forceinline__ void seg_post(int* flag, int v) {
#ifdef BFVI_EMU
  *flag = v;
#else
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void seg_wait(const int* flag, int want) {
#ifdef BFVI_EMU
  (void)flag; (void)want;
#else
  int v;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= want) break;
    __nanosleep(256);
  }
#endif
}
__device__ __forceinline__ float ld_cg(const float* p) {
#ifdef BFVI_EMU
  return *p;
#else
  return __ldcg(p);
#endif
}

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
