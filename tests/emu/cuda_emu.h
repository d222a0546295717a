// cuda_emu.h — a tiny single-threaded SIMT emulator (DEVELOPMENT / TEST TOOL ONLY).
//
// The build container has no GPU and every `gpurun` call costs minutes, so the
// kernels in multimodal-dmm_b200/csrc are ALSO compiled as plain C++ against this
// shim (tests/emu/build_emu.sh -> tests/emu/_build/libbfvi_emu.so) and the
// `-m "not gpu"` tests run them on tiny shapes against the oracle.  This checks
// kernel LOGIC (indexing, reductions, gradient math, workspace carving) before
// GPU time is spent.  It is not a product path: the shipped package only ever
// loads the nvcc-built libbfvi_b200.so and raises without a CUDA device.
//
// Model: every CUDA thread of a block is a ucontext fiber; blocks run one after
// another; __syncthreads / warp collectives are cooperative barriers that yield
// to the scheduler.  All 32 lanes of a warp must reach every warp collective and
// all threads of a block every __syncthreads (the kernels are written that way).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

#define BFVI_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
using std::min;
using std::max;
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, int,
                                            cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memcpy((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
typedef int cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace emu {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
};

struct State {
  ucontext_t main_ctx;
  std::vector<Fiber> fibers;
  std::function<void()> body;
  int nthreads = 0, cur = 0;
  dim3 grid, block;
  uint3 bidx;
  // block barrier
  int bar_count = 0; unsigned bar_gen = 0;
  // per-warp barrier + exchange
  std::vector<int> wbar_count; std::vector<unsigned> wbar_gen;
  std::vector<uint64_t> xchg;   // [warp][2][32]
  std::vector<unsigned char> dyn_smem;
  unsigned long progress = 0;
};
inline State& st() { static State s; return s; }
static const size_t kStack = 256 * 1024;

inline void yield_to_main() {
  State& s = st();
  swapcontext(&s.fibers[s.cur].ctx, &s.main_ctx);
}

inline void trampoline() {
  State& s = st();
  s.body();
  s.fibers[s.cur].done = true;
  s.progress++;
  swapcontext(&s.fibers[s.cur].ctx, &s.main_ctx);
}

inline void block_barrier() {
  State& s = st();
  unsigned gen = s.bar_gen;
  s.progress++;
  if (++s.bar_count == s.nthreads) { s.bar_count = 0; s.bar_gen++; return; }
  while (s.bar_gen == gen) yield_to_main();
}

inline void warp_barrier() {
  State& s = st();
  int w = s.cur / 32;
  unsigned gen = s.wbar_gen[w];
  s.progress++;
  int lanes = std::min(32, s.nthreads - w * 32);
  if (++s.wbar_count[w] == lanes) { s.wbar_count[w] = 0; s.wbar_gen[w]++; return; }
  while (s.wbar_gen[w] == gen) yield_to_main();
}

// exchange: every lane deposits a 64-bit payload, then reads lane `src`'s.
inline uint64_t warp_exchange(uint64_t mine, int src) {
  State& s = st();
  int w = s.cur / 32, lane = s.cur % 32;
  unsigned par = s.wbar_gen[w] & 1;
  s.xchg[(size_t)(w * 2 + par) * 32 + lane] = mine;
  warp_barrier();
  return s.xchg[(size_t)(w * 2 + par) * 32 + (src & 31)];
}

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& fn) {
  State& s = st();
  s.grid = grid; s.block = block;
  s.nthreads = (int)(block.x * block.y * block.z);
  if ((int)s.fibers.size() < s.nthreads) {
    size_t old = s.fibers.size();
    s.fibers.resize(s.nthreads);
    for (size_t i = old; i < s.fibers.size(); ++i) s.fibers[i].stack = (char*)malloc(kStack);
  }
  int nwarps = (s.nthreads + 31) / 32;
  s.wbar_count.assign(nwarps, 0); s.wbar_gen.assign(nwarps, 0);
  s.xchg.assign((size_t)nwarps * 64, 0);
  s.dyn_smem.assign(smem_bytes + 64, 0xCD);   // poison: catches reads of unwritten smem
  s.body = fn;
  for (unsigned bz = 0; bz < grid.z; ++bz)
  for (unsigned by = 0; by < grid.y; ++by)
  for (unsigned bx = 0; bx < grid.x; ++bx) {
    s.bidx = uint3{bx, by, bz};
    s.bar_count = 0;
    std::fill(s.wbar_count.begin(), s.wbar_count.end(), 0);
    for (int i = 0; i < s.nthreads; ++i) {
      Fiber& f = s.fibers[i];
      f.done = false;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    int alive = s.nthreads;
    while (alive > 0) {
      unsigned long before = s.progress;
      alive = 0;
      for (int i = 0; i < s.nthreads; ++i) {
        if (s.fibers[i].done) continue;
        s.cur = i;
        swapcontext(&s.main_ctx, &s.fibers[i].ctx);
        if (!s.fibers[i].done) alive++;
      }
      if (alive > 0 && s.progress == before) {
        fprintf(stderr, "cuda_emu: deadlock (a barrier / warp collective was not reached by all threads)\n");
        abort();
      }
    }
  }
}

struct Tid {
  operator uint3() const { return get(); }
  static uint3 get() {
    State& s = st();
    unsigned i = (unsigned)s.cur;
    return uint3{i % s.block.x, (i / s.block.x) % s.block.y, i / (s.block.x * s.block.y)};
  }
};
struct TidProxy { struct C { int which; operator unsigned() const { uint3 t = Tid::get(); return which == 0 ? t.x : which == 1 ? t.y : t.z; } }; C x{0}, y{1}, z{2}; };
struct BidProxy { struct C { int which; operator unsigned() const { uint3 t = st().bidx; return which == 0 ? t.x : which == 1 ? t.y : t.z; } }; C x{0}, y{1}, z{2}; };
struct BdimProxy { struct C { int which; operator unsigned() const { dim3 t = st().block; return which == 0 ? t.x : which == 1 ? t.y : t.z; } }; C x{0}, y{1}, z{2}; };
struct GdimProxy { struct C { int which; operator unsigned() const { dim3 t = st().grid; return which == 0 ? t.x : which == 1 ? t.y : t.z; } }; C x{0}, y{1}, z{2}; };
inline void* dyn_smem_base() {
  uintptr_t p = (uintptr_t)st().dyn_smem.data();
  return (void*)((p + 15) & ~(uintptr_t)15);
}
}  // namespace emu

static emu::TidProxy threadIdx;
static emu::BidProxy blockIdx;
static emu::BdimProxy blockDim;
static emu::GdimProxy gridDim;

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T> static inline uint64_t emu_pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T emu_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
static inline int emu_lane() { return emu::st().cur % 32; }

template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_unpack<T>(emu::warp_exchange(emu_pack(v), src)); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_unpack<T>(emu::warp_exchange(emu_pack(v), emu_lane() ^ m)); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  int src = emu_lane() + (int)d; if (src > 31) src = emu_lane();
  return emu_unpack<T>(emu::warp_exchange(emu_pack(v), src));
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  // 32 exchanges would be slow; use one exchange per lane via a gather loop
  emu::State& s = emu::st();
  int w = s.cur / 32, lane = s.cur % 32;
  unsigned par = s.wbar_gen[w] & 1;
  s.xchg[(size_t)(w * 2 + par) * 32 + lane] = pred ? 1 : 0;
  emu::warp_barrier();
  int lanes = std::min(32, s.nthreads - w * 32);
  for (int i = 0; i < lanes; ++i) r |= (unsigned)(s.xchg[(size_t)(w * 2 + par) * 32 + i] & 1) << i;
  return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }

static inline float atomicAdd(float* a, float v) { float o = *a; *a = o + v; return o; }
static inline double atomicAdd(double* a, double v) { double o = *a; *a = o + v; return o; }
static inline int atomicAdd(int* a, int v) { int o = *a; *a = o + v; return o; }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { unsigned o = *a; *a = o + v; return o; }
static inline unsigned atomicMax(unsigned* a, unsigned v) { unsigned o = *a; *a = o > v ? o : v; return o; }

template <typename T> static inline T __ldg(const T* p) { return *p; }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline void emu_sincosf(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#define __sincosf(x, s, c) emu_sincosf(x, s, c)
static inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979323846f * x); *c = cosf(3.14159265358979323846f * x); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __saturatef(float x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

#define BFVI_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define BFVI_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem_base())
