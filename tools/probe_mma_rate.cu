// Development probe (not part of the library): what does ONE tcgen05.mma instruction cost on B200 as a function of
// kind, N and the A-operand source, when a single thread issues a back-to-back stream of them (the situation of the
// fused GTF kernels, bfvi_fused.cuh)?  Also: the commit -> mbarrier -> wait round trip, and tcgen05.ld / tcgen05.st
// throughput from 8 warps.  Operand contents are irrelevant (shared memory is zero-filled).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_variants/probe_mma_rate tools/probe_mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t* m, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(m)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* m, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(smem_u32(m)), "r"(parity));
}
__device__ __forceinline__ void commit(uint64_t* m) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(m))); }
template <int KIND>   // 0 = f16, 1 = tf32
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc));
}
template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ uint32_t idesc_of(int kind, int M, int N) {
  return (1u << 4) | ((kind ? 2u : 0u) << 7) | ((kind ? 2u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Result { long long cyc[16]; };

template <int KIND, bool TS>
__device__ long long run_stream(uint32_t tb, uint32_t sA, uint32_t sB, int N, int n_mma, uint64_t* bar, uint32_t& parity) {
  const uint32_t idesc = idesc_of(KIND, 128, N);
  const long long t0 = clock64();
  for (int i = 0; i < n_mma; ++i) {
    const uint64_t b = desc_sw128(sB + (i & 3) * 32);
    if (TS) mma_ts<KIND>(tb, tb + 256 + (i & 3) * 8, b, idesc, 1u);
    else mma_ss<KIND>(tb, desc_sw128(sA + (i & 3) * 32), b, idesc, 1u);
  }
  commit(bar);
  mbar_wait(bar, parity);
  parity ^= 1u;
  return clock64() - t0;
}

__global__ void __launch_bounds__(320) probe(Result* out, int n_mma) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_s;
  Result r;
  for (int i = 0; i < 16; ++i) r.cyc[i] = 0;
  if (warp == 8 && lane == 0) {
    uint32_t parity = 0;
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 32768;
    run_stream<0, true>(tb, sA, sB, 64, 8, &bar, parity);                       // warm-up
    r.cyc[0] = run_stream<0, true>(tb, sA, sB, 64, n_mma, &bar, parity);        // f16 TS N=64
    r.cyc[1] = run_stream<0, true>(tb, sA, sB, 128, n_mma, &bar, parity);
    r.cyc[2] = run_stream<0, true>(tb, sA, sB, 256, n_mma, &bar, parity);
    r.cyc[3] = run_stream<0, false>(tb, sA, sB, 64, n_mma, &bar, parity);       // f16 SS
    r.cyc[4] = run_stream<0, false>(tb, sA, sB, 128, n_mma, &bar, parity);
    r.cyc[5] = run_stream<0, false>(tb, sA, sB, 256, n_mma, &bar, parity);
    r.cyc[6] = run_stream<1, true>(tb, sA, sB, 64, n_mma, &bar, parity);        // tf32 TS
    r.cyc[7] = run_stream<1, true>(tb, sA, sB, 128, n_mma, &bar, parity);
    r.cyc[8] = run_stream<1, true>(tb, sA, sB, 256, n_mma, &bar, parity);
    // round trip: 16 x (12 MMAs N=64, commit, wait)
    const long long t0 = clock64();
    for (int rep = 0; rep < 16; ++rep) run_stream<0, true>(tb, sA, sB, 64, 12, &bar, parity);
    r.cyc[9] = (clock64() - t0) / 16;
    // empty commit round trip
    const long long t1 = clock64();
    for (int rep = 0; rep < 16; ++rep) { commit(&bar); mbar_wait(&bar, parity); parity ^= 1u; }
    r.cyc[10] = (clock64() - t1) / 16;
  }
  __syncthreads();
  // tcgen05.ld / st throughput: 8 warps, each 16 x (32 lanes x 32 columns)
  if (warp < 8) {
    const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
    uint32_t v[32];
    uint32_t accv = 0;
    asm volatile("bar.sync 1, 256;");
    const long long t0 = clock64();
    for (int it = 0; it < 16; ++it) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tl + (uint32_t)((it & 3) * 64)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) accv ^= v[j];
    }
    asm volatile("bar.sync 1, 256;");
    const long long t1 = clock64();
    for (int it = 0; it < 16; ++it) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
          "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tl + (uint32_t)((it & 3) * 64)),
          "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
          "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
          "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
          "r"(v[31])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("bar.sync 1, 256;");
    const long long t2 = clock64();
    if (threadIdx.x == 0) { r.cyc[11] = (t1 - t0) / 16; r.cyc[12] = (t2 - t1) / 16; r.cyc[13] = accv & 1; }
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[1].cyc[11] = r.cyc[11]; out[1].cyc[12] = r.cyc[12]; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8 && lane == 0 && blockIdx.x == 0) { for (int i = 0; i < 11; ++i) out[0].cyc[i] = r.cyc[i]; }
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
}  // namespace

int main(int argc, char** argv) {
  const int n_mma = argc > 1 ? atoi(argv[1]) : 96;
  Result* d;
  cudaMalloc(&d, 2 * sizeof(Result));
  cudaMemset(d, 0, 2 * sizeof(Result));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  for (int blocks : {1, 148}) {
    probe<<<blocks, 320, 66560>>>(d, n_mma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    Result h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"f16 TS N=64", "f16 TS N=128", "f16 TS N=256", "f16 SS N=64", "f16 SS N=128", "f16 SS N=256",
                           "tf32 TS N=64", "tf32 TS N=128", "tf32 TS N=256"};
    printf("== %d CTA(s), %d back-to-back MMAs (M=128) + commit + wait: total cycles, cycles per MMA\n", blocks, n_mma);
    for (int i = 0; i < 9; ++i) printf("  %-14s %8lld  %7.1f\n", names[i], h[0].cyc[i], (double)h[0].cyc[i] / n_mma);
    printf("  12 x (f16 TS N=64) + commit + wait round trip: %lld cycles\n", h[0].cyc[9]);
    printf("  empty commit + wait round trip: %lld cycles\n", h[0].cyc[10]);
    printf("  8 warps x tcgen05.ld 32x32b.x32 (32 KB per round) + wait: %lld cycles per round; st: %lld\n", h[1].cyc[11], h[1].cyc[12]);
  }
  return 0;
}
