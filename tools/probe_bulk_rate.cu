// Development probe: throughput of cp.async.bulk (global -> shared, mbarrier complete_tx) per SM when the source is
// L2-resident (a 544 KB weight pack re-read by every CTA, as in bfvi_fused.cuh), as a function of the copy size and of
// how many threads issue the pieces of a 32 KB block.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_variants/probe_bulk_rate tools/probe_bulk_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* m, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(m)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* m, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(smem_u32(m)), "r"(parity));
}
__device__ __forceinline__ void expect_tx(uint64_t* m, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(m)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* m) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(m)) : "memory");
}
constexpr int kStages = 6, kBlock = 32768;
// pieces: the 32 KB block is issued as `pieces` copies by `pieces` lanes of one warp (1, 2, 4, 8, 16, 32)
__global__ void __launch_bounds__(64) probe(const unsigned char* src, int n_src_blocks, int n_blocks, int pieces, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[kStages];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 0) {
    // keep kStages blocks in flight: issue block g, wait block g - kStages + 1 ... (consumer = same warp, no compute)
    for (int g = 0; g < n_blocks + kStages - 1; ++g) {
      if (g < n_blocks) {
        const int s = g % kStages;
        if (lane == 0) expect_tx(&full[s], kBlock);
        __syncwarp();
        if (lane < pieces) {
          const uint32_t sz = kBlock / pieces;
          bulk(smem + (size_t)s * kBlock + lane * sz, src + (size_t)(g % n_src_blocks) * kBlock + lane * sz, sz, &full[s]);
        }
      }
      const int w = g - (kStages - 1);
      if (w >= 0) mbar_wait(&full[w % kStages], (uint32_t)((w / kStages) & 1));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
}  // namespace
int main() {
  const int n_src = 17, n_blocks = 17 * 8;
  unsigned char* src; long long* out;
  cudaMalloc(&src, (size_t)n_src * kBlock); cudaMemset(src, 1, (size_t)n_src * kBlock);
  cudaMalloc(&out, 8);
  const int smem = kStages * kBlock + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int blocks : {1, 18, 148}) for (int pieces : {1, 2, 4, 8, 16, 32}) {
    probe<<<blocks, 64, smem>>>(src, n_src, n_blocks, pieces, out);
    cudaDeviceSynchronize();
    probe<<<blocks, 64, smem>>>(src, n_src, n_blocks, pieces, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("CTAs %3d  pieces %2d (%5d B each): %8lld cycles for %d x 32 KB  -> %.1f B/clk/SM, %.0f cycles per block\n", blocks, pieces, kBlock / pieces, c, n_blocks, (double)n_blocks * kBlock / c, (double)c / n_blocks);
  }
  return 0;
}
