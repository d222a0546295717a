// Development probe (not part of the library): does tcgen05.mma.kind::tf32 take MN-major SWIZZLE_128B
// operands, and with which descriptor encoding?  This is the open question behind DESIGN §6 item (0a):
// weight-gradient GEMMs dW = dY^T X reduce over ROWS, so with MN-major operand descriptors they can read
// the row-major activations directly and the transposed copies (h1T, dh1T, ...) need not be written.
//
//   C[128][64] = sum_k X[k][m] * Y[k][n]        X: [K][128] row-major,  Y: [K][64] row-major,  K = 32
//
// Shared-memory layout (cute::UMMA canonical MN-major SWIZZLE_128B_BASE32B, layout type 1, in bytes):
//   off(mn, k) = (mn / 32) * ATOM_MN + (k / 4) * ATOM_K + (k % 4) * 128 + (((mn % 32) / 8) ^ (k % 4)) * 32 + (mn % 8) * 4
// (plain SWIZZLE_128B with a_major = MN was tried first: the instruction then returns zeros)
// Variants (argv[1] bit 0): which of LBO / SBO carries ATOM_MN;  the instruction descriptor sets
// a_major (bit 15) and b_major (bit 16) to MN.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_variants/probe_mn_major tools/probe_mn_major.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

namespace {
constexpr int kM = 128, kN = 64, kK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// layout type (bits [61,64)): 2 = SWIZZLE_128B (16-byte atomicity), 1 = SWIZZLE_128B_BASE32B (32-byte atomicity)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type = 2) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}

__global__ void __launch_bounds__(128) probe_kernel(const float* X, const float* Y, float* C, int variant) {
  __shared__ __align__(1024) float sA[kM * kK];      // 4 MN atoms x 4 k-groups x 1024 B
  __shared__ __align__(1024) float sB[kN * kK];      // 2 MN atoms x 4 k-groups
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // MN-major TF32 operands: cutlass sm100_common.inl:92 "for mn-major tf32 operands, SW128_32B is the only available
  // smem layout" = Swizzle<2,5,2> over atoms of 4 k-rows x 128 B (32 floats along MN): the 32-byte chunk index of a
  // line (address bits 5-6) is XORed with the k-row index (bits 7-8)
  constexpr uint32_t kAtomMnA = 512, kAtomKA = (kM / 32) * 512;       // MN atoms of one 4-row k-group are adjacent
  constexpr uint32_t kAtomMnB = 512, kAtomKB = (kN / 32) * 512;

  auto off = [](int mn, int k, uint32_t atom_mn, uint32_t atom_k) -> uint32_t {
    return (uint32_t)(mn / 32) * atom_mn + (uint32_t)(k / 4) * atom_k + (uint32_t)(k % 4) * 128u +
           (uint32_t)((((mn % 32) / 8) ^ (k % 4)) * 32) + (uint32_t)(mn % 8) * 4u;
  };
  const bool a_kmajor = (variant & 8) != 0;          // sanity mode: A in the known-good K-major SW128 layout as well
  for (int i = threadIdx.x; i < kM * kK; i += 128) {
    const int k = i / kM, m = i % kM;
    const uint32_t o = a_kmajor ? (uint32_t)(m / 8) * 1024u + (uint32_t)(m % 8) * 128u + (uint32_t)(((k / 4) ^ (m % 8)) * 16) + (uint32_t)(k % 4) * 4u
                                : off(m, k, kAtomMnA, kAtomKA);
    *reinterpret_cast<float*>(reinterpret_cast<char*>(sA) + o) = X[k * kM + m];
  }
  const bool b_kmajor = (variant & 2) != 0;          // decode mode: B in the known-good K-major SW128 layout
  for (int i = threadIdx.x; i < kN * kK; i += 128) {
    const int k = i / kN, n = i % kN;
    const uint32_t o = b_kmajor ? (uint32_t)(n / 8) * 1024u + (uint32_t)(n % 8) * 128u + (uint32_t)(((k / 4) ^ (n % 8)) * 16) + (uint32_t)(k % 4) * 4u
                                : off(n, k, kAtomMnB, kAtomKB);
    *reinterpret_cast<float*>(reinterpret_cast<char*>(sB) + o) = Y[k * kN + n];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  if (threadIdx.x == 0) {
    // kind::tf32, FP32 accumulate, a_major = b_major = MN
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_kmajor ? 0u : (1u << 15)) | (b_kmajor ? 0u : (1u << 16)) |
                           ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const int ksteps = (variant & 4) ? 1 : kK / 8;      // decode mode: a single K = 8 instruction
    for (int j = 0; j < ksteps; ++j) {
      const uint32_t a0 = smem_u32(sA) + j * 2 * kAtomKA, b0 = smem_u32(sB) + j * 2 * kAtomKB;   // K = 8: two 4-row groups
      const uint64_t da = a_kmajor ? desc_sw128(smem_u32(sA) + j * 32, 16, 1024) : (variant & 1) ? desc_sw128(a0, kAtomKA, kAtomMnA, 1) : desc_sw128(a0, kAtomMnA, kAtomKA, 1);
      const uint64_t db = b_kmajor ? desc_sw128(smem_u32(sB) + j * 32, 16, 1024)
                                   : (variant & 1) ? desc_sw128(b0, kAtomKB, kAtomMnB, 1) : desc_sw128(b0, kAtomMnB, kAtomKB, 1);
      const uint32_t acc = j > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(&mbar)), "r"(0));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < kN; c += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) C[(warp * 32 + lane) * kN + c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64));
}
}  // namespace

int main(int argc, char** argv) {
  std::vector<float> X(kK * kM), Y(kK * kN), Cref(kM * kN, 0.f), Cout(kM * kN);
  srand(7);
  // values exactly representable in TF32 (small integers / 8) so that a correct layout gives an exact match
  for (auto& v : X) v = (float)(rand() % 33 - 16) / 8.f;
  for (auto& v : Y) v = (float)(rand() % 33 - 16) / 8.f;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      float s = 0.f;
      for (int k = 0; k < kK; ++k) s += X[k * kM + m] * Y[k * kN + n];
      Cref[m * kN + n] = s;
    }
  float *dX, *dY, *dC;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, Y.size() * 4); cudaMalloc(&dC, Cout.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dC, 0, Cout.size() * 4);
    probe_kernel<<<1, 128>>>(dX, dY, dC, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(Cout.data(), dC, Cout.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0.0; int bad = 0;
    for (int i = 0; i < kM * kN; ++i) { const double d = fabs((double)Cout[i] - Cref[i]); if (d > worst) worst = d; bad += d > 1e-5; }
    printf("variant %d (%s): max |err| %.6f, mismatching elements %d of %d -> %s\n", variant,
           variant ? "LBO = k-group stride, SBO = MN-atom stride" : "LBO = MN-atom stride, SBO = k-group stride", worst, bad,
           kM * kN, bad == 0 ? "EXACT" : "wrong");
  }
  // ---- decode: what does the hardware read for A(m, k)?  X holds a unique code per element, Y is one-hot in k
  // (K-major, known-good), a single K = 8 instruction: C[m][n] = code the hardware associates with (m, k = n)
  for (int k = 0; k < kK; ++k)
    for (int m = 0; m < kM; ++m) X[k * kM + m] = k < 8 ? (float)(k * kM + m) : 0.f;
  for (int k = 0; k < kK; ++k)
    for (int n = 0; n < kN; ++n) Y[k * kN + n] = (k < 8 && n == k) ? 1.f : 0.f;
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
  for (int swap = 0; swap < 3; ++swap) {            // swap == 2: sanity run, A K-major too
    cudaMemset(dC, 0xff, Cout.size() * 4);
    probe_kernel<<<1, 128>>>(dX, dY, dC, swap == 2 ? (2 | 4 | 8) : (2 | 4 | swap));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("decode %d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(Cout.data(), dC, Cout.size() * 4, cudaMemcpyDeviceToHost);
    int good = 0;
    for (int m = 0; m < kM; ++m) for (int n = 0; n < 8; ++n) good += Cout[m * kN + n] == (float)(n * kM + m);
    printf("decode (A MN-major, B K-major one-hot, swap=%d): %d of %d as assumed\n", swap, good, kM * 8);
    const int ms[] = {0, 1, 3, 4, 5, 8, 31, 32, 33, 64, 127};
    for (int m : ms) {
      printf("  m=%3d:", m);
      for (int n = 0; n < 8; ++n) { const int c = (int)Cout[m * kN + n]; printf(" (k%d,m%3d)", c / kM, c % kM); }
      printf("\n");
    }
  }
  return 0;
}
