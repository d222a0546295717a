// bfvi_tc.cuh — tcgen05 (5th-generation tensor core) building blocks for the large-dim
// kernel family: a TF32 GEMM tile kernel with a fused bias / activation epilogue,
//
//     C[M, N] = act( A[M, K] · W[N, K]^T + bias[N] ),
//
// which is exactly an nn.Linear (weights stay in PyTorch's (out, in) row-major layout,
// i.e. the K-major B operand of the MMA — no transposition anywhere).  Dense
// contractions of the BFVI step at large batch — GaussianMLP encoder / decoder layers
// (models/common.py:38-41) and the six GaussianGTF layers evaluated for all particles of
// a time step (models/common.py:62-68) — run through it.
//
// Precision: kind::tf32 with an FP32 accumulator in TMEM.  Operands are rounded to TF32
// with round-to-nearest while they are staged into shared memory; letting the
// MMA truncate raw FP32 bits would bias every product low by ~2^-11 and break the 1e-4
// ELBO tolerance (rounded TF32 meets 1e-4 / 1e-3, plain BF16 does not).
//
// Two kernels live here:
//  * gemm_tf32_kernel (round-1, BFVI_GEMM_V1=1): one 128 x BN tile per 128-thread CTA, operands staged
//    ld.global -> cvt.rna -> st.shared in the canonical no-swizzle K-major UMMA layout, double buffered.
//    Kept as the measured baseline of tools/time_gemm.py.
//  * gemm_tf32_p_kernel (the product): persistent, grouped (several independent GEMMs per launch),
//    warp-specialised — 8 converter warps (cp.async ring + rounding), 1 MMA warp, 8 epilogue warps —
//    SWIZZLE_128B operand tiles, up to 8 accumulators in TMEM.  See the banner above it.
//
// Both: tcgen05.mma cta_group::1, M = 128, N = BN, K = 8 per instruction, issued by one elected
// thread; tcgen05.commit to mbarriers; accumulators read back with tcgen05.ld (32 lanes x 32 columns).
//
// Round-1 kernel's shared-memory operand layout: element (row r, float k) of a [ROWS x 32] chunk lives at
// byte (k/4) * ROWS*16 + r*16 + (k%4)*4 — an array [k/4][r] of 16-byte vectors.  In UMMA
// terms: 8-row x 16-byte core matrices, contiguous along rows (SBO = 128 B), K chunks
// ROWS*16 B apart (LBO).  A warp storing 32 consecutive rows writes 512 contiguous bytes.
#pragma once
#include "bfvi_platform.cuh"

namespace bfvi {
namespace tc {

constexpr int kBM = 128;          // rows per CTA tile (UMMA M)
constexpr int kBK = 32;           // floats of K per shared-memory stage (4 MMA k-steps)
constexpr int kThreads = 128;

enum { ACT_NONE = 0, ACT_RELU = 1 };

struct GemmParams {
  // C[M,N] = A[M,K] · W[N,K]^T: both operands K-major.  Forward and input-gradient layers use
  // it directly; a weight gradient dW = dY^T X is the same form on the TRANSPOSED copies
  // dY^T [out, rows], X^T [in, rows] that the producing kernels write next to dY and X.
  const float* A; int64_t lda;
  const float* W; int64_t ldw;
  const float* bias;                // [N] or null, added before the activation
  float* C; int64_t ldc;            // [M, N] row-major
  int64_t M;
  int N;
  int64_t K;
  int act;                          // ACT_*
  int accumulate;                   // C += result instead of C = result
  float* Ct; int64_t ldct;          // optional transposed copy of the result: Ct[n][m]  (nullable)
  int64_t k_split;                  // > 0: blockIdx.z owns K range [z * k_split, (z+1) * k_split) and ADDS its
                                    // partial tile into C with atomics (weight gradients: few output tiles,
                                    // long contraction over the rows); requires accumulate semantics
  const float* mask_aux;            // epilogue: result *= (mask_aux[m][n] > 0)  (ReLU backward), nullable
  int64_t ldaux;
  float* colsum;                    // epilogue: colsum[n] += sum_m result[m][n] (bias gradients), nullable
  int trans_out;                    // 1: ADD the product into C TRANSPOSED, C[n * ldc + m] += result[m][n] (atomics; no
                                    // bias / activation / mask).  Weight gradients with few outputs and many inputs
                                    // run as dW^T = X^T dY: the long side fills the 128 MMA rows instead of padding them
};
enum { PREC_TF32X3 = 0, PREC_TF32 = 1 };   // operand precision of the large-dim family

#ifndef BFVI_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bits:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor for kind::tf32, FP32 accumulate, both operands K-major
// (cute::UMMA::InstrDescriptor: c_format=F32 [4,6), a/b_format=TF32 [7,10)/[10,13),
//  N>>3 [17,23), M>>4 [24,29))
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
      smem_u32(mbar)));
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"(smem_u32(mbar)),
      "r"(parity));
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {       // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {          // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
// 32 lanes x 32 consecutive fp32 columns: thread (lane) gets its row's 32 values
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one [ROWS x 32] K-chunk of a row-major matrix -> registers (zero outside the matrix)
// thread t owns 16-byte vectors v = t, t + 128, ... of the chunk; vector v = (row v % ROWS, k4 v / ROWS)
template <int ROWS>
__device__ __forceinline__ void load_chunk(const float* __restrict__ P, int64_t ld, int64_t row0, int64_t n_rows,
                                           int64_t k0, int64_t K, bool vec_ok,
                                           float4 (&reg)[ROWS * 8 / kThreads]) {
  constexpr int NV = ROWS * 8 / kThreads;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = threadIdx.x + i * kThreads;
    const int r = v % ROWS;
    const int64_t k = k0 + (v / ROWS) * 4;
    const int64_t row = row0 + r;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n_rows && k < K) {
      const float* src = P + row * ld + k;
      if (vec_ok && k + 3 < K) {
        x = *reinterpret_cast<const float4*>(src);
      } else {
        x.x = src[0];
        if (k + 1 < K) x.y = src[1];
        if (k + 2 < K) x.z = src[2];
        if (k + 3 < K) x.w = src[3];
      }
    }
    reg[i] = x;
  }
}
// registers -> shared memory; `lo` (nullable) receives the TF32-rounded residual x - hi
// for the error-compensated 3xTF32 product
template <int ROWS>
__device__ __forceinline__ void store_chunk(float* __restrict__ hi, float* __restrict__ lo,
                                            const float4 (&reg)[ROWS * 8 / kThreads]) {
  constexpr int NV = ROWS * 8 / kThreads;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = threadIdx.x + i * kThreads;          // [k4][row] vector index == canonical layout
    const float4 x = reg[i];
    float4 h;
    h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
    reinterpret_cast<float4*>(hi)[v] = h;
    if (lo != nullptr) {
      float4 l;
      l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
      reinterpret_cast<float4*>(lo)[v] = l;
    }
  }
}

// C tile = act(A W^T + bias); BN in {32, 64, 128, 256}.  SPLIT: error-compensated 3xTF32
// (A_hi W_hi + A_lo W_hi + A_hi W_lo, all into the same TMEM accumulator): FP32-class
// accuracy at three MMAs per k-step, for the parity mode of the large-dim family.
template <int BN, bool SPLIT>
__global__ void __launch_bounds__(kThreads) gemm_tf32_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(128) unsigned char tc_smem_raw[];
  float* sA[2];
  float* sW[2];
  float* sAl[2] = {nullptr, nullptr};
  float* sWl[2] = {nullptr, nullptr};
  sA[0] = reinterpret_cast<float*>(tc_smem_raw);
  sA[1] = sA[0] + kBM * kBK;
  sW[0] = sA[1] + kBM * kBK;
  sW[1] = sW[0] + BN * kBK;
  if (SPLIT) {
    sAl[0] = sW[1] + BN * kBK;
    sAl[1] = sAl[0] + kBM * kBK;
    sWl[0] = sAl[1] + kBM * kBK;
    sWl[1] = sWl[0] + BN * kBK;
  }
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kBM;
  const int col0 = blockIdx.y * BN;
  constexpr uint32_t kCols = BN < 32 ? 32 : BN;

  if (warp == 0) tmem_alloc(&tmem_base_s, kCols);
  if (threadIdx.x == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  const bool a_vec = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
  const bool w_vec = (p.ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);
  const int64_t k_begin = p.k_split > 0 ? (int64_t)blockIdx.z * p.k_split : 0;
  const int64_t k_end = p.k_split > 0 ? (k_begin + p.k_split < p.K ? k_begin + p.k_split : p.K) : p.K;
  const int n_chunks = (int)((k_end - k_begin + kBK - 1) / kBK);
  const uint32_t idesc = umma_idesc_tf32(kBM, BN);
  float4 ra[kBM * 8 / kThreads], rw[BN * 8 / kThreads];
  auto load_ab = [&](int chunk) {
    load_chunk<kBM>(p.A, p.lda, row0, p.M, k_begin + (int64_t)chunk * kBK, k_end, a_vec, ra);
    load_chunk<BN>(p.W, p.ldw, col0, p.N, k_begin + (int64_t)chunk * kBK, k_end, w_vec, rw);
  };
  auto store_ab = [&](int s) {
    store_chunk<kBM>(sA[s], sAl[s], ra);
    store_chunk<BN>(sW[s], sWl[s], rw);
  };
  // per MMA k-step (8 floats of K) the operands advance by two 16-byte K vectors of ROWS rows
  constexpr uint32_t kStepA = 2 * kBM * 16, kStepW = 2 * BN * 16, kLboA = kBM * 16, kLboW = BN * 16;

  load_ab(0);
  store_ab(0);
  fence_async_smem();
  __syncthreads();

  uint32_t phase[2] = {0u, 0u};
  for (int i = 0; i < n_chunks; ++i) {
    const int s = i & 1;
    if (i + 1 < n_chunks) load_ab(i + 1);         // global loads of the next chunk fly during the MMAs
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA[s]), w0 = smem_u32(sW[s]);
#pragma unroll
      for (int j = 0; j < kBK / 8; ++j) {         // one MMA consumes K = 8 floats = two 16-byte K vectors
        const uint64_t da = umma_desc(a0 + j * kStepA, kLboA, 128);
        const uint64_t dw = umma_desc(w0 + j * kStepW, kLboW, 128);
        umma_tf32(tmem_d, da, dw, idesc, (i > 0 || j > 0) ? 1u : 0u);
        if (SPLIT) {
          const uint64_t dal = umma_desc(smem_u32(sAl[s]) + j * kStepA, kLboA, 128);
          const uint64_t dwl = umma_desc(smem_u32(sWl[s]) + j * kStepW, kLboW, 128);
          umma_tf32(tmem_d, dal, dw, idesc, 1u);
          umma_tf32(tmem_d, da, dwl, idesc, 1u);
        }
      }
      umma_commit(&mbar[s]);                      // arrives when these MMAs have read their operands
    }
    if (i + 1 < n_chunks) {
      if (i >= 1) { mbar_wait(&mbar[s ^ 1], phase[s ^ 1]); phase[s ^ 1] ^= 1u; }   // chunk i-1 done with its stage
      store_ab(s ^ 1);
      fence_async_smem();
      __syncthreads();
    }
  }
  // drain: the last one or two commits
  if (n_chunks >= 2) { const int s = (n_chunks - 2) & 1; mbar_wait(&mbar[s], phase[s]); phase[s] ^= 1u; }
  { const int s = (n_chunks - 1) & 1; mbar_wait(&mbar[s], phase[s]); phase[s] ^= 1u; }
  tc_fence_after();

  // epilogue: warp w owns accumulator rows (TMEM lanes) 32w .. 32w+31
  const int64_t row = row0 + warp * 32 + lane;
  const bool row_ok = row < p.M;
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    if (col0 + c >= p.N) break;
    float v[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    float* dst = p.C + row * p.ldc + col0 + c;
    const float* aux = p.mask_aux != nullptr ? p.mask_aux + row * p.ldaux + col0 + c : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = col0 + c + j;
      float x = 0.f;
      if (row_ok && col < p.N) {
        x = v[j] + (p.bias != nullptr ? p.bias[col] : 0.f);
        if (p.act == ACT_RELU) x = x < 0.f ? 0.f : x;
        if (aux != nullptr) x = aux[j] > 0.f ? x : 0.f;
        if (p.k_split > 0) {
          atomicAdd(dst + j, x);                                          // split-K partial tile
        } else {
          const float y = p.accumulate ? dst[j] + x : x;
          dst[j] = y;
          if (p.Ct != nullptr) p.Ct[(int64_t)col * p.ldct + row] = y;   // lanes = consecutive rows: coalesced
        }
      }
      if (p.colsum != nullptr) {                  // column sums of THIS product over the tile's rows (bias gradients)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && col < p.N) atomicAdd(p.colsum + col, x);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, kCols);
}

// =====================================================================================
// Building blocks of the product tile kernel (gemm_tf32_p_kernel below).  It was rebuilt
// around what the launch list and ncu captures of a C3-dims step showed for the round-1
// kernel above (profiles/r1_c3_launches_before.csv, r1_gemm_tf32_full.txt): 2.8 us per
// 32-float K chunk on the latency-bound single-particle GEMMs, 13x the HBM time on the
// 57 600-row particle GEMMs, and 4 warps per SM that were issue-bound on an emulated cvt.rna.
//
//  * operands travel global -> shared memory with 16-byte cp.async (LDGSTS, zero-filling
//    out-of-range rows / the K tail) into a ring of 2-4 stages;
//  * shared-memory operand layout: K-major SWIZZLE_128B (a row's 32-float chunk is one
//    128-byte line whose 16-byte vectors are XOR-permuted by row % 8; 8-row groups 1024 B
//    apart) — eight lanes fetch one full global line and store one full shared line;
//  * each thread rounds ITS vectors in place (hi = cvt.rn.tf32, one F2FP instruction; lo = x - hi
//    into the twin tile), so no barrier is needed between the copy and the conversion;
//  * epilogue per 32 x 32 accumulator block: bias / ReLU / ReLU-mask / accumulate / column sums /
//    split-K atomics / transposed copy, 128-bit row segments on the aligned path, a padded shared
//    patch (rows -> columns) wherever a full-line access pattern needs the transposition.
// =====================================================================================
constexpr int kRowBytes = kBK * 4;            // 128 B: one swizzle-128B line per operand row and chunk
constexpr int kMaxStages = 4;

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  // K-major SWIZZLE_128B: LBO unused (1), SBO = 1024 B between 8-row groups, version 1, layout type 2
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kThreadsV2 = 256;        // converter threads of the tile kernel
// Operand preparation by the thread that copied the vectors (no barrier in between):
// hi = rn_tf32(x) in place (round-to-nearest-even, ONE F2FP.TF32.F32 instruction on sm_100; cvt.rna is a
// four-instruction emulation, and this pass is issue-bound); for 3xTF32 the twin tile gets lo = x - hi,
// exact in fp32 and left unrounded — kind::tf32 reads its top 11 significant bits, 2^-22 |x| at worst.
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rn.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
template <int ROWS, bool SPLIT>
__device__ __forceinline__ void convert_tile(unsigned char* hi, unsigned char* lo, uint32_t off0) {
#pragma unroll
  for (int i = 0; i < ROWS / 32; ++i) {
    const uint32_t off = off0 + i * (32 * kRowBytes);
    const float4 x = *reinterpret_cast<const float4*>(hi + off);
    const float4 h = make_float4(rn_tf32(x.x), rn_tf32(x.y), rn_tf32(x.z), rn_tf32(x.w));
    *reinterpret_cast<float4*>(hi + off) = h;
    if (SPLIT) *reinterpret_cast<float4*>(lo + off) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
  }
}

template <int BN, bool SPLIT>
struct TileCfg {
  static constexpr int kABytes = kBM * kRowBytes;                    // 16 KB
  static constexpr int kWBytes = BN * kRowBytes;                     // multiple of 1024 (BN >= 32)
  static constexpr int kStageBytes = (SPLIT ? 2 : 1) * (kABytes + kWBytes);
  static_assert(BN % 32 == 0, "tile width");
  };

// 32 lanes x 16 consecutive fp32 columns, registers -> TMEM: thread (lane) writes its row's 16 values
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 lanes x 8 columns of tf32 per instruction) from tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
// the mbarrier receives one arrival when all cp.async of this thread issued so far have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* mbar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

// Epilogue of one 32-row x 32-column block of the accumulator held by one warp (thread = row, v[j] =
// column c + j): bias, ReLU, ReLU mask, column sums, accumulate / split-K atomics, C and its
// transposed copy.  `patch` is the warp's padded 32 x 33 shared staging area.
//  fast path (16-byte aligned rows, full block): the thread moves its 32 consecutive columns as
//   128-bit vectors (eight instructions fill each row's 128-byte line); the transposed copy leaves
//   as full lines (lane = row); only the column sums go through the patch;
//  general path: rows -> patch, then lane = column, so that every global access is a full line.
// (A variant staging C through the patch for full-line float4 stores measured SLOWER — 81 -> 101 us on the
// 57 600 x 64 -> 512 layer: the four epilogue warps are issue/latency-bound, not store-bound.)
constexpr int kPatchLd = 33;      // floats per patch row (conflict-free row <-> column transposition)
constexpr int kPatchLdBulk = 36;  // bulk-store mode: 144-byte rows (16-byte aligned, conflict-free 128-bit row writes)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void epilogue_block(const GemmParams& p, float (&v)[32], float* patch, int64_t wrow0,
                                               int rows_valid, int cbase, int lane, bool fast_c, bool bulk = false) {
  if (bulk) {                                  // the bulk stores of the previous block have read the patch
    bulk_wait_read();
    __syncwarp();
  }
  if (p.trans_out) {                           // lane = row: consecutive lanes add into consecutive addresses
    if (lane < rows_valid) {
      float* dt = p.C + (int64_t)cbase * p.ldc + wrow0 + lane;
      const int cols_valid = p.N - cbase < 32 ? p.N - cbase : 32;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cols_valid) atomicAdd(dt + (int64_t)j * p.ldc, v[j]);
    }
    return;
  }
  if (fast_c && cbase + 32 <= p.N) {
    const int64_t row = wrow0 + lane;
    const bool row_ok = lane < rows_valid;
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + cbase);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b = __ldg(b4 + q);
        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
      }
    }
    if (p.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] < 0.f ? 0.f : v[j];
    }
    if (p.mask_aux != nullptr && row_ok) {
      const float4* a4 = reinterpret_cast<const float4*>(p.mask_aux + row * p.ldaux + cbase);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 m = a4[q];
        v[4 * q] = m.x > 0.f ? v[4 * q] : 0.f; v[4 * q + 1] = m.y > 0.f ? v[4 * q + 1] : 0.f;
        v[4 * q + 2] = m.z > 0.f ? v[4 * q + 2] : 0.f; v[4 * q + 3] = m.w > 0.f ? v[4 * q + 3] : 0.f;
      }
    }
    if (!row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    if (p.colsum != nullptr) {                 // column sums of THIS product: transpose through the patch
#pragma unroll
      for (int j = 0; j < 32; ++j) patch[lane * 33 + j] = v[j];
      __syncwarp();
      float cs = 0.f;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) cs += patch[rr * 33 + lane];
      atomicAdd(p.colsum + cbase + lane, cs);
      __syncwarp();
    }
    if (bulk && !p.accumulate) {
      // thread = row: the row's 32 values go to the thread's own 144-byte patch row, then ONE 128-byte bulk copy
      // takes the row to global memory (a full line, off the LSU path) instead of eight 16-byte stores that
      // touch 32 lines per instruction
      float4* s4 = reinterpret_cast<float4*>(patch + lane * kPatchLdBulk);
#pragma unroll
      for (int q = 0; q < 8; ++q) s4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      fence_async_smem();
      if (row_ok)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(p.C + row * p.ldc + cbase),
                     "r"(smem_u32(s4))
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (row_ok && p.Ct != nullptr) {
        float* dt = p.Ct + (int64_t)cbase * p.ldct + row;
#pragma unroll
        for (int j = 0; j < 32; ++j) dt[(int64_t)j * p.ldct] = v[j];
      }
      return;
    }
    if (row_ok) {
      float4* d4 = reinterpret_cast<float4*>(p.C + row * p.ldc + cbase);
      if (p.accumulate) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 o = d4[q];
          v[4 * q] += o.x; v[4 * q + 1] += o.y; v[4 * q + 2] += o.z; v[4 * q + 3] += o.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (p.Ct != nullptr) {
        float* dt = p.Ct + (int64_t)cbase * p.ldct + row;
#pragma unroll
        for (int j = 0; j < 32; ++j) dt[(int64_t)j * p.ldct] = v[j];
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) patch[lane * 33 + j] = v[j];      // thread = row
  __syncwarp();
  const int col = cbase + lane;                                   // lane = column from here on
  const bool col_ok = col < p.N;
  const int cols_valid = p.N - cbase < 32 ? p.N - cbase : 32;
  const float bias = (p.bias != nullptr && col_ok) ? p.bias[col] : 0.f;
  float csum = 0.f;
  if (col_ok) {
    float* dst = p.C + wrow0 * p.ldc + col;
    const float* aux = p.mask_aux != nullptr ? p.mask_aux + wrow0 * p.ldaux + col : nullptr;
    const bool relu = p.act == ACT_RELU, split = p.k_split > 0, acc = p.accumulate != 0;
#pragma unroll 4
    for (int rr = 0; rr < rows_valid; ++rr) {
      float x = patch[rr * 33 + lane] + bias;
      if (relu) x = x < 0.f ? 0.f : x;
      if (aux != nullptr) { x = *aux > 0.f ? x : 0.f; aux += p.ldaux; }
      csum += x;
      if (split) {
        atomicAdd(dst, x);                                        // split-K partial tile
      } else {
        if (acc) x += *dst;
        *dst = x;
      }
      dst += p.ldc;
      patch[rr * 33 + lane] = x;                                  // final value, for the transposed copy
    }
    if (p.colsum != nullptr) atomicAdd(p.colsum + col, csum);
  }
  __syncwarp();
  if (p.Ct != nullptr && p.k_split == 0 && lane < rows_valid) {   // lane = row again: Ct[col][row], full lines
    float* dt = p.Ct + (int64_t)cbase * p.ldct + wrow0 + lane;
#pragma unroll 4
    for (int j = 0; j < cols_valid; ++j) { *dt = patch[lane * 33 + j]; dt += p.ldct; }
  }
  __syncwarp();
}
__device__ __forceinline__ bool epilogue_fast_ok(const GemmParams& p) {
  return (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && p.k_split == 0 &&
         (p.mask_aux == nullptr || ((p.ldaux % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.mask_aux) & 15) == 0))) &&
         (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
}

// =====================================================================================
// The tile kernel proper: persistent and GROUPED.  One launch runs up to kMaxGroup
// independent GEMM problems (the layers of a time step that do not depend on each other:
// three forward layers that read the same particles, six weight-gradient products, ...);
// one CTA per SM walks over the tiles of all problems (grid-stride; column tile fastest so
// that co-running CTAs share A rows in L2).  TMEM holds up to 8 accumulators (all 512 columns),
// so the epilogue of tile i (8 dedicated warps) overlaps the copies, rounding and MMAs of the
// following tiles, and the operand ring runs across tile and problem boundaries.
// Warp roles: warps 0..7 "converters" (cp.async + rounding), warp 8 MMA issuer, warps 9..16
// epilogue.  Barriers (arrivals per phase):
//   full[s]   8 (converter warps)   stage s holds rounded operands
//   empty[s]  tcgen05.commit        the MMAs that read stage s are done
//   tfull[a]  tcgen05.commit        accumulator a holds a finished tile
//   tempty[a] 8 (epilogue warps)    accumulator a has been read out
// =====================================================================================
constexpr int kEpiWarps = 8;
constexpr int kThreadsP = kThreadsV2 + 32 + kEpiWarps * 32;      // 544
#endif  // !BFVI_EMU
constexpr int kThreadsPhost = 544;
constexpr int kMaxGroup = 8;

struct GemmGroup {
  GemmParams g[kMaxGroup];
  int tile_end[kMaxGroup];                    // exclusive prefix sums of the tile counts
  int tiles_m[kMaxGroup], tiles_n[kMaxGroup];
  int chunks[kMaxGroup];                      // K chunks per tile (a short last split-K slice is zero padded)
  int n, total;
  int bulk_store;                             // 1: aligned epilogue rows leave through cp.async.bulk from a padded patch (BFVI_GEMM_BULK)
};
#ifndef BFVI_EMU

struct TileInfo {
  int pi, chunks, col0;
  int64_t row0, k_begin, k_end;
};
__device__ __forceinline__ TileInfo decode_tile(const GemmGroup& grp, int t, int BN) {
  TileInfo ti;
  int pi = 0, start = 0;
#pragma unroll 1
  while (pi + 1 < grp.n && t >= grp.tile_end[pi]) { start = grp.tile_end[pi]; ++pi; }
  const int local = t - start;
  const int tn = local % grp.tiles_n[pi];
  const int rest = local / grp.tiles_n[pi];
  const int tm = rest % grp.tiles_m[pi], tz = rest / grp.tiles_m[pi];
  const GemmParams& q = grp.g[pi];
  ti.pi = pi; ti.chunks = grp.chunks[pi];
  ti.row0 = (int64_t)tm * kBM; ti.col0 = tn * BN;
  ti.k_begin = q.k_split > 0 ? (int64_t)tz * q.k_split : 0;
  ti.k_end = q.k_split > 0 ? (ti.k_begin + q.k_split < q.K ? ti.k_begin + q.k_split : q.K) : q.K;
  return ti;
}

// K chunks of tile t (no coordinates: what the MMA issuer and the convert cursor need)
__device__ __forceinline__ int tile_chunks(const GemmGroup& grp, int t) {
  int pi = 0;
#pragma unroll 1
  while (pi + 1 < grp.n && t >= grp.tile_end[pi]) ++pi;
  return grp.chunks[pi];
}

template <int BN, bool SPLIT, bool VEC>
__global__ void __launch_bounds__(kThreadsP, 1) gemm_tf32_p_kernel(const __grid_constant__ GemmGroup grp, int n_stages, int lag_arg) {
  using Cfg = TileCfg<BN, SPLIT>;
  extern __shared__ unsigned char tc_smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  // accumulators in TMEM: as many as fit in its 512 columns (up to 8), so the MMA issuer can run that many
  // tiles ahead of the epilogue (with two, short-K tiles serialised: copies + MMAs + stores ADDED up)
  constexpr int kAcc = 512 / BN > 8 ? 8 : 512 / BN;
  __shared__ __align__(8) uint64_t tfull_bar[kAcc];
  __shared__ __align__(8) uint64_t tempty_bar[kAcc];
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = tc_smem_dyn + ((1024u - (smem_u32(tc_smem_dyn) & 1023u)) & 1023u);
  auto tileA = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes; };
  auto tileW = [&](int s) { return tileA(s) + Cfg::kABytes; };
  auto tileAl = [&](int s) { return tileW(s) + Cfg::kWBytes; };
  auto tileWl = [&](int s) { return tileAl(s) + Cfg::kABytes; };
  float* patches = reinterpret_cast<float*>(smem + (size_t)n_stages * Cfg::kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = kAcc * BN;                    // 256 (BN = 32) or 512, a power of two
  constexpr int kMmaWarp = kThreadsV2 / 32;

  if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, kCols);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], kThreadsV2 / 32); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < kAcc; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = ((int)blockIdx.x < grp.total) ? (grp.total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < kMmaWarp) {
    // ================= converters: copy + round, one "job" = one K chunk of one tile =================
    // Thread t owns the 16-byte vectors (row r0 + 32 i, column c4) of every chunk, r0 = t >> 3, c4 = t & 7.
    // Per-tile state of the load cursor is precomputed so that an interior job costs two pointer bumps
    // and six LDGSTS (the pass is issue-bound: profiles/r1_gemm_p_full.txt).
    const int r0 = threadIdx.x >> 3, c4 = threadIdx.x & 7;
    const uint32_t off0 = (uint32_t)(r0 * kRowBytes + ((c4 ^ (r0 & 7)) << 4));
    const uint32_t ring0 = smem_u32(smem) + off0;
    constexpr int kVA = kBM / 32, kVW = BN / 32;          // vectors per thread and chunk: A, W
    const float* a_ptr = nullptr; const float* w_ptr = nullptr;
    const float* a_base = nullptr; const float* w_base = nullptr;
    int64_t a_step = 0, w_step = 0;
    int na = 0, nw = 0;                                   // valid vectors (rows inside the matrix)
    int64_t k_left = 0;                                   // floats from this thread's column to the end of K
    int l_tile = -1, l_chunk = 0, l_chunks = 0;           // load cursor: next job to copy
    int c_tile = 0, c_chunk = 0, c_chunks = 0;            // convert cursor
    auto load_more = [&]() { return l_chunk < l_chunks || l_tile + 1 < my_tiles; };
    auto issue_next_load = [&](int s) {                   // copies the job under the load cursor into stage s
      if (l_chunk == l_chunks) {
        l_chunk = 0; ++l_tile;
        const TileInfo ti = decode_tile(grp, (int)blockIdx.x + l_tile * (int)gridDim.x, BN);
        const GemmParams& q = grp.g[ti.pi];
        l_chunks = ti.chunks;
        a_base = q.A; w_base = q.W;
        a_ptr = q.A + (ti.row0 + r0) * q.lda + ti.k_begin + c4 * 4;
        w_ptr = q.W + ((int64_t)ti.col0 + r0) * q.ldw + ti.k_begin + c4 * 4;
        a_step = 32 * q.lda; w_step = 32 * q.ldw;
        const int64_t ra = q.M - ti.row0 - r0, rw = (int64_t)q.N - ti.col0 - r0;      // rows from r0 to the edge
        na = ra <= 0 ? 0 : ra >= kBM ? kVA : (int)((ra + 31) / 32);
        nw = rw <= 0 ? 0 : rw >= BN ? kVW : (int)((rw + 31) / 32);
        if (na > kVA) na = kVA;
        if (nw > kVW) nw = kVW;
        k_left = ti.k_end - ti.k_begin - c4 * 4;
      }
      const uint32_t dst_a = ring0 + (uint32_t)s * Cfg::kStageBytes, dst_w = dst_a + Cfg::kABytes;
      if (VEC && k_left >= 4 && na == kVA && nw == kVW) {            // interior job
#pragma unroll
        for (int i = 0; i < kVA; ++i) cp_async16(dst_a + i * (32 * kRowBytes), a_ptr + i * a_step, 16);
#pragma unroll
        for (int i = 0; i < kVW; ++i) cp_async16(dst_w + i * (32 * kRowBytes), w_ptr + i * w_step, 16);
      } else {
        const int kbytes = k_left >= 4 ? 16 : k_left > 0 ? (int)k_left * 4 : 0;
#pragma unroll
        for (int i = 0; i < kVA + kVW; ++i) {
          const bool is_a = i < kVA;
          const int ii = is_a ? i : i - kVA;
          const bool in = ii < (is_a ? na : nw) && kbytes > 0;
          const float* src = is_a ? a_ptr + ii * a_step : w_ptr + ii * w_step;
          const float* dummy = is_a ? a_base : w_base;
          const uint32_t dst = (is_a ? dst_a : dst_w) + ii * (32 * kRowBytes);
          if (VEC) {
            cp_async16(dst, in ? src : dummy, in ? kbytes : 0);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool ine = in && e * 4 < kbytes;
              cp_async4(dst + e * 4, ine ? src + e : dummy, ine ? 4 : 0);
            }
          }
        }
      }
      a_ptr += kBK; w_ptr += kBK; k_left -= kBK;
      ++l_chunk;
    };
    // A stage is refilled `lag` jobs after the job that used it: with a 4-stage ring lag = 2, so the wait
    // for that job's MMAs (hand-over latency MMA warp -> tensor pipe -> commit -> this warp, ~0.5 us) has a
    // whole iteration of slack and two jobs stay in flight; shallower rings refill at once (lag 1).
    const int lag = lag_arg > 0 && lag_arg < n_stages ? lag_arg : (n_stages >= 4 ? 2 : 1);
    for (int c = 0; c < n_stages - lag; ++c) {            // prologue: jobs 0 .. n_stages-lag-1
      if (load_more()) issue_next_load(c);
      cp_async_commit();
    }
    if (my_tiles > 0) c_chunks = tile_chunks(grp, (int)blockIdx.x);
    int s = 0;
    uint32_t round = 0;                                   // how many times the ring has wrapped (parity source)
    int done = 0;                                         // jobs converted so far
    while (c_tile < my_tiles) {
      // job `done` is the oldest copy group but (n_stages - lag - 1) younger ones
      if (n_stages - lag == 1) cp_async_wait<0>(); else cp_async_wait<1>();
#ifndef BFVI_DBG_NO_CONVERT
      convert_tile<kBM, SPLIT>(tileA(s), tileAl(s), off0);
      convert_tile<BN, SPLIT>(tileW(s), tileWl(s), off0);
#endif
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      if (load_more()) {                                  // job done + n_stages - lag goes where job done - lag was
        const int sp = s - lag < 0 ? s - lag + n_stages : s - lag;
        if (done >= lag) mbar_wait(&empty_bar[sp], (s - lag < 0 ? round - 1u : round) & 1u);
        issue_next_load(sp);
      }
      cp_async_commit();
      ++done;
      if (++s == n_stages) { s = 0; ++round; }
      if (++c_chunk == c_chunks) {
        c_chunk = 0;
        if (++c_tile < my_tiles) c_chunks = tile_chunks(grp, (int)blockIdx.x + c_tile * (int)gridDim.x);
      }
    }
    cp_async_wait<0>();
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kBM, BN);
      int s = 0;
      uint32_t round = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int a = lt % kAcc;
        const int chunks = tile_chunks(grp, (int)blockIdx.x + lt * (int)gridDim.x);
        mbar_wait(&tempty_bar[a], (uint32_t)(((lt / kAcc) & 1) ^ 1));   // epilogue has drained accumulator a
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN);
        for (int i = 0; i < chunks; ++i) {
          mbar_wait(&full_bar[s], round & 1u);
          tc_fence_after();
          const uint32_t a0 = smem_u32(tileA(s)), w0 = smem_u32(tileW(s));
          const uint32_t al0 = smem_u32(tileAl(s)), wl0 = smem_u32(tileWl(s));
#ifndef BFVI_DBG_NO_MMA
#pragma unroll
          for (int q = 0; q < kBK / 8; ++q) {
            const uint64_t da = umma_desc_sw128(a0 + q * 32), dw = umma_desc_sw128(w0 + q * 32);
            umma_tf32(tmem_acc, da, dw, idesc, (i > 0 || q > 0) ? 1u : 0u);
            if (SPLIT) {
              umma_tf32(tmem_acc, umma_desc_sw128(al0 + q * 32), dw, idesc, 1u);
              umma_tf32(tmem_acc, da, umma_desc_sw128(wl0 + q * 32), idesc, 1u);
            }
          }
#endif
          umma_commit(&empty_bar[s]);
          if (++s == n_stages) { s = 0; ++round; }
        }
        umma_commit(&tfull_bar[a]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps: two per TMEM lane group 32 (warp % 4) .., alternating 32-column blocks
    const int wq = warp & 3, half = (warp - kMmaWarp - 1) >> 2;
    float* patch = patches + (warp - kMmaWarp - 1) * (32 * kPatchLd);
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int a = lt % kAcc;
      const TileInfo ti = decode_tile(grp, (int)blockIdx.x + lt * (int)gridDim.x, BN);
      const GemmParams& p = grp.g[ti.pi];
      const bool fast_c = VEC && epilogue_fast_ok(p);
      const int64_t wrow0 = ti.row0 + wq * 32;
      const int rows_valid = (int)(p.M - wrow0 < 32 ? (p.M - wrow0 < 0 ? 0 : p.M - wrow0) : 32);
      const int limit = p.N - ti.col0 < BN ? p.N - ti.col0 : BN;      // valid columns of this tile
      mbar_wait(&tfull_bar[a], (uint32_t)((lt / kAcc) & 1));
      tc_fence_after();
      bool released = false;
#pragma unroll 1
      for (int c = half * 32; c < limit; c += 64) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * BN + c), v);
        if (c + 64 >= limit) {                            // this warp's last block: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[a]);
          released = true;
        }
#ifndef BFVI_DBG_NO_EPI_STORE
        epilogue_block(p, v, patch, wrow0, rows_valid, ti.col0 + c, lane, fast_c);
#else
        if (v[0] == 123.456f) patch[lane] = v[1];
#endif
      }
      if (!released) {                                    // no block for this warp in a narrow tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[a]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kCols);
}
__host__ __device__ constexpr int ts_max_stages(int bn) { return bn <= 64 ? 6 : 4; }
// A-from-TMEM variant of the tile kernel: the tile kernel above is bound by the SM's shared-memory pipe
// (DESIGN.md 4b); here the rounded A chunk never returns to shared memory — converter warps 0-3 (thread =
// row) read the raw chunk once, round it and tcgen05.st hi / lo into a TMEM ring beside the accumulators, and
// the MMAs take A from tensor memory (tcgen05.mma [d], [a], b_desc): per job and BN = 128, 144 KB of
// shared-memory traffic instead of 224 KB.  Copies are tracked by an mbarrier (cp.async.mbarrier.arrive), so
// the copying and the converting thread of a vector need not be the same; warps 4-7 round W in place.
template <int BN, bool SPLIT, bool VEC>
__global__ void __launch_bounds__(kThreadsP, 1) gemm_tf32_ts_kernel(const __grid_constant__ GemmGroup grp, int n_stages, int lag_arg) {
  using Cfg = TileCfg<BN, SPLIT>;
  static_assert(BN <= 128, "A ring + accumulators must fit in 512 TMEM columns");
  constexpr int kStageTs = Cfg::kABytes + (SPLIT ? 2 : 1) * Cfg::kWBytes;      // raw A | W hi | W lo
  extern __shared__ unsigned char tc_smem_dyn[];
  // TMEM budget (512 columns): BN = 128: 2 accumulators + a 4-stage A ring (64 columns per stage: hi | lo);
  // BN <= 64: 128 columns of accumulators + a 6-stage ring — the narrow-output GEMMs stream their A operand
  // from HBM and want the bytes in flight (4 jobs x 32 KB per SM with lag 2)
  constexpr int kTsStages = ts_max_stages(BN);
  constexpr uint32_t kARing = BN <= 64 ? 128 : 256;        // first column of the A ring
  __shared__ __align__(8) uint64_t landed_bar[kTsStages];
  __shared__ __align__(8) uint64_t full_bar[kTsStages];
  __shared__ __align__(8) uint64_t empty_bar[kTsStages];
  // accumulators in TMEM: as many as fit in its 512 columns (up to 8), so the MMA issuer can run that many
  // tiles ahead of the epilogue (with two, short-K tiles serialised: copies + MMAs + stores ADDED up)
  constexpr int kAcc = (int)kARing / BN;
  __shared__ __align__(8) uint64_t tfull_bar[kAcc];
  __shared__ __align__(8) uint64_t tempty_bar[kAcc];
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = tc_smem_dyn + ((1024u - (smem_u32(tc_smem_dyn) & 1023u)) & 1023u);
  auto tileA = [&](int s) { return smem + (size_t)s * kStageTs; };
  auto tileW = [&](int s) { return tileA(s) + Cfg::kABytes; };
  auto tileWl = [&](int s) { return tileW(s) + Cfg::kWBytes; };
  float* patches = reinterpret_cast<float*>(smem + (size_t)n_stages * kStageTs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = 512;
  constexpr int kMmaWarp = kThreadsV2 / 32;

  if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, kCols);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTsStages; ++s) {
      mbar_init(&landed_bar[s], kThreadsV2); mbar_init(&full_bar[s], kThreadsV2 / 32); mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kAcc; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = ((int)blockIdx.x < grp.total) ? (grp.total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < kMmaWarp) {
    // ================= converters: copy + round, one "job" = one K chunk of one tile =================
    // Thread t owns the 16-byte vectors (row r0 + 32 i, column c4) of every chunk, r0 = t >> 3, c4 = t & 7.
    // Per-tile state of the load cursor is precomputed so that an interior job costs two pointer bumps
    // and six LDGSTS (the pass is issue-bound: profiles/r1_gemm_p_full.txt).
    const int r0 = threadIdx.x >> 3, c4 = threadIdx.x & 7;
    const uint32_t off0 = (uint32_t)(r0 * kRowBytes + ((c4 ^ (r0 & 7)) << 4));
    const uint32_t ring0 = smem_u32(smem) + off0;
    constexpr int kVA = kBM / 32, kVW = BN / 32;          // vectors per thread and chunk: A, W
    const float* a_ptr = nullptr; const float* w_ptr = nullptr;
    const float* a_base = nullptr; const float* w_base = nullptr;
    int64_t a_step = 0, w_step = 0;
    int na = 0, nw = 0;                                   // valid vectors (rows inside the matrix)
    int64_t k_left = 0;                                   // floats from this thread's column to the end of K
    int l_tile = -1, l_chunk = 0, l_chunks = 0;           // load cursor: next job to copy
    int c_tile = 0, c_chunk = 0, c_chunks = 0;            // convert cursor
    auto load_more = [&]() { return l_chunk < l_chunks || l_tile + 1 < my_tiles; };
    auto issue_next_load = [&](int s) {                   // copies the job under the load cursor into stage s
      if (l_chunk == l_chunks) {
        l_chunk = 0; ++l_tile;
        const TileInfo ti = decode_tile(grp, (int)blockIdx.x + l_tile * (int)gridDim.x, BN);
        const GemmParams& q = grp.g[ti.pi];
        l_chunks = ti.chunks;
        a_base = q.A; w_base = q.W;
        a_ptr = q.A + (ti.row0 + r0) * q.lda + ti.k_begin + c4 * 4;
        w_ptr = q.W + ((int64_t)ti.col0 + r0) * q.ldw + ti.k_begin + c4 * 4;
        a_step = 32 * q.lda; w_step = 32 * q.ldw;
        const int64_t ra = q.M - ti.row0 - r0, rw = (int64_t)q.N - ti.col0 - r0;      // rows from r0 to the edge
        na = ra <= 0 ? 0 : ra >= kBM ? kVA : (int)((ra + 31) / 32);
        nw = rw <= 0 ? 0 : rw >= BN ? kVW : (int)((rw + 31) / 32);
        if (na > kVA) na = kVA;
        if (nw > kVW) nw = kVW;
        k_left = ti.k_end - ti.k_begin - c4 * 4;
      }
      const uint32_t dst_a = ring0 + (uint32_t)s * kStageTs, dst_w = dst_a + Cfg::kABytes;
      if (VEC && k_left >= 4 && na == kVA && nw == kVW) {            // interior job
#pragma unroll
        for (int i = 0; i < kVA; ++i) cp_async16(dst_a + i * (32 * kRowBytes), a_ptr + i * a_step, 16);
#pragma unroll
        for (int i = 0; i < kVW; ++i) cp_async16(dst_w + i * (32 * kRowBytes), w_ptr + i * w_step, 16);
      } else {
        const int kbytes = k_left >= 4 ? 16 : k_left > 0 ? (int)k_left * 4 : 0;
#pragma unroll
        for (int i = 0; i < kVA + kVW; ++i) {
          const bool is_a = i < kVA;
          const int ii = is_a ? i : i - kVA;
          const bool in = ii < (is_a ? na : nw) && kbytes > 0;
          const float* src = is_a ? a_ptr + ii * a_step : w_ptr + ii * w_step;
          const float* dummy = is_a ? a_base : w_base;
          const uint32_t dst = (is_a ? dst_a : dst_w) + ii * (32 * kRowBytes);
          if (VEC) {
            cp_async16(dst, in ? src : dummy, in ? kbytes : 0);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool ine = in && e * 4 < kbytes;
              cp_async4(dst + e * 4, ine ? src + e : dummy, ine ? 4 : 0);
            }
          }
        }
      }
      a_ptr += kBK; w_ptr += kBK; k_left -= kBK;
      ++l_chunk;
      cp_async_arrive(&landed_bar[s]);
    };
    // A stage is refilled `lag` jobs after the job that used it: with a 4-stage ring lag = 2, so the wait
    // for that job's MMAs (hand-over latency MMA warp -> tensor pipe -> commit -> this warp, ~0.5 us) has a
    // whole iteration of slack and two jobs stay in flight; shallower rings refill at once (lag 1).
    const int lag = lag_arg > 0 && lag_arg < n_stages ? lag_arg : (n_stages >= 4 ? 2 : 1);
    for (int c = 0; c < n_stages - lag; ++c)              // prologue: jobs 0 .. n_stages-lag-1
      if (load_more()) issue_next_load(c);
    if (my_tiles > 0) c_chunks = tile_chunks(grp, (int)blockIdx.x);
    int s = 0;
    uint32_t round = 0;                                   // how many times the ring has wrapped (parity source)
    int done = 0;                                         // jobs converted so far
    const int wt = threadIdx.x - 128;                     // W converters: vectors (row (wt >> 3) + 16 i, column wt & 7)
    const uint32_t woff0 = (uint32_t)((wt >> 3) * kRowBytes + (((wt & 7) ^ ((wt >> 3) & 7)) << 4));
    while (c_tile < my_tiles) {
      mbar_wait(&landed_bar[s], round & 1u);              // every thread's copies of this job have landed
      if (warp < 4) {
        // thread = row: read the raw chunk (8 swizzled 16-byte vectors), round, store hi | lo to the TMEM ring
        const unsigned char* rowp = tileA(s) + threadIdx.x * kRowBytes;
        const uint32_t tdst = tmem_base + ((uint32_t)(warp * 32) << 16) + kARing + (uint32_t)s * 64u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int cc = h * 4 + c;
            const float4 x = *reinterpret_cast<const float4*>(rowp + ((cc ^ (threadIdx.x & 7)) << 4));
            hi[4 * c] = rn_tf32(x.x); hi[4 * c + 1] = rn_tf32(x.y); hi[4 * c + 2] = rn_tf32(x.z); hi[4 * c + 3] = rn_tf32(x.w);
            lo[4 * c] = x.x - hi[4 * c]; lo[4 * c + 1] = x.y - hi[4 * c + 1];
            lo[4 * c + 2] = x.z - hi[4 * c + 2]; lo[4 * c + 3] = x.w - hi[4 * c + 3];
          }
          tmem_st16(tdst + h * 16, hi);
          if (SPLIT) tmem_st16(tdst + 32 + h * 16, lo);
        }
        tmem_wait_st();
        tc_fence_before();
      } else {
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
          const uint32_t off = woff0 + i * (16 * kRowBytes);
          const float4 x = *reinterpret_cast<const float4*>(tileW(s) + off);
          const float4 h = make_float4(rn_tf32(x.x), rn_tf32(x.y), rn_tf32(x.z), rn_tf32(x.w));
          *reinterpret_cast<float4*>(tileW(s) + off) = h;
          if (SPLIT) *reinterpret_cast<float4*>(tileWl(s) + off) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
        fence_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      if (load_more()) {                                  // job done + n_stages - lag goes where job done - lag was
        const int sp = s - lag < 0 ? s - lag + n_stages : s - lag;
        if (done >= lag) mbar_wait(&empty_bar[sp], (s - lag < 0 ? round - 1u : round) & 1u);
        issue_next_load(sp);
      }
      ++done;
      if (++s == n_stages) { s = 0; ++round; }
      if (++c_chunk == c_chunks) {
        c_chunk = 0;
        if (++c_tile < my_tiles) c_chunks = tile_chunks(grp, (int)blockIdx.x + c_tile * (int)gridDim.x);
      }
    }
    cp_async_wait<0>();
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kBM, BN);
      int s = 0;
      uint32_t round = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int a = lt % kAcc;
        const int chunks = tile_chunks(grp, (int)blockIdx.x + lt * (int)gridDim.x);
        mbar_wait(&tempty_bar[a], (uint32_t)(((lt / kAcc) & 1) ^ 1));   // epilogue has drained accumulator a
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN);
        for (int i = 0; i < chunks; ++i) {
          mbar_wait(&full_bar[s], round & 1u);
          tc_fence_after();
          const uint32_t w0 = smem_u32(tileW(s)), wl0 = smem_u32(tileWl(s));
          const uint32_t ah = tmem_base + kARing + (uint32_t)s * 64u, al = ah + 32u;
#pragma unroll
          for (int q = 0; q < kBK / 8; ++q) {
            const uint64_t dw = umma_desc_sw128(w0 + q * 32);
            umma_tf32_ts(tmem_acc, ah + q * 8, dw, idesc, (i > 0 || q > 0) ? 1u : 0u);
            if (SPLIT) {
              umma_tf32_ts(tmem_acc, al + q * 8, dw, idesc, 1u);
              umma_tf32_ts(tmem_acc, ah + q * 8, umma_desc_sw128(wl0 + q * 32), idesc, 1u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (++s == n_stages) { s = 0; ++round; }
        }
        umma_commit(&tfull_bar[a]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps: two per TMEM lane group 32 (warp % 4) .., alternating 32-column blocks
    const int wq = warp & 3, half = (warp - kMmaWarp - 1) >> 2;
    const bool bulk = grp.bulk_store != 0;
    float* patch = patches + (warp - kMmaWarp - 1) * (32 * (bulk ? kPatchLdBulk : kPatchLd));
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int a = lt % kAcc;
      const TileInfo ti = decode_tile(grp, (int)blockIdx.x + lt * (int)gridDim.x, BN);
      const GemmParams& p = grp.g[ti.pi];
      const bool fast_c = VEC && epilogue_fast_ok(p);
      const int64_t wrow0 = ti.row0 + wq * 32;
      const int rows_valid = (int)(p.M - wrow0 < 32 ? (p.M - wrow0 < 0 ? 0 : p.M - wrow0) : 32);
      const int limit = p.N - ti.col0 < BN ? p.N - ti.col0 : BN;      // valid columns of this tile
      mbar_wait(&tfull_bar[a], (uint32_t)((lt / kAcc) & 1));
      tc_fence_after();
      bool released = false;
#pragma unroll 1
      for (int c = half * 32; c < limit; c += 64) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * BN + c), v);
        if (c + 64 >= limit) {                            // this warp's last block: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[a]);
          released = true;
        }
#ifndef BFVI_DBG_NO_EPI_STORE
        epilogue_block(p, v, patch, wrow0, rows_valid, ti.col0 + c, lane, fast_c, bulk);
#else
        if (v[0] == 123.456f) patch[lane] = v[1];
#endif
      }
      if (!released) {                                    // no block for this warp in a narrow tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[a]);
      }
    }
    if (bulk) bulk_wait_read();                           // shared memory stays valid until the last rows are read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kCols);
}
#endif  // !BFVI_EMU

// CPU stand-in used only by the SIMT-emulator test build (exact fp32, no TF32 rounding)
#ifdef BFVI_EMU
inline void gemm_reference_emu(const GemmParams& p) {
  for (int n = 0; n < p.N; ++n) {
    float cs = 0.f;
    for (int64_t m = 0; m < p.M; ++m) {
      float acc = 0.f;
      for (int64_t k = 0; k < p.K; ++k) acc = fmaf(p.A[m * p.lda + k], p.W[(int64_t)n * p.ldw + k], acc);
      acc += p.bias ? p.bias[n] : 0.f;
      if (p.act == ACT_RELU) acc = acc < 0.f ? 0.f : acc;
      if (p.mask_aux != nullptr) acc = p.mask_aux[m * p.ldaux + n] > 0.f ? acc : 0.f;
      if (p.trans_out) { p.C[(int64_t)n * p.ldc + m] += acc; continue; }
      cs += acc;
      acc = p.accumulate ? p.C[m * p.ldc + n] + acc : acc;
      p.C[m * p.ldc + n] = acc;
      if (p.Ct != nullptr) p.Ct[(int64_t)n * p.ldct + m] = acc;
    }
    if (p.colsum != nullptr) p.colsum[n] += cs;
  }
}
#endif

template <int BN, bool SPLIT>
inline size_t gemm_smem_bytes() { return sizeof(float) * (SPLIT ? 4 : 2) * (size_t)(kBM + BN) * kBK; }
// v2: ring of `stages` stages (+ slack for the 1024-byte alignment of the swizzled tiles)
template <int BN, bool SPLIT>
inline size_t gemm_v2_stage_bytes() { return (size_t)(SPLIT ? 2 : 1) * (size_t)(kBM + BN) * kBK * sizeof(float); }
template <int BN, bool SPLIT>
inline size_t gemm_v2_smem_bytes(int stages) { return gemm_v2_stage_bytes<BN, SPLIT>() * (size_t)stages + 1024; }
// persistent kernel: ring + four epilogue patches
// A-from-TMEM variant: raw A | W hi | W lo per stage
template <int BN, bool SPLIT>
inline size_t gemm_ts_stage_bytes() { return (size_t)(kBM + (SPLIT ? 2 : 1) * BN) * kBK * sizeof(float); }
template <int BN, bool SPLIT>
inline size_t gemm_ts_smem_bytes(int stages, bool bulk = false) {
  return gemm_ts_stage_bytes<BN, SPLIT>() * (size_t)stages + 1024 + 8 * 32 * (bulk ? 36 : 33) * sizeof(float);
}
template <int BN, bool SPLIT>
inline size_t gemm_p_smem_bytes(int stages) { return gemm_v2_smem_bytes<BN, SPLIT>(stages) + 8 * 32 * 33 * sizeof(float); }

}  // namespace tc
}  // namespace bfvi
