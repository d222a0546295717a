// bfvi_fused.cuh — fused on-chip GaussianGTF kernels of the large-dim family (precision mode BFVI_PREC_TF32).
//
// One transition evaluation (models/common.py:62-68) on a tile of 128 latent rows (rows = chains x particles of one
// time step, models/dmm.py:239-258) is a chain of dense contractions
//
//     z[128,Z] -> relu(z W0^T + b0)[128,H] -> (.) W2^T [128,Z]        (gate and nonlinear branch)
//
// whose hidden activations are 8x larger than its inputs and outputs.  The launch-sequence path (bfvi_tc.cuh) writes
// them to HBM and reads them back between launches: ~10 MB of DRAM traffic per sequence-timestep at the C3 shape
// against 1 KB of algorithmic bytes (profiles/r1_particle_pass_ncu.txt).  Here they never leave the SM:
//
//   gtf_fwd_kernel   per row tile, per unit of 64 hidden columns:  tcgen05.mma D = z W0u^T into TMEM  ->  row warps
//                    tcgen05.ld, + bias, ReLU, round to TF32, tcgen05.st back IN PLACE as the A operand  ->
//                    tcgen05.mma head += A W2u^T.  z itself is the A operand from TMEM (one tcgen05.st per tile).
//                    Heads (pre-sigmoid gate, nonlinear, linear, pre-softplus std) leave as (R, Z) fp32 rows.
//                    KEEP mode (the backward recompute): also writes the hidden activations as FP16 operand tiles
//                    for the weight-gradient GEMM, the ReLU sign bits (64 per row and unit) and an FP16 copy of z.
//   gtf_bwd_kernel   same pipeline for the input gradient:  D = d_head W2u  ->  mask by the ReLU bits, column sums
//                    (bias gradients), FP16 tile for the weight gradients, round, A in place  ->  dz += A W0u.
//   wgrad16_kernel   dW^T[H, Z] += X^T Y over all rows: X (hidden activations / their gradients) and Y (z / head
//                    gradients) are the FP16 tiles the two kernels above wrote in the MN-major SWIZZLE_128B shared-memory
//                    image, so a stage is three cp.async.bulk copies; kind::f16 MMAs, FP32 accumulation in TMEM over a
//                    slice of the rows, one red.global.add pass per work item.
//
// Weights: tf32-rounded ONCE per step into the exact shared-memory image of every unit (pack_gtf_kernel: K-major
// SWIZZLE_128B tiles, 32 KB per unit) and streamed through a ring of stages by ONE thread with cp.async.bulk
// (global -> shared, mbarrier complete_tx) — no per-tile rounding pass, no LDGSTS, no converter warps.
//
// Warp roles (320 threads): warps 0-7 "row warps" (thread = tile row = TMEM lane; two warps per 32-lane quadrant, one
// per 32-column half), warp 8 MMA issuer (one elected thread), warp 9 weight loader (one elected thread).
#pragma once
#include "bfvi_platform.cuh"
#include "bfvi_tc.cuh"

#ifndef BFVI_EMU
#include <cuda_fp16.h>
#endif

namespace bfvi {
namespace fused {

constexpr int kZ = 64;                 // latent width served by the fused kernels
constexpr int kHU = 64;                // hidden columns per unit
constexpr int kTileRows = 128;         // rows per tile (UMMA M)
constexpr int kBlockBytes = 32768;     // one weight block: two 64 x 64 tf32 operand tiles
constexpr int kTileBytes = 16384;
constexpr int kRowWarps = 8;
constexpr int kThreads = (kRowWarps + 2) * 32;
constexpr int kMaxStages = 6;
constexpr int kRowGroup = 64;          // rows per FP16 operand tile of the weight-gradient GEMM (one K stage)
constexpr int kAtomBytes = 8192;       // 64 rows x 64 halves, MN-major SWIZZLE_128B

// shared-memory image of a 64 x 64 fp32 operand tile, K-major SWIZZLE_128B: two K halves of 32 floats; row r of a
// half is one 128-byte line whose 16-byte chunks are XOR-permuted by r % 8 (8-row groups 1024 B apart)
__host__ __device__ inline int tile_offset(int r, int k) {
  return (k >> 5) * 8192 + r * 128 + (((((k & 31) >> 2) ^ (r & 7))) << 4) + (k & 3) * 4;
}

// Weight packs of one GaussianGTF.  FWD: blocks 0..2U-1 = units (gate branch first), block 2U = tail.
//   unit block: tile 0 = W0[unit rows, :] (N = hidden, K = z), tile 1 = W2[:, unit cols] (N = z, K = hidden)
//   tail block: tile 0 = W_lin, tile 1 = W_std
// BWD: block 0 = head (tile 1 = W_lin^T), blocks 1..2U = units:
//   tile 0 = W2[:, unit cols]^T (N = hidden, K = z_out), tile 1 = W0[unit rows, :]^T (N = z_in, K = hidden)
// followed by the bias table: b0 gate (H), b0 nonlin (H), gate2_b, nonlin2_b, lin_b, std_b (Z each).
struct PackParams {
  const float* w_gate0; const float* b_gate0; const float* w_gate2; const float* b_gate2;
  const float* w_lin; const float* b_lin; const float* w_non0; const float* b_non0;
  const float* w_non2; const float* b_non2; const float* w_std; const float* b_std;
  unsigned char* fwd; unsigned char* bwd; float* bias;
  int H;
};
inline size_t pack_blocks(int H) { return (size_t)(2 * (H / kHU) + 1); }
inline size_t pack_bytes(int H) { return pack_blocks(H) * kBlockBytes; }
inline size_t bias_floats(int H) { return (size_t)2 * H + 4 * kZ; }

#ifndef BFVI_EMU
using tc::smem_u32;
using tc::mbar_init; using tc::mbar_wait; using tc::mbar_arrive; using tc::umma_commit;
using tc::tmem_alloc; using tc::tmem_dealloc; using tc::tmem_ld32; using tc::tc_fence_before; using tc::tc_fence_after;
using tc::umma_desc_sw128; using tc::umma_idesc_tf32; using tc::umma_tf32_ts; using tc::rn_tf32; using tc::tmem_wait_st;

__global__ void __launch_bounds__(256) pack_gtf_kernel(const __grid_constant__ PackParams p) {
  const int H = p.H, U = H / kHU, nb = 2 * U + 1;
  const int b = blockIdx.x;                                  // block index, both packs
  float* fw = reinterpret_cast<float*>(p.fwd + (size_t)b * kBlockBytes);
  float* bw = reinterpret_cast<float*>(p.bwd + (size_t)b * kBlockBytes);
  for (int e = threadIdx.x; e < 2 * 64 * 64; e += blockDim.x) {
    const int t = e >> 12, r = (e >> 6) & 63, k = e & 63;
    const int off = (t * kTileBytes + tile_offset(r, k)) >> 2;
    // ---- forward pack
    float v;
    if (b < 2 * U) {
      const int br = b / U, c = b % U;
      const float* w0 = br ? p.w_non0 : p.w_gate0;
      const float* w2 = br ? p.w_non2 : p.w_gate2;
      v = t == 0 ? w0[(size_t)(c * kHU + r) * kZ + k] : w2[(size_t)r * H + c * kHU + k];
    } else {
      v = t == 0 ? p.w_lin[r * kZ + k] : p.w_std[r * kZ + k];
    }
    fw[off] = rn_tf32(v);
    // ---- backward pack
    if (b == 0) {
      v = t == 0 ? 0.f : p.w_lin[k * kZ + r];
    } else {
      const int u = b - 1, br = u / U, c = u % U;
      const float* w0 = br ? p.w_non0 : p.w_gate0;
      const float* w2 = br ? p.w_non2 : p.w_gate2;
      v = t == 0 ? w2[(size_t)k * H + c * kHU + r] : w0[(size_t)(c * kHU + k) * kZ + r];
    }
    bw[off] = rn_tf32(v);
  }
  if (b == nb - 1) {
    for (int i = threadIdx.x; i < H; i += blockDim.x) { p.bias[i] = p.b_gate0[i]; p.bias[H + i] = p.b_non0[i]; }
    for (int i = threadIdx.x; i < kZ; i += blockDim.x) {
      p.bias[2 * H + i] = p.b_gate2[i]; p.bias[2 * H + kZ + i] = p.b_non2[i];
      p.bias[2 * H + 2 * kZ + i] = p.b_lin[i]; p.bias[2 * H + 3 * kZ + i] = p.b_std[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// small PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP.S.G)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns, registers -> TMEM
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void load_row32(const float* __restrict__ src, bool ok, float (&v)[32]) {
  if (ok) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 x = __ldg(s4 + q);
      v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
}
__device__ __forceinline__ void store_row32(float* __restrict__ dst, const float (&v)[32]) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int q = 0; q < 8; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
// FP16 operand tiles of the weight-gradient GEMM: (row group of 64 rows) x (atom of 64 columns) = 8 KB, row r of the
// group is a 128-byte line, 16-byte chunk c stored at c ^ (r % 8): the MN-major SWIZZLE_128B shared-memory image, so
// the GEMM loads a tile with one bulk copy.  A thread (row, 32-column half) writes its four chunks.
__device__ __forceinline__ void store_half32(__half* __restrict__ base, int64_t row, int n_atoms, int atom, int hf,
                                             const float (&v)[32]) {
  unsigned char* tile = reinterpret_cast<unsigned char*>(base) +
                        ((size_t)(row / kRowGroup) * n_atoms + atom) * kAtomBytes + (size_t)(row % kRowGroup) * 128;
  const int sw = (int)(row & 7);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 pk;
    __half2 h0 = __floats2half2_rn(v[8 * c], v[8 * c + 1]), h1 = __floats2half2_rn(v[8 * c + 2], v[8 * c + 3]);
    __half2 h2 = __floats2half2_rn(v[8 * c + 4], v[8 * c + 5]), h3 = __floats2half2_rn(v[8 * c + 6], v[8 * c + 7]);
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(tile + (((hf * 4 + c) ^ sw) << 4)) = pk;
  }
}

struct FwdParams {
  const unsigned char* pack;     // forward pack
  const float* bias;             // bias table (see PackParams)
  const float* z;                // (R, 64)
  float* g; float* nl; float* lin; float* as;      // (R, 64) heads, biases added
  __half* h16;                   // KEEP: hidden activations, FP16 tiles [row group][2U atoms]
  uint32_t* relu_bits;           // KEEP: sign bits of the hidden activations, [row tile][unit][half][128 rows] words
  __half* z16;                   // KEEP: FP16 tiles of z [row group][1 atom]
  int64_t R;
  int H;
  int n_stages;
};

// TMEM columns of the forward kernel
constexpr uint32_t kFZ = 0, kFG = 64, kFNL = 128, kFLIN = 192, kFAS = 256, kFHB = 320;   // 3 hidden buffers of 64

template <bool KEEP>
__global__ void __launch_bounds__(kThreads, 1) gtf_fwd_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ unsigned char fused_smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t z_full, units_done, tail_a, heads_full, heads_empty;
  __shared__ __align__(8) uint64_t d_full[3];
  __shared__ __align__(8) uint64_t a_full[3];
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = fused_smem_dyn + ((1024u - (smem_u32(fused_smem_dyn) & 1023u)) & 1023u);
  const int n_stages = p.n_stages;
  float* bias_s = reinterpret_cast<float*>(smem + (size_t)n_stages * kBlockBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, U = H / kHU, U2 = 2 * U, n_blocks = U2 + 1;
  const int64_t n_tiles = (p.R + kTileRows - 1) / kTileRows;
  const int my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;

  if (warp == kRowWarps) tmem_alloc(&tmem_base_s, 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&z_full, kRowWarps); mbar_init(&units_done, 1); mbar_init(&tail_a, kRowWarps);
    mbar_init(&heads_full, 1); mbar_init(&heads_empty, kRowWarps);
    for (int i = 0; i < 3; ++i) { mbar_init(&d_full[i], 1); mbar_init(&a_full[i], kRowWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * H + 4 * kZ; i += blockDim.x) bias_s[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;

  if (warp < kRowWarps) {
    // ================= row warps =================
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t tl = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32);   // my lane group, my column half
    uint32_t par_d = 0, par_misc = 0;                 // phase bits: d_full[i] in bit i; misc flips once per tile
    float zreg[32];
    {
      const int64_t row = (int64_t)blockIdx.x * kTileRows + q * 32 + lane;
      load_row32(p.z + row * kZ + hf * 32, my_tiles > 0 && row < p.R, zreg);
    }
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)lt * gridDim.x;
      const int64_t row = tile * kTileRows + q * 32 + lane;
      const bool row_ok = row < p.R;
      // ---- z -> TMEM (A operand of the hidden layers and of the linear head), rounded once
      // (all MMAs of the previous tile are complete: this warp has passed heads_full of that tile)
      if (KEEP) store_half32(p.z16, row, 1, 0, hf, zreg);
#pragma unroll
      for (int j = 0; j < 32; ++j) zreg[j] = rn_tf32(zreg[j]);
      tmem_st32(tl + kFZ, zreg);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&z_full);
      if (lt + 1 < my_tiles) {                        // next tile's rows fly during this tile's units
        const int64_t nrow = (tile + gridDim.x) * kTileRows + q * 32 + lane;
        load_row32(p.z + nrow * kZ + hf * 32, nrow < p.R, zreg);
      }
      // ---- hidden units: D -> + bias, ReLU, round -> A, in place
#pragma unroll 1
      for (int u = 0; u < U2; ++u) {
        const int hb = u % 3;
        mbar_wait(&d_full[hb], (par_d >> hb) & 1u);
        par_d ^= 1u << hb;
        tc_fence_after();
        float v[32];
        tmem_ld32(tl + kFHB + hb * 64, v);
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + u * kHU + hf * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b = b4[c];
          v[4 * c] = fmaxf(v[4 * c] + b.x, 0.f); v[4 * c + 1] = fmaxf(v[4 * c + 1] + b.y, 0.f);
          v[4 * c + 2] = fmaxf(v[4 * c + 2] + b.z, 0.f); v[4 * c + 3] = fmaxf(v[4 * c + 3] + b.w, 0.f);
        }
        if (KEEP) {
          uint32_t bits = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) bits |= (v[j] > 0.f ? 1u : 0u) << j;
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;      // rows past the end contribute nothing to the weight gradients
          }
          p.relu_bits[((tile * U2 + u) * 2 + hf) * kTileRows + q * 32 + lane] = row_ok ? bits : 0u;   // coalesced
          store_half32(p.h16, row, U2, u, hf, v);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        tmem_st32(tl + kFHB + hb * 64, v);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[hb]);
      }
      // ---- tail: the finished nonlinear head is the A operand of the std head
      {
        mbar_wait(&units_done, par_misc & 1u);
        tc_fence_after();
        float v[32];
        tmem_ld32(tl + kFNL, v);
        const float* b = bias_s + 2 * H + kZ + hf * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        if (row_ok) store_row32(p.nl + row * kZ + hf * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        tmem_st32(tl + kFHB, v);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail_a);
      }
      // ---- heads out
      {
        mbar_wait(&heads_full, par_misc & 1u);
        tc_fence_after();
        float v[32];
        tmem_ld32(tl + kFG, v);
        const float* b = bias_s + 2 * H + hf * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        if (row_ok) store_row32(p.g + row * kZ + hf * 32, v);
        tmem_ld32(tl + kFLIN, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[2 * kZ + j];
        if (row_ok) store_row32(p.lin + row * kZ + hf * 32, v);
        tmem_ld32(tl + kFAS, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[3 * kZ + j];
        if (row_ok) store_row32(p.as + row * kZ + hf * 32, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&heads_empty);
      }
      par_misc ^= 1u;
    }
  } else if (warp == kRowWarps) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kTileRows, 64);
      const uint32_t ring = smem_u32(smem);
      uint32_t par_a = 0;
      int64_t gblk = 0;                               // blocks consumed so far (ring position)
      auto stage_of = [&](int64_t g) { return (int)(g % n_stages); };
      auto wait_block = [&](int64_t g) { mbar_wait(&full_bar[stage_of(g)], (uint32_t)((g / n_stages) & 1)); tc_fence_after(); };
      auto mma8 = [&](uint32_t d_col, uint32_t a_col, uint32_t tile_addr, bool fresh) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_tf32_ts(tb + d_col, tb + a_col + k * 8, umma_desc_sw128(tile_addr + (k >> 2) * 8192 + (k & 3) * 32), idesc,
                       (fresh && k == 0) ? 0u : 1u);
      };
      for (int lt = 0; lt < my_tiles; ++lt) {
        mbar_wait(&z_full, (uint32_t)(lt & 1));
        tc_fence_after();
        auto issue1 = [&](int u) {                    // hidden pre-activations of unit u
          wait_block(gblk + u);
          mma8(kFHB + (u % 3) * 64, kFZ, ring + stage_of(gblk + u) * kBlockBytes, true);
          umma_commit(&d_full[u % 3]);
        };
        issue1(0);
        if (U2 > 1) issue1(1);
        for (int u = 0; u < U2; ++u) {
          const int hb = u % 3;
          mbar_wait(&a_full[hb], (par_a >> hb) & 1u);
          par_a ^= 1u << hb;
          tc_fence_after();
          if (u == 0) { mbar_wait(&heads_empty, (uint32_t)((lt & 1) ^ 1)); tc_fence_after(); }   // previous heads read out
          mma8(u < U ? kFG : kFNL, kFHB + hb * 64, ring + stage_of(gblk + u) * kBlockBytes + kTileBytes, u == 0 || u == U);
          umma_commit(&empty_bar[stage_of(gblk + u)]);
          if (u + 2 < U2) issue1(u + 2);
        }
        umma_commit(&units_done);
        const int64_t gt = gblk + U2;
        wait_block(gt);
        mma8(kFLIN, kFZ, ring + stage_of(gt) * kBlockBytes, true);
        mbar_wait(&tail_a, (uint32_t)(lt & 1));
        tc_fence_after();
        mma8(kFAS, kFHB, ring + stage_of(gt) * kBlockBytes + kTileBytes, true);
        umma_commit(&empty_bar[stage_of(gt)]);
        umma_commit(&heads_full);
        gblk += n_blocks;
      }
    }
    __syncwarp();
  } else {
    // ================= weight loader =================
    if (lane == 0) {
      const int64_t total = (int64_t)my_tiles * n_blocks;
      for (int64_t g = 0; g < total; ++g) {
        const int s = (int)(g % n_stages);
        if (g >= n_stages) mbar_wait(&empty_bar[s], (uint32_t)(((g / n_stages) - 1) & 1));
        mbar_expect_tx(&full_bar[s], kBlockBytes);
        bulk_g2s(smem + (size_t)s * kBlockBytes, p.pack + (size_t)(g % n_blocks) * kBlockBytes, kBlockBytes, &full_bar[s]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kRowWarps) tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------------------------------------------------
// input gradient of one transition on a row tile
// ---------------------------------------------------------------------------------------------------------
struct BwdParams {
  const unsigned char* pack;     // backward pack
  const float* d_g; const float* d_nl; const float* d_lin;   // (R, 64): gradients at the gate / nonlinear (complete,
                                                              // std path included) / linear heads
  const uint32_t* relu_bits;     // [row tile][unit][half][128 rows] from the KEEP forward
  float* dz;                     // (R, 64) out
  __half* dh16;                  // masked hidden gradients, FP16 tiles [row group][2U atoms]
  __half* dg16; __half* dnl16;   // FP16 tiles of d_g / d_nl [row group][1 atom]
  float* gb_gate0; float* gb_non0;   // bias gradients of the two hidden layers (H each), accumulated; both null = skip
  int64_t R;
  int H;
  int n_stages;
};
constexpr uint32_t kBDG = 0, kBDNL = 64, kBDZ = 128, kBHB = 192;        // 4 hidden buffers + d_lin buffer (index 4)

__global__ void __launch_bounds__(kThreads, 1) gtf_bwd_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ unsigned char fused_smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t in_full, dz_full, dz_empty;
  __shared__ __align__(8) uint64_t d_full[4];
  __shared__ __align__(8) uint64_t a_full[4];
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = fused_smem_dyn + ((1024u - (smem_u32(fused_smem_dyn) & 1023u)) & 1023u);
  const int n_stages = p.n_stages;
  float* gb_s = reinterpret_cast<float*>(smem + (size_t)n_stages * kBlockBytes);      // (2H) column sums of this CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, U = H / kHU, U2 = 2 * U, n_blocks = U2 + 1;
  const int64_t n_tiles = (p.R + kTileRows - 1) / kTileRows;
  const int my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;

  if (warp == kRowWarps) tmem_alloc(&tmem_base_s, 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&in_full, kRowWarps); mbar_init(&dz_full, 1); mbar_init(&dz_empty, kRowWarps);
    for (int i = 0; i < 4; ++i) { mbar_init(&d_full[i], 1); mbar_init(&a_full[i], kRowWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) gb_s[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;

  if (warp < kRowWarps) {
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t tl = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32);
    uint32_t par_d = 0, par_misc = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)lt * gridDim.x;
      const int64_t row = tile * kTileRows + q * 32 + lane;
      const bool row_ok = row < p.R;
      // ---- head gradients -> TMEM (A operands), FP16 tiles for the weight-gradient GEMM
      {
        float v[32];
        load_row32(p.d_g + row * kZ + hf * 32, row_ok, v);
        store_half32(p.dg16, row, 1, 0, hf, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        tmem_st32(tl + kBDG, v);
        load_row32(p.d_nl + row * kZ + hf * 32, row_ok, v);
        store_half32(p.dnl16, row, 1, 0, hf, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        tmem_st32(tl + kBDNL, v);
        load_row32(p.d_lin + row * kZ + hf * 32, row_ok, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        tmem_st32(tl + kBHB + 4 * 64, v);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&in_full);
      }
#pragma unroll 1
      for (int u = 0; u < U2; ++u) {
        const int hb = u & 3;
        const uint32_t bits = __ldg(p.relu_bits + ((tile * U2 + u) * 2 + hf) * kTileRows + q * 32 + lane);
        mbar_wait(&d_full[hb], (par_d >> hb) & 1u);
        par_d ^= 1u << hb;
        tc_fence_after();
        float v[32];
        tmem_ld32(tl + kBHB + hb * 64, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.f;
        store_half32(p.dh16, row, U2, u, hf, v);
        if (p.gb_gate0 != nullptr) {
          // column sums over the warp's 32 rows by a transposing butterfly: after 5 exchange steps lane j holds the
          // sum of column j (31 shuffles instead of 160)
          float s[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) s[j] = v[j];
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1) {
            const bool up = (lane & w) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              const float mine = up ? s[j + w] : s[j];            // the half this lane keeps
              const float give = up ? s[j] : s[j + w];            // the half its partner keeps
              s[j] = mine + __shfl_xor_sync(0xffffffffu, give, w);
            }
          }
          // lane's column: bit-reversal-free order — after the steps lane l holds column index equal to l's bits
          // consumed high to low, i.e. column l
          atomicAdd(gb_s + u * kHU + hf * 32 + lane, s[0]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        tmem_st32(tl + kBHB + hb * 64, v);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[hb]);
      }
      {
        mbar_wait(&dz_full, par_misc & 1u);
        tc_fence_after();
        float v[32];
        tmem_ld32(tl + kBDZ, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dz_empty);
        if (row_ok) store_row32(p.dz + row * kZ + hf * 32, v);
      }
      par_misc ^= 1u;
    }
  } else if (warp == kRowWarps) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kTileRows, 64);
      const uint32_t ring = smem_u32(smem);
      uint32_t par_a = 0;
      int64_t gblk = 0;
      auto stage_of = [&](int64_t g) { return (int)(g % n_stages); };
      auto wait_block = [&](int64_t g) { mbar_wait(&full_bar[stage_of(g)], (uint32_t)((g / n_stages) & 1)); tc_fence_after(); };
      auto mma8 = [&](uint32_t d_col, uint32_t a_col, uint32_t tile_addr, bool fresh) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_tf32_ts(tb + d_col, tb + a_col + k * 8, umma_desc_sw128(tile_addr + (k >> 2) * 8192 + (k & 3) * 32), idesc,
                       (fresh && k == 0) ? 0u : 1u);
      };
      for (int lt = 0; lt < my_tiles; ++lt) {
        mbar_wait(&in_full, (uint32_t)(lt & 1));
        tc_fence_after();
        // head block: dz = d_lin W_lin (starts the accumulator)
        mbar_wait(&dz_empty, (uint32_t)((lt & 1) ^ 1));
        tc_fence_after();
        wait_block(gblk);
        mma8(kBDZ, kBHB + 4 * 64, ring + stage_of(gblk) * kBlockBytes + kTileBytes, true);
        umma_commit(&empty_bar[stage_of(gblk)]);
        auto issue1 = [&](int u) {                    // hidden gradients of unit u: d_head W2u
          wait_block(gblk + 1 + u);
          mma8(kBHB + (u & 3) * 64, u < U ? kBDG : kBDNL, ring + stage_of(gblk + 1 + u) * kBlockBytes, true);
          umma_commit(&d_full[u & 3]);
        };
        issue1(0);
        if (U2 > 1) issue1(1);
        if (U2 > 2) issue1(2);
        for (int u = 0; u < U2; ++u) {
          const int hb = u & 3;
          mbar_wait(&a_full[hb], (par_a >> hb) & 1u);
          par_a ^= 1u << hb;
          tc_fence_after();
          mma8(kBDZ, kBHB + hb * 64, ring + stage_of(gblk + 1 + u) * kBlockBytes + kTileBytes, false);
          umma_commit(&empty_bar[stage_of(gblk + 1 + u)]);
          if (u + 3 < U2) issue1(u + 3);
        }
        umma_commit(&dz_full);
        gblk += n_blocks;
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      const int64_t total = (int64_t)my_tiles * n_blocks;
      for (int64_t g = 0; g < total; ++g) {
        const int s = (int)(g % n_stages);
        if (g >= n_stages) mbar_wait(&empty_bar[s], (uint32_t)(((g / n_stages) - 1) & 1));
        mbar_expect_tx(&full_bar[s], kBlockBytes);
        bulk_g2s(smem + (size_t)s * kBlockBytes, p.pack + (size_t)(g % n_blocks) * kBlockBytes, kBlockBytes, &full_bar[s]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (p.gb_gate0 != nullptr)
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x)
      if (gb_s[i] != 0.f) atomicAdd(i < H ? p.gb_gate0 + i : p.gb_non0 + (i - H), gb_s[i]);
  if (warp == kRowWarps) tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------------------------------------------------
// weight gradients of the four H-wide layers from the FP16 operand tiles
// ---------------------------------------------------------------------------------------------------------
// out^T[h, zc] += sum_rows X[row, h] * Y[row, zc].  X: FP16 tiles [row group][n_atoms_x atoms], the problem uses atoms
// atom0 .. (H/64 of them); Y: FP16 tiles [row group][1 atom].  out is either (H, Z) row-major (dW of a z -> hidden
// layer: direct) or (Z, H) row-major (dW of a hidden -> head layer: transposed add).
struct Wgrad16Problem {
  const __half* X; int n_atoms_x; int atom0;
  const __half* Y;
  float* out; int transposed;
};
struct Wgrad16Params {
  Wgrad16Problem pr[4];
  int n_problems;
  int H;
  int64_t n_groups;            // row groups of 64 rows (R rounded up to 128 rows, tiles past R are zero)
  int groups_per_slice;        // K split: a work item contracts this many row groups
  int n_slices;
  int n_stages;
};
constexpr int kWgStageBytes = 3 * kAtomBytes;        // X atoms (2) + Y atom
constexpr int kWgThreads = 6 * 32;                   // 4 epilogue warps, MMA warp, loader warp

// MN-major SWIZZLE_128B shared-memory descriptor: LBO = stride between 64-element MN atoms, SBO = stride between
// 8-row K groups (cute::UMMA canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)), T = 8 halves)
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16, FP16 operands, FP32 accumulate, both operands MN-major (bits 15 / 16)
__device__ __forceinline__ uint32_t umma_idesc_f16_mn(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad16_kernel(const __grid_constant__ Wgrad16Params p) {
  extern __shared__ unsigned char fused_smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = fused_smem_dyn + ((1024u - (smem_u32(fused_smem_dyn) & 1023u)) & 1023u);
  const int n_stages = p.n_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.H / 128;
  const int n_items = p.n_problems * m_tiles * p.n_slices;
  const int my_items = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto item_of = [&](int li, int& pi, int& mt, int64_t& g0, int64_t& g1) {
    const int it = (int)blockIdx.x + li * (int)gridDim.x;
    const int sl = it % p.n_slices, rest = it / p.n_slices;
    mt = rest % m_tiles; pi = rest / m_tiles;
    g0 = (int64_t)sl * p.groups_per_slice;
    g1 = g0 + p.groups_per_slice < p.n_groups ? g0 + p.groups_per_slice : p.n_groups;
  };
  if (warp == 4) tmem_alloc(&tmem_base_s, 64);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1); mbar_init(&acc_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;

  if (warp < 4) {
    // ===== epilogue: accumulator (128 hidden rows x 64 z columns) -> red.global.add =====
    for (int li = 0; li < my_items; ++li) {
      int pi, mt; int64_t g0, g1;
      item_of(li, pi, mt, g0, g1);
      const Wgrad16Problem& pr = p.pr[pi];
      mbar_wait(&acc_full, (uint32_t)(li & 1));
      tc_fence_after();
      const int h = mt * 128 + warp * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        if (c == 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty);
        }
        if (g1 > g0) {
          if (pr.transposed) {                       // out (Z, H): lanes = consecutive h -> coalesced per column
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(pr.out + (size_t)(c + j) * p.H + h, v[j]);
          } else {                                   // out (H, Z): a thread owns 32 consecutive columns of its row
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(pr.out + (size_t)h * kZ + c + j, v[j]);
          }
        }
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16_mn(128, 64);
      const uint32_t ring = smem_u32(smem);
      int64_t gpos = 0;
      for (int li = 0; li < my_items; ++li) {
        int pi, mt; int64_t g0, g1;
        item_of(li, pi, mt, g0, g1);
        mbar_wait(&acc_empty, (uint32_t)((li & 1) ^ 1));
        tc_fence_after();
        for (int64_t g = g0; g < g1; ++g, ++gpos) {
          const int s = (int)(gpos % n_stages);
          mbar_wait(&full_bar[s], (uint32_t)((gpos / n_stages) & 1));
          tc_fence_after();
          const uint32_t xa = ring + s * kWgStageBytes, ya = xa + 2 * kAtomBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)                 // 16 rows (two 8-row K groups = 2 KB) per instruction
            umma_f16_ss(tb, umma_desc_mn128(xa + k * 2048, kAtomBytes, 1024), umma_desc_mn128(ya + k * 2048, kAtomBytes, 1024),
                        idesc, (g > g0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full);
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      int64_t gpos = 0;
      for (int li = 0; li < my_items; ++li) {
        int pi, mt; int64_t g0, g1;
        item_of(li, pi, mt, g0, g1);
        const Wgrad16Problem& pr = p.pr[pi];
        for (int64_t g = g0; g < g1; ++g, ++gpos) {
          const int s = (int)(gpos % n_stages);
          if (gpos >= n_stages) mbar_wait(&empty_bar[s], (uint32_t)(((gpos / n_stages) - 1) & 1));
          unsigned char* dst = smem + (size_t)s * kWgStageBytes;
          const unsigned char* xs = reinterpret_cast<const unsigned char*>(pr.X) +
                                    ((size_t)g * pr.n_atoms_x + pr.atom0 + mt * 2) * kAtomBytes;
          mbar_expect_tx(&full_bar[s], kWgStageBytes);
          bulk_g2s(dst, xs, 2 * kAtomBytes, &full_bar[s]);                   // two adjacent 64-column atoms
          bulk_g2s(dst + 2 * kAtomBytes, reinterpret_cast<const unsigned char*>(pr.Y) + (size_t)g * kAtomBytes, kAtomBytes,
                   &full_bar[s]);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, 64);
}
#endif  // !BFVI_EMU

}  // namespace fused
}  // namespace bfvi
