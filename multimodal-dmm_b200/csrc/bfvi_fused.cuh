// bfvi_fused.cuh — fused on-chip GaussianGTF kernels of the large-dim family (precision mode BFVI_PREC_FUSED).
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
//   gtf_fwd_kernel   per row tile, per PAIR of 64-column hidden units (gate unit c, nonlinear unit c):
//                    tcgen05.mma D[128,128] = z [W0g_c ; W0n_c]^T into TMEM (one N = 128 instruction stream for both
//                    branches)  ->  row warps tcgen05.ld, scale + bias, ReLU, split into FP16 hi | lo, tcgen05.st back
//                    IN PLACE as the A operands  ->  tcgen05.mma gate head += A_g W2g_c^T, nonlinear head += A_n W2n_c^T.
//                    z itself is an A operand in TMEM (one tcgen05.st per tile).  Heads (pre-sigmoid gate, nonlinear,
//                    linear, pre-softplus std) leave as (R, Z) fp32 rows.  KEEP mode (the backward recompute) also writes
//                    the hidden activations as FP16 operand tiles for the weight-gradient GEMM, the ReLU sign bits
//                    (64 per row and unit) and an FP16 tile of z.
//   gtf_bwd_kernel   same pipeline for the input gradient:  D = d_head W2u  ->  mask by the ReLU bits, FP16 tile for the
//                    weight gradients, round to TF32, A in place  ->  dz += A W0u.
//   wgrad16_kernel   dW^T[H, Z] += X^T Y over all rows: X (hidden activations / their gradients) and Y (z / head
//                    gradients) are the FP16 tiles the two kernels above wrote in the MN-major SWIZZLE_128B shared-memory
//                    image, so a stage is two cp.async.bulk copies; kind::f16 MMAs, FP32 accumulation in TMEM over a
//                    slice of the rows, one red.global.add pass per work item; a constant all-ones column appended to
//                    the Y operand (N = 80) makes the MMA produce the column sums of X (bias gradients) for free.
//
// Precision.  Forward contractions are error-compensated FP16 products: every operand x is split into hi = fp16(x),
// lo = fp16(x - hi) and a_hi b_hi + a_lo b_hi + a_hi b_lo accumulates in FP32 (~2^-21 relative per product; weights are
// pre-scaled by a power of two per layer so that their residuals are normal FP16 numbers).  That is FP32-class, and it has
// to be: the ReLU derivative is discontinuous, so a hidden pre-activation whose SIGN differs from the reference's flips a
// whole gradient term — with single-pass TF32 hidden layers d_z was off by 8e-3 (tests/test_gpu_fused.py).  Gradients
// contract single-pass TF32 (input gradient) and FP16 (H-wide weight gradients) operands with FP32 accumulation.
//
// Weights are laid out ONCE per step as the exact shared-memory images of every operand tile (pack_gtf_kernel, K-major
// SWIZZLE_128B) and streamed through a ring of 32 KB stages by ONE thread with cp.async.bulk (global -> shared, mbarrier
// complete_tx; measured 57 B/clk per SM from L2 with all 148 SMs pulling, tools/probe_bulk_rate.cu) — no per-tile
// rounding pass, no LDGSTS, no converter warps.  Row results leave through padded shared-memory patches and one
// cp.async.bulk (shared -> global) per row segment: a thread = row layout otherwise touches 32 cache lines per store
// instruction.
//
// Warp roles (320 threads): warps 0-7 "row warps" (thread = tile row = TMEM lane; two warps per 32-lane quadrant, one
// per 32-column half), warp 8 MMA issuer (one elected thread), warp 9 weight loader (one elected thread).
// Measured on B200 (tools/probe_mma_rate.cu): one M = 128 tcgen05.mma costs 52 / 67 / 131 cycles at N = 64 / 128 / 256
// whatever the kind or the A source — N = 64 instructions run the tensor pipe at 61 % — hence the N = 128 hidden layers.
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
constexpr int kBlockBytes = 32768;     // one weight block (one ring stage)
constexpr int kTileBytes = 16384;
constexpr int kRowWarps = 8;
constexpr int kThreads = (kRowWarps + 2) * 32;
constexpr int kMaxStages = 9;
constexpr int kRowGroup = 64;          // rows per FP16 operand tile of the weight-gradient GEMM (one K stage)
constexpr int kAtomBytes = 8192;       // 64 rows x 64 halves, MN-major SWIZZLE_128B
constexpr int kPatchBytes = 16384;
// development ablations (timing only, results are garbage): skip FP16 tile stores / fp32 row stores / ReLU bits /
// second-level MMAs (heads, dz) / first-level MMAs (hidden) / row-warp arithmetic
// (masks are compiled in only with -DBFVI_FUSED_ABLATE: python tools/variants.py ablate:BFVI_FUSED_ABLATE, then
// BFVI_LIB_PATH=tools/_variants/libbfvi_ablate.so python tools/probe_fused_ablate.py; the product build tests nothing)
#ifdef BFVI_FUSED_ABLATE
#define BFVI_ABL(f) (p.abl & (f))
#else
#define BFVI_ABL(f) 0
#endif
constexpr int kAblTile16 = 1, kAblRows = 2, kAblBits = 4, kAblMma2 = 8, kAblMma1 = 16, kAblMath = 32;
// kStoreLsu (64, a real variant, results valid): results leave through coalesced st.global by the pair's 64 threads instead
// of cp.async.bulk (whose requests queue behind the weight ring's bulk loads in the SM's one copy engine)
constexpr int kStoreLsu = 64;
// 128 (timing only): the loaders stop copying after the first lap of the ring (stale operands): is the ring the limit?
constexpr int kAblRing = 128;
// 256 (timing only): no proxy fence before the bulk stores (fence.proxy.async = MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in SASS)
constexpr int kAblFence = 256;     // per 32-lane quadrant (two row warps): two 32 rows x 256 B slots = four 32 x 128 B slots

// shared-memory image of a 64 x 64 fp32 operand tile, K-major SWIZZLE_128B: two K halves of 32 floats; row r of a
// half is one 128-byte line whose 16-byte chunks are XOR-permuted by r % 8 (8-row groups 1024 B apart)
__host__ __device__ inline int tile_offset(int r, int k) {
  return (k >> 5) * 8192 + r * 128 + (((((k & 31) >> 2) ^ (r & 7))) << 4) + (k & 3) * 4;
}
// FP16 operand tile with 64 K values per row (one 128-byte line per row), element (row r, k)
__host__ __device__ inline int tile16_offset(int r, int k) { return r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2; }

// Weight packs of one GaussianGTF, 2U + 1 blocks of 32 KB each (U = H / 64).
// FWD (FP16 hi | lo tiles, weights scaled by a power of two per layer):
//   block 2c     hidden layers of pair c: 128 rows (64 gate units, 64 nonlinear units) x K = z: hi 16 KB | lo 16 KB
//   block 2c + 1 head layers of pair c: gate W2[:, unit cols] (N = z, K = hidden) hi 8 KB | lo 8 KB, then nonlinear
//   block 2U     tail: W_lin hi | lo, W_std hi | lo
// BWD (TF32 tiles): block 0 = head (tile 1 = W_lin^T), blocks 1..2U = units (gate branch first):
//   tile 0 = W2[:, unit cols]^T (N = hidden, K = z_out), tile 1 = W0[unit rows, :]^T (N = z_in, K = hidden)
// Bias table: b0 gate (H), b0 nonlin (H), gate2_b, nonlin2_b, lin_b, std_b (Z each), then 8 floats: 1 / scale of
// gate0, nonlin0, gate2, nonlin2, lin, std (+ 2 spares).
struct PackParams {
  const float* w_gate0; const float* b_gate0; const float* w_gate2; const float* b_gate2;
  const float* w_lin; const float* b_lin; const float* w_non0; const float* b_non0;
  const float* w_non2; const float* b_non2; const float* w_std; const float* b_std;
  unsigned char* fwd; unsigned char* bwd; float* bias;
  int H;
};
inline size_t pack_blocks(int H) { return (size_t)(2 * (H / kHU) + 1); }
inline size_t pack_bytes(int H) { return pack_blocks(H) * kBlockBytes; }
inline size_t bias_floats(int H) { return (size_t)2 * H + 4 * kZ + 8; }

#ifndef BFVI_EMU
using tc::smem_u32;
using tc::mbar_init; using tc::mbar_wait; using tc::mbar_arrive; using tc::umma_commit;
using tc::tmem_alloc; using tc::tmem_dealloc; using tc::tmem_ld32; using tc::tc_fence_before; using tc::tc_fence_after;
using tc::umma_desc_sw128; using tc::umma_idesc_tf32; using tc::umma_tf32_ts; using tc::rn_tf32; using tc::tmem_wait_st;
using tc::fence_async_smem;

// power-of-two scale of one layer: max |w| * s in [2^13, 2^14)
// blocks 0..5: power-of-two scale of one layer; blocks 6, 7: column L1 norms that bound what a head gradient can become
// downstream (slot 6: max_h sum_zc |W2[zc, h]| over both hidden -> head layers bounds |dh| / max|d_head|; slot 7:
// max_j sum_i |W_std[i, j]| bounds the std path's share of d_nl / max|d_as|) — the input-gradient kernel scales its
// operands by a power of two derived from them so that the FP16 weight-gradient tiles use the FP16 range whatever the
// scale of the loss (tools/probe_f16_range.py)
__global__ void __launch_bounds__(256) gtf_scale_kernel(const __grid_constant__ PackParams p) {
  __shared__ float red[256];
  const int H = p.H, layer = blockIdx.x;
  if (layer >= 6) {
    float m = 0.f;
    if (layer == 6) {
      for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float sg = 0.f, sn = 0.f;
        for (int zc = 0; zc < kZ; ++zc) { sg += fabsf(p.w_gate2[(size_t)zc * H + h]); sn += fabsf(p.w_non2[(size_t)zc * H + h]); }
        m = fmaxf(m, fmaxf(sg, sn));
      }
    } else {
      for (int j = threadIdx.x; j < kZ; j += blockDim.x) {
        float sa = 0.f;
        for (int i = 0; i < kZ; ++i) sa += fabsf(p.w_std[i * kZ + j]);
        m = fmaxf(m, sa);
      }
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
      __syncthreads();
    }
    if (threadIdx.x == 0) p.bias[2 * H + 4 * kZ + layer] = red[0];
    return;
  }
  const float* w = layer == 0 ? p.w_gate0 : layer == 1 ? p.w_non0 : layer == 2 ? p.w_gate2 : layer == 3 ? p.w_non2
                   : layer == 4 ? p.w_lin : p.w_std;
  const int n = layer < 4 ? H * kZ : kZ * kZ;
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = fabsf(w[i]);
    m = (a > m && a < 3.0e38f) ? a : m;              // ignore inf / NaN weights: they poison the step anyway
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float inv = 1.f;
    if (red[0] > 0.f) {
      int e;
      frexpf(red[0], &e);                            // red[0] = f * 2^e, f in [0.5, 1)  ->  max * 2^(14 - e) in [2^13, 2^14)
      inv = ldexpf(1.f, e - 14);
    }
    p.bias[2 * H + 4 * kZ + layer] = inv;
  }
}

__device__ __forceinline__ void put_split(unsigned char* hi_tile, int lo_off, int off, float x) {
  const __half hi = __float2half_rn(x);
  *reinterpret_cast<__half*>(hi_tile + off) = hi;
  *reinterpret_cast<__half*>(hi_tile + lo_off + off) = __float2half_rn(x - __half2float(hi));
}

__global__ void __launch_bounds__(256) pack_gtf_kernel(const __grid_constant__ PackParams p) {
  const int H = p.H, U = H / kHU, nb = 2 * U + 1;
  const int b = blockIdx.x;                                  // block index, both packs
  unsigned char* fw = p.fwd + (size_t)b * kBlockBytes;
  float* bw = reinterpret_cast<float*>(p.bwd + (size_t)b * kBlockBytes);
  const float* inv = p.bias + 2 * H + 4 * kZ;                // written by gtf_scale_kernel (previous launch)
  for (int e = threadIdx.x; e < 2 * 64 * 64; e += blockDim.x) {
    const int t = e >> 12, r = (e >> 6) & 63, k = e & 63;
    // ---- forward pack (x / inv is exact: inv is a power of two)
    if (b == 2 * U) {                                        // tail: lin | std
      const float v = t == 0 ? p.w_lin[r * kZ + k] / inv[4] : p.w_std[r * kZ + k] / inv[5];
      put_split(fw + t * kTileBytes, 8192, tile16_offset(r, k), v);
    } else if ((b & 1) == 0) {                               // hidden layers of pair c: rows 0-63 gate, 64-127 nonlinear
      const int c = b >> 1;
      const float v = t == 0 ? p.w_gate0[(size_t)(c * kHU + r) * kZ + k] / inv[0] : p.w_non0[(size_t)(c * kHU + r) * kZ + k] / inv[1];
      put_split(fw, kTileBytes, tile16_offset(t * 64 + r, k), v);
    } else {                                                 // head layers of pair c: gate | nonlinear
      const int c = b >> 1;
      const float v = t == 0 ? p.w_gate2[(size_t)r * H + c * kHU + k] / inv[2] : p.w_non2[(size_t)r * H + c * kHU + k] / inv[3];
      put_split(fw + t * kTileBytes, 8192, tile16_offset(r, k), v);
    }
    // ---- backward pack
    float v;
    if (b == 0) {
      v = t == 0 ? 0.f : p.w_lin[k * kZ + r];
    } else {
      const int u = b - 1, br = u / U, c = u % U;
      const float* w0 = br ? p.w_non0 : p.w_gate0;
      const float* w2 = br ? p.w_non2 : p.w_gate2;
      v = t == 0 ? w2[(size_t)k * H + c * kHU + r] : w0[(size_t)(c * kHU + k) * kZ + r];
    }
    bw[(t * kTileBytes + tile_offset(r, k)) >> 2] = rn_tf32(v);
  }
  if (b == nb - 1) {
    for (int i = threadIdx.x; i < H; i += blockDim.x) { p.bias[i] = p.b_gate0[i]; p.bias[H + i] = p.b_non0[i]; }
    for (int i = threadIdx.x; i < kZ; i += blockDim.x) {
      p.bias[2 * H + i] = p.b_gate2[i]; p.bias[2 * H + kZ + i] = p.b_non2[i];
      p.bias[2 * H + 2 * kZ + i] = p.b_lin[i]; p.bias[2 * H + 3 * kZ + i] = p.b_std[i];
    }
  }
}

// maxima of |a| and |b| over n floats as float bit patterns (stand-alone entries: inside a step bwd_rows_kernel keeps them)
__global__ void __launch_bounds__(256) absmax2_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                      unsigned* __restrict__ out) {
  float ma = 0.f, mb = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    ma = fmaxf(ma, fabsf(a[i])); mb = fmaxf(mb, fabsf(b[i]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (__float_as_uint(ma) > *(volatile unsigned*)(out + 0)) atomicMax(out + 0, __float_as_uint(ma));
    if (__float_as_uint(mb) > *(volatile unsigned*)(out + 1)) atomicMax(out + 1, __float_as_uint(mb));
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
// TMA bulk copy shared -> global (SASS: UBLKCP.G.S), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One elected lane of a CONVERGED warp (ptxas knows the guarded region runs in a single thread and keeps the tcgen05
// operands in uniform registers; a `lane == 0` test gives it no such guarantee).  Always elects the same lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float relu_nan(float x) {
  float y;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(y) : "f"(x));
  return y;
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
// Row results leave through "pair patches": the two row warps of a 32-lane quadrant (column halves 0 and 1) share an
// 8 KB shared-memory patch that holds the 32 rows exactly as they lie in global memory, so ONE elected thread sends them
// with ONE cp.async.bulk (per-thread bulk copies were measured at ~50 cycles of TMA issue each: 256 of them per unit).
// Named barrier 1 + q orders the two warps around the patch.
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); }

// 32 rows x 64 fp32 columns (256 B per row, 8 KB): thread (row = lane, half hf) writes its 128 bytes.  `n_valid` rows of
// the quadrant exist (the copy is clipped to them).  The array is stored in the "swz64" layout (bfvi_generic.cuh
// row64): the 16-byte chunks of each 128-byte half row are XOR-permuted by row % 8 — a linear row layout puts all 32
// lanes of a store on the same four banks (rows are 64 words apart; measured 2 500 cycles per 8 KB) — and the consumers
// (step / bwd_rows / carry / match kernels) apply the same permutation when they read.
__device__ __forceinline__ void store_rows_f32(unsigned char* pp, int& slot, int q, int lane, int hf, bool elected,
                                               float* __restrict__ dst_row0, int n_valid, const float (&v)[32],
                                               bool lsu = false) {
  unsigned char* base = pp + slot * 8192;
  unsigned char* prow = base + lane * 256 + hf * 128;
  const int sw = lane & 7;                           // = row % 8 (row0 is a multiple of 32)
  if (lsu) {
    // two slots alternate and every call has one barrier between its writes and its reads: a thread can run at most
    // one barrier ahead, i.e. write the OTHER slot while a slower thread still reads this one
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(prow + ((i ^ sw) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    pair_sync(q);
    const int tid = hf * 32 + lane;
    unsigned char* dst = reinterpret_cast<unsigned char*>(dst_row0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {                    // 64 threads x 16 B = 1 KB (four rows) per round, full lines
      const int chunk = i * 64 + tid;
      if ((chunk >> 4) < n_valid) __stcs(reinterpret_cast<float4*>(dst + chunk * 16), *reinterpret_cast<const float4*>(base + chunk * 16));
    }
    slot ^= 1;
    return;
  }
  if (elected) bulk_wait_read<1>();                  // the copy that used this slot two stores ago has read it
  pair_sync(q);
#pragma unroll
  for (int i = 0; i < 8; ++i)                        // swz64: chunk i of the half row sits at i ^ (row % 8)
    *reinterpret_cast<float4*>(prow + ((i ^ sw) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  fence_async_smem();
  pair_sync(q);
  if (elected) {
    if (n_valid > 0) bulk_s2g(dst_row0, base, (uint32_t)n_valid * 256u);
    bulk_commit();
  }
  slot ^= 1;
}
// the patch changes hands between FP16 tile slots and fp32 row slots (they overlap): every earlier copy has read it
__device__ __forceinline__ void patch_switch(int q, bool elected, bool lsu) {
  if (lsu) pair_sync(q);
  else if (elected) bulk_wait_read<0>();
}
// FP16 operand tiles of the weight-gradient GEMM: (row group of 64 rows) x (atom of 64 columns) = 8 KB, row r of the
// group is a 128-byte line, 16-byte chunk c stored at c ^ (r % 8): the MN-major SWIZZLE_128B shared-memory image, so
// the GEMM loads a tile with one bulk copy.  Tiles of a multi-atom array (hidden-sized: 2U atoms) are laid out
// [atom pair][row group][2 atoms]: a weight-gradient work item (128 hidden columns = one atom pair, a slice of the row
// groups) then streams ONE contiguous run of 16 KB blocks — with row-group-major tiles the same item read 16 KB every
// 128 KB, 304 such streams at once, and the launch reached 45 % of the DRAM bandwidth
// (profiles/r2_wgrad16_big_full.txt).  The quadrant's 32 rows are 4 KB contiguous in that image; thread (row,
// half) writes its four chunks (conflict-free thanks to the XOR) into half `buf` of the pair patch, which leaves as one
// 4 KB bulk copy.  Tiles are padded to whole row tiles, so rows past the end are written (as zeros) too.
__device__ __forceinline__ void store_tile_f16(unsigned char* pp, int& buf, int q, int lane, int hf, bool elected,
                                               __half* __restrict__ base, int64_t row0, int64_t n_groups, int n_atoms,
                                               int atom, const float (&v)[32], bool lsu = false, bool nofence = false) {
  if (!lsu) {
    if (elected) bulk_wait_read<3>();                // the copy that used this slot four stores ago has read it
    pair_sync(q);
  }
  unsigned char* prow = pp + buf * 4096 + lane * 128;
  const int sw = lane & 7;                           // row0 is a multiple of 32
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 pk;
    __half2 h0 = __floats2half2_rn(v[8 * c], v[8 * c + 1]), h1 = __floats2half2_rn(v[8 * c + 2], v[8 * c + 3]);
    __half2 h2 = __floats2half2_rn(v[8 * c + 4], v[8 * c + 5]), h3 = __floats2half2_rn(v[8 * c + 6], v[8 * c + 7]);
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(prow + (((hf * 4 + c) ^ sw) << 4)) = pk;
  }
  const size_t g = (size_t)(row0 / kRowGroup);
  const size_t idx = n_atoms == 1 ? g : ((size_t)(atom >> 1) * (size_t)n_groups + g) * 2 + (size_t)(atom & 1);
  unsigned char* tile = reinterpret_cast<unsigned char*>(base) + idx * kAtomBytes + (size_t)(row0 % kRowGroup) * 128;
  if (lsu) {                                         // (slots rotate: see store_rows_f32)
    pair_sync(q);
    const int tid = hf * 32 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int chunk = i * 64 + tid;
      __stcs(reinterpret_cast<uint4*>(tile + chunk * 16), *reinterpret_cast<const uint4*>(pp + buf * 4096 + chunk * 16));
    }
    buf = (buf + 1) & 3;
    return;
  }
  if (!nofence) fence_async_smem();
  pair_sync(q);
  if (elected) {
    bulk_s2g(tile, pp + buf * 4096, 4096);
    bulk_commit();
  }
  buf = (buf + 1) & 3;
}

// kind::f16 with FP16 operands, FP32 accumulate, both operands K-major; the A operand from tensor memory (two halves
// per 32-bit column: 8 columns = one K = 16 instruction)
__device__ __forceinline__ uint32_t umma_idesc_f16_k(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
// A thread's 32 fp32 values -> the A operand of a 64-wide contraction in tensor memory as FP16 pairs: 16 columns of
// hi = fp16(x) followed by 16 columns of lo = fp16(x - hi), INSIDE the thread's own 32-column half (so the in-place
// D -> A rewrite never touches columns the other half-warp still has to read).  k-step ks (16 values) reads hi at
// (ks >> 1) * 32 + (ks & 1) * 8 and lo 16 columns further.
__device__ __forceinline__ void store_a_split(uint32_t taddr_half, const float (&v)[32]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
  }
  tmem_st16u(taddr_half, hi);
  tmem_st16u(taddr_half + 16, lo);
}
// D[128, N] (+)= A[128, 64] B[N, 64]^T as three FP16 products; A (hi | lo pairs) from tensor memory, B tiles hi at
// tile_addr and lo at tile_addr + lo_off (one 128-byte line per row)
__device__ __forceinline__ void mma_split(uint32_t tb, uint32_t d_col, uint32_t a_col, uint32_t tile_addr, uint32_t lo_off,
                                          bool fresh, uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t ah = tb + a_col + (ks >> 1) * 32 + (ks & 1) * 8, al = ah + 16;
    const uint64_t bh = umma_desc_sw128(tile_addr + ks * 32), bl = umma_desc_sw128(tile_addr + lo_off + ks * 32);
    umma_f16_ts(tb + d_col, ah, bh, idesc, (fresh && ks == 0) ? 0u : 1u);
    umma_f16_ts(tb + d_col, al, bh, idesc, 1u);
    umma_f16_ts(tb + d_col, ah, bl, idesc, 1u);
  }
}

// Ring positions are small unsigned counters divided by a launch constant (ring depth, blocks per tile): a multiply
// by the rounded-up reciprocal is exact for g < 2^32 / n.  The 64-bit `%` and `/` these loops used first are ~50
// dependent instructions EACH in the one thread that issues the MMAs / the bulk copies (measured: 720 cycles per
// ring step with nothing else left in the loop, tools/probe_fused_ablate.py masks 191 / 0).
struct FastDiv {
  uint32_t n, m;
  __device__ __forceinline__ explicit FastDiv(int n_) : n((uint32_t)n_), m((uint32_t)(0x100000000ull / (uint32_t)n_) + 1u) {}
  __device__ __forceinline__ uint32_t div(uint32_t g) const { return __umulhi(g, m); }
  __device__ __forceinline__ uint32_t mod(uint32_t g) const { return g - __umulhi(g, m) * n; }
};

struct FwdParams {
  const unsigned char* pack;     // forward pack
  const float* bias;             // bias + scale table (see PackParams)
  const float* z;                // (R, 64)
  float* g; float* nl; float* lin; float* as;      // (R, 64) heads, biases added
  __half* h16;                   // KEEP: hidden activations, FP16 tiles [atom pair][row group][2 atoms]
  uint32_t* relu_bits;           // KEEP: sign bits of the hidden activations, [row tile][unit][half][128 rows] words
  __half* z16;                   // KEEP: FP16 tiles of z [row group][1 atom]
  int64_t R;
  int H;
  int n_stages;
  long long* dbg;                // development: cycle counters of CTA 0 (BFVI_FUSED_DBG=1), nullable
  int abl;                       // development: ablation mask (BFVI_FUSED_ABL, tools/probe_fused_ablate.py); 0 in the product
};
// cycle counters of CTA 0: compiled in only with -DBFVI_FUSED_COUNTERS (python tools/variants.py counters:BFVI_FUSED_COUNTERS;
// BFVI_LIB_PATH=tools/_variants/libbfvi_counters.so BFVI_FUSED_DBG=1 ...): the accumulators cost registers and a spill
#ifdef BFVI_FUSED_COUNTERS
#define BFVI_DBG_T(var) const long long var = p.dbg ? clock64() : 0
#define BFVI_DBG_ADD(i, t0) do { if (p.dbg && blockIdx.x == 0 && lane == 0) dbg_acc[i] += clock64() - (t0); } while (0)
#define BFVI_DBG_ONLY(...) __VA_ARGS__
#else
#define BFVI_DBG_T(var) do { } while (0)
#define BFVI_DBG_ADD(i, t0) do { } while (0)
#define BFVI_DBG_ONLY(...)
#endif

// TMEM columns of the forward kernel: z (A operand), gate / nonlinear head accumulators, two 128-column pair buffers
// (hidden pre-activations, rewritten in place as A operands), the linear head accumulator.  The tail reuses buffer 0
// (the nonlinear head as A operand) and buffer 1 (std head accumulator).
constexpr uint32_t kFZ = 0, kFG = 64, kFNL = 128, kFB = 192, kFLIN = 448;

template <bool KEEP>
__global__ void __launch_bounds__(kThreads, 1) gtf_fwd_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ unsigned char fused_smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t z_full, units_done, tail_a, heads_full, heads_empty;
  __shared__ __align__(8) uint64_t d_full[2];
  __shared__ __align__(8) uint64_t a_full[2];
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = fused_smem_dyn + ((1024u - (smem_u32(fused_smem_dyn) & 1023u)) & 1023u);
  const int n_stages = p.n_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, U = H / kHU, U2 = 2 * U, n_blocks = U2 + 1;
  unsigned char* patches = smem + (size_t)n_stages * kBlockBytes;
  float* bias_s = reinterpret_cast<float*>(patches + 4 * kPatchBytes);
  const int64_t n_tiles = (p.R + kTileRows - 1) / kTileRows;
  const int my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;

  if (warp == kRowWarps) tmem_alloc(&tmem_base_s, 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&z_full, kRowWarps); mbar_init(&units_done, 1); mbar_init(&tail_a, kRowWarps);
    mbar_init(&heads_full, 1); mbar_init(&heads_empty, kRowWarps);
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&a_full[i], kRowWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * H + 4 * kZ + 8; i += blockDim.x) bias_s[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const float* inv_s = bias_s + 2 * H + 4 * kZ;        // 1 / scale of gate0, nonlin0, gate2, nonlin2, lin, std

  if (warp < kRowWarps) {
    // ================= row warps =================
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t tl = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32);   // my lane group, my column half
    unsigned char* pp = patches + q * kPatchBytes;    // the quadrant's pair patch
    const bool elected = hf == 0 && lane == 0;
    const bool lsu = BFVI_ABL(kStoreLsu) != 0;
    uint32_t par_d = 0, par_misc = 0;                 // phase bits: d_full[i] in bit i; misc flips once per tile
    int hp = 0, fs = 0;                               // next FP16 tile slot (4 x 4 KB) / fp32 rows slot (2 x 8 KB) of the patch
    BFVI_DBG_ONLY(long long dbg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};)
    BFVI_DBG_T(t_begin);
    float zreg[32];
    {
      const int64_t row = (int64_t)blockIdx.x * kTileRows + q * 32 + lane;
      load_row32(p.z + row * kZ + hf * 32, my_tiles > 0 && row < p.R, zreg);
    }
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)lt * gridDim.x;
      const int64_t row = tile * kTileRows + q * 32 + lane;
      const bool row_ok = row < p.R;
      // ---- z -> TMEM (A operand of the hidden layers and of the linear head)
      // (all MMAs of the previous tile are complete: this warp has passed heads_full of that tile)
      const int64_t row0 = tile * kTileRows + q * 32;
      const int n_valid = (int)(p.R - row0 < 32 ? (p.R - row0 < 0 ? 0 : p.R - row0) : 32);
      if (KEEP && !BFVI_ABL(kAblTile16)) {
        patch_switch(q, elected, lsu);                // the fp32 rows of the previous tile used the same shared memory
        store_tile_f16(pp, hp, q, lane, hf, elected, p.z16, row0, 2 * n_tiles, 1, 0, zreg, lsu, BFVI_ABL(kAblFence) != 0);
      }
      if (lt == 0) {                                  // (later tiles: written during the previous tile's tail)
        store_a_split(tl + kFZ, zreg);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&z_full);
      }
      if (lt + 1 < my_tiles) {                        // next tile's rows fly during this tile's units
        const int64_t nrow = (tile + gridDim.x) * kTileRows + q * 32 + lane;
        load_row32(p.z + nrow * kZ + hf * 32, nrow < p.R, zreg);
      }
      // ---- hidden pairs: D -> scale + bias, ReLU -> A (FP16 hi | lo), in place
#pragma unroll 1
      for (int c = 0; c < U; ++c) {
        const int b = c & 1;
        BFVI_DBG_T(t0);
        mbar_wait(&d_full[b], (par_d >> b) & 1u);
        par_d ^= 1u << b;
        tc_fence_after();
        BFVI_DBG_ADD(0, t0);
#pragma unroll
        for (int br = 0; br < 2; ++br) {              // gate unit c, nonlinear unit c
          BFVI_DBG_T(t1);
          float v[32];
          tmem_ld32(tl + kFB + b * 128 + br * 64, v);
          BFVI_DBG_ADD(1, t1);
          BFVI_DBG_T(t2);
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + br * H + c * kHU + hf * 32);
          const float sc = inv_s[br];
          if (!BFVI_ABL(kAblMath))
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 bb = b4[cc];
            // ReLU that keeps NaN like torch.relu (fmaxf would drop it): max.NaN is ONE instruction (FMNMX.NAN)
            // where `x < 0 ? 0 : x` is a compare and a select
            v[4 * cc] = relu_nan(fmaf(v[4 * cc], sc, bb.x));
            v[4 * cc + 1] = relu_nan(fmaf(v[4 * cc + 1], sc, bb.y));
            v[4 * cc + 2] = relu_nan(fmaf(v[4 * cc + 2], sc, bb.z));
            v[4 * cc + 3] = relu_nan(fmaf(v[4 * cc + 3], sc, bb.w));
          }
          if (KEEP) {
            const int u = br * U + c;
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) bits |= (v[j] > 0.f ? 1u : 0u) << j;
            // rows past the end contribute nothing to the weight gradients (their A operand is irrelevant: no output
            // row is stored for them).  Only the last tile has such rows: a warp-uniform test skips the 32 selects
            // elsewhere (and lets the FP16 conversions of the tile store and of the A operand share their work); the
            // selects themselves stay branch-free, the named barriers further down are reached convergently.
            if (n_valid < 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = row_ok ? v[j] : 0.f;
            }
            if (!BFVI_ABL(kAblTile16)) store_tile_f16(pp, hp, q, lane, hf, elected, p.h16, row0, 2 * n_tiles, U2, u, v, lsu, BFVI_ABL(kAblFence) != 0);
            // AFTER the tile store: its proxy fence is a MEMBAR.ALL.CTA in SASS, which waits for every global access this
            // thread has in flight — a store issued just before it cost the fence a round trip to L2 per unit
            if (!BFVI_ABL(kAblBits)) p.relu_bits[((tile * U2 + u) * 2 + hf) * kTileRows + q * 32 + lane] = row_ok ? bits : 0u;   // coalesced
          }
          BFVI_DBG_ADD(2, t2);
          BFVI_DBG_T(t3);
          if (BFVI_ABL(kAblMath)) tmem_st32(tl + kFB + b * 128 + br * 64, v); else
          store_a_split(tl + kFB + b * 128 + br * 64, v);
          BFVI_DBG_ADD(3, t3);
        }
        BFVI_DBG_T(t3b);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[b]);
        BFVI_DBG_ADD(3, t3b);
      }
      // ---- tail: the finished nonlinear head is the A operand of the std head
      BFVI_DBG_T(t4);
      {
        mbar_wait(&units_done, par_misc & 1u);
        tc_fence_after();
        BFVI_DBG_ADD(6, t4);
        BFVI_DBG_T(t8);
        float v[32];
        tmem_ld32(tl + kFNL, v);
        const float* bb = bias_s + 2 * H + kZ + hf * 32;
        const float sc = inv_s[3];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], sc, bb[j]);
        store_a_split(tl + kFB, v);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail_a);
        BFVI_DBG_ADD(8, t8);
        BFVI_DBG_T(t9);
        if (KEEP) patch_switch(q, elected, lsu);      // FP16 tile slots and fp32 row slots share the patch
        if (!BFVI_ABL(kAblRows)) store_rows_f32(pp, fs, q, lane, hf, elected, p.nl + row0 * kZ, n_valid, v, lsu);
        BFVI_DBG_ADD(9, t9);
      }
      // ---- heads out.  First hand the tensor pipe its next tile: z of the next tile goes to TMEM (the linear head has
      // consumed the old one) and the three accumulators move to registers, so that the issuer starts the next hidden
      // layers while this warp still scales, biases and stores the heads.
      {
        BFVI_DBG_T(t5);
        mbar_wait(&heads_full, par_misc & 1u);
        tc_fence_after();
        BFVI_DBG_ADD(7, t5);
        BFVI_DBG_T(t10);
        if (lt + 1 < my_tiles) {
          store_a_split(tl + kFZ, zreg);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&z_full);
        }
        const float* bb = bias_s + 2 * H + hf * 32;
        float vg[32], vl[32], va[32];
        tmem_ld32(tl + kFG, vg);
        tmem_ld32(tl + kFLIN, vl);
        tmem_ld32(tl + kFB + 128, va);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&heads_empty);     // every accumulator has been read
        BFVI_DBG_ADD(10, t10);
        BFVI_DBG_T(t11);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          vg[j] = fmaf(vg[j], inv_s[2], bb[j]);
          vl[j] = fmaf(vl[j], inv_s[4], bb[2 * kZ + j]);
          va[j] = fmaf(va[j], inv_s[5], bb[3 * kZ + j]);
        }
        if (!BFVI_ABL(kAblRows)) {
          store_rows_f32(pp, fs, q, lane, hf, elected, p.g + row0 * kZ, n_valid, vg, lsu);
          store_rows_f32(pp, fs, q, lane, hf, elected, p.lin + row0 * kZ, n_valid, vl, lsu);
          store_rows_f32(pp, fs, q, lane, hf, elected, p.as + row0 * kZ, n_valid, va, lsu);
        }
        BFVI_DBG_ADD(11, t11);
      }
      BFVI_DBG_ADD(4, t4);
      par_misc ^= 1u;
    }
    if (elected) bulk_wait_all();                     // results are in global memory before the kernel ends
    BFVI_DBG_ADD(5, t_begin);
    BFVI_DBG_ONLY(if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0)
      for (int i = 0; i < 12; ++i) p.dbg[i] = dbg_acc[i];)
  } else if (warp == kRowWarps) {
    // ================= MMA issuer =================
    // The WHOLE warp walks the schedule with warp-uniform values (TMEM base broadcast by a shuffle, ring addresses
    // from uniform loop counters) and one elected lane issues: under an `if (lane == 0)` around the loop the compiler cannot prove
    // the tcgen05 operands uniform and wraps every instruction in an ELECT / R2UR.BROADCAST waterfall loop (measured:
    // 81 cycles per MMA issued instead of the tensor pipe's 52-67).
    {
      BFVI_DBG_ONLY(long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};)
      const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
      const uint32_t idesc64 = umma_idesc_f16_k(kTileRows, 64), idesc128 = umma_idesc_f16_k(kTileRows, 128);
      const uint32_t ring = smem_u32(smem);
      uint32_t par_a = 0;
      uint32_t gblk = 0;                              // blocks consumed so far (ring position)
      const FastDiv ring_div(n_stages);
      auto stage_of = [&](uint32_t g) { return (int)ring_div.mod(g); };
      auto wait_block = [&](uint32_t g) { const uint32_t lap = ring_div.div(g); mbar_wait(&full_bar[g - lap * ring_div.n], lap & 1u); tc_fence_after(); };
      for (int lt = 0; lt < my_tiles; ++lt) {
        mbar_wait(&z_full, (uint32_t)(lt & 1));
        tc_fence_after();
        auto issue1 = [&](int c) {                    // hidden pre-activations of pair c: one N = 128 stream
          BFVI_DBG_T(tw);
          wait_block(gblk + 2 * c);
          BFVI_DBG_ADD(1, tw);
          BFVI_DBG_T(ti);
          if (elect_one()) {
            if (!BFVI_ABL(kAblMma1))
            mma_split(tbu, kFB + (c & 1) * 128, kFZ, ring + stage_of(gblk + 2 * c) * kBlockBytes, kTileBytes, true, idesc128);
            umma_commit(&d_full[c & 1]);
            umma_commit(&empty_bar[stage_of(gblk + 2 * c)]);
          }
          __syncwarp();
          BFVI_DBG_ADD(2, ti);
        };
        issue1(0);                                    // buffer 0: free since the std head's MMAs (in order) completed
        mbar_wait(&heads_empty, (uint32_t)((lt & 1) ^ 1));   // previous tile's accumulators (one sits in buffer 1) read out
        tc_fence_after();
        for (int c = 0; c < U; ++c) {
          const int b = c & 1;
          if (c + 1 < U) issue1(c + 1);               // overlaps the row warps' work on pair c
          BFVI_DBG_T(ta);
          mbar_wait(&a_full[b], (par_a >> b) & 1u);
          par_a ^= 1u << b;
          tc_fence_after();
          BFVI_DBG_ADD(0, ta);
          BFVI_DBG_T(tw);
          wait_block(gblk + 2 * c + 1);
          BFVI_DBG_ADD(1, tw);
          BFVI_DBG_T(ti);
          const uint32_t blk = ring + stage_of(gblk + 2 * c + 1) * kBlockBytes;
          if (elect_one()) {
            if (!BFVI_ABL(kAblMma2)) {
            mma_split(tbu, kFG, kFB + b * 128, blk, 8192, c == 0, idesc64);
            mma_split(tbu, kFNL, kFB + b * 128 + 64, blk + kTileBytes, 8192, c == 0, idesc64);
            }
            umma_commit(&empty_bar[stage_of(gblk + 2 * c + 1)]);
          }
          __syncwarp();
          BFVI_DBG_ADD(2, ti);
        }
        if (elect_one()) umma_commit(&units_done);
        __syncwarp();
        const uint32_t gt = gblk + (uint32_t)U2;
        wait_block(gt);
        const uint32_t blk = ring + stage_of(gt) * kBlockBytes;
        if (elect_one()) mma_split(tbu, kFLIN, kFZ, blk, 8192, true, idesc64);
        __syncwarp();
        mbar_wait(&tail_a, (uint32_t)(lt & 1));
        tc_fence_after();
        if (elect_one()) {
          mma_split(tbu, kFB + 128, kFB, blk + kTileBytes, 8192, true, idesc64);
          umma_commit(&empty_bar[stage_of(gt)]);
          umma_commit(&heads_full);
        }
        __syncwarp();
        gblk += n_blocks;
      }
      BFVI_DBG_ONLY(if (p.dbg && blockIdx.x == 0 && lane == 0)
        for (int i = 0; i < 8; ++i) p.dbg[16 + i] = dbg_acc[i];)
    }
    __syncwarp();
  } else {
    // ================= weight loader =================
    if (lane == 0) {
      const uint32_t total = (uint32_t)my_tiles * (uint32_t)n_blocks;
      int s = 0, blk = 0;                             // ring stage / block of the pack, advanced with the counter
      uint32_t lap = 0;
      for (uint32_t g = 0; g < total; ++g, s = (s + 1 == n_stages ? 0 : s + 1), blk = (blk + 1 == n_blocks ? 0 : blk + 1)) {
        if (g > 0 && s == 0) ++lap;
        if (g >= (uint32_t)n_stages) mbar_wait(&empty_bar[s], (lap - 1u) & 1u);
        if (BFVI_ABL(kAblRing) && g >= n_stages) { mbar_arrive(&full_bar[s]); continue; }
        mbar_expect_tx(&full_bar[s], kBlockBytes);
        bulk_g2s(smem + (size_t)s * kBlockBytes, p.pack + (size_t)blk * kBlockBytes, kBlockBytes, &full_bar[s]);
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
  __half* dh16;                  // masked hidden gradients, FP16 tiles [atom pair][row group][2 atoms]
  __half* dg16; __half* dnl16;   // FP16 tiles of d_g / d_nl [row group][1 atom]
  const unsigned* gmax;          // nullable: maxima [|d_g|, |d_nl| before the std path, |d_as|] of the launch's head gradients (float bits)
  const float* l1;               // 2 floats: the pack's column L1 norms (gtf_scale_kernel slots 6, 7)
  float* gscale;                 // out (nullable): the power-of-two scale s the FP16 tiles carry; wgrad16_kernel divides by it
  int64_t R;
  int H;
  int n_stages;
  int abl;                       // development: ablation mask, 0 in the product
};
constexpr uint32_t kBDG = 0, kBDNL = 64, kBDZ = 128, kBHB = 192;        // 4 hidden buffers + d_lin buffer (index 4)

__device__ __forceinline__ void mma8_tf32(uint32_t tb, uint32_t d_col, uint32_t a_col, uint32_t tile_addr, bool fresh,
                                          uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    umma_tf32_ts(tb + d_col, tb + a_col + k * 8, umma_desc_sw128(tile_addr + (k >> 2) * 8192 + (k & 3) * 32), idesc,
                 (fresh && k == 0) ? 0u : 1u);
}

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
  unsigned char* patches = smem + (size_t)n_stages * kBlockBytes;
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  // Power-of-two scale of this launch's gradients: head gradients (and with them the hidden gradients the MMAs produce,
  // the FP16 tiles of both, and dz) are multiplied by s so that the largest hidden gradient this launch can produce sits
  // at ~2^14; dz is divided by s on the way out, the weight-gradient kernel divides its accumulators.  Exact (a power of
  // two), and the TF32 operands have exponent range to spare.  Without it FP16 tiles of gradients below ~1e-4 went
  // subnormal: 3e-3 error at 2^-14 of O(1), 5e-2 at 2^-18 (tools/probe_f16_range.py).
  float gs = 1.f;
  if (p.gmax != nullptr) {
    const float mg = __uint_as_float(p.gmax[0]), mn = __uint_as_float(p.gmax[1]), ma = __uint_as_float(p.gmax[2]);
    const float m_head = fmaxf(mg, fmaf(ma, p.l1[1], mn));
    const float m = fmaxf(m_head, m_head * p.l1[0]);
    if (m > 0.f && m < 3.0e38f) {
      int e;
      frexpf(m, &e);                                  // m = f * 2^e, f in [0.5, 1)
      e = 14 - e;
      e = e < -60 ? -60 : (e > 60 ? 60 : e);
      gs = ldexpf(1.f, e);
    }
  }
  const float inv_gs = 1.f / gs;
  if (p.gscale != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.gscale[0] = gs;

  if (warp < kRowWarps) {
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t tl = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32);
    unsigned char* pp = patches + q * kPatchBytes;
    const bool elected = hf == 0 && lane == 0;
    const bool lsu = BFVI_ABL(kStoreLsu) != 0;
    uint32_t par_d = 0, par_misc = 0;
    int hp = 0, fs = 0;
    float rg[32], rn[32];                             // next tile's d_g / d_nl half rows (prefetched during the units)
    {
      const int64_t row = (int64_t)blockIdx.x * kTileRows + q * 32 + lane;
      const bool ok = my_tiles > 0 && row < p.R;
      load_row32(p.d_g + row * kZ + hf * 32, ok, rg);
      load_row32(p.d_nl + row * kZ + hf * 32, ok, rn);
    }
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)lt * gridDim.x;
      const int64_t row = tile * kTileRows + q * 32 + lane;
      const bool row_ok = row < p.R;
      // ---- head gradients -> TMEM (A operands), FP16 tiles for the weight-gradient GEMM
      const int64_t row0 = tile * kTileRows + q * 32;
      const int n_valid = (int)(p.R - row0 < 32 ? (p.R - row0 < 0 ? 0 : p.R - row0) : 32);
      {
        float v[32];                                  // d_lin rows: in flight while the FP16 tiles below are written
        load_row32(p.d_lin + row * kZ + hf * 32, row_ok, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) { rg[j] *= gs; rn[j] *= gs; v[j] *= gs; }
        patch_switch(q, elected, lsu);                // dz of the previous tile used the whole patch
        if (!BFVI_ABL(kAblTile16)) {
          store_tile_f16(pp, hp, q, lane, hf, elected, p.dg16, row0, 2 * n_tiles, 1, 0, rg, lsu, BFVI_ABL(kAblFence) != 0);
          store_tile_f16(pp, hp, q, lane, hf, elected, p.dnl16, row0, 2 * n_tiles, 1, 0, rn, lsu, BFVI_ABL(kAblFence) != 0);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { rg[j] = rn_tf32(rg[j]); rn[j] = rn_tf32(rn[j]); }
        tmem_st32(tl + kBDG, rg);
        tmem_st32(tl + kBDNL, rn);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rn_tf32(v[j]);
        tmem_st32(tl + kBHB + 4 * 64, v);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&in_full);
        if (lt + 1 < my_tiles) {
          const int64_t nrow = (tile + gridDim.x) * kTileRows + q * 32 + lane;
          load_row32(p.d_g + nrow * kZ + hf * 32, nrow < p.R, rg);
          load_row32(p.d_nl + nrow * kZ + hf * 32, nrow < p.R, rn);
        }
      }
      // ReLU bits of unit u + 1 are fetched while unit u is processed (a load issued at the top of its own iteration
      // was exposed whenever the hidden gradients were already waiting: ~4 000 cycles per tile)
      const uint32_t* bits_p = p.relu_bits + (tile * U2 * 2 + hf) * kTileRows + q * 32 + lane;
      uint32_t bits_next = BFVI_ABL(kAblBits) ? 0xffffffffu : __ldg(bits_p);
#pragma unroll 1
      for (int u = 0; u < U2; ++u) {
        const int hb = u & 3;
        const uint32_t bits = bits_next;
        // (issued here, a whole iteration ahead: after the tile store — behind its MEMBAR — measured 170 -> 196 us)
        if (u + 1 < U2 && !BFVI_ABL(kAblBits)) bits_next = __ldg(bits_p + (size_t)(u + 1) * 2 * kTileRows);
        mbar_wait(&d_full[hb], (par_d >> hb) & 1u);
        par_d ^= 1u << hb;
        tc_fence_after();
        float v[32];
        tmem_ld32(tl + kBHB + hb * 64, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.f;
        if (!BFVI_ABL(kAblTile16)) store_tile_f16(pp, hp, q, lane, hf, elected, p.dh16, row0, 2 * n_tiles, U2, u, v, lsu, BFVI_ABL(kAblFence) != 0);
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
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= inv_gs;
        patch_switch(q, elected, lsu);                // FP16 tile slots and fp32 row slots share the patch
        if (!BFVI_ABL(kAblRows)) store_rows_f32(pp, fs, q, lane, hf, elected, p.dz + row0 * kZ, n_valid, v, lsu);
      }
      par_misc ^= 1u;
    }
    if (elected) bulk_wait_all();
  } else if (warp == kRowWarps) {
    // whole warp walks the schedule, one elected lane issues (see gtf_fwd_kernel)
    {
      const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
      const uint32_t idesc = umma_idesc_tf32(kTileRows, 64);
      const uint32_t ring = smem_u32(smem);
      uint32_t par_a = 0;
      uint32_t gblk = 0;
      const FastDiv ring_div(n_stages);
      auto stage_of = [&](uint32_t g) { return (int)ring_div.mod(g); };
      auto wait_block = [&](uint32_t g) { const uint32_t lap = ring_div.div(g); mbar_wait(&full_bar[g - lap * ring_div.n], lap & 1u); tc_fence_after(); };
      for (int lt = 0; lt < my_tiles; ++lt) {
        mbar_wait(&in_full, (uint32_t)(lt & 1));
        tc_fence_after();
        // head block: dz = d_lin W_lin (starts the accumulator)
        mbar_wait(&dz_empty, (uint32_t)((lt & 1) ^ 1));
        tc_fence_after();
        wait_block(gblk);
        if (elect_one()) {
          mma8_tf32(tbu, kBDZ, kBHB + 4 * 64, ring + stage_of(gblk) * kBlockBytes + kTileBytes, true, idesc);
          umma_commit(&empty_bar[stage_of(gblk)]);
        }
        __syncwarp();
        auto issue1 = [&](int u) {                    // hidden gradients of unit u: d_head W2u
          wait_block(gblk + 1 + u);
          if (elect_one()) {
            if (!BFVI_ABL(kAblMma1))
            mma8_tf32(tbu, kBHB + (u & 3) * 64, u < U ? kBDG : kBDNL, ring + stage_of(gblk + 1 + u) * kBlockBytes, true, idesc);
            umma_commit(&d_full[u & 3]);
          }
          __syncwarp();
        };
        issue1(0);
        if (U2 > 1) issue1(1);
        for (int u = 0; u < U2; ++u) {
          const int hb = u & 3;
          if (u + 2 < U2) issue1(u + 2);
          mbar_wait(&a_full[hb], (par_a >> hb) & 1u);
          par_a ^= 1u << hb;
          tc_fence_after();
          if (elect_one()) {
            if (!BFVI_ABL(kAblMma2))
            mma8_tf32(tbu, kBDZ, kBHB + hb * 64, ring + stage_of(gblk + 1 + u) * kBlockBytes + kTileBytes, false, idesc);
            umma_commit(&empty_bar[stage_of(gblk + 1 + u)]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&dz_full);
        __syncwarp();
        gblk += n_blocks;
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      const uint32_t total = (uint32_t)my_tiles * (uint32_t)n_blocks;
      int s = 0, blk = 0;                             // ring stage / block of the pack, advanced with the counter
      uint32_t lap = 0;
      for (uint32_t g = 0; g < total; ++g, s = (s + 1 == n_stages ? 0 : s + 1), blk = (blk + 1 == n_blocks ? 0 : blk + 1)) {
        if (g > 0 && s == 0) ++lap;
        if (g >= (uint32_t)n_stages) mbar_wait(&empty_bar[s], (lap - 1u) & 1u);
        if (BFVI_ABL(kAblRing) && g >= n_stages) { mbar_arrive(&full_bar[s]); continue; }
        mbar_expect_tx(&full_bar[s], kBlockBytes);
        bulk_g2s(smem + (size_t)s * kBlockBytes, p.pack + (size_t)blk * kBlockBytes, kBlockBytes, &full_bar[s]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kRowWarps) tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------------------------------------------------
// weight gradients of the four H-wide layers from the FP16 operand tiles
// ---------------------------------------------------------------------------------------------------------
// out^T[h, zc] += sum_rows X[row, h] * Y[row, zc].  X: FP16 tiles [row group][n_atoms_x atoms], the problem uses atoms
// atom0 .. (H/64 of them; [atom pair][row group][2 atoms], n_groups = the launch's); Y: FP16 tiles [row group][1 atom].  out is either (H, Z) row-major (dW of a z -> hidden
// layer: direct) or (Z, H) row-major (dW of a hidden -> head layer: transposed add).  bias (nullable): (H) += column
// sums of X (the bias gradient of a z -> hidden layer whose X is the hidden gradient).
struct Wgrad16Problem {
  const __half* X; int n_atoms_x; int atom0;
  const __half* Y;
  float* out; int transposed;
  float* bias;
};
struct Wgrad16Params {
  Wgrad16Problem pr[4];
  int n_problems;
  int H;
  int64_t n_groups;            // row groups of 64 rows (R rounded up to 128 rows, tiles past R are zero)
  int groups_per_slice;        // K split: a work item contracts this many row groups
  int n_slices;
  int n_stages;
  int abl;                     // development: ablation mask, 0 in the product
  const float* gscale;         // nullable: the scale the gradient tiles carry (gtf_bwd_kernel); accumulators are divided by it
};
constexpr int kWgStageBytes = 3 * kAtomBytes;        // X atoms (2) + Y atom
constexpr int kWgThreads = 6 * 32;                   // 4 epilogue / column-sum warps, MMA warp, loader warp

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
  // Bias gradients ride on the MMA: a constant MN-major atom whose column 0 is all ones is appended to the Y operand
  // (N = 80 instead of 64: the B descriptor's leading-dimension offset points from the stage's Y atom to it), so
  // accumulator column 64 = sum over the rows of X = the column sums the z -> hidden bias gradients need.
  unsigned char* ones = smem + (size_t)n_stages * kWgStageBytes;
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
  if (warp == 4) tmem_alloc(&tmem_base_s, 128);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1); mbar_init(&acc_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kAtomBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0u;
  __syncthreads();
  if (threadIdx.x < kRowGroup) {                     // element (row k, column 0) of the atom: chunk 0 ^ (k % 8), first half
    const int k = threadIdx.x;
    *reinterpret_cast<__half*>(ones + k * 128 + ((k & 7) << 4)) = __float2half_rn(1.f);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;

  if (warp < 4) {
    // ===== epilogue: accumulator (128 hidden rows x 64 z columns [+ column sums]) -> red.global.add =====
    for (int li = 0; li < my_items; ++li) {
      int pi, mt; int64_t g0, g1;
      item_of(li, pi, mt, g0, g1);
      const Wgrad16Problem& pr = p.pr[pi];
      mbar_wait(&acc_full, (uint32_t)(li & 1));
      tc_fence_after();
      const int h = mt * 128 + warp * 32 + lane;
      float v0[32], v1[32], vb[32];
      tmem_ld32(tb + ((uint32_t)(warp * 32) << 16), v0);
      tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + 32u, v1);
      if (pr.bias != nullptr) tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + 64u, vb);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
      if (g1 > g0 && !BFVI_ABL(kAblRows)) {
        const float inv = p.gscale != nullptr ? 1.f / p.gscale[0] : 1.f;      // a power of two: exact
#pragma unroll
        for (int j = 0; j < 32; ++j) { v0[j] *= inv; v1[j] *= inv; }
        if (pr.bias != nullptr) vb[0] *= inv;
        if (pr.transposed) {                         // out (Z, H): lanes = consecutive h -> coalesced per column
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(pr.out + (size_t)j * p.H + h, v0[j]);
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(pr.out + (size_t)(32 + j) * p.H + h, v1[j]);
        } else {                                     // out (H, Z): a thread owns the 64 columns of its row
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(pr.out + (size_t)h * kZ + j, v0[j]);
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(pr.out + (size_t)h * kZ + 32 + j, v1[j]);
        }
        if (pr.bias != nullptr) atomicAdd(pr.bias + h, vb[0]);
      }
    }
  } else if (warp == 4) {
    {
      const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
      const uint32_t idesc64 = umma_idesc_f16_mn(128, 64), idesc80 = umma_idesc_f16_mn(128, 80);
      const uint32_t ring = smem_u32(smem), ones_addr = smem_u32(ones);
      int s = 0;                                      // ring stage and lap parity, advanced with the position
      uint32_t lap = 0;
      for (int li = 0; li < my_items; ++li) {
        int pi, mt; int64_t g0, g1;
        item_of(li, pi, mt, g0, g1);
        const bool with_bias = p.pr[pi].bias != nullptr;
        const uint32_t idesc = with_bias ? idesc80 : idesc64;
        mbar_wait(&acc_empty, (uint32_t)((li & 1) ^ 1));
        tc_fence_after();
        for (int64_t g = g0; g < g1; ++g) {
          mbar_wait(&full_bar[s], lap & 1u);
          tc_fence_after();
          const uint32_t xa = ring + s * kWgStageBytes, ya = xa + 2 * kAtomBytes;
          const uint32_t y_lbo = with_bias ? ones_addr - ya : (uint32_t)kAtomBytes;
          if (elect_one()) {
            if (!BFVI_ABL(kAblMma1))
#pragma unroll
            for (int k = 0; k < 4; ++k)               // 16 rows (two 8-row K groups = 2 KB) per instruction
              umma_f16_ss(tbu, umma_desc_mn128(xa + k * 2048, kAtomBytes, 1024), umma_desc_mn128(ya + k * 2048, y_lbo, 1024),
                          idesc, (g > g0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[s]);
          }
          __syncwarp();
          if (++s == n_stages) { s = 0; ++lap; }
        }
        if (elect_one()) umma_commit(&acc_full);
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      int s = 0;
      uint32_t lap = 0, gpos = 0;
      for (int li = 0; li < my_items; ++li) {
        int pi, mt; int64_t g0, g1;
        item_of(li, pi, mt, g0, g1);
        const Wgrad16Problem& pr = p.pr[pi];
        const unsigned char* xs = reinterpret_cast<const unsigned char*>(pr.X) +
                                  ((size_t)((pr.atom0 >> 1) + mt) * (size_t)p.n_groups + (size_t)g0) * 2 * kAtomBytes;
        const unsigned char* ys = reinterpret_cast<const unsigned char*>(pr.Y) + (size_t)g0 * kAtomBytes;
        for (int64_t g = g0; g < g1; ++g, ++gpos, xs += 2 * kAtomBytes, ys += kAtomBytes) {
          if (gpos >= (uint32_t)n_stages) mbar_wait(&empty_bar[s], (lap - 1u) & 1u);
          if (!(BFVI_ABL(kAblRing) && gpos >= (uint32_t)n_stages)) {
            unsigned char* dst = smem + (size_t)s * kWgStageBytes;
            mbar_expect_tx(&full_bar[s], kWgStageBytes);
            bulk_g2s(dst, xs, 2 * kAtomBytes, &full_bar[s]);                 // two adjacent 64-column atoms
            bulk_g2s(dst + 2 * kAtomBytes, ys, kAtomBytes, &full_bar[s]);
          } else {
            mbar_arrive(&full_bar[s]);
          }
          if (++s == n_stages) { s = 0; ++lap; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, 128);
}
#endif  // !BFVI_EMU

}  // namespace fused
}  // namespace bfvi
