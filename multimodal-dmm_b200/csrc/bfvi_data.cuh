// bfvi_data.cuh — the batch preparation either side of the BFVI step, on the device
// (SURVEY.md §8f-2): collation (pad_and_merge, len_to_mask) and the input corruptions the
// trainers apply to every batch (burst_delete / rand_delete / keep_segment / del_segment =
// func_delete with different index sets), datasets/multiseq.py:321-353, 405-448, called at
// trainer.py:235, 284-287.
//
// The reference walks the batch in a Python loop (one numpy draw and one indexed device write
// per sequence and modality).  Here a corruption is ONE streaming pass over the (T, B, D...)
// tensor — read 4 B, write 4 B per element, NaN where the row (t, b) is deleted — and the row
// predicate is either
//   * an explicit (T, B) flag tensor / per-sequence span built by the host from the SAME numpy
//     draws the reference makes (bit-exact drop-in for seeded runs), or
//   * drawn on the device from the Philox stream of bfvi_rng.cuh with integer-exact rules
//     (restated in oracle/multiseq_oracle.py), so no host work and no H2D at all.
#pragma once
#include "bfvi_rng.cuh"

namespace bfvi {

struct DeleteParams {
  const float* x;
  float* out;
  int T, B;
  int64_t D;                 // elements per (t, b) row
  const uint8_t* del_mask;   // (T, B) flags, or null
  const int32_t* lo;         // per-sequence span [lo, hi), or null
  const int32_t* hi;
  const int32_t* lengths;    // (B) or null = T
  int invert;                // 1: delete what lies OUTSIDE the span, within [0, length)
};

__device__ __forceinline__ bool row_deleted(const DeleteParams& p, int64_t row) {
  if (p.del_mask != nullptr) return p.del_mask[row] != 0;
  const int t = (int)(row / p.B), b = (int)(row - (int64_t)t * p.B);
  const bool inside = t >= p.lo[b] && t < p.hi[b];
  if (!p.invert) return inside;
  const int len = p.lengths != nullptr ? p.lengths[b] : p.T;
  return !inside && t < len;
}

// out = x with NaN rows where deleted (func_delete, datasets/multiseq.py:405-420).
template <bool VEC>
__global__ void __launch_bounds__(256) delete_rows_kernel(const DeleteParams p) {
  constexpr int W = VEC ? 4 : 1;
  const int64_t n = (int64_t)p.T * p.B * p.D;
  const float nan = __int_as_float(0x7fc00000);
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * W; i < n;
       i += (int64_t)gridDim.x * blockDim.x * W) {
    const bool del = row_deleted(p, i / p.D);                // VEC: D % 4 == 0, one row per vector
    if (VEC) {
      float4 v = *reinterpret_cast<const float4*>(p.x + i);
      if (del) v = make_float4(nan, nan, nan, nan);
      *reinterpret_cast<float4*>(p.out + i) = v;
    } else {
      p.out[i] = del ? nan : p.x[i];
    }
  }
}

// len_to_mask (datasets/multiseq.py:321-327), time first: mask[t, b] = t < lengths[b]
__global__ void __launch_bounds__(256)
len_to_mask_kernel(const int32_t* __restrict__ lengths, int T, int B, uint8_t* __restrict__ mask) {
  const int64_t n = (int64_t)T * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / B), b = (int)(i - (int64_t)t * B);
    mask[i] = t < lengths[b] ? 1 : 0;
  }
}

// pad_and_merge (datasets/multiseq.py:342-353): sequences packed back to back, `row_start[b]`
// = first packed row of sequence b (B + 1 entries) -> (T, B, D) padded with NaN.
template <bool VEC>
__global__ void __launch_bounds__(256)
pad_merge_kernel(const float* __restrict__ packed, const int64_t* __restrict__ row_start, int T, int B, int64_t D,
                 float* __restrict__ out) {
  constexpr int W = VEC ? 4 : 1;
  const int64_t n = (int64_t)T * B * D;
  const float nan = __int_as_float(0x7fc00000);
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * W; i < n;
       i += (int64_t)gridDim.x * blockDim.x * W) {
    const int64_t row = i / D, e = i - row * D;
    const int t = (int)(row / B), b = (int)(row - (int64_t)t * B);
    const int64_t r0 = row_start[b], len = row_start[b + 1] - r0;
    const bool on = t < len;
    const float* src = packed + (r0 + t) * D + e;
    if (VEC) {
      *reinterpret_cast<float4*>(out + i) = on ? *reinterpret_cast<const float4*>(src) : make_float4(nan, nan, nan, nan);
    } else {
      out[i] = on ? *src : nan;
    }
  }
}

// seq_decoll (datasets/multiseq.py:388-398): the inverse of pad_merge — de-pad and reorder.  Output
// sequence j = batch column src[j] (= the caller's `order`), its first lengths[src[j]] steps, packed back
// to back at row_start[j] (B + 1 entries, rows of D floats).  One streaming pass over the OUTPUT:
// read 4 B, write 4 B per kept element; the host then splits ONE D2H copy into per-sequence views.
template <bool VEC>
__global__ void __launch_bounds__(256)
unpad_kernel(const float* __restrict__ x, const int64_t* __restrict__ row_start, const int32_t* __restrict__ src,
             int B, int n_out, int64_t D, float* __restrict__ packed) {
  constexpr int W = VEC ? 4 : 1;
  const int64_t n = row_start[n_out] * D;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * W; i < n;
       i += (int64_t)gridDim.x * blockDim.x * W) {
    const int64_t row = i / D, e = i - row * D;
    int lo = 0, hi = n_out;                                   // sequence j with row_start[j] <= row < row_start[j+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (row_start[mid] <= row) lo = mid; else hi = mid;
    }
    const int64_t t = row - row_start[lo];
    const float* s = x + (t * B + src[lo]) * D + e;
    if (VEC) *reinterpret_cast<float4*>(packed + i) = *reinterpret_cast<const float4*>(s);
    else packed[i] = *s;
  }
}

// wide rows (images): one block per output row at a time — the row -> sequence search once per row instead
// of once per vector, then a straight 128-bit copy of the row
__global__ void __launch_bounds__(256)
unpad_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ row_start, const int32_t* __restrict__ src,
                  int B, int n_out, int64_t D, float* __restrict__ packed) {
  __shared__ int64_t s_src;
  const int64_t n_rows = row_start[n_out];
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    if (threadIdx.x == 0) {
      int lo = 0, hi = n_out;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row_start[mid] <= row) lo = mid; else hi = mid;
      }
      s_src = ((row - row_start[lo]) * B + src[lo]) * D;
    }
    __syncthreads();
    const float4* s4 = reinterpret_cast<const float4*>(x + s_src);
    float4* d4 = reinterpret_cast<float4*>(packed + row * D);
    for (int64_t i = threadIdx.x; i < D / 4; i += blockDim.x) d4[i] = s4[i];
    __syncthreads();
  }
}

// Per-sequence mean squared error of the evaluation metrics (spirals.py:105-111):
//   mse[t, b] = sum_m sum_d (recon_m[t, b, d] - target_m[t, b, d])^2, zeroed where the sequence mask is off,
//   out[b] = sum_t mse[t, b] / lengths[b].
// One 128-thread block per sequence: threads stride over the (t, d) elements of every modality, feature
// index fastest (coalesced rows, 128-bit loads when rows are 16-byte multiples), then one block reduction.
constexpr int kMaxMseMods = 8;
struct SeqMseParams {
  const float* recon[kMaxMseMods];
  const float* target[kMaxMseMods];
  int64_t D[kMaxMseMods];
  int n_mods, T, B;
  const uint8_t* mask;         // (T, B)
  const float* lengths;        // (B), the divisor (reference: FloatTensor(lengths))
  float* out;                  // (B)
  float* scratch;              // (B, n_split) partial sums when n_split > 1 (few long sequences: more blocks)
  int n_split;                 // blockIdx.y owns time steps [y * t_per, (y + 1) * t_per)
  int t_per;
};
__global__ void __launch_bounds__(128) seq_mse_kernel(const SeqMseParams p) {
  __shared__ float part[4];
  const int b = blockIdx.x;
  const int t_lo = blockIdx.y * p.t_per, t_hi = t_lo + p.t_per < p.T ? t_lo + p.t_per : p.T;
  const int nt = t_hi - t_lo;
  float acc = 0.f;
  for (int m = 0; m < p.n_mods; ++m) {
    const int64_t D = p.D[m];
    const float* __restrict__ r = p.recon[m];
    const float* __restrict__ x = p.target[m];
    const bool vec = D % 4 == 0 && ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
    if (vec) {
      const int64_t n4 = (int64_t)nt * (D / 4);
      for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        const int tl = (int)(i / (D / 4)), t = t_lo + tl;
        if (p.mask[(int64_t)t * p.B + b] == 0) continue;
        const int64_t o = ((int64_t)t * p.B + b) * D + (i - (int64_t)tl * (D / 4)) * 4;
        const float4 a = *reinterpret_cast<const float4*>(r + o), c = *reinterpret_cast<const float4*>(x + o);
        const float d0 = a.x - c.x, d1 = a.y - c.y, d2 = a.z - c.z, d3 = a.w - c.w;
        acc += fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
      }
    } else {
      const int64_t n = (int64_t)nt * D;
      for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int tl = (int)(i / D), t = t_lo + tl;
        if (p.mask[(int64_t)t * p.B + b] == 0) continue;
        const int64_t o = ((int64_t)t * p.B + b) * D + (i - (int64_t)tl * D);
        const float df = r[o] - x[o];
        acc = fmaf(df, df, acc);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float v = part[0] + part[1] + part[2] + part[3];
    if (p.n_split > 1) p.scratch[(int64_t)b * p.n_split + blockIdx.y] = v;
    else p.out[b] = v / p.lengths[b];
  }
}
// second stage of a split reduction: fixed summation order (deterministic)
__global__ void __launch_bounds__(256) seq_mse_finish_kernel(const float* __restrict__ scratch, int n_split,
                                                             const float* __restrict__ lengths, int B, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float v = 0.f;
  for (int s = 0; s < n_split; ++s) v += scratch[(int64_t)b * n_split + s];
  out[b] = v / lengths[b];
}

// SSIM of the evaluation metrics (utils.py:110-212, called at weizmann.py:133,141): separable Gaussian blur
// (valid padding) of X, Y, X^2, Y^2, XY, the SSIM / contrast-structure maps, and their mean over (C, H', W')
// per image — fused: one block per (32 x 32 output tile, channel, image) keeps the haloed X / Y tile and the
// row-blurred maps in shared memory; the reference materialises five blurred (N, 5C, H, W) tensors.
// Block partial sums go to scratch (image, channel, tile); ssim_finish_kernel adds them in a fixed order.
constexpr int kSsimTile = 32, kSsimMaxWin = 15, kSsimIn = kSsimTile + kSsimMaxWin - 1;   // 46
struct SsimParams {
  const float* x; const float* y;   // (N, C, H, W)
  int N, C, H, W, win;
  float w[kSsimMaxWin];
  float c1, c2;
  int tiles_x, tiles_y;
  float* scratch;                   // (N, C * tiles, 2)
};
__global__ void __launch_bounds__(256) ssim_kernel(const SsimParams p) {
  __shared__ float sx[kSsimIn][kSsimIn + 1], sy[kSsimIn][kSsimIn + 1];
  __shared__ float hb[5][kSsimIn][kSsimTile + 1];
  __shared__ float red[2][8];
  const int tile = blockIdx.x, c = blockIdx.y, n = blockIdx.z;
  const int oy0 = (tile / p.tiles_x) * kSsimTile, ox0 = (tile % p.tiles_x) * kSsimTile;
  const int Ho = p.H - p.win + 1, Wo = p.W - p.win + 1;
  const int in_n = kSsimTile + p.win - 1;
  const float* X = p.x + ((int64_t)n * p.C + c) * p.H * p.W;
  const float* Y = p.y + ((int64_t)n * p.C + c) * p.H * p.W;
  for (int i = threadIdx.x; i < in_n * in_n; i += blockDim.x) {
    const int r = i / in_n, q = i - r * in_n;
    const int gy = oy0 + r, gx = ox0 + q;
    const bool in = gy < p.H && gx < p.W;
    sx[r][q] = in ? X[(int64_t)gy * p.W + gx] : 0.f;
    sy[r][q] = in ? Y[(int64_t)gy * p.W + gx] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < in_n * kSsimTile; i += blockDim.x) {      // blur along W
    const int r = i / kSsimTile, q = i - r * kSsimTile;
    float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
    for (int k = 0; k < p.win; ++k) {
      const float a = sx[r][q + k], b = sy[r][q + k], w = p.w[k];
      m1 = fmaf(w, a, m1); m2 = fmaf(w, b, m2);
      xx = fmaf(w, a * a, xx); yy = fmaf(w, b * b, yy); xy = fmaf(w, a * b, xy);
    }
    hb[0][r][q] = m1; hb[1][r][q] = m2; hb[2][r][q] = xx; hb[3][r][q] = yy; hb[4][r][q] = xy;
  }
  __syncthreads();
  float ssim = 0.f, cs = 0.f;
  for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += blockDim.x) { // blur along H + the maps
    const int r = i / kSsimTile, q = i - r * kSsimTile;
    if (oy0 + r >= Ho || ox0 + q >= Wo) continue;
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < p.win; ++k) {
      const float w = p.w[k];
#pragma unroll
      for (int j = 0; j < 5; ++j) v[j] = fmaf(w, hb[j][r + k][q], v[j]);
    }
    const float mu1_sq = v[0] * v[0], mu2_sq = v[1] * v[1], mu12 = v[0] * v[1];
    const float s1 = v[2] - mu1_sq, s2 = v[3] - mu2_sq, s12 = v[4] - mu12;
    const float cs_v = (2.f * s12 + p.c2) / (s1 + s2 + p.c2);
    cs += cs_v;
    ssim += ((2.f * mu12 + p.c1) / (mu1_sq + mu2_sq + p.c1)) * cs_v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ssim += __shfl_xor_sync(0xffffffffu, ssim, o); cs += __shfl_xor_sync(0xffffffffu, cs, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ssim; red[1][threadIdx.x >> 5] = cs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int wv = 0; wv < 8; ++wv) { a += red[0][wv]; b += red[1][wv]; }
    float* o = p.scratch + (((int64_t)n * p.C + c) * (p.tiles_x * p.tiles_y) + tile) * 2;
    o[0] = a; o[1] = b;
  }
}
// The 11-tap window (the reference default) with register blocking: a thread produces FOUR adjacent outputs
// of a blur from 14 loaded values (3.1x fewer shared-memory loads per FMA: the generic kernel is bound by one
// LDS per FMA at 16 % of the FP32 peak).
__global__ void __launch_bounds__(256) ssim11_kernel(const SsimParams p) {
  constexpr int WIN = 11, IN = kSsimTile + WIN - 1, NV = WIN + 3;             // 42, 14
  __shared__ float sx[IN][kSsimIn + 1], sy[IN][kSsimIn + 1];
  __shared__ float hb[5][IN][kSsimTile + 1];
  __shared__ float red[2][8];
  const int tile = blockIdx.x, c = blockIdx.y, n = blockIdx.z;
  const int oy0 = (tile / p.tiles_x) * kSsimTile, ox0 = (tile % p.tiles_x) * kSsimTile;
  const int Ho = p.H - WIN + 1, Wo = p.W - WIN + 1;
  const float* X = p.x + ((int64_t)n * p.C + c) * p.H * p.W;
  const float* Y = p.y + ((int64_t)n * p.C + c) * p.H * p.W;
  float w[WIN];
#pragma unroll
  for (int k = 0; k < WIN; ++k) w[k] = p.w[k];
  for (int i = threadIdx.x; i < IN * IN; i += blockDim.x) {
    const int r = i / IN, q = i - r * IN;
    const int gy = oy0 + r, gx = ox0 + q;
    const bool in = gy < p.H && gx < p.W;
    sx[r][q] = in ? X[(int64_t)gy * p.W + gx] : 0.f;
    sy[r][q] = in ? Y[(int64_t)gy * p.W + gx] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IN * (kSsimTile / 4); i += blockDim.x) {      // blur along W, 4 outputs per thread
    const int r = i / (kSsimTile / 4), q0 = (i - r * (kSsimTile / 4)) * 4;
    float a[NV], b[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) { a[j] = sx[r][q0 + j]; b[j] = sy[r][q0 + j]; }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
        const float u = a[o + k], v = b[o + k];
        m1 = fmaf(w[k], u, m1); m2 = fmaf(w[k], v, m2);
        xx = fmaf(w[k], u * u, xx); yy = fmaf(w[k], v * v, yy); xy = fmaf(w[k], u * v, xy);
      }
      hb[0][r][q0 + o] = m1; hb[1][r][q0 + o] = m2; hb[2][r][q0 + o] = xx; hb[3][r][q0 + o] = yy; hb[4][r][q0 + o] = xy;
    }
  }
  __syncthreads();
  float ssim = 0.f, cs = 0.f;
  {                                                                            // blur along H: 4 rows x 1 column per thread
    const int q = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
    float acc[4][5];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int j = 0; j < 5; ++j) acc[o][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      float col[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) col[k] = hb[j][r0 + k][q];
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k = 0; k < WIN; ++k) acc[o][j] = fmaf(w[k], col[o + k], acc[o][j]);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (oy0 + r0 + o >= Ho || ox0 + q >= Wo) continue;
      const float* v = acc[o];
      const float mu1_sq = v[0] * v[0], mu2_sq = v[1] * v[1], mu12 = v[0] * v[1];
      const float s1 = v[2] - mu1_sq, s2 = v[3] - mu2_sq, s12 = v[4] - mu12;
      const float cs_v = (2.f * s12 + p.c2) / (s1 + s2 + p.c2);
      cs += cs_v;
      ssim += ((2.f * mu12 + p.c1) / (mu1_sq + mu2_sq + p.c1)) * cs_v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ssim += __shfl_xor_sync(0xffffffffu, ssim, o); cs += __shfl_xor_sync(0xffffffffu, cs, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ssim; red[1][threadIdx.x >> 5] = cs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int wv = 0; wv < 8; ++wv) { a += red[0][wv]; b += red[1][wv]; }
    float* o = p.scratch + (((int64_t)n * p.C + c) * (p.tiles_x * p.tiles_y) + tile) * 2;
    o[0] = a; o[1] = b;
  }
}
__global__ void __launch_bounds__(128) ssim_finish_kernel(const float* __restrict__ scratch, int N, int per_image,
                                                          float inv_count, float* __restrict__ ssim, float* __restrict__ cs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a = 0.f, b = 0.f;
  for (int i = 0; i < per_image; ++i) { a += scratch[((int64_t)n * per_image + i) * 2]; b += scratch[((int64_t)n * per_image + i) * 2 + 1]; }
  ssim[n] = a * inv_count;
  if (cs != nullptr) cs[n] = b * inv_count;
}

// Seeded device draws of the deleted rows of every sequence (one thread per sequence):
//   mode 0 (rand_delete, datasets/multiseq.py:422-426): exactly k = int(frac * length) of the
//     `length` steps, a uniformly random subset — selection sampling (Knuth 3.4.2 S): step t is
//     deleted iff r_t * (length - t) < (k - deleted so far) * 2^32 with r_t a 32-bit Philox word;
//   mode 1 (burst_delete, :428-434): t_start = (r * length) >> 32, span of int(frac * length)
//     steps clipped at length.
// Philox counter (b + b_offset, t / 4, mode, stream_id), key = seed; word t % 4.
__global__ void __launch_bounds__(128)
draw_deletions_kernel(const int32_t* __restrict__ lengths, int T, int B, double frac, int mode, uint64_t seed,
                      unsigned stream_id, unsigned b_offset, uint8_t* __restrict__ del_mask) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int len = lengths != nullptr ? lengths[b] : T;
  len = len < T ? len : T;
  const int k = (int)(frac * (double)len);                  // Python: int(del_frac * length)
  const uint2 key{(unsigned)seed, (unsigned)(seed >> 32)};
  if (mode == 1) {
    const uint4 r = philox4x32_10(make_uint4(b + b_offset, 0u, 1u, stream_id), key);
    const int t0 = (int)(((uint64_t)r.x * (uint64_t)(len > 0 ? len : 0)) >> 32);
    const int t1 = t0 + k < len ? t0 + k : len;
    for (int t = 0; t < T; ++t) del_mask[(int64_t)t * B + b] = (t >= t0 && t < t1) ? 1 : 0;
    return;
  }
  int need = k;
  uint4 r = make_uint4(0, 0, 0, 0);
  for (int t = 0; t < T; ++t) {
    if ((t & 3) == 0) r = philox4x32_10(make_uint4(b + b_offset, (unsigned)(t >> 2), 0u, stream_id), key);
    const unsigned w = (t & 3) == 0 ? r.x : (t & 3) == 1 ? r.y : (t & 3) == 2 ? r.z : r.w;
    bool del = false;
    if (t < len && need > 0) {
      del = (uint64_t)w * (uint64_t)(len - t) < ((uint64_t)need << 32);
      need -= del ? 1 : 0;
    }
    del_mask[(int64_t)t * B + b] = del ? 1 : 0;
  }
}

}  // namespace bfvi
