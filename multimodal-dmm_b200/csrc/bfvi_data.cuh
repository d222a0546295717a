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
