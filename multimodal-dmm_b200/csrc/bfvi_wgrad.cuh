// bfvi_wgrad.cuh — warp-level weight-gradient accumulation for the small-dim path.
//
// Every lane of a warp owns one "row" (a particle or a sequence) and has the row's
// layer inputs X and pre-activation gradients D in registers.  The weight gradient
// dW = sum_rows D (x) X is a rank-32 update per warp step.  Rows are staged in two
// column-major shared-memory panels (panel_at(col, row)) and each lane then owns a
// TD x TX register tile of dW, walks the 32 rows with 128-bit loads, and adds its
// tile into a per-warp accumulator G that mirrors the flat parameter block.  The
// (lane -> tile) assignment and the (tile element -> parameter index) table are
// built once per kernel in shared memory from a small block description.
#pragma once
#include "bfvi_platform.cuh"

namespace bfvi {

constexpr int kRS = 36;   // panel column stride in floats: 32 rows + 4 pad (keeps float4 alignment
                          // and spreads the tile loads of 8 lanes on 8 columns over 8 bank groups)

// element (col, row) of a panel (an XOR swizzle without padding was tried: it saves 11 %
// of the shared memory but costs two integer instructions per tile load)
__host__ __device__ __forceinline__ int panel_at(int col, int row) { return col * kRS + row; }

struct WgBlock {          // one Linear layer: dW (nd x (nx-1)) and db (nd)
  int d0, nd;             // D-panel columns (pre-activation gradients)
  int x0, nx;             // X-panel columns; column x0 holds 1.0 (bias), then the inputs
  int w_off, b_off;       // offsets of weight / bias inside the accumulator G
};
struct WgSpec {
  int n_blocks;
  WgBlock blk[6];
};

template <int TD, int TX>
__host__ __device__ inline int wg_num_tasks(const WgSpec& s) {
  int n = 0;
  for (int i = 0; i < s.n_blocks; ++i)
    n += ((s.blk[i].nd + TD - 1) / TD) * ((s.blk[i].nx + TX - 1) / TX);
  return n;
}
template <int TD, int TX>
__host__ __device__ inline int wg_rounds(const WgSpec& s) { return (wg_num_tasks<TD, TX>(s) + 31) / 32; }
// shared-memory ints needed for the tables
template <int TD, int TX>
__host__ __device__ inline int wg_table_ints(const WgSpec& s) {
  return wg_rounds<TD, TX>(s) * 32 * (4 + TD * TX);
}

// tasks: [rounds*32] x {dcol0, ndc, xcol0, nxc};  oidx: [rounds][TD*TX][32].
// n_out = size of the parameter block; indices n_out + lane are per-lane dump slots.
template <int TD, int TX>
__device__ inline void wg_build_tables(const WgSpec& s, int* tasks, int* oidx, int rounds, int n_out) {
  for (int e = threadIdx.x; e < rounds * 32; e += blockDim.x) {
    const int lane = e & 31, round = e >> 5;
    int blk = -1, rem = e, n_xt = 1;
    for (int i = 0; i < s.n_blocks; ++i) {
      const int xt = (s.blk[i].nx + TX - 1) / TX;
      const int cnt = ((s.blk[i].nd + TD - 1) / TD) * xt;
      if (rem < cnt) { blk = i; n_xt = xt; break; }
      rem -= cnt;
    }
    int dcol0 = 0, ndc = 1, xcol0 = 0, nxc = 1, dt = 0, xt = 0;
    if (blk >= 0) {
      dt = rem / n_xt; xt = rem % n_xt;
      dcol0 = s.blk[blk].d0 + dt * TD;
      ndc = min(TD, s.blk[blk].nd - dt * TD);
      xcol0 = s.blk[blk].x0 + xt * TX;
      nxc = min(TX, s.blk[blk].nx - xt * TX);
    }
    tasks[e * 4 + 0] = dcol0; tasks[e * 4 + 1] = ndc;
    tasks[e * 4 + 2] = xcol0; tasks[e * 4 + 3] = nxc;
    for (int k = 0; k < TD * TX; ++k) {
      const int i = k / TX, j = k % TX;
      int idx = n_out + lane;
      if (blk >= 0 && i < ndc && j < nxc) {
        const int a = dt * TD + i, c = xt * TX + j;
        idx = (c == 0) ? s.blk[blk].b_off + a
                       : s.blk[blk].w_off + a * (s.blk[blk].nx - 1) + (c - 1);
      }
      oidx[(round * TD * TX + k) * 32 + lane] = idx;
    }
  }
}

// G[...] += sum over the warp's 32 staged rows.  Dp / Xp are this warp's panels.
template <int TD, int TX>
__device__ __forceinline__ void wg_accumulate(const float* __restrict__ Dp, const float* __restrict__ Xp,
                                              const int* __restrict__ tasks, const int* __restrict__ oidx,
                                              int rounds, float* __restrict__ G, int lane) {
  for (int r = 0; r < rounds; ++r) {
    const int4 t = reinterpret_cast<const int4*>(tasks)[r * 32 + lane];
    float acc[TD][TX];
#pragma unroll
    for (int i = 0; i < TD; ++i)
#pragma unroll
      for (int j = 0; j < TX; ++j) acc[i][j] = 0.f;
    const float4* dcol[TD];
    const float4* xcol[TX];
  #pragma unroll
    for (int i = 0; i < TD; ++i) {
      const int c = t.x + min(i, t.y - 1);
      dcol[i] = reinterpret_cast<const float4*>(Dp + c * kRS);
    }
#pragma unroll
    for (int j = 0; j < TX; ++j) {
      const int c = t.z + min(j, t.w - 1);
      xcol[j] = reinterpret_cast<const float4*>(Xp + c * kRS);
    }
#pragma unroll 2
    for (int q = 0; q < 8; ++q) {
      float4 dv[TD], xv[TX];
#pragma unroll
      for (int i = 0; i < TD; ++i) dv[i] = dcol[i][q];
#pragma unroll
      for (int j = 0; j < TX; ++j) xv[j] = xcol[j][q];
#pragma unroll
      for (int i = 0; i < TD; ++i)
#pragma unroll
        for (int j = 0; j < TX; ++j) {
          acc[i][j] = fmaf(dv[i].x, xv[j].x, acc[i][j]);
          acc[i][j] = fmaf(dv[i].y, xv[j].y, acc[i][j]);
          acc[i][j] = fmaf(dv[i].z, xv[j].z, acc[i][j]);
          acc[i][j] = fmaf(dv[i].w, xv[j].w, acc[i][j]);
        }
    }
    const int* oi = oidx + (size_t)r * TD * TX * 32 + lane;
#pragma unroll
    for (int i = 0; i < TD; ++i)
#pragma unroll
      for (int j = 0; j < TX; ++j) G[oi[(i * TX + j) * 32]] += acc[i][j];
  }
}

// Register-tile variant for kernels whose task list fits ONE round (<= 32 tiles):
// the lane keeps its TD x TX tile in registers across several staged row slices
// (wg_tile_fma) and folds it into the per-warp accumulator G once (wg_tile_flush).
// Each accumulator is a float2 {even rows, odd rows}: the products of one 128-bit
// panel load pair up as two packed FFMA2 (sm_100 fma.rn.f32x2) instead of four FFMA,
// halving the issue slots of the rank-32 update.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#ifdef BFVI_EMU
  return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#else
  return __ffma2_rn(a, b, c);
#endif
}
template <int TD, int TX>
__device__ __forceinline__ void wg_tile_fma(float2 (&acc)[TD][TX], const float* __restrict__ Dp,
                                            const float* __restrict__ Xp, const int4 t) {
  const float4* dcol[TD];
  const float4* xcol[TX];
#pragma unroll
  for (int i = 0; i < TD; ++i) dcol[i] = reinterpret_cast<const float4*>(Dp + (t.x + min(i, t.y - 1)) * kRS);
#pragma unroll
  for (int j = 0; j < TX; ++j) xcol[j] = reinterpret_cast<const float4*>(Xp + (t.z + min(j, t.w - 1)) * kRS);
#pragma unroll 2
  for (int q = 0; q < 8; ++q) {
    float4 dv[TD], xv[TX];
#pragma unroll
    for (int i = 0; i < TD; ++i) dv[i] = dcol[i][q];
#pragma unroll
    for (int j = 0; j < TX; ++j) xv[j] = xcol[j][q];
#pragma unroll
    for (int i = 0; i < TD; ++i)
#pragma unroll
      for (int j = 0; j < TX; ++j) {
        acc[i][j] = ffma2(float2{dv[i].x, dv[i].y}, float2{xv[j].x, xv[j].y}, acc[i][j]);
        acc[i][j] = ffma2(float2{dv[i].z, dv[i].w}, float2{xv[j].z, xv[j].w}, acc[i][j]);
      }
  }
}
template <int TD, int TX>
__device__ __forceinline__ void wg_tile_flush(float2 (&acc)[TD][TX], const int* __restrict__ oidx,
                                              float* __restrict__ G, int lane) {
  const int* oi = oidx + lane;
#pragma unroll
  for (int i = 0; i < TD; ++i)
#pragma unroll
    for (int j = 0; j < TX; ++j) {
      G[oi[(i * TX + j) * 32]] += acc[i][j].x + acc[i][j].y;
      acc[i][j] = float2{0.f, 0.f};
    }
}

// Block-level flush: sum the per-warp accumulators and add them to the global
// gradient block with one atomic per parameter per CTA.
__device__ inline void wg_flush(const float* __restrict__ G_all, int g_stride, int n_warps, int n_out,
                                float* __restrict__ grads) {
  __syncthreads();
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < n_warps; ++w) v += G_all[(size_t)w * g_stride + i];
    if (v != 0.f) atomicAdd(grads + i, v);
  }
}

}  // namespace bfvi
