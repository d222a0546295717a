// bfvi_conv.cuh — the image modules of the Weizmann / vidTIMIT models (models/common.py:70-175: Conv = Conv2d ->
// BatchNorm2d -> ReLU, Deconv = ConvTranspose2d -> BatchNorm2d -> ReLU, the decoder's final sigmoid) as FP32 kernels,
// forward and backward, NCHW like the reference's tensors.
//
// One geometry describes both layer kinds: a SMALL map (conv output / deconv input, Cs x Hs x Ws) and a BIG map (conv
// input / deconv output, Cb x Hb x Wb) tied by  hb = hs * stride - padding + kh,  and ONE weight layout
// w[Cs][Cb][k][k] — which is nn.Conv2d's (out, in, k, k) and nn.ConvTranspose2d's (in, out, k, k).  Three kernels
// serve every pass:
//   conv_gather_kernel   small <- big   (Conv2d forward,          ConvTranspose2d input gradient)
//   conv_scatter_kernel  big   <- small (ConvTranspose2d forward, Conv2d input gradient), written as a gather over
//                                       the taps that land on the output pixel (no atomics)
//   conv_wgrad_kernel    dw[cs][cb][kh][kw] += sum over images and small pixels (both layers' weight gradients)
// These are direct convolutions on the FP32 pipe: at the reference's sizes (3..64 channels, 64x64 images) one step's
// convolutions are ~30 GFLOP, well under the temporal core's share, and FP32 keeps the reference's (non-TF32) numerics.
// BatchNorm2d in training mode needs whole-batch statistics: chan_reduce_kernel (per-channel partial sums in double,
// deterministic two-stage) + a finish kernel (mean / rstd, running statistics updated like torch: biased variance to
// normalise, unbiased into running_var) + an apply kernel fused with the ReLU; the backward mirrors it.
#pragma once
#include "bfvi_platform.cuh"

namespace bfvi {
namespace conv {

struct Geom {
  int N;
  int Cs, Hs, Ws;
  int Cb, Hb, Wb;
  int k, s, p;
};

constexpr int kThreads = 128;       // pixels per CTA of the gather / scatter kernels
constexpr int kWgradThreads = 256;
constexpr int kMaxSplit = 64;       // partial sums per channel of the reductions
constexpr int kMaxSmemFloats = 8192;  // 32 KB weight stage

__device__ __forceinline__ float act_apply(float y, int act) {
  return act == 2 ? 1.f / (1.f + expf(-y)) : y;
}

// small[n, cs, oh, ow] = bias[cs] + sum_{cb, kh, kw} big[n, cb, oh*s - p + kh, ow*s - p + kw] * w[cs][cb][kh][kw]
// grid (ceil(N*Hs*Ws / 128), ceil(Cs / CT)); a thread owns one output pixel and CT output channels; the weights of
// its channel tile are staged in shared memory, cb_chunk input channels at a time, as [cb][tap][CT] so that one
// input value meets CT weights through 128-bit broadcast reads.
template <int KT, int CT>
__global__ void __launch_bounds__(kThreads) conv_gather_kernel(Geom g, const float* __restrict__ big,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               float* __restrict__ small, int cb_chunk, int act) {
  BFVI_DYN_SMEM(float, wsm);
  const int k = KT ? KT : g.k;
  const int kk = k * k;
  const int cs0 = (int)blockIdx.y * CT;
  const long long total = (long long)g.N * g.Hs * g.Ws;
  const long long pix = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool live = pix < total;
  int n = 0, oh = 0, ow = 0;
  if (live) {
    ow = (int)(pix % g.Ws);
    const long long t = pix / g.Ws;
    oh = (int)(t % g.Hs);
    n = (int)(t / g.Hs);
  }
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = (bias != nullptr && cs0 + c < g.Cs) ? bias[cs0 + c] : 0.f;
  const int h0 = oh * g.s - g.p, w0 = ow * g.s - g.p;
  const float* bimg = big + (size_t)n * g.Cb * g.Hb * g.Wb;
  for (int c0 = 0; c0 < g.Cb; c0 += cb_chunk) {
    const int cn = g.Cb - c0 < cb_chunk ? g.Cb - c0 : cb_chunk;
    __syncthreads();
    for (int i = (int)threadIdx.x; i < cn * kk * CT; i += kThreads) {
      const int c = i % CT, r = i / CT;
      const int tap = r % kk, cb = r / kk;
      wsm[i] = cs0 + c < g.Cs ? w[((size_t)(cs0 + c) * g.Cb + c0 + cb) * kk + tap] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int cb = 0; cb < cn; ++cb) {
      const float* bch = bimg + (size_t)(c0 + cb) * g.Hb * g.Wb;
      const float* wcb = wsm + (size_t)cb * kk * CT;
#pragma unroll
      for (int kh = 0; kh < k; ++kh) {
        const int h = h0 + kh;
        if ((unsigned)h >= (unsigned)g.Hb) continue;
#pragma unroll
        for (int kw = 0; kw < k; ++kw) {
          const int x = w0 + kw;
          if ((unsigned)x >= (unsigned)g.Wb) continue;
          const float v = bch[(size_t)h * g.Wb + x];
          const float4* wp = reinterpret_cast<const float4*>(wcb + (kh * k + kw) * CT);
#pragma unroll
          for (int c4 = 0; c4 < CT / 4; ++c4) {
            const float4 q = wp[c4];
            acc[4 * c4 + 0] = fmaf(v, q.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(v, q.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(v, q.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(v, q.w, acc[4 * c4 + 3]);
          }
        }
      }
    }
  }
  if (!live) return;
  const size_t plane = (size_t)g.Hs * g.Ws;
  float* out = small + ((size_t)n * g.Cs + cs0) * plane + (size_t)oh * g.Ws + ow;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (cs0 + c < g.Cs) out[c * plane] = act_apply(acc[c], act);
}

// big[n, cb, H, W] = bias[cb] + sum_{cs, kh, kw : (H + p - kh) = s*h, (W + p - kw) = s*w} small[n, cs, h, w] * w[cs][cb][kh][kw]
// grid (ceil(N*Hb*Wb / 128), ceil(Cb / CT)); weights staged as [cs][tap][CT], cs_chunk small channels at a time.
// ST = compile-time stride (0: run time).
template <int KT, int ST, int CT>
__global__ void __launch_bounds__(kThreads) conv_scatter_kernel(Geom g, const float* __restrict__ small,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ big, int cs_chunk, int act) {
  BFVI_DYN_SMEM(float, wsm);
  const int k = KT ? KT : g.k;
  const int s = ST ? ST : g.s;
  const int kk = k * k;
  const int cb0 = (int)blockIdx.y * CT;
  const long long total = (long long)g.N * g.Hb * g.Wb;
  const long long pix = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool live = pix < total;
  int n = 0, H = 0, W = 0;
  if (live) {
    W = (int)(pix % g.Wb);
    const long long t = pix / g.Wb;
    H = (int)(t % g.Hb);
    n = (int)(t / g.Hb);
  }
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = (bias != nullptr && cb0 + c < g.Cb) ? bias[cb0 + c] : 0.f;
  const float* simg = small + (size_t)n * g.Cs * g.Hs * g.Ws;
  // the taps that land on this pixel: kh = kh0 + a*s with (H + p - kh) = s*h, 0 <= h < Hs (and the same along W);
  // at k4 s2 that is 2 x 2 of the 16 taps, found once per thread, not once per channel
  constexpr int TAPS = (KT && ST) ? (KT + ST - 1) / ST : 7;
  const int kh0 = (H + g.p) % s, kw0 = (W + g.p) % s;
  int hh[TAPS], ww[TAPS];
  bool hv[TAPS], wv[TAPS];
#pragma unroll
  for (int a = 0; a < TAPS; ++a) {
    const int kh = kh0 + a * s, th = H + g.p - kh;
    hh[a] = th / s;
    hv[a] = kh < k && th >= 0 && hh[a] < g.Hs;
    const int kw = kw0 + a * s, tw = W + g.p - kw;
    ww[a] = tw / s;
    wv[a] = kw < k && tw >= 0 && ww[a] < g.Ws;
  }
  for (int c0 = 0; c0 < g.Cs; c0 += cs_chunk) {
    const int cn = g.Cs - c0 < cs_chunk ? g.Cs - c0 : cs_chunk;
    __syncthreads();
    for (int i = (int)threadIdx.x; i < cn * kk * CT; i += kThreads) {
      const int c = i % CT, r = i / CT;
      const int tap = r % kk, cs = r / kk;
      wsm[i] = cb0 + c < g.Cb ? w[((size_t)(c0 + cs) * g.Cb + cb0 + c) * kk + tap] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int cs = 0; cs < cn; ++cs) {
      const float* sch = simg + (size_t)(c0 + cs) * g.Hs * g.Ws;
      const float* wcs = wsm + (size_t)cs * kk * CT;
#pragma unroll
      for (int a = 0; a < TAPS; ++a) {
        if (!hv[a]) continue;
#pragma unroll
        for (int b = 0; b < TAPS; ++b) {
          if (!wv[b]) continue;
          const float v = sch[(size_t)hh[a] * g.Ws + ww[b]];
          const float4* wp = reinterpret_cast<const float4*>(wcs + ((kh0 + a * s) * k + kw0 + b * s) * CT);
#pragma unroll
          for (int c4 = 0; c4 < CT / 4; ++c4) {
            const float4 q = wp[c4];
            acc[4 * c4 + 0] = fmaf(v, q.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(v, q.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(v, q.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(v, q.w, acc[4 * c4 + 3]);
          }
        }
      }
    }
  }
  if (!live) return;
  const size_t plane = (size_t)g.Hb * g.Wb;
  float* out = big + ((size_t)n * g.Cb + cb0) * plane + (size_t)H * g.Wb + W;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (cb0 + c < g.Cb) out[c * plane] = act_apply(acc[c], act);
}

// dw[cs][cb][kh][kw] += sum_{n, h, w} small[n, cs, h, w] * big[n, cb, h*s - p + kh, w*s - p + kw]
// grid (Cb, ceil(Cs / CSR), splits): a CTA owns one big channel, CSR small channels and a contiguous range of the
// N*Hs*Ws small pixels; a thread keeps CSR x k*k partial sums in registers (one big value meets CSR small values),
// warps reduce by shuffles, the CTA through shared memory, CTAs by atomicAdd (the gradient buffers accumulate).
template <int KT, int CSR>
__global__ void __launch_bounds__(kWgradThreads) conv_wgrad_kernel(Geom g, const float* __restrict__ small,
                                                                   const float* __restrict__ big,
                                                                   float* __restrict__ dw, long long pix_per_cta) {
  constexpr int KK = KT ? KT * KT : 49;
  const int k = KT ? KT : g.k;
  const int kk = k * k;
  const int cb = (int)blockIdx.x, cs0 = (int)blockIdx.y * CSR;
  const long long total = (long long)g.N * g.Hs * g.Ws;
  const long long lo = (long long)blockIdx.z * pix_per_cta;
  const long long hi = lo + pix_per_cta < total ? lo + pix_per_cta : total;
  float acc[CSR][KK];
#pragma unroll
  for (int r = 0; r < CSR; ++r)
#pragma unroll
    for (int t = 0; t < KK; ++t) acc[r][t] = 0.f;
  const size_t splane = (size_t)g.Hs * g.Ws, bplane = (size_t)g.Hb * g.Wb;
  // (n, oh, ow) of this thread's pixel, advanced by 256 pixels per iteration in mixed radix (no division in the loop)
  const int step_w = kWgradThreads % g.Ws, step_h = (kWgradThreads / g.Ws) % g.Hs, step_n = kWgradThreads / (g.Ws * g.Hs);
  int ow, oh, n;
  {
    const long long first = lo + threadIdx.x;
    ow = (int)(first % g.Ws);
    const long long t = first / g.Ws;
    oh = (int)(t % g.Hs);
    n = (int)(t / g.Hs);
  }
  for (long long pix = lo + threadIdx.x; pix < hi; pix += kWgradThreads) {
    float sv[CSR];
#pragma unroll
    for (int r = 0; r < CSR; ++r)
      sv[r] = cs0 + r < g.Cs ? small[((size_t)n * g.Cs + cs0 + r) * splane + (size_t)oh * g.Ws + ow] : 0.f;
    const float* bch = big + ((size_t)n * g.Cb + cb) * bplane;
    const int h0 = oh * g.s - g.p, w0 = ow * g.s - g.p;
#pragma unroll
    for (int kh = 0; kh < (KT ? KT : 7); ++kh) {
      if (kh >= k) break;
      const int h = h0 + kh;
      if ((unsigned)h >= (unsigned)g.Hb) continue;
#pragma unroll
      for (int kw = 0; kw < (KT ? KT : 7); ++kw) {
        if (kw >= k) break;
        const int x = w0 + kw;
        if ((unsigned)x >= (unsigned)g.Wb) continue;
        const float v = bch[(size_t)h * g.Wb + x];
#pragma unroll
        for (int r = 0; r < CSR; ++r) acc[r][kh * (KT ? KT : 7) + kw] = fmaf(sv[r], v, acc[r][kh * (KT ? KT : 7) + kw]);
      }
    }
    ow += step_w;
    if (ow >= g.Ws) { ow -= g.Ws; ++oh; }
    oh += step_h;
    if (oh >= g.Hs) { oh -= g.Hs; ++n; }
    n += step_n;
  }
  __shared__ float red[kWgradThreads / 32][CSR * KK];
  const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < CSR; ++r)
#pragma unroll
    for (int t = 0; t < KK; ++t) {
      float v = acc[r][t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][r * KK + t] = v;
    }
  __syncthreads();
  constexpr int KS = KT ? KT : 7;       // row pitch of the accumulator tile
  for (int i = (int)threadIdx.x; i < CSR * KK; i += kWgradThreads) {
    const int r = i / KK, t = i % KK;
    const int kh = t / KS, kw = t % KS;
    if (cs0 + r >= g.Cs || kh >= k || kw >= k) continue;
    float v = 0.f;
#pragma unroll
    for (int wi = 0; wi < kWgradThreads / 32; ++wi) v += red[wi][i];
    atomicAdd(dw + ((size_t)(cs0 + r) * g.Cb + cb) * kk + kh * k + kw, v);
  }
}

// Per-channel sums over the N x HW elements of an NCHW tensor, partials in double: grid (C, nsplit), CTA (c, z)
// covers the linear range [z * per, (z+1) * per) of the channel's N*HW elements and writes partial[(c*nsplit+z)*2 ..].
//   mode 0: sum a, sum a^2                      (BatchNorm statistics; bias gradients use the first)
//   mode 1: g = relu ? b * (y > 0) : b; xhat = (a - mean) * rstd:  sum g, sum g * xhat     (BatchNorm backward; a = x, b = dy)
struct ReduceParams {
  const float* a; const float* b; const float* y; const float* mean_rstd;   // mean_rstd: [C][2]
  int N, C; long long HW; int mode, relu; long long per;
  double* partial;
};
__global__ void __launch_bounds__(256) chan_reduce_kernel(ReduceParams p) {
  const int c = (int)blockIdx.x, z = (int)blockIdx.y, nsplit = (int)gridDim.y;
  const long long total = (long long)p.N * p.HW;
  const long long lo = (long long)z * p.per, hi = lo + p.per < total ? lo + p.per : total;
  double s0 = 0.0, s1 = 0.0;
  float mean = 0.f, rstd = 0.f;
  if (p.mode == 1) { mean = p.mean_rstd[2 * c]; rstd = p.mean_rstd[2 * c + 1]; }
  // (image, offset in the plane) advanced by 256 elements per iteration without a division
  long long n = (lo + threadIdx.x) / p.HW, q = (lo + threadIdx.x) - n * p.HW;
  const long long step_n = 256 / p.HW, step_q = 256 - step_n * p.HW;
  for (long long i = lo + threadIdx.x; i < hi; i += 256, n += step_n, q += step_q) {
    if (q >= p.HW) { q -= p.HW; ++n; }
    const size_t at = ((size_t)n * p.C + c) * (size_t)p.HW + (size_t)q;
    if (p.mode == 0) {
      const float v = p.a[at];
      s0 += (double)v;
      s1 += (double)v * (double)v;
    } else {
      float gq = p.b[at];
      if (p.relu && !(p.y[at] > 0.f)) gq = 0.f;
      const float xh = (p.a[at] - mean) * rstd;
      s0 += (double)gq;
      s1 += (double)gq * (double)xh;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  __shared__ double red[8][2];
  const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
  if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int wi = 0; wi < 8; ++wi) { t0 += red[wi][0]; t1 += red[wi][1]; }
    p.partial[((size_t)c * nsplit + z) * 2 + 0] = t0;
    p.partial[((size_t)c * nsplit + z) * 2 + 1] = t1;
  }
}

// BatchNorm2d statistics (torch.nn.functional.batch_norm, training): mean, biased variance -> rstd; running_mean /
// running_var (unbiased) blended with `momentum`.  Evaluation mode: statistics from the running buffers.
__global__ void bn_stats_finish_kernel(const double* __restrict__ partial, int nsplit, int C, double count, float eps,
                                       float momentum, int training, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, float* __restrict__ mean_rstd) {
  const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= C) return;
  if (!training) {
    mean_rstd[2 * c] = running_mean[c];
    mean_rstd[2 * c + 1] = 1.f / sqrtf(running_var[c] + eps);
    return;
  }
  double s0 = 0.0, s1 = 0.0;
  for (int z = 0; z < nsplit; ++z) { s0 += partial[((size_t)c * nsplit + z) * 2]; s1 += partial[((size_t)c * nsplit + z) * 2 + 1]; }
  const double mean = s0 / count;
  double var = s1 / count - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[2 * c] = (float)mean;
  mean_rstd[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
    running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
  }
}

// y = [relu]((x - mean) * rstd * gamma + beta); grid (N*C planes, chunks of the plane)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean_rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int C, long long HW, int relu, float* __restrict__ y) {
  const long long plane = blockIdx.x;
  const int c = (int)(plane % C);
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1];
  const float ga = gamma != nullptr ? gamma[c] : 1.f, be = beta != nullptr ? beta[c] : 0.f;
  const size_t base = (size_t)plane * (size_t)HW;
  for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < HW; i += (long long)gridDim.y * 256) {
    float v = (x[base + i] - mean) * rstd * ga + be;
    if (relu) v = v > 0.f ? v : (v != v ? v : 0.f);
    y[base + i] = v;
  }
}

// sums of chan_reduce mode 1 -> d_gamma += sum g*xhat, d_beta += sum g, coef[c] = (sum g / M, sum g*xhat / M)
__global__ void bn_bwd_finish_kernel(const double* __restrict__ partial, int nsplit, int C, double count,
                                     float* __restrict__ d_gamma, float* __restrict__ d_beta, float* __restrict__ coef) {
  const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int z = 0; z < nsplit; ++z) { s0 += partial[((size_t)c * nsplit + z) * 2]; s1 += partial[((size_t)c * nsplit + z) * 2 + 1]; }
  if (d_beta != nullptr) d_beta[c] += (float)s0;
  if (d_gamma != nullptr) d_gamma[c] += (float)s1;
  coef[2 * c] = (float)(s0 / count);
  coef[2 * c + 1] = (float)(s1 / count);
}

// dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat))   (training)   |   gamma * rstd * g   (evaluation)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ y, const float* __restrict__ mean_rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ coef,
                                                           int C, long long HW, int relu, int training,
                                                           float* __restrict__ dx) {
  const long long plane = blockIdx.x;
  const int c = (int)(plane % C);
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1];
  const float ga = gamma != nullptr ? gamma[c] : 1.f;
  const float a = training ? coef[2 * c] : 0.f, b = training ? coef[2 * c + 1] : 0.f;
  const size_t base = (size_t)plane * (size_t)HW;
  for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < HW; i += (long long)gridDim.y * 256) {
    float gq = dy[base + i];
    if (relu && !(y[base + i] > 0.f)) gq = 0.f;
    const float xh = (x[base + i] - mean) * rstd;
    dx[base + i] = ga * rstd * (gq - a - xh * b);
  }
}

// bias gradient from chan_reduce mode 0: db[c] += sum
__global__ void bias_grad_finish_kernel(const double* __restrict__ partial, int nsplit, int C, float* __restrict__ db) {
  const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= C) return;
  double s0 = 0.0;
  for (int z = 0; z < nsplit; ++z) s0 += partial[((size_t)c * nsplit + z) * 2];
  db[c] += (float)s0;
}

// d_pre = d_p * p * (1 - p): backward of the decoder's final sigmoid (models/common.py:148)
__global__ void __launch_bounds__(256) sigmoid_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                          long long n, float* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float q = p[i];
    dx[i] = dp[i] * q * (1.f - q);
  }
}


// ---- dense layers of the image modules (feat_to_z_mean / feat_to_z_std.0 / z_to_feat.0, models/common.py:127-133, 146-149)
// C[i, j] (+)= act(sum_l A(i, l) * B(j, l) + bias[j]) in FP32 on the FFMA pipe, operands addressed by two strides each so
// that one kernel serves y = x W^T + b, dx = dy W and dW = dy^T x without transposed copies.  Why not the tcgen05 TF32
// GEMM of bfvi_tc.cuh: the contraction over feat_dim = 4096 leaves the error-compensated 3xTF32 product at 2e-5 of the
// result (tensor-core accumulation, measured: profiles/r2_dense_accuracy.json), which flips enough BatchNorm -> ReLU masks
// downstream to move the gradients by 3e-3; the reference computes these layers in FP32.
// 64 x 64 output tile per CTA, 16-deep operand tiles in shared memory (l-major), 4 x 4 outputs per thread.
struct DenseParams {
  const float* A; long long sai, sal; int a_l_contig;
  const float* B; long long sbj, sbl; int b_l_contig;
  float* C; long long ldc;
  int M, N, K;
  const float* bias; int relu, accumulate;
  int k_per_split;      // gridDim.z > 1: each z-slice contracts k_per_split of K and adds its tile with atomicAdd
};
constexpr int kDenseTile = 64, kDenseK = 16, kDensePitch = 68;
__global__ void __launch_bounds__(256) dense_gemm_kernel(DenseParams p) {
  __shared__ __align__(16) float As[kDenseK][kDensePitch];
  __shared__ __align__(16) float Bs[kDenseK][kDensePitch];
  const int i0 = (int)blockIdx.y * kDenseTile, j0 = (int)blockIdx.x * kDenseTile;
  const int tx = (int)threadIdx.x & 15, ty = (int)threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  const bool split = gridDim.z > 1;
  const int l_lo = split ? (int)blockIdx.z * p.k_per_split : 0;
  const int l_hi = split ? (l_lo + p.k_per_split < p.K ? l_lo + p.k_per_split : p.K) : p.K;
  for (int l0 = l_lo; l0 < l_hi; l0 += kDenseK) {
#pragma unroll
    for (int e = (int)threadIdx.x; e < kDenseTile * kDenseK; e += 256) {
      int i, l;
      if (p.a_l_contig) { l = e % kDenseK; i = e / kDenseK; } else { i = e % kDenseTile; l = e / kDenseTile; }
      As[l][i] = (i0 + i < p.M && l0 + l < l_hi) ? p.A[(long long)(i0 + i) * p.sai + (long long)(l0 + l) * p.sal] : 0.f;
      int j, m;
      if (p.b_l_contig) { m = e % kDenseK; j = e / kDenseK; } else { j = e % kDenseTile; m = e / kDenseTile; }
      Bs[m][j] = (j0 + j < p.N && l0 + m < l_hi) ? p.B[(long long)(j0 + j) * p.sbj + (long long)(l0 + m) * p.sbl] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < kDenseK; ++l) {
      const float4 a = *reinterpret_cast<const float4*>(&As[l][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[l][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= p.M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= p.N) continue;
      float v = acc[r][c] + ((p.bias != nullptr && blockIdx.z == 0) ? p.bias[j] : 0.f);
      float* out = p.C + (long long)i * p.ldc + j;
      if (split) { atomicAdd(out, v); continue; }             // host side: no ReLU, C zeroed unless accumulating
      if (p.relu) v = v > 0.f ? v : (v != v ? v : 0.f);
      *out = p.accumulate ? *out + v : v;
    }
  }
}

// out = dy where y > 0, else 0: the gradient through the ReLU that follows z_to_feat.0 (models/common.py:147)
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n,
                                                        float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    out[i] = y[i] > 0.f ? dy[i] : 0.f;
}

}  // namespace conv
}  // namespace bfvi
