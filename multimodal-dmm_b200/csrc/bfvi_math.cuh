// bfvi_math.cuh — per-row device math of the BFVI step (forward and hand-derived
// backward).  Everything here is register-level code on compile-time (Z, H) so the
// loops unroll completely; weights are read from a shared-memory copy of the flat
// parameter block as 128-bit broadcast loads.
//
// Reference formulas (paths relative to the reference repository):
//   GaussianGTF           models/common.py:43-68
//   GaussianMLP           models/common.py:25-41
//   product_of_experts    models/dgts.py:15-51
//   mean_of_experts       models/dgts.py:53-83
//   kld_gauss / nll_gauss models/losses.py:14-21, 68-89
#pragma once
#include "bfvi_platform.cuh"

namespace bfvi {

constexpr float kPoeEps = 1e-8f;             // models/dgts.py:15
constexpr float kMlpMinStd = 1e-3f;          // models/common.py:27
constexpr float kHalfLog2Pi = 0.91893853320467274178f;

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }

// Offsets (floats) of one GaussianGTF inside the flat buffer; every tensor starts
// on a 16-byte boundary.  Must match bfvi_param_layout() (checked at dispatch).
template <int Z, int H>
struct GtfLayout {
  static constexpr int G0W = 0;                      // z_to_gate.0.weight (H,Z)
  static constexpr int G0B = G0W + pad4(H * Z);      // z_to_gate.0.bias   (H)
  static constexpr int G2W = G0B + pad4(H);          // z_to_gate.2.weight (Z,H)
  static constexpr int G2B = G2W + pad4(Z * H);
  static constexpr int LW = G2B + pad4(Z);           // z_lin.weight (Z,Z)
  static constexpr int LB = LW + pad4(Z * Z);
  static constexpr int N0W = LB + pad4(Z);           // z_nonlin.0.weight (H,Z)
  static constexpr int N0B = N0W + pad4(H * Z);
  static constexpr int N2W = N0B + pad4(H);          // z_nonlin.2.weight (Z,H)
  static constexpr int N2B = N2W + pad4(Z * H);
  static constexpr int SW = N2B + pad4(Z);           // z_to_std.0.weight (Z,Z)
  static constexpr int SB = SW + pad4(Z * Z);
  static constexpr int SIZE = SB + pad4(Z);
};

// ---------------------------------------------------------------- activations
__device__ __forceinline__ float softplus_f(float x) {      // torch Softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
// NaN-propagating ReLU like torch.relu (fmaxf would swallow a NaN)
__device__ __forceinline__ float relu_f(float x) { return x < 0.f ? 0.f : x; }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplus_grad(float x) { return x > 20.f ? 1.f : sigmoid_f(x); }
__device__ __forceinline__ float sign_f(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// y[o] = b[o] + sum_i W[o][i] x[i]; W row-major (OUT, IN) at a 16-byte aligned
// shared-memory address.  The flat walk over W lets 128-bit loads serve any IN.
template <int OUT, int IN>
__device__ __forceinline__ void matvec(const float* __restrict__ W, const float* __restrict__ b,
                                       const float (&x)[IN], float (&y)[OUT]) {
#pragma unroll
  for (int o = 0; o < OUT; ++o) y[o] = b[o];
  constexpr int N = OUT * IN;
  const float4* W4 = reinterpret_cast<const float4*>(W);
#pragma unroll
  for (int q = 0; q < (N + 3) / 4; ++q) {
    const float4 w = W4[q];
    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = 4 * q + j;
      if (idx < N) y[idx / IN] = fmaf(wv[j], x[idx % IN], y[idx / IN]);
    }
  }
}

// dx[i] += sum_o W[o][i] dy[o]  (transposed product, same flat walk)
template <int OUT, int IN>
__device__ __forceinline__ void matvec_t_acc(const float* __restrict__ W, const float (&dy)[OUT],
                                             float (&dx)[IN]) {
  constexpr int N = OUT * IN;
  const float4* W4 = reinterpret_cast<const float4*>(W);
#pragma unroll
  for (int q = 0; q < (N + 3) / 4; ++q) {
    const float4 w = W4[q];
    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = 4 * q + j;
      if (idx < N) dx[idx % IN] = fmaf(wv[j], dy[idx / IN], dx[idx % IN]);
    }
  }
}

// ------------------------------------------------------------------ GTF
template <int Z, int H>
struct GtfAct {       // activations kept for the backward pass
  float h1[H];        // relu(gate hidden)
  float h3[H];        // relu(nonlin hidden)
  float g[Z];         // gate
  float lin[Z];
  float nl[Z];
  float as[Z];        // pre-softplus std
};

template <int Z, int H>
__device__ __forceinline__ void gtf_forward(const float* __restrict__ W, float min_std,
                                            const float (&z)[Z], GtfAct<Z, H>& a,
                                            float (&qm)[Z], float (&qs)[Z]) {
  using L = GtfLayout<Z, H>;
  matvec<H, Z>(W + L::G0W, W + L::G0B, z, a.h1);
#pragma unroll
  for (int h = 0; h < H; ++h) a.h1[h] = relu_f(a.h1[h]);
  matvec<Z, H>(W + L::G2W, W + L::G2B, a.h1, a.g);
  matvec<Z, Z>(W + L::LW, W + L::LB, z, a.lin);
  matvec<H, Z>(W + L::N0W, W + L::N0B, z, a.h3);
#pragma unroll
  for (int h = 0; h < H; ++h) a.h3[h] = relu_f(a.h3[h]);
  matvec<Z, H>(W + L::N2W, W + L::N2B, a.h3, a.nl);
  matvec<Z, Z>(W + L::SW, W + L::SB, a.nl, a.as);
#pragma unroll
  for (int i = 0; i < Z; ++i) {
    a.g[i] = sigmoid_f(a.g[i]);
    qs[i] = softplus_f(a.as[i]) + min_std;
    qm[i] = (1.f - a.g[i]) * a.lin[i] + a.g[i] * a.nl[i];
  }
}

// Gradients at every pre-activation (what the weight-gradient panels need) and dz.
template <int Z, int H>
struct GtfGrad {
  float d_a1[H];   // gate hidden pre-relu
  float d_a3[H];   // nonlin hidden pre-relu
  float d_lin[Z];
  float d_ag[Z];   // gate pre-sigmoid
  float d_nl[Z];
  float d_as[Z];   // std pre-softplus
};

template <int Z, int H>
__device__ __forceinline__ void gtf_backward(const float* __restrict__ W, const GtfAct<Z, H>& a,
                                             const float (&d_qm)[Z], const float (&d_qs)[Z],
                                             GtfGrad<Z, H>& g, float (&dz)[Z]) {
  using L = GtfLayout<Z, H>;
#pragma unroll
  for (int i = 0; i < Z; ++i) {
    g.d_as[i] = d_qs[i] * softplus_grad(a.as[i]);
    g.d_nl[i] = d_qm[i] * a.g[i];
    g.d_ag[i] = d_qm[i] * (a.nl[i] - a.lin[i]) * a.g[i] * (1.f - a.g[i]);
    g.d_lin[i] = d_qm[i] * (1.f - a.g[i]);
    dz[i] = 0.f;
  }
  matvec_t_acc<Z, Z>(W + L::SW, g.d_as, g.d_nl);           // nl feeds the std head too
#pragma unroll
  for (int h = 0; h < H; ++h) { g.d_a1[h] = 0.f; g.d_a3[h] = 0.f; }
  matvec_t_acc<Z, H>(W + L::N2W, g.d_nl, g.d_a3);
  matvec_t_acc<Z, H>(W + L::G2W, g.d_ag, g.d_a1);
#pragma unroll
  for (int h = 0; h < H; ++h) {
    g.d_a3[h] = a.h3[h] > 0.f ? g.d_a3[h] : 0.f;
    g.d_a1[h] = a.h1[h] > 0.f ? g.d_a1[h] : 0.f;
  }
  matvec_t_acc<H, Z>(W + L::G0W, g.d_a1, dz);
  matvec_t_acc<H, Z>(W + L::N0W, g.d_a3, dz);
  matvec_t_acc<Z, Z>(W + L::LW, g.d_lin, dz);
}

// ------------------------------------------------ product / mixture of experts
// precision with the sign trick of models/dgts.py:40-42
__device__ __forceinline__ float poe_prec(float std) {
  return 1.f / (std * std + kPoeEps) * sign_f(std);
}
// d prec / d std
__device__ __forceinline__ float poe_prec_grad(float std, float prec) {
  return -2.f * std * prec / (std * std + kPoeEps);
}

// p(z|z_prev) = p(z) * q'(z|z_prev)   (models/dmm.py:239-245), one component
__device__ __forceinline__ void poe2_forward(float gm, float gs, float qm, float qs,
                                             float& pm, float& ps) {
  const float tg = poe_prec(gs), tq = poe_prec(qs);
  const float s = tg + tq;
  float m = (gm * tg + qm * tq) / s;
  pm = (m != m) ? 0.f : m;
  ps = sqrtf(1.f / s);
}
__device__ __forceinline__ void poe2_backward(float gm, float gs, float qm, float qs, float pm,
                                              float ps, float d_pm, float d_ps, float& d_gm,
                                              float& d_gs, float& d_qm, float& d_qs) {
  const float tg = poe_prec(gs), tq = poe_prec(qs);
  const float s = tg + tq;
  const float d_n = d_pm / s;
  const float d_s = -d_pm * pm / s - 0.5f * d_ps * ps / s;
  d_gm = d_n * tg;
  d_qm = d_n * tq;
  d_gs = (d_n * gm + d_s) * poe_prec_grad(gs, tg);
  d_qs = (d_n * qm + d_s) * poe_prec_grad(qs, tq);
}

// ------------------------------------------------------------------ losses
// one element of losses.kld_gauss (before the 0.5 factor is applied: included here)
__device__ __forceinline__ float kld_elem(float m1, float s1, float m2, float s2) {
  const float dm = m1 - m2;
  return 0.5f * (2.f * logf(s2) - 2.f * logf(s1) + (s1 * s1 + dm * dm) / (s2 * s2) - 1.f);
}
__device__ __forceinline__ void kld_elem_grad(float m1, float s1, float m2, float s2, float c,
                                              float& d_m1, float& d_s1, float& d_m2, float& d_s2) {
  const float dm = m1 - m2, iv = 1.f / (s2 * s2);
  d_m1 = c * dm * iv;
  d_m2 = -d_m1;
  d_s1 = c * (s1 * iv - 1.f / s1);
  d_s2 = c * (1.f / s2 - (s1 * s1 + dm * dm) * iv / s2);
}
__device__ __forceinline__ float nll_gauss_elem(float mean, float std, float x) {
  const float r = (x - mean) / std;
  return 0.5f * r * r + logf(std) + kHalfLog2Pi;
}
__device__ __forceinline__ void nll_gauss_elem_grad(float mean, float std, float x, float c,
                                                    float& d_mean, float& d_std) {
  const float r = (x - mean) / std;
  d_mean = -c * r / std;
  d_std = c * (1.f - r * r) / std;
}

// ------------------------------------------------------------ warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace bfvi
