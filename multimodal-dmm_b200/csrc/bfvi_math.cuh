// bfvi_math.cuh — per-row device math of the BFVI step (forward and hand-derived
// backward) for the register-resident small-dim path.  Everything here is
// register-level code on compile-time (Z, H) so the loops unroll completely.
//
// The gated transition is evaluated "unit-streaming": for every hidden unit the
// pre-activation (a dot product with z), its ReLU and the unit's contribution to
// the Z outputs are computed back to back, so no H-sized activation array ever
// lives in registers (the v0 kernels spilled 1.5-2 KB per thread on exactly those
// arrays).  Weights are read from a re-packed, unit-major shared-memory copy
// (GtfPack) with 128-bit broadcast loads; R rows per thread share each load.
//
// Reference formulas (paths relative to the reference repository):
//   GaussianGTF           models/common.py:43-68
//   GaussianMLP           models/common.py:25-41
//   product_of_experts    models/dgts.py:15-51
//   mean_of_experts       models/dgts.py:53-83
//   kld_gauss / nll_gauss models/losses.py:14-21, 68-89
#pragma once
#include "bfvi_platform.cuh"
#include "bfvi_wgrad.cuh"

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

// Unit-major shared-memory packing of one GTF:
//   gate unit h   : [ b0[h], W0[h][0..Z), W2[0..Z)[h] ]   (U floats, 16-byte aligned)
//   nonlin unit h : same for the z_nonlin branch
//   lin row o     : [ b[o], W[o][0..Z) ]                  (RW floats)
//   std row o     : same for z_to_std
//   B2G / B2N     : output biases of z_to_gate.2 / z_nonlin.2
template <int Z, int H>
struct GtfPack {
  static constexpr int U = pad4(1 + 2 * Z);
  static constexpr int RW = pad4(1 + Z);
  static constexpr int GATE = 0;
  static constexpr int NONLIN = GATE + H * U;
  static constexpr int LIN = NONLIN + H * U;
  static constexpr int STD = LIN + Z * RW;
  static constexpr int B2G = STD + Z * RW;
  static constexpr int B2N = B2G + pad4(Z);
  // the two hidden branches interleaved: unit h, column c -> (gate, nonlin) adjacent, so one
  // 128-bit load yields two register PAIRS for packed FFMA2 (bfvi_wgrad.cuh: ffma2)
  static constexpr int U2 = 2 * U;
  static constexpr int PAIR = B2N + pad4(Z);
  static constexpr int SIZE = PAIR + H * U2;
};

// cooperative re-pack global flat GTF block -> shared GtfPack (all threads of the CTA)
template <int Z, int H>
__device__ inline void gtf_pack_load(const float* __restrict__ w, float* __restrict__ sP) {
  using L = GtfLayout<Z, H>;
  using P = GtfPack<Z, H>;
  for (int i = threadIdx.x; i < P::SIZE; i += blockDim.x) {
    float v = 0.f;
    if (i < P::LIN) {
      const int br = i >= P::NONLIN;                      // 0 gate, 1 nonlin
      const int j = i - (br ? P::NONLIN : P::GATE);
      const int h = j / P::U, c = j % P::U;
      const int w0 = br ? L::N0W : L::G0W, b0 = br ? L::N0B : L::G0B, w2 = br ? L::N2W : L::G2W;
      if (c == 0) v = w[b0 + h];
      else if (c <= Z) v = w[w0 + h * Z + (c - 1)];
      else if (c <= 2 * Z) v = w[w2 + (c - 1 - Z) * H + h];
    } else if (i < P::B2G) {
      const int st = i >= P::STD;
      const int j = i - (st ? P::STD : P::LIN);
      const int o = j / P::RW, c = j % P::RW;
      const int ww = st ? L::SW : L::LW, bb = st ? L::SB : L::LB;
      if (c == 0) v = w[bb + o];
      else if (c <= Z) v = w[ww + o * Z + (c - 1)];
    } else if (i < P::B2N) {
      const int o = i - P::B2G;
      if (o < Z) v = w[L::G2B + o];
    } else if (i < P::PAIR) {
      const int o = i - P::B2N;
      if (o < Z) v = w[L::N2B + o];
    } else {
      const int j = i - P::PAIR;
      const int h = j / P::U2, c = (j % P::U2) >> 1, br = j & 1;
      const int w0 = br ? L::N0W : L::G0W, b0 = br ? L::N0B : L::G0B, w2 = br ? L::N2W : L::G2W;
      if (c == 0) v = w[b0 + h];
      else if (c <= Z) v = w[w0 + h * Z + (c - 1)];
      else if (c <= 2 * Z) v = w[w2 + (c - 1 - Z) * H + h];
    }
    sP[i] = v;
  }
}

// N floats from a 16-byte aligned shared address with 128-bit loads (N % 4 == 0)
template <int N>
__device__ __forceinline__ void lds_vec(const float* __restrict__ p, float (&v)[N]) {
  static_assert(N % 4 == 0, "vector loads need a multiple of 4 floats");
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    const float4 t = p4[q];
    v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

// L1 prefetch of one address (no register result, no scoreboard wait)
__device__ __forceinline__ void prefetch_l1(const void* p) {
#ifndef BFVI_EMU
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// ---------------------------------------------------------------- fast scalar ops
// Single-instruction MUFU approximations (about 1-2 ulp, denormals flushed) for the
// WELL-CONDITIONED per-particle math.  The product-of-experts step keeps IEEE-rounded
// operations: its inverse-prior expert cancels precisions and amplifies every
// rounding error.
#ifdef BFVI_EMU
__device__ __forceinline__ float fast_rcp(float x) { return 1.f / x; }
__device__ __forceinline__ float fast_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float fast_ex2(float x) { return exp2f(x); }
__device__ __forceinline__ float fast_lg2(float x) { return log2f(x); }
// NaN-propagating max like torch.relu (fmaxf would swallow a NaN)
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
#else
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float max_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
#endif
__device__ __forceinline__ float fast_div(float a, float b) { return a * fast_rcp(b); }
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------- activations
// torch Softplus(beta=1, threshold=20) = max(x,0) + log1p(exp(-|x|)) (above the
// threshold the log term is below half an ulp of x, so no select is needed).
// log1p(u) for u in (0,1] is evaluated as 2*atanh(s), s = u/(2+u) <= 1/3, with the odd
// series up to s^15: relative error < 1e-7 over the whole range (MUFU lg2 has an
// ABSOLUTE error of 2^-22, i.e. 1e-5 relative for small results, which the reference's
// cancelling product of experts amplifies beyond the parity tolerance).
__device__ __forceinline__ float log1p_unit(float u) {
  const float s = u * fast_rcp(2.f + u), q = s * s;
  float p = 1.f / 15.f;
  p = fmaf(p, q, 1.f / 13.f);
  p = fmaf(p, q, 1.f / 11.f);
  p = fmaf(p, q, 1.f / 9.f);
  p = fmaf(p, q, 1.f / 7.f);
  p = fmaf(p, q, 1.f / 5.f);
  p = fmaf(p, q, 1.f / 3.f);
  p = fmaf(p, q, 1.f);
  return (s + s) * p;
}
__device__ __forceinline__ float softplus_f(float x) {
  return max_nan(x, 0.f) + log1p_unit(fast_ex2(-fabsf(x) * kLog2e));
}
__device__ __forceinline__ float relu_f(float x) { return max_nan(x, 0.f); }
__device__ __forceinline__ float sigmoid_f(float x) { return fast_rcp(1.f + fast_ex2(-x * kLog2e)); }
// d softplus / dx (equals 1 to fp32 precision above torch's threshold of 20)
__device__ __forceinline__ float softplus_grad(float x) { return sigmoid_f(x); }
__device__ __forceinline__ float sign_f(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// ------------------------------------------------------------------ GTF forward
// R rows per thread; q'(z_next | z) mean / std (models/common.py:62-68) of row r,
// component o are handed to `epi(r, o, mean, std)` as soon as they are complete.
#ifndef BFVI_GTF_PAIR
#define BFVI_GTF_PAIR 1          // packed-FFMA2 hidden layers in the backward kernel (A/B build knob)
#endif
__device__ __forceinline__ float2 ffma2_rn(float2 a, float2 b, float2 c) {
#ifdef BFVI_EMU
  return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#else
  return __ffma2_rn(a, b, c);
#endif
}
// (R rows per lane stay scalar FFMA: pairing the branches needs every z duplicated into a
//  register pair, which spills at 128 registers; pairing only the second layer measured
//  1 % SLOWER at C2 — the kernel is bound by dependent-issue latency, not FMA issue slots.)
template <int Z, int H, int R, typename Epi>
__device__ __forceinline__ void gtf_rows_forward(const float* __restrict__ sP, float min_std,
                                                 const float (&z)[R][Z], Epi&& epi) {
  using P = GtfPack<Z, H>;
  float g[R][Z], nl[R][Z];
#pragma unroll
  for (int o = 0; o < Z; ++o) {
    const float bg = sP[P::B2G + o], bn = sP[P::B2N + o];
#pragma unroll
    for (int r = 0; r < R; ++r) { g[r][o] = bg; nl[r][o] = bn; }
  }
#pragma unroll 2
  for (int h = 0; h < H; ++h) {
    float wg[P::U], wn[P::U];
    lds_vec<P::U>(sP + P::GATE + h * P::U, wg);
    lds_vec<P::U>(sP + P::NONLIN + h * P::U, wn);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float a = wg[0], c = wn[0];
#pragma unroll
      for (int i = 0; i < Z; ++i) { a = fmaf(wg[1 + i], z[r][i], a); c = fmaf(wn[1 + i], z[r][i], c); }
      a = relu_f(a); c = relu_f(c);
#pragma unroll
      for (int o = 0; o < Z; ++o) {
        g[r][o] = fmaf(wg[1 + Z + o], a, g[r][o]);
        nl[r][o] = fmaf(wn[1 + Z + o], c, nl[r][o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < Z; ++o) {
    float wl[P::RW], ws[P::RW];
    lds_vec<P::RW>(sP + P::LIN + o * P::RW, wl);
    lds_vec<P::RW>(sP + P::STD + o * P::RW, ws);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float lin = wl[0], as = ws[0];
#pragma unroll
      for (int i = 0; i < Z; ++i) { lin = fmaf(wl[1 + i], z[r][i], lin); as = fmaf(ws[1 + i], nl[r][i], as); }
      const float gate = sigmoid_f(g[r][o]);
      epi(r, o, fmaf(gate, nl[r][o] - lin, lin), softplus_f(as) + min_std);   // (1-g)*lin + g*nl
    }
  }
}

// Panel columns of the weight-gradient staging area (bfvi_wgrad.cuh): X holds the
// layer inputs [1, z] [1, h1] [1, h3] [1, nl]; D the pre-activation gradients.
template <int Z, int H>
struct GtfCols {
  static constexpr int XZ = 0, XH1 = 1 + Z, XH3 = XH1 + 1 + H, XNL = XH3 + 1 + H, NXC = XNL + 1 + Z;
  static constexpr int DA1 = 0, DA3 = H, DLIN = 2 * H, DAG = 2 * H + Z, DNL = 2 * H + 2 * Z,
                       DAS = 2 * H + 3 * Z, NDC = 2 * H + 4 * Z;
};

// One row forward for the backward pass: returns gate (post-sigmoid), lin, nl and
// the pre-softplus std, and writes the hidden activations relu(a1), relu(a3) straight
// into this lane's slot of the X staging panel (panel_at(column, lane)), where the
// weight-gradient tiles AND the backward unit loop below read them back.
template <int Z, int H>
__device__ __forceinline__ void gtf_row_forward_stage(const float* __restrict__ sP, const float (&z)[Z],
                                                      float (&g)[Z], float (&lin)[Z], float (&nl)[Z],
                                                      float (&as)[Z], float* __restrict__ Xp, int lane) {
  using P = GtfPack<Z, H>;
  using C = GtfCols<Z, H>;
#if BFVI_GTF_PAIR
  // one row per lane: (gate, nonlin) register pairs through both hidden layers, FFMA2 throughout
  float2 gn[Z], zz[Z];
#pragma unroll
  for (int o = 0; o < Z; ++o) { gn[o] = float2{sP[P::B2G + o], sP[P::B2N + o]}; zz[o] = float2{z[o], z[o]}; }
#pragma unroll 4
  for (int h = 0; h < H; ++h) {
    float w[P::U2];
    lds_vec<P::U2>(sP + P::PAIR + h * P::U2, w);
    float2 ac{w[0], w[1]};
#pragma unroll
    for (int i = 0; i < Z; ++i) ac = ffma2_rn(float2{w[2 + 2 * i], w[3 + 2 * i]}, zz[i], ac);
    ac.x = relu_f(ac.x); ac.y = relu_f(ac.y);
    Xp[panel_at(C::XH1 + 1 + h, lane)] = ac.x;
    Xp[panel_at(C::XH3 + 1 + h, lane)] = ac.y;
#pragma unroll
    for (int o = 0; o < Z; ++o) gn[o] = ffma2_rn(float2{w[2 + 2 * (Z + o)], w[3 + 2 * (Z + o)]}, ac, gn[o]);
  }
#pragma unroll
  for (int o = 0; o < Z; ++o) { g[o] = gn[o].x; nl[o] = gn[o].y; }
#else
#pragma unroll
  for (int o = 0; o < Z; ++o) { g[o] = sP[P::B2G + o]; nl[o] = sP[P::B2N + o]; }
#pragma unroll 4
  for (int h = 0; h < H; ++h) {
    float wg[P::U], wn[P::U];
    lds_vec<P::U>(sP + P::GATE + h * P::U, wg);
    lds_vec<P::U>(sP + P::NONLIN + h * P::U, wn);
    float a = wg[0], c = wn[0];
#pragma unroll
    for (int i = 0; i < Z; ++i) { a = fmaf(wg[1 + i], z[i], a); c = fmaf(wn[1 + i], z[i], c); }
    a = relu_f(a); c = relu_f(c);
    Xp[panel_at(C::XH1 + 1 + h, lane)] = a;
    Xp[panel_at(C::XH3 + 1 + h, lane)] = c;
#pragma unroll
    for (int o = 0; o < Z; ++o) { g[o] = fmaf(wg[1 + Z + o], a, g[o]); nl[o] = fmaf(wn[1 + Z + o], c, nl[o]); }
  }
#endif
#pragma unroll
  for (int o = 0; o < Z; ++o) {
    float wl[P::RW], ws[P::RW];
    lds_vec<P::RW>(sP + P::LIN + o * P::RW, wl);
    lds_vec<P::RW>(sP + P::STD + o * P::RW, ws);
    float l = wl[0], s = ws[0];
#pragma unroll
    for (int i = 0; i < Z; ++i) { l = fmaf(wl[1 + i], z[i], l); s = fmaf(ws[1 + i], nl[i], s); }
    lin[o] = l; as[o] = s;
    g[o] = sigmoid_f(g[o]);
    Xp[panel_at(C::XZ + 1 + o, lane)] = z[o];
    Xp[panel_at(C::XNL + 1 + o, lane)] = nl[o];
  }
}

// Backward of one row given d_qm / d_qs.  Hidden activations come back from the X
// panel (gtf_row_forward_stage); the pre-activation gradients go to the D panel,
// all zero when the row is padding (so the X entries of padding rows, copies of a
// real row, contribute nothing).  Returns dz.
template <int Z, int H>
__device__ __forceinline__ void gtf_row_backward_stage(const float* __restrict__ sP, const float (&g)[Z],
                                                       const float (&lin)[Z], const float (&nl)[Z],
                                                       const float (&as)[Z], const float (&d_qm)[Z],
                                                       const float (&d_qs)[Z], float (&dz)[Z],
                                                       const float* __restrict__ Xp, float* __restrict__ Dp,
                                                       int lane, bool valid) {
  using P = GtfPack<Z, H>;
  using C = GtfCols<Z, H>;
  const float vm = valid ? 1.f : 0.f;
  float d_as[Z], d_nl[Z], d_ag[Z], d_lin[Z];
#pragma unroll
  for (int o = 0; o < Z; ++o) {
    const float dq = d_qm[o] * vm;
    d_as[o] = d_qs[o] * vm * softplus_grad(as[o]);
    d_nl[o] = dq * g[o];
    d_lin[o] = dq - d_nl[o];                               // dq * (1 - g)
    d_ag[o] = d_lin[o] * g[o] * (nl[o] - lin[o]);          // dq * (nl - lin) * g * (1 - g)
    dz[o] = 0.f;
  }
#pragma unroll
  for (int o = 0; o < Z; ++o) {
    float wl[P::RW], ws[P::RW];
    lds_vec<P::RW>(sP + P::LIN + o * P::RW, wl);
    lds_vec<P::RW>(sP + P::STD + o * P::RW, ws);
#pragma unroll
    for (int i = 0; i < Z; ++i) {
      d_nl[i] = fmaf(ws[1 + i], d_as[o], d_nl[i]);       // nl feeds the std head too
      dz[i] = fmaf(wl[1 + i], d_lin[o], dz[i]);
    }
  }
#pragma unroll
  for (int o = 0; o < Z; ++o) {
    Dp[panel_at(C::DLIN + o, lane)] = d_lin[o];
    Dp[panel_at(C::DAG + o, lane)] = d_ag[o];
    Dp[panel_at(C::DNL + o, lane)] = d_nl[o];
    Dp[panel_at(C::DAS + o, lane)] = d_as[o];
  }
#if BFVI_GTF_PAIR
  float2 dd[Z], dzz[Z];                                    // (gate path, nonlin path) pairs
#pragma unroll
  for (int o = 0; o < Z; ++o) { dd[o] = float2{d_ag[o], d_nl[o]}; dzz[o] = float2{0.f, 0.f}; }
#pragma unroll 4
  for (int h = 0; h < H; ++h) {
    float w[P::U2];
    lds_vec<P::U2>(sP + P::PAIR + h * P::U2, w);
    const float h1 = Xp[panel_at(C::XH1 + 1 + h, lane)], h3 = Xp[panel_at(C::XH3 + 1 + h, lane)];
    float2 dac{0.f, 0.f};
#pragma unroll
    for (int o = 0; o < Z; ++o) dac = ffma2_rn(float2{w[2 + 2 * (Z + o)], w[3 + 2 * (Z + o)]}, dd[o], dac);
    dac.x = h1 > 0.f ? dac.x : 0.f;
    dac.y = h3 > 0.f ? dac.y : 0.f;
#pragma unroll
    for (int i = 0; i < Z; ++i) dzz[i] = ffma2_rn(float2{w[2 + 2 * i], w[3 + 2 * i]}, dac, dzz[i]);
    Dp[panel_at(C::DA1 + h, lane)] = dac.x;
    Dp[panel_at(C::DA3 + h, lane)] = dac.y;
  }
#pragma unroll
  for (int i = 0; i < Z; ++i) dz[i] += dzz[i].x + dzz[i].y;
#else
#pragma unroll 4
  for (int h = 0; h < H; ++h) {
    float wg[P::U], wn[P::U];
    lds_vec<P::U>(sP + P::GATE + h * P::U, wg);
    lds_vec<P::U>(sP + P::NONLIN + h * P::U, wn);
    const float h1 = Xp[panel_at(C::XH1 + 1 + h, lane)], h3 = Xp[panel_at(C::XH3 + 1 + h, lane)];
    float da = 0.f, dc = 0.f;
#pragma unroll
    for (int o = 0; o < Z; ++o) { da = fmaf(wg[1 + Z + o], d_ag[o], da); dc = fmaf(wn[1 + Z + o], d_nl[o], dc); }
    da = h1 > 0.f ? da : 0.f;
    dc = h3 > 0.f ? dc : 0.f;
#pragma unroll
    for (int i = 0; i < Z; ++i) { dz[i] = fmaf(wg[1 + i], da, dz[i]); dz[i] = fmaf(wn[1 + i], dc, dz[i]); }
    Dp[panel_at(C::DA1 + h, lane)] = da;
    Dp[panel_at(C::DA3 + h, lane)] = dc;
  }
#endif
}

// ------------------------------------------------ product / mixture of experts
// precision with the sign trick of models/dgts.py:40-42, IEEE-rounded operations in
// the reference's order (1/var, then * sign)
__device__ __forceinline__ float poe_prec(float std) {
  return __fmul_rn(__fdiv_rn(1.f, __fadd_rn(__fmul_rn(std, std), kPoeEps)), sign_f(std));
}
// d prec / d std
__device__ __forceinline__ float poe_prec_grad(float std, float prec) {
  return -2.f * std * prec / (std * std + kPoeEps);
}
// Backward-pass form of both: gradients only need the reciprocal to an ulp or two, so the MUFU
// reciprocal replaces two IEEE divisions (whose slow paths were 7 % of the dominant kernel's
// stall samples); the FORWARD keeps poe_prec, its values feed the inverse-prior cancellation.
__device__ __forceinline__ void poe_prec_bwd(float std, float& prec, float& dprec) {
  const float r = fast_rcp(fmaf(std, std, kPoeEps));
  prec = r * sign_f(std);
  dprec = -2.f * std * prec * r;
}

// p(z|z_prev) = p(z) * q'(z|z_prev)   (models/dmm.py:239-245), one component, both
// stds positive (global prior and a GTF output).  Same value as the reference's
// sum of precisions, written with ONE division: with vg = gs^2+eps, vq = qs^2+eps,
//   mean = (gm*vq + qm*vg) / (vg+vq),   var = vg*vq / (vg+vq).
__device__ __forceinline__ void poe2_forward(float gm, float gs, float qm, float qs,
                                             float& pm, float& ps) {
  const float vg = fmaf(gs, gs, kPoeEps), vq = fmaf(qs, qs, kPoeEps);
  const float r = fast_rcp(vg + vq);
  const float m = (gm * vq + qm * vg) * r;
  pm = (m != m) ? 0.f : m;                    // product_mean[isnan] = 0, models/dgts.py:49
  ps = fast_sqrt(vg * vq * r);
}
__device__ __forceinline__ void poe2_backward(float gm, float gs, float qm, float qs, float pm,
                                              float ps, float d_pm, float d_ps, float& d_gm,
                                              float& d_gs, float& d_qm, float& d_qs) {
  const float vg = fmaf(gs, gs, kPoeEps), vq = fmaf(qs, qs, kPoeEps);
  const float r = fast_rcp(vg + vq);
  const float wg = vq * r, wq = vg * r;             // weights of gm / qm in the mean
  const float d_var = 0.5f * d_ps * fast_rcp(ps);
  d_gm = d_pm * wg;
  d_qm = d_pm * wq;
  d_gs = (d_pm * (qm - pm) * r + d_var * wg * wg) * 2.f * gs;
  d_qs = (d_pm * (gm - pm) * r + d_var * wq * wq) * 2.f * qs;
}

// ------------------------------------------------------------------ losses
// one element of losses.kld_gauss (the 0.5 factor included)
__device__ __forceinline__ float kld_elem(float m1, float s1, float m2, float s2) {
  const float dm = m1 - m2;
  return 0.5f * (2.f * logf(s2) - 2.f * logf(s1) + (s1 * s1 + dm * dm) / (s2 * s2) - 1.f);
}
// the same on the MUFU path for the fused per-step KL of the chain kernels: one lg2 of
// the ratio (absolute error 2^-22 per element, far inside the 1e-4 ELBO tolerance)
__device__ __forceinline__ float kld_elem_fast(float m1, float s1, float m2, float s2) {
  const float dm = m1 - m2, r2 = fast_rcp(s2);
  return 0.6931471805599453f * fast_lg2(s2 * fast_rcp(s1)) + 0.5f * ((s1 * s1 + dm * dm) * r2 * r2 - 1.f);
}
__device__ __forceinline__ void kld_elem_grad(float m1, float s1, float m2, float s2, float c,
                                              float& d_m1, float& d_s1, float& d_m2, float& d_s2) {
  const float dm = m1 - m2, r2 = fast_rcp(s2), iv = r2 * r2;
  d_m1 = c * dm * iv;
  d_m2 = -d_m1;
  d_s1 = c * (s1 * iv - fast_rcp(s1));
  d_s2 = c * (r2 - (s1 * s1 + dm * dm) * iv * r2);
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

// Sum over the L consecutive lanes [base, base+L) of a lane group, in lane order so
// that every lane of the group gets the bit-identical result.  All 32 lanes call it.
__device__ __forceinline__ float group_sum(float v, int base, int L) {
  float s = __shfl_sync(0xffffffffu, v, base);
  for (int j = 1; j < L; ++j) s += __shfl_sync(0xffffffffu, v, base + j);
  return s;
}

// the same for N values at once (one shuffle loop instead of N)
template <int N>
__device__ __forceinline__ void group_sum_vec(float (&v)[N], int base, int L) {
  float s[N];
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = __shfl_sync(0xffffffffu, v[i], base);
  for (int j = 1; j < L; ++j) {
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] += __shfl_sync(0xffffffffu, v[i], base + j);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = s[i];
}

}  // namespace bfvi
