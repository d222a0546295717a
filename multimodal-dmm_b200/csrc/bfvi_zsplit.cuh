// bfvi_zsplit.cuh — single-particle (K = 1) z_filter kernels of the small-dim family.
//
// The f_mode pass and the smoothing pass of a step have ONE particle per chain: only
// (1+M)*B chains of T serial steps each, far too few warps to fill a B200, so what matters is
// the LATENCY of one step.  Here a chain owns Z consecutive lanes ("z-split"):
//
//   * lane j owns latent component j for everything that is per-component: the product of
//     experts (expert loads are coalesced across the Z lanes), the sigmoid / softplus heads,
//     the product with the global prior, KL, sampling (its own Philox component), the stores;
//   * lane j also owns H/Z hidden units of both GTF branches: it computes their
//     pre-activations from the gathered z (Z shuffles), their ReLU, and their contribution
//     to all Z outputs; a reduce-scatter over the Z lanes (shuffles) leaves output o on lane o;
//   * in the backward kernel the weight gradients of a lane's units (and of its rows of
//     z_lin / z_to_std) are LOCAL to the lane: they accumulate in registers over the whole
//     time loop and every chain the lane serves, and are flushed once per CTA — no staging
//     panels, no rank-32 updates.
//
// A step costs ~600 (forward) / ~900 (backward) dependent instructions per lane instead of
// ~2400 / ~5100 with one chain per lane (bfvi_chain.cuh, R = 1).
//
// Reference semantics as in bfvi_chain.cuh: MultiDMM.z_filter / z_next (models/dmm.py:214-258,
// 319-412), product_of_experts (models/dgts.py:15-51), GaussianGTF (models/common.py:43-68).
#pragma once
#include "bfvi_chain.cuh"
#include "bfvi_generic.cuh"

namespace bfvi {

constexpr int kZsplitThreads = 64;
// (a one-step-ahead prefetch.global.L1 of the per-step operands was measured SLOWER here:
//  forward 0.44 -> 0.49 ms at C2, 0.31 -> 0.39 ms at C1 — DESIGN.md "tried and measured")

// partial[idx] for a runtime idx in [0, Z)
template <int Z>
__device__ __forceinline__ float pick(const float (&v)[Z], int idx) {
  float r = v[0];
#pragma unroll
  for (int i = 1; i < Z; ++i) r = idx == i ? v[i] : r;
  return r;
}
// every lane of a Z-lane group holds partial[0..Z): lane j returns sum over the group of partial[j]
template <int Z>
__device__ __forceinline__ float reduce_scatter(const float (&partial)[Z], int base, int j) {
  float tot = pick<Z>(partial, j);
#pragma unroll
  for (int s = 1; s < Z; ++s) {
    const int dst = j + s < Z ? j + s : j + s - Z;         // the lane this value is for
    const int src = j - s >= 0 ? j - s : j - s + Z;        // the lane whose value is for me
    tot += __shfl_sync(0xffffffffu, pick<Z>(partial, dst), base + src);
  }
  return tot;
}
template <int Z>
__device__ __forceinline__ void gather(float own, int base, float (&v)[Z]) {
#pragma unroll
  for (int i = 0; i < Z; ++i) v[i] = __shfl_sync(0xffffffffu, own, base + i);
}

// one component of the product of experts at (s, t, b): prior first, then the chain set's
// experts in order, IEEE-rounded like poe_step_forward
__device__ __forceinline__ void poe_component(const bfvi_filter_args& a, unsigned bits, int s, int t, int b, int j,
                                              float gm, float gs, float pm, float ps, float& mu, float& sd) {
  float S = poe_prec(ps);
  float N = __fmul_rn(pm, S);
  for (int e = 0; e < a.n_experts; ++e) {
    if (!((bits >> e) & 1u)) continue;
    const bfvi_expert& ex = a.experts[e];
    bool m = true;
    if (ex.mask != nullptr) m = ex.mask[s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b] != 0;
    if (ex.zero_mask_last_t && t == a.T - 1) m = false;
    const float w = m ? 1.f : 0.f;
    float mean, std;
    if (ex.kind == BFVI_EXPERT_INV_PRIOR) { mean = gm; std = -gs; }
    else {
      const int64_t off = s * ex.stride_s + t * ex.stride_t + b * ex.stride_b + j;
      mean = ex.mean[off]; std = ex.std[off];
    }
    const float te = __fmul_rn(poe_prec(std), w);
    S = __fadd_rn(S, te);
    N = __fadd_rn(N, __fmul_rn(__fmul_rn(mean, w), te));
  }
  const float mq = __fdiv_rn(N, S);
  mu = (mq != mq) ? 0.f : mq;
  sd = __fsqrt_rn(__fdiv_rn(1.f, S));
}

// GTF forward for one chain spread over Z lanes.  zv = gathered input; returns this lane's
// component of gate (post-sigmoid), lin, nl, pre-softplus std, and (optionally) the ReLU
// activations of the lane's units.
template <int Z, int H, bool KEEP>
__device__ __forceinline__ void zsplit_gtf_forward(const float* __restrict__ sP, int base, int j,
                                                   const float (&zv)[Z], float& gate, float& lin, float& nl,
                                                   float& as, float (&nlv)[Z], float (&ag)[H / Z],
                                                   float (&an)[H / Z]) {
  using P = GtfPack<Z, H>;
  constexpr int UPL = H / Z;
  float pg[Z], pn[Z];
#pragma unroll
  for (int o = 0; o < Z; ++o) pg[o] = pn[o] = 0.f;
#pragma unroll
  for (int u = 0; u < UPL; ++u) {
    const int h = j * UPL + u;
    float wg[P::U], wn[P::U];
    lds_vec<P::U>(sP + P::GATE + h * P::U, wg);
    lds_vec<P::U>(sP + P::NONLIN + h * P::U, wn);
    float a = wg[0], c = wn[0];
#pragma unroll
    for (int i = 0; i < Z; ++i) { a = fmaf(wg[1 + i], zv[i], a); c = fmaf(wn[1 + i], zv[i], c); }
    a = relu_f(a); c = relu_f(c);
    if (KEEP) { ag[u] = a; an[u] = c; }
#pragma unroll
    for (int o = 0; o < Z; ++o) { pg[o] = fmaf(wg[1 + Z + o], a, pg[o]); pn[o] = fmaf(wn[1 + Z + o], c, pn[o]); }
  }
  const float g = reduce_scatter<Z>(pg, base, j) + sP[P::B2G + j];
  nl = reduce_scatter<Z>(pn, base, j) + sP[P::B2N + j];
  gather<Z>(nl, base, nlv);
  float wl[P::RW], ws[P::RW];
  lds_vec<P::RW>(sP + P::LIN + j * P::RW, wl);
  lds_vec<P::RW>(sP + P::STD + j * P::RW, ws);
  lin = wl[0]; as = ws[0];
#pragma unroll
  for (int i = 0; i < Z; ++i) { lin = fmaf(wl[1 + i], zv[i], lin); as = fmaf(ws[1 + i], nlv[i], as); }
  gate = sigmoid_f(g);
}

struct ZGroup {
  int cpw, cig, j, base;
  bool on;
  __device__ __forceinline__ explicit ZGroup(int Z) {
    const int lane = threadIdx.x & 31;
    cpw = 32 / Z; cig = lane / Z; j = lane - cig * Z;
    on = cig < cpw;
    base = on ? cig * Z : 0;
  }
};

// =========================================================================
// forward
// =========================================================================
template <int Z, int H>
__global__ void __launch_bounds__(kZsplitThreads) zsplit_fwd_kernel(const __grid_constant__ FilterParams p) {
  static_assert(H % Z == 0, "z-split needs H to be a multiple of Z");
  using P = GtfPack<Z, H>;
  __shared__ __align__(16) float sP[P::SIZE];
  const bfvi_filter_args& a = p.a;
  gtf_pack_load<Z, H>(p.trans_w, sP);
  __syncthreads();
  const ZGroup zg(Z);
  const int j = zg.j;
  const float gm = p.z0_mean[j], gs = expf(p.z0_log_std[j]) + p.min_std;        // models/dmm.py:126-127
  const int T = a.T, B = a.B;
  const int Bc = p.bc > 0 ? p.bc : B;
  const int n_chains = a.S * Bc;
  const int n_tasks = (n_chains + zg.cpw - 1) / zg.cpw;
  const int wpb = blockDim.x >> 5;
  float kl_sum = 0.f;
  for (int task = blockIdx.x * wpb + (threadIdx.x >> 5); task < n_tasks; task += gridDim.x * wpb) {
    const int chain_raw = task * zg.cpw + zg.cig;
    const bool ok = zg.on && chain_raw < n_chains;
    const int chain = chain_raw < n_chains ? chain_raw : n_chains - 1;
    const int s = chain / Bc, b = p.b0 + chain % Bc;
    const unsigned bits = a.set_expert_bits[s];
    float z_own = 0.f;
    for (int i = 0; i < T; ++i) {
      const int t = pass_time(i, T, a.direction);
      float pm, ps;
      if (i == 0) { pm = gm; ps = gs; }
      else {
        float zv[Z], nlv[Z], gate, lin, nl, as, ag[H / Z], an[H / Z];
        gather<Z>(z_own, zg.base, zv);
        zsplit_gtf_forward<Z, H, false>(sP + opaque_zero(), zg.base, j, zv, gate, lin, nl, as, nlv, ag, an);
        poe2_forward(gm, gs, fmaf(gate, nl - lin, lin), softplus_f(as) + p.min_std, pm, ps);
      }
      float mu, sd;
      poe_component(a, bits, s, t, b, j, gm, gs, pm, ps, mu, sd);
      const int64_t o = (((int64_t)s * T + t) * B + b) * Z + j;
      const bool sampled = pass_samples(a, i);
      z_own = sampled ? fmaf(gen::eps_at(a.noise, s, t, b, 0, j, T, B, 1, Z), sd, mu) : mu;
      if (ok) {
        a.infer_mean[o] = mu; a.infer_std[o] = sd;
        a.prior_mean[o] = pm; a.prior_std[o] = ps;
        if (a.samples != nullptr) a.samples[o] = z_own;
        if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b]))
          kl_sum += kld_elem_fast(mu, sd, pm, ps);
      }
    }
  }
  if (a.loss_acc != nullptr && a.kl_weight != 0.f)
    block_reduce_add_double(kl_sum * a.kl_weight, a.loss_acc);
}

// =========================================================================
// backward
// =========================================================================
// Per-lane weight-gradient accumulators: registers (lowest latency; 8 warps/SM at ~230
// registers) or one shared-memory column per lane (element i of lane l at i*32 + l: conflict
// free; 128 registers, 14 warps/SM, so twice as many chains are resident at once).  The
// shared-memory form is a tuning variant (BFVI_ZSPLIT_BWD=2): at C2 it shortens the smoothing
// backward 1.18 -> 0.97 ms stand-alone, but its 214 KB/SM of shared memory keeps the f_mode
// backward pass (side stream) from running beside it and the STEP gets 0.5 ms slower.
template <int N, bool SMEM>
struct LaneAcc;
template <int N>
struct LaneAcc<N, false> {
  float v[N];
  __device__ __forceinline__ explicit LaneAcc(float*) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ void fma(int i, float a, float b) { v[i] = fmaf(a, b, v[i]); }
  __device__ __forceinline__ void add(int i, float a) { v[i] += a; }
  __device__ __forceinline__ float get(int i) const { return v[i]; }
};
template <int N>
struct LaneAcc<N, true> {
  float* col;
  __device__ __forceinline__ explicit LaneAcc(float* warp_base) : col(warp_base + (threadIdx.x & 31)) {
#pragma unroll
    for (int i = 0; i < N; ++i) col[i * 32] = 0.f;
  }
  __device__ __forceinline__ void fma(int i, float a, float b) { col[i * 32] = fmaf(a, b, col[i * 32]); }
  __device__ __forceinline__ void add(int i, float a) { col[i * 32] += a; }
  __device__ __forceinline__ float get(int i) const { return col[i * 32]; }
};
template <int Z, int H>
struct ZAccLayout {
  static constexpr int UPL = H / Z;
  static constexpr int W0G = 0, W2G = W0G + UPL * Z, W0N = W2G + UPL * Z, W2N = W0N + UPL * Z,
                       B0G = W2N + UPL * Z, B0N = B0G + UPL, WL = B0N + UPL, WS = WL + Z, BL = WS + Z, BS = BL + 1,
                       B2G = BS + 1, B2N = B2G + 1, N = B2N + 1;
};

template <int Z, int H, bool ACC_SMEM>
__global__ void __launch_bounds__(kZsplitThreads, ACC_SMEM ? 7 : 1)
zsplit_bwd_kernel(const __grid_constant__ FilterParams p) {
  static_assert(H % Z == 0, "z-split needs H to be a multiple of Z");
  using P = GtfPack<Z, H>;
  using L_ = GtfLayout<Z, H>;
  constexpr int UPL = H / Z;
  using A_ = ZAccLayout<Z, H>;
  __shared__ __align__(16) float sP[P::SIZE];
  __shared__ float sG[L_::SIZE + 2 * Z];            // CTA-level gradient accumulator (+ global prior)
  __shared__ float sAcc[ACC_SMEM ? (kZsplitThreads / 32) * A_::N * 32 : 1];
  const bfvi_filter_args& a = p.a;
  gtf_pack_load<Z, H>(p.trans_w, sP);
  for (int i = threadIdx.x; i < L_::SIZE + 2 * Z; i += blockDim.x) sG[i] = 0.f;
  __syncthreads();
  const ZGroup zg(Z);
  const int j = zg.j;
  const float gm = p.z0_mean[j], gs = expf(p.z0_log_std[j]) + p.min_std;
  const int T = a.T, B = a.B;
  const int Bc = p.bc > 0 ? p.bc : B;
  const int n_chains = a.S * Bc;
  const int n_tasks = (n_chains + zg.cpw - 1) / zg.cpw;
  const int wpb = blockDim.x >> 5;

  // weight-gradient accumulators of this lane's units and rows, alive for the whole kernel
  LaneAcc<A_::N, ACC_SMEM> acc(sAcc + (ACC_SMEM ? (threadIdx.x >> 5) * A_::N * 32 : 0));
  float d_gm = 0.f, d_gs = 0.f;

  for (int task = blockIdx.x * wpb + (threadIdx.x >> 5); task < n_tasks; task += gridDim.x * wpb) {
    const int chain_raw = task * zg.cpw + zg.cig;
    const bool ok = zg.on && chain_raw < n_chains;
    const float vm = ok ? 1.f : 0.f;
    const int chain = chain_raw < n_chains ? chain_raw : n_chains - 1;
    const int s = chain / Bc, b = p.b0 + chain % Bc;
    const unsigned bits = a.set_expert_bits[s];
    float c_mu = 0.f, c_sd = 0.f, eps_cur = 0.f;
    bool have_eps_cur = false;
    float mu_c, sd_c;
    {
      const int64_t o0 = (((int64_t)s * T + pass_time(T - 1, T, a.direction)) * B + b) * Z + j;
      mu_c = a.infer_mean[o0]; sd_c = a.infer_std[o0];
    }
    for (int i = T - 1; i >= 0; --i) {
      const int t = pass_time(i, T, a.direction);
      const int64_t o = (((int64_t)s * T + t) * B + b) * Z + j;
      const float mu = mu_c, sd = sd_c, pm = a.prior_mean[o], ps = a.prior_std[o];
      float d_mu = c_mu + (a.d_infer_mean ? a.d_infer_mean[o] : 0.f);
      float d_sd = c_sd + (a.d_infer_std ? a.d_infer_std[o] : 0.f);
      float d_pm = a.d_prior_mean ? a.d_prior_mean[o] : 0.f;
      float d_ps = a.d_prior_std ? a.d_prior_std[o] : 0.f;
      if (a.d_samples != nullptr) {
        const float ds = a.d_samples[o];
        d_mu += ds;
        if (pass_samples(a, i)) {
          if (!have_eps_cur) eps_cur = gen::eps_at(a.noise, s, t, b, 0, j, T, B, 1, Z);
          d_sd = fmaf(ds, eps_cur, d_sd);
        }
      }
      if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b])) {
        float g1, g2, g3, g4;
        kld_elem_grad(mu, sd, pm, ps, a.kl_weight, g1, g2, g3, g4);
        d_mu += g1; d_sd += g2; d_pm += g3; d_ps += g4;
      }
      // product of experts backward, component j (models/dgts.py:40-51)
      {
        const float inv_s = sd * sd;
        const float d_n = d_mu * inv_s;
        const float d_s = -d_mu * mu * inv_s - 0.5f * d_sd * sd * inv_s;
        float tp, dtp;
        poe_prec_bwd(ps, tp, dtp);
        d_pm += d_n * tp;
        d_ps += (d_n * pm + d_s) * dtp;
        for (int e = 0; e < a.n_experts; ++e) {
          if (!((bits >> e) & 1u)) continue;
          const bfvi_expert& ex = a.experts[e];
          bool m = true;
          if (ex.mask != nullptr) m = ex.mask[s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b] != 0;
          if (ex.zero_mask_last_t && t == T - 1) m = false;
          if (!m) continue;
          if (ex.kind == BFVI_EXPERT_INV_PRIOR) {
            float te, dte;
            poe_prec_bwd(-gs, te, dte);
            d_gm += vm * d_n * te;
            d_gs -= vm * (d_n * gm + d_s) * dte;
          } else if (ex.d_mean != nullptr && ok) {
            const int64_t off = s * ex.stride_s + t * ex.stride_t + b * ex.stride_b + j;
            const float mean = ex.mean[off];
            float te, dte;
            poe_prec_bwd(ex.std[off], te, dte);
            atomicAdd(ex.d_mean + off, d_n * te);
            atomicAdd(ex.d_std + off, (d_n * mean + d_s) * dte);
          }
        }
      }
      if (i == 0) { d_gm += vm * d_pm; d_gs += vm * d_ps; continue; }
      // ---- transition from step i-1 (single particle: prior = product with the global prior) ----
      const int t_prev = pass_time(i - 1, T, a.direction);
      const int64_t op = (((int64_t)s * T + t_prev) * B + b) * Z + j;
      const float mu_p = a.infer_mean[op], sd_p = a.infer_std[op];
      mu_c = mu_p; sd_c = sd_p;
      const bool sampled_prev = pass_samples(a, i - 1);
      const float eps = sampled_prev ? gen::eps_at(a.noise, s, t_prev, b, 0, j, T, B, 1, Z) : 0.f;
      const float z_own = fmaf(eps, sd_p, mu_p);
      float zv[Z], nlv[Z], gate, lin, nl, as, ag[UPL], an[UPL];
      gather<Z>(z_own, zg.base, zv);
      const float* sPw = sP + opaque_zero();
      zsplit_gtf_forward<Z, H, true>(sPw, zg.base, j, zv, gate, lin, nl, as, nlv, ag, an);
      const float qm = fmaf(gate, nl - lin, lin), qs = softplus_f(as) + p.min_std;
      float m_k, s_k, g_gm, g_gs, d_qm, d_qs;
      poe2_forward(gm, gs, qm, qs, m_k, s_k);
      poe2_backward(gm, gs, qm, qs, m_k, s_k, d_pm * vm, d_ps * vm, g_gm, g_gs, d_qm, d_qs);
      d_gm += g_gm; d_gs += g_gs;
      const float d_as = d_qs * softplus_grad(as);
      float d_nl = d_qm * gate;
      const float d_lin = d_qm - d_nl;
      const float d_ag = d_lin * gate * (nl - lin);
      float d_asv[Z], d_agv[Z], d_nlv[Z], d_linv[Z];
      gather<Z>(d_as, zg.base, d_asv);
      gather<Z>(d_lin, zg.base, d_linv);
      float dz = 0.f;
#pragma unroll
      for (int o = 0; o < Z; ++o) {                        // column j of z_to_std / z_lin
        d_nl = fmaf(sPw[P::STD + o * P::RW + 1 + j], d_asv[o], d_nl);
        dz = fmaf(sPw[P::LIN + o * P::RW + 1 + j], d_linv[o], dz);
      }
      gather<Z>(d_ag, zg.base, d_agv);
      gather<Z>(d_nl, zg.base, d_nlv);
      float pz[Z];
#pragma unroll
      for (int i2 = 0; i2 < Z; ++i2) pz[i2] = 0.f;
#pragma unroll
      for (int u = 0; u < UPL; ++u) {
        const int h = j * UPL + u;
        float wg[P::U], wn[P::U];
        lds_vec<P::U>(sPw + P::GATE + h * P::U, wg);
        lds_vec<P::U>(sPw + P::NONLIN + h * P::U, wn);
        float da = 0.f, dc = 0.f;
#pragma unroll
        for (int o = 0; o < Z; ++o) { da = fmaf(wg[1 + Z + o], d_agv[o], da); dc = fmaf(wn[1 + Z + o], d_nlv[o], dc); }
        da = ag[u] > 0.f ? da : 0.f;
        dc = an[u] > 0.f ? dc : 0.f;
        acc.add(A_::B0G + u, da); acc.add(A_::B0N + u, dc);
#pragma unroll
        for (int i2 = 0; i2 < Z; ++i2) {
          pz[i2] = fmaf(wg[1 + i2], da, pz[i2]);
          pz[i2] = fmaf(wn[1 + i2], dc, pz[i2]);
          acc.fma(A_::W0G + u * Z + i2, da, zv[i2]);
          acc.fma(A_::W0N + u * Z + i2, dc, zv[i2]);
          acc.fma(A_::W2G + u * Z + i2, d_agv[i2], ag[u]);
          acc.fma(A_::W2N + u * Z + i2, d_nlv[i2], an[u]);
        }
      }
      dz += reduce_scatter<Z>(pz, zg.base, j);
#pragma unroll
      for (int i2 = 0; i2 < Z; ++i2) {
        acc.fma(A_::WL + i2, d_lin, zv[i2]);
        acc.fma(A_::WS + i2, d_as, nlv[i2]);
      }
      acc.add(A_::BL, d_lin); acc.add(A_::BS, d_as); acc.add(A_::B2G, d_ag); acc.add(A_::B2N, d_nl);
      c_mu = dz; c_sd = dz * eps;
      eps_cur = eps; have_eps_cur = sampled_prev;
    }
  }
  // ---- flush: lane accumulators -> CTA accumulator (shared atomics) -> global ------------------
  if (zg.on) {
#pragma unroll
    for (int u = 0; u < UPL; ++u) {
      const int h = j * UPL + u;
      atomicAdd(&sG[L_::G0B + h], acc.get(A_::B0G + u)); atomicAdd(&sG[L_::N0B + h], acc.get(A_::B0N + u));
#pragma unroll
      for (int i = 0; i < Z; ++i) {
        atomicAdd(&sG[L_::G0W + h * Z + i], acc.get(A_::W0G + u * Z + i));
        atomicAdd(&sG[L_::N0W + h * Z + i], acc.get(A_::W0N + u * Z + i));
        atomicAdd(&sG[L_::G2W + i * H + h], acc.get(A_::W2G + u * Z + i));
        atomicAdd(&sG[L_::N2W + i * H + h], acc.get(A_::W2N + u * Z + i));
      }
    }
#pragma unroll
    for (int i = 0; i < Z; ++i) {
      atomicAdd(&sG[L_::LW + j * Z + i], acc.get(A_::WL + i));
      atomicAdd(&sG[L_::SW + j * Z + i], acc.get(A_::WS + i));
    }
    atomicAdd(&sG[L_::LB + j], acc.get(A_::BL)); atomicAdd(&sG[L_::SB + j], acc.get(A_::BS));
    atomicAdd(&sG[L_::G2B + j], acc.get(A_::B2G)); atomicAdd(&sG[L_::N2B + j], acc.get(A_::B2N));
    atomicAdd(&sG[L_::SIZE + j], d_gm); atomicAdd(&sG[L_::SIZE + Z + j], d_gs);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L_::SIZE; i += blockDim.x)
    if (sG[i] != 0.f) atomicAdd(p.g_trans + i, sG[i]);
  if (threadIdx.x < Z) {
    if (sG[L_::SIZE + threadIdx.x] != 0.f) atomicAdd(p.g_z0_mean + threadIdx.x, sG[L_::SIZE + threadIdx.x]);
    const float v = sG[L_::SIZE + Z + threadIdx.x] * expf(p.z0_log_std[threadIdx.x]);   // gs = exp(log_std) + min_std
    if (v != 0.f) atomicAdd(p.g_z0_log_std + threadIdx.x, v);
  }
}

}  // namespace bfvi
