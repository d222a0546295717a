// bfvi_chain.cuh — the time-recurrent kernels of the small-dim path: z_filter
// forward / backward (MultiDMM.z_filter + z_next, models/dmm.py:214-258,319-412) and
// the prior-matching term (kld_prior, models/dmm.py:496-501).
//
// Thread mapping ("lane groups"): a chain (one sequence of one chain set) owns L
// consecutive lanes of a warp and each lane owns R particles, so a warp walks
// 32/L chains through all T steps on-chip.  K = 25 particles -> L = 5, R = 5: 30 of
// 32 lanes busy and every shared-memory weight load feeds 5 rows.  K = 1 -> L = 1:
// 32 sequences per warp.  The mixture moment-matching over particles is an
// in-thread sum over R followed by an L-lane shuffle sum; the product of experts of
// a step is computed redundantly by the L lanes (it is tiny) and written by lane 0
// of the group.
//
// The backward kernel walks the same chains in reverse, one row slice (32 rows, one
// per lane) at a time: it regenerates the particle from the saved (infer mean, std)
// and the counter-based / external noise, recomputes the GTF, back-propagates, and
// stages the slice's layer inputs and pre-activation gradients in per-warp panels
// from which each lane accumulates a register tile of the weight gradient
// (bfvi_wgrad.cuh).
#pragma once
#include "../../include/bfvi.h"
#include "bfvi_math.cuh"
#include "bfvi_rng.cuh"
#include "bfvi_wgrad.cuh"

namespace bfvi {

constexpr int kTX = 4;                       // weight-gradient tile width
// launch geometry (overridable at build time for tuning runs)
#ifndef BFVI_FWD_THREADS
#define BFVI_FWD_THREADS 128
#endif
#ifndef BFVI_FWD_MINB
#define BFVI_FWD_MINB 4          // resident CTAs per SM the R = 5 forward kernel is compiled for
#endif
#ifndef BFVI_BWD_WARPS
#define BFVI_BWD_WARPS 4
#endif
#ifndef BFVI_BWD_MINB
#define BFVI_BWD_MINB 2
#endif
constexpr int kChainFwdThreads = BFVI_FWD_THREADS;
constexpr int kChainBwdWarps = BFVI_BWD_WARPS;

struct FilterParams {
  bfvi_filter_args a;
  const float* trans_w;      // flat GTF block of the pass direction
  const float* z0_mean;
  const float* z0_log_std;
  float* g_trans;            // gradient block of the same GTF   (backward only)
  float* g_z0_mean;
  float* g_z0_log_std;
  float min_std;
  int lanes;                 // L: lanes per chain (1..32)
  int rounds;                // forward: ceil(K / (L * R))
  int slices;                // backward: ceil(K / L)
  int b0, bc;                // batch chunk [b0, b0 + bc) served by this launch (bc = 0: all of B)
  // backward, K > 1: the time loop cut into seg_count segments that are scheduled as separate
  // work items (task, segment) — see chain_bwd_kernel.  This launch serves [seg_lo, seg_hi).
  int seg_count, seg_lo, seg_hi;
  int* seg_done;             // per task: segments completed (several segments per launch only)
  float* seg_carry;          // (S, B, 2Z): gradient carried into the next (earlier) segment
};

// -------------------------------------------------------------------------
// small helpers
// -------------------------------------------------------------------------
__device__ __forceinline__ int pass_time(int i, int T, int direction) {
  return direction == BFVI_DIR_BWD ? T - 1 - i : i;
}
__device__ __forceinline__ bool pass_samples(const bfvi_filter_args& a, int i) {
  return a.sample || a.n_particles > 1 || (i == 0 && a.sample_init);   // models/dmm.py:398
}

__device__ inline void block_reduce_add_double(float v, double* target) {
  __shared__ float red[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0 && target != nullptr && t != 0.f) atomicAdd(target, (double)t);
  }
  __syncthreads();
}

// product of experts at (s, t, b): prior first, then the chain set's experts, in
// expert order (models/dmm.py:388-395 / models/dgts.py:40-51).  IEEE-rounded ops.
template <int Z>
__device__ __forceinline__ void poe_step_forward(const bfvi_filter_args& a, unsigned bits, int s, int t,
                                                 int b, const float* gm, const float* gs,
                                                 const float (&pm)[Z], const float (&ps)[Z],
                                                 float (&mu)[Z], float (&sd)[Z]) {
  float N[Z], S[Z];
#pragma unroll
  for (int i = 0; i < Z; ++i) {
    const float tp = poe_prec(ps[i]);
    S[i] = tp;
    N[i] = __fmul_rn(pm[i], tp);
  }
  for (int e = 0; e < a.n_experts; ++e) {
    if (!((bits >> e) & 1u)) continue;
    const bfvi_expert& ex = a.experts[e];
    bool m = true;
    if (ex.mask != nullptr) m = ex.mask[s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b] != 0;
    if (ex.zero_mask_last_t && t == a.T - 1) m = false;
    const float w = m ? 1.f : 0.f;
    const float* pmean = ex.mean + s * ex.stride_s + t * ex.stride_t + b * ex.stride_b;
    const float* pstd = ex.std + s * ex.stride_s + t * ex.stride_t + b * ex.stride_b;
#pragma unroll
    for (int i = 0; i < Z; ++i) {
      float mean, std;
      if (ex.kind == BFVI_EXPERT_INV_PRIOR) { mean = gm[i]; std = -gs[i]; }   // models/dmm.py:476-477
      else { mean = pmean[i]; std = pstd[i]; }
      // products are rounded before they are summed (no FMA contraction), like the
      // reference's `sum(mean * T)`: when the inverse-prior expert cancels the prior
      // exactly (0/0 -> NaN -> 0, models/dgts.py:48-49) the cancellation must be exact
      const float te = __fmul_rn(poe_prec(std), w);
      S[i] = __fadd_rn(S[i], te);
      N[i] = __fadd_rn(N[i], __fmul_rn(__fmul_rn(mean, w), te));
    }
  }
#pragma unroll
  for (int i = 0; i < Z; ++i) {
    const float m = __fdiv_rn(N[i], S[i]);
    mu[i] = (m != m) ? 0.f : m;                // models/dgts.py:49
    sd[i] = __fsqrt_rn(__fdiv_rn(1.f, S[i]));
  }
}

// backward of the above.  d_mu / d_sd in; prior gradients accumulate into
// d_pm / d_ps; expert gradients are scattered (atomic) when `emit`; the inverse
// prior expert feeds d_gm / d_gs.
template <int Z>
__device__ __forceinline__ void poe_step_backward(const bfvi_filter_args& a, unsigned bits, int s, int t,
                                                  int b, const float* gm, const float* gs,
                                                  const float (&pm)[Z], const float (&ps)[Z],
                                                  const float (&mu)[Z], const float (&sd)[Z],
                                                  const float (&d_mu)[Z], const float (&d_sd)[Z],
                                                  float (&d_pm)[Z], float (&d_ps)[Z],
                                                  float (&d_gm)[Z], float (&d_gs)[Z], bool emit) {
  float d_n[Z], d_s[Z];
#pragma unroll
  for (int i = 0; i < Z; ++i) {
    const float inv_s = sd[i] * sd[i];          // 1 / sum of precisions
    d_n[i] = d_mu[i] * inv_s;
    d_s[i] = -d_mu[i] * mu[i] * inv_s - 0.5f * d_sd[i] * sd[i] * inv_s;
    float tp, dtp;
    poe_prec_bwd(ps[i], tp, dtp);
    d_pm[i] += d_n[i] * tp;
    d_ps[i] += (d_n[i] * pm[i] + d_s[i]) * dtp;
  }
  for (int e = 0; e < a.n_experts; ++e) {
    if (!((bits >> e) & 1u)) continue;
    const bfvi_expert& ex = a.experts[e];
    bool m = true;
    if (ex.mask != nullptr) m = ex.mask[s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b] != 0;
    if (ex.zero_mask_last_t && t == a.T - 1) m = false;
    if (!m) continue;
    const int64_t off = s * ex.stride_s + t * ex.stride_t + b * ex.stride_b;
    if (ex.kind == BFVI_EXPERT_INV_PRIOR) {
#pragma unroll
      for (int i = 0; i < Z; ++i) {
        if (emit) {
          float te, dte;
          poe_prec_bwd(-gs[i], te, dte);
          d_gm[i] += d_n[i] * te;
          d_gs[i] -= (d_n[i] * gm[i] + d_s[i]) * dte;
        }
      }
    } else if (ex.d_mean != nullptr && emit) {
#pragma unroll
      for (int i = 0; i < Z; ++i) {
        const float mean = ex.mean[off + i];
        float te, dte;
        poe_prec_bwd(ex.std[off + i], te, dte);
        atomicAdd(ex.d_mean + off + i, d_n[i] * te);
        atomicAdd(ex.d_std + off + i, (d_n[i] * mean + d_s[i]) * dte);
      }
    }
  }
}

// Warm L1 with everything the product of experts of (s, t, b) will read, one time
// step ahead of its use: the single-particle passes are latency-bound and their
// expert loads sit on the critical path of every step.
template <int Z>
__device__ __forceinline__ void poe_step_prefetch(const bfvi_filter_args& a, unsigned bits, int s, int t,
                                                  int b) {
  for (int e = 0; e < a.n_experts; ++e) {
    if (!((bits >> e) & 1u)) continue;
    const bfvi_expert& ex = a.experts[e];
    if (ex.kind != BFVI_EXPERT_TENSOR) continue;
    if (ex.mask != nullptr) prefetch_l1(ex.mask + s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b);
    const int64_t off = s * ex.stride_s + t * ex.stride_t + b * ex.stride_b;
    prefetch_l1(ex.mean + off); prefetch_l1(ex.mean + off + Z - 1);
    prefetch_l1(ex.std + off); prefetch_l1(ex.std + off + Z - 1);
  }
}

// lane-group geometry of one warp
struct LaneGroup {
  int L, cpw, cig, lig, base;
  bool on;                      // this lane belongs to a chain slot of the warp task
  __device__ __forceinline__ explicit LaneGroup(int lanes) {
    const int lane = threadIdx.x & 31;
    L = lanes;
    cpw = 32 / L;
    cig = lane / L;
    lig = lane - cig * L;
    on = cig < cpw;
    base = on ? cig * L : 0;    // idle tail lanes shadow group 0 (their rows are never valid)
  }
};

// =========================================================================
// z_filter forward
// =========================================================================
template <int Z, int H, int R>
__global__ void __launch_bounds__(kChainFwdThreads, R == 1 ? (512 / kChainFwdThreads) : BFVI_FWD_MINB)
chain_fwd_kernel(const __grid_constant__ FilterParams p) {
  using P = GtfPack<Z, H>;
  __shared__ __align__(16) float sP[P::SIZE];
  __shared__ float sGm[Z], sGs[Z];
  const bfvi_filter_args& a = p.a;
  gtf_pack_load<Z, H>(p.trans_w, sP);
  if (threadIdx.x < Z) {
    sGm[threadIdx.x] = p.z0_mean[threadIdx.x];
    sGs[threadIdx.x] = expf(p.z0_log_std[threadIdx.x]) + p.min_std;       // models/dmm.py:126-127
  }
  __syncthreads();

  const LaneGroup lg(p.lanes);
  const int L = lg.L, rounds = p.rounds;
  const int T = a.T, B = a.B, K = a.n_particles;
  const int Bc = p.bc > 0 ? p.bc : B;
  const int n_chains = a.S * Bc;
  const int n_tasks = (n_chains + lg.cpw - 1) / lg.cpw;
  const int wpb = blockDim.x >> 5;
  const float inv_k = 1.f / (float)K;
  float kl_sum = 0.f;

  for (int task = blockIdx.x * wpb + (threadIdx.x >> 5); task < n_tasks; task += gridDim.x * wpb) {
    const int chain_raw = task * lg.cpw + lg.cig;
    const bool chain_ok = lg.on && chain_raw < n_chains;
    const int chain = chain_raw < n_chains ? chain_raw : n_chains - 1;
    const int s = chain / Bc, b = p.b0 + chain % Bc;
    const unsigned bits = a.set_expert_bits[s];
    const bool writer = chain_ok && lg.lig == 0;
    float mu_p[Z], sd_p[Z];
    float zc[R][Z];                                   // particles of the previous step (rounds == 1)
    int t_prev = 0;
#pragma unroll
    for (int j = 0; j < Z; ++j) mu_p[j] = sd_p[j] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < Z; ++j) zc[r][j] = 0.f;

    for (int i = 0; i < T; ++i) {
      const int t = pass_time(i, T, a.direction);
      if (R == 1 && i + 1 < T) poe_step_prefetch<Z>(a, bits, s, pass_time(i + 1, T, a.direction), b);
      float pm[Z], ps[Z];
      if (i == 0) {
#pragma unroll
        for (int j = 0; j < Z; ++j) { pm[j] = sGm[j]; ps[j] = sGs[j]; }
      } else {
        float sm[Z], sv[Z], sq[Z];
#pragma unroll
        for (int j = 0; j < Z; ++j) sm[j] = sv[j] = sq[j] = 0.f;
        const bool sampled_prev = pass_samples(a, i - 1);
        for (int round = 0; round < rounds; ++round) {
          float z[R][Z];
          if (rounds == 1) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int j = 0; j < Z; ++j) z[r][j] = zc[r][j];
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const int k = (round * R + r) * L + lg.lig;
              float eps[Z];
              load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t_prev, b, a.noise.b_offset,
                          k < K ? k : K - 1, T, B, K, eps);
#pragma unroll
              for (int j = 0; j < Z; ++j) z[r][j] = sampled_prev ? fmaf(eps[j], sd_p[j], mu_p[j]) : mu_p[j];
            }
          }
          // opaque offset: keeps the compiler from hoisting the (time-loop invariant)
          // weight loads out of the loop into registers it then has to spill
          const float* sPw = sP + opaque_zero();
          gtf_rows_forward<Z, H, R>(sPw, p.min_std, z, [&](int r, int j, float qm, float qs) {
            const int k = (round * R + r) * L + lg.lig;
            float m_k, s_k;
            poe2_forward(sGm[j], sGs[j], qm, qs, m_k, s_k);
            if (K == 1) {
              if (k == 0) { sm[j] = m_k; sv[j] = s_k; }
            } else {
              const float on = k < K ? 1.f : 0.f;
              sm[j] = fmaf(on, m_k, sm[j]);
              sv[j] = fmaf(on * s_k, s_k, sv[j]);
              sq[j] = fmaf(on * m_k, m_k, sq[j]);
            }
          });
        }
        if (K == 1) {
#pragma unroll
          for (int j = 0; j < Z; ++j) { pm[j] = sm[j]; ps[j] = sv[j]; }
        } else {
          float mom[3 * Z];
#pragma unroll
          for (int j = 0; j < Z; ++j) { mom[j] = sm[j]; mom[Z + j] = sv[j]; mom[2 * Z + j] = sq[j]; }
          group_sum_vec<3 * Z>(mom, lg.base, L);
#pragma unroll
          for (int j = 0; j < Z; ++j) {                     // models/dgts.py:78-83
            const float m = mom[j] * inv_k;
            const float v = mom[Z + j] * inv_k + (mom[2 * Z + j] * inv_k - m * m);
            pm[j] = m; ps[j] = sqrtf(v);
          }
        }
      }
      float mu[Z], sd[Z];
      poe_step_forward<Z>(a, bits, s, t, b, sGm, sGs, pm, ps, mu, sd);
      const int64_t o = (((int64_t)s * T + t) * B + b) * Z;
      if (writer) {
#pragma unroll
        for (int j = 0; j < Z; ++j) {
          a.infer_mean[o + j] = mu[j]; a.infer_std[o + j] = sd[j];
          a.prior_mean[o + j] = pm[j]; a.prior_std[o + j] = ps[j];
        }
        if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b])) {
#pragma unroll
          for (int j = 0; j < Z; ++j) kl_sum += kld_elem_fast(mu[j], sd[j], pm[j], ps[j]);
        }
      }
      // ---- particles of this step: next step's GTF input and the `samples` output ----
      const bool sampled = pass_samples(a, i);
      if (rounds == 1) {
        if (i + 1 < T || a.samples != nullptr) {
          float se[Z];
#pragma unroll
          for (int j = 0; j < Z; ++j) se[j] = 0.f;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int k = r * L + lg.lig;
            if (sampled) {
              float eps[Z];
              load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t, b, a.noise.b_offset,
                          k < K ? k : K - 1, T, B, K, eps);
#pragma unroll
              for (int j = 0; j < Z; ++j) zc[r][j] = fmaf(eps[j], sd[j], mu[j]);
            } else {
#pragma unroll
              for (int j = 0; j < Z; ++j) zc[r][j] = mu[j];
            }
            if (k < K) {
#pragma unroll
              for (int j = 0; j < Z; ++j) se[j] += zc[r][j];
            }
          }
          if (a.samples != nullptr) {                       // mean over particles, models/dmm.py:402
            if (K > 1) {
#pragma unroll
              for (int j = 0; j < Z; ++j) se[j] = group_sum(se[j], lg.base, L) * inv_k;
            }
            if (writer) {
#pragma unroll
              for (int j = 0; j < Z; ++j) a.samples[o + j] = se[j];
            }
          }
        }
      } else if (a.samples != nullptr) {
        float se[Z];
#pragma unroll
        for (int j = 0; j < Z; ++j) se[j] = 0.f;
        for (int k = lg.lig; k < K; k += L) {
          float eps[Z];
          load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t, b, a.noise.b_offset, k, T, B, K,
                      eps);
#pragma unroll
          for (int j = 0; j < Z; ++j) se[j] += fmaf(eps[j], sd[j], mu[j]);
        }
#pragma unroll
        for (int j = 0; j < Z; ++j) se[j] = group_sum(se[j], lg.base, L) * inv_k;
        if (writer) {
#pragma unroll
          for (int j = 0; j < Z; ++j) a.samples[o + j] = se[j];
        }
      }
#pragma unroll
      for (int j = 0; j < Z; ++j) { mu_p[j] = mu[j]; sd_p[j] = sd[j]; }
      t_prev = t;
    }
  }
  if (a.loss_acc != nullptr && a.kl_weight != 0.f)
    block_reduce_add_double(kl_sum * a.kl_weight, a.loss_acc);
}

// =========================================================================
// z_filter backward (reverse pass order; GTF activations recomputed from the
// saved infer (mu, sd) and the regenerated / external noise)
// =========================================================================
template <int Z, int H>
struct GtfPanels {
  using C = GtfCols<Z, H>;
  static constexpr int TD = Z <= 8 ? Z : 8;
  __host__ __device__ static WgSpec spec() {
    using L = GtfLayout<Z, H>;
    WgSpec s;
    s.n_blocks = 6;
    s.blk[0] = WgBlock{C::DA1, H, C::XZ, 1 + Z, L::G0W, L::G0B};
    s.blk[1] = WgBlock{C::DA3, H, C::XZ, 1 + Z, L::N0W, L::N0B};
    s.blk[2] = WgBlock{C::DLIN, Z, C::XZ, 1 + Z, L::LW, L::LB};
    s.blk[3] = WgBlock{C::DAG, Z, C::XH1, 1 + H, L::G2W, L::G2B};
    s.blk[4] = WgBlock{C::DNL, Z, C::XH3, 1 + H, L::N2W, L::N2B};
    s.blk[5] = WgBlock{C::DAS, Z, C::XNL, 1 + Z, L::SW, L::SB};
    return s;
  }
  // floats of dynamic shared memory per warp: panels + accumulator (+32 dump slots)
  static constexpr int WARP_FLOATS = (C::NXC + C::NDC) * kRS + GtfLayout<Z, H>::SIZE + 32;
};

// One row slice of the transition backward: regenerate -> GTF forward -> mixture /
// PoE backward -> GTF backward + staging -> register-tile weight gradient.
// d_pm / d_v: gradient at the prior mean and 0.5 * d_ps / ps.
template <int Z, int H>
__device__ __forceinline__ void transition_slice_backward(
    const float* __restrict__ sP, const float* __restrict__ sGm, const float* __restrict__ sGs,
    float min_std, float inv_k, const float (&z)[Z], const float (&pm)[Z], const float (&d_pm)[Z],
    const float (&d_v)[Z], bool valid, float (&d_gm)[Z], float (&d_gs)[Z], float (&dz)[Z],
    float* __restrict__ Xp, float* __restrict__ Dp, int lane,
    float2 (&acc)[GtfPanels<Z, H>::TD][kTX], const int4 task) {
  float g[Z], lin[Z], nl[Z], as[Z], d_qm[Z], d_qs[Z];
  __syncwarp();                        // the previous slice's tile reads are done
  gtf_row_forward_stage<Z, H>(sP, z, g, lin, nl, as, Xp, lane);
#pragma unroll
  for (int j = 0; j < Z; ++j) {
    const float qs = softplus_f(as[j]) + min_std;
    const float qm = fmaf(g[j], nl[j] - lin[j], lin[j]);
    float m_k, s_k, g_gm, g_gs;
    poe2_forward(sGm[j], sGs[j], qm, qs, m_k, s_k);
    const float d_mk = (d_pm[j] + 2.f * d_v[j] * (m_k - pm[j])) * inv_k;   // models/dgts.py:78-83
    const float d_sk = 2.f * d_v[j] * s_k * inv_k;
    poe2_backward(sGm[j], sGs[j], qm, qs, m_k, s_k, d_mk, d_sk, g_gm, g_gs, d_qm[j], d_qs[j]);
    if (valid) { d_gm[j] += g_gm; d_gs[j] += g_gs; }
  }
  gtf_row_backward_stage<Z, H>(sP, g, lin, nl, as, d_qm, d_qs, dz, Xp, Dp, lane, valid);
  __syncwarp();
  wg_tile_fma<GtfPanels<Z, H>::TD, kTX>(acc, Dp, Xp, task);
}

// SEG = false compiles the time segmentation out (single-particle and small-batch launches).
template <int Z, int H, bool SEG>
__global__ void __launch_bounds__(kChainBwdWarps * 32, BFVI_BWD_MINB)
chain_bwd_kernel(const __grid_constant__ FilterParams p) {
  using L_ = GtfLayout<Z, H>;
  using P = GtfPack<Z, H>;
  using PN = GtfPanels<Z, H>;
  using C = GtfCols<Z, H>;
  constexpr int TD = PN::TD;
  BFVI_DYN_SMEM(float, smem);
  __shared__ __align__(16) float sP[P::SIZE];
  __shared__ float sGm[Z], sGs[Z];
  __shared__ float sRed[kChainBwdWarps][2 * Z];
  const bfvi_filter_args& a = p.a;
  const WgSpec spec = PN::spec();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;

  float* Xp = smem + (size_t)warp * PN::WARP_FLOATS;
  float* Dp = Xp + C::NXC * kRS;
  float* G = Dp + C::NDC * kRS;
  int* tasks = reinterpret_cast<int*>(smem + (size_t)n_warps * PN::WARP_FLOATS);
  int* oidx = tasks + 32 * 4;

  gtf_pack_load<Z, H>(p.trans_w, sP);
  if (threadIdx.x < Z) {
    sGm[threadIdx.x] = p.z0_mean[threadIdx.x];
    sGs[threadIdx.x] = expf(p.z0_log_std[threadIdx.x]) + p.min_std;
  }
  wg_build_tables<TD, kTX>(spec, tasks, oidx, 1, L_::SIZE);
  for (int i = lane; i < L_::SIZE + 32; i += 32) G[i] = 0.f;
  for (int i = lane; i < (C::NXC + C::NDC) * kRS; i += 32) Xp[i] = 0.f;
  __syncwarp();
  Xp[panel_at(C::XZ, lane)] = 1.f;
  Xp[panel_at(C::XH1, lane)] = 1.f;
  Xp[panel_at(C::XH3, lane)] = 1.f;
  Xp[panel_at(C::XNL, lane)] = 1.f;
  __syncthreads();
  const int4 task = reinterpret_cast<const int4*>(tasks)[lane];
  float2 acc[TD][kTX];
#pragma unroll
  for (int i = 0; i < TD; ++i)
#pragma unroll
    for (int j = 0; j < kTX; ++j) acc[i][j] = float2{0.f, 0.f};

  const LaneGroup lg(p.lanes);
  const int L = lg.L, slices = p.slices;
  const int T = a.T, B = a.B, K = a.n_particles;
  const int Bc = p.bc > 0 ? p.bc : B;
  const int n_chains = a.S * Bc;
  const int n_tasks = (n_chains + lg.cpw - 1) / lg.cpw;
  const float inv_k = 1.f / (float)K;
  float d_gm[Z], d_gs[Z];                 // global-prior gradient, per thread
#pragma unroll
  for (int j = 0; j < Z; ++j) d_gm[j] = d_gs[j] = 0.f;

  // Work items.  A warp task (32 / L chains, all T steps) is long — 2.6 ms at C2 — and a B200 holds
  // 1184 such warps, so 2048 tasks take two full rounds with the second 27 % empty.  Cutting the
  // time loop into segments makes the items short enough to pack the machine: item q = (segment,
  // task), segment-major, dealt round-robin to the resident warps; (seg, task) starts from the
  // gradient carry (c_mu, c_sd) that (seg - 1, task) left in seg_carry.  When one launch serves
  // several segments (cooperative launch: every warp resident) the hand-over is a per-task flag.
  const int n_seg = SEG && p.seg_count > 1 ? p.seg_count : 1;
  const int seg_lo = SEG && n_seg > 1 ? p.seg_lo : 0, seg_n = SEG && n_seg > 1 ? p.seg_hi - p.seg_lo : 1;
  const bool seg_flags = SEG && seg_n > 1;
  const int n_items = n_tasks * seg_n;
  const int gwarp = blockIdx.x * n_warps + warp, total_warps = gridDim.x * n_warps;
  for (int q = gwarp; q < n_items; q += total_warps) {
    const int seg = seg_lo + q / n_tasks, wt = q - (q / n_tasks) * n_tasks;
    const int i_hi = T - 1 - (int)(((int64_t)seg * T) / n_seg);
    const int i_lo = seg + 1 < n_seg ? T - (int)(((int64_t)(seg + 1) * T) / n_seg) : 0;
    const int chain_raw = wt * lg.cpw + lg.cig;
    const bool chain_ok = lg.on && chain_raw < n_chains;
    const int chain = chain_raw < n_chains ? chain_raw : n_chains - 1;
    const int s = chain / Bc, b = p.b0 + chain % Bc;
    const unsigned bits = a.set_expert_bits[s];
    const bool emit = chain_ok && lg.lig == 0;        // one writer per chain
    float c_mu[Z], c_sd[Z], eps_cur[Z];
#pragma unroll
    for (int j = 0; j < Z; ++j) c_mu[j] = c_sd[j] = eps_cur[j] = 0.f;
    bool have_eps_cur = false;
    if (SEG && seg > 0) {
      if (seg_flags) {
        if (lane == 0) seg_wait(p.seg_done + wt, seg);
        __syncwarp();
      }
      const float* cr = p.seg_carry + ((int64_t)s * B + b) * 2 * Z;
#pragma unroll
      for (int j = 0; j < Z; ++j) { c_mu[j] = ld_cg(cr + j); c_sd[j] = ld_cg(cr + Z + j); }
    }

    float mu_c[Z], sd_c[Z];               // infer (mean, std) of the step being processed
    {
      const int64_t o0 = (((int64_t)s * T + pass_time(i_hi, T, a.direction)) * B + b) * Z;
#pragma unroll
      for (int j = 0; j < Z; ++j) { mu_c[j] = a.infer_mean[o0 + j]; sd_c[j] = a.infer_std[o0 + j]; }
    }
    for (int i = i_hi; i >= i_lo; --i) {
      const int t = pass_time(i, T, a.direction);
      const int64_t o = (((int64_t)s * T + t) * B + b) * Z;
      if (K == 1 && i > 0) {              // warm L1 for the next (earlier) step of this chain
        const int tn = pass_time(i - 1, T, a.direction);
        const int64_t on = (((int64_t)s * T + tn) * B + b) * Z;
        prefetch_l1(a.prior_mean + on); prefetch_l1(a.prior_mean + on + Z - 1);
        prefetch_l1(a.prior_std + on); prefetch_l1(a.prior_std + on + Z - 1);
        if (a.d_samples) { prefetch_l1(a.d_samples + on); prefetch_l1(a.d_samples + on + Z - 1); }
        if (a.d_prior_mean) { prefetch_l1(a.d_prior_mean + on); prefetch_l1(a.d_prior_std + on); }
        if (i > 1) {
          const int64_t o2 = (((int64_t)s * T + pass_time(i - 2, T, a.direction)) * B + b) * Z;
          prefetch_l1(a.infer_mean + o2); prefetch_l1(a.infer_mean + o2 + Z - 1);
          prefetch_l1(a.infer_std + o2); prefetch_l1(a.infer_std + o2 + Z - 1);
        }
        poe_step_prefetch<Z>(a, bits, s, tn, b);
      }
      float mu[Z], sd[Z], pm[Z], ps[Z], d_mu[Z], d_sd[Z], d_pm[Z], d_ps[Z];
#pragma unroll
      for (int j = 0; j < Z; ++j) {
        mu[j] = mu_c[j]; sd[j] = sd_c[j];
        pm[j] = a.prior_mean[o + j]; ps[j] = a.prior_std[o + j];
        d_mu[j] = c_mu[j] + (a.d_infer_mean ? a.d_infer_mean[o + j] : 0.f);
        d_sd[j] = c_sd[j] + (a.d_infer_std ? a.d_infer_std[o + j] : 0.f);
        d_pm[j] = a.d_prior_mean ? a.d_prior_mean[o + j] : 0.f;
        d_ps[j] = a.d_prior_std ? a.d_prior_std[o + j] : 0.f;
      }
      // --- gradient arriving through `samples` (mean over particles of z_t) ----
      if (a.d_samples != nullptr) {
        if (pass_samples(a, i)) {
          float me[Z];
          if (K > 1) {
#pragma unroll
            for (int j = 0; j < Z; ++j) me[j] = 0.f;
            for (int k = lg.lig; k < K; k += L) {
              float eps[Z];
              load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t, b, a.noise.b_offset, k, T, B,
                          K, eps);
#pragma unroll
              for (int j = 0; j < Z; ++j) me[j] += eps[j];
            }
#pragma unroll
            for (int j = 0; j < Z; ++j) me[j] = group_sum(me[j], lg.base, L) * inv_k;
          } else {
            if (!have_eps_cur)
              load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t, b, a.noise.b_offset, 0, T,
                          B, 1, eps_cur);
#pragma unroll
            for (int j = 0; j < Z; ++j) me[j] = eps_cur[j];
          }
#pragma unroll
          for (int j = 0; j < Z; ++j) {
            const float ds = a.d_samples[o + j];
            d_mu[j] += ds; d_sd[j] = fmaf(ds, me[j], d_sd[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < Z; ++j) d_mu[j] += a.d_samples[o + j];
        }
      }
      // --- fused KL(infer || prior) term ------------------------------------
      if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b])) {
#pragma unroll
        for (int j = 0; j < Z; ++j) {
          float g1, g2, g3, g4;
          kld_elem_grad(mu[j], sd[j], pm[j], ps[j], a.kl_weight, g1, g2, g3, g4);
          d_mu[j] += g1; d_sd[j] += g2; d_pm[j] += g3; d_ps[j] += g4;
        }
      }
      // --- product of experts ---------------------------------------------
      poe_step_backward<Z>(a, bits, s, t, b, sGm, sGs, pm, ps, mu, sd, d_mu, d_sd, d_pm, d_ps, d_gm,
                           d_gs, emit);
      // --- prior -------------------------------------------------------------
      if (i == 0) {
        if (emit) {
#pragma unroll
          for (int j = 0; j < Z; ++j) { d_gm[j] += d_pm[j]; d_gs[j] += d_ps[j]; }
        }
        continue;
      }
      const int t_prev = pass_time(i - 1, T, a.direction);
      const int64_t op = (((int64_t)s * T + t_prev) * B + b) * Z;
      float mu_p[Z], sd_p[Z], d_v[Z];
#pragma unroll
      for (int j = 0; j < Z; ++j) {
        mu_p[j] = a.infer_mean[op + j]; sd_p[j] = a.infer_std[op + j];
        mu_c[j] = mu_p[j]; sd_c[j] = sd_p[j];
        d_v[j] = d_ps[j] * 0.5f * fast_rcp(ps[j]);
        c_mu[j] = c_sd[j] = 0.f;
      }
      const bool sampled_prev = pass_samples(a, i - 1);
      for (int sl = 0; sl < slices; ++sl) {
        const int k = sl * L + lg.lig;
        const bool valid = chain_ok && k < K;
        float eps[Z], z[Z], dz[Z];
        if (sampled_prev) {
          load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t_prev, b, a.noise.b_offset,
                      k < K ? k : K - 1, T, B, K, eps);
#pragma unroll
          for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sd_p[j], mu_p[j]);
        } else {
#pragma unroll
          for (int j = 0; j < Z; ++j) { eps[j] = 0.f; z[j] = mu_p[j]; }
        }
        transition_slice_backward<Z, H>(sP + opaque_zero(), sGm, sGs, p.min_std, inv_k, z, pm, d_pm, d_v, valid,
                                        d_gm, d_gs, dz, Xp, Dp, lane, acc, task);
        if (valid) {
#pragma unroll
          for (int j = 0; j < Z; ++j) { c_mu[j] += dz[j]; c_sd[j] = fmaf(dz[j], eps[j], c_sd[j]); }
        }
        if (K == 1) {
#pragma unroll
          for (int j = 0; j < Z; ++j) eps_cur[j] = eps[j];
        }
      }
      if (L > 1) {
        float cc[2 * Z];
#pragma unroll
        for (int j = 0; j < Z; ++j) { cc[j] = c_mu[j]; cc[Z + j] = c_sd[j]; }
        group_sum_vec<2 * Z>(cc, lg.base, L);
#pragma unroll
        for (int j = 0; j < Z; ++j) { c_mu[j] = cc[j]; c_sd[j] = cc[Z + j]; }
      }
      have_eps_cur = sampled_prev;
    }
    wg_tile_flush<TD, kTX>(acc, oidx, G, lane);
    if (SEG && seg + 1 < n_seg) {         // hand the gradient carry to the next (earlier) segment
      if (emit) {
        float* cw = p.seg_carry + ((int64_t)s * B + b) * 2 * Z;
#pragma unroll
        for (int j = 0; j < Z; ++j) { cw[j] = c_mu[j]; cw[Z + j] = c_sd[j]; }
      }
      if (seg_flags) {
        __threadfence();
        __syncwarp();
        if (lane == 0) seg_post(p.seg_done + wt, seg + 1);
      }
    }
  }

  // ---- flush: transition weights, global prior --------------------------------
  wg_flush(smem + (C::NXC + C::NDC) * kRS, PN::WARP_FLOATS, n_warps, L_::SIZE, p.g_trans);
#pragma unroll
  for (int j = 0; j < Z; ++j) {
    const float m = warp_sum(d_gm[j]), sgs = warp_sum(d_gs[j]);
    if (lane == 0) { sRed[warp][j] = m; sRed[warp][Z + j] = sgs; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * Z) {
    float v = 0.f;
    for (int w = 0; w < n_warps; ++w) v += sRed[w][threadIdx.x];
    if (threadIdx.x < Z) {
      if (v != 0.f) atomicAdd(p.g_z0_mean + threadIdx.x, v);
    } else {
      const int j = threadIdx.x - Z;      // gs = exp(z0_log_std) + min_std
      v *= expf(p.z0_log_std[j]);
      if (v != 0.f) atomicAdd(p.g_z0_log_std + j, v);
    }
  }
}

template <int Z, int H>
inline size_t chain_bwd_smem_bytes(int warps) {
  using PN = GtfPanels<Z, H>;
  const WgSpec spec = PN::spec();
  return sizeof(float) * ((size_t)warps * PN::WARP_FLOATS) +
         sizeof(int) * (size_t)wg_table_ints<PN::TD, kTX>(spec);
}

// =========================================================================
// prior-matching term  kld_prior (models/dmm.py:496-501) forward + backward,
// both directions in one launch (block 0 = fwd, block 1 = bwd), one warp each.
// loss += coef * KL( p(z) || E_k[p(z_next | z_k)] ),  z_k ~ p(z)
// =========================================================================
struct MatchParams {
  const float* trans_w[2];
  float* g_trans[2];
  const float* z0_mean;
  const float* z0_log_std;
  float* g_z0_mean;
  float* g_z0_log_std;
  const float* eps;          // (2, K, Z) or null
  uint64_t seed;
  int K;
  float min_std;
  float coef_static;         // match_mult * kld_mult
  const float* count;        // device scalar mask.sum() (nullable: folded into coef_static)
  double* loss_acc;
  int with_grad;
};

template <int Z, int H>
__global__ void __launch_bounds__(32) match_kernel(const __grid_constant__ MatchParams p) {
  using L_ = GtfLayout<Z, H>;
  using P = GtfPack<Z, H>;
  using PN = GtfPanels<Z, H>;
  using C = GtfCols<Z, H>;
  constexpr int TD = PN::TD;
  BFVI_DYN_SMEM(float, smem);
  __shared__ __align__(16) float sP[P::SIZE];
  __shared__ float sGm[Z], sGs[Z];
  const int dir = blockIdx.x, lane = threadIdx.x;
  const WgSpec spec = PN::spec();
  float* Xp = smem;
  float* Dp = Xp + C::NXC * kRS;
  float* G = Dp + C::NDC * kRS;
  int* tasks = reinterpret_cast<int*>(smem + PN::WARP_FLOATS);
  int* oidx = tasks + 32 * 4;
  gtf_pack_load<Z, H>(p.trans_w[dir], sP);
  if (lane < Z) { sGm[lane] = p.z0_mean[lane]; sGs[lane] = expf(p.z0_log_std[lane]) + p.min_std; }
  wg_build_tables<TD, kTX>(spec, tasks, oidx, 1, L_::SIZE);
  for (int i = lane; i < L_::SIZE + 32; i += 32) G[i] = 0.f;
  for (int i = lane; i < (C::NXC + C::NDC) * kRS; i += 32) Xp[i] = 0.f;
  __syncwarp();
  Xp[panel_at(C::XZ, lane)] = 1.f; Xp[panel_at(C::XH1, lane)] = 1.f;
  Xp[panel_at(C::XH3, lane)] = 1.f; Xp[panel_at(C::XNL, lane)] = 1.f;
  __syncthreads();
  const int4 task = reinterpret_cast<const int4*>(tasks)[lane];
  const int K = p.K;
  const float coef = p.coef_static * (p.count != nullptr ? p.count[0] : 1.f);
  const float* eps_ext = p.eps ? p.eps + (size_t)dir * K * Z : nullptr;
  const int n_sl = (K + 31) / 32;

  // forward: moments over particles
  float sm[Z], sv[Z], sq[Z];
#pragma unroll
  for (int j = 0; j < Z; ++j) sm[j] = sv[j] = sq[j] = 0.f;
  for (int sl = 0; sl < n_sl; ++sl) {
    const int k = sl * 32 + lane;
    float eps[Z], z[1][Z];
    load_eps<Z>(eps_ext, p.seed, 100u + dir, 0, 0, 0, 0u, k < K ? k : K - 1, 1, 1, K, eps);
#pragma unroll
    for (int j = 0; j < Z; ++j) z[0][j] = fmaf(eps[j], sGs[j], sGm[j]);
    gtf_rows_forward<Z, H, 1>(sP, p.min_std, z, [&](int, int j, float qm, float qs) {
      if (k < K) {
        float m_k, s_k;
        poe2_forward(sGm[j], sGs[j], qm, qs, m_k, s_k);
        sm[j] += m_k; sv[j] = fmaf(s_k, s_k, sv[j]); sq[j] = fmaf(m_k, m_k, sq[j]);
      }
    });
  }
  const float inv_k = 1.f / (float)K;
  float nm[Z], ns[Z], kl = 0.f;
#pragma unroll
  for (int j = 0; j < Z; ++j) {
    nm[j] = warp_sum(sm[j]) * inv_k;
    ns[j] = sqrtf(warp_sum(sv[j]) * inv_k + (warp_sum(sq[j]) * inv_k - nm[j] * nm[j]));
    kl += kld_elem(sGm[j], sGs[j], nm[j], ns[j]);       // KL(global || next)
  }
  if (lane == 0 && p.loss_acc != nullptr) atomicAdd(p.loss_acc, (double)(coef * kl));
  if (!p.with_grad) return;

  // backward
  float d_gm[Z], d_gs[Z], d_nm[Z], d_v[Z];
#pragma unroll
  for (int j = 0; j < Z; ++j) {
    float g1, g2, d_ns;
    kld_elem_grad(sGm[j], sGs[j], nm[j], ns[j], coef, g1, g2, d_nm[j], d_ns);
    d_gm[j] = lane == 0 ? g1 : 0.f;
    d_gs[j] = lane == 0 ? g2 : 0.f;
    d_v[j] = d_ns * 0.5f * fast_rcp(ns[j]);
  }
  float2 acc[TD][kTX];
#pragma unroll
  for (int i = 0; i < TD; ++i)
#pragma unroll
    for (int j = 0; j < kTX; ++j) acc[i][j] = float2{0.f, 0.f};
  for (int sl = 0; sl < n_sl; ++sl) {
    const int k = sl * 32 + lane;
    const bool valid = k < K;
    float eps[Z], z[Z], dz[Z];
    load_eps<Z>(eps_ext, p.seed, 100u + dir, 0, 0, 0, 0u, valid ? k : K - 1, 1, 1, K, eps);
#pragma unroll
    for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sGs[j], sGm[j]);
    transition_slice_backward<Z, H>(sP, sGm, sGs, p.min_std, inv_k, z, nm, d_nm, d_v, valid, d_gm, d_gs, dz,
                                    Xp, Dp, lane, acc, task);
    if (valid) {                        // z_k = gm + eps * gs
#pragma unroll
      for (int j = 0; j < Z; ++j) { d_gm[j] += dz[j]; d_gs[j] = fmaf(dz[j], eps[j], d_gs[j]); }
    }
  }
  wg_tile_flush<TD, kTX>(acc, oidx, G, lane);
  wg_flush(G, PN::WARP_FLOATS, 1, L_::SIZE, p.g_trans[dir]);
#pragma unroll
  for (int j = 0; j < Z; ++j) {
    const float m = warp_sum(d_gm[j]), sgs = warp_sum(d_gs[j]);
    if (lane == 0) {
      atomicAdd(p.g_z0_mean + j, m);
      atomicAdd(p.g_z0_log_std + j, sgs * expf(p.z0_log_std[j]));
    }
  }
}

template <int Z, int H>
inline size_t match_smem_bytes() {
  using PN = GtfPanels<Z, H>;
  const WgSpec spec = PN::spec();
  return sizeof(float) * (size_t)PN::WARP_FLOATS + sizeof(int) * (size_t)wg_table_ints<PN::TD, kTX>(spec);
}

}  // namespace bfvi
