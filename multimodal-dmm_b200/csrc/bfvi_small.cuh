// bfvi_small.cuh — register-resident kernels for small latent / hidden sizes
// (spirals configuration Z=5, H=20 and neighbours).
//
// Matrices this small cannot feed tcgen05 tiles, so every "row" (a particle of the
// backward filter, or a whole sequence when K == 1) lives in the registers of one
// thread and the six GTF layers are fully unrolled FFMA chains over weights held
// in shared memory.  The time recurrence is serial and stays on-chip: one kernel
// launch walks all T steps of its sequences.
//
// Two thread mappings share the code:
//   WPC = true  ("warp per chain", K > 1): lanes are particles; the mixture
//               moment-matching over particles is a warp-shuffle reduction.
//   WPC = false ("thread per chain", K == 1): lanes are 32 different sequences.
//
// Reference semantics: MultiDMM.z_filter / z_next (models/dmm.py:214-258,319-412),
// MultiDGTS.product_of_experts / mean_of_experts (models/dgts.py:15-83),
// losses.kld_gauss / nll_gauss (models/losses.py:14-21,68-89),
// GaussianMLP / GaussianGTF (models/common.py:25-68).
#pragma once
#include "../../include/bfvi.h"
#include "bfvi_math.cuh"
#include "bfvi_rng.cuh"
#include "bfvi_wgrad.cuh"

namespace bfvi {

constexpr int kTX = 4;                       // weight-gradient tile width
constexpr int kFilterFwdThreads = 128;
constexpr int kFilterBwdWarps = 4;
constexpr int kMlpWarps = 4;
constexpr int kMlpTD = 4;

struct FilterParams {
  bfvi_filter_args a;
  const float* trans_w;      // flat GTF block of the pass direction
  const float* z0_mean;
  const float* z0_log_std;
  float* g_trans;            // gradient block of the same GTF   (backward only)
  float* g_z0_mean;
  float* g_z0_log_std;
  float min_std;
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
// expert order (models/dmm.py:388-395 / models/dgts.py:40-51).
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
      const float te = poe_prec(std) * w;
      S[i] = __fadd_rn(S[i], te);
      N[i] = __fadd_rn(N[i], __fmul_rn(mean * w, te));
    }
  }
#pragma unroll
  for (int i = 0; i < Z; ++i) {
    const float m = N[i] / S[i];
    mu[i] = (m != m) ? 0.f : m;                // models/dgts.py:49
    sd[i] = sqrtf(1.f / S[i]);
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
    const float tp = poe_prec(ps[i]);
    d_pm[i] += d_n[i] * tp;
    d_ps[i] += (d_n[i] * pm[i] + d_s[i]) * poe_prec_grad(ps[i], tp);
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
        const float std = -gs[i], te = poe_prec(std);
        if (emit) {
          d_gm[i] += d_n[i] * te;
          d_gs[i] -= (d_n[i] * gm[i] + d_s[i]) * poe_prec_grad(std, te);
        }
      }
    } else if (ex.d_mean != nullptr && emit) {
#pragma unroll
      for (int i = 0; i < Z; ++i) {
        const float mean = ex.mean[off + i], std = ex.std[off + i], te = poe_prec(std);
        atomicAdd(ex.d_mean + off + i, d_n[i] * te);
        atomicAdd(ex.d_std + off + i, (d_n[i] * mean + d_s[i]) * poe_prec_grad(std, te));
      }
    }
  }
}

// =========================================================================
// z_filter forward
// =========================================================================
template <int Z, int H, bool WPC>
__global__ void __launch_bounds__(kFilterFwdThreads)
filter_fwd_kernel(const __grid_constant__ FilterParams p) {
  using L = GtfLayout<Z, H>;
  __shared__ __align__(16) float sW[L::SIZE];
  __shared__ float sGm[Z], sGs[Z];
  const bfvi_filter_args& a = p.a;
  for (int i = threadIdx.x; i < L::SIZE; i += blockDim.x) sW[i] = p.trans_w[i];
  if (threadIdx.x < Z) {
    sGm[threadIdx.x] = p.z0_mean[threadIdx.x];
    sGs[threadIdx.x] = expf(p.z0_log_std[threadIdx.x]) + p.min_std;       // models/dmm.py:126-127
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int T = a.T, B = a.B, K = a.n_particles;
  const int n_chains = a.S * B;
  float kl_sum = 0.f;

  if (WPC) {
    const int wpb = blockDim.x >> 5;
    for (int chain = blockIdx.x * wpb + (threadIdx.x >> 5); chain < n_chains; chain += gridDim.x * wpb) {
      const int s = chain / B, b = chain % B;
      const unsigned bits = a.set_expert_bits[s];
      float mu_p[Z], sd_p[Z];
      int t_prev = 0;
      for (int i = 0; i < T; ++i) {
        const int t = pass_time(i, T, a.direction);
        float pm[Z], ps[Z];
        if (i == 0) {
#pragma unroll
          for (int j = 0; j < Z; ++j) { pm[j] = sGm[j]; ps[j] = sGs[j]; }
        } else {
          float sm[Z], sv[Z], sq[Z];
#pragma unroll
          for (int j = 0; j < Z; ++j) sm[j] = sv[j] = sq[j] = 0.f;
          for (int k = lane; k < ((K + 31) & ~31); k += 32) {
            float eps[Z], z[Z], qm[Z], qs[Z];
            const int kk = k < K ? k : K - 1;
            load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t_prev, b, a.noise.b_offset,
                        kk, T, B, K, eps);
#pragma unroll
            for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sd_p[j], mu_p[j]);
            GtfAct<Z, H> act;
            gtf_forward<Z, H>(sW, p.min_std, z, act, qm, qs);
            if (k < K) {
#pragma unroll
              for (int j = 0; j < Z; ++j) {
                float m_k, s_k;
                poe2_forward(sGm[j], sGs[j], qm[j], qs[j], m_k, s_k);
                sm[j] += m_k; sv[j] = fmaf(s_k, s_k, sv[j]); sq[j] = fmaf(m_k, m_k, sq[j]);
              }
            }
          }
          const float inv_k = 1.f / (float)K;
#pragma unroll
          for (int j = 0; j < Z; ++j) {                       // models/dgts.py:78-83
            const float m = warp_sum(sm[j]) * inv_k;
            const float v = warp_sum(sv[j]) * inv_k + (warp_sum(sq[j]) * inv_k - m * m);
            pm[j] = m; ps[j] = sqrtf(v);
          }
        }
        float mu[Z], sd[Z];
        poe_step_forward<Z>(a, bits, s, t, b, sGm, sGs, pm, ps, mu, sd);
        const int64_t o = (((int64_t)s * T + t) * B + b) * Z;
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < Z; ++j) {
            a.infer_mean[o + j] = mu[j]; a.infer_std[o + j] = sd[j];
            a.prior_mean[o + j] = pm[j]; a.prior_std[o + j] = ps[j];
          }
          if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b])) {
#pragma unroll
            for (int j = 0; j < Z; ++j) kl_sum += kld_elem(mu[j], sd[j], pm[j], ps[j]);
          }
        }
        if (a.samples != nullptr) {                           // mean over particles, models/dmm.py:402
          float se[Z];
#pragma unroll
          for (int j = 0; j < Z; ++j) se[j] = 0.f;
          for (int k = lane; k < ((K + 31) & ~31); k += 32) {
            float eps[Z];
            const int kk = k < K ? k : K - 1;
            load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t, b, a.noise.b_offset, kk, T,
                        B, K, eps);
            if (k < K) {
#pragma unroll
              for (int j = 0; j < Z; ++j) se[j] += eps[j];
            }
          }
#pragma unroll
          for (int j = 0; j < Z; ++j) se[j] = warp_sum(se[j]) / (float)K;
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < Z; ++j) a.samples[o + j] = fmaf(se[j], sd[j], mu[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < Z; ++j) { mu_p[j] = mu[j]; sd_p[j] = sd[j]; }
        t_prev = t;
      }
    }
  } else {
    for (int chain = blockIdx.x * blockDim.x + threadIdx.x; chain < n_chains;
         chain += gridDim.x * blockDim.x) {
      const int s = chain / B, b = chain % B;
      const unsigned bits = a.set_expert_bits[s];
      float z[Z];
      for (int i = 0; i < T; ++i) {
        const int t = pass_time(i, T, a.direction);
        float pm[Z], ps[Z];
        if (i == 0) {
#pragma unroll
          for (int j = 0; j < Z; ++j) { pm[j] = sGm[j]; ps[j] = sGs[j]; }
        } else {
          float qm[Z], qs[Z];
          GtfAct<Z, H> act;
          gtf_forward<Z, H>(sW, p.min_std, z, act, qm, qs);
#pragma unroll
          for (int j = 0; j < Z; ++j) poe2_forward(sGm[j], sGs[j], qm[j], qs[j], pm[j], ps[j]);
        }
        float mu[Z], sd[Z];
        poe_step_forward<Z>(a, bits, s, t, b, sGm, sGs, pm, ps, mu, sd);
        const int64_t o = (((int64_t)s * T + t) * B + b) * Z;
#pragma unroll
        for (int j = 0; j < Z; ++j) {
          a.infer_mean[o + j] = mu[j]; a.infer_std[o + j] = sd[j];
          a.prior_mean[o + j] = pm[j]; a.prior_std[o + j] = ps[j];
        }
        if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b])) {
#pragma unroll
          for (int j = 0; j < Z; ++j) kl_sum += kld_elem(mu[j], sd[j], pm[j], ps[j]);
        }
        if (pass_samples(a, i)) {
          float eps[Z];
          load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t, b, a.noise.b_offset, 0, T, B,
                      1, eps);
#pragma unroll
          for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sd[j], mu[j]);
        } else {
#pragma unroll
          for (int j = 0; j < Z; ++j) z[j] = mu[j];
        }
        if (a.samples != nullptr) {
#pragma unroll
          for (int j = 0; j < Z; ++j) a.samples[o + j] = z[j];
        }
      }
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
  // X panel: [1, z] [1, h1] [1, h3] [1, nl]
  static constexpr int XZ = 0, XH1 = 1 + Z, XH3 = XH1 + 1 + H, XNL = XH3 + 1 + H, NXC = XNL + 1 + Z;
  // D panel: d_a1 | d_a3 | d_lin | d_ag | d_nl | d_as
  static constexpr int DA1 = 0, DA3 = H, DLIN = 2 * H, DAG = 2 * H + Z, DNL = 2 * H + 2 * Z,
                       DAS = 2 * H + 3 * Z, NDC = 2 * H + 4 * Z;
  static constexpr int TD = Z <= 8 ? Z : 8;
  __host__ __device__ static WgSpec spec() {
    using L = GtfLayout<Z, H>;
    WgSpec s;
    s.n_blocks = 6;
    s.blk[0] = WgBlock{DA1, H, XZ, 1 + Z, L::G0W, L::G0B};
    s.blk[1] = WgBlock{DA3, H, XZ, 1 + Z, L::N0W, L::N0B};
    s.blk[2] = WgBlock{DLIN, Z, XZ, 1 + Z, L::LW, L::LB};
    s.blk[3] = WgBlock{DAG, Z, XH1, 1 + H, L::G2W, L::G2B};
    s.blk[4] = WgBlock{DNL, Z, XH3, 1 + H, L::N2W, L::N2B};
    s.blk[5] = WgBlock{DAS, Z, XNL, 1 + Z, L::SW, L::SB};
    return s;
  }
  // floats of dynamic shared memory per warp: panels + accumulator (+32 dump slots)
  static constexpr int WARP_FLOATS = (NXC + NDC) * kRS + GtfLayout<Z, H>::SIZE + 32;
};

// stage one row (this lane) of the GTF backward into the panels
template <int Z, int H>
__device__ __forceinline__ void gtf_stage_row(float* Xp, float* Dp, int lane, bool valid,
                                              const float (&z)[Z], const GtfAct<Z, H>& act,
                                              const GtfGrad<Z, H>& g) {
  using P = GtfPanels<Z, H>;
  const float v = valid ? 1.f : 0.f;
#pragma unroll
  for (int j = 0; j < Z; ++j) {
    Xp[(P::XZ + 1 + j) * kRS + lane] = valid ? z[j] : 0.f;
    Xp[(P::XNL + 1 + j) * kRS + lane] = valid ? act.nl[j] : 0.f;
    Dp[(P::DLIN + j) * kRS + lane] = g.d_lin[j] * v;
    Dp[(P::DAG + j) * kRS + lane] = g.d_ag[j] * v;
    Dp[(P::DNL + j) * kRS + lane] = g.d_nl[j] * v;
    Dp[(P::DAS + j) * kRS + lane] = g.d_as[j] * v;
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
    Xp[(P::XH1 + 1 + h) * kRS + lane] = valid ? act.h1[h] : 0.f;
    Xp[(P::XH3 + 1 + h) * kRS + lane] = valid ? act.h3[h] : 0.f;
    Dp[(P::DA1 + h) * kRS + lane] = valid ? g.d_a1[h] : 0.f;
    Dp[(P::DA3 + h) * kRS + lane] = valid ? g.d_a3[h] : 0.f;
  }
}

template <int Z, int H, bool WPC>
__global__ void __launch_bounds__(kFilterBwdWarps * 32)
filter_bwd_kernel(const __grid_constant__ FilterParams p) {
  using L = GtfLayout<Z, H>;
  using P = GtfPanels<Z, H>;
  constexpr int TD = P::TD;
  BFVI_DYN_SMEM(float, smem);
  __shared__ __align__(16) float sW[L::SIZE];
  __shared__ float sGm[Z], sGs[Z];
  __shared__ float sRed[kFilterBwdWarps][2 * Z];
  const bfvi_filter_args& a = p.a;
  const WgSpec spec = P::spec();
  const int rounds = wg_rounds<TD, kTX>(spec);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;

  float* Xp = smem + (size_t)warp * P::WARP_FLOATS;
  float* Dp = Xp + P::NXC * kRS;
  float* G = Dp + P::NDC * kRS;
  int* tasks = reinterpret_cast<int*>(smem + (size_t)n_warps * P::WARP_FLOATS);
  int* oidx = tasks + rounds * 32 * 4;

  for (int i = threadIdx.x; i < L::SIZE; i += blockDim.x) sW[i] = p.trans_w[i];
  if (threadIdx.x < Z) {
    sGm[threadIdx.x] = p.z0_mean[threadIdx.x];
    sGs[threadIdx.x] = expf(p.z0_log_std[threadIdx.x]) + p.min_std;
  }
  wg_build_tables<TD, kTX>(spec, tasks, oidx, rounds, L::SIZE);
  for (int i = lane; i < L::SIZE + 32; i += 32) G[i] = 0.f;
  Xp[(P::XZ) * kRS + lane] = 1.f;
  Xp[(P::XH1) * kRS + lane] = 1.f;
  Xp[(P::XH3) * kRS + lane] = 1.f;
  Xp[(P::XNL) * kRS + lane] = 1.f;
  __syncthreads();

  const int T = a.T, B = a.B, K = a.n_particles;
  const int n_chains = a.S * B;
  float d_gm[Z], d_gs[Z];                 // global-prior gradient, per thread
#pragma unroll
  for (int j = 0; j < Z; ++j) d_gm[j] = d_gs[j] = 0.f;

  const int gwarp = blockIdx.x * n_warps + warp, total_warps = gridDim.x * n_warps;
  const int per_iter = WPC ? 1 : 32;
  for (int base = gwarp * per_iter; base < n_chains; base += total_warps * per_iter) {
    const int chain_raw = WPC ? base : base + lane;
    const bool chain_ok = chain_raw < n_chains;
    const int chain = chain_ok ? chain_raw : n_chains - 1;
    const int s = chain / B, b = chain % B;
    const unsigned bits = a.set_expert_bits[s];
    const bool emit = chain_ok && (!WPC || lane == 0);   // one writer per chain
    float c_mu[Z], c_sd[Z], eps_cur[Z];
#pragma unroll
    for (int j = 0; j < Z; ++j) c_mu[j] = c_sd[j] = eps_cur[j] = 0.f;
    bool have_eps_cur = false;

    for (int i = T - 1; i >= 0; --i) {
      const int t = pass_time(i, T, a.direction);
      const int64_t o = (((int64_t)s * T + t) * B + b) * Z;
      float mu[Z], sd[Z], pm[Z], ps[Z], d_mu[Z], d_sd[Z], d_pm[Z], d_ps[Z];
#pragma unroll
      for (int j = 0; j < Z; ++j) {
        mu[j] = a.infer_mean[o + j]; sd[j] = a.infer_std[o + j];
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
          if (WPC) {
#pragma unroll
            for (int j = 0; j < Z; ++j) me[j] = 0.f;
            for (int k = lane; k < ((K + 31) & ~31); k += 32) {
              float eps[Z];
              const int kk = k < K ? k : K - 1;
              load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t, b, a.noise.b_offset, kk,
                          T, B, K, eps);
              if (k < K) {
#pragma unroll
                for (int j = 0; j < Z; ++j) me[j] += eps[j];
              }
            }
#pragma unroll
            for (int j = 0; j < Z; ++j) me[j] = warp_sum(me[j]) / (float)K;
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
      float mu_p[Z], sd_p[Z];
#pragma unroll
      for (int j = 0; j < Z; ++j) { mu_p[j] = a.infer_mean[op + j]; sd_p[j] = a.infer_std[op + j]; }
      const bool sampled_prev = pass_samples(a, i - 1);
#pragma unroll
      for (int j = 0; j < Z; ++j) c_mu[j] = c_sd[j] = 0.f;

      if (WPC) {
        float d_v[Z];
#pragma unroll
        for (int j = 0; j < Z; ++j) d_v[j] = d_ps[j] * 0.5f / ps[j];
        const float inv_k = 1.f / (float)K;
        for (int k = lane; k < ((K + 31) & ~31); k += 32) {
          const bool valid = k < K;
          const int kk = valid ? k : K - 1;
          float eps[Z], z[Z], qm[Z], qs[Z], d_qm[Z], d_qs[Z], dz[Z];
          load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t_prev, b, a.noise.b_offset, kk,
                      T, B, K, eps);
#pragma unroll
          for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sd_p[j], mu_p[j]);
          GtfAct<Z, H> act;
          gtf_forward<Z, H>(sW, p.min_std, z, act, qm, qs);
#pragma unroll
          for (int j = 0; j < Z; ++j) {
            float m_k, s_k, g_gm, g_gs;
            poe2_forward(sGm[j], sGs[j], qm[j], qs[j], m_k, s_k);
            const float d_mk = (d_pm[j] + 2.f * d_v[j] * (m_k - pm[j])) * inv_k;   // models/dgts.py:78-83
            const float d_sk = 2.f * d_v[j] * s_k * inv_k;
            poe2_backward(sGm[j], sGs[j], qm[j], qs[j], m_k, s_k, d_mk, d_sk, g_gm, g_gs, d_qm[j],
                          d_qs[j]);
            if (valid) { d_gm[j] += g_gm; d_gs[j] += g_gs; }
          }
          GtfGrad<Z, H> gg;
          gtf_backward<Z, H>(sW, act, d_qm, d_qs, gg, dz);
          if (valid) {
#pragma unroll
            for (int j = 0; j < Z; ++j) { c_mu[j] += dz[j]; c_sd[j] = fmaf(dz[j], eps[j], c_sd[j]); }
          }
          __syncwarp();
          gtf_stage_row<Z, H>(Xp, Dp, lane, valid, z, act, gg);
          __syncwarp();
          wg_accumulate<TD, kTX>(Dp, Xp, tasks, oidx, rounds, G, lane);
        }
#pragma unroll
        for (int j = 0; j < Z; ++j) { c_mu[j] = warp_sum(c_mu[j]); c_sd[j] = warp_sum(c_sd[j]); }
      } else {
        float z[Z], qm[Z], qs[Z], d_qm[Z], d_qs[Z], dz[Z], eps[Z];
        if (sampled_prev) {
          load_eps<Z>(a.noise.eps, a.noise.seed, a.noise.stream_id, s, t_prev, b, a.noise.b_offset, 0, T,
                      B, 1, eps);
#pragma unroll
          for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sd_p[j], mu_p[j]);
        } else {
#pragma unroll
          for (int j = 0; j < Z; ++j) { eps[j] = 0.f; z[j] = mu_p[j]; }
        }
        GtfAct<Z, H> act;
        gtf_forward<Z, H>(sW, p.min_std, z, act, qm, qs);
#pragma unroll
        for (int j = 0; j < Z; ++j) {
          float g_gm, g_gs;
          poe2_backward(sGm[j], sGs[j], qm[j], qs[j], pm[j], ps[j], d_pm[j], d_ps[j], g_gm, g_gs,
                        d_qm[j], d_qs[j]);
          if (chain_ok) { d_gm[j] += g_gm; d_gs[j] += g_gs; }
        }
        GtfGrad<Z, H> gg;
        gtf_backward<Z, H>(sW, act, d_qm, d_qs, gg, dz);
#pragma unroll
        for (int j = 0; j < Z; ++j) { c_mu[j] = dz[j]; c_sd[j] = dz[j] * eps[j]; eps_cur[j] = eps[j]; }
        have_eps_cur = sampled_prev;
        __syncwarp();
        gtf_stage_row<Z, H>(Xp, Dp, lane, chain_ok, z, act, gg);
        __syncwarp();
        wg_accumulate<TD, kTX>(Dp, Xp, tasks, oidx, rounds, G, lane);
      }
    }
  }

  // ---- flush: transition weights, global prior --------------------------------
  wg_flush(smem + L::SIZE * 0 + (P::NXC + P::NDC) * kRS, P::WARP_FLOATS, n_warps, L::SIZE, p.g_trans);
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
inline size_t filter_bwd_smem_bytes() {
  using P = GtfPanels<Z, H>;
  const WgSpec spec = P::spec();
  return sizeof(float) * ((size_t)kFilterBwdWarps * P::WARP_FLOATS) +
         sizeof(int) * (size_t)wg_table_ints<P::TD, kTX>(spec);
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
  using L = GtfLayout<Z, H>;
  using P = GtfPanels<Z, H>;
  constexpr int TD = P::TD;
  BFVI_DYN_SMEM(float, smem);
  __shared__ __align__(16) float sW[L::SIZE];
  __shared__ float sGm[Z], sGs[Z];
  const int dir = blockIdx.x, lane = threadIdx.x;
  const WgSpec spec = P::spec();
  const int rounds = wg_rounds<TD, kTX>(spec);
  float* Xp = smem;
  float* Dp = Xp + P::NXC * kRS;
  float* G = Dp + P::NDC * kRS;
  int* tasks = reinterpret_cast<int*>(smem + P::WARP_FLOATS);
  int* oidx = tasks + rounds * 32 * 4;
  for (int i = lane; i < L::SIZE; i += 32) sW[i] = p.trans_w[dir][i];
  if (lane < Z) { sGm[lane] = p.z0_mean[lane]; sGs[lane] = expf(p.z0_log_std[lane]) + p.min_std; }
  wg_build_tables<TD, kTX>(spec, tasks, oidx, rounds, L::SIZE);
  for (int i = lane; i < L::SIZE + 32; i += 32) G[i] = 0.f;
  Xp[(P::XZ) * kRS + lane] = 1.f; Xp[(P::XH1) * kRS + lane] = 1.f;
  Xp[(P::XH3) * kRS + lane] = 1.f; Xp[(P::XNL) * kRS + lane] = 1.f;
  __syncthreads();
  const int K = p.K;
  const float coef = p.coef_static * (p.count != nullptr ? p.count[0] : 1.f);
  const float* eps_ext = p.eps ? p.eps + (size_t)dir * K * Z : nullptr;

  // forward: moments over particles
  float sm[Z], sv[Z], sq[Z];
#pragma unroll
  for (int j = 0; j < Z; ++j) sm[j] = sv[j] = sq[j] = 0.f;
  for (int k = lane; k < ((K + 31) & ~31); k += 32) {
    float eps[Z], z[Z], qm[Z], qs[Z];
    const int kk = k < K ? k : K - 1;
    load_eps<Z>(eps_ext, p.seed, 100u + dir, 0, 0, 0, 0u, kk, 1, 1, K, eps);
#pragma unroll
    for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sGs[j], sGm[j]);
    GtfAct<Z, H> act;
    gtf_forward<Z, H>(sW, p.min_std, z, act, qm, qs);
    if (k < K) {
#pragma unroll
      for (int j = 0; j < Z; ++j) {
        float m_k, s_k;
        poe2_forward(sGm[j], sGs[j], qm[j], qs[j], m_k, s_k);
        sm[j] += m_k; sv[j] = fmaf(s_k, s_k, sv[j]); sq[j] = fmaf(m_k, m_k, sq[j]);
      }
    }
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
  float d_gm[Z], d_gs[Z], d_nm[Z], d_ns[Z], d_v[Z];
#pragma unroll
  for (int j = 0; j < Z; ++j) {
    float g1, g2;
    kld_elem_grad(sGm[j], sGs[j], nm[j], ns[j], coef, g1, g2, d_nm[j], d_ns[j]);
    d_gm[j] = lane == 0 ? g1 : 0.f;
    d_gs[j] = lane == 0 ? g2 : 0.f;
    d_v[j] = d_ns[j] * 0.5f / ns[j];
  }
  for (int k = lane; k < ((K + 31) & ~31); k += 32) {
    const bool valid = k < K;
    const int kk = valid ? k : K - 1;
    float eps[Z], z[Z], qm[Z], qs[Z], d_qm[Z], d_qs[Z], dz[Z];
    load_eps<Z>(eps_ext, p.seed, 100u + dir, 0, 0, 0, 0u, kk, 1, 1, K, eps);
#pragma unroll
    for (int j = 0; j < Z; ++j) z[j] = fmaf(eps[j], sGs[j], sGm[j]);
    GtfAct<Z, H> act;
    gtf_forward<Z, H>(sW, p.min_std, z, act, qm, qs);
#pragma unroll
    for (int j = 0; j < Z; ++j) {
      float m_k, s_k, g_gm, g_gs;
      poe2_forward(sGm[j], sGs[j], qm[j], qs[j], m_k, s_k);
      const float d_mk = (d_nm[j] + 2.f * d_v[j] * (m_k - nm[j])) * inv_k;
      const float d_sk = 2.f * d_v[j] * s_k * inv_k;
      poe2_backward(sGm[j], sGs[j], qm[j], qs[j], m_k, s_k, d_mk, d_sk, g_gm, g_gs, d_qm[j], d_qs[j]);
      if (valid) { d_gm[j] += g_gm; d_gs[j] += g_gs; }
    }
    GtfGrad<Z, H> gg;
    gtf_backward<Z, H>(sW, act, d_qm, d_qs, gg, dz);
    if (valid) {
#pragma unroll
      for (int j = 0; j < Z; ++j) { d_gm[j] += dz[j]; d_gs[j] = fmaf(dz[j], eps[j], d_gs[j]); }
    }
    __syncwarp();
    gtf_stage_row<Z, H>(Xp, Dp, lane, valid, z, act, gg);
    __syncwarp();
    wg_accumulate<TD, kTX>(Dp, Xp, tasks, oidx, rounds, G, lane);
  }
  wg_flush(G, P::WARP_FLOATS, 1, L::SIZE, p.g_trans[dir]);
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
  using P = GtfPanels<Z, H>;
  const WgSpec spec = P::spec();
  return sizeof(float) * (size_t)P::WARP_FLOATS + sizeof(int) * (size_t)wg_table_ints<P::TD, kTX>(spec);
}

// =========================================================================
// GaussianMLP kernels (encoder / decoder), runtime in/out width, compile-time H
// =========================================================================
struct MlpOffsets { int w1, b1, wm, bm, ws, bs, size; };   // relative to the block start

struct MlpParams {
  const float* w;         // flat MLP block (global)
  float* g;               // gradient block (nullable)
  MlpOffsets off;
  int n_in, n_out;
  int64_t n_rows;
  // encoder
  const float* x;         // (n_rows, n_in); NaN = missing (encoder) / latent z (decoder)
  float* mean; float* std; uint8_t* mask;          // outputs
  const float* d_mean; const float* d_std;         // upstream grads (encoder backward)
  // decoder + NLL
  const float* target;    // (n_rows, n_out), NaN = unobserved
  const uint8_t* row_mask;
  float weight;
  double* loss_acc;
  float* d_x;             // (n_rows, n_in), += (nullable)
};

// hidden layer: h = relu(W1 x + b1); x read from global with NaN -> 0 when `nan_to_zero`
template <int H>
__device__ __forceinline__ bool mlp_hidden(const float* sW, const MlpOffsets& o, const float* xrow,
                                           int n_in, bool nan_to_zero, float (&h)[H]) {
  bool any_nan = false;
#pragma unroll
  for (int j = 0; j < H; ++j) h[j] = sW[o.b1 + j];
  for (int i = 0; i < n_in; ++i) {
    float xi = xrow[i];
    if (xi != xi) { any_nan = true; if (nan_to_zero) xi = 0.f; }
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = fmaf(sW[o.w1 + j * n_in + i], xi, h[j]);
  }
#pragma unroll
  for (int j = 0; j < H; ++j) h[j] = relu_f(h[j]);
  return any_nan;
}
template <int H>
__device__ __forceinline__ float mlp_head(const float* sW, int w_off, int b_off, int o, const float (&h)[H]) {
  float v = sW[b_off + o];
#pragma unroll
  for (int j = 0; j < H; ++j) v = fmaf(sW[w_off + o * H + j], h[j], v);
  return v;
}

// MultiDMM.encode for one modality (models/dmm.py:165-173) / MultiDMM.decode
template <int H>
__global__ void __launch_bounds__(128) mlp_fwd_kernel(const __grid_constant__ MlpParams p) {
  BFVI_DYN_SMEM(float, sW);
  for (int i = threadIdx.x; i < p.off.size; i += blockDim.x) sW[i] = p.w[i];
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < p.n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    float h[H];
    const bool any_nan = mlp_hidden<H>(sW, p.off, p.x + r * p.n_in, p.n_in, p.mask != nullptr, h);
    if (p.mask != nullptr) p.mask[r] = any_nan ? 0 : 1;
    for (int o = 0; o < p.n_out; ++o) {
      p.mean[r * p.n_out + o] = mlp_head<H>(sW, p.off.wm, p.off.bm, o, h);
      p.std[r * p.n_out + o] = softplus_f(mlp_head<H>(sW, p.off.ws, p.off.bs, o, h)) + kMlpMinStd;
    }
  }
}

struct MlpPanels {
  int n_in, n_out, H;
  __host__ __device__ int xh() const { return 1 + n_in; }                 // [1, x] [1, h]
  __host__ __device__ int nxc() const { return 2 + n_in + H; }
  __host__ __device__ int da() const { return 0; }                         // d_a | d_mean | d_pstd
  __host__ __device__ int dm() const { return H; }
  __host__ __device__ int ds() const { return H + n_out; }
  __host__ __device__ int ndc() const { return H + 2 * n_out; }
  __host__ __device__ WgSpec spec(const MlpOffsets& o) const {
    WgSpec s;
    s.n_blocks = 3;
    s.blk[0] = WgBlock{da(), H, 0, 1 + n_in, o.w1, o.b1};
    s.blk[1] = WgBlock{dm(), n_out, xh(), 1 + H, o.wm, o.bm};
    s.blk[2] = WgBlock{ds(), n_out, xh(), 1 + H, o.ws, o.bs};
    return s;
  }
  __host__ __device__ int warp_floats(int size) const { return (nxc() + ndc()) * kRS + size + 32; }
};

// Shared backward tail: given dh (gradient at the hidden activations, already
// accumulated) and the staged d_mean / d_pstd columns, finish the row.
template <int H, bool DECODER>
__global__ void __launch_bounds__(kMlpWarps * 32) mlp_bwd_kernel(const __grid_constant__ MlpParams p) {
  BFVI_DYN_SMEM(float, smem);
  const MlpPanels pn{p.n_in, p.n_out, H};
  const WgSpec spec = pn.spec(p.off);
  const int rounds = wg_rounds<kMlpTD, kTX>(spec);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int size4 = pad4(p.off.size);
  float* sW = smem;
  const int wf = pn.warp_floats(p.off.size);
  float* Xp = smem + size4 + (size_t)warp * wf;
  float* Dp = Xp + pn.nxc() * kRS;
  float* G = Dp + pn.ndc() * kRS;
  int* tasks = reinterpret_cast<int*>(smem + size4 + (size_t)n_warps * wf);
  int* oidx = tasks + rounds * 32 * 4;
  for (int i = threadIdx.x; i < p.off.size; i += blockDim.x) sW[i] = p.w[i];
  const bool with_grad = p.g != nullptr;
  if (with_grad) {
    wg_build_tables<kMlpTD, kTX>(spec, tasks, oidx, rounds, p.off.size);
    for (int i = lane; i < p.off.size + 32; i += 32) G[i] = 0.f;
    Xp[0 * kRS + lane] = 1.f;
    Xp[pn.xh() * kRS + lane] = 1.f;
  }
  __syncthreads();

  float loss = 0.f;
  const int64_t gwarp = (int64_t)blockIdx.x * n_warps + warp, total_warps = (int64_t)gridDim.x * n_warps;
  for (int64_t base = gwarp * 32; base < p.n_rows; base += total_warps * 32) {
    const int64_t r_raw = base + lane;
    const bool ok = r_raw < p.n_rows;
    const int64_t r = ok ? r_raw : p.n_rows - 1;
    const float* xrow = p.x + r * p.n_in;
    float h[H], dh[H];
    mlp_hidden<H>(sW, p.off, xrow, p.n_in, !DECODER, h);
#pragma unroll
    for (int j = 0; j < H; ++j) dh[j] = 0.f;
    const bool row_on = ok && (!DECODER || p.row_mask == nullptr || p.row_mask[r] != 0);
    for (int o = 0; o < p.n_out; ++o) {
      float d_m = 0.f, d_p = 0.f;
      if (DECODER) {
        // losses.nll_gauss (models/losses.py:78-89): elementwise observed mask
        const float xt = p.target[r * p.n_out + o];
        if (row_on && xt == xt) {
          const float mean = mlp_head<H>(sW, p.off.wm, p.off.bm, o, h);
          const float a_s = mlp_head<H>(sW, p.off.ws, p.off.bs, o, h);
          const float std = softplus_f(a_s) + kMlpMinStd;
          loss += nll_gauss_elem(mean, std, xt);
          float g_std;
          nll_gauss_elem_grad(mean, std, xt, p.weight, d_m, g_std);
          d_p = g_std * softplus_grad(a_s);
        }
      } else {
        if (ok) {
          const float a_s = mlp_head<H>(sW, p.off.ws, p.off.bs, o, h);
          d_m = p.d_mean[r * p.n_out + o];
          d_p = p.d_std[r * p.n_out + o] * softplus_grad(a_s);
        }
      }
      if (with_grad) {
        Dp[(pn.dm() + o) * kRS + lane] = d_m;
        Dp[(pn.ds() + o) * kRS + lane] = d_p;
      }
#pragma unroll
      for (int j = 0; j < H; ++j)
        dh[j] = fmaf(sW[p.off.wm + o * H + j], d_m, fmaf(sW[p.off.ws + o * H + j], d_p, dh[j]));
    }
#pragma unroll
    for (int j = 0; j < H; ++j) dh[j] = h[j] > 0.f ? dh[j] : 0.f;
    if (p.d_x != nullptr && ok) {
      for (int i = 0; i < p.n_in; ++i) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) v = fmaf(sW[p.off.w1 + j * p.n_in + i], dh[j], v);
        p.d_x[r * p.n_in + i] += v;
      }
    }
    if (with_grad) {
#pragma unroll
      for (int j = 0; j < H; ++j) {
        Dp[(pn.da() + j) * kRS + lane] = ok ? dh[j] : 0.f;
        Xp[(pn.xh() + 1 + j) * kRS + lane] = ok ? h[j] : 0.f;
      }
      for (int i = 0; i < p.n_in; ++i) {
        float xi = xrow[i];
        if (xi != xi) xi = 0.f;
        Xp[(1 + i) * kRS + lane] = ok ? xi : 0.f;
      }
      __syncwarp();
      wg_accumulate<kMlpTD, kTX>(Dp, Xp, tasks, oidx, rounds, G, lane);
      __syncwarp();
    }
  }
  if (with_grad)
    wg_flush(smem + size4 + (pn.nxc() + pn.ndc()) * kRS, wf, n_warps, p.off.size, p.g);
  if (DECODER && p.loss_acc != nullptr) block_reduce_add_double(loss * p.weight, p.loss_acc);
}

inline size_t mlp_bwd_smem_bytes(int n_in, int n_out, int H, const MlpOffsets& off) {
  const MlpPanels pn{n_in, n_out, H};
  const WgSpec spec = pn.spec(off);
  return sizeof(float) * ((size_t)pad4(off.size) + (size_t)kMlpWarps * pn.warp_floats(off.size)) +
         sizeof(int) * (size_t)wg_table_ints<kMlpTD, kTX>(spec);
}

// =========================================================================
// stand-alone loss kernels (public losses.* API) and utilities
// =========================================================================
__global__ void __launch_bounds__(256)
kld_kernel(const float* m1, const float* s1, const float* m2, const float* s2, const uint8_t* row_mask,
           int64_t n_rows, int z_dim, double* out, float g, float* d_m1, float* d_s1, float* d_m2,
           float* d_s2) {
  float acc = 0.f;
  const int64_t n = n_rows * z_dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool on = row_mask == nullptr || row_mask[i / z_dim] != 0;
    if (out != nullptr) { if (on) acc += kld_elem(m1[i], s1[i], m2[i], s2[i]); }
    else {
      float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
      if (on) kld_elem_grad(m1[i], s1[i], m2[i], s2[i], g, a, b, c, d);
      d_m1[i] = a; d_s1[i] = b; d_m2[i] = c; d_s2[i] = d;
    }
  }
  if (out != nullptr) block_reduce_add_double(acc, out);
}

__global__ void __launch_bounds__(256)
nll_gauss_kernel(const float* mean, const float* std, const float* x, const uint8_t* row_mask,
                 int64_t n_rows, int d, double* out, float g, float* d_mean, float* d_std) {
  float acc = 0.f;
  const int64_t n = n_rows * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const bool on = (xv == xv) && (row_mask == nullptr || row_mask[i / d] != 0);
    if (out != nullptr) { if (on) acc += nll_gauss_elem(mean[i], std[i], xv); }
    else {
      float a = 0.f, b = 0.f;
      if (on) nll_gauss_elem_grad(mean[i], std[i], xv, g, a, b);
      d_mean[i] = a; d_std[i] = b;
    }
  }
  if (out != nullptr) block_reduce_add_double(acc, out);
}

__global__ void __launch_bounds__(256) count_mask_kernel(const uint8_t* mask, int64_t n, float* out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    acc += mask[i] ? 1.f : 0.f;
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

__global__ void finalize_loss_kernel(const double* acc, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)acc[0];
}

__global__ void __launch_bounds__(256)
dump_noise_kernel(uint64_t seed, unsigned stream_id, unsigned b_offset, int S, int T, int B, int K, int Z,
                  float* out) {
  const int64_t n = (int64_t)S * T * B * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int b = (int)((i / K) % B);
    const int t = (int)((i / ((int64_t)K * B)) % T);
    const int s = (int)(i / ((int64_t)K * B * T));
    for (int c = 0; c < (Z + 3) / 4; ++c) {
      float nrm[4];
      normal4(seed, stream_id, (unsigned)s, (unsigned)t, (unsigned)b + b_offset, (unsigned)k, (unsigned)c,
              nrm);
      for (int j = 0; j < 4; ++j)
        if (4 * c + j < Z) out[i * Z + 4 * c + j] = nrm[j];
    }
  }
}

// FP32 FFMA peak probe for the roofline denominator of the small-dim path: every
// thread runs 16 independent FMA chains (register operands only).
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
  const float m = 1.0000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;       // keeps the chains alive, practically never true
}

}  // namespace bfvi
