// bfvi_generic.cuh — elementwise / reduction kernels of the large-dim kernel family
// (any z_dim, h_dim: default MultiDMM(32, 32), Weizmann 256/256, the scaled 64/512 model).
//
// At these sizes every Linear layer is a dense contraction and runs on the tcgen05
// tensor cores (bfvi_tc.cuh), one launch per layer over ALL rows of a time step
// (rows = sequences x particles).  What is left is the per-component math between the
// GEMMs; it is written for runtime Z with one thread per (sequence, latent component):
// coalesced along the component axis, particles reduced in-thread.
//
//   prep_rows_kernel    NaN -> mask + zero fill of an encoder input      (models/dmm.py:165-167)
//   softplus_kernel     std = softplus(a) + min_std, in place            (models/common.py:41,66)
//   sample_rows_kernel  z[b,k,:] = mu[b,:] + eps[b,k,:] * sd[b,:]        (models/dgts.py:177-180)
//   step_kernel         gate / mean mix of the GTF heads, product with the global prior,
//                       mixture moments over particles, product of experts with the
//                       observations, outputs of the step                (models/common.py:62-68,
//                       models/dmm.py:239-258,388-405, models/dgts.py:40-51,78-83)
//
// Same numerics as the small-dim family: IEEE-rounded product-of-experts step, accurate
// softplus, the same Philox stream (so both families draw identical noise for a seed).
#pragma once
#include "../../include/bfvi.h"
#include "bfvi_math.cuh"
#include "bfvi_rng.cuh"
#include "bfvi_chain.cuh"

namespace bfvi {
namespace gen {

// one N(0,1) draw, component zi of particle k (same stream as load_eps<Z>)
__device__ __forceinline__ float eps_at(const bfvi_noise& nz, int s, int t, int b, int k, int zi, int T, int B,
                                        int K, int Z) {
  if (nz.eps != nullptr) return nz.eps[((((int64_t)s * T + t) * B + b) * K + k) * Z + zi];
  float n[4];
  normal4(nz.seed, nz.stream_id, (unsigned)s, (unsigned)t, (unsigned)b + nz.b_offset, (unsigned)k,
          (unsigned)(zi >> 2), n);
  return n[zi & 3];
}

__global__ void __launch_bounds__(256)
prep_rows_kernel(const float* __restrict__ x, int64_t n_rows, int d, float* __restrict__ x0,
                 uint8_t* __restrict__ mask) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    bool any_nan = false;
    for (int i = 0; i < d; ++i) {
      float v = x[r * d + i];
      if (v != v) { any_nan = true; v = 0.f; }
      x0[r * d + i] = v;
    }
    mask[r] = any_nan ? 0 : 1;
  }
}

__global__ void __launch_bounds__(256) softplus_kernel(float* __restrict__ a, int64_t n, float min_std) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    a[i] = softplus_f(a[i]) + min_std;
}

struct StepParams {
  bfvi_filter_args a;          // S = 1; experts / outputs / noise of the pass
  const float* z0_mean;
  const float* z0_log_std;
  float min_std;
  int Z;
  int i;                       // step index of the pass (0 = first), t = pass_time(i)
  // GTF heads of this step's transition, rows (b, k): pre-sigmoid gate, nonlinear, linear,
  // pre-softplus std; null when i == 0
  const float* g; const float* nl; const float* lin; const float* as;
  float* zrows;                // (B, K, Z) particles of step i for the next transition (nullable)
};

__device__ __forceinline__ int gen_pass_time(int i, int T, int direction) {
  return direction == BFVI_DIR_BWD ? T - 1 - i : i;
}

// particles of the PREVIOUS step -> GEMM input rows (used when the step kernel did not
// already write them, i.e. never in the current pipeline; kept for the op-level API)
__global__ void __launch_bounds__(256) sample_rows_kernel(StepParams p, int t_src, int sampled) {
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles;
  const int64_t n = (int64_t)a.B * K * Z;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int zi = (int)(idx % Z), k = (int)((idx / Z) % K), b = (int)(idx / ((int64_t)Z * K));
    const int64_t o = ((int64_t)t_src * a.B + b) * Z + zi;
    const float mu = a.infer_mean[o], sd = a.infer_std[o];
    p.zrows[idx] = sampled ? fmaf(eps_at(a.noise, 0, t_src, b, k, zi, a.T, a.B, K, Z), sd, mu) : mu;
  }
}

// one filtering step for every (b, zi)
__global__ void __launch_bounds__(128) step_kernel(const __grid_constant__ StepParams p) {
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T;
  const int t = gen_pass_time(p.i, T, a.direction);
  const unsigned bits = a.set_expert_bits[0];
  const float inv_k = 1.f / (float)K;
  float kl_sum = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)B * Z;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int zi = (int)(idx % Z), b = (int)(idx / Z);
    const float gm = p.z0_mean[zi], gs = expf(p.z0_log_std[zi]) + p.min_std;      // models/dmm.py:126-127
    float pm, ps;
    if (p.i == 0) { pm = gm; ps = gs; }
    else {
      float sm = 0.f, sv = 0.f, sq = 0.f;
      for (int k = 0; k < K; ++k) {
        const int64_t r = ((int64_t)b * K + k) * Z + zi;
        const float gate = sigmoid_f(p.g[r]);
        const float nl = p.nl[r], lin = p.lin[r];
        const float qm = fmaf(gate, nl - lin, lin);
        const float qs = softplus_f(p.as[r]) + p.min_std;
        float m_k, s_k;
        poe2_forward(gm, gs, qm, qs, m_k, s_k);
        if (K == 1) { sm = m_k; sv = s_k; }
        else { sm += m_k; sv = fmaf(s_k, s_k, sv); sq = fmaf(m_k, m_k, sq); }
      }
      if (K == 1) { pm = sm; ps = sv; }
      else {                                                // models/dgts.py:78-83
        pm = sm * inv_k;
        ps = sqrtf(sv * inv_k + (sq * inv_k - pm * pm));
      }
    }
    // product of experts, prior first then the experts in order (models/dgts.py:40-51)
    float S = poe_prec(ps);
    float N = __fmul_rn(pm, S);
    for (int e = 0; e < a.n_experts; ++e) {
      if (!((bits >> e) & 1u)) continue;
      const bfvi_expert& ex = a.experts[e];
      bool m = true;
      if (ex.mask != nullptr) m = ex.mask[t * ex.mstride_t + b * ex.mstride_b] != 0;
      if (ex.zero_mask_last_t && t == T - 1) m = false;
      const float w = m ? 1.f : 0.f;
      float mean, std;
      if (ex.kind == BFVI_EXPERT_INV_PRIOR) { mean = gm; std = -gs; }
      else {
        const int64_t off = t * ex.stride_t + b * ex.stride_b + zi;
        mean = ex.mean[off]; std = ex.std[off];
      }
      const float te = __fmul_rn(poe_prec(std), w);
      S = __fadd_rn(S, te);
      N = __fadd_rn(N, __fmul_rn(__fmul_rn(mean, w), te));
    }
    const float mq = __fdiv_rn(N, S);
    const float mu = (mq != mq) ? 0.f : mq;
    const float sd = __fsqrt_rn(__fdiv_rn(1.f, S));
    const int64_t o = ((int64_t)t * B + b) * Z + zi;
    a.infer_mean[o] = mu; a.infer_std[o] = sd;
    a.prior_mean[o] = pm; a.prior_std[o] = ps;
    if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b]))
      kl_sum += kld_elem_fast(mu, sd, pm, ps);
    // particles of this step: input rows of the next transition and the `samples` output
    const bool sampled = a.sample || K > 1 || (p.i == 0 && a.sample_init);          // models/dmm.py:398
    float se = 0.f;
    for (int k = 0; k < K; ++k) {
      const float z = sampled ? fmaf(eps_at(a.noise, 0, t, b, k, zi, T, B, K, Z), sd, mu) : mu;
      if (p.zrows != nullptr) p.zrows[((int64_t)b * K + k) * Z + zi] = z;
      se += z;
    }
    if (a.samples != nullptr) a.samples[o] = se * inv_k;
  }
  if (a.loss_acc != nullptr && a.kl_weight != 0.f) block_reduce_add_double(kl_sum * a.kl_weight, a.loss_acc);
}

}  // namespace gen
}  // namespace bfvi
