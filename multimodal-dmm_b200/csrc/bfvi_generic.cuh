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

// Philox seed of a pass: by value, or read from device memory (CUDA-graph replay, bfvi_noise.seed_dev)
__device__ __forceinline__ uint64_t noise_seed(const bfvi_noise& nz) {
  return nz.seed_dev != nullptr ? *nz.seed_dev : nz.seed;
}
// one N(0,1) draw, component zi of particle k (same stream as load_eps<Z>)
__device__ __forceinline__ float eps_at(const bfvi_noise& nz, int s, int t, int b, int k, int zi, int T, int B,
                                        int K, int Z) {
  if (nz.eps != nullptr) return nz.eps[((((int64_t)s * T + t) * B + b) * K + k) * Z + zi];
  float n[4];
  normal4(noise_seed(nz), nz.stream_id, (unsigned)s, (unsigned)t, (unsigned)b + nz.b_offset, (unsigned)k,
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

// (R, 64) fp32 arrays written by the fused transition kernels (heads, dz) are stored with the 16-byte chunks of every
// 128-byte half row XOR-permuted by row % 8 ("swz64"): the producing thread = row layout then writes shared memory
// without bank conflicts and the 32 rows of a warp quadrant leave as ONE contiguous bulk copy (bfvi_fused.cuh).
__device__ __forceinline__ int64_t row64(int64_t row, int zi, int Z, int swz) {
  return swz ? row * 64 + (zi & 32) + ((((zi & 31) >> 2) ^ (int)(row & 7)) << 2) + (zi & 3) : row * Z + zi;
}

struct StepParams {
  int swz;                     // heads (g, nl, lin, as) and dz are in the swz64 layout (fused path, Z = 64)
  bfvi_filter_args a;          // experts / outputs / noise of the pass (S chain sets)
  const float* z0_mean;
  const float* z0_log_std;
  float* g_z0_mean;            // gradient slots of the global prior (backward)
  float* g_z0_log_std;
  float min_std;
  int Z;
  int i;                       // step index of the pass (0 = first), t = pass_time(i)
  int64_t R;                   // rows = S * B * K
  // GTF heads of this step's transition, rows (chain, k): pre-sigmoid gate, nonlinear,
  // linear, pre-softplus std (valid when i > 0)
  const float* g; const float* nl; const float* lin; const float* as;
  float* zrows;                // (R, Z) particles of step i for the next transition (nullable)
  float* zrowsT;               // (Z, R) transposed copy for the weight-gradient GEMMs (nullable)
  float* samplesT;             // (S, Z, T*B) transposed copy of `samples` (nullable)
  // ---- backward ----
  float* c_mu; float* c_sd;    // (C, Z) gradient flowing into infer(i) from the later step
  float* d_pm; float* d_v;     // (C, Z) gradient at the prior mean, 0.5 * d_ps / ps
  float* d_as; float* d_asT;   // (R, Z) / (Z, R) pre-activation gradients of the four GTF heads
  float* d_g; float* d_gT;
  float* d_lin; float* d_linT;
  float* d_nl; float* d_nlT;
  float* gb_std; float* gb_gate2; float* gb_lin; float* gb_nonlin2;   // bias gradient slots (Z each)
  const float* dz;             // (R, Z) gradient at the particles of the previous step
  const float* dz2;            // nullable second partial of the same gradient (summed by bwd_carry_kernel)
  // fused path (nullable): running maxima [|d_g|, |d_nl| (before the std path), |d_as|] of this time step's head gradients as
  // float bit patterns: the fused input-gradient kernel derives the power-of-two scale of its FP16 operand tiles from them;
  // bwd_rows_kernel raises them, bwd_carry_kernel (the last kernel of the time step) clears them
  unsigned* gmax;
};

__device__ __forceinline__ int gen_pass_time(int i, int T, int direction) {
  return direction == BFVI_DIR_BWD ? T - 1 - i : i;
}
__device__ __forceinline__ bool gen_samples(const bfvi_filter_args& a, int i) {
  return a.sample || a.n_particles > 1 || (i == 0 && a.sample_init);      // models/dmm.py:398
}

// particles of step i_src -> GEMM input rows (+ transposed copy); the backward pass uses it
// to regenerate what the forward step kernel wrote
__global__ void __launch_bounds__(256) sample_rows_kernel(const __grid_constant__ StepParams p, int i_src) {
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T;
  const int t = gen_pass_time(i_src, T, a.direction);
  const bool sampled = gen_samples(a, i_src);
  const int64_t n = p.R * Z;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int zi = (int)(idx % Z);
    const int64_t r = idx / Z;
    const int k = (int)(r % K);
    const int64_t c = r / K;
    const int s = (int)(c / B), b = (int)(c % B);
    const int64_t o = (((int64_t)s * T + t) * B + b) * Z + zi;
    const float mu = a.infer_mean[o], sd = a.infer_std[o];
    const float z = sampled ? fmaf(eps_at(a.noise, s, t, b, k, zi, T, B, K, Z), sd, mu) : mu;
    p.zrows[idx] = z;
    if (p.zrowsT != nullptr) p.zrowsT[(int64_t)zi * p.R + r] = z;
  }
}

// Same as sample_rows_kernel for Z % 4 == 0: a thread owns FOUR consecutive components of a row — one
// Philox call instead of four, 128-bit loads / stores — and the transposed copy leaves through a swizzled
// shared tile as full 128-byte lines (a direct store is one 4-byte transaction per element).
// Block = 64 threads = 16 component quads x 4 rows, walking kRowsPerBlock (32) rows.
__global__ void __launch_bounds__(64) sample_rows4_kernel(const __grid_constant__ StepParams p, int i_src) {
  __shared__ float tile[64][32];
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T;
  const int t = gen_pass_time(i_src, T, a.direction);
  const bool sampled = gen_samples(a, i_src);
  const int quad = threadIdx.x & 15, rsub = threadIdx.x >> 4;
  const int zl = quad * 4;                                   // first component inside the block's 64
  const int zi = blockIdx.y * 64 + zl;
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int64_t r1 = r0 + 32 < p.R ? r0 + 32 : p.R;
  if (zi < Z) {
#pragma unroll 2
    for (int rr = 0; rr < 8; ++rr) {
      const int rl = rr * 4 + rsub;
      const int64_t r = r0 + rl;
      if (r >= r1) break;
      const int k = (int)(r % K);
      const int64_t c = r / K;
      const int s = (int)(c / B), b = (int)(c % B);
      const int64_t o = (((int64_t)s * T + t) * B + b) * Z + zi;
      const float4 mu = *reinterpret_cast<const float4*>(a.infer_mean + o);
      float4 z = mu;
      if (sampled) {
        const float4 sd = *reinterpret_cast<const float4*>(a.infer_std + o);
        float e[4];
        if (a.noise.eps != nullptr) {
          const float* ep = a.noise.eps + ((((int64_t)s * T + t) * B + b) * K + k) * Z + zi;
          e[0] = ep[0]; e[1] = ep[1]; e[2] = ep[2]; e[3] = ep[3];
        } else {
          normal4(noise_seed(a.noise), a.noise.stream_id, (unsigned)s, (unsigned)t, (unsigned)b + a.noise.b_offset,
                  (unsigned)k, (unsigned)(zi >> 2), e);
        }
        z.x = fmaf(e[0], sd.x, mu.x); z.y = fmaf(e[1], sd.y, mu.y);
        z.z = fmaf(e[2], sd.z, mu.z); z.w = fmaf(e[3], sd.w, mu.w);
      }
      *reinterpret_cast<float4*>(p.zrows + r * Z + zi) = z;
      tile[zl][(rl ^ zl) & 31] = z.x; tile[zl + 1][(rl ^ (zl + 1)) & 31] = z.y;
      tile[zl + 2][(rl ^ (zl + 2)) & 31] = z.z; tile[zl + 3][(rl ^ (zl + 3)) & 31] = z.w;
    }
  }
  if (p.zrowsT == nullptr) return;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = r0 + lane;
  if (r < r1) {
    for (int zz = warp; zz < 64; zz += 2) {
      const int zg = blockIdx.y * 64 + zz;
      if (zg >= Z) break;
      p.zrowsT[(int64_t)zg * p.R + r] = tile[zz][(lane ^ zz) & 31];
    }
  }
}

// one filtering step for every (chain, zi)
__global__ void __launch_bounds__(128) step_kernel(const __grid_constant__ StepParams p) {
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T;
  const int t = gen_pass_time(p.i, T, a.direction);
  const float inv_k = 1.f / (float)K;
  const int64_t n_chains = (int64_t)a.S * B;
  float kl_sum = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_chains * Z;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int zi = (int)(idx % Z);
    const int64_t c = idx / Z;
    const int s = (int)(c / B), b = (int)(c % B);
    const unsigned bits = a.set_expert_bits[s];
    const float gm = p.z0_mean[zi], gs = expf(p.z0_log_std[zi]) + p.min_std;      // models/dmm.py:126-127
    float pm, ps;
    if (p.i == 0) { pm = gm; ps = gs; }
    else {
      float sm = 0.f, sv = 0.f, sq = 0.f;
      for (int k = 0; k < K; ++k) {
        const int64_t r = row64(c * K + k, zi, Z, p.swz);
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
      if (ex.mask != nullptr) m = ex.mask[s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b] != 0;
      if (ex.zero_mask_last_t && t == T - 1) m = false;
      const float w = m ? 1.f : 0.f;
      float mean, std;
      if (ex.kind == BFVI_EXPERT_INV_PRIOR) { mean = gm; std = -gs; }
      else {
        const int64_t off = s * ex.stride_s + t * ex.stride_t + b * ex.stride_b + zi;
        mean = ex.mean[off]; std = ex.std[off];
      }
      const float te = __fmul_rn(poe_prec(std), w);
      S = __fadd_rn(S, te);
      N = __fadd_rn(N, __fmul_rn(__fmul_rn(mean, w), te));
    }
    const float mq = __fdiv_rn(N, S);
    const float mu = (mq != mq) ? 0.f : mq;
    const float sd = __fsqrt_rn(__fdiv_rn(1.f, S));
    const int64_t o = (((int64_t)s * T + t) * B + b) * Z + zi;
    a.infer_mean[o] = mu; a.infer_std[o] = sd;
    a.prior_mean[o] = pm; a.prior_std[o] = ps;
    if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b]))
      kl_sum += kld_elem_fast(mu, sd, pm, ps);
    // particles of this step: input rows of the next transition and the `samples` output
    const bool sampled = gen_samples(a, p.i);
    float se = 0.f;
    for (int k = 0; k < K; ++k) {
      const float z = sampled ? fmaf(eps_at(a.noise, s, t, b, k, zi, T, B, K, Z), sd, mu) : mu;
      if (p.zrows != nullptr) p.zrows[(c * K + k) * Z + zi] = z;
      se += z;
    }
    if (a.samples != nullptr) {
      a.samples[o] = se * inv_k;
      if (p.samplesT != nullptr) p.samplesT[((int64_t)s * Z + zi) * ((int64_t)T * B) + (int64_t)t * B + b] = se * inv_k;
    }
  }
  if (a.loss_acc != nullptr && a.kl_weight != 0.f) block_reduce_add_double(kl_sum * a.kl_weight, a.loss_acc);
}

// step_kernel for Z % 4 == 0: a thread owns FOUR consecutive components of a chain — the GTF heads arrive as
// 128-bit loads, and the K new particles cost one Philox call each instead of four (the scalar kernel draws a
// whole normal4 per component and keeps one value: at K = 25 the generator was most of its 38 us).  Same
// per-component arithmetic in the same order: results are bit-identical to step_kernel.
__global__ void __launch_bounds__(128) step4_kernel(const __grid_constant__ StepParams p) {
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T, Q = Z >> 2;
  const int t = gen_pass_time(p.i, T, a.direction);
  const float inv_k = 1.f / (float)K;
  const int64_t n_chains = (int64_t)a.S * B;
  float kl_sum = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_chains * Q;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int zi = (int)(idx % Q) * 4;
    const int64_t c = idx / Q;
    const int s = (int)(c / B), b = (int)(c % B);
    const unsigned bits = a.set_expert_bits[s];
    float gm[4], gs[4], pm[4], ps[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { gm[j] = p.z0_mean[zi + j]; gs[j] = expf(p.z0_log_std[zi + j]) + p.min_std; }
    if (p.i == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { pm[j] = gm[j]; ps[j] = gs[j]; }
    } else {
      float sm[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < K; ++k) {
        const int64_t r = row64(c * K + k, zi, Z, p.swz);
        const float4 g4 = *reinterpret_cast<const float4*>(p.g + r), n4 = *reinterpret_cast<const float4*>(p.nl + r);
        const float4 l4 = *reinterpret_cast<const float4*>(p.lin + r), a4 = *reinterpret_cast<const float4*>(p.as + r);
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, nv[4] = {n4.x, n4.y, n4.z, n4.w};
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gate = sigmoid_f(gv[j]);
          const float qm = fmaf(gate, nv[j] - lv[j], lv[j]);
          const float qs = softplus_f(av[j]) + p.min_std;
          float m_k, s_k;
          poe2_forward(gm[j], gs[j], qm, qs, m_k, s_k);
          if (K == 1) { sm[j] = m_k; sv[j] = s_k; }
          else { sm[j] += m_k; sv[j] = fmaf(s_k, s_k, sv[j]); sq[j] = fmaf(m_k, m_k, sq[j]); }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (K == 1) { pm[j] = sm[j]; ps[j] = sv[j]; }
        else {                                              // models/dgts.py:78-83
          pm[j] = sm[j] * inv_k;
          ps[j] = sqrtf(sv[j] * inv_k + (sq[j] * inv_k - pm[j] * pm[j]));
        }
      }
    }
    // product of experts, prior first then the experts in order (models/dgts.py:40-51)
    float S[4], N[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { S[j] = poe_prec(ps[j]); N[j] = __fmul_rn(pm[j], S[j]); }
    for (int e = 0; e < a.n_experts; ++e) {
      if (!((bits >> e) & 1u)) continue;
      const bfvi_expert& ex = a.experts[e];
      bool m = true;
      if (ex.mask != nullptr) m = ex.mask[s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b] != 0;
      if (ex.zero_mask_last_t && t == T - 1) m = false;
      const float w = m ? 1.f : 0.f;
      const int64_t off = ex.kind == BFVI_EXPERT_INV_PRIOR ? 0 : s * ex.stride_s + t * ex.stride_t + b * ex.stride_b + zi;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float mean, std;
        if (ex.kind == BFVI_EXPERT_INV_PRIOR) { mean = gm[j]; std = -gs[j]; }
        else { mean = ex.mean[off + j]; std = ex.std[off + j]; }
        const float te = __fmul_rn(poe_prec(std), w);
        S[j] = __fadd_rn(S[j], te);
        N[j] = __fadd_rn(N[j], __fmul_rn(__fmul_rn(mean, w), te));
      }
    }
    const int64_t o = (((int64_t)s * T + t) * B + b) * Z + zi;
    const bool kl_on = a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b]);
    float mu[4], sd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mq = __fdiv_rn(N[j], S[j]);
      mu[j] = (mq != mq) ? 0.f : mq;
      sd[j] = __fsqrt_rn(__fdiv_rn(1.f, S[j]));
      a.infer_mean[o + j] = mu[j]; a.infer_std[o + j] = sd[j];
      a.prior_mean[o + j] = pm[j]; a.prior_std[o + j] = ps[j];
      if (kl_on) kl_sum += kld_elem_fast(mu[j], sd[j], pm[j], ps[j]);
    }
    // particles of this step: input rows of the next transition and the `samples` output
    const bool sampled = gen_samples(a, p.i);
    float se[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k) {
      float z[4] = {mu[0], mu[1], mu[2], mu[3]};
      if (sampled) {
        float e[4];
        if (a.noise.eps != nullptr) {
          const float* ep = a.noise.eps + ((((int64_t)s * T + t) * B + b) * K + k) * Z + zi;
          e[0] = ep[0]; e[1] = ep[1]; e[2] = ep[2]; e[3] = ep[3];
        } else {
          normal4(noise_seed(a.noise), a.noise.stream_id, (unsigned)s, (unsigned)t, (unsigned)b + a.noise.b_offset,
                  (unsigned)k, (unsigned)(zi >> 2), e);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = fmaf(e[j], sd[j], mu[j]);
      }
      if (p.zrows != nullptr) *reinterpret_cast<float4*>(p.zrows + (c * K + k) * Z + zi) = make_float4(z[0], z[1], z[2], z[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) se[j] += z[j];
    }
    if (a.samples != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a.samples[o + j] = se[j] * inv_k;
        if (p.samplesT != nullptr)
          p.samplesT[((int64_t)s * Z + zi + j) * ((int64_t)T * B) + (int64_t)t * B + b] = se[j] * inv_k;
      }
    }
  }
  if (a.loss_acc != nullptr && a.kl_weight != 0.f) block_reduce_add_double(kl_sum * a.kl_weight, a.loss_acc);
}

// ---------------------------------------------------------------------------------------
// backward of one step, part 1 — everything per (chain, zi) ABOVE the transition: upstream
// gradients, the fused KL term, the product of experts.  Writes d_pm / d_v for the particle
// kernel (bwd_rows_kernel); expert gradients are scattered with atomics.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bwd_head_kernel(const __grid_constant__ StepParams p) {
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T;
  const int t = gen_pass_time(p.i, T, a.direction);
  const int64_t n_chains = (int64_t)a.S * B;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_chains * Z;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int zi = (int)(idx % Z);
    const int64_t c = idx / Z;
    const int s = (int)(c / B), b = (int)(c % B);
    const unsigned bits = a.set_expert_bits[s];
    const float gm = p.z0_mean[zi], gs = expf(p.z0_log_std[zi]) + p.min_std;
    const int64_t o = (((int64_t)s * T + t) * B + b) * Z + zi;
    const float mu = a.infer_mean[o], sd = a.infer_std[o], pm = a.prior_mean[o], ps = a.prior_std[o];
    float d_mu = p.c_mu[idx] + (a.d_infer_mean ? a.d_infer_mean[o] : 0.f);
    float d_sd = p.c_sd[idx] + (a.d_infer_std ? a.d_infer_std[o] : 0.f);
    float d_pm = a.d_prior_mean ? a.d_prior_mean[o] : 0.f;
    float d_ps = a.d_prior_std ? a.d_prior_std[o] : 0.f;
    float d_gm = 0.f, d_gs = 0.f;
    if (a.d_samples != nullptr) {                         // samples = mean_k (mu + eps_k sd)
      const float ds = a.d_samples[o];
      d_mu += ds;
      if (gen_samples(a, p.i)) {
        float me = 0.f;
        for (int k = 0; k < K; ++k) me += eps_at(a.noise, s, t, b, k, zi, T, B, K, Z);
        d_sd = fmaf(ds, me / (float)K, d_sd);
      }
    }
    if (a.kl_weight != 0.f && (a.seq_mask == nullptr || a.seq_mask[t * B + b])) {
      float g1, g2, g3, g4;
      kld_elem_grad(mu, sd, pm, ps, a.kl_weight, g1, g2, g3, g4);
      d_mu += g1; d_sd += g2; d_pm += g3; d_ps += g4;
    }
    // product of experts backward (models/dgts.py:40-51)
    const float inv_s = sd * sd;
    const float d_n = d_mu * inv_s;
    const float d_s = -d_mu * mu * inv_s - 0.5f * d_sd * sd * inv_s;
    {
      const float tp = poe_prec(ps);
      d_pm += d_n * tp;
      d_ps += (d_n * pm + d_s) * poe_prec_grad(ps, tp);
    }
    for (int e = 0; e < a.n_experts; ++e) {
      if (!((bits >> e) & 1u)) continue;
      const bfvi_expert& ex = a.experts[e];
      bool m = true;
      if (ex.mask != nullptr) m = ex.mask[s * ex.mstride_s + t * ex.mstride_t + b * ex.mstride_b] != 0;
      if (ex.zero_mask_last_t && t == T - 1) m = false;
      if (!m) continue;
      if (ex.kind == BFVI_EXPERT_INV_PRIOR) {
        const float std = -gs, te = poe_prec(std);
        d_gm += d_n * te;
        d_gs -= (d_n * gm + d_s) * poe_prec_grad(std, te);
      } else if (ex.d_mean != nullptr) {
        const int64_t off = s * ex.stride_s + t * ex.stride_t + b * ex.stride_b + zi;
        const float mean = ex.mean[off], std = ex.std[off], te = poe_prec(std);
        atomicAdd(ex.d_mean + off, d_n * te);
        atomicAdd(ex.d_std + off, (d_n * mean + d_s) * poe_prec_grad(std, te));
      }
    }
    if (p.i == 0) { d_gm += d_pm; d_gs += d_ps; }          // the first prior is the global prior
    else { p.d_pm[idx] = d_pm; p.d_v[idx] = d_ps * 0.5f / ps; }
    if (d_gm != 0.f) atomicAdd(p.g_z0_mean + zi, d_gm);
    if (d_gs != 0.f) atomicAdd(p.g_z0_log_std + zi, d_gs * expf(p.z0_log_std[zi]));
  }
}

// backward of one step, part 2 — per particle row: mixture moment matching, the product
// with the global prior and the four GTF heads (gate sigmoid, mean mix, softplus std).
// Thread = one latent component zi walking a chunk of rows (coalesced along zi); writes the
// pre-activation gradients and their transposed copies, accumulates bias gradients.
constexpr int kRowsPerBlock = 32;
constexpr int kRowsZ = 64;          // latent components per block of bwd_rows_kernel (blockDim.x)
constexpr int kRowsSplit = 4;       // row groups per block (blockDim.y): 8 rows per thread instead of 32 — the
                                    // single-particle passes launch only rows / 32 blocks and were latency-bound
                                    // on the 32-row serial walk (35 us per launch for 2 304 rows)
__global__ void __launch_bounds__(kRowsZ * kRowsSplit) bwd_rows_kernel(const __grid_constant__ StepParams p) {
  // transposed copies leave through shared memory: tile[a][zi][(r ^ zi) & 31] is conflict-free for the
  // writer (thread = zi, fixed r) and the reader (lane = r, fixed zi), and the reader stores 32 consecutive
  // rows = one 128-byte line per instruction (a direct store is one 4-byte transaction per element)
  __shared__ float tile[3][kRowsZ][kRowsPerBlock];
  __shared__ float part[kRowsSplit][6][kRowsZ];
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int zi = blockIdx.y * kRowsZ + tx;
  const bool active = zi < Z;
  const int t = gen_pass_time(p.i, T, a.direction);
  const int64_t r0 = (int64_t)blockIdx.x * kRowsPerBlock;
  const int64_t r1 = r0 + kRowsPerBlock < p.R ? r0 + kRowsPerBlock : p.R;
  constexpr int kPer = kRowsPerBlock / kRowsSplit;
  float d_gm = 0.f, d_gs = 0.f, b_s = 0.f, b_g = 0.f, b_l = 0.f, b_n = 0.f;
  float m_g = 0.f, m_n = 0.f, m_a = 0.f;
  if (active) {
    const float gm = p.z0_mean[zi], gs = expf(p.z0_log_std[zi]) + p.min_std;
    const float inv_k = 1.f / (float)K;
    const int64_t ra = r0 + ty * kPer, rb = ra + kPer < r1 ? ra + kPer : r1;
    for (int64_t r = ra; r < rb; ++r) {
      const int64_t c = r / K;
      const int s = (int)(c / B), b = (int)(c % B);
      const int64_t o = (((int64_t)s * T + t) * B + b) * Z + zi;
      const float pm = a.prior_mean[o];
      const float d_pm = p.d_pm[c * Z + zi], d_v = p.d_v[c * Z + zi];
      const int64_t q = r * Z + zi, qh = row64(r, zi, Z, p.swz);
      const float gate = sigmoid_f(p.g[qh]), nl = p.nl[qh], lin = p.lin[qh], as = p.as[qh];
      const float qm = fmaf(gate, nl - lin, lin), qs = softplus_f(as) + p.min_std;
      float m_k, s_k, g_gm, g_gs, d_qm, d_qs;
      poe2_forward(gm, gs, qm, qs, m_k, s_k);
      const float d_mk = (d_pm + 2.f * d_v * (m_k - pm)) * inv_k;      // models/dgts.py:78-83
      const float d_sk = 2.f * d_v * s_k * inv_k;
      poe2_backward(gm, gs, qm, qs, m_k, s_k, d_mk, d_sk, g_gm, g_gs, d_qm, d_qs);
      d_gm += g_gm; d_gs += g_gs;
      const float d_as = d_qs * softplus_grad(as);
      const float d_nl = d_qm * gate;
      const float d_lin = d_qm - d_nl;                                   // d_qm * (1 - gate)
      const float d_g = d_lin * gate * (nl - lin);                       // through the sigmoid
      p.d_as[q] = d_as; p.d_g[q] = d_g; p.d_lin[q] = d_lin; p.d_nl[q] = d_nl;
      const int col = ((int)(r - r0) ^ tx) & (kRowsPerBlock - 1);
      tile[0][tx][col] = d_as; tile[1][tx][col] = d_g; tile[2][tx][col] = d_lin;
      b_s += d_as; b_g += d_g; b_l += d_lin; b_n += d_nl;
      m_g = fmaxf(m_g, fabsf(d_g)); m_n = fmaxf(m_n, fabsf(d_nl)); m_a = fmaxf(m_a, fabsf(d_as));
    }
  }
  if (p.gmax != nullptr) {                       // (NaN never wins fmaxf: a NaN gradient leaves the maxima alone)
    const float mm[3] = {m_g, m_n, m_a};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float m = mm[k];
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      // test before the atomic: after the first blocks almost nobody raises the maximum (same-address atomics serialise)
      if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && __float_as_uint(m) > *(volatile unsigned*)(p.gmax + k))
        atomicMax(p.gmax + k, __float_as_uint(m));
    }
  }
  part[ty][0][tx] = b_s; part[ty][1][tx] = b_g; part[ty][2][tx] = b_l; part[ty][3][tx] = b_n;
  part[ty][4][tx] = d_gm; part[ty][5][tx] = d_gs;
  __syncthreads();
  if (active && ty == 0) {                       // one set of atomics per component and block
    float v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      v[j] = 0.f;
#pragma unroll
      for (int g = 0; g < kRowsSplit; ++g) v[j] += part[g][j][tx];
    }
    atomicAdd(p.gb_std + zi, v[0]); atomicAdd(p.gb_gate2 + zi, v[1]);
    atomicAdd(p.gb_lin + zi, v[2]); atomicAdd(p.gb_nonlin2 + zi, v[3]);
    if (v[4] != 0.f) atomicAdd(p.g_z0_mean + zi, v[4]);
    if (v[5] != 0.f) atomicAdd(p.g_z0_log_std + zi, v[5] * expf(p.z0_log_std[zi]));
  }
  const int tid = ty * kRowsZ + tx;
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t r = r0 + lane;
  if (r < r1) {
    for (int zl = warp; zl < kRowsZ; zl += kRowsZ * kRowsSplit / 32) {
      const int zz = blockIdx.y * kRowsZ + zl;
      if (zz >= Z) break;
      const int col = (lane ^ zl) & (kRowsPerBlock - 1);
      const int64_t qt = (int64_t)zz * p.R + r;
      p.d_asT[qt] = tile[0][zl][col]; p.d_gT[qt] = tile[1][zl][col]; p.d_linT[qt] = tile[2][zl][col];
    }
  }
}

// backward of one step, part 3 — particles of the previous step back to its (mu, sd):
// c_mu = sum_k dz, c_sd = sum_k dz * eps   (z = mu + eps * sd)
__global__ void __launch_bounds__(128) bwd_carry_kernel(const __grid_constant__ StepParams p, int i_prev) {
  const bfvi_filter_args& a = p.a;
  const int Z = p.Z, K = a.n_particles, B = a.B, T = a.T;
  const int t = gen_pass_time(i_prev, T, a.direction);
  const bool sampled = gen_samples(a, i_prev);
  const int64_t n_chains = (int64_t)a.S * B;
  if (p.gmax != nullptr && blockIdx.x == 0 && threadIdx.x < 3) p.gmax[threadIdx.x] = 0u;   // next time step starts afresh
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_chains * Z;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int zi = (int)(idx % Z);
    const int64_t c = idx / Z;
    const int s = (int)(c / B), b = (int)(c % B);
    float cm = 0.f, cs = 0.f;
    for (int k = 0; k < K; ++k) {
      const int64_t qd = row64(c * K + k, zi, Z, p.swz);
      const float d = p.dz[qd] + (p.dz2 != nullptr ? p.dz2[(c * K + k) * Z + zi] : 0.f);
      cm += d;
      if (sampled) cs = fmaf(d, eps_at(a.noise, s, t, b, k, zi, T, B, K, Z), cs);
    }
    p.c_mu[idx] = cm; p.c_sd[idx] = cs;
  }
}

// ---------------------------------------------------------------------------------------
// GaussianMLP heads
// ---------------------------------------------------------------------------------------
// decoder: Gaussian NLL forward + backward at the heads (models/losses.py:68-89).  In place:
// mean -> d_mean, stdpre -> d_stdpre (gradient at the pre-softplus std), transposed copies
// for the weight gradients, bias gradients.  Thread = one feature d walking a chunk of rows.
struct HeadParams {
  float* mean; float* stdpre;        // (n_rows, D) in: values; out: gradients
  float* meanT; float* stdpreT;      // (D, n_rows) transposed gradients
  const float* target;               // decoder: (n_rows, D), NaN = unobserved
  const uint8_t* row_mask;           // decoder: (n_rows) nullable
  const float* d_mean_in; const float* d_std_in;   // encoder: upstream gradients (n_rows, D)
  float* gb_mean; float* gb_std;     // bias gradient slots (D each)
  int64_t n_rows;
  int D;
  float weight;
  double* loss_acc;
};
__global__ void __launch_bounds__(128) head_kernel(const __grid_constant__ HeadParams p) {
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  float loss = 0.f, b_m = 0.f, b_s = 0.f;
  if (d < p.D) {
    const int64_t r0 = (int64_t)blockIdx.x * kRowsPerBlock;
    const int64_t r1 = r0 + kRowsPerBlock < p.n_rows ? r0 + kRowsPerBlock : p.n_rows;
    for (int64_t r = r0; r < r1; ++r) {
      const int64_t q = r * p.D + d;
      const float a_s = p.stdpre[q];
      float d_m = 0.f, d_p = 0.f;
      if (p.target != nullptr) {                      // decoder + NLL
        const float xt = p.target[q];
        if (xt == xt && (p.row_mask == nullptr || p.row_mask[r] != 0)) {
          const float mean = p.mean[q], std = softplus_f(a_s) + kMlpMinStd;
          loss += nll_gauss_elem(mean, std, xt);
          float g_std;
          nll_gauss_elem_grad(mean, std, xt, p.weight, d_m, g_std);
          d_p = g_std * softplus_grad(a_s);
        }
      } else {                                        // encoder: chain rule through the softplus
        d_m = p.d_mean_in[q];
        d_p = p.d_std_in[q] * softplus_grad(a_s);
      }
      p.mean[q] = d_m; p.stdpre[q] = d_p;
      p.meanT[(int64_t)d * p.n_rows + r] = d_m; p.stdpreT[(int64_t)d * p.n_rows + r] = d_p;
      b_m += d_m; b_s += d_p;
    }
    atomicAdd(p.gb_mean + d, b_m); atomicAdd(p.gb_std + d, b_s);
  }
  if (p.loss_acc != nullptr) block_reduce_add_double(loss * p.weight, p.loss_acc);
}

// plain 2-D transpose (rows x cols) -> (cols x rows), used for per-step weight transposes
// and the encoder inputs
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int64_t rows, int cols,
                                                        float* __restrict__ out, int swz = 0) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j;
    const int c = c0 + tx;
    tile[j][tx] = (r < rows && c < cols) ? in[row64(r, c, cols, swz)] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t r = r0 + tx;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = tile[tx][j];
  }
}

// ---------------------------------------------------------------------------------------------------------
// elementwise pieces of the per-modality MLPs of the composed path (bfvi_mlp_fwd / bfvi_mlp_bwd: GaussianMLP and
// CategoricalMLP at any size, the categorical encoder's Embedding -> ReLU; models/common.py:9-41, models/dmm.py:78-82)
// ---------------------------------------------------------------------------------------------------------
// class index of a row: the reference casts the (NaN-zero-filled) float label with .long() (models/dmm.py:96-98)
__device__ __forceinline__ int label_of(const float* __restrict__ labels, int64_t r, int n_classes) {
  const float v = labels[r];
  int c = (v != v) ? 0 : (int)v;
  return c < 0 ? 0 : (c >= n_classes ? n_classes - 1 : c);
}
// out[r, :] = relu(emb[label[r], :])
__global__ void __launch_bounds__(256) embed_relu_kernel(const float* __restrict__ emb, const float* __restrict__ labels,
                                                         int64_t n_rows, int H, int n_classes, float* __restrict__ out) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_rows * H;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / H;
    const int j = (int)(idx % H);
    const float v = emb[(int64_t)label_of(labels, r, n_classes) * H + j];
    out[idx] = v < 0.f ? 0.f : v;
  }
}
// d_emb[label[r], :] += de[r, :]  (de already carries the ReLU mask)
__global__ void __launch_bounds__(256) embed_bwd_kernel(const float* __restrict__ de, const float* __restrict__ labels,
                                                        int64_t n_rows, int H, int n_classes, float* __restrict__ d_emb) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_rows * H;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / H;
    const int j = (int)(idx % H);
    const float g = de[idx];
    if (g != 0.f) atomicAdd(d_emb + (int64_t)label_of(labels, r, n_classes) * H + j, g);
  }
}
// softmax over the columns of every row, in place (one warp per row; torch.softmax: max-shifted)
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, int64_t n_rows, int D) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * wpb) {
    float* row = x + r * D;
    float mx = -INFINITY;
    for (int j = lane; j < D; j += 32) mx = fmaxf(mx, row[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < D; j += 32) sum += expf(row[j] - mx);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int j = lane; j < D; j += 32) row[j] = expf(row[j] - mx) * inv;
  }
}
// head backward (one warp per row), from the forward OUTPUTS:
//  Gaussian head: d_a = d_mean, d_b = d_std * sigmoid(pre) with sigmoid(pre) = 1 - exp(-(std - min_std))
//                 (std = softplus(pre) + min_std, models/common.py:41; torch's softplus is the identity above 20)
//  softmax head:  d_a = p * (d_p - sum_j d_p[j] p[j])
__global__ void __launch_bounds__(256) mlp_head_bwd_kernel(const float* __restrict__ out_a, const float* __restrict__ out_b,
                                                           const float* __restrict__ d_out_a, const float* __restrict__ d_out_b,
                                                           int64_t n_rows, int D, int softmax, float min_std,
                                                           float* __restrict__ d_a, float* __restrict__ d_b) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * wpb) {
    if (softmax) {
      float dot = 0.f;
      for (int j = lane; j < D; j += 32) dot = fmaf(d_out_a[r * D + j], out_a[r * D + j], dot);
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      for (int j = lane; j < D; j += 32) d_a[r * D + j] = out_a[r * D + j] * (d_out_a[r * D + j] - dot);
    } else {
      for (int j = lane; j < D; j += 32) {
        d_a[r * D + j] = d_out_a ? d_out_a[r * D + j] : 0.f;
        const float sp = out_b[r * D + j] - min_std;                 // softplus(pre)
        const float sg = sp > 20.f ? 1.f : -expm1f(-sp);             // sigmoid(pre)
        d_b[r * D + j] = d_out_b ? d_out_b[r * D + j] * sg : 0.f;
      }
    }
  }
}
// out[j] += sum_r x[r, j]  (bias gradients): thread = column, blockIdx.y = slice of the rows
__global__ void __launch_bounds__(128) colsum_kernel(const float* __restrict__ x, int64_t n_rows, int D, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  const int64_t per = (n_rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per < n_rows ? r0 + per : n_rows;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += x[r * D + j];
  if (r1 > r0) atomicAdd(out + j, s);
}

// prior-matching term (models/dmm.py:496-501,541-545), one direction: moments of the K
// propagated particles of the global prior, KL(global || next), and its gradient wired into
// the same d_pm / d_v slots bwd_rows_kernel reads (a single chain, no experts)
struct MatchHeadParams {
  const float* z0_mean; const float* z0_log_std; float* g_z0_mean; float* g_z0_log_std;
  const float* g; const float* nl; const float* lin; const float* as;   // (K, Z) GTF heads
  float* pm; float* d_pm; float* d_v;                                    // (Z)
  float min_std, coef_static;
  int swz;                 // heads in the swz64 layout (fused path)
  const float* count;      // device scalar mask.sum() (nullable)
  double* loss_acc;
  int K, Z, with_grad;
};
__global__ void __launch_bounds__(128) match_head_kernel(const __grid_constant__ MatchHeadParams p) {
  const int zi = blockIdx.x * blockDim.x + threadIdx.x;
  float kl = 0.f;
  if (zi < p.Z) {
    const float coef = p.coef_static * (p.count != nullptr ? p.count[0] : 1.f);
    const float gm = p.z0_mean[zi], gs = expf(p.z0_log_std[zi]) + p.min_std;
    float sm = 0.f, sv = 0.f, sq = 0.f;
    for (int k = 0; k < p.K; ++k) {
      const int64_t q = row64(k, zi, p.Z, p.swz);
      const float gate = sigmoid_f(p.g[q]), nl = p.nl[q], lin = p.lin[q];
      const float qm = fmaf(gate, nl - lin, lin), qs = softplus_f(p.as[q]) + p.min_std;
      float m_k, s_k;
      poe2_forward(gm, gs, qm, qs, m_k, s_k);
      sm += m_k; sv = fmaf(s_k, s_k, sv); sq = fmaf(m_k, m_k, sq);
    }
    const float inv_k = 1.f / (float)p.K;
    const float nm = sm * inv_k, ns = sqrtf(sv * inv_k + (sq * inv_k - nm * nm));
    kl = coef * kld_elem(gm, gs, nm, ns);
    if (p.with_grad) {
      float g1, g2, d_nm, d_ns;
      kld_elem_grad(gm, gs, nm, ns, coef, g1, g2, d_nm, d_ns);
      p.pm[zi] = nm; p.d_pm[zi] = d_nm; p.d_v[zi] = d_ns * 0.5f / ns;
      atomicAdd(p.g_z0_mean + zi, g1);
      atomicAdd(p.g_z0_log_std + zi, g2 * expf(p.z0_log_std[zi]));
    }
  }
  if (p.loss_acc != nullptr) block_reduce_add_double(kl, p.loss_acc);
}

}  // namespace gen
}  // namespace bfvi
