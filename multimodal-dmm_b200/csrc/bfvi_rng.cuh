// bfvi_rng.cuh — counter-based reparameterisation noise (throughput mode).
//
// Replaces `torch.FloatTensor(size).normal_()` of MultiDGTS._sample_gauss
// (models/dgts.py:177-180) when no external noise tensor is given.  Philox4x32-10
// keyed by the 64-bit seed; the 128-bit counter is (b, k | chunk<<24, t, s |
// stream_id<<16), so a draw is a pure function of its logical index and the
// backward pass can regenerate it instead of storing it.  bfvi_dump_noise runs the
// SAME function to hand the identical stream to the oracle.
#pragma once
#include "bfvi_math.cuh"

namespace bfvi {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float u01(unsigned x) {            // (0, 1)
  return __fmaf_rn((float)(x >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);
}

// four N(0,1) draws: components 4*chunk .. 4*chunk+3 of particle k
__device__ __forceinline__ void normal4(uint64_t seed, unsigned stream_id, unsigned s, unsigned t,
                                        unsigned b, unsigned k, unsigned chunk, float (&out)[4]) {
  const uint4 r = philox4x32_10(make_uint4(b, k | (chunk << 24), t, s | (stream_id << 16)),
                                uint2{(unsigned)seed, (unsigned)(seed >> 32)});
  // Box-Muller radius sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u) on the MUFU path
  const float r0 = fast_sqrt(-1.3862943611198906f * fast_lg2(u01(r.x)));
  const float r1 = fast_sqrt(-1.3862943611198906f * fast_lg2(u01(r.z)));
  float s0, c0, s1, c1;
  __sincosf(__fmul_rn(6.283185307179586f, u01(r.y)), &s0, &c0);
  __sincosf(__fmul_rn(6.283185307179586f, u01(r.w)), &s1, &c1);
  out[0] = __fmul_rn(r0, c0);
  out[1] = __fmul_rn(r0, s0);
  out[2] = __fmul_rn(r1, c1);
  out[3] = __fmul_rn(r1, s1);
}

// Z draws for (s, t, b, k): external tensor (S,T,B,K,Z) or the Philox stream.
template <int Z>
__device__ __forceinline__ void load_eps(const float* __restrict__ eps_ext, uint64_t seed,
                                         unsigned stream_id, int s, int t, int b, unsigned b_offset,
                                         int k, int T, int B, int K, float (&e)[Z]) {
  if (eps_ext != nullptr) {
    const float* p = eps_ext + ((((int64_t)s * T + t) * B + b) * K + k) * Z;
#pragma unroll
    for (int i = 0; i < Z; ++i) e[i] = p[i];
  } else {
#pragma unroll
    for (int c = 0; c < (Z + 3) / 4; ++c) {
      float n[4];
      normal4(seed, stream_id, (unsigned)s, (unsigned)t, (unsigned)b + b_offset, (unsigned)k,
              (unsigned)c, n);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * c + j < Z) e[4 * c + j] = n[j];
    }
  }
}

}  // namespace bfvi
