// bfvi_api.cu — the C ABI declared in include/bfvi.h: argument checking, kernel
// dispatch on (z_dim, h_dim), workspace carving and the orchestration of one whole
// MultiDMM.step (models/dmm.py:503-554) + backward as a fixed sequence of launches
// on the caller's stream.  No device allocation, no global mutable state.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "bfvi_small.cuh"
#include "bfvi_tc.cuh"
#include "bfvi_generic.cuh"
#include "bfvi_fused.cuh"
#include "bfvi_data.cuh"
#include "bfvi_internal.h"

namespace {

thread_local std::string g_err;
// Which kernel variants the last top-level call of this host thread launched (bfvi_last_dispatch):
// the parity tests assert that the variant they mean to check is the one that ran.
thread_local std::string g_dispatch;
void note_dispatch(const char* fmt, ...) {
  char buf[160];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (g_dispatch.find(buf) != std::string::npos) return;       // distinct entries only
  if (g_dispatch.size() > 4000) return;
  if (!g_dispatch.empty()) g_dispatch += ';';
  g_dispatch += buf;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

}  // namespace
namespace bfvi {
int report_error(int code, const char* fmt, ...) {        // the other translation units' way to bfvi_last_error()
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace bfvi
namespace {

#define BFVI_CHECK_CUDA()                                                                 \
  do {                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess) return fail(BFVI_ERR_CUDA, "CUDA error: %s (%s:%d)",           \
                                       cudaGetErrorString(e_), __FILE__, __LINE__);       \
  } while (0)

using bfvi::pad4;

int check_model(const bfvi_model* m) {
  if (m == nullptr) return fail(BFVI_ERR_ARG, "model is null");
  if (m->n_mods < 1 || m->n_mods > BFVI_MAX_MODS) return fail(BFVI_ERR_ARG, "n_mods %d out of range", m->n_mods);
  if (m->z_dim < 1 || m->h_dim < 1) return fail(BFVI_ERR_ARG, "bad z_dim/h_dim");
  for (int i = 0; i < m->n_mods; ++i)
    if (m->dims[i] < 1) return fail(BFVI_ERR_ARG, "dims[%d] < 1", i);
  return BFVI_OK;
}

int64_t take(int64_t& cur, int n) { int64_t o = cur; cur += pad4(n); return o; }

void mlp_layout(int64_t& cur, int n_in, int n_out, int H, bfvi_mlp_layout* l) {
  l->begin = cur;
  l->in_to_h_w = take(cur, H * n_in);
  l->in_to_h_b = take(cur, H);
  l->mean_w = take(cur, n_out * H);
  l->mean_b = take(cur, n_out);
  l->std_w = take(cur, n_out * H);
  l->std_b = take(cur, n_out);
  l->end = cur;
}

void gtf_layout(int64_t& cur, int Z, int H, bfvi_gtf_layout* l) {
  l->begin = cur;
  l->gate0_w = take(cur, H * Z); l->gate0_b = take(cur, H);
  l->gate2_w = take(cur, Z * H); l->gate2_b = take(cur, Z);
  l->lin_w = take(cur, Z * Z);   l->lin_b = take(cur, Z);
  l->nonlin0_w = take(cur, H * Z); l->nonlin0_b = take(cur, H);
  l->nonlin2_w = take(cur, Z * H); l->nonlin2_b = take(cur, Z);
  l->std_w = take(cur, Z * Z);   l->std_b = take(cur, Z);
  l->end = cur;
}

bfvi::MlpOffsets mlp_offsets(const bfvi_mlp_layout& l) {
  bfvi::MlpOffsets o;
  o.w1 = (int)(l.in_to_h_w - l.begin); o.b1 = (int)(l.in_to_h_b - l.begin);
  o.wm = (int)(l.mean_w - l.begin);    o.bm = (int)(l.mean_b - l.begin);
  o.ws = (int)(l.std_w - l.begin);     o.bs = (int)(l.std_b - l.begin);
  o.size = (int)(l.end - l.begin);
  return o;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs microseconds of host time per call; a step makes ~35 000 launches of
// the large-dim kernels, so the attribute is raised only when a kernel needs more than it was last given (per host thread
// and device)
#ifndef BFVI_EMU
template <typename K>
inline void ensure_dyn_smem(K kernel, size_t bytes) {
  struct Slot { const void* fn; int dev; size_t have; };
  static thread_local Slot tab[64];
  static thread_local int n_slots = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  const void* fn = (const void*)kernel;
  Slot* sl = nullptr;
  for (int i = 0; i < n_slots; ++i)
    if (tab[i].fn == fn && tab[i].dev == dev) { sl = &tab[i]; break; }
  if (sl == nullptr && n_slots < 64) { sl = &tab[n_slots++]; sl->fn = fn; sl->dev = dev; sl->have = 0; }
  if (sl == nullptr || bytes > sl->have) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (sl != nullptr) sl->have = bytes;
  }
}
#endif
int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

int grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
  const int sms = num_sms();
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)(sms > 0 ? sms : 1) * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- (Z, H) dispatch for the register-resident family -------------------------
#define BFVI_SMALL_DIMS(X) X(5, 20) X(4, 8) X(3, 6) X(6, 12)

// 1 = register-resident small-dim family, 2 = tcgen05 large-dim family.  BFVI_FAMILY=2 forces the
// large-dim family on small models (test knob: the golden fixtures then exercise it end to end).
int family_of(int Z, int H);
bool small_supported(int Z, int H) {
#define X(z, h) if (Z == z && H == h) return true;
  BFVI_SMALL_DIMS(X)
#undef X
  return false;
}

#define BFVI_MLP_H(X) X(20) X(8) X(6) X(12)

int family_of(int Z, int H) {
  if (!small_supported(Z, H)) return 2;
  const char* env = getenv("BFVI_FAMILY");
  return (env != nullptr && atoi(env) == 2) ? 2 : 1;
}

template <int Z, int H>
int layout_matches(const bfvi_gtf_layout& l) {
  using L = bfvi::GtfLayout<Z, H>;
  const int64_t b = l.begin;
  return l.gate0_w - b == L::G0W && l.gate0_b - b == L::G0B && l.gate2_w - b == L::G2W &&
         l.gate2_b - b == L::G2B && l.lin_w - b == L::LW && l.lin_b - b == L::LB &&
         l.nonlin0_w - b == L::N0W && l.nonlin0_b - b == L::N0B && l.nonlin2_w - b == L::N2W &&
         l.nonlin2_b - b == L::N2B && l.std_w - b == L::SW && l.std_b - b == L::SB &&
         l.end - b == L::SIZE;
}

// Lane-group geometry for K particles with R rows per thread (bfvi_chain.cuh): among
// the lane counts whose utilisation is within 10 % of the best, take the smallest
// that still yields enough warp tasks to fill the machine, else the most parallel.
void choose_lanes(int K, int R, int64_t chains, int* lanes, int* rounds) {
  double best = 0.0;
  double util[33];
  for (int L = 1; L <= 32; ++L) {
    const int cpw = 32 / L, rd = (K + L * R - 1) / (L * R);
    util[L] = (double)cpw * K / (32.0 * R * rd);
    if (util[L] > best) best = util[L];
  }
  const int sms = num_sms();
  const int64_t want = (int64_t)(sms > 0 ? sms : 1) * 16;
  int pick = 0;
  if (const char* env = getenv("BFVI_LANES")) {          // test / tuning knob
    const int v = atoi(env);
    if (v >= 1 && v <= 32 && K > 1) pick = v;
  }
  for (int L = 1; L <= 32 && !pick; ++L)
    if (util[L] >= 0.9 * best && (chains + 32 / L - 1) / (32 / L) >= want) pick = L;
  if (!pick)
    for (int L = 32; L >= 1 && !pick; --L)
      if (util[L] >= 0.9 * best) pick = L;
  *lanes = pick;
  *rounds = (K + pick * R - 1) / (pick * R);
}

// Few chains (small batches such as the reference's default spirals run, B = 100): the serial
// time loop is latency-bound and a warp per chain with ONE particle per lane minimises the work
// of a step; throughput mappings (5 particles per lane) only pay once the GPU is filled.
bool latency_bound_pass(int64_t chains, int K) {
  if (K <= 1 || getenv("BFVI_LANES") != nullptr) return false;
  const int sms = num_sms() > 0 ? num_sms() : 1;
  return chains <= (int64_t)sms * 16;
}

int task_grid(int64_t chains, int lanes, int warps_per_block) {
  const int cpw = 32 / lanes;
  const int64_t tasks = (chains + cpw - 1) / cpw;
  int64_t blocks = (tasks + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)(num_sms() > 0 ? num_sms() : 1) * 32;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

size_t align_up(size_t v, size_t a);
// scratch of the time-segmented backward filter: per-task flags + the (S, B, 2Z) gradient carry
size_t seg_scratch_bytes(int S, int B, int Z) {
  return align_up(sizeof(int) * (size_t)S * B, 256) + align_up(sizeof(float) * (size_t)S * B * 2 * Z, 256);
}

// Segment count for the time-segmented particle kernels: `tasks` warp tasks on `slots` resident
// warps take ceil(tasks / slots) rounds of a whole task; with sg segments the rounds shrink to
// ceil(tasks * sg / slots) / sg.  Smallest sg within 1 % of the best, if that saves >= 5 %.
int pick_segments(int64_t tasks, int64_t slots, int T, const char* env_name) {
  int n_seg = 1;
  if (slots > 0 && tasks > 0) {
    const double base = (double)((tasks + slots - 1) / slots);
    double best = base;
    for (int sg = 2; sg <= 8 && sg <= T; ++sg) {
      const double r = (double)((tasks * sg + slots - 1) / slots) / sg;
      if (r < 0.99 * best) { best = r; n_seg = sg; }
    }
    if (best > 0.95 * base) n_seg = 1;
  }
  if (const char* env = getenv(env_name)) {                  // tuning / test knob (1 = off)
    const int v = atoi(env);
    if (v >= 1 && v <= 16 && v <= T) n_seg = v;
  }
  return n_seg;
}

// Launch a time-segmented kernel: one cooperative launch serving every segment (all warps resident,
// per-task flags order the segments), else one stream-ordered launch per segment.
template <typename Kernel>
int launch_segmented(Kernel k, bfvi::FilterParams& fp, int n_seg, int64_t tasks, int warps, int64_t resident_blocks,
                     dim3 plain_grid, int threads, size_t smem, cudaStream_t st) {
#ifndef BFVI_EMU
  const char* coop_env = getenv("BFVI_COOPERATIVE");           // 0: stream-ordered launches per segment
  if ((coop_env == nullptr || atoi(coop_env) != 0) && resident_blocks > 0) {
    cudaMemsetAsync(fp.seg_done, 0, sizeof(int) * (size_t)tasks, st);
    fp.seg_lo = 0; fp.seg_hi = n_seg;
    int64_t blocks = (tasks * n_seg + warps - 1) / warps;
    if (blocks > resident_blocks) blocks = resident_blocks;
    void* args[] = {&fp};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)k, dim3((unsigned)blocks), dim3(threads), args, smem, st);
    if (e == cudaSuccess) { note_dispatch("segmented:cooperative seg=%d", n_seg); return BFVI_OK; }
    cudaGetLastError();                                        // not launchable cooperatively here: fall through
  }
#else
  (void)tasks; (void)warps; (void)resident_blocks;
#endif
  note_dispatch("segmented:per-launch seg=%d", n_seg);
  for (int sg = 0; sg < n_seg; ++sg) {                          // same work, ordered by the stream
    fp.seg_lo = sg; fp.seg_hi = sg + 1;
    BFVI_LAUNCH(k, plain_grid, dim3(threads), smem, st, fp);
  }
  return BFVI_OK;
}

template <typename Kernel>
int64_t resident_blocks_of(Kernel k, int threads, size_t smem) {
#ifdef BFVI_EMU
  (void)k; (void)threads; (void)smem;
  return 2;                                                    // exercises the multi-launch form on the CPU
#else
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem) != cudaSuccess) return 0;
  return (int64_t)per_sm * (num_sms() > 0 ? num_sms() : 1);
#endif
}

template <int Z, int H>
int launch_filter_fwd(bfvi::FilterParams fp, cudaStream_t st) {
  const bfvi_filter_args& a = fp.a;
  const int64_t chains = (int64_t)a.S * (fp.bc > 0 ? fp.bc : a.B);
  const int wpb = bfvi::kChainFwdThreads / 32;
  if (latency_bound_pass(chains, a.n_particles)) {
    fp.lanes = a.n_particles < 32 ? a.n_particles : 32;
    fp.rounds = (a.n_particles + fp.lanes - 1) / fp.lanes;
    auto k = bfvi::chain_fwd_kernel<Z, H, 1>;
    note_dispatch("chain_fwd<%d,%d,1> lanes=%d", Z, H, fp.lanes);
    BFVI_LAUNCH(k, dim3(task_grid(chains, fp.lanes, 2)), dim3(64), 0, st, fp);
  } else if (a.n_particles > 1) {
    // (time segmentation as in the backward kernel was measured here and not kept: 2048 tasks fit
    //  the 2368 resident warps in one round, the kernel is issue-bound rather than short of warps,
    //  and a cooperative launch has to wait for pass A's kernels on the side stream: step +0.4 ms)
    constexpr int R = 5;
    choose_lanes(a.n_particles, R, chains, &fp.lanes, &fp.rounds);
    auto k = bfvi::chain_fwd_kernel<Z, H, R>;
    note_dispatch("chain_fwd<%d,%d,%d> lanes=%d", Z, H, R, fp.lanes);
    BFVI_LAUNCH(k, dim3(task_grid(chains, fp.lanes, wpb)), dim3(bfvi::kChainFwdThreads), 0, st, fp);
  } else {
    // single particle: latency-bound; a chain is spread over Z lanes (bfvi_zsplit.cuh), 2-warp CTAs
    // put the few warps on all SMs
    fp.lanes = Z; fp.rounds = 1;
    auto k = bfvi::zsplit_fwd_kernel<Z, H>;
    note_dispatch("zsplit_fwd<%d,%d>", Z, H);
    BFVI_LAUNCH(k, dim3(task_grid(chains, Z, 2)), dim3(bfvi::kZsplitThreads), 0, st, fp);
  }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

template <int Z, int H>
int launch_filter_bwd(bfvi::FilterParams fp, cudaStream_t st) {
  const bfvi_filter_args& a = fp.a;
  const int64_t chains = (int64_t)a.S * (fp.bc > 0 ? fp.bc : a.B);
  // z-split single-particle kernel (bfvi_zsplit.cuh) while all its warps are resident at once
  // (register accumulators: 8 warps/SM); beyond that the one-chain-per-lane kernel below does the
  // same work with 5x fewer warps, leaves room for the other pass running beside it, and wins
  if (a.n_particles == 1) {
    const int64_t warps_needed = (chains + 32 / Z - 1) / (32 / Z);
    const int64_t sms = num_sms() > 0 ? num_sms() : 1;
    const char* force = getenv("BFVI_ZSPLIT_BWD");          // tuning knob: 0 off, 1 registers, 2 shared
    const int mode = force ? atoi(force) : (warps_needed <= sms * 8 ? 1 : 0);
    if (mode == 1 || mode == 2) {
      fp.lanes = Z; fp.slices = 1;
      note_dispatch("zsplit_bwd<%d,%d,%d>", Z, H, mode == 2 ? 1 : 0);
      const dim3 grid(task_grid(chains, Z, 2)), block(bfvi::kZsplitThreads);
      if (mode == 1) { auto k = bfvi::zsplit_bwd_kernel<Z, H, false>; BFVI_LAUNCH(k, grid, block, 0, st, fp); }
      else { auto k = bfvi::zsplit_bwd_kernel<Z, H, true>; BFVI_LAUNCH(k, grid, block, 0, st, fp); }
      BFVI_CHECK_CUDA();
      return BFVI_OK;
    }
  }
  // single-particle passes are latency-bound with few warps: 2-warp CTAs reach all SMs
  const int warps = (a.n_particles > 1 && !latency_bound_pass(chains, a.n_particles)) ? bfvi::kChainBwdWarps : 2;
  const size_t smem = bfvi::chain_bwd_smem_bytes<Z, H>(warps);
  const int threads = warps * 32;
  const bfvi::WgSpec spec = bfvi::GtfPanels<Z, H>::spec();
  if (bfvi::wg_rounds<bfvi::GtfPanels<Z, H>::TD, bfvi::kTX>(spec) != 1)
    return fail(BFVI_ERR_UNSUPPORTED, "internal: weight-gradient tiling needs one round");
  int rounds = 1;
  if (latency_bound_pass(chains, a.n_particles)) fp.lanes = a.n_particles < 32 ? a.n_particles : 32;
  else choose_lanes(a.n_particles, 1, chains, &fp.lanes, &rounds);
  fp.slices = (a.n_particles + fp.lanes - 1) / fp.lanes;
  auto k1 = bfvi::chain_bwd_kernel<Z, H, false>;
  auto k = bfvi::chain_bwd_kernel<Z, H, true>;
  cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // Time segmentation (chain_bwd_kernel): when the warp tasks do not fill a whole number of rounds
  // of the resident warps, cut the time loop into the segment count that packs best.
  const int cpw = 32 / fp.lanes;
  const int64_t tasks = (chains + cpw - 1) / cpw;
  const dim3 grid(task_grid(chains, fp.lanes, warps));
  int n_seg = 1;
  int64_t resident = 0;
  if (a.n_particles > 1 && warps == bfvi::kChainBwdWarps && a.workspace != nullptr &&
      a.workspace_bytes >= seg_scratch_bytes(a.S, a.B, Z)) {
    resident = resident_blocks_of(k, threads, smem);
    n_seg = pick_segments(tasks, resident * warps, a.T, "BFVI_BWD_SEGMENTS");
  }
  note_dispatch("chain_bwd<%d,%d,%d> K=%d lanes=%d warps=%d", Z, H, n_seg == 1 ? 0 : 1, a.n_particles, fp.lanes, warps);
  if (n_seg == 1) {
    BFVI_LAUNCH(k1, grid, dim3(threads), smem, st, fp);
    BFVI_CHECK_CUDA();
    return BFVI_OK;
  }
  // scratch: per-task flags (S * B ints, this chunk's share) then the (S, B, 2Z) carry
  fp.seg_count = n_seg;
  fp.seg_done = reinterpret_cast<int*>(a.workspace) + (size_t)a.S * fp.b0;
  fp.seg_carry = reinterpret_cast<float*>(reinterpret_cast<char*>(a.workspace) + align_up(sizeof(int) * (size_t)a.S * a.B, 256));
  if (int rc = launch_segmented(k, fp, n_seg, tasks, warps, resident, grid, threads, smem, st)) return rc;
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

template <int Z, int H>
int launch_match(const bfvi::MatchParams& mp, cudaStream_t st) {
  auto k = bfvi::match_kernel<Z, H>;
  const size_t smem = bfvi::match_smem_bytes<Z, H>();
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  BFVI_LAUNCH(k, dim3(2), dim3(32), smem, st, mp);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int launch_mlp_fwd(int H, const bfvi::MlpParams& mp, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)mp.off.size;
  const int grid = grid_for(mp.n_rows, 128, 8);
#define X(h)                                                                       \
  if (H == h) {                                                                    \
    auto k = bfvi::mlp_fwd_kernel<h>;                                              \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    BFVI_LAUNCH(k, dim3(grid), dim3(128), smem, st, mp);                           \
    BFVI_CHECK_CUDA();                                                             \
    return BFVI_OK;                                                                \
  }
  BFVI_MLP_H(X)
#undef X
  return fail(BFVI_ERR_UNSUPPORTED, "no MLP kernel for h_dim=%d", H);
}

int launch_mlp_bwd(int H, bool decoder, const bfvi::MlpParams& mp, cudaStream_t st) {
  const size_t smem = bfvi::mlp_bwd_smem_bytes(mp.n_in, mp.n_out, H, mp.off);
  if (smem > 200 * 1024) return fail(BFVI_ERR_UNSUPPORTED, "MLP too wide for the small-dim path");
  const int threads = bfvi::kMlpWarps * 32;
  const int grid = grid_for(mp.n_rows, threads, 2);
#define X(h)                                                                         \
  if (H == h) {                                                                      \
    if (decoder) {                                                                   \
      auto k = bfvi::mlp_bwd_kernel<h, true>;                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      BFVI_LAUNCH(k, dim3(grid), dim3(threads), smem, st, mp);                       \
    } else {                                                                         \
      auto k = bfvi::mlp_bwd_kernel<h, false>;                                       \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      BFVI_LAUNCH(k, dim3(grid), dim3(threads), smem, st, mp);                       \
    }                                                                                \
    BFVI_CHECK_CUDA();                                                               \
    return BFVI_OK;                                                                  \
  }
  BFVI_MLP_H(X)
#undef X
  return fail(BFVI_ERR_UNSUPPORTED, "no MLP kernel for h_dim=%d", H);
}

int dispatch_filter(const bfvi_model* m, const bfvi_layout& lay, bool backward,
                    const bfvi::FilterParams& fp, cudaStream_t st) {
#define X(z, h)                                                                      \
  if (m->z_dim == z && m->h_dim == h) {                                              \
    if (!layout_matches<z, h>(lay.trans[0]))                                         \
      return fail(BFVI_ERR_ARG, "internal: GTF layout mismatch");                    \
    return backward ? launch_filter_bwd<z, h>(fp, st) : launch_filter_fwd<z, h>(fp, st); \
  }
  BFVI_SMALL_DIMS(X)
#undef X
  return fail(BFVI_ERR_UNSUPPORTED, "no filter kernel for z_dim=%d h_dim=%d", m->z_dim, m->h_dim);
}

int dispatch_match(const bfvi_model* m, const bfvi::MatchParams& mp, cudaStream_t st) {
#define X(z, h) if (m->z_dim == z && m->h_dim == h) return launch_match<z, h>(mp, st);
  BFVI_SMALL_DIMS(X)
#undef X
  return fail(BFVI_ERR_UNSUPPORTED, "no match kernel for z_dim=%d h_dim=%d", m->z_dim, m->h_dim);
}

int check_filter_args(const bfvi_model* m, const bfvi_filter_args* a) {
  if (a == nullptr) return fail(BFVI_ERR_ARG, "filter args null");
  if (a->T < 1 || a->B < 1 || a->S < 1 || a->S > BFVI_MAX_SETS) return fail(BFVI_ERR_ARG, "bad T/B/S");
  if (a->n_experts < 0 || a->n_experts > BFVI_MAX_EXPERTS) return fail(BFVI_ERR_ARG, "bad n_experts");
  if (a->n_particles < 1) return fail(BFVI_ERR_ARG, "n_particles < 1");
  if ((int64_t)a->S * a->B > 0x7fffffff / 2) return fail(BFVI_ERR_ARG, "too many chains for one call");
  if (!a->infer_mean || !a->infer_std || !a->prior_mean || !a->prior_std)
    return fail(BFVI_ERR_ARG, "filter outputs must be non-null");
  for (int e = 0; e < a->n_experts; ++e)
    if (a->experts[e].kind == BFVI_EXPERT_TENSOR && (!a->experts[e].mean || !a->experts[e].std))
      return fail(BFVI_ERR_ARG, "expert %d has null tensors", e);
  (void)m;
  return BFVI_OK;
}

bfvi::FilterParams make_filter_params(const bfvi_model* m, const bfvi_layout& lay, const float* params,
                                      float* grads, const bfvi_filter_args* a) {
  bfvi::FilterParams fp;
  fp.a = *a;
  const int d = a->direction == BFVI_DIR_BWD ? 1 : 0;
  fp.trans_w = params + lay.trans[d].begin;
  fp.z0_mean = params + lay.z0_mean;
  fp.z0_log_std = params + lay.z0_log_std;
  fp.g_trans = grads ? grads + lay.trans[d].begin : nullptr;
  fp.g_z0_mean = grads ? grads + lay.z0_mean : nullptr;
  fp.g_z0_log_std = grads ? grads + lay.z0_log_std : nullptr;
  fp.min_std = m->min_std;
  fp.lanes = 1; fp.rounds = 1; fp.slices = 1;
  fp.b0 = 0; fp.bc = 0;
  fp.seg_count = 1; fp.seg_lo = 0; fp.seg_hi = 1; fp.seg_done = nullptr; fp.seg_carry = nullptr;
  return fp;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Lazily created non-blocking side streams + fork/join events per (host thread,
// device).  Inside one step the f_mode filtering ELBO (pass A) is independent of the
// s_mode passes (B, C), and the batch chunks of B / C are independent of each other:
// the latency-bound single-particle kernels of one branch overlap the throughput-bound
// particle kernels of another.  Everything forks from and joins back into the caller's
// stream, so the call stays asynchronous and stream-ordered for the caller.
constexpr int kSideStreams = 4;
struct SideStreams {
  cudaStream_t stream[kSideStreams];
  cudaEvent_t fork, join[kSideStreams];
  bool ok;
};
constexpr int kTileLanes = 2;            // batch tiles in flight (step_large_tiled)
SideStreams* side_streams(int lane = 0) {
  static thread_local SideStreams cache_l[64 * kTileLanes];
  static thread_local bool made_l[64 * kTileLanes];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || lane < 0 || lane >= kTileLanes) return nullptr;
  SideStreams* cache = cache_l + (size_t)lane * 64;
  bool* made = made_l + (size_t)lane * 64;
  if (!made[dev]) {
    made[dev] = true;
    SideStreams& c = cache[dev];
    c.ok = cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kSideStreams && c.ok; ++i)
      c.ok = cudaStreamCreateWithFlags(&c.stream[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming) == cudaSuccess;
  }
  return cache[dev].ok ? &cache[dev] : nullptr;
}

struct StepPlan {
  int S;
  unsigned set_bits[BFVI_MAX_SETS];     // modalities in each input set
  size_t n_tbz, n_obs;                  // T*B*Z, M*T*B*Z
  // workspace offsets (bytes)
  size_t off_acc, off_count, off_obs_mean, off_obs_std, off_obs_mask, off_dobs_mean, off_dobs_std;
  size_t off_a[6], off_b[6], off_c[6];  // infer_m, infer_s, prior_m, prior_s, samples / d_prior_m, d_samples / d_prior_s
  size_t zero_begin, zero_end;          // region cleared at step start
  size_t off_seg, seg_bytes;            // scratch of the time-segmented pass-B backward
  size_t total;
};

// sets of one DGTS step in evaluation order (models/dgts.py:119-129)
void plan_step(const bfvi_model* m, const bfvi_step_args* a, bool with_grad, StepPlan* pl) {
  int S = 0;
  const int M = m->n_mods;
  if (M > 1) pl->set_bits[S++] = (M >= 32) ? 0xffffffffu : ((1u << M) - 1u);
  if (a->uni_loss)
    for (int i = 0; i < M; ++i) pl->set_bits[S++] = 1u << i;
  pl->S = S;
  const size_t tb = (size_t)a->T * a->B;
  pl->n_tbz = tb * m->z_dim;
  pl->n_obs = pl->n_tbz * M;
  size_t cur = 0;
  auto carve = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
  const size_t fS = sizeof(float) * pl->n_tbz * (size_t)(S > 0 ? S : 1);
  // --- zeroed region: accumulators and gradient scratch ---
  pl->zero_begin = cur;
  pl->off_acc = carve(sizeof(double));
  pl->off_count = carve(sizeof(float));
  pl->off_dobs_mean = carve(with_grad ? sizeof(float) * pl->n_obs : 0);
  pl->off_dobs_std = carve(with_grad ? sizeof(float) * pl->n_obs : 0);
  pl->off_a[5] = carve(with_grad ? fS : 0);   // d_samples A
  pl->off_b[4] = carve(with_grad ? fS : 0);   // d_prior_mean B
  pl->off_b[5] = carve(with_grad ? fS : 0);   // d_prior_std B
  pl->off_c[5] = carve(with_grad ? fS : 0);   // d_samples C
  pl->zero_end = cur;
  pl->off_obs_mean = carve(sizeof(float) * pl->n_obs);
  pl->off_obs_std = carve(sizeof(float) * pl->n_obs);
  pl->off_obs_mask = carve(tb * M);
  for (int i = 0; i < 5; ++i) pl->off_a[i] = carve(fS);
  for (int i = 0; i < 4; ++i) pl->off_b[i] = carve(fS);
  for (int i = 0; i < 5; ++i) pl->off_c[i] = carve(fS);
  pl->seg_bytes = seg_scratch_bytes(S > 0 ? S : 1, a->B, m->z_dim);
  pl->off_seg = carve(pl->seg_bytes);
  pl->total = cur;
}

// One launch of the grouped persistent tile kernel over `n` independent problems (no problem may read
// what another one of the group writes, and no two may accumulate into the same C without split-K atomics).
template <int BN, bool SPLIT>
int launch_gemm_group(const bfvi::tc::GemmParams* gps, int n, cudaStream_t st) {
#ifdef BFVI_EMU
  (void)st;
  for (int i = 0; i < n; ++i) bfvi::tc::gemm_reference_emu(gps[i]);
#else
  static const bool v1 = [] { const char* e = getenv("BFVI_GEMM_V1"); return e && atoi(e) != 0; }();
  if (v1) {                          // round-1 kernel kept for A/B timing (tools/time_gemm.py)
    for (int i = 0; i < n; ++i) {
      const bfvi::tc::GemmParams& gp = gps[i];
      const unsigned gz = gp.k_split > 0 ? (unsigned)((gp.K + gp.k_split - 1) / gp.k_split) : 1u;
      const dim3 grid((unsigned)((gp.M + bfvi::tc::kBM - 1) / bfvi::tc::kBM), (unsigned)((gp.N + BN - 1) / BN), gz);
      auto k = bfvi::tc::gemm_tf32_kernel<BN, SPLIT>;
      const size_t smem = bfvi::tc::gemm_smem_bytes<BN, SPLIT>();
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<grid, dim3(bfvi::tc::kThreads), smem, st>>>(gp);
    }
  } else {
    static const int cap = [] { const char* e = getenv("BFVI_GEMM_STAGES"); return e ? atoi(e) : bfvi::tc::kMaxStages; }();
    static const int ctas_per_sm = [] { const char* e = getenv("BFVI_GEMM_CTAS"); return e ? atoi(e) : 1; }();
    bfvi::tc::GemmGroup grp;
    memset(&grp, 0, sizeof(grp));
    bool vec = true;
    int total = 0;
    for (int i = 0; i < n; ++i) {
      const bfvi::tc::GemmParams& gp = gps[i];
      grp.g[i] = gp;
      const int64_t k_len = gp.k_split > 0 ? gp.k_split : gp.K;
      grp.chunks[i] = (int)((k_len + bfvi::tc::kBK - 1) / bfvi::tc::kBK);
      grp.tiles_m[i] = (int)((gp.M + bfvi::tc::kBM - 1) / bfvi::tc::kBM);
      grp.tiles_n[i] = (gp.N + BN - 1) / BN;
      const int tz = gp.k_split > 0 ? (int)((gp.K + gp.k_split - 1) / gp.k_split) : 1;
      total += grp.tiles_m[i] * grp.tiles_n[i] * tz;
      grp.tile_end[i] = total;
      vec = vec && gp.lda % 4 == 0 && gp.ldw % 4 == 0 && (((uintptr_t)gp.A | (uintptr_t)gp.W) & 15) == 0;
    }
    grp.n = n; grp.total = total;
    // the operand ring runs across tile boundaries, so it is as deep as fits (4 stages / 200 kB)
    const size_t budget = (ctas_per_sm > 1 ? 96u : 200u) * 1024u;
    int stages = (int)(budget / bfvi::tc::gemm_v2_stage_bytes<BN, SPLIT>());
    if (stages > bfvi::tc::kMaxStages) stages = bfvi::tc::kMaxStages;
    if (stages > cap) stages = cap;
    if (stages < 2) stages = 2;
    const int slots = (num_sms() > 0 ? num_sms() : 1) * (ctas_per_sm > 1 ? 2 : 1);
    static const bool a_in_tmem = [] { const char* e = getenv("BFVI_GEMM_TS"); return !e || atoi(e) != 0; }();
    static const int ts_cap = [] { const char* e = getenv("BFVI_GEMM_TS_STAGES"); return e ? atoi(e) : 8; }();
    static const int lag_env = [] { const char* e = getenv("BFVI_GEMM_LAG"); return e ? atoi(e) : 0; }();   // 0: kernel default
    if constexpr (BN <= 128 && SPLIT) {
      if (a_in_tmem && vec) {        // A operand from tensor memory (aligned 3xTF32 problems)
        // aligned epilogue rows leave through cp.async.bulk (one 128-byte line per row from a padded patch; the ring gives
        // up the shared memory the larger patches need: 3 stages at 128-wide tiles, measured equal to 4);
        // BFVI_GEMM_BULK=0 restores the per-thread 16-byte global stores
        static const bool bulk = [] { const char* e = getenv("BFVI_GEMM_BULK"); return !e || atoi(e) != 0; }();
        grp.bulk_store = bulk ? 1 : 0;
        const size_t ts_budget = bulk ? (size_t)232448 - 1024 - 8 * 32 * 36 * sizeof(float) : budget;
        int ts_stages = (int)(ts_budget / bfvi::tc::gemm_ts_stage_bytes<BN, SPLIT>());
        if (ts_stages > bfvi::tc::ts_max_stages(BN)) ts_stages = bfvi::tc::ts_max_stages(BN);
        if (ts_stages > ts_cap) ts_stages = ts_cap;
        if (ts_stages < 2) ts_stages = 2;
        const size_t smem_ts = bfvi::tc::gemm_ts_smem_bytes<BN, SPLIT>(ts_stages, bulk);
        auto kt = bfvi::tc::gemm_tf32_ts_kernel<BN, SPLIT, true>;
        ensure_dyn_smem(kt, smem_ts);
        kt<<<dim3((unsigned)(total < slots ? total : slots)), dim3(bfvi::tc::kThreadsPhost), smem_ts, st>>>(grp, ts_stages, lag_env);
        BFVI_CHECK_CUDA();
        return BFVI_OK;
      }
    }
    const size_t smem = bfvi::tc::gemm_p_smem_bytes<BN, SPLIT>(stages);
    auto k = vec ? bfvi::tc::gemm_tf32_p_kernel<BN, SPLIT, true> : bfvi::tc::gemm_tf32_p_kernel<BN, SPLIT, false>;
    ensure_dyn_smem(k, smem);
    k<<<dim3((unsigned)(total < slots ? total : slots)), dim3(bfvi::tc::kThreadsPhost), smem, st>>>(grp, stages, lag_env);
  }
#endif
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

// Tile width of a launch: the narrowest of 32 / 64 / 128 that covers the widest problem.  The kernel is bound
// by the SM's shared-memory pipe (cp.async writes, the rounding pass, three UMMA operand reads per k-step
// all share 128 B/clk), so a wider tile — the A chunk is copied and rounded once per 128 instead of per
// 64 columns — is worth more than the extra ring stage a narrow one affords: C3-dims step 205 ms (cap 64),
// 192 ms (128), 196 ms (256).  BFVI_GEMM_BN overrides the cap.
template <bool SPLIT>
int dispatch_gemm_bn(const bfvi::tc::GemmParams* gps, int n, cudaStream_t st) {
  static const int cap = [] { const char* e = getenv("BFVI_GEMM_BN"); return e ? atoi(e) : 128; }();
  int n_max = 0;
  for (int i = 0; i < n; ++i) n_max = gps[i].N > n_max ? gps[i].N : n_max;
  if (n_max <= 32 || cap <= 32) return launch_gemm_group<32, SPLIT>(gps, n, st);
  if (n_max <= 64 || cap <= 64) return launch_gemm_group<64, SPLIT>(gps, n, st);
  if (n_max <= 128 || cap <= 128) return launch_gemm_group<128, SPLIT>(gps, n, st);
  return launch_gemm_group<256, SPLIT>(gps, n, st);
}

int gemm_group_tc(const bfvi::tc::GemmParams* gps, int n, int prec, cudaStream_t st) {
  if (n <= 0) return BFVI_OK;
  return prec == bfvi::tc::PREC_TF32 ? dispatch_gemm_bn<false>(gps, n, st) : dispatch_gemm_bn<true>(gps, n, st);
}
int gemm_tc(const bfvi::tc::GemmParams& gp, int prec, cudaStream_t st) { return gemm_group_tc(&gp, 1, prec, st); }

// A weight gradient dW (n_out, n_in) = dY^T X with few outputs and more inputs (the hidden -> head layers:
// 64 x 512, decoder heads 16 x 512) would pad the 128 MMA rows with zeros; it runs transposed instead
// (GemmParams::trans_out), the same shape as the fast 512 x 64 case.  BFVI_WGRAD_SWAP=0 disables.
bool wgrad_swapped(int n_out, int n_in) {
  static const bool on = [] {
    const char* e = getenv("BFVI_WGRAD_SWAP");
    const char* v1 = getenv("BFVI_GEMM_V1");            // the round-1 kernel has no transposed output
    return (!e || atoi(e) != 0) && !(v1 && atoi(v1) != 0);
  }();
  return on && n_out < n_in && n_out <= 64;
}

// Split of the (long) contraction over the rows of a weight-gradient GEMM: the output has only a
// few 128 x BN tiles, so K is cut into enough slices to put ~2 CTAs on every SM (slices are a
// multiple of the 32-float stage and at least 256 long).
int64_t wgrad_k_split(int64_t rows, int n_out, int n_in) {
  const int bn = n_in <= 32 ? 32 : n_in <= 64 ? 64 : n_in <= 128 ? 128 : 256;
  const int64_t tiles = ((n_out + 127) / 128) * (int64_t)((n_in + bn - 1) / bn);
  const int sms = num_sms() > 0 ? num_sms() : 1;
  int64_t slices = (2 * sms + tiles - 1) / tiles;
  if (slices < 1) slices = 1;
  int64_t len = (rows + slices - 1) / slices;
  len = (len + 31) / 32 * 32;
  if (len < 256) len = 256;
  return len >= rows ? 0 : len;
}

// y = act(x W^T + b) on tcgen05; prec: PREC_TF32X3 (error-compensated) or PREC_TF32
int linear_tc(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y, int64_t ldy,
              int64_t n_rows, int n_in, int n_out, int act, int prec, cudaStream_t st) {
  bfvi::tc::GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.A = x; gp.lda = ldx; gp.W = w; gp.ldw = ldw; gp.bias = bias; gp.C = y; gp.ldc = ldy;
  gp.M = n_rows; gp.N = n_out; gp.K = n_in; gp.act = act;
  return gemm_tc(gp, prec, st);
}

// a few independent y = act(x W^T + b) layers as ONE grouped launch
struct LinearGroup {
  bfvi::tc::GemmParams g[bfvi::tc::kMaxGroup];
  int n = 0;
  void add(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y, int64_t ldy,
           int64_t n_rows, int n_in, int n_out, int act) {
    bfvi::tc::GemmParams& gp = g[n++];
    memset(&gp, 0, sizeof(gp));
    gp.A = x; gp.lda = ldx; gp.W = w; gp.ldw = ldw; gp.bias = bias; gp.C = y; gp.ldc = ldy;
    gp.M = n_rows; gp.N = n_out; gp.K = n_in; gp.act = act;
  }
  int run(int prec, cudaStream_t st) {
    const int rc = gemm_group_tc(g, n, prec, st);
    n = 0;
    return rc;
  }
};


// ===========================================================================================
// fused on-chip transition kernels (bfvi_fused.cuh): host side
// ===========================================================================================
// served shapes: Z = 64, H a multiple of 128 (C3: 64 / 512); anything else takes the launch-sequence path
bool fused_supported(int Z, int H) {
#ifdef BFVI_EMU
  (void)Z; (void)H;
  return false;
#else
  static const bool off = [] { const char* e = getenv("BFVI_FUSED"); return e && atoi(e) == 0; }();
  return !off && Z == bfvi::fused::kZ && H >= 128 && H % 128 == 0 && H <= 2048;
#endif
}
struct FusedBufs {
  unsigned char* pack_fwd[2]; unsigned char* pack_bwd[2]; float* bias[2];      // per direction, shared by all passes
  void* h16; void* dh16; void* z16; void* dg16; void* dnl16; uint32_t* bits;   // per-row scratch (rows padded to 128)
  unsigned* gstat;               // 8 words: [0..2] maxima of the head gradients (float bits), [4] the scale gtf_bwd_kernel chose
};
int64_t fused_rows_pad(int64_t rows) { return (rows + 127) / 128 * 128; }
// per-row scratch bytes: hidden activations + their gradients as FP16 tiles, three Z-wide FP16 tiles, the ReLU bits
struct FusedRowScratch { size_t h16, dh16, z16, dg16, dnl16, bits, gstat, total; };
FusedRowScratch fused_row_scratch(int H, int64_t rows) {
  const size_t rp = (size_t)fused_rows_pad(rows);
  FusedRowScratch r;
  size_t cur = 0;
  auto carve = [&](size_t bytes) { size_t o = cur; cur = (cur + bytes + 1023) / 1024 * 1024; return o; };
  r.h16 = carve(rp * 2 * H * 2); r.dh16 = carve(rp * 2 * H * 2);
  r.z16 = carve(rp * 64 * 2); r.dg16 = carve(rp * 64 * 2); r.dnl16 = carve(rp * 64 * 2);
  r.bits = carve(rp * (2 * H / 64) * 2 * 4);
  r.gstat = carve(256);
  r.total = cur;
  return r;
}
size_t fused_pack_total(int H) {          // both directions: forward pack, backward pack, bias table
  return 2 * (2 * bfvi::fused::pack_bytes(H) + (bfvi::fused::bias_floats(H) * 4 + 1023) / 1024 * 1024);
}
void fused_carve_packs(char* base, int H, FusedBufs* fb) {
  const size_t pb = bfvi::fused::pack_bytes(H), bb = (bfvi::fused::bias_floats(H) * 4 + 1023) / 1024 * 1024;
  for (int d = 0; d < 2; ++d) {
    char* o = base + (size_t)d * (2 * pb + bb);
    fb->pack_fwd[d] = (unsigned char*)o; fb->pack_bwd[d] = (unsigned char*)(o + pb); fb->bias[d] = (float*)(o + 2 * pb);
  }
}
void fused_carve_rows(char* base, int H, int64_t rows, FusedBufs* fb) {
  const FusedRowScratch r = fused_row_scratch(H, rows);
  fb->h16 = base + r.h16; fb->dh16 = base + r.dh16; fb->z16 = base + r.z16; fb->dg16 = base + r.dg16;
  fb->dnl16 = base + r.dnl16; fb->bits = (uint32_t*)(base + r.bits);
  fb->gstat = (unsigned*)(base + r.gstat);
}
#ifndef BFVI_EMU
int fused_smem_limit() { return 232448; }
// development: ablation mask of the fused kernels (timing probes only; read per call so a probe can switch it)
int fused_abl() { const char* e = getenv("BFVI_FUSED_ABL"); return e ? atoi(e) : 0; }
int fused_stages(size_t extra_bytes) {
  int st = (int)((fused_smem_limit() - 1024 - extra_bytes) / bfvi::fused::kBlockBytes);
  if (st > 6) st = 6;
  static const int cap = [] { const char* e = getenv("BFVI_FUSED_STAGES"); return e ? atoi(e) : 99; }();
  if (st > cap) st = cap;
  return st;
}
int fused_pack(const bfvi_gtf_layout& g, const float* params, int dir, int H, const FusedBufs& fb, cudaStream_t st) {
  bfvi::fused::PackParams pp;
  pp.w_gate0 = params + g.gate0_w; pp.b_gate0 = params + g.gate0_b; pp.w_gate2 = params + g.gate2_w; pp.b_gate2 = params + g.gate2_b;
  pp.w_lin = params + g.lin_w; pp.b_lin = params + g.lin_b; pp.w_non0 = params + g.nonlin0_w; pp.b_non0 = params + g.nonlin0_b;
  pp.w_non2 = params + g.nonlin2_w; pp.b_non2 = params + g.nonlin2_b; pp.w_std = params + g.std_w; pp.b_std = params + g.std_b;
  pp.fwd = fb.pack_fwd[dir]; pp.bwd = fb.pack_bwd[dir]; pp.bias = fb.bias[dir]; pp.H = H;
  auto ks = bfvi::fused::gtf_scale_kernel;
  ks<<<dim3(8), dim3(256), 0, st>>>(pp);
  auto k = bfvi::fused::pack_gtf_kernel;
  k<<<dim3((unsigned)bfvi::fused::pack_blocks(H)), dim3(256), 0, st>>>(pp);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}
constexpr size_t kFusedPatches = (size_t)4 * bfvi::fused::kPatchBytes;
// heads of one transition for `rows` latent rows; keep = also write the operand tiles / ReLU bits the backward needs
int fused_fwd(const FusedBufs& fb, int dir, int H, const float* z, int64_t rows, float* g, float* nl, float* lin, float* as,
              bool keep, cudaStream_t st) {
  bfvi::fused::FwdParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.pack = fb.pack_fwd[dir]; fp.bias = fb.bias[dir]; fp.z = z; fp.g = g; fp.nl = nl; fp.lin = lin; fp.as = as;
  fp.h16 = (__half*)fb.h16; fp.relu_bits = fb.bits; fp.z16 = (__half*)fb.z16;
  fp.R = rows; fp.H = H; fp.abl = fused_abl();
  const size_t extra = bfvi::fused::bias_floats(H) * 4 + kFusedPatches;
  fp.n_stages = fused_stages(extra);
  if (fp.n_stages < 4) return fail(BFVI_ERR_UNSUPPORTED, "fused transition kernel: h_dim %d too wide", H);
  const size_t smem = (size_t)fp.n_stages * bfvi::fused::kBlockBytes + extra + 1024;
  const int64_t tiles = (rows + 127) / 128;
  const int sms = num_sms() > 0 ? num_sms() : 1;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  static const bool dbg = [] { const char* e = getenv("BFVI_FUSED_DBG"); return e && atoi(e) != 0; }();
  long long* dbg_dev = nullptr;
  if (dbg) { cudaMalloc(&dbg_dev, 32 * sizeof(long long)); cudaMemsetAsync(dbg_dev, 0, 32 * sizeof(long long), st); fp.dbg = dbg_dev; }
  auto k = keep ? bfvi::fused::gtf_fwd_kernel<true> : bfvi::fused::gtf_fwd_kernel<false>;
  ensure_dyn_smem(k, smem);
  k<<<dim3(grid), dim3(bfvi::fused::kThreads), smem, st>>>(fp);
  note_dispatch("gtf_fwd_fused%s f16x3 stages=%d", keep ? "<keep>" : "", fp.n_stages);
  if (dbg) {                           // development: where do the cycles of CTA 0 go?
    cudaStreamSynchronize(st);
    long long h[32];
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    const double tpc = (double)((tiles + grid - 1) / grid), pairs = tpc * (H / 64);
    fprintf(stderr, "[fused fwd dbg] rows %lld tiles/CTA %.0f | row warp 0 per PAIR: wait_d %.0f ld %.0f math %.0f st+arrive %.0f | tail+heads per tile %.0f (wait units_done %.0f, std-head operand %.0f, nl rows %.0f, wait heads_full %.0f, next z + heads to registers %.0f, scale + 3 row stores %.0f) | total %.0f per tile\n"
                    "                issuer per PAIR: wait_a %.0f wait_blk %.0f issue %.0f\n",
            (long long)rows, tpc, h[0] / pairs, h[1] / pairs, h[2] / pairs, h[3] / pairs, h[4] / tpc, h[6] / tpc, h[8] / tpc, h[9] / tpc, h[7] / tpc, h[10] / tpc, h[11] / tpc, h[5] / tpc,
            h[16] / pairs, h[17] / pairs, h[18] / pairs);
  }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}
// input gradient dz of one transition (the KEEP forward of the same rows must have run)
int fused_bwd(const FusedBufs& fb, int dir, int H, const float* d_g, const float* d_nl, const float* d_lin, int64_t rows,
              float* dz, cudaStream_t st) {
  static const bool skip_bwd = getenv("BFVI_DBG_SKIP_BWD") != nullptr;      // development: bisect a hang
  if (skip_bwd) return BFVI_OK;
  bfvi::fused::BwdParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.pack = fb.pack_bwd[dir]; bp.d_g = d_g; bp.d_nl = d_nl; bp.d_lin = d_lin; bp.relu_bits = fb.bits; bp.dz = dz;
  bp.dh16 = (__half*)fb.dh16; bp.dg16 = (__half*)fb.dg16; bp.dnl16 = (__half*)fb.dnl16;
  bp.R = rows; bp.H = H; bp.abl = fused_abl();
  // power-of-two gradient scale from the maxima the producer of the head gradients keeps (BFVI_FUSED_GSCALE=0: none)
  static const bool gscale_on = [] { const char* e = getenv("BFVI_FUSED_GSCALE"); return !e || atoi(e) != 0; }();
  bp.gmax = gscale_on ? fb.gstat : nullptr;
  bp.l1 = fb.bias[dir] + 2 * H + 4 * bfvi::fused::kZ + 6;
  bp.gscale = reinterpret_cast<float*>(fb.gstat + 4);
  bp.n_stages = fused_stages(kFusedPatches);
  if (bp.n_stages < 4) return fail(BFVI_ERR_UNSUPPORTED, "fused transition kernel: no room for the weight ring");
  const size_t smem = (size_t)bp.n_stages * bfvi::fused::kBlockBytes + kFusedPatches + 1024;
  const int64_t tiles = (rows + 127) / 128;
  const int sms = num_sms() > 0 ? num_sms() : 1;
  auto k = bfvi::fused::gtf_bwd_kernel;
  ensure_dyn_smem(k, smem);
  k<<<dim3((unsigned)(tiles < sms ? tiles : sms)), dim3(bfvi::fused::kThreads), smem, st>>>(bp);
  note_dispatch("gtf_bwd_fused tf32 stages=%d", bp.n_stages);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}
// weight gradients of the four H-wide layers (+ the two hidden bias gradients) from the FP16 operand tiles of
// fused_fwd(keep) / fused_bwd
int fused_wgrad(const FusedBufs& fb, int H, int64_t rows, float* dw_gate0, float* dw_non0, float* dw_gate2, float* dw_non2,
                float* gb_gate0, float* gb_non0, cudaStream_t st) {
  static const bool skip_wgrad = getenv("BFVI_DBG_SKIP_WGRAD") != nullptr;  // development: bisect a hang
  if (skip_wgrad) return BFVI_OK;
  bfvi::fused::Wgrad16Params wp;
  memset(&wp, 0, sizeof(wp));
  const int U = H / 64;
  auto prob = [&](int i, const void* X, int atom0, const void* Y, float* out, int transposed, float* bias) {
    wp.pr[i].X = (const __half*)X; wp.pr[i].n_atoms_x = 2 * U; wp.pr[i].atom0 = atom0; wp.pr[i].Y = (const __half*)Y;
    wp.pr[i].out = out; wp.pr[i].transposed = transposed; wp.pr[i].bias = bias;
  };
  prob(0, fb.dh16, 0, fb.z16, dw_gate0, 0, gb_gate0);       // dW_gate0 (H, Z) = dh1^T z, b_gate0 = column sums of dh1
  prob(1, fb.dh16, U, fb.z16, dw_non0, 0, gb_non0);         // dW_nonlin0 (H, Z) = dh3^T z
  prob(2, fb.h16, 0, fb.dg16, dw_gate2, 1, nullptr);        // dW_gate2 (Z, H) = d_g^T h1, computed as h1^T d_g
  prob(3, fb.h16, U, fb.dnl16, dw_non2, 1, nullptr);        // dW_nonlin2 (Z, H) = d_nl^T h3
  wp.n_problems = 4; wp.H = H; wp.abl = fused_abl();
  wp.gscale = reinterpret_cast<const float*>(fb.gstat + 4);      // written by the fused_bwd launch before this one
  wp.n_groups = fused_rows_pad(rows) / 64;
  const int sms = num_sms() > 0 ? num_sms() : 1;
  const int tiles = 4 * (H / 128);
  int64_t slices = (2 * sms + tiles - 1) / tiles;
  if (slices > wp.n_groups) slices = wp.n_groups;
  if (slices < 1) slices = 1;
  wp.groups_per_slice = (int)((wp.n_groups + slices - 1) / slices);
  wp.n_slices = (int)((wp.n_groups + wp.groups_per_slice - 1) / wp.groups_per_slice);
  wp.n_stages = 6;
  if (const char* e = getenv("BFVI_WGRAD_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= bfvi::fused::kMaxStages) wp.n_stages = v; }
  const size_t smem = (size_t)wp.n_stages * bfvi::fused::kWgStageBytes + bfvi::fused::kAtomBytes + 1024;
  const int items = tiles * wp.n_slices;
  auto k = bfvi::fused::wgrad16_kernel;
  ensure_dyn_smem(k, smem);
  k<<<dim3((unsigned)(items < sms ? items : sms)), dim3(bfvi::fused::kWgThreads), smem, st>>>(wp);
  note_dispatch("wgrad16 f16-mn");
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}
#endif  // !BFVI_EMU

// ===========================================================================================
// large-dim family: the whole MultiDMM.step + backward as a stream-ordered launch sequence of
// tcgen05 GEMMs (every Linear layer: forward, input gradient, weight gradient) and fused
// elementwise kernels (bfvi_generic.cuh).  Same passes, chain sets and workspace roles as the
// small-dim step_impl below; rows of a GEMM = chains x particles of ONE time step.
// ===========================================================================================
constexpr size_t kMlpChunkRows = 262144;          // rows per chunk of the encoder / decoder passes (BFVI_MLP_CHUNK overrides)
size_t mlp_chunk_rows() {
  const char* e = getenv("BFVI_MLP_CHUNK");
  const long v = e ? atol(e) : 0;
  return v >= 32 ? (size_t)v : kMlpChunkRows;
}
struct LargePlan {
  int S, k_b;
  unsigned set_bits[BFVI_MAX_SETS];
  size_t tb, tbz, C, R, d_max;
  size_t rc;                                       // rows per chunk of the encoder / decoder scratch (<= tb)
  size_t zero_begin, zero_end, total;
  size_t acc, count, dobs_mean, dobs_std, a_dsamp, c_dsamp, b_dpm, b_dps;        // zeroed
  size_t paramsT, x0[BFVI_MAX_MODS], x0T[BFVI_MAX_MODS], mask, henc, hencT, obs_mean, obs_stdpre, obs_std;
  size_t pa[6], pb[4], pc[6];                       // infer m/s, prior m/s, samples, samplesT
  size_t zrows, zrowsT, h1, h1T, h3, h3T, dh1, dh1T, dh3, dh3T;
  size_t g, nl, nlT, lin, as, d_as, d_asT, d_g, d_gT, d_lin, d_linT, d_nl, d_nlT, dz, dz2;
  size_t c_mu, c_sd, d_pm, d_v, zvec;
  size_t hdec, hdecT, dhd, dhdT, dmean, dstd, dmeanT, dstdT;
  size_t side_grads;                                 // zeroed: parameter gradients of the side stream (pass A)
  size_t fused_packs, fused_rows;                    // fused transition kernels: weight packs (shared), per-row scratch
  int fused;
};

// `fonly` != null plans the workspace of a stand-alone z_filter call (bfvi_filter_fwd / _bwd of the
// large-dim family): only the per-step row scratch, the chain scratch and the transposed weights.
int plan_large(const bfvi_model* m, const bfvi_step_args* a, const bfvi_filter_args* fonly, LargePlan* pl) {
  const int M = m->n_mods, Z = m->z_dim, H = m->h_dim;
  const int T = fonly ? fonly->T : a->T, B = fonly ? fonly->B : a->B;
  int S = 0;
  if (fonly) {
    S = fonly->S;
  } else {
    if (M > 1) pl->set_bits[S++] = (M >= 32) ? 0xffffffffu : ((1u << M) - 1u);
    if (a->uni_loss)
      for (int i = 0; i < M; ++i) pl->set_bits[S++] = 1u << i;
  }
  pl->S = S;
  pl->k_b = fonly ? fonly->n_particles : a->train_particles;
  pl->tb = fonly ? 0 : (size_t)T * B;
  pl->tbz = pl->tb * Z;
  pl->C = (size_t)(S > 0 ? S : 1) * B;
  size_t R = pl->C * (size_t)pl->k_b;
  if (!fonly && (size_t)a->match_particles > R) R = a->match_particles;
  pl->R = R;
  pl->d_max = Z;                     // head scratch serves decoders (D_m wide) and encoders (Z wide)
  for (int i = 0; i < M; ++i) if ((size_t)m->dims[i] > pl->d_max) pl->d_max = m->dims[i];
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  size_t cur = 0;
  auto carve = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
  const size_t f = sizeof(float), fS = f * pl->tbz * (S > 0 ? S : 1);
  const size_t Mx = fonly ? 0 : M;                    // no encoder / decoder scratch for a lone filter
  pl->zero_begin = cur;
  pl->acc = carve(sizeof(double)); pl->count = carve(f);
  pl->dobs_mean = carve(f * pl->tbz * M); pl->dobs_std = carve(f * pl->tbz * M);
  pl->a_dsamp = carve(fS); pl->c_dsamp = carve(fS); pl->b_dpm = carve(fS); pl->b_dps = carve(fS);
  pl->zero_end = cur;
  pl->paramsT = carve(f * lay.total);
  (void)Mx;
  for (int i = 0; i < M; ++i) { pl->x0[i] = carve(f * pl->tb * m->dims[i]); pl->x0T[i] = carve(f * pl->tb * m->dims[i]); }
  pl->mask = carve(pl->tb * M);
  // encoder / decoder hidden activations live in a ROW-CHUNK scratch (hdec, hdecT, dhd, dhdT: rc x H each): the encoder
  // backward recomputes its hidden layer per chunk instead of keeping (T*B, H) x M activations and their transposed copies
  // (32 KB per sequence-timestep at the C3 shape, 29 % of the step's workspace), and the decoders walk T*B in chunks
  pl->rc = pl->tb < mlp_chunk_rows() ? pl->tb : mlp_chunk_rows();
  pl->henc = pl->hencT = 0;
  pl->obs_mean = carve(f * pl->tbz * M); pl->obs_stdpre = carve(f * pl->tbz * M); pl->obs_std = carve(f * pl->tbz * M);
  for (int i = 0; i < 6; ++i) pl->pa[i] = carve(fS);
  for (int i = 0; i < 4; ++i) pl->pb[i] = carve(fS);
  for (int i = 0; i < 6; ++i) pl->pc[i] = carve(fS);
  pl->fused = (a != nullptr && a->precision == BFVI_PREC_FUSED && fused_supported(Z, H)) ? 1 : 0;
  // hidden activations / gradients of a time step's transition (+ transposed copies): launch-sequence path only
  // (the fused kernels keep them on-chip and write FP16 operand tiles into `fused_rows` instead)
  const size_t rz = f * R * Z, rh = pl->fused ? 256 : f * R * H;
  pl->zrows = carve(rz); pl->zrowsT = carve(rz);
  pl->h1 = carve(rh); pl->h1T = carve(rh); pl->h3 = carve(rh); pl->h3T = carve(rh);
  pl->dh1 = carve(rh); pl->dh1T = carve(rh); pl->dh3 = carve(rh); pl->dh3T = carve(rh);
  size_t* zbufs[] = {&pl->g, &pl->nl, &pl->nlT, &pl->lin, &pl->as, &pl->d_as, &pl->d_asT, &pl->d_g, &pl->d_gT,
                     &pl->d_lin, &pl->d_linT, &pl->d_nl, &pl->d_nlT, &pl->dz, &pl->dz2};
  for (size_t* z : zbufs) *z = carve(rz);
  pl->c_mu = carve(f * pl->C * Z); pl->c_sd = carve(f * pl->C * Z);
  pl->d_pm = carve(f * pl->C * Z); pl->d_v = carve(f * pl->C * Z);
  pl->zvec = carve(f * Z * 4);
  pl->hdec = carve(f * pl->rc * H); pl->hdecT = carve(f * pl->rc * H);
  pl->dhd = carve(f * pl->rc * H); pl->dhdT = carve(f * pl->rc * H);
  pl->dmean = carve(f * pl->rc * pl->d_max); pl->dstd = carve(f * pl->rc * pl->d_max);
  pl->dmeanT = carve(f * pl->rc * pl->d_max); pl->dstdT = carve(f * pl->rc * pl->d_max);
  pl->fused_packs = pl->fused_rows = 0;
  if (pl->fused) {
    cur = align_up(cur, 1024);
    pl->fused_packs = carve(fused_pack_total(H));
    cur = align_up(cur, 1024);
    pl->fused_rows = carve(fused_row_scratch(H, (int64_t)R).total);
  }
  pl->total = cur;
  return BFVI_OK;
}

// Second scratch set for the filter-only pass ("pass A": f_mode, one particle), which the step runs on a
// side stream beside the particle pass and the smoother: its own row scratch (rows = chains), decoder
// scratch, observation-gradient and parameter-gradient buffers, appended to the main plan's workspace.
void plan_large_side(const bfvi_model* m, const bfvi_step_args* a, const LargePlan& pl, LargePlan* ps) {
  *ps = pl;
  const int M = m->n_mods, Z = m->z_dim, H = m->h_dim;
  (void)a;
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  size_t cur = pl.total;
  auto carve = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
  const size_t f = sizeof(float);
  ps->zero_begin = cur;
  ps->dobs_mean = carve(f * pl.tbz * M); ps->dobs_std = carve(f * pl.tbz * M);
  ps->side_grads = carve(f * lay.total);
  ps->zero_end = cur;
  const size_t R = pl.C;                               // one particle per chain
  ps->R = R;
  const size_t rz = f * R * Z, rh = pl.fused ? 256 : f * R * H;
  ps->zrows = carve(rz); ps->zrowsT = carve(rz);
  ps->h1 = carve(rh); ps->h1T = carve(rh); ps->h3 = carve(rh); ps->h3T = carve(rh);
  ps->dh1 = carve(rh); ps->dh1T = carve(rh); ps->dh3 = carve(rh); ps->dh3T = carve(rh);
  size_t* zbufs[] = {&ps->g, &ps->nl, &ps->nlT, &ps->lin, &ps->as, &ps->d_as, &ps->d_asT, &ps->d_g, &ps->d_gT,
                     &ps->d_lin, &ps->d_linT, &ps->d_nl, &ps->d_nlT, &ps->dz, &ps->dz2};
  for (size_t* z : zbufs) *z = carve(rz);
  ps->c_mu = carve(f * pl.C * Z); ps->c_sd = carve(f * pl.C * Z);
  ps->d_pm = carve(f * pl.C * Z); ps->d_v = carve(f * pl.C * Z);
  ps->hdec = carve(f * pl.rc * H); ps->hdecT = carve(f * pl.rc * H);
  ps->dhd = carve(f * pl.rc * H); ps->dhdT = carve(f * pl.rc * H);
  ps->dmean = carve(f * pl.rc * pl.d_max); ps->dstd = carve(f * pl.rc * pl.d_max);
  ps->dmeanT = carve(f * pl.rc * pl.d_max); ps->dstdT = carve(f * pl.rc * pl.d_max);
  if (pl.fused) {                                      // packs are shared with the main plan; own row scratch
    cur = align_up(cur, 1024);
    ps->fused_rows = carve(fused_row_scratch(H, (int64_t)R).total);
  }
  ps->total = cur;
}

// y[i] += x[i]  (join of the side stream's gradient buffers)
__global__ void __launch_bounds__(256) add_into_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += x[i];
}

// fills mu[z] = z0_mean, sd[z] = exp(z0_log_std) + min_std (the "infer" of the prior-matching chain)
__global__ void prior_fill_kernel(const float* z0_mean, const float* z0_log_std, float min_std, int Z, float* mu,
                                  float* sd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Z) { mu[i] = z0_mean[i]; sd[i] = expf(z0_log_std[i]) + min_std; }
}
// z_k = gm + eps_k gs: gradient of the propagated particles back to the global prior
__global__ void match_tail_kernel(const float* c_mu, const float* c_sd, const float* z0_log_std, int Z,
                                  float* g_z0_mean, float* g_z0_log_std) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Z) { atomicAdd(g_z0_mean + i, c_mu[i]); atomicAdd(g_z0_log_std + i, c_sd[i] * expf(z0_log_std[i])); }
}

// step_kernel, or its four-components-per-thread form when the rows are 16-byte multiples
static bool step4_ok(const bfvi::gen::StepParams& sp) {
  return sp.Z % 4 == 0 && (((uintptr_t)sp.g | (uintptr_t)sp.nl | (uintptr_t)sp.lin | (uintptr_t)sp.as | (uintptr_t)sp.zrows) & 15) == 0;
}

// One batch tile of a step that walks its batch in tiles (step_large_tiled): the loss accumulator and the mask count
// live outside the tile's workspace, the gradient buffer is cleared by the first tile only, the prior-matching term
// (linear in the GLOBAL mask count, models/dmm.py:541-545) is added by the first tile, the loss is finalised by the last.
// block width of head_kernel (thread = output column, 32 rows per block): the narrowest warp multiple that covers D, so a
// 16-wide decoder head does not idle 112 of 128 threads
inline int head_block(int D) { return D <= 32 ? 32 : (D <= 64 ? 64 : 128); }
struct TileCtx { bool first, last; double* acc; const float* count; bool match; int lane; };

// fonly != null: run only z_filter forward (fonly_backward = false) or backward on `fonly`
int step_large(const bfvi_model* m, const float* params, float* grads, const bfvi_step_args* a,
               const bfvi_filter_args* fonly, bool fonly_backward, void* workspace, size_t workspace_bytes,
               float* loss_out, int32_t* launches, cudaStream_t st, const TileCtx* tile = nullptr) {
  // `pl`, `grads` and `st` are the CURRENT context of every helper below (captured by reference):
  // the main one, or — while pass A is being queued — the side scratch set, gradient buffer and stream
  LargePlan pl_main, pl_side;
  plan_large(m, a, fonly, &pl_main);
  pl_side = pl_main;
  if (!fonly) plan_large_side(m, a, pl_main, &pl_side);
  LargePlan pl = pl_main;
  float* const grads_main = grads;
  const cudaStream_t st_main = st;
  if (workspace_bytes < pl_side.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, pl_side.total);
  if (((uintptr_t)workspace & 255) != 0) return fail(BFVI_ERR_ARG, "workspace must be 256-byte aligned");
  char* ws = (char*)workspace;
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  const int M = m->n_mods, Z = m->z_dim, H = m->h_dim, S = pl.S;
  const int T = fonly ? fonly->T : a->T, B = fonly ? fonly->B : a->B;
  const bool with_grad = fonly ? fonly_backward : grads != nullptr;
  const int prec = (a != nullptr && a->precision == BFVI_PREC_TF32) ? bfvi::tc::PREC_TF32 : bfvi::tc::PREC_TF32X3;
  const int64_t tb = (int64_t)pl.tb;
  int n_launch = 0;
  auto F = [&](size_t off) { return (float*)(ws + off); };
  float* paramsT = F(pl.paramsT);
  double* acc = tile ? tile->acc : (double*)(ws + pl.acc);
  float* count = F(pl.count);
  const bool first_tile = tile == nullptr || tile->first, last_tile = tile == nullptr || tile->last;
  auto ew_grid = [&](int64_t n, int per) { return dim3((unsigned)grid_for(n, per, 16)); };

  if (!fonly) {                      // a lone filter ACCUMULATES into the caller's (zeroed) gradient buffers
    cudaMemsetAsync(ws + pl.zero_begin, 0, pl.zero_end - pl.zero_begin, st);
    cudaMemsetAsync(ws + pl_side.zero_begin, 0, pl_side.zero_end - pl_side.zero_begin, st);
    if (with_grad && first_tile) cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)lay.total, st);
  }
  BFVI_CHECK_CUDA();

  // ---- GEMM helpers ---------------------------------------------------------------------
  // GEMMs queue into `pending`; flush() runs everything queued as ONE grouped launch.  The code below
  // flushes wherever a later GEMM (or elementwise kernel) reads what a queued one writes, or two would
  // accumulate into the same matrix.  BFVI_GEMM_GROUP=0: one launch per GEMM (A/B timing, bisecting).
  static const bool grouping = [] { const char* e = getenv("BFVI_GEMM_GROUP"); return !e || atoi(e) != 0; }();
  bfvi::tc::GemmParams pending[bfvi::tc::kMaxGroup];
  int n_pending = 0;
  auto flush = [&]() -> int {
    if (n_pending == 0) return BFVI_OK;
    ++n_launch;
    const int rc = gemm_group_tc(pending, n_pending, prec, st);
    n_pending = 0;
    return rc;
  };
  auto gemm = [&](const bfvi::tc::GemmParams& gp) -> int {
    pending[n_pending++] = gp;
    if (!grouping || n_pending == bfvi::tc::kMaxGroup) return flush();
    return BFVI_OK;
  };
  // y = act(x W^T + b), optional transposed copy yT
  auto lin = [&](const float* x, int64_t ldx, int64_t w_off, int64_t b_off, float* y, float* yT, int64_t rows,
                 int n_in, int n_out, int act) -> int {
    bfvi::tc::GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.A = x; gp.lda = ldx; gp.W = params + w_off; gp.ldw = n_in; gp.bias = params + b_off;
    gp.C = y; gp.ldc = n_out; gp.M = rows; gp.N = n_out; gp.K = n_in; gp.act = act;
    gp.Ct = yT; gp.ldct = rows;
    return gemm(gp);
  };
  // dx (+)= dy W  (W stored (n_out, n_in); its transposed copy is the K-major B operand)
  auto dgrad = [&](const float* dy, int64_t w_off, float* dx, float* dxT, int64_t rows, int n_out, int n_in,
                   bool accumulate, const float* relu_aux, float* bias_grad) -> int {
    bfvi::tc::GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.A = dy; gp.lda = n_out; gp.W = paramsT + w_off; gp.ldw = n_out;
    gp.C = dx; gp.ldc = n_in; gp.M = rows; gp.N = n_in; gp.K = n_out;
    gp.accumulate = accumulate ? 1 : 0; gp.mask_aux = relu_aux; gp.ldaux = n_in;
    gp.Ct = dxT; gp.ldct = rows; gp.colsum = bias_grad;
    return gemm(gp);
  };
  // dW (n_out, n_in) += dy^T x from the transposed copies
  auto wgrad = [&](const float* dyT, const float* xT, int64_t rows, int n_out, int n_in, int64_t w_off, int64_t ld_dy = 0,
                   int64_t ld_x = 0) -> int {
    bfvi::tc::GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    if (ld_dy == 0) ld_dy = rows;               // leading dimensions of the transposed operands (a row chunk of a
    if (ld_x == 0) ld_x = rows;                 // (width, T*B) array keeps the array's)
    if (wgrad_swapped(n_out, n_in)) {           // dW^T = X^T dY, added into dW transposed
      gp.A = xT; gp.lda = ld_x; gp.W = dyT; gp.ldw = ld_dy;
      gp.C = grads + w_off; gp.ldc = n_in; gp.M = n_in; gp.N = n_out; gp.K = rows; gp.accumulate = 1;
      gp.trans_out = 1;
      gp.k_split = wgrad_k_split(rows, n_in, n_out);
      return gemm(gp);
    }
    gp.A = dyT; gp.lda = ld_dy; gp.W = xT; gp.ldw = ld_x;
    gp.C = grads + w_off; gp.ldc = n_in; gp.M = n_out; gp.N = n_in; gp.K = rows; gp.accumulate = 1;
    gp.k_split = wgrad_k_split(rows, n_out, n_in);
    return gemm(gp);
  };
  auto transpose = [&](const float* in, int64_t rows, int cols, float* out, int swz = 0) {
    auto k = bfvi::gen::transpose_kernel;
    BFVI_LAUNCH(k, dim3((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32)), dim3(256), 0, st, in, rows, cols, out, swz);
    ++n_launch;
  };

  // ---- transposed weights (input-gradient GEMMs) ------------------------------------------
  if (with_grad) {
    auto tw = [&](int64_t w_off, int n_out, int n_in) { transpose(params + w_off, n_out, n_in, paramsT + w_off); };
    for (int i = 0; i < (fonly ? 0 : M); ++i) {
      tw(lay.enc[i].in_to_h_w, H, m->dims[i]); tw(lay.enc[i].mean_w, Z, H); tw(lay.enc[i].std_w, Z, H);
      tw(lay.dec[i].in_to_h_w, H, Z); tw(lay.dec[i].mean_w, m->dims[i], H); tw(lay.dec[i].std_w, m->dims[i], H);
    }
    for (int d = 0; d < 2; ++d) {
      const bfvi_gtf_layout& g = lay.trans[d];
      tw(g.gate0_w, H, Z); tw(g.gate2_w, Z, H); tw(g.lin_w, Z, Z);
      tw(g.nonlin0_w, H, Z); tw(g.nonlin2_w, Z, H); tw(g.std_w, Z, Z);
    }
    BFVI_CHECK_CUDA();
  }

  // ---- fused on-chip transition kernels (precision BFVI_PREC_TF32, bfvi_fused.cuh) ---------------
  const bool fused = pl_main.fused != 0;
#ifndef BFVI_EMU
  auto fbufs = [&]() {                // packs are shared; the per-row scratch belongs to the CURRENT context (main / side)
    FusedBufs fb;
    fused_carve_packs(ws + pl_main.fused_packs, H, &fb);
    fused_carve_rows(ws + pl.fused_rows, H, (int64_t)pl.R, &fb);
    return fb;
  };
  if (fused) {                        // weights rounded to TF32 and laid out as shared-memory images ONCE per step
    const FusedBufs fb = fbufs();
    {                                 // gradient-scale statistics of both contexts start at zero
      FusedBufs fs;
      fused_carve_rows(ws + pl_side.fused_rows, H, (int64_t)pl_side.R, &fs);
      cudaMemsetAsync(fb.gstat, 0, 32, st);
      if (!fonly) cudaMemsetAsync(fs.gstat, 0, 32, st);
    }
    for (int d = 0; d < 2; ++d) { if (int rc = fused_pack(lay.trans[d], params, d, H, fb, st)) return rc; n_launch += 2; }
  }
#endif
  // ---- one transition: 6 forward GEMMs over `rows` particles --------------------------------
  // Groups are width-homogeneous (a launch has ONE tile width: a 64-wide problem in a 128-wide launch copies,
  // rounds and multiplies a half-empty W tile): the two z -> hidden layers (N = H) go out alone, everything
  // that is Z wide rides with the next level.
  auto trans_fwd = [&](const bfvi_gtf_layout& g, int64_t rows, bool keep) -> int {
#ifndef BFVI_EMU
    if (fused) {                      // ONE launch: hidden activations stay in tensor memory
      if (int rc = flush()) return rc;
      const int dir = (&g == &lay.trans[1]) ? 1 : 0;
      if (int rc = fused_fwd(fbufs(), dir, H, F(pl.zrows), rows, F(pl.g), F(pl.nl), F(pl.lin), F(pl.as), keep, st)) return rc;
      ++n_launch;
      if (keep) transpose(F(pl.nl), rows, Z, F(pl.nlT), 1);     // operand of the Z x Z std weight gradient (nl is swz64)
      return BFVI_OK;
    }
#endif
    if (int rc = lin(F(pl.zrows), Z, g.gate0_w, g.gate0_b, F(pl.h1), keep ? F(pl.h1T) : nullptr, rows, Z, H, 1)) return rc;
    if (int rc = lin(F(pl.zrows), Z, g.nonlin0_w, g.nonlin0_b, F(pl.h3), keep ? F(pl.h3T) : nullptr, rows, Z, H, 1)) return rc;
    if (int rc = flush()) return rc;
    if (int rc = lin(F(pl.h1), H, g.gate2_w, g.gate2_b, F(pl.g), nullptr, rows, H, Z, 0)) return rc;
    if (int rc = lin(F(pl.h3), H, g.nonlin2_w, g.nonlin2_b, F(pl.nl), keep ? F(pl.nlT) : nullptr, rows, H, Z, 0)) return rc;
    if (int rc = lin(F(pl.zrows), Z, g.lin_w, g.lin_b, F(pl.lin), nullptr, rows, Z, Z, 0)) return rc;
    if (int rc = flush()) return rc;
    if (int rc = lin(F(pl.nl), Z, g.std_w, g.std_b, F(pl.as), nullptr, rows, Z, Z, 0)) return rc;
    return flush();
  };
  // backward of one transition given d_as / d_g / d_lin / d_nl(partial) rows: input gradient dz (+ dz2, its
  // second partial: summed by bwd_carry_kernel, so that the two H -> Z input gradients can share a launch)
  // and the weight gradients (bias gradients come from the elementwise kernels / GEMM column sums):
  // twelve GEMMs in three dependency levels
  auto trans_bwd = [&](const bfvi_gtf_layout& g, int64_t rows) -> int {
#ifndef BFVI_EMU
    if (fused) {
      const int dir = (&g == &lay.trans[1]) ? 1 : 0;
      // Z-wide level on the launch-sequence GEMMs: d_nl += d_as W_std (its column sums complete nonlin2_b's
      // gradient) and the two Z x Z weight gradients
      if (int rc = dgrad(F(pl.d_as), g.std_w, F(pl.d_nl), nullptr, rows, Z, Z, true, nullptr, grads + g.nonlin2_b)) return rc;
      if (int rc = wgrad(F(pl.d_linT), F(pl.zrowsT), rows, Z, Z, g.lin_w)) return rc;
      if (int rc = wgrad(F(pl.d_asT), F(pl.nlT), rows, Z, Z, g.std_w)) return rc;
      if (int rc = flush()) return rc;
      const FusedBufs fb = fbufs();
      // input gradient with the hidden gradients on-chip (+ hidden bias gradients, FP16 operand tiles) ...
      if (int rc = fused_bwd(fb, dir, H, F(pl.d_g), F(pl.d_nl), F(pl.d_lin), rows, F(pl.dz), st)) return rc;
      // ... and the four H-wide weight gradients (+ the two hidden bias gradients) from those tiles
      if (int rc = fused_wgrad(fb, H, rows, grads + g.gate0_w, grads + g.nonlin0_w, grads + g.gate2_w, grads + g.nonlin2_w,
                               grads + g.gate0_b, grads + g.nonlin0_b, st)) return rc;
      n_launch += 2;
      return BFVI_OK;
    }
#endif
    // level 1 (Z wide): what needs only the head gradients and the saved activations
    if (int rc = dgrad(F(pl.d_as), g.std_w, F(pl.d_nl), F(pl.d_nlT), rows, Z, Z, true, nullptr, grads + g.nonlin2_b)) return rc;
    if (int rc = dgrad(F(pl.d_lin), g.lin_w, F(pl.dz), nullptr, rows, Z, Z, false, nullptr, nullptr)) return rc;
    if (int rc = wgrad(F(pl.d_linT), F(pl.zrowsT), rows, Z, Z, g.lin_w)) return rc;
    if (int rc = wgrad(F(pl.d_gT), F(pl.h1T), rows, Z, H, g.gate2_w)) return rc;
    if (int rc = wgrad(F(pl.d_asT), F(pl.nlT), rows, Z, Z, g.std_w)) return rc;
    if (int rc = flush()) return rc;
    // level 2 (H wide): the two head -> hidden input gradients (d_nl is complete now)
    if (int rc = dgrad(F(pl.d_g), g.gate2_w, F(pl.dh1), F(pl.dh1T), rows, Z, H, false, F(pl.h1), grads + g.gate0_b)) return rc;
    if (int rc = dgrad(F(pl.d_nl), g.nonlin2_w, F(pl.dh3), F(pl.dh3T), rows, Z, H, false, F(pl.h3), grads + g.nonlin0_b)) return rc;
    if (int rc = flush()) return rc;
    // level 3 (Z wide): hidden -> particle input gradients and the remaining weight gradients
    if (int rc = dgrad(F(pl.dh1), g.gate0_w, F(pl.dz), nullptr, rows, H, Z, true, nullptr, nullptr)) return rc;
    if (int rc = dgrad(F(pl.dh3), g.nonlin0_w, F(pl.dz2), nullptr, rows, H, Z, false, nullptr, nullptr)) return rc;
    if (int rc = wgrad(F(pl.dh1T), F(pl.zrowsT), rows, H, Z, g.gate0_w)) return rc;
    if (int rc = wgrad(F(pl.dh3T), F(pl.zrowsT), rows, H, Z, g.nonlin0_w)) return rc;
    if (int rc = wgrad(F(pl.d_nlT), F(pl.h3T), rows, Z, H, g.nonlin2_w)) return rc;
    return flush();
  };
  auto step_params = [&](const bfvi_filter_args& f, int i) {
    bfvi::gen::StepParams sp;
    memset(&sp, 0, sizeof(sp));
    const bfvi_gtf_layout& g = lay.trans[f.direction == BFVI_DIR_BWD ? 1 : 0];
    sp.a = f;
    sp.swz = fused ? 1 : 0;
    sp.z0_mean = params + lay.z0_mean; sp.z0_log_std = params + lay.z0_log_std;
    sp.g_z0_mean = grads ? grads + lay.z0_mean : nullptr; sp.g_z0_log_std = grads ? grads + lay.z0_log_std : nullptr;
    sp.min_std = m->min_std; sp.Z = Z; sp.i = i; sp.R = (int64_t)f.S * f.B * f.n_particles;
    sp.g = F(pl.g); sp.nl = F(pl.nl); sp.lin = F(pl.lin); sp.as = F(pl.as);
    sp.zrows = F(pl.zrows); sp.zrowsT = F(pl.zrowsT);
    sp.c_mu = F(pl.c_mu); sp.c_sd = F(pl.c_sd); sp.d_pm = F(pl.d_pm); sp.d_v = F(pl.d_v);
    sp.d_as = F(pl.d_as); sp.d_asT = F(pl.d_asT); sp.d_g = F(pl.d_g); sp.d_gT = F(pl.d_gT);
    sp.d_lin = F(pl.d_lin); sp.d_linT = F(pl.d_linT); sp.d_nl = F(pl.d_nl); sp.d_nlT = F(pl.d_nlT);
    if (grads) {
      sp.gb_std = grads + g.std_b; sp.gb_gate2 = grads + g.gate2_b;
      sp.gb_lin = grads + g.lin_b; sp.gb_nonlin2 = grads + g.nonlin2_b;
    }
    sp.dz = F(pl.dz); sp.dz2 = fused ? nullptr : F(pl.dz2);      // the fused backward writes the complete gradient
#ifndef BFVI_EMU
    sp.gmax = fused ? fbufs().gstat : nullptr;                   // maxima of the head gradients for the FP16 tile scale
#endif
    return sp;
  };
  // particles of step i_src as GEMM input rows (+ transposed copy)
  auto sample_rows = [&](const bfvi::gen::StepParams& sp, int i_src, int64_t rows) {
    const bool vec4 = Z % 4 == 0 && (((uintptr_t)sp.a.infer_mean | (uintptr_t)sp.a.infer_std | (uintptr_t)sp.zrows) & 15) == 0;
    if (vec4) {
      auto ks = bfvi::gen::sample_rows4_kernel;
      BFVI_LAUNCH(ks, dim3((unsigned)((rows + 31) / 32), (unsigned)((Z + 63) / 64)), dim3(64), 0, st, sp, i_src);
    } else {
      auto ks = bfvi::gen::sample_rows_kernel;
      BFVI_LAUNCH(ks, ew_grid(rows * Z, 256), dim3(256), 0, st, sp, i_src);
    }
    ++n_launch;
  };
  auto pass_fwd = [&](const bfvi_filter_args& f, float* samplesT) -> int {
    const bfvi_gtf_layout& g = lay.trans[f.direction == BFVI_DIR_BWD ? 1 : 0];
    const int64_t chains = (int64_t)f.S * B, rows = chains * f.n_particles;
    for (int i = 0; i < T; ++i) {
      if (i > 0) { if (int rc = trans_fwd(g, rows, false)) return rc; }
      bfvi::gen::StepParams sp = step_params(f, i);
      sp.zrowsT = nullptr; sp.samplesT = samplesT;
      if (step4_ok(sp)) { auto k = bfvi::gen::step4_kernel; BFVI_LAUNCH(k, ew_grid(chains * (Z / 4), 128), dim3(128), 0, st, sp); }
      else { auto k = bfvi::gen::step_kernel; BFVI_LAUNCH(k, ew_grid(chains * Z, 128), dim3(128), 0, st, sp); }
      ++n_launch;
    }
    BFVI_CHECK_CUDA();
    return BFVI_OK;
  };
  auto pass_bwd = [&](const bfvi_filter_args& f) -> int {
    const bfvi_gtf_layout& g = lay.trans[f.direction == BFVI_DIR_BWD ? 1 : 0];
    const int64_t chains = (int64_t)f.S * B, rows = chains * f.n_particles;
    cudaMemsetAsync(F(pl.c_mu), 0, sizeof(float) * chains * Z, st);
    cudaMemsetAsync(F(pl.c_sd), 0, sizeof(float) * chains * Z, st);
    for (int i = T - 1; i >= 0; --i) {
      bfvi::gen::StepParams sp = step_params(f, i);
      auto kh = bfvi::gen::bwd_head_kernel;
      BFVI_LAUNCH(kh, ew_grid(chains * Z, 128), dim3(128), 0, st, sp);
      ++n_launch;
      if (i == 0) break;
      sample_rows(sp, i - 1, rows);                            // particles of step i-1
      if (int rc = trans_fwd(g, rows, true)) return rc;
      auto kr = bfvi::gen::bwd_rows_kernel;
      BFVI_LAUNCH(kr, dim3((unsigned)((rows + bfvi::gen::kRowsPerBlock - 1) / bfvi::gen::kRowsPerBlock),
                           (unsigned)((Z + bfvi::gen::kRowsZ - 1) / bfvi::gen::kRowsZ)), dim3(bfvi::gen::kRowsZ, bfvi::gen::kRowsSplit), 0, st, sp);
      ++n_launch;
      if (int rc = trans_bwd(g, rows)) return rc;
      auto kc = bfvi::gen::bwd_carry_kernel;
      BFVI_LAUNCH(kc, ew_grid(chains * Z, 128), dim3(128), 0, st, sp, i - 1);
      ++n_launch;
    }
    BFVI_CHECK_CUDA();
    return BFVI_OK;
  };

  if (fonly) {                       // stand-alone MultiDMM.z_filter (models/dmm.py:319-412)
    if (int rc = fonly_backward ? pass_bwd(*fonly) : pass_fwd(*fonly, nullptr)) return rc;
    if (int rc = flush()) return rc;
    if (launches) *launches = n_launch;
    return BFVI_OK;
  }
  const bool external = a->eps_filt != nullptr || a->eps_sflt != nullptr || a->eps_ssmt != nullptr ||
                        a->eps_match != nullptr;
  // ---- prior-matching term (models/dmm.py:540-545) -----------------------------------------
  if (a->match_mult > 0.f && (tile == nullptr || tile->match)) {
    if (external && !a->eps_match) return fail(BFVI_ERR_ARG, "eps_match missing");
    const float* cnt = nullptr;
    float coef = a->match_mult * a->kld_mult;
    if (tile != nullptr && tile->count != nullptr) {
      cnt = tile->count;                            // mask.sum() over the WHOLE batch, counted by the tile walker
    } else if (a->match_count < 0.f) {
      auto k = bfvi::count_mask_kernel;
      BFVI_LAUNCH(k, dim3(grid_for(tb, 256, 4)), dim3(256), 0, st, a->seq_mask, tb, count);
      ++n_launch;
      cnt = count;
    } else {
      coef *= a->match_count;
    }
    const int Km = a->match_particles;
    float* zvec = F(pl.zvec);                       // [mu | sd | pm | unused] x Z
    auto kf = prior_fill_kernel;
    BFVI_LAUNCH(kf, dim3((Z + 127) / 128), dim3(128), 0, st, params + lay.z0_mean, params + lay.z0_log_std,
                m->min_std, Z, zvec, zvec + Z);
    ++n_launch;
    for (int dir = 0; dir < 2; ++dir) {
      bfvi_filter_args f;
      memset(&f, 0, sizeof(f));
      f.T = 1; f.B = 1; f.S = 1; f.n_particles = Km; f.sample = 1; f.direction = BFVI_DIR_FWD;
      f.noise.eps = a->eps_match ? a->eps_match + (size_t)dir * Km * Z : nullptr;
      f.noise.seed = a->seed; f.noise.seed_dev = a->seed_dev; f.noise.stream_id = 100u + dir;
      f.infer_mean = zvec; f.infer_std = zvec + Z; f.prior_mean = zvec + 2 * Z; f.prior_std = zvec + 3 * Z;
      bfvi::gen::StepParams sp = step_params(f, 0);
      const bfvi_gtf_layout& g = lay.trans[dir];
      if (grads) {
        sp.gb_std = grads + g.std_b; sp.gb_gate2 = grads + g.gate2_b;
        sp.gb_lin = grads + g.lin_b; sp.gb_nonlin2 = grads + g.nonlin2_b;
      }
      sample_rows(sp, 0, Km);
      if (int rc = trans_fwd(g, Km, with_grad)) return rc;
      bfvi::gen::MatchHeadParams mh;
      memset(&mh, 0, sizeof(mh));
      mh.z0_mean = sp.z0_mean; mh.z0_log_std = sp.z0_log_std; mh.g_z0_mean = sp.g_z0_mean; mh.g_z0_log_std = sp.g_z0_log_std;
      mh.g = F(pl.g); mh.nl = F(pl.nl); mh.lin = F(pl.lin); mh.as = F(pl.as);
      mh.pm = zvec + 2 * Z; mh.d_pm = F(pl.d_pm); mh.d_v = F(pl.d_v);
      mh.min_std = m->min_std; mh.coef_static = coef; mh.count = cnt; mh.loss_acc = acc;
      mh.K = Km; mh.Z = Z; mh.with_grad = with_grad ? 1 : 0; mh.swz = fused ? 1 : 0;
      auto km = bfvi::gen::match_head_kernel;
      BFVI_LAUNCH(km, dim3((Z + 127) / 128), dim3(128), 0, st, mh);
      ++n_launch;
      if (with_grad) {
        auto kr = bfvi::gen::bwd_rows_kernel;
        BFVI_LAUNCH(kr, dim3((unsigned)((Km + bfvi::gen::kRowsPerBlock - 1) / bfvi::gen::kRowsPerBlock),
                             (unsigned)((Z + bfvi::gen::kRowsZ - 1) / bfvi::gen::kRowsZ)), dim3(bfvi::gen::kRowsZ, bfvi::gen::kRowsSplit), 0, st, sp);
        ++n_launch;
        if (int rc = trans_bwd(g, Km)) return rc;
        auto kc = bfvi::gen::bwd_carry_kernel;
        BFVI_LAUNCH(kc, dim3((Z + 127) / 128), dim3(128), 0, st, sp, 0);
        auto kt = match_tail_kernel;
        BFVI_LAUNCH(kt, dim3((Z + 127) / 128), dim3(128), 0, st, (const float*)F(pl.c_mu), (const float*)F(pl.c_sd),
                    params + lay.z0_log_std, Z, grads + lay.z0_mean, grads + lay.z0_log_std);
        n_launch += 2;
      }
    }
    BFVI_CHECK_CUDA();
  }

  if (S > 0 && (a->f_mult != 0.f || a->s_mult != 0.f)) {
    float* obs_mean = F(pl.obs_mean); float* obs_stdpre = F(pl.obs_stdpre); float* obs_std = F(pl.obs_std);
    uint8_t* obs_mask = (uint8_t*)(ws + pl.mask);
    float* dobs_mean = F(pl.dobs_mean); float* dobs_std = F(pl.dobs_std);
    // ---- encoders (models/dmm.py:165-173) ---------------------------------------------------
    for (int i = 0; i < M; ++i) {
      const bfvi_mlp_layout& l = lay.enc[i];
      const int D = m->dims[i];
      auto kp = bfvi::gen::prep_rows_kernel;
      BFVI_LAUNCH(kp, ew_grid(tb, 256), dim3(256), 0, st, a->inputs[i], tb, D, F(pl.x0[i]), obs_mask + (size_t)i * pl.tb);
      ++n_launch;
      if (with_grad) transpose(F(pl.x0[i]), tb, D, F(pl.x0T[i]));
      for (int64_t r0 = 0; r0 < tb; r0 += (int64_t)pl.rc) {        // row chunks: the hidden layer is scratch
        const int64_t n = tb - r0 < (int64_t)pl.rc ? tb - r0 : (int64_t)pl.rc;
        if (int rc = lin(F(pl.x0[i]) + r0 * D, D, l.in_to_h_w, l.in_to_h_b, F(pl.hdec), nullptr, n, D, H, 1)) return rc;
        if (int rc = flush()) return rc;
        if (int rc = lin(F(pl.hdec), H, l.mean_w, l.mean_b, obs_mean + (size_t)i * pl.tbz + r0 * Z, nullptr, n, H, Z, 0)) return rc;
        if (int rc = lin(F(pl.hdec), H, l.std_w, l.std_b, obs_stdpre + (size_t)i * pl.tbz + r0 * Z, nullptr, n, H, Z, 0)) return rc;
        if (int rc = flush()) return rc;
      }
      cudaMemcpyAsync(obs_std + (size_t)i * pl.tbz, obs_stdpre + (size_t)i * pl.tbz, sizeof(float) * pl.tbz,
                      cudaMemcpyDeviceToDevice, st);
      auto ks = bfvi::gen::softplus_kernel;
      BFVI_LAUNCH(ks, ew_grid(tb * Z, 256), dim3(256), 0, st, obs_std + (size_t)i * pl.tbz, tb * Z, bfvi::kMlpMinStd);
      ++n_launch;
    }
    BFVI_CHECK_CUDA();
    auto obs_expert = [&](int i) {
      bfvi_expert e;
      memset(&e, 0, sizeof(e));
      e.mean = obs_mean + (size_t)i * pl.tbz; e.std = obs_std + (size_t)i * pl.tbz;
      e.mask = obs_mask + (size_t)i * pl.tb;
      e.stride_s = 0; e.stride_t = (int64_t)B * Z; e.stride_b = Z;
      e.mstride_s = 0; e.mstride_t = B; e.mstride_b = 1;
      e.d_mean = with_grad ? dobs_mean + (size_t)i * pl.tbz : nullptr;
      e.d_std = with_grad ? dobs_std + (size_t)i * pl.tbz : nullptr;
      e.kind = BFVI_EXPERT_TENSOR;
      return e;
    };
    auto base_args = [&]() {
      bfvi_filter_args f;
      memset(&f, 0, sizeof(f));
      f.T = T; f.B = B; f.S = S;
      f.n_experts = M;
      for (int i = 0; i < M; ++i) f.experts[i] = obs_expert(i);
      for (int s = 0; s < S; ++s) f.set_expert_bits[s] = pl.set_bits[s];
      f.sample = a->sample; f.sample_init = a->sample_init;
      f.noise.seed = a->seed; f.noise.seed_dev = a->seed_dev; f.noise.b_offset = a->b_offset;
      f.seq_mask = a->seq_mask;
      f.loss_acc = acc;
      return f;
    };
    bfvi_filter_args fa = base_args();
    fa.direction = a->f_mode == BFVI_MODE_BFILTER ? BFVI_DIR_BWD : BFVI_DIR_FWD;
    fa.n_particles = 1;
    fa.noise.eps = a->eps_filt; fa.noise.stream_id = 1;
    fa.infer_mean = F(pl.pa[0]); fa.infer_std = F(pl.pa[1]); fa.prior_mean = F(pl.pa[2]); fa.prior_std = F(pl.pa[3]);
    fa.samples = F(pl.pa[4]);
    fa.kl_weight = a->f_mult * a->kld_mult;
    bfvi_filter_args fb = base_args();
    fb.direction = a->s_mode == BFVI_MODE_FSMOOTH ? BFVI_DIR_BWD : BFVI_DIR_FWD;
    fb.n_particles = a->train_particles;
    fb.sample_init = 0;
    fb.noise.eps = a->eps_sflt; fb.noise.stream_id = 2;
    fb.infer_mean = F(pl.pb[0]); fb.infer_std = F(pl.pb[1]); fb.prior_mean = F(pl.pb[2]); fb.prior_std = F(pl.pb[3]);
    fb.samples = nullptr; fb.kl_weight = 0.f; fb.loss_acc = nullptr;
    bfvi_filter_args fc = base_args();
    fc.direction = a->s_mode == BFVI_MODE_FSMOOTH ? BFVI_DIR_FWD : BFVI_DIR_BWD;
    fc.n_particles = 1;
    fc.noise.eps = a->eps_ssmt; fc.noise.stream_id = 3;
    {
      bfvi_expert e;
      memset(&e, 0, sizeof(e));
      e.mean = F(pl.pb[2]); e.std = F(pl.pb[3]); e.mask = nullptr;
      e.stride_s = (int64_t)pl.tbz; e.stride_t = (int64_t)B * Z; e.stride_b = Z;
      e.d_mean = with_grad ? F(pl.b_dpm) : nullptr; e.d_std = with_grad ? F(pl.b_dps) : nullptr;
      e.kind = BFVI_EXPERT_TENSOR; e.zero_mask_last_t = 1;
      fc.experts[M] = e;
      memset(&e, 0, sizeof(e));
      e.kind = BFVI_EXPERT_INV_PRIOR;
      fc.experts[M + 1] = e;
      fc.n_experts = M + 2;
      for (int s = 0; s < S; ++s) fc.set_expert_bits[s] = pl.set_bits[s] | (1u << M) | (1u << (M + 1));
    }
    fc.infer_mean = F(pl.pc[0]); fc.infer_std = F(pl.pc[1]); fc.prior_mean = F(pl.pc[2]); fc.prior_std = F(pl.pc[3]);
    fc.samples = F(pl.pc[4]);
    fc.kl_weight = a->s_mult * a->kld_mult;
    if (external) {
      if (a->f_mult != 0.f && !fa.noise.eps && (a->sample || a->sample_init)) return fail(BFVI_ERR_ARG, "eps_filt missing");
      if (a->s_mult != 0.f && !fb.noise.eps) return fail(BFVI_ERR_ARG, "eps_sflt missing");
      if (a->s_mult != 0.f && !fc.noise.eps && (a->sample || a->sample_init)) return fail(BFVI_ERR_ARG, "eps_ssmt missing");
    }
    const bool do_f = a->f_mult != 0.f, do_s = a->s_mult != 0.f;
    // ---- decoders + NLL, forward and backward, on the samples of one pass
    //      (models/dmm.py:207-211, models/losses.py:68-89) ----
    auto decode_pass = [&](int pass) -> int {
      const float mult = pass == 0 ? a->f_mult : a->s_mult;
      float* samp = pass == 0 ? F(pl.pa[4]) : F(pl.pc[4]);
      float* sampT = pass == 0 ? F(pl.pa[5]) : F(pl.pc[5]);
      float* dsamp = pass == 0 ? F(pl.a_dsamp) : F(pl.c_dsamp);
      for (int s = 0; s < S; ++s)
        for (int i = 0; i < M; ++i) {
          if (!((pl.set_bits[s] >> i) & 1u) || a->rec_mults[i] == 0.f) continue;
          const bfvi_mlp_layout& l = lay.dec[i];
          const int D = m->dims[i];
          for (int64_t r0 = 0; r0 < tb; r0 += (int64_t)pl.rc) {   // row chunks of the (T*B) samples: hidden layers are scratch
            const int64_t n = tb - r0 < (int64_t)pl.rc ? tb - r0 : (int64_t)pl.rc;
            const float* zs = samp + (size_t)s * pl.tbz + r0 * Z;
            if (int rc = lin(zs, Z, l.in_to_h_w, l.in_to_h_b, F(pl.hdec), with_grad ? F(pl.hdecT) : nullptr, n, Z, H, 1)) return rc;
            if (int rc = flush()) return rc;
            if (int rc = lin(F(pl.hdec), H, l.mean_w, l.mean_b, F(pl.dmean), nullptr, n, H, D, 0)) return rc;
            if (int rc = lin(F(pl.hdec), H, l.std_w, l.std_b, F(pl.dstd), nullptr, n, H, D, 0)) return rc;
            if (int rc = flush()) return rc;
            bfvi::gen::HeadParams hp;
            memset(&hp, 0, sizeof(hp));
            hp.mean = F(pl.dmean); hp.stdpre = F(pl.dstd); hp.meanT = F(pl.dmeanT); hp.stdpreT = F(pl.dstdT);
            hp.target = a->targets[i] + r0 * D; hp.row_mask = a->seq_mask ? a->seq_mask + r0 : nullptr;
            hp.gb_mean = with_grad ? grads + l.mean_b : F(pl.dz); hp.gb_std = with_grad ? grads + l.std_b : F(pl.dz);
            hp.n_rows = n; hp.D = D; hp.weight = mult * a->rec_mults[i]; hp.loss_acc = acc;
            auto kh = bfvi::gen::head_kernel;
            const int hb = head_block(D);             // thread = column: narrow heads get narrow blocks (D = 16: 32 threads)
            BFVI_LAUNCH(kh, dim3((unsigned)((n + bfvi::gen::kRowsPerBlock - 1) / bfvi::gen::kRowsPerBlock),
                                 (unsigned)((D + hb - 1) / hb)), dim3(hb), 0, st, hp);
            ++n_launch;
            if (!with_grad) continue;
            if (int rc = dgrad(F(pl.dmean), l.mean_w, F(pl.dhd), nullptr, n, D, H, false, F(pl.hdec), grads + l.in_to_h_b)) return rc;
            if (int rc = wgrad(F(pl.dmeanT), F(pl.hdecT), n, D, H, l.mean_w)) return rc;
            if (int rc = wgrad(F(pl.dstdT), F(pl.hdecT), n, D, H, l.std_w)) return rc;
            if (int rc = flush()) return rc;
            if (int rc = dgrad(F(pl.dstd), l.std_w, F(pl.dhd), F(pl.dhdT), n, D, H, true, F(pl.hdec), grads + l.in_to_h_b)) return rc;
            if (int rc = flush()) return rc;
            if (int rc = dgrad(F(pl.dhd), l.in_to_h_w, dsamp + (size_t)s * pl.tbz + r0 * Z, nullptr, n, H, Z, true, nullptr, nullptr)) return rc;
            // the transposed samples are a (Z, T*B) array per chain set: a chunk keeps its leading dimension
            if (int rc = wgrad(F(pl.dhdT), sampT + (size_t)s * pl.tbz + r0, n, H, Z, l.in_to_h_w, n, tb)) return rc;
            if (int rc = flush()) return rc;
          }
        }
      BFVI_CHECK_CUDA();
      return BFVI_OK;
    };
    // ---- pass A (f_mode filter, its decoders and its backward) is independent of passes B -> C until
    //      the encoder backward: it runs on a side stream with its own scratch set and gradient buffers,
    //      beside the particle pass; its GEMMs are latency-bound (S*B rows) and fill SMs the other
    //      stream leaves idle.  BFVI_LARGE_SIDE=0 keeps everything on the caller's stream. ----
    static const bool want_side = [] { const char* e = getenv("BFVI_LARGE_SIDE"); return !e || atoi(e) != 0; }();
    SideStreams* side = (want_side && do_f && do_s && with_grad) ? side_streams(tile ? tile->lane : 0) : nullptr;
    // fork AFTER the particle pass forward (BFVI_LARGE_SIDE=2: before it): pass A then runs beside pass C — two
    // latency-bound kernel sequences of <= 148 CTAs that share the SMs — instead of beside the 148-CTA
    // persistent launches of pass B, whose statically assigned tiles a co-running kernel only delays
    static const bool fork_early = [] { const char* e = getenv("BFVI_LARGE_SIDE"); return e && atoi(e) == 2; }();
    bool b_fwd_done = false;
    if (do_s && side && !fork_early) {
      if (int rc = pass_fwd(fb, nullptr)) return rc;
      b_fwd_done = true;
    }
    if (do_f) {
      if (side) {
        if (int rc = flush()) return rc;
        cudaEventRecord(side->fork, st_main);
        cudaStreamWaitEvent(side->stream[0], side->fork, 0);
        pl = pl_side; grads = F(pl_side.side_grads); st = side->stream[0];
        for (int i = 0; i < M; ++i) {
          fa.experts[i].d_mean = F(pl.dobs_mean) + (size_t)i * pl.tbz;
          fa.experts[i].d_std = F(pl.dobs_std) + (size_t)i * pl.tbz;
        }
      }
      if (int rc = pass_fwd(fa, with_grad ? F(pl.pa[5]) : nullptr)) return rc;
      if (int rc = decode_pass(0)) return rc;
      if (side) {                    // the whole backward of pass A follows on the side stream
        fa.d_samples = F(pl.a_dsamp);
        if (int rc = pass_bwd(fa)) return rc;
        if (int rc = flush()) return rc;
        cudaEventRecord(side->join[0], st);
        pl = pl_main; grads = grads_main; st = st_main;
      }
    }
    if (do_s) {
      if (!b_fwd_done) { if (int rc = pass_fwd(fb, nullptr)) return rc; }
      if (int rc = pass_fwd(fc, with_grad ? F(pl.pc[5]) : nullptr)) return rc;
      if (int rc = decode_pass(1)) return rc;
    }
    BFVI_CHECK_CUDA();
    // ---- backward through the three passes and the encoders ----------------------------------
    if (with_grad) {
      if (do_s) {
        fc.d_samples = F(pl.c_dsamp);
        if (int rc = pass_bwd(fc)) return rc;
        fb.d_prior_mean = F(pl.b_dpm); fb.d_prior_std = F(pl.b_dps);
        if (int rc = pass_bwd(fb)) return rc;
      }
      if (do_f && !side) {
        fa.d_samples = F(pl.a_dsamp);
        if (int rc = pass_bwd(fa)) return rc;
      }
      if (side) {                    // join: add the side stream's parameter / observation gradients
        if (int rc = flush()) return rc;
        cudaStreamWaitEvent(st_main, side->join[0], 0);
        auto ka = add_into_kernel;
        BFVI_LAUNCH(ka, dim3((unsigned)grid_for(lay.total, 256, 4)), dim3(256), 0, st, grads, (const float*)F(pl_side.side_grads), (int64_t)lay.total);
        const int64_t n_obs = (int64_t)pl.tbz * M;
        BFVI_LAUNCH(ka, dim3((unsigned)grid_for(n_obs, 256, 8)), dim3(256), 0, st, dobs_mean, (const float*)F(pl_side.dobs_mean), n_obs);
        BFVI_LAUNCH(ka, dim3((unsigned)grid_for(n_obs, 256, 8)), dim3(256), 0, st, dobs_std, (const float*)F(pl_side.dobs_std), n_obs);
        n_launch += 3;
      }
      for (int i = 0; i < M; ++i) {
        const bfvi_mlp_layout& l = lay.enc[i];
        const int D = m->dims[i];
        // d_mean / d_stdpre live in the obs_mean / obs_stdpre slots from here on (their values are spent).  Row chunks: the
        // hidden activations (and their transposed copy) are RECOMPUTED into the decoder scratch, free by now
        for (int64_t r0 = 0; r0 < tb; r0 += (int64_t)pl.rc) {
          const int64_t n = tb - r0 < (int64_t)pl.rc ? tb - r0 : (int64_t)pl.rc;
          float* henc = F(pl.hdec);
          float* hencT = F(pl.hdecT);
          if (int rc = lin(F(pl.x0[i]) + r0 * D, D, l.in_to_h_w, l.in_to_h_b, henc, hencT, n, D, H, 1)) return rc;
          if (int rc = flush()) return rc;
          float* dm = obs_mean + (size_t)i * pl.tbz + r0 * Z;
          float* dsp = obs_stdpre + (size_t)i * pl.tbz + r0 * Z;
          bfvi::gen::HeadParams hp;
          memset(&hp, 0, sizeof(hp));
          hp.mean = dm; hp.stdpre = dsp; hp.meanT = F(pl.dmeanT); hp.stdpreT = F(pl.dstdT);
          hp.d_mean_in = dobs_mean + (size_t)i * pl.tbz + r0 * Z; hp.d_std_in = dobs_std + (size_t)i * pl.tbz + r0 * Z;
          hp.gb_mean = grads + l.mean_b; hp.gb_std = grads + l.std_b;
          hp.n_rows = n; hp.D = Z; hp.weight = 1.f;
          auto kh = bfvi::gen::head_kernel;
          const int hb = head_block(Z);
          BFVI_LAUNCH(kh, dim3((unsigned)((n + bfvi::gen::kRowsPerBlock - 1) / bfvi::gen::kRowsPerBlock),
                               (unsigned)((Z + hb - 1) / hb)), dim3(hb), 0, st, hp);
          ++n_launch;
          if (int rc = dgrad(dm, l.mean_w, F(pl.dhd), nullptr, n, Z, H, false, henc, grads + l.in_to_h_b)) return rc;
          if (int rc = wgrad(F(pl.dmeanT), hencT, n, Z, H, l.mean_w)) return rc;
          if (int rc = wgrad(F(pl.dstdT), hencT, n, Z, H, l.std_w)) return rc;
          if (int rc = flush()) return rc;
          if (int rc = dgrad(dsp, l.std_w, F(pl.dhd), F(pl.dhdT), n, Z, H, true, henc, grads + l.in_to_h_b)) return rc;
          if (int rc = flush()) return rc;
          if (int rc = wgrad(F(pl.dhdT), F(pl.x0T[i]) + r0, n, H, D, l.in_to_h_w, n, tb)) return rc;
          if (int rc = flush()) return rc;
        }
      }
      BFVI_CHECK_CUDA();
    }
  }
  if (int rc = flush()) return rc;
  if (last_tile) {
    auto kfin = bfvi::finalize_loss_kernel;
    BFVI_LAUNCH(kfin, dim3(1), dim3(32), 0, st, (const double*)acc, loss_out);
    BFVI_CHECK_CUDA();
    ++n_launch;
  }
  if (launches) *launches = n_launch;
  return BFVI_OK;
}

// ---- batch tiles -------------------------------------------------------------------------------------------
// The workspace of a large-dim step is O(T * B) (saved posteriors / priors of three passes, encoder and decoder
// activations: ~100 KB per sequence-timestep at the C3 shape), so B = 8 192 x T = 1 000 cannot be one piece.  Sequences
// are independent: the step walks the batch in tiles of `bt` sequences — inputs / targets / mask of a tile are staged
// into contiguous (T, bt, D) buffers (strided 2-D copies), the tile runs as a step of its own with the noise indexed by
// the global sequence index, gradients and loss accumulate across tiles.
// Workspace budget of one batch tile: 78 % of the device's memory (a B200: 139 GB -> 1 152 sequences of T = 1 000 at
// the C3 shape; larger tiles amortise the latency-bound single-particle passes: a time step of a tile costs
// ~0.6 ms + 1.6 us per sequence, gpurun_out/r2_pass_breakdown.txt), BFVI_TILE_GB overrides.  A function of
// the TOTAL memory, so bfvi_step_workspace and bfvi_step_fwd_bwd always agree.
size_t tile_budget_bytes() {
  static const size_t b = [] {
    if (const char* e = getenv("BFVI_TILE_GB")) return (size_t)(atof(e) * (double)(1ull << 30));
#ifndef BFVI_EMU
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0) return (size_t)(0.78 * (double)total_b);
#endif
    return (size_t)40 << 30;
  }();
  return b;
}
// Optionally two tiles are in flight at a time ("lanes", BFVI_TILE_LANES=2; each lane has its own streams, staging
// buffers, tile workspace and — lane 1 — gradient buffer): a time step of a tile is ~0.6 ms of dependent latency-bound
// launches (the single-particle passes sit on the critical path B fwd -> C fwd -> C bwd -> B bwd) plus 1.6 us per sequence
// of throughput-bound work, so a second lane can fill the SMs the first one leaves idle.  Measured (C3 dims, T = 40,
// tools/r2_gpu27.sh): at EQUAL tile size two lanes win (2 x 512: 98.9 -> 84.6 ms; 2 x 1 024: 171.8 -> 158.5 ms), but the
// lanes share the memory budget, and four tiles of 512 in two lanes (168.1 ms) barely beat two tiles of 1 024 one after
// the other (171.8 ms) — the persistent 148-CTA kernels of the two lanes cannot co-reside, and at T = 1 000 the host
// enqueues a whole tile (~0.4 s of launches) before the other lane gets its first kernel.  Default: one lane.
struct TiledPlan {
  int bt, n_tiles, n_lanes;
  size_t acc, count, lane0, lane_bytes, lane_grads;       // lane l: [lane0 + l * lane_bytes, ...)
  size_t stage_in[BFVI_MAX_MODS], stage_tg[BFVI_MAX_MODS], stage_mask, tile_ws, tile_bytes;   // offsets INSIDE a lane
  size_t total;
};
int tile_lanes_wanted() {                 // read per call: bfvi_step_workspace and bfvi_step_fwd_bwd see the same value
  const char* e = getenv("BFVI_TILE_LANES");
  const int v = e ? atoi(e) : 1;
  return v < 1 ? 1 : (v > kTileLanes ? kTileLanes : v);
}
int plan_tiled(const bfvi_model* m, const bfvi_step_args* a, TiledPlan* tp) {
  const int B = a->B, T = a->T;
  auto tile_bytes_of = [&](int bt) {
    bfvi_step_args at = *a;
    at.B = bt;
    LargePlan lp, ls;
    plan_large(m, &at, nullptr, &lp);
    plan_large_side(m, &at, lp, &ls);
    return ls.total;
  };
  int bt = a->batch_tile > 0 ? a->batch_tile : B;
  if (bt > B) bt = B;
  int lanes = 1;
  if (a->batch_tile <= 0 && tile_bytes_of(B) > tile_budget_bytes()) {
    // workspace is affine in the batch: fit the budget (shared by the lanes), keep tiles a multiple of 64 sequences
    lanes = tile_lanes_wanted();
    const size_t b1 = tile_bytes_of(64), b2 = tile_bytes_of(128);
    const double per = (double)(b2 - b1) / 64.0;
    const double fixed = (double)b1 - 64.0 * per;
    int fit = (int)(((double)tile_budget_bytes() / lanes - fixed) / per);
    fit = fit / 64 * 64;
    if (fit < 64) fit = 64;
    bt = fit < B ? fit : B;
    // equalise: same tile count (a multiple of the lane count), smaller last-tile imbalance
    int n = (B + bt - 1) / bt;
    n = (n + lanes - 1) / lanes * lanes;
    bt = ((B + n - 1) / n + 63) / 64 * 64;
    if (bt > B) bt = B;
  } else if (a->batch_tile > 0 && bt < B) {
    lanes = tile_lanes_wanted();                       // explicit tile size (tests): the caller sized it
  }
  tp->bt = bt;
  tp->n_tiles = (B + bt - 1) / bt;
  if (tp->n_tiles < 2) lanes = 1;
  tp->n_lanes = lanes;
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  size_t cur = 0;
  auto carve = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
  tp->acc = carve(sizeof(double)); tp->count = carve(sizeof(float));
  tp->lane_grads = carve(lanes > 1 ? sizeof(float) * (size_t)lay.total * (lanes - 1) : 0);
  tp->lane0 = carve(0);
  cur = 0;                                             // offsets inside one lane
  if (tp->n_tiles > 1) {
    for (int i = 0; i < m->n_mods; ++i) {
      tp->stage_in[i] = carve(sizeof(float) * (size_t)T * bt * m->dims[i]);
      tp->stage_tg[i] = carve(sizeof(float) * (size_t)T * bt * m->dims[i]);
    }
    tp->stage_mask = carve((size_t)T * bt);
  }
  tp->tile_ws = carve(0);
  tp->tile_bytes = tile_bytes_of(bt);
  tp->lane_bytes = align_up(tp->tile_ws + tp->tile_bytes, 256);
  tp->total = tp->lane0 + (size_t)lanes * tp->lane_bytes;
  return BFVI_OK;
}
int step_large_tiled(const bfvi_model* m, const float* params, float* grads, const bfvi_step_args* a, void* workspace,
                     size_t workspace_bytes, float* loss_out, int32_t* launches, cudaStream_t st) {
  TiledPlan tp;
  plan_tiled(m, a, &tp);
  if (workspace_bytes < tp.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, tp.total);
  if (((uintptr_t)workspace & 255) != 0) return fail(BFVI_ERR_ARG, "workspace must be 256-byte aligned");
  char* ws = (char*)workspace;
  if (tp.n_tiles == 1)
    return step_large(m, params, grads, a, nullptr, false, ws + tp.lane0 + tp.tile_ws, tp.tile_bytes, loss_out, launches, st);
  if (a->eps_match || a->eps_filt || a->eps_sflt || a->eps_ssmt)
    return fail(BFVI_ERR_UNSUPPORTED, "external noise tensors with batch tiles: pass batch_tile >= B (parity runs are small)");
  const int T = a->T, B = a->B;
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  // lane 1 runs on a stream of its own (forked from / joined into the caller's stream by events: the call stays
  // asynchronous and stream-ordered, and capturable in a CUDA graph)
  int lanes = tp.n_lanes;
  SideStreams* lane_side = lanes > 1 ? side_streams(1) : nullptr;
  if (lanes > 1 && (lane_side == nullptr || grads == nullptr)) lanes = 1;        // (no streams / forward only: one lane)
  note_dispatch("step:batch_tiles=%d x %d lanes=%d", tp.n_tiles, tp.bt, lanes);
  double* acc = (double*)(ws + tp.acc);
  float* count = (float*)(ws + tp.count);
  cudaMemsetAsync(acc, 0, sizeof(double), st);
  cudaMemsetAsync(count, 0, sizeof(float), st);
  int n_launch = 0;
  if (a->match_mult > 0.f) {                          // mask.sum() of the whole batch (or the caller's count)
    if (a->match_count < 0.f) {
      auto k = bfvi::count_mask_kernel;
      BFVI_LAUNCH(k, dim3(grid_for((int64_t)T * B, 256, 4)), dim3(256), 0, st, a->seq_mask, (int64_t)T * B, count);
      ++n_launch;
    }                                                 // else: the tiles use the caller's count (static coefficient)
  }
  BFVI_CHECK_CUDA();
  cudaStream_t lane_st[kTileLanes] = {st, st};
  float* lane_grads[kTileLanes] = {grads, grads};
  if (lanes > 1) {
    lane_st[1] = lane_side->stream[1];                // (stream[0] of a lane's set is its pass-A side stream)
    lane_grads[1] = (float*)(ws + tp.lane_grads);
    cudaEventRecord(lane_side->fork, st);
    cudaStreamWaitEvent(lane_st[1], lane_side->fork, 0);
  }
  for (int i = 0; i < tp.n_tiles; ++i) {
    const int lane = i % lanes;
    char* lw = ws + tp.lane0 + (size_t)lane * tp.lane_bytes;
    cudaStream_t ls = lane_st[lane];
    const int b0 = i * tp.bt, bc = b0 + tp.bt <= B ? tp.bt : B - b0;
    bfvi_step_args at = *a;
    at.B = bc;
    at.b_offset = a->b_offset + (uint32_t)b0;
    for (int k = 0; k < m->n_mods; ++k) {
      const size_t D = (size_t)m->dims[k];
      float* si = (float*)(lw + tp.stage_in[k]);
      float* sg = (float*)(lw + tp.stage_tg[k]);
      cudaMemcpy2DAsync(si, bc * D * 4, a->inputs[k] + (size_t)b0 * D, (size_t)B * D * 4, bc * D * 4, T, cudaMemcpyDeviceToDevice, ls);
      cudaMemcpy2DAsync(sg, bc * D * 4, a->targets[k] + (size_t)b0 * D, (size_t)B * D * 4, bc * D * 4, T, cudaMemcpyDeviceToDevice, ls);
      at.inputs[k] = si; at.targets[k] = sg;
    }
    uint8_t* sm = (uint8_t*)(lw + tp.stage_mask);
    cudaMemcpy2DAsync(sm, bc, a->seq_mask + b0, B, bc, T, cudaMemcpyDeviceToDevice, ls);
    at.seq_mask = sm;
    BFVI_CHECK_CUDA();
    TileCtx tc;
    tc.first = i < lanes;                             // first tile of its lane: clears the lane's gradient buffer
    tc.last = lanes == 1 && i == tp.n_tiles - 1;      // one lane: the last tile finalises the loss
    tc.match = i == 0;                                // the prior-matching term is computed once
    tc.lane = lane;
    tc.acc = acc; tc.count = a->match_count < 0.f ? count : nullptr;
    int32_t l = 0;
    if (int rc = step_large(m, params, lane_grads[lane], &at, nullptr, false, lw + tp.tile_ws, tp.tile_bytes, loss_out, &l, ls, &tc)) return rc;
    n_launch += l;
  }
  if (lanes > 1) {                                    // join: add lane 1's gradients, then the loss
    cudaEventRecord(lane_side->join[1], lane_st[1]);
    cudaStreamWaitEvent(st, lane_side->join[1], 0);
    auto ka = add_into_kernel;
    BFVI_LAUNCH(ka, dim3((unsigned)grid_for(lay.total, 256, 4)), dim3(256), 0, st, grads, (const float*)lane_grads[1], (int64_t)lay.total);
    auto kfin = bfvi::finalize_loss_kernel;
    BFVI_LAUNCH(kfin, dim3(1), dim3(32), 0, st, (const double*)acc, loss_out);
    BFVI_CHECK_CUDA();
    n_launch += 2;
  }
  if (launches) *launches = n_launch;
  return BFVI_OK;
}

// batch chunk [b0, b0 + bc) of the (T, B) problem; bc = 0 means the whole batch
struct Chunk { int b0, bc; };

int decode_nll_impl(const bfvi_model* m, const bfvi_layout& lay, const float* params, float* grads, int32_t mod,
                    const float* z, const float* target, const uint8_t* row_mask, int T, int B, Chunk ck,
                    float weight, double* loss_acc, float* d_z, cudaStream_t st) {
  bfvi::MlpParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.w = params + lay.dec[mod].begin;
  mp.g = grads ? grads + lay.dec[mod].begin : nullptr;
  mp.off = mlp_offsets(lay.dec[mod]);
  mp.n_in = m->z_dim; mp.n_out = m->dims[mod];
  mp.n_rows = (int64_t)T * (ck.bc > 0 ? ck.bc : B);
  mp.B = B; mp.b0 = ck.b0; mp.bc = ck.bc;
  mp.x = z; mp.target = target; mp.row_mask = row_mask; mp.weight = weight;
  mp.loss_acc = loss_acc; mp.d_x = d_z;
  return launch_mlp_bwd(m->h_dim, true, mp, st);
}

int filter_impl(const bfvi_model* m, const bfvi_layout& lay, const float* params, float* grads,
                const bfvi_filter_args* a, Chunk ck, bool backward, cudaStream_t st) {
  bfvi::FilterParams fp = make_filter_params(m, lay, params, grads, a);
  fp.b0 = ck.b0; fp.bc = ck.bc;
  return dispatch_filter(m, lay, backward, fp, st);
}

}  // namespace

// ===========================================================================
extern "C" {

int bfvi_version(void) { return BFVI_VERSION; }
#ifndef BFVI_SOURCE_ID
#define BFVI_SOURCE_ID "unknown"
#endif
const char* bfvi_build_id(void) { return BFVI_SOURCE_ID; }
const char* bfvi_last_error(void) { return g_err.c_str(); }
const char* bfvi_last_dispatch(void) { return g_dispatch.c_str(); }
size_t bfvi_sizeof(int32_t which) {
  switch (which) {
    case BFVI_STRUCT_MODEL: return sizeof(bfvi_model);
    case BFVI_STRUCT_LAYOUT: return sizeof(bfvi_layout);
    case BFVI_STRUCT_EXPERT: return sizeof(bfvi_expert);
    case BFVI_STRUCT_NOISE: return sizeof(bfvi_noise);
    case BFVI_STRUCT_FILTER_ARGS: return sizeof(bfvi_filter_args);
    case BFVI_STRUCT_STEP_ARGS: return sizeof(bfvi_step_args);
    case BFVI_STRUCT_FORWARD_ARGS: return sizeof(bfvi_forward_args);
    case BFVI_STRUCT_CONV_GEOM: return sizeof(bfvi_conv_geom);
    default: return 0;
  }
}

int bfvi_param_layout(const bfvi_model* m, bfvi_layout* out) {
  if (int rc = check_model(m)) return rc;
  if (out == nullptr) return fail(BFVI_ERR_ARG, "layout out is null");
  memset(out, 0, sizeof(*out));
  int64_t cur = 0;
  // registration order of the reference module tree (models/dmm.py:75-116):
  // direct parameters first, then enc.*, dec.*, trans.fwd, trans.bwd
  out->z0_mean = take(cur, m->z_dim);
  out->z0_log_std = take(cur, m->z_dim);
  for (int i = 0; i < m->n_mods; ++i) mlp_layout(cur, m->dims[i], m->z_dim, m->h_dim, &out->enc[i]);
  for (int i = 0; i < m->n_mods; ++i) mlp_layout(cur, m->z_dim, m->dims[i], m->h_dim, &out->dec[i]);
  gtf_layout(cur, m->z_dim, m->h_dim, &out->trans[0]);
  gtf_layout(cur, m->z_dim, m->h_dim, &out->trans[1]);
  out->total = cur;
  return BFVI_OK;
}

int bfvi_kernel_family(const bfvi_model* m) {
  if (check_model(m)) return 0;
  return family_of(m->z_dim, m->h_dim);
}

int bfvi_encode_fwd(const bfvi_model* m, const float* params, int32_t mod, const float* x,
                    int64_t n_rows, float* mean, float* std, uint8_t* mask, void* stream) {
  if (int rc = check_model(m)) return rc;
  if (mod < 0 || mod >= m->n_mods) return fail(BFVI_ERR_ARG, "bad modality index");
  if (m->dists[mod] == BFVI_DIST_CATEGORICAL) return fail(BFVI_ERR_UNSUPPORTED, "categorical encoder is a host-side module");
  if (!params || !x || !mean || !std || !mask || n_rows < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  bfvi::MlpParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.w = params + lay.enc[mod].begin;
  mp.off = mlp_offsets(lay.enc[mod]);
  mp.n_in = m->dims[mod]; mp.n_out = m->z_dim; mp.n_rows = n_rows;
  mp.x = x; mp.mean = mean; mp.std = std; mp.mask = mask;
  return launch_mlp_fwd(m->h_dim, mp, (cudaStream_t)stream);
}

int bfvi_encode_bwd(const bfvi_model* m, const float* params, float* grads, int32_t mod,
                    const float* x, int64_t n_rows, const float* d_mean, const float* d_std,
                    void* stream) {
  if (int rc = check_model(m)) return rc;
  if (mod < 0 || mod >= m->n_mods) return fail(BFVI_ERR_ARG, "bad modality index");
  if (!params || !grads || !x || !d_mean || !d_std || n_rows < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  bfvi::MlpParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.w = params + lay.enc[mod].begin;
  mp.g = grads + lay.enc[mod].begin;
  mp.off = mlp_offsets(lay.enc[mod]);
  mp.n_in = m->dims[mod]; mp.n_out = m->z_dim; mp.n_rows = n_rows;
  mp.x = x; mp.d_mean = d_mean; mp.d_std = d_std;
  return launch_mlp_bwd(m->h_dim, false, mp, (cudaStream_t)stream);
}

int bfvi_decode_fwd(const bfvi_model* m, const float* params, int32_t mod, const float* z,
                    int64_t n_rows, float* mean, float* std, void* stream) {
  if (int rc = check_model(m)) return rc;
  if (mod < 0 || mod >= m->n_mods) return fail(BFVI_ERR_ARG, "bad modality index");
  if (m->dists[mod] != BFVI_DIST_NORMAL) return fail(BFVI_ERR_UNSUPPORTED, "only Normal decoders are fused");
  if (!params || !z || !mean || !std || n_rows < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  bfvi::MlpParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.w = params + lay.dec[mod].begin;
  mp.off = mlp_offsets(lay.dec[mod]);
  mp.n_in = m->z_dim; mp.n_out = m->dims[mod]; mp.n_rows = n_rows;
  mp.x = z; mp.mean = mean; mp.std = std; mp.mask = nullptr;
  return launch_mlp_fwd(m->h_dim, mp, (cudaStream_t)stream);
}

int bfvi_decode_bwd(const bfvi_model* m, const float* params, float* grads, int32_t mod,
                    const float* z, int64_t n_rows, const float* d_mean, const float* d_std,
                    float* d_z, void* stream) {
  if (int rc = check_model(m)) return rc;
  if (mod < 0 || mod >= m->n_mods) return fail(BFVI_ERR_ARG, "bad modality index");
  if (m->dists[mod] != BFVI_DIST_NORMAL) return fail(BFVI_ERR_UNSUPPORTED, "only Normal decoders are fused");
  if (!params || !grads || !z || !d_mean || !d_std || n_rows < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  bfvi::MlpParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.w = params + lay.dec[mod].begin;
  mp.g = grads + lay.dec[mod].begin;
  mp.off = mlp_offsets(lay.dec[mod]);
  mp.n_in = m->z_dim; mp.n_out = m->dims[mod]; mp.n_rows = n_rows;
  mp.x = z; mp.d_mean = d_mean; mp.d_std = d_std; mp.d_x = d_z;
  return launch_mlp_bwd(m->h_dim, false, mp, (cudaStream_t)stream);
}

int bfvi_decode_nll(const bfvi_model* m, const float* params, float* grads, int32_t mod,
                    const float* z, const float* target, const uint8_t* row_mask, int64_t n_rows,
                    float weight, double* loss_acc, float* d_z, void* stream) {
  if (int rc = check_model(m)) return rc;
  if (mod < 0 || mod >= m->n_mods) return fail(BFVI_ERR_ARG, "bad modality index");
  if (m->dists[mod] != BFVI_DIST_NORMAL) return fail(BFVI_ERR_UNSUPPORTED, "only Normal decoders are fused");
  if (!params || !z || !target || n_rows < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  bfvi::MlpParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.w = params + lay.dec[mod].begin;
  mp.g = grads ? grads + lay.dec[mod].begin : nullptr;
  mp.off = mlp_offsets(lay.dec[mod]);
  mp.n_in = m->z_dim; mp.n_out = m->dims[mod]; mp.n_rows = n_rows;
  mp.x = z; mp.target = target; mp.row_mask = row_mask; mp.weight = weight;
  mp.loss_acc = loss_acc; mp.d_x = d_z;
  return launch_mlp_bwd(m->h_dim, true, mp, (cudaStream_t)stream);
}

int bfvi_filter_workspace(const bfvi_model* m, const bfvi_filter_args* a, size_t* bytes) {
  if (int rc = check_model(m)) return rc;
  if (int rc = check_filter_args(m, a)) return rc;
  if (!bytes) return fail(BFVI_ERR_ARG, "bytes null");
  *bytes = 0;
  if (family_of(m->z_dim, m->h_dim) == 2) {
    LargePlan pl;
    plan_large(m, nullptr, a, &pl);
    *bytes = pl.total;
  } else if (a->n_particles > 1) {
    *bytes = seg_scratch_bytes(a->S, a->B, m->z_dim);     // optional: lets the backward pack its time segments
  }
  return BFVI_OK;
}

static int filter_large(const bfvi_model* m, const float* params, float* grads, const bfvi_filter_args* a,
                        bool backward, void* stream) {
  if (!a->workspace) return fail(BFVI_ERR_WORKSPACE, "the large-dim family needs args->workspace (bfvi_filter_workspace)");
  return step_large(m, params, grads, nullptr, a, backward, a->workspace, a->workspace_bytes, nullptr, nullptr,
                    (cudaStream_t)stream);
}

int bfvi_filter_fwd(const bfvi_model* m, const float* params, const bfvi_filter_args* a, void* stream) {
  g_dispatch.clear();
  if (int rc = check_model(m)) return rc;
  if (int rc = check_filter_args(m, a)) return rc;
  if (!params) return fail(BFVI_ERR_ARG, "params null");
  if (family_of(m->z_dim, m->h_dim) == 2) return filter_large(m, params, nullptr, a, false, stream);
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  return dispatch_filter(m, lay, false, make_filter_params(m, lay, params, nullptr, a), (cudaStream_t)stream);
}

int bfvi_filter_bwd(const bfvi_model* m, const float* params, float* grads, const bfvi_filter_args* a,
                    void* stream) {
  g_dispatch.clear();
  if (int rc = check_model(m)) return rc;
  if (int rc = check_filter_args(m, a)) return rc;
  if (!params || !grads) return fail(BFVI_ERR_ARG, "params/grads null");
  if (family_of(m->z_dim, m->h_dim) == 2) return filter_large(m, params, grads, a, true, stream);
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  return dispatch_filter(m, lay, true, make_filter_params(m, lay, params, grads, a), (cudaStream_t)stream);
}

int bfvi_kld_fwd(const float* m1, const float* s1, const float* m2, const float* s2,
                 const uint8_t* row_mask, int64_t n_rows, int32_t z_dim, double* out, void* stream) {
  if (!m1 || !s1 || !m2 || !s2 || !out || n_rows < 1 || z_dim < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double), st);
  auto k = bfvi::kld_kernel;
  BFVI_LAUNCH(k, dim3(grid_for(n_rows * z_dim, 256, 8)), dim3(256), 0, st, m1, s1, m2, s2, row_mask, n_rows,
              (int)z_dim, out, 0.f, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_kld_bwd(const float* m1, const float* s1, const float* m2, const float* s2,
                 const uint8_t* row_mask, int64_t n_rows, int32_t z_dim, float g, float* d_m1,
                 float* d_s1, float* d_m2, float* d_s2, void* stream) {
  if (!m1 || !s1 || !m2 || !s2 || !d_m1 || !d_s1 || !d_m2 || !d_s2 || n_rows < 1 || z_dim < 1)
    return fail(BFVI_ERR_ARG, "null/empty argument");
  auto k = bfvi::kld_kernel;
  BFVI_LAUNCH(k, dim3(grid_for(n_rows * z_dim, 256, 8)), dim3(256), 0, (cudaStream_t)stream, m1, s1, m2, s2,
              row_mask, n_rows, (int)z_dim, (double*)nullptr, g, d_m1, d_s1, d_m2, d_s2);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_nll_gauss_fwd(const float* mean, const float* std, const float* x, const uint8_t* row_mask,
                       int64_t n_rows, int32_t d, double* out, void* stream) {
  if (!mean || !std || !x || !out || n_rows < 1 || d < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double), st);
  auto k = bfvi::nll_gauss_kernel;
  BFVI_LAUNCH(k, dim3(grid_for(n_rows * d, 256, 8)), dim3(256), 0, st, mean, std, x, row_mask, n_rows, (int)d,
              out, 0.f, (float*)nullptr, (float*)nullptr);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_nll_gauss_bwd(const float* mean, const float* std, const float* x, const uint8_t* row_mask,
                       int64_t n_rows, int32_t d, float g, float* d_mean, float* d_std, void* stream) {
  if (!mean || !std || !x || !d_mean || !d_std || n_rows < 1 || d < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  auto k = bfvi::nll_gauss_kernel;
  BFVI_LAUNCH(k, dim3(grid_for(n_rows * d, 256, 8)), dim3(256), 0, (cudaStream_t)stream, mean, std, x, row_mask,
              n_rows, (int)d, (double*)nullptr, g, d_mean, d_std);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

namespace {
int launch_bernoulli(const float* theta, const float* x, const uint8_t* row_mask, int64_t n_rows, int d,
                     double* out, float g, float* d_theta, cudaStream_t st) {
  const int64_t n = n_rows * d;
  const bool vec = d % 4 == 0 && ((uintptr_t)theta % 16 == 0) && ((uintptr_t)x % 16 == 0) &&
                   (d_theta == nullptr || (uintptr_t)d_theta % 16 == 0);
  // 16 resident 256-thread CTAs per SM, 4 elements per thread and trip
  const dim3 grid(grid_for((n + (vec ? 3 : 0)) / (vec ? 4 : 1), 256, 8)), block(256);
  if (vec) { auto k = bfvi::nll_bernoulli_kernel<true>; BFVI_LAUNCH(k, grid, block, 0, st, theta, x, row_mask, n_rows, d, out, g, d_theta); }
  else { auto k = bfvi::nll_bernoulli_kernel<false>; BFVI_LAUNCH(k, grid, block, 0, st, theta, x, row_mask, n_rows, d, out, g, d_theta); }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}
}  // namespace

int bfvi_nll_bernoulli_fwd(const float* theta, const float* x, const uint8_t* row_mask, int64_t n_rows,
                           int32_t d, double* out, void* stream) {
  if (!theta || !x || !out || n_rows < 1 || d < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double), st);
  return launch_bernoulli(theta, x, row_mask, n_rows, d, out, 0.f, nullptr, st);
}

int bfvi_nll_bernoulli_bwd(const float* theta, const float* x, const uint8_t* row_mask, int64_t n_rows,
                           int32_t d, float g, float* d_theta, void* stream) {
  if (!theta || !x || !d_theta || n_rows < 1 || d < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  return launch_bernoulli(theta, x, row_mask, n_rows, d, nullptr, g, d_theta, (cudaStream_t)stream);
}

int bfvi_nll_categorical_fwd(const float* probs, const float* x, const uint8_t* row_mask, int64_t n_rows,
                             int32_t n_cat, double* out, void* stream) {
  if (!probs || !x || !out || n_rows < 1 || n_cat < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double), st);
  auto k = bfvi::nll_categorical_kernel;
  BFVI_LAUNCH(k, dim3(grid_for(n_rows, 256, 8)), dim3(256), 0, st, probs, x, row_mask, n_rows, (int)n_cat, out, 0.f,
              (float*)nullptr);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_nll_categorical_bwd(const float* probs, const float* x, const uint8_t* row_mask, int64_t n_rows,
                             int32_t n_cat, float g, float* d_probs, void* stream) {
  if (!probs || !x || !d_probs || n_rows < 1 || n_cat < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  auto k = bfvi::nll_categorical_kernel;
  BFVI_LAUNCH(k, dim3(grid_for(n_rows, 256, 8)), dim3(256), 0, (cudaStream_t)stream, probs, x, row_mask, n_rows,
              (int)n_cat, (double*)nullptr, g, d_probs);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

// ---- batch preparation (datasets/multiseq.py), bfvi_data.cuh ------------------------------------
namespace {
int launch_delete(bfvi::DeleteParams dp, cudaStream_t st) {
  const int64_t n = (int64_t)dp.T * dp.B * dp.D;
  const bool vec = dp.D % 4 == 0 && (uintptr_t)dp.x % 16 == 0 && (uintptr_t)dp.out % 16 == 0;
  const dim3 grid(grid_for(vec ? n / 4 : n, 256, 8)), block(256);
  if (vec) { auto k = bfvi::delete_rows_kernel<true>; BFVI_LAUNCH(k, grid, block, 0, st, dp); }
  else { auto k = bfvi::delete_rows_kernel<false>; BFVI_LAUNCH(k, grid, block, 0, st, dp); }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}
}  // namespace

int bfvi_len_to_mask(const int32_t* lengths, int32_t T, int32_t B, uint8_t* mask, void* stream) {
  if (!lengths || !mask || T < 1 || B < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  auto k = bfvi::len_to_mask_kernel;
  BFVI_LAUNCH(k, dim3(grid_for((int64_t)T * B, 256, 8)), dim3(256), 0, (cudaStream_t)stream, lengths, (int)T, (int)B, mask);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_pad_merge(const float* packed, const int64_t* row_start, int32_t T, int32_t B, int64_t D, float* out,
                   void* stream) {
  if (!packed || !row_start || !out || T < 1 || B < 1 || D < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  const int64_t n = (int64_t)T * B * D;
  const bool vec = D % 4 == 0 && (uintptr_t)packed % 16 == 0 && (uintptr_t)out % 16 == 0;
  const dim3 grid(grid_for(vec ? n / 4 : n, 256, 8)), block(256);
  if (vec) { auto k = bfvi::pad_merge_kernel<true>; BFVI_LAUNCH(k, grid, block, 0, (cudaStream_t)stream, packed, row_start, (int)T, (int)B, D, out); }
  else { auto k = bfvi::pad_merge_kernel<false>; BFVI_LAUNCH(k, grid, block, 0, (cudaStream_t)stream, packed, row_start, (int)T, (int)B, D, out); }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_unpad(const float* x, const int64_t* row_start, const int32_t* src, int32_t B, int32_t n_out,
               int64_t total_rows, int64_t D, float* packed, void* stream) {
  if (!x || !row_start || !src || !packed || B < 1 || n_out < 1 || D < 1 || total_rows < 0)
    return fail(BFVI_ERR_ARG, "null/empty argument");
  if (total_rows == 0) return BFVI_OK;
  const int64_t n = total_rows * D;
  const bool vec = D % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)packed % 16 == 0;
  const dim3 grid(grid_for(vec ? n / 4 : n, 256, 8)), block(256);
  if (vec && D >= 1024) {             // wide rows: search once per row
    auto k = bfvi::unpad_rows_kernel;
    BFVI_LAUNCH(k, dim3((unsigned)grid_for(total_rows, 1, 8)), block, 0, (cudaStream_t)stream, x, row_start, src, (int)B, (int)n_out, D, packed);
  } else if (vec) { auto k = bfvi::unpad_kernel<true>; BFVI_LAUNCH(k, grid, block, 0, (cudaStream_t)stream, x, row_start, src, (int)B, (int)n_out, D, packed); }
  else { auto k = bfvi::unpad_kernel<false>; BFVI_LAUNCH(k, grid, block, 0, (cudaStream_t)stream, x, row_start, src, (int)B, (int)n_out, D, packed); }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

// time splits of bfvi_seq_mse: enough blocks (~16 per SM) when the batch has few, long sequences
int bfvi_seq_mse_splits(int32_t T, int32_t B) {
  if (T < 1 || B < 1) return 1;
  const int sms = num_sms() > 0 ? num_sms() : 1;
  int n = (16 * sms + B - 1) / B;
  if (n > T) n = T;
  return n < 1 ? 1 : n;
}

int bfvi_seq_mse(const float* const* recon, const float* const* target, const int64_t* dims, int32_t n_mods,
                 const uint8_t* mask, const float* lengths, int32_t T, int32_t B, float* out, float* scratch,
                 int32_t n_split, void* stream) {
  if (!recon || !target || !dims || !mask || !lengths || !out || T < 1 || B < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (n_mods < 1 || n_mods > bfvi::kMaxMseMods) return fail(BFVI_ERR_ARG, "n_mods must be 1..%d", bfvi::kMaxMseMods);
  if (n_split > 1 && !scratch) return fail(BFVI_ERR_WORKSPACE, "n_split > 1 needs scratch of B * n_split floats");
  if (n_split < 1) n_split = 1;
  if (n_split > T) n_split = T;
  bfvi::SeqMseParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n_mods; ++i) {
    if (!recon[i] || !target[i] || dims[i] < 1) return fail(BFVI_ERR_ARG, "modality %d: null tensor or empty row", i);
    p.recon[i] = recon[i]; p.target[i] = target[i]; p.D[i] = dims[i];
  }
  p.n_mods = n_mods; p.T = T; p.B = B; p.mask = mask; p.lengths = lengths; p.out = out;
  p.scratch = scratch; p.n_split = n_split; p.t_per = (T + n_split - 1) / n_split;
  auto k = bfvi::seq_mse_kernel;
  BFVI_LAUNCH(k, dim3((unsigned)B, (unsigned)n_split), dim3(128), 0, (cudaStream_t)stream, p);
  if (n_split > 1) {
    auto kf = bfvi::seq_mse_finish_kernel;
    BFVI_LAUNCH(kf, dim3((unsigned)((B + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (const float*)scratch, (int)n_split,
                lengths, (int)B, out);
  }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

size_t bfvi_ssim_scratch(int32_t N, int32_t C, int32_t H, int32_t W, int32_t win) {
  if (N < 1 || C < 1 || win < 1 || H < win || W < win) return 0;
  const int tx = (W - win + 1 + bfvi::kSsimTile - 1) / bfvi::kSsimTile, ty = (H - win + 1 + bfvi::kSsimTile - 1) / bfvi::kSsimTile;
  return sizeof(float) * 2 * (size_t)N * C * tx * ty;
}

int bfvi_ssim(const float* x, const float* y, int32_t N, int32_t C, int32_t H, int32_t W, const float* win,
              int32_t win_size, float data_range, float* ssim, float* cs, float* scratch, void* stream) {
  if (!x || !y || !win || !ssim || !scratch || N < 1 || C < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (win_size < 1 || win_size > bfvi::kSsimMaxWin || win_size % 2 == 0)
    return fail(BFVI_ERR_UNSUPPORTED, "window size must be odd and <= %d", bfvi::kSsimMaxWin);
  if (H < win_size || W < win_size) return fail(BFVI_ERR_ARG, "image smaller than the window");
  if (C > 65535 || N > 65535) return fail(BFVI_ERR_UNSUPPORTED, "N and C up to 65535 per call");
  bfvi::SsimParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.y = y; p.N = N; p.C = C; p.H = H; p.W = W; p.win = win_size;
  for (int i = 0; i < win_size; ++i) p.w[i] = win[i];           // HOST array
  p.c1 = (0.01f * data_range) * (0.01f * data_range);
  p.c2 = (0.03f * data_range) * (0.03f * data_range);
  p.tiles_x = (W - win_size + 1 + bfvi::kSsimTile - 1) / bfvi::kSsimTile;
  p.tiles_y = (H - win_size + 1 + bfvi::kSsimTile - 1) / bfvi::kSsimTile;
  p.scratch = scratch;
  const dim3 grid((unsigned)(p.tiles_x * p.tiles_y), (unsigned)C, (unsigned)N);
  if (win_size == 11) { auto k = bfvi::ssim11_kernel; BFVI_LAUNCH(k, grid, dim3(256), 0, (cudaStream_t)stream, p); }
  else { auto k = bfvi::ssim_kernel; BFVI_LAUNCH(k, grid, dim3(256), 0, (cudaStream_t)stream, p); }
  const float inv = 1.f / ((float)C * (float)(H - win_size + 1) * (float)(W - win_size + 1));
  auto kf = bfvi::ssim_finish_kernel;
  BFVI_LAUNCH(kf, dim3((unsigned)((N + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, (const float*)scratch, (int)N,
              (int)(C * p.tiles_x * p.tiles_y), inv, ssim, cs);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_delete_rows(const float* x, const uint8_t* del_mask, int32_t T, int32_t B, int64_t D, float* out,
                     void* stream) {
  if (!x || !del_mask || !out || T < 1 || B < 1 || D < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  bfvi::DeleteParams dp{};
  dp.x = x; dp.out = out; dp.T = T; dp.B = B; dp.D = D; dp.del_mask = del_mask;
  return launch_delete(dp, (cudaStream_t)stream);
}

int bfvi_delete_spans(const float* x, const int32_t* lo, const int32_t* hi, const int32_t* lengths, int32_t invert,
                      int32_t T, int32_t B, int64_t D, float* out, void* stream) {
  if (!x || !lo || !hi || !out || T < 1 || B < 1 || D < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  bfvi::DeleteParams dp{};
  dp.x = x; dp.out = out; dp.T = T; dp.B = B; dp.D = D; dp.lo = lo; dp.hi = hi; dp.lengths = lengths;
  dp.invert = invert != 0;
  return launch_delete(dp, (cudaStream_t)stream);
}

int bfvi_draw_deletions(const int32_t* lengths, int32_t T, int32_t B, double frac, int32_t mode, uint64_t seed,
                        uint32_t stream_id, uint32_t b_offset, uint8_t* del_mask, void* stream) {
  if (!del_mask || T < 1 || B < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (mode != BFVI_DELETE_UNIFORM && mode != BFVI_DELETE_BURST) return fail(BFVI_ERR_ARG, "unknown deletion mode %d", (int)mode);
  if (!(frac >= 0.0 && frac <= 1.0)) return fail(BFVI_ERR_ARG, "deleted fraction %g outside [0, 1]", frac);
  auto k = bfvi::draw_deletions_kernel;
  BFVI_LAUNCH(k, dim3((B + 127) / 128), dim3(128), 0, (cudaStream_t)stream, lengths, (int)T, (int)B, frac, (int)mode, seed,
              (unsigned)stream_id, (unsigned)b_offset, del_mask);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_dump_noise(uint64_t seed, uint32_t stream_id, uint32_t b_offset, int32_t S, int32_t T, int32_t B,
                    int32_t K, int32_t Z, float* out, void* stream) {
  if (!out || S < 1 || T < 1 || B < 1 || K < 1 || Z < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  auto k = bfvi::dump_noise_kernel;
  BFVI_LAUNCH(k, dim3(grid_for((int64_t)S * T * B * K, 256, 8)), dim3(256), 0, (cudaStream_t)stream, seed,
              (unsigned)stream_id, (unsigned)b_offset, (int)S, (int)T, (int)B, (int)K, (int)Z, out);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_linear_tf32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                     int64_t ldy, int64_t n_rows, int32_t n_in, int32_t n_out, int32_t act, void* stream) {
  if (!x || !w || !y || n_rows < 1 || n_in < 1 || n_out < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (ldx < n_in || ldw < n_in || ldy < n_out) return fail(BFVI_ERR_ARG, "leading dimension too small");
  if ((act & ~0x11) != 0) return fail(BFVI_ERR_ARG, "unknown activation / precision flags %d", act);
  if ((n_rows + bfvi::tc::kBM - 1) / bfvi::tc::kBM > 0x7fffffff) return fail(BFVI_ERR_ARG, "too many rows");
  return linear_tc(x, ldx, w, ldw, bias, y, ldy, n_rows, n_in, n_out, act & 1, (act >> 4) & 1, (cudaStream_t)stream);
}

int bfvi_wgrad_tf32(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw, int64_t lddw,
                    int64_t n_rows, int32_t n_out, int32_t n_in, int32_t accumulate, int32_t flags,
                    void* stream) {
  if (!dy || !x || !dw || n_rows < 1 || n_in < 1 || n_out < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (lddy < n_rows || ldx < n_rows || lddw < n_in) return fail(BFVI_ERR_ARG, "leading dimension too small");
  bfvi::tc::GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  if (accumulate && wgrad_swapped(n_out, n_in)) {
    gp.A = x; gp.lda = ldx; gp.W = dy; gp.ldw = lddy; gp.C = dw; gp.ldc = lddw;
    gp.M = n_in; gp.N = n_out; gp.K = n_rows; gp.accumulate = 1; gp.trans_out = 1;
    gp.k_split = wgrad_k_split(n_rows, n_in, n_out);
    return gemm_tc(gp, (flags >> 4) & 1, (cudaStream_t)stream);
  }
  gp.A = dy; gp.lda = lddy; gp.W = x; gp.ldw = ldx; gp.C = dw; gp.ldc = lddw;
  gp.M = n_out; gp.N = n_in; gp.K = n_rows; gp.accumulate = accumulate ? 1 : 0;
  if (accumulate) gp.k_split = wgrad_k_split(n_rows, n_out, n_in);
  return gemm_tc(gp, (flags >> 4) & 1, (cudaStream_t)stream);
}


// ---- per-modality MLPs of the composed path (any size, weights by pointer): bfvi_mlp_* -----------------------
namespace {
struct MlpPlan { size_t x0, x0T, h, hT, da, db, daT, dbT, dh, dhT, w1T, waT, wbT, mask, total; };
int check_mlp(const bfvi_mlp_desc* d) {
  if (!d) return fail(BFVI_ERR_ARG, "mlp desc null");
  if (!d->w1 || !d->b1 || !d->wa || !d->ba) return fail(BFVI_ERR_ARG, "mlp weights null");
  if (d->n_in < 1 || d->h_dim < 1 || d->n_out < 1) return fail(BFVI_ERR_ARG, "bad mlp dims");
  if (d->head != BFVI_HEAD_GAUSSIAN && d->head != BFVI_HEAD_SOFTMAX) return fail(BFVI_ERR_ARG, "bad head");
  if (d->head == BFVI_HEAD_GAUSSIAN && (!d->wb || !d->bb)) return fail(BFVI_ERR_ARG, "Gaussian head needs wb / bb");
  if (d->emb && (d->n_classes < 1 || d->n_in != d->h_dim)) return fail(BFVI_ERR_ARG, "embedding: n_classes >= 1, n_in == h_dim");
  return BFVI_OK;
}
void plan_mlp(const bfvi_mlp_desc* d, int64_t n, bool backward, MlpPlan* pl) {
  size_t cur = 0;
  auto carve = [&](size_t floats) { size_t o = cur; cur = align_up(cur + floats * sizeof(float), 256); return o; };
  const size_t H = d->h_dim, I = d->n_in, O = d->n_out, R = (size_t)n;
  memset(pl, 0, sizeof(*pl));
  pl->x0 = carve(R * I);
  pl->h = carve(R * H);
  pl->mask = carve((R + 3) / 4);                     // u8 row mask scratch (the backward's recompute)
  if (backward) {
    pl->x0T = carve(R * I); pl->hT = carve(R * H);
    pl->da = carve(R * O); pl->db = carve(R * O); pl->daT = carve(R * O); pl->dbT = carve(R * O);
    pl->dh = carve(R * H); pl->dhT = carve(R * H);
    pl->w1T = carve(H * I); pl->waT = carve(O * H); pl->wbT = carve(O * H);
  }
  pl->total = cur;
}
// x -> MLP input rows x0 (NaN -> mask + zero fill; Embedding -> ReLU) and hidden activations h (+ transposed copies)
int mlp_hidden(const bfvi_mlp_desc* d, const float* x, int64_t n, uint8_t* mask, char* ws, const MlpPlan& pl, bool backward,
               const float** x0_out, cudaStream_t st) {
  float* x0 = (float*)(ws + pl.x0);
  const float* in = x;
  if (mask == nullptr && !d->emb) mask = (uint8_t*)(ws + pl.mask);
  auto ew = [&](int64_t work) { return dim3((unsigned)grid_for(work, 256, 8)); };
  if (d->emb) {
    auto k = bfvi::gen::embed_relu_kernel;
    BFVI_LAUNCH(k, ew(n * d->h_dim), dim3(256), 0, st, d->emb, x, n, d->h_dim, d->n_classes, x0);
    if (mask != nullptr && d->nan_mask) {            // a NaN label masks the row (models/dmm.py:165)
      auto kp = bfvi::gen::prep_rows_kernel;
      BFVI_LAUNCH(kp, ew(n), dim3(256), 0, st, x, n, 1, (float*)(ws + pl.h), mask);   // zero-filled copy is scratch
    }
    in = x0;
  } else if (d->nan_mask) {
    auto kp = bfvi::gen::prep_rows_kernel;
    BFVI_LAUNCH(kp, ew(n), dim3(256), 0, st, x, n, d->n_in, x0, mask);
    in = x0;
  }
  BFVI_CHECK_CUDA();
  bfvi::tc::GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.A = in; gp.lda = d->n_in; gp.W = d->w1; gp.ldw = d->n_in; gp.bias = d->b1;
  gp.C = (float*)(ws + pl.h); gp.ldc = d->h_dim; gp.M = n; gp.N = d->h_dim; gp.K = d->n_in; gp.act = bfvi::tc::ACT_RELU;
  if (backward) { gp.Ct = (float*)(ws + pl.hT); gp.ldct = n; }
  *x0_out = in;
  return gemm_tc(gp, bfvi::tc::PREC_TF32X3, st);
}
}  // namespace

int bfvi_mlp_workspace(const bfvi_mlp_desc* d, int64_t n_rows, int32_t backward, size_t* bytes) {
  if (int rc = check_mlp(d)) return rc;
  if (!bytes || n_rows < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  MlpPlan pl;
  plan_mlp(d, n_rows, backward != 0, &pl);
  *bytes = pl.total;
  return BFVI_OK;
}

int bfvi_mlp_fwd(const bfvi_mlp_desc* d, const float* x, int64_t n_rows, float* out_a, float* out_b, uint8_t* mask,
                 void* workspace, size_t workspace_bytes, void* stream) {
  g_dispatch.clear();
  if (int rc = check_mlp(d)) return rc;
  if (!x || !out_a || n_rows < 1 || !workspace) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (d->head == BFVI_HEAD_GAUSSIAN && !out_b) return fail(BFVI_ERR_ARG, "out_b null");
  MlpPlan pl;
  plan_mlp(d, n_rows, false, &pl);
  if (workspace_bytes < pl.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, pl.total);
  if (((uintptr_t)workspace & 255) != 0) return fail(BFVI_ERR_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const float* x0 = nullptr;
  if (int rc = mlp_hidden(d, x, n_rows, mask, ws, pl, false, &x0, st)) return rc;
  LinearGroup lg;
  lg.add((const float*)(ws + pl.h), d->h_dim, d->wa, d->h_dim, d->ba, out_a, d->n_out, n_rows, d->h_dim, d->n_out, 0);
  if (d->head == BFVI_HEAD_GAUSSIAN)
    lg.add((const float*)(ws + pl.h), d->h_dim, d->wb, d->h_dim, d->bb, out_b, d->n_out, n_rows, d->h_dim, d->n_out, 0);
  if (int rc = lg.run(bfvi::tc::PREC_TF32X3, st)) return rc;
  const int64_t no = n_rows * d->n_out;
  if (d->head == BFVI_HEAD_GAUSSIAN) {
    auto ks = bfvi::gen::softplus_kernel;
    BFVI_LAUNCH(ks, dim3((unsigned)grid_for(no, 256, 8)), dim3(256), 0, st, out_b, no, d->min_std);
  } else {
    auto ks = bfvi::gen::softmax_rows_kernel;
    BFVI_LAUNCH(ks, dim3((unsigned)grid_for(n_rows, 8, 8)), dim3(256), 0, st, out_a, n_rows, d->n_out);
  }
  note_dispatch("mlp_fwd %s%s tf32x3", d->emb ? "embed+" : "", d->head == BFVI_HEAD_SOFTMAX ? "softmax" : "gaussian");
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_mlp_bwd(const bfvi_mlp_desc* d, const bfvi_mlp_grads* g, const float* x, int64_t n_rows, const float* out_a,
                 const float* out_b, const float* d_out_a, const float* d_out_b, float* d_x, void* workspace,
                 size_t workspace_bytes, void* stream) {
  g_dispatch.clear();
  if (int rc = check_mlp(d)) return rc;
  if (!g || !x || !out_a || n_rows < 1 || !workspace) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (!g->w1 || !g->b1 || !g->wa || !g->ba) return fail(BFVI_ERR_ARG, "gradient slots null");
  const bool gauss = d->head == BFVI_HEAD_GAUSSIAN;
  if (gauss && (!out_b || !g->wb || !g->bb)) return fail(BFVI_ERR_ARG, "Gaussian head: out_b / wb / bb slots null");
  if (!gauss && !d_out_a) return fail(BFVI_ERR_ARG, "d_out_a null");
  if (d->emb && !g->emb) return fail(BFVI_ERR_ARG, "embedding gradient slot null");
  MlpPlan pl;
  plan_mlp(d, n_rows, true, &pl);
  if (workspace_bytes < pl.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, pl.total);
  if (((uintptr_t)workspace & 255) != 0) return fail(BFVI_ERR_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int H = d->h_dim, I = d->n_in, O = d->n_out;
  const int64_t n = n_rows;
  auto F = [&](size_t off) { return (float*)(ws + off); };
  auto transpose = [&](const float* in, int64_t rows, int cols, float* out) {
    auto k = bfvi::gen::transpose_kernel;
    BFVI_LAUNCH(k, dim3((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32)), dim3(256), 0, st, in, rows, cols, out, 0);
  };
  // recompute: MLP input rows and hidden activations (+ transposed copies for the weight gradients)
  const float* x0 = nullptr;
  if (int rc = mlp_hidden(d, x, n, nullptr, ws, pl, true, &x0, st)) return rc;
  transpose(x0, n, I, F(pl.x0T));
  transpose(d->w1, H, I, F(pl.w1T));
  transpose(d->wa, O, H, F(pl.waT));
  if (gauss) transpose(d->wb, O, H, F(pl.wbT));
  // head backward from the forward outputs; bias gradients
  {
    auto k = bfvi::gen::mlp_head_bwd_kernel;
    BFVI_LAUNCH(k, dim3((unsigned)grid_for(n, 8, 8)), dim3(256), 0, st, out_a, out_b, d_out_a, d_out_b, n, O, gauss ? 0 : 1,
                d->min_std, F(pl.da), F(pl.db));
    auto kc = bfvi::gen::colsum_kernel;
    const unsigned gy = (unsigned)(n / 256 + 1 < 64 ? n / 256 + 1 : 64);
    BFVI_LAUNCH(kc, dim3((unsigned)((O + 127) / 128), gy), dim3(128), 0, st, (const float*)F(pl.da), n, O, g->ba);
    transpose(F(pl.da), n, O, F(pl.daT));
    if (gauss) {
      BFVI_LAUNCH(kc, dim3((unsigned)((O + 127) / 128), gy), dim3(128), 0, st, (const float*)F(pl.db), n, O, g->bb);
      transpose(F(pl.db), n, O, F(pl.dbT));
    }
  }
  BFVI_CHECK_CUDA();
  auto wgrad = [&](const float* dyT, const float* xT, int n_out, int n_in, float* dw) {
    bfvi::tc::GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    if (wgrad_swapped(n_out, n_in)) {
      gp.A = xT; gp.lda = n; gp.W = dyT; gp.ldw = n; gp.C = dw; gp.ldc = n_in; gp.M = n_in; gp.N = n_out; gp.K = n;
      gp.accumulate = 1; gp.trans_out = 1; gp.k_split = wgrad_k_split(n, n_in, n_out);
    } else {
      gp.A = dyT; gp.lda = n; gp.W = xT; gp.ldw = n; gp.C = dw; gp.ldc = n_in; gp.M = n_out; gp.N = n_in; gp.K = n;
      gp.accumulate = 1; gp.k_split = wgrad_k_split(n, n_out, n_in);
    }
    return gp;
  };
  // level 1: head weight gradients and dh = (d_a Wa) masked by h > 0
  {
    bfvi::tc::GemmParams grp[3];
    int k = 0;
    grp[k++] = wgrad(F(pl.daT), F(pl.hT), O, H, g->wa);
    if (gauss) grp[k++] = wgrad(F(pl.dbT), F(pl.hT), O, H, g->wb);
    if (int rc = gemm_group_tc(grp, k, bfvi::tc::PREC_TF32X3, st)) return rc;
    bfvi::tc::GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.A = F(pl.da); gp.lda = O; gp.W = F(pl.waT); gp.ldw = O; gp.C = F(pl.dh); gp.ldc = H; gp.M = n; gp.N = H; gp.K = O;
    gp.mask_aux = F(pl.h); gp.ldaux = H;
    gp.colsum = g->b1;                              // column sums of every (masked) partial: the first-layer bias gradient
    if (!gauss) { gp.Ct = F(pl.dhT); gp.ldct = n; }
    if (int rc = gemm_tc(gp, bfvi::tc::PREC_TF32X3, st)) return rc;
    if (gauss) {                                    // += d_b Wb: complete now, transposed copy for the weight gradient
      gp.A = F(pl.db); gp.W = F(pl.wbT); gp.accumulate = 1; gp.Ct = F(pl.dhT); gp.ldct = n;
      if (int rc = gemm_tc(gp, bfvi::tc::PREC_TF32X3, st)) return rc;
    }
  }
  // level 2: first-layer weight gradient, input gradient
  {
    bfvi::tc::GemmParams gw = wgrad(F(pl.dhT), F(pl.x0T), H, I, g->w1);
    if (int rc = gemm_tc(gw, bfvi::tc::PREC_TF32X3, st)) return rc;
    if (d->emb || d_x) {
      bfvi::tc::GemmParams gp;
      memset(&gp, 0, sizeof(gp));
      float* dx = d->emb ? F(pl.hT) : d_x;          // embedding rows: scratch (hT is spent after the head weight gradients)
      gp.A = F(pl.dh); gp.lda = H; gp.W = F(pl.w1T); gp.ldw = H; gp.C = dx; gp.ldc = I; gp.M = n; gp.N = I; gp.K = H;
      if (d->emb) { gp.mask_aux = x0; gp.ldaux = I; }            // ReLU of the embedding rows
      if (int rc = gemm_tc(gp, bfvi::tc::PREC_TF32X3, st)) return rc;
      if (d->emb) {
        auto ke = bfvi::gen::embed_bwd_kernel;
        BFVI_LAUNCH(ke, dim3((unsigned)grid_for(n * H, 256, 8)), dim3(256), 0, st, (const float*)dx, x, n, H, d->n_classes, g->emb);
      }
    }
  }
  note_dispatch("mlp_bwd %s%s tf32x3", d->emb ? "embed+" : "", gauss ? "gaussian" : "softmax");
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

// ---- GaussianGTF on latent rows through the fused on-chip kernels (bfvi_fused.cuh) --------------------------
namespace {
#ifndef BFVI_EMU
// the Z x Z corners of a transition's backward for the stand-alone entry (exact fp32, CUDA cores; inside a step
// these ride on the grouped tcgen05 launches):  d_nl = d_nonlin + d_std_pre W_std
__global__ void __launch_bounds__(64) gtf_small_dnl_kernel(const float* __restrict__ d_nonlin, const float* __restrict__ d_std_pre,
                                                          const float* __restrict__ w_std, int64_t rows, float* __restrict__ d_nl) {
  __shared__ float ds[64];
  const int j = threadIdx.x;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    ds[j] = d_std_pre[r * 64 + j];
    __syncthreads();
    float acc = d_nonlin[r * 64 + j];
    for (int i = 0; i < 64; ++i) acc = fmaf(ds[i], w_std[i * 64 + j], acc);
    d_nl[r * 64 + j] = acc;
    __syncthreads();
  }
}
// dW (64, 64) += dY^T X, bias (64) += column sums of dY; one block per slice of rows, thread = (o, 4 inputs)
__global__ void __launch_bounds__(256) gtf_small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t rows,
                                                             float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sy[64], sx[64];
  const int o = threadIdx.x >> 2, i0 = (threadIdx.x & 3) * 16;
  float acc[16], bsum = 0.f;
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x, r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = r0 + per < rows ? r0 + per : rows;
  for (int64_t r = r0; r < r1; ++r) {
    if (threadIdx.x < 64) sy[threadIdx.x] = dy[r * 64 + threadIdx.x];
    else if (threadIdx.x < 128) sx[threadIdx.x - 64] = x ? x[r * 64 + threadIdx.x - 64] : 0.f;
    __syncthreads();
    const float y = sy[o];
    bsum += y;
    if (dw != nullptr)
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(y, sx[i0 + i], acc[i]);
    __syncthreads();
  }
  if (dw != nullptr)
    for (int i = 0; i < 16; ++i) atomicAdd(dw + o * 64 + i0 + i, acc[i]);
  if (db != nullptr && (threadIdx.x & 3) == 0) atomicAdd(db + o, bsum);
}
#endif
// in-place swz64 -> row-major of an (R, 64) array (the stand-alone entries hand plain rows to the caller): a warp per
// row, lane = 16-byte chunk of the row (16 chunks; lanes 16-31 idle)
__global__ void __launch_bounds__(256) unswz64_kernel(float* __restrict__ x, int64_t rows) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* row = reinterpret_cast<float4*>(x + r * 64);
    if (lane < 16) v = row[(lane & 8) | ((lane & 7) ^ (int)(r & 7))];       // chunk `lane` of the row-major row
    __syncwarp();
    if (lane < 16) row[lane] = v;
  }
}
struct GtfWs { size_t packs, rows, d_nl, total; };
GtfWs gtf_ws_plan(int H, int64_t n_rows) {
  GtfWs w;
  size_t cur = 0;
  auto carve = [&](size_t bytes) { size_t o = cur; cur = (cur + bytes + 1023) / 1024 * 1024; return o; };
  w.packs = carve(fused_pack_total(H));
  w.rows = carve(fused_row_scratch(H, n_rows).total);
  w.d_nl = carve(sizeof(float) * (size_t)n_rows * 64);
  w.total = cur;
  return w;
}
}  // namespace

size_t bfvi_gtf_workspace(const bfvi_model* m, int64_t n_rows) {
  if (check_model(m) || n_rows < 1 || !fused_supported(m->z_dim, m->h_dim)) return 0;
  return gtf_ws_plan(m->h_dim, n_rows).total;
}

int bfvi_gtf_fwd(const bfvi_model* m, const float* params, int32_t direction, const float* z, int64_t n_rows,
                 float* gate_pre, float* nonlin, float* lin, float* std_pre, int32_t keep, void* workspace,
                 size_t workspace_bytes, void* stream) {
  g_dispatch.clear();
  if (int rc = check_model(m)) return rc;
  if (!params || !z || !gate_pre || !nonlin || !lin || !std_pre || n_rows < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (direction != BFVI_DIR_FWD && direction != BFVI_DIR_BWD) return fail(BFVI_ERR_ARG, "bad direction");
  if (!fused_supported(m->z_dim, m->h_dim))
    return fail(BFVI_ERR_UNSUPPORTED, "fused transition kernels serve z_dim 64 with h_dim a multiple of 128 (got %d / %d)", m->z_dim, m->h_dim);
#ifdef BFVI_EMU
  (void)keep; (void)workspace; (void)workspace_bytes; (void)stream;
  return fail(BFVI_ERR_UNSUPPORTED, "tcgen05 kernels do not exist in the emulator build");
#else
  const GtfWs w = gtf_ws_plan(m->h_dim, n_rows);
  if (!workspace || workspace_bytes < w.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, w.total);
  if (((uintptr_t)workspace & 255) != 0) return fail(BFVI_ERR_ARG, "workspace must be 256-byte aligned");
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  FusedBufs fb;
  fused_carve_packs((char*)workspace + w.packs, m->h_dim, &fb);
  fused_carve_rows((char*)workspace + w.rows, m->h_dim, n_rows, &fb);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = fused_pack(lay.trans[direction], params, direction, m->h_dim, fb, st)) return rc;
  if (int rc = fused_fwd(fb, direction, m->h_dim, z, n_rows, gate_pre, nonlin, lin, std_pre, keep != 0, st)) return rc;
  for (float* x : {gate_pre, nonlin, lin, std_pre})      // the kernels write the swz64 layout; callers get plain rows
    unswz64_kernel<<<dim3((unsigned)grid_for(n_rows, 8, 8)), dim3(256), 0, st>>>(x, n_rows);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
#endif
}

int bfvi_gtf_probe(const bfvi_model* m, const float* params, int32_t direction, int32_t which, const float* z, int64_t n_rows,
                   int32_t iters, float* scratch_rows, void* workspace, size_t workspace_bytes, float* ms_per_launch, void* stream) {
  g_dispatch.clear();
  if (int rc = check_model(m)) return rc;
  if (!params || !z || !scratch_rows || !ms_per_launch || n_rows < 1 || iters < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (which < 0 || which > 3) return fail(BFVI_ERR_ARG, "which: 0 forward, 1 forward<keep>, 2 input gradient, 3 weight gradients");
  if (!fused_supported(m->z_dim, m->h_dim)) return fail(BFVI_ERR_UNSUPPORTED, "fused transition kernels do not serve this shape");
#ifdef BFVI_EMU
  (void)direction; (void)workspace; (void)workspace_bytes; (void)stream;
  return fail(BFVI_ERR_UNSUPPORTED, "tcgen05 kernels do not exist in the emulator build");
#else
  const GtfWs w = gtf_ws_plan(m->h_dim, n_rows);
  if (!workspace || workspace_bytes < w.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, w.total);
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  FusedBufs fb;
  fused_carve_packs((char*)workspace + w.packs, m->h_dim, &fb);
  fused_carve_rows((char*)workspace + w.rows, m->h_dim, n_rows, &fb);
  cudaStream_t st = (cudaStream_t)stream;
  const int H = m->h_dim;
  float* o[5];
  for (int i = 0; i < 5; ++i) o[i] = scratch_rows + (size_t)i * n_rows * 64;
  if (int rc = fused_pack(lay.trans[direction], params, direction, H, fb, st)) return rc;
  cudaMemsetAsync(fb.gstat, 0, 32, st);             // no gradient statistics in the probe: scale 1
  // operands of the later kernels: one KEEP forward (whose heads double as head gradients: any finite values do)
  if (int rc = fused_fwd(fb, direction, H, z, n_rows, o[0], o[1], o[2], o[3], true, st)) return rc;
  if (which >= 3) { if (int rc = fused_bwd(fb, direction, H, o[0], o[1], o[2], n_rows, o[4], st)) return rc; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = BFVI_OK;
  // throw-away gradient buffer for which = 3: the first floats of the row scratch are not read by that kernel
  cudaEventRecord(e0, st);
  for (int it = 0; it < iters && rc == BFVI_OK; ++it) {
    if (which == 0) rc = fused_fwd(fb, direction, H, z, n_rows, o[0], o[1], o[2], o[3], false, st);
    else if (which == 1) rc = fused_fwd(fb, direction, H, z, n_rows, o[0], o[1], o[2], o[3], true, st);
    else if (which == 2) rc = fused_bwd(fb, direction, H, o[0], o[1], o[2], n_rows, o[4], st);
    else rc = fused_wgrad(fb, H, n_rows, o[4], o[4] + (size_t)H * 64, o[4] + (size_t)2 * H * 64, o[4] + (size_t)3 * H * 64, o[3], o[3] + H, st);
  }
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc != BFVI_OK) return rc;
  *ms_per_launch = ms / (float)iters;
  BFVI_CHECK_CUDA();
  return BFVI_OK;
#endif
}

int bfvi_gtf_bwd(const bfvi_model* m, const float* params, float* grads, int32_t direction, const float* z,
                 const float* nonlin, int64_t n_rows, const float* d_gate_pre, const float* d_nonlin, const float* d_lin,
                 const float* d_std_pre, float* d_z, void* workspace, size_t workspace_bytes, void* stream) {
  g_dispatch.clear();
  if (int rc = check_model(m)) return rc;
  if (!params || !grads || !z || !nonlin || !d_gate_pre || !d_nonlin || !d_lin || !d_std_pre || !d_z || n_rows < 1)
    return fail(BFVI_ERR_ARG, "null/empty argument");
  if (direction != BFVI_DIR_FWD && direction != BFVI_DIR_BWD) return fail(BFVI_ERR_ARG, "bad direction");
  if (!fused_supported(m->z_dim, m->h_dim))
    return fail(BFVI_ERR_UNSUPPORTED, "fused transition kernels serve z_dim 64 with h_dim a multiple of 128 (got %d / %d)", m->z_dim, m->h_dim);
#ifdef BFVI_EMU
  (void)workspace; (void)workspace_bytes; (void)stream;
  return fail(BFVI_ERR_UNSUPPORTED, "tcgen05 kernels do not exist in the emulator build");
#else
  const GtfWs w = gtf_ws_plan(m->h_dim, n_rows);
  if (!workspace || workspace_bytes < w.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, w.total);
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  const bfvi_gtf_layout& g = lay.trans[direction];
  FusedBufs fb;
  fused_carve_packs((char*)workspace + w.packs, m->h_dim, &fb);
  fused_carve_rows((char*)workspace + w.rows, m->h_dim, n_rows, &fb);
  float* d_nl = (float*)((char*)workspace + w.d_nl);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)grid_for(n_rows, 8, 4);
  gtf_small_dnl_kernel<<<dim3(blocks), dim3(64), 0, st>>>(d_nonlin, d_std_pre, params + g.std_w, n_rows, d_nl);
  gtf_small_wgrad_kernel<<<dim3(blocks), dim3(256), 0, st>>>(d_std_pre, nonlin, n_rows, grads + g.std_w, grads + g.std_b);
  gtf_small_wgrad_kernel<<<dim3(blocks), dim3(256), 0, st>>>(d_lin, z, n_rows, grads + g.lin_w, grads + g.lin_b);
  gtf_small_wgrad_kernel<<<dim3(blocks), dim3(256), 0, st>>>(d_gate_pre, nullptr, n_rows, nullptr, grads + g.gate2_b);
  gtf_small_wgrad_kernel<<<dim3(blocks), dim3(256), 0, st>>>(d_nl, nullptr, n_rows, nullptr, grads + g.nonlin2_b);
  // maxima of the head gradients for the FP16 tile scale (inside a step bwd_rows_kernel keeps them)
  cudaMemsetAsync(fb.gstat, 0, 32, st);
  bfvi::fused::absmax2_kernel<<<dim3((unsigned)grid_for(n_rows * 64, 256, 8)), dim3(256), 0, st>>>(d_gate_pre, d_nl, n_rows * 64, fb.gstat);
  BFVI_CHECK_CUDA();
  if (int rc = fused_bwd(fb, direction, m->h_dim, d_gate_pre, d_nl, d_lin, n_rows, d_z, st)) return rc;
  unswz64_kernel<<<dim3((unsigned)grid_for(n_rows, 8, 8)), dim3(256), 0, st>>>(d_z, n_rows);
  return fused_wgrad(fb, m->h_dim, n_rows, grads + g.gate0_w, grads + g.nonlin0_w, grads + g.gate2_w, grads + g.nonlin2_w,
                     grads + g.gate0_b, grads + g.nonlin0_b, st);
#endif
}

int bfvi_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                   float max_norm, float* norm_scratch, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  if (max_norm > 0.f && !norm_scratch) return fail(BFVI_ERR_ARG, "norm_scratch null");
  cudaStream_t st = (cudaStream_t)stream;
  bfvi::AdamParams ap;
  ap.p = params; ap.g = grads; ap.m = exp_avg; ap.v = exp_avg_sq; ap.n = n;
  ap.lr = lr; ap.beta1 = beta1; ap.beta2 = beta2; ap.eps = eps; ap.weight_decay = weight_decay;
  ap.grad_scale = grad_scale; ap.max_norm = max_norm;
  ap.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  ap.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  ap.sqnorm = nullptr;
  if (max_norm > 0.f) {
    cudaMemsetAsync(norm_scratch, 0, sizeof(float), st);
    auto kn = bfvi::sqnorm_kernel;
    BFVI_LAUNCH(kn, dim3(grid_for(n, 256, 4)), dim3(256), 0, st, grads, n, grad_scale, norm_scratch);
    ap.sqnorm = norm_scratch;
  }
  auto k = bfvi::adam_kernel;
  BFVI_LAUNCH(k, dim3(grid_for(n, 256, 8)), dim3(256), 0, st, ap);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_ffma_probe(float* out, int32_t iters, int32_t blocks, void* stream) {
  if (!out || iters < 1 || blocks < 1) return fail(BFVI_ERR_ARG, "null/empty argument");
  auto k = bfvi::ffma_peak_kernel;
  BFVI_LAUNCH(k, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, out, (int)iters);
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

static int check_step(const bfvi_model* m, const bfvi_step_args* a) {
  if (int rc = check_model(m)) return rc;
  if (a == nullptr) return fail(BFVI_ERR_ARG, "step args null");
  if (a->T < 1 || a->B < 1) return fail(BFVI_ERR_ARG, "bad T/B");
  for (int i = 0; i < m->n_mods; ++i)
    if (m->dists[i] != BFVI_DIST_NORMAL)
      return fail(BFVI_ERR_UNSUPPORTED, "fused step covers Normal modalities; compose the ops for others");
  if (a->f_mode != BFVI_MODE_BFILTER && a->f_mode != BFVI_MODE_FFILTER) return fail(BFVI_ERR_ARG, "bad f_mode");
  if (a->s_mode != BFVI_MODE_FSMOOTH && a->s_mode != BFVI_MODE_BSMOOTH) return fail(BFVI_ERR_ARG, "bad s_mode");
  if (a->train_particles < 1 || a->match_particles < 1) return fail(BFVI_ERR_ARG, "particle counts must be >= 1");
  if (a->precision < BFVI_PREC_TF32X3 || a->precision > BFVI_PREC_FUSED) return fail(BFVI_ERR_ARG, "bad precision");
  if (a->batch_tile < 0) return fail(BFVI_ERR_ARG, "batch_tile < 0");
  return BFVI_OK;
}

int bfvi_step_workspace(const bfvi_model* m, const bfvi_step_args* a, size_t* bytes) {
  if (int rc = check_step(m, a)) return rc;
  if (!bytes) return fail(BFVI_ERR_ARG, "bytes null");
  if (family_of(m->z_dim, m->h_dim) == 2) {
    TiledPlan tp;
    plan_tiled(m, a, &tp);
    *bytes = tp.total;
    return BFVI_OK;
  }
  StepPlan pl;
  plan_step(m, a, true, &pl);
  *bytes = pl.total;
  return BFVI_OK;
}

// Marks the start of a phase on the stream when a profile is being recorded.
struct PhaseMarks {
  cudaEvent_t ev[BFVI_N_PHASES + 1];
  int used[BFVI_N_PHASES + 1];
  bool on;
  cudaStream_t st;
  void begin(int phase) {
    if (on && !used[phase]) { cudaEventRecord(ev[phase], st); used[phase] = 1; }
  }
};

static int step_impl(const bfvi_model* m, const float* params, float* grads, const bfvi_step_args* a,
                     void* workspace, size_t workspace_bytes, float* loss_out, int32_t* launches,
                     void* stream, PhaseMarks& pm) {
  g_dispatch.clear();
  if (int rc = check_step(m, a)) return rc;
  if (!params || !workspace || !loss_out) return fail(BFVI_ERR_ARG, "null argument");
  if (!a->seq_mask) return fail(BFVI_ERR_ARG, "seq_mask null");
  const int M = m->n_mods, Z = m->z_dim, T = a->T, B = a->B;
  for (int i = 0; i < M; ++i)
    if (!a->inputs[i] || !a->targets[i]) return fail(BFVI_ERR_ARG, "inputs/targets[%d] null", i);
  if (family_of(m->z_dim, m->h_dim) == 2) {
    if (pm.on) return fail(BFVI_ERR_UNSUPPORTED, "phase profile exists for the small-dim family only");
    return step_large_tiled(m, params, grads, a, workspace, workspace_bytes, loss_out, launches, (cudaStream_t)stream);
  }
  if (a->seed_dev != nullptr) return fail(BFVI_ERR_UNSUPPORTED, "seed_dev (graph-replayable seed) exists for the large-dim family only");
  const bool with_grad = grads != nullptr;
  StepPlan pl;
  plan_step(m, a, with_grad, &pl);
  if (workspace_bytes < pl.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, pl.total);
  if (((uintptr_t)workspace & 255) != 0) return fail(BFVI_ERR_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  int n_launch = 0;

  double* acc = (double*)(ws + pl.off_acc);
  float* count = (float*)(ws + pl.off_count);
  float* obs_mean = (float*)(ws + pl.off_obs_mean);
  float* obs_std = (float*)(ws + pl.off_obs_std);
  uint8_t* obs_mask = (uint8_t*)(ws + pl.off_obs_mask);
  float* dobs_mean = with_grad ? (float*)(ws + pl.off_dobs_mean) : nullptr;
  float* dobs_std = with_grad ? (float*)(ws + pl.off_dobs_std) : nullptr;
  auto A = [&](int i) { return (float*)(ws + pl.off_a[i]); };
  auto Bf = [&](int i) { return (float*)(ws + pl.off_b[i]); };
  auto C = [&](int i) { return (float*)(ws + pl.off_c[i]); };

  cudaMemsetAsync(ws + pl.zero_begin, 0, pl.zero_end - pl.zero_begin, st);
  if (with_grad) cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)lay.total, st);
  BFVI_CHECK_CUDA();

  const int64_t tb = (int64_t)T * B;
  const bool external = a->eps_filt != nullptr || a->eps_sflt != nullptr || a->eps_ssmt != nullptr ||
                        a->eps_match != nullptr;
  const int S = pl.S;

  // ---- prior-matching term (models/dmm.py:540-545) ------------------------------
  pm.begin(BFVI_PHASE_MATCH);
  if (a->match_mult > 0.f) {
    if (external && !a->eps_match) return fail(BFVI_ERR_ARG, "eps_match missing");
    bfvi::MatchParams mp;
    memset(&mp, 0, sizeof(mp));
    for (int d = 0; d < 2; ++d) {
      mp.trans_w[d] = params + lay.trans[d].begin;
      mp.g_trans[d] = with_grad ? grads + lay.trans[d].begin : nullptr;
    }
    mp.z0_mean = params + lay.z0_mean; mp.z0_log_std = params + lay.z0_log_std;
    mp.g_z0_mean = with_grad ? grads + lay.z0_mean : nullptr;
    mp.g_z0_log_std = with_grad ? grads + lay.z0_log_std : nullptr;
    mp.eps = a->eps_match; mp.seed = a->seed; mp.K = a->match_particles; mp.min_std = m->min_std;
    mp.coef_static = a->match_mult * a->kld_mult;
    if (a->match_count < 0.f) {
      auto k = bfvi::count_mask_kernel;
      BFVI_LAUNCH(k, dim3(grid_for(tb, 256, 4)), dim3(256), 0, st, a->seq_mask, tb, count);
      BFVI_CHECK_CUDA();
      ++n_launch;
      mp.count = count;
    } else {
      mp.count = nullptr;
      mp.coef_static *= a->match_count;
    }
    mp.loss_acc = acc; mp.with_grad = with_grad ? 1 : 0;
    if (int rc = dispatch_match(m, mp, st)) return rc;
    ++n_launch;
  }

  if (S > 0 && (a->f_mult != 0.f || a->s_mult != 0.f)) {
    // ---- encode every modality once (models/dmm.py:165-173) ----------------------
    pm.begin(BFVI_PHASE_ENCODE_FWD);
    for (int i = 0; i < M; ++i) {
      if (int rc = bfvi_encode_fwd(m, params, i, a->inputs[i], tb, obs_mean + (size_t)i * pl.n_tbz,
                                   obs_std + (size_t)i * pl.n_tbz, obs_mask + (size_t)i * tb, stream))
        return rc;
      ++n_launch;
    }
    auto obs_expert = [&](int i) {
      bfvi_expert e;
      memset(&e, 0, sizeof(e));
      e.mean = obs_mean + (size_t)i * pl.n_tbz; e.std = obs_std + (size_t)i * pl.n_tbz;
      e.mask = obs_mask + (size_t)i * tb;
      e.stride_s = 0; e.stride_t = (int64_t)B * Z; e.stride_b = Z;
      e.mstride_s = 0; e.mstride_t = B; e.mstride_b = 1;
      e.d_mean = with_grad ? dobs_mean + (size_t)i * pl.n_tbz : nullptr;
      e.d_std = with_grad ? dobs_std + (size_t)i * pl.n_tbz : nullptr;
      e.kind = BFVI_EXPERT_TENSOR;
      return e;
    };
    auto base_args = [&]() {
      bfvi_filter_args f;
      memset(&f, 0, sizeof(f));
      f.T = T; f.B = B; f.S = S;
      f.n_experts = M;
      for (int i = 0; i < M; ++i) f.experts[i] = obs_expert(i);
      for (int s = 0; s < S; ++s) f.set_expert_bits[s] = pl.set_bits[s];
      f.sample = a->sample; f.sample_init = a->sample_init;
      f.noise.seed = a->seed; f.noise.b_offset = a->b_offset;
      f.seq_mask = a->seq_mask;
      f.loss_acc = acc;
      return f;
    };

    // pass A: f_mode filtering ELBO (models/dmm.py:547-549)
    bfvi_filter_args fa = base_args();
    fa.direction = a->f_mode == BFVI_MODE_BFILTER ? BFVI_DIR_BWD : BFVI_DIR_FWD;
    fa.n_particles = 1;
    fa.noise.eps = a->eps_filt; fa.noise.stream_id = 1;
    fa.infer_mean = A(0); fa.infer_std = A(1); fa.prior_mean = A(2); fa.prior_std = A(3); fa.samples = A(4);
    fa.kl_weight = a->f_mult * a->kld_mult;
    // pass B: filtering pass of s_mode with K particles (models/dmm.py:465-470,551-553)
    bfvi_filter_args fb = base_args();
    fb.direction = a->s_mode == BFVI_MODE_FSMOOTH ? BFVI_DIR_BWD : BFVI_DIR_FWD;
    fb.n_particles = a->train_particles;
    fb.sample_init = 0;
    fb.noise.eps = a->eps_sflt; fb.noise.stream_id = 2;
    fb.infer_mean = Bf(0); fb.infer_std = Bf(1); fb.prior_mean = Bf(2); fb.prior_std = Bf(3);
    fb.samples = nullptr; fb.kl_weight = 0.f; fb.loss_acc = nullptr;
    // pass C: smoothing pass (models/dmm.py:473-489)
    bfvi_filter_args fc = base_args();
    fc.direction = a->s_mode == BFVI_MODE_FSMOOTH ? BFVI_DIR_FWD : BFVI_DIR_BWD;
    fc.n_particles = 1;
    fc.noise.eps = a->eps_ssmt; fc.noise.stream_id = 3;
    {
      bfvi_expert e;
      memset(&e, 0, sizeof(e));
      e.mean = Bf(2); e.std = Bf(3); e.mask = nullptr;
      e.stride_s = (int64_t)pl.n_tbz; e.stride_t = (int64_t)B * Z; e.stride_b = Z;
      e.d_mean = with_grad ? Bf(4) : nullptr; e.d_std = with_grad ? Bf(5) : nullptr;
      e.kind = BFVI_EXPERT_TENSOR; e.zero_mask_last_t = 1;
      fc.experts[M] = e;
      memset(&e, 0, sizeof(e));
      e.kind = BFVI_EXPERT_INV_PRIOR;
      fc.experts[M + 1] = e;
      fc.n_experts = M + 2;
      for (int s = 0; s < S; ++s) fc.set_expert_bits[s] = pl.set_bits[s] | (1u << M) | (1u << (M + 1));
    }
    fc.infer_mean = C(0); fc.infer_std = C(1); fc.prior_mean = C(2); fc.prior_std = C(3); fc.samples = C(4);
    fc.kl_weight = a->s_mult * a->kld_mult;
    if (external) {
      if (a->f_mult != 0.f && !fa.noise.eps && (a->sample || a->sample_init)) return fail(BFVI_ERR_ARG, "eps_filt missing");
      if (a->s_mult != 0.f && a->train_particles > 0 && !fb.noise.eps) return fail(BFVI_ERR_ARG, "eps_sflt missing");
      if (a->s_mult != 0.f && !fc.noise.eps && (a->sample || a->sample_init)) return fail(BFVI_ERR_ARG, "eps_ssmt missing");
    }

    const bool do_f = a->f_mult != 0.f, do_s = a->s_mult != 0.f;
    // decoders + NLL (+ their backward) on the samples of one pass, one batch chunk
    auto decode_pass = [&](int pass, Chunk ck, cudaStream_t strm) -> int {
      const float mult = pass == 0 ? a->f_mult : a->s_mult;
      float* samp = pass == 0 ? A(4) : C(4);
      float* dsamp = with_grad ? (pass == 0 ? A(5) : C(5)) : nullptr;
      for (int s = 0; s < S; ++s)
        for (int i = 0; i < M; ++i) {
          if (!((pl.set_bits[s] >> i) & 1u) || a->rec_mults[i] == 0.f) continue;
          if (int rc = decode_nll_impl(m, lay, params, grads, i, samp + (size_t)s * pl.n_tbz, a->targets[i],
                                       a->seq_mask, T, B, ck, mult * a->rec_mults[i], acc,
                                       dsamp ? dsamp + (size_t)s * pl.n_tbz : nullptr, strm))
            return rc;
          ++n_launch;
        }
      return BFVI_OK;
    };
    fa.d_samples = with_grad ? A(5) : nullptr;
    fc.d_samples = with_grad ? C(5) : nullptr;
    fb.d_prior_mean = with_grad ? Bf(4) : nullptr;
    fb.d_prior_std = with_grad ? Bf(5) : nullptr;
    fb.workspace = ws + pl.off_seg; fb.workspace_bytes = pl.seg_bytes;
    const Chunk all{0, 0};

    // ---- stream plan ------------------------------------------------------------------
    // Pass A and every batch chunk of passes B / C are independent branches.  While a
    // phase profile is recorded everything stays on the caller's stream, one phase at a
    // time; otherwise branches fork onto side streams and join before the encoders'
    // backward.
    int n_chunks = 1;
    SideStreams* side = pm.on ? nullptr : side_streams();
    if (side != nullptr && do_s) {
      // two chunks overlap one chunk's latency-bound single-particle passes with the other's particle
      // kernels, which pays once each chunk alone fills the GPU (measured on B200: B = 4096 and 8192 are
      // 1 % faster unchunked, B = 16384 is 2.5 % faster in two chunks)
      n_chunks = B >= 16384 ? 2 : 1;
      if (const char* env = getenv("BFVI_CHUNKS")) {       // tuning knob
        const int v = atoi(env);
        if (v >= 1 && v <= kSideStreams) n_chunks = v;
      }
      if (n_chunks > B) n_chunks = B;
    }
    note_dispatch("step:chunks=%d", n_chunks);
    const bool fork_a = side != nullptr && do_f && do_s;
    const bool forked = fork_a || n_chunks > 1;
    if (forked) {
      cudaEventRecord(side->fork, st);
      for (int i = 0; i < kSideStreams; ++i) cudaStreamWaitEvent(side->stream[i], side->fork, 0);
    }
    // chunk c of B / C runs on the caller's stream (c = 0) or side stream c; pass A on
    // side stream 0 when it is forked
    auto chunk_stream = [&](int c) { return c == 0 ? st : side->stream[c]; };
    auto chunk_of = [&](int c) {
      if (n_chunks == 1) return all;
      const int per = (B + n_chunks - 1) / n_chunks, b0 = c * per;
      return Chunk{b0, (b0 + per <= B ? per : B - b0)};
    };
    cudaStream_t st_a = fork_a ? side->stream[0] : st;

    pm.begin(BFVI_PHASE_FILTER_F_FWD);
    if (do_f) { if (int rc = filter_impl(m, lay, params, nullptr, &fa, all, false, st_a)) return rc; ++n_launch; }
    if (do_f && fork_a) {                        // the whole A pipeline runs beside B / C
      if (int rc = decode_pass(0, all, st_a)) return rc;
      if (with_grad) { if (int rc = filter_impl(m, lay, params, grads, &fa, all, true, st_a)) return rc; ++n_launch; }
    }
    if (do_s) {
      pm.begin(BFVI_PHASE_FILTER_S_FLT_FWD);
      for (int c = 0; c < n_chunks; ++c) {
        if (int rc = filter_impl(m, lay, params, nullptr, &fb, chunk_of(c), false, chunk_stream(c))) return rc;
        ++n_launch;
      }
      pm.begin(BFVI_PHASE_FILTER_S_SMT_FWD);
      for (int c = 0; c < n_chunks; ++c) {
        if (int rc = filter_impl(m, lay, params, nullptr, &fc, chunk_of(c), false, chunk_stream(c))) return rc;
        ++n_launch;
      }
    }
    pm.begin(BFVI_PHASE_DECODE_NLL);
    if (do_f && !fork_a) { if (int rc = decode_pass(0, all, st)) return rc; }
    if (do_s)
      for (int c = 0; c < n_chunks; ++c)
        if (int rc = decode_pass(1, chunk_of(c), chunk_stream(c))) return rc;
    // ---- backward through the three passes and the encoders ------------------------
    if (with_grad) {
      if (do_s) {
        pm.begin(BFVI_PHASE_FILTER_S_SMT_BWD);
        for (int c = 0; c < n_chunks; ++c) {
          if (int rc = filter_impl(m, lay, params, grads, &fc, chunk_of(c), true, chunk_stream(c))) return rc;
          ++n_launch;
        }
        pm.begin(BFVI_PHASE_FILTER_S_FLT_BWD);
        for (int c = 0; c < n_chunks; ++c) {
          if (int rc = filter_impl(m, lay, params, grads, &fb, chunk_of(c), true, chunk_stream(c))) return rc;
          ++n_launch;
        }
      }
      if (do_f && !fork_a) {
        pm.begin(BFVI_PHASE_FILTER_F_BWD);
        if (int rc = filter_impl(m, lay, params, grads, &fa, all, true, st)) return rc; ++n_launch;
      }
    }
    if (forked) {
      for (int i = 0; i < kSideStreams; ++i) {
        const bool used = (i == 0 && fork_a) || (i > 0 && i < n_chunks);
        if (!used) continue;
        cudaEventRecord(side->join[i], side->stream[i]);
        cudaStreamWaitEvent(st, side->join[i], 0);
      }
    }
    if (with_grad) {
      pm.begin(BFVI_PHASE_ENCODE_BWD);
      for (int i = 0; i < M; ++i) {
        if (int rc = bfvi_encode_bwd(m, params, grads, i, a->inputs[i], tb, dobs_mean + (size_t)i * pl.n_tbz,
                                     dobs_std + (size_t)i * pl.n_tbz, stream))
          return rc;
        ++n_launch;
      }
    }
    BFVI_CHECK_CUDA();
  }
  pm.begin(BFVI_PHASE_FINALIZE);
  auto k = bfvi::finalize_loss_kernel;
  BFVI_LAUNCH(k, dim3(1), dim3(32), 0, st, (const double*)acc, loss_out);
  BFVI_CHECK_CUDA();
  ++n_launch;
  if (launches) *launches = n_launch;
  return BFVI_OK;
}

int bfvi_step_fwd_bwd(const bfvi_model* m, const float* params, float* grads, const bfvi_step_args* a,
                      void* workspace, size_t workspace_bytes, float* loss_out, int32_t* launches,
                      void* stream) {
  PhaseMarks pm;
  pm.on = false;
  return step_impl(m, params, grads, a, workspace, workspace_bytes, loss_out, launches, stream, pm);
}

int bfvi_step_profile(const bfvi_model* m, const float* params, float* grads, const bfvi_step_args* a,
                      void* workspace, size_t workspace_bytes, float* loss_out, float* phase_ms,
                      void* stream) {
  if (!phase_ms) return fail(BFVI_ERR_ARG, "phase_ms null");
  PhaseMarks pm;
  pm.on = true;
  pm.st = (cudaStream_t)stream;
  for (int i = 0; i <= BFVI_N_PHASES; ++i) { cudaEventCreate(&pm.ev[i]); pm.used[i] = 0; }
  int rc = step_impl(m, params, grads, a, workspace, workspace_bytes, loss_out, nullptr, stream, pm);
  if (rc == BFVI_OK) {
    cudaEventRecord(pm.ev[BFVI_N_PHASES], pm.st);
    cudaEventSynchronize(pm.ev[BFVI_N_PHASES]);
    // phase i lasts from its marker to the next marker that was recorded
    for (int i = 0; i < BFVI_N_PHASES; ++i) {
      phase_ms[i] = 0.f;
      if (!pm.used[i]) continue;
      int j = i + 1;
      while (j < BFVI_N_PHASES && !pm.used[j]) ++j;
      cudaEventElapsedTime(&phase_ms[i], pm.ev[i], pm.ev[j]);
    }
    if (cudaGetLastError() != cudaSuccess) rc = fail(BFVI_ERR_CUDA, "event timing failed");
  }
  for (int i = 0; i <= BFVI_N_PHASES; ++i) cudaEventDestroy(pm.ev[i]);
  return rc;
}

// ---------------------------------------------------------------------------
// large-dim family: MultiDMM.forward as a launch sequence of tcgen05 GEMMs and fused
// elementwise kernels
// ---------------------------------------------------------------------------
struct ForwardPlan {
  int n_present, k_max, d_max;
  size_t tb, tbz;
  size_t off_obs_mean, off_obs_std, off_obs_mask, off_x0, off_h, off_flt[4], off_samples;
  size_t off_zrows, off_hrows, off_hrows2, off_g, off_nl, off_lin, off_as, off_packs;
  size_t total;
  int fused;
};

static int plan_forward(const bfvi_model* m, const bfvi_forward_args* a, ForwardPlan* pl) {
  if (int rc = check_model(m)) return rc;
  if (a == nullptr) return fail(BFVI_ERR_ARG, "forward args null");
  if (a->T < 1 || a->B < 1) return fail(BFVI_ERR_ARG, "bad T/B");
  if (a->mode < BFVI_MODE_BFILTER || a->mode > BFVI_MODE_BSMOOTH) return fail(BFVI_ERR_ARG, "bad mode");
  if (a->flt_particles < 1 || a->smt_particles < 1) return fail(BFVI_ERR_ARG, "particle counts must be >= 1");
  if (a->precision < BFVI_PREC_TF32X3 || a->precision > BFVI_PREC_FUSED) return fail(BFVI_ERR_ARG, "bad precision");
  for (int i = 0; i < m->n_mods; ++i)
    if (m->dists[i] != BFVI_DIST_NORMAL)
      return fail(BFVI_ERR_UNSUPPORTED, "bfvi_forward covers Normal modalities; compose the ops for others");
  pl->n_present = 0; pl->d_max = 1;
  for (int i = 0; i < m->n_mods; ++i) {
    if (a->inputs[i]) ++pl->n_present;
    if (m->dims[i] > pl->d_max) pl->d_max = m->dims[i];
  }
  if (pl->n_present == 0) return fail(BFVI_ERR_ARG, "at least one modality must be given");
  pl->k_max = a->flt_particles > a->smt_particles ? a->flt_particles : a->smt_particles;
  pl->tb = (size_t)a->T * a->B;
  pl->tbz = pl->tb * m->z_dim;
  const size_t rows = (size_t)a->B * pl->k_max;
  size_t cur = 0;
  auto carve = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
  pl->off_obs_mean = carve(sizeof(float) * pl->tbz * pl->n_present);
  pl->off_obs_std = carve(sizeof(float) * pl->tbz * pl->n_present);
  pl->off_obs_mask = carve(pl->tb * pl->n_present);
  pl->off_x0 = carve(sizeof(float) * pl->tb * pl->d_max);
  pl->off_h = carve(sizeof(float) * pl->tb * m->h_dim);
  for (int i = 0; i < 4; ++i) pl->off_flt[i] = carve(sizeof(float) * pl->tbz);
  pl->off_samples = carve(sizeof(float) * pl->tbz);
  pl->off_zrows = carve(sizeof(float) * rows * m->z_dim);
  // BFVI_PREC_FUSED: the transitions run in the fused on-chip kernels (hidden activations never reach HBM); shapes
  // they do not serve fall back to the 3xTF32 launch sequence (same numerics class)
  pl->fused = (a->precision == BFVI_PREC_FUSED && fused_supported(m->z_dim, m->h_dim)) ? 1 : 0;
  pl->off_hrows = carve(pl->fused ? 256 : sizeof(float) * rows * m->h_dim);
  pl->off_hrows2 = carve(pl->fused ? 256 : sizeof(float) * rows * m->h_dim);
  pl->off_g = carve(sizeof(float) * rows * m->z_dim);
  pl->off_nl = carve(sizeof(float) * rows * m->z_dim);
  pl->off_lin = carve(sizeof(float) * rows * m->z_dim);
  pl->off_as = carve(sizeof(float) * rows * m->z_dim);
  pl->off_packs = 0;
  if (pl->fused) {
    cur = align_up(cur, 1024);
    pl->off_packs = carve(fused_pack_total(m->h_dim));
  }
  pl->total = cur;
  return BFVI_OK;
}

int bfvi_forward_workspace(const bfvi_model* m, const bfvi_forward_args* a, size_t* bytes) {
  ForwardPlan pl;
  if (int rc = plan_forward(m, a, &pl)) return rc;
  if (!bytes) return fail(BFVI_ERR_ARG, "bytes null");
  *bytes = pl.total;
  return BFVI_OK;
}

int bfvi_forward(const bfvi_model* m, const float* params, const bfvi_forward_args* a, void* workspace,
                 size_t workspace_bytes, void* stream) {
  ForwardPlan pl;
  if (int rc = plan_forward(m, a, &pl)) return rc;
  if (!params || !workspace) return fail(BFVI_ERR_ARG, "null argument");
  if (!a->infer_mean || !a->infer_std || !a->prior_mean || !a->prior_std) return fail(BFVI_ERR_ARG, "outputs null");
  if (workspace_bytes < pl.total) return fail(BFVI_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, pl.total);
  if (((uintptr_t)workspace & 255) != 0) return fail(BFVI_ERR_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  bfvi_layout lay;
  bfvi_param_layout(m, &lay);
  const int M = m->n_mods, Z = m->z_dim, H = m->h_dim, T = a->T, B = a->B;
  const int prec = a->precision == BFVI_PREC_TF32 ? bfvi::tc::PREC_TF32 : bfvi::tc::PREC_TF32X3;
  const int64_t tb = (int64_t)pl.tb;
  float* obs_mean = (float*)(ws + pl.off_obs_mean);
  float* obs_std = (float*)(ws + pl.off_obs_std);
  uint8_t* obs_mask = (uint8_t*)(ws + pl.off_obs_mask);
  float* x0 = (float*)(ws + pl.off_x0);
  float* hbuf = (float*)(ws + pl.off_h);
  float* samples = (float*)(ws + pl.off_samples);
  float* zrows = (float*)(ws + pl.off_zrows);
  float* hrows = (float*)(ws + pl.off_hrows);
  float* hrows2 = (float*)(ws + pl.off_hrows2);
  LinearGroup lg;
  float* gbuf = (float*)(ws + pl.off_g);
  float* nlbuf = (float*)(ws + pl.off_nl);
  float* linbuf = (float*)(ws + pl.off_lin);
  float* asbuf = (float*)(ws + pl.off_as);
  auto ew_grid = [&](int64_t n, int per) { return dim3((unsigned)grid_for(n, per, 16)); };

  // ---- encode every given modality (models/dmm.py:160-173) ------------------------------
  int slot = 0;
  for (int i = 0; i < M; ++i) {
    if (!a->inputs[i]) continue;
    const bfvi_mlp_layout& l = lay.enc[i];
    const int D = m->dims[i];
    float* mean = obs_mean + (size_t)slot * pl.tbz;
    float* std = obs_std + (size_t)slot * pl.tbz;
    auto kp = bfvi::gen::prep_rows_kernel;
    BFVI_LAUNCH(kp, ew_grid(tb, 256), dim3(256), 0, st, a->inputs[i], tb, D, x0, obs_mask + (size_t)slot * pl.tb);
    if (int rc = linear_tc(x0, D, params + l.in_to_h_w, D, params + l.in_to_h_b, hbuf, H, tb, D, H, 1, prec, st)) return rc;
    lg.add(hbuf, H, params + l.mean_w, H, params + l.mean_b, mean, Z, tb, H, Z, 0);
    lg.add(hbuf, H, params + l.std_w, H, params + l.std_b, std, Z, tb, H, Z, 0);
    if (int rc = lg.run(prec, st)) return rc;
    auto ks = bfvi::gen::softplus_kernel;
    BFVI_LAUNCH(ks, ew_grid(tb * Z, 256), dim3(256), 0, st, std, tb * Z, bfvi::kMlpMinStd);
    ++slot;
  }
  BFVI_CHECK_CUDA();

  // ---- one filtering pass: T steps, each = 6 GEMMs over (B*K) rows + one fused kernel ----
#ifndef BFVI_EMU
  FusedBufs fb;
  memset(&fb, 0, sizeof(fb));
  bool packed[2] = {false, false};
  if (pl.fused) fused_carve_packs(ws + pl.off_packs, H, &fb);
#endif
  auto run_pass = [&](bfvi_filter_args& f) -> int {
    const int dir = f.direction == BFVI_DIR_BWD ? 1 : 0;
    const bfvi_gtf_layout& g = lay.trans[dir];
    const int64_t rows = (int64_t)B * f.n_particles;
#ifndef BFVI_EMU
    if (pl.fused && !packed[dir]) {              // weights as shared-memory images, once per direction and call
      if (int rc = fused_pack(g, params, dir, H, fb, st)) return rc;
      packed[dir] = true;
    }
#endif
    for (int i = 0; i < T; ++i) {
#ifndef BFVI_EMU
      if (i > 0 && pl.fused) {                     // ONE launch per transition
        if (int rc = fused_fwd(fb, dir, H, zrows, rows, gbuf, nlbuf, linbuf, asbuf, false, st)) return rc;
      } else
#endif
      if (i > 0) {
        // three grouped, width-homogeneous launches: the two z -> hidden layers; the hidden -> head layers and the
        // linear branch; the std head
        lg.add(zrows, Z, params + g.gate0_w, Z, params + g.gate0_b, hrows, H, rows, Z, H, 1);
        lg.add(zrows, Z, params + g.nonlin0_w, Z, params + g.nonlin0_b, hrows2, H, rows, Z, H, 1);
        if (int rc = lg.run(prec, st)) return rc;
        lg.add(hrows, H, params + g.gate2_w, H, params + g.gate2_b, gbuf, Z, rows, H, Z, 0);
        lg.add(hrows2, H, params + g.nonlin2_w, H, params + g.nonlin2_b, nlbuf, Z, rows, H, Z, 0);
        lg.add(zrows, Z, params + g.lin_w, Z, params + g.lin_b, linbuf, Z, rows, Z, Z, 0);   // Z wide: rides with level 2
        if (int rc = lg.run(prec, st)) return rc;
        if (int rc = linear_tc(nlbuf, Z, params + g.std_w, Z, params + g.std_b, asbuf, Z, rows, Z, Z, 0, prec, st)) return rc;
      }
      bfvi::gen::StepParams sp;
      memset(&sp, 0, sizeof(sp));
      sp.a = f;
      sp.z0_mean = params + lay.z0_mean; sp.z0_log_std = params + lay.z0_log_std;
      sp.min_std = m->min_std; sp.Z = Z; sp.i = i; sp.R = rows;
      sp.g = gbuf; sp.nl = nlbuf; sp.lin = linbuf; sp.as = asbuf;
      sp.zrows = zrows;
      sp.swz = pl.fused ? 1 : 0;                   // the fused kernels write their heads in the swz64 layout
      if (step4_ok(sp)) { auto k = bfvi::gen::step4_kernel; BFVI_LAUNCH(k, ew_grid((int64_t)B * (Z / 4), 128), dim3(128), 0, st, sp); }
      else { auto k = bfvi::gen::step_kernel; BFVI_LAUNCH(k, ew_grid((int64_t)B * Z, 128), dim3(128), 0, st, sp); }
    }
    BFVI_CHECK_CUDA();
    return BFVI_OK;
  };
  auto obs_experts = [&](bfvi_filter_args& f) {
    memset(&f, 0, sizeof(f));
    f.T = T; f.B = B; f.S = 1;
    for (int e = 0; e < pl.n_present; ++e) {
      bfvi_expert& ex = f.experts[e];
      ex.mean = obs_mean + (size_t)e * pl.tbz; ex.std = obs_std + (size_t)e * pl.tbz;
      ex.mask = obs_mask + (size_t)e * pl.tb;
      ex.stride_t = (int64_t)B * Z; ex.stride_b = Z; ex.mstride_t = B; ex.mstride_b = 1;
      ex.kind = BFVI_EXPERT_TENSOR;
    }
    f.n_experts = pl.n_present;
    f.sample = a->sample;
    f.noise.seed = a->seed; f.noise.b_offset = a->b_offset;
  };
  const bool smooth = a->mode == BFVI_MODE_FSMOOTH || a->mode == BFVI_MODE_BSMOOTH;
  // filtering pass (models/dmm.py:462-470): forward for ffilter / bsmooth, backward otherwise
  bfvi_filter_args f1;
  obs_experts(f1);
  f1.set_expert_bits[0] = (1u << pl.n_present) - 1u;
  f1.direction = (a->mode == BFVI_MODE_FFILTER || a->mode == BFVI_MODE_BSMOOTH) ? BFVI_DIR_FWD : BFVI_DIR_BWD;
  f1.n_particles = a->flt_particles;
  f1.sample_init = smooth ? 0 : a->sample_init;
  f1.noise.eps = a->eps_flt; f1.noise.stream_id = 7;
  if (smooth) {
    f1.infer_mean = (float*)(ws + pl.off_flt[0]); f1.infer_std = (float*)(ws + pl.off_flt[1]);
    f1.prior_mean = (float*)(ws + pl.off_flt[2]); f1.prior_std = (float*)(ws + pl.off_flt[3]);
  } else {
    f1.infer_mean = a->infer_mean; f1.infer_std = a->infer_std;
    f1.prior_mean = a->prior_mean; f1.prior_std = a->prior_std;
  }
  f1.samples = samples;
  if (int rc = run_pass(f1)) return rc;
  if (smooth) {                                   // smoothing pass (models/dmm.py:473-489)
    bfvi_filter_args f2;
    obs_experts(f2);
    const int P = pl.n_present;
    bfvi_expert& fe = f2.experts[P];
    fe.mean = f1.prior_mean; fe.std = f1.prior_std; fe.mask = nullptr;
    fe.stride_t = (int64_t)B * Z; fe.stride_b = Z;
    fe.kind = BFVI_EXPERT_TENSOR; fe.zero_mask_last_t = 1;          // flt_mask[-1] = 0
    f2.experts[P + 1].kind = BFVI_EXPERT_INV_PRIOR;
    f2.n_experts = P + 2;
    f2.set_expert_bits[0] = (1u << (P + 2)) - 1u;
    f2.direction = a->mode == BFVI_MODE_FSMOOTH ? BFVI_DIR_FWD : BFVI_DIR_BWD;
    f2.n_particles = a->smt_particles;
    f2.sample_init = a->sample_init;
    f2.noise.eps = a->eps_smt; f2.noise.stream_id = 8;
    f2.infer_mean = a->infer_mean; f2.infer_std = a->infer_std;
    f2.prior_mean = a->prior_mean; f2.prior_std = a->prior_std;
    f2.samples = samples;
    if (int rc = run_pass(f2)) return rc;
  }

  // ---- decode every modality from the pass's samples (models/dmm.py:192-212) -------------
  for (int i = 0; i < M; ++i) {
    if (!a->recon_mean[i] || !a->recon_std[i]) continue;
    const bfvi_mlp_layout& l = lay.dec[i];
    const int D = m->dims[i];
    if (int rc = linear_tc(samples, Z, params + l.in_to_h_w, Z, params + l.in_to_h_b, hbuf, H, tb, Z, H, 1, prec, st)) return rc;
    lg.add(hbuf, H, params + l.mean_w, H, params + l.mean_b, a->recon_mean[i], D, tb, H, D, 0);
    lg.add(hbuf, H, params + l.std_w, H, params + l.std_b, a->recon_std[i], D, tb, H, D, 0);
    if (int rc = lg.run(prec, st)) return rc;
    auto ks = bfvi::gen::softplus_kernel;
    BFVI_LAUNCH(ks, ew_grid(tb * D, 256), dim3(256), 0, st, a->recon_std[i], tb * D, bfvi::kMlpMinStd);
  }
  BFVI_CHECK_CUDA();
  return BFVI_OK;
}

}  // extern "C"
