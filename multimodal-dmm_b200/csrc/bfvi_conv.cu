// bfvi_conv.cu — C entries of the image-module kernels (bfvi_conv.cuh): Conv2d / ConvTranspose2d passes, BatchNorm2d
// (+ ReLU) forward / backward, the sigmoid backward and per-channel bias gradients.  Second translation unit of
// libbfvi_b200.so; errors go through the library's bfvi_last_error().
#include <cstdlib>
#include <cstring>

#include "../../include/bfvi.h"
#include "bfvi_conv.cuh"
#include "bfvi_internal.h"

namespace {

using bfvi::conv::Geom;
using bfvi::report_error;

#define BFVI_CONV_CHECK_CUDA()                                                                       \
  do {                                                                                               \
    cudaError_t e_ = cudaGetLastError();                                                             \
    if (e_ != cudaSuccess) return report_error(BFVI_ERR_CUDA, "CUDA error: %s (%s:%d)",              \
                                               cudaGetErrorString(e_), __FILE__, __LINE__);          \
  } while (0)

// BFVI_DETERMINISTIC=1: no cross-CTA float atomics (one pixel split per weight-gradient tile, no K slices in the dense
// GEMMs): run-to-run bit-identical results at the price of parallelism on the small problems.  Read per call.
bool deterministic() {
  const char* e = getenv("BFVI_DETERMINISTIC");
  return e != nullptr && atoi(e) != 0;
}

int check_geom(const bfvi_conv_geom* g, Geom* out) {
  if (g == nullptr) return report_error(BFVI_ERR_ARG, "geometry is null");
  if (g->n < 1 || g->c_small < 1 || g->h_small < 1 || g->w_small < 1 || g->c_big < 1 || g->h_big < 1 || g->w_big < 1)
    return report_error(BFVI_ERR_ARG, "empty convolution geometry");
  if (g->kernel < 1 || g->stride < 1 || g->padding < 0) return report_error(BFVI_ERR_ARG, "bad kernel / stride / padding");
  if (g->kernel > 7) return report_error(BFVI_ERR_UNSUPPORTED, "kernel sizes up to 7");
  // conv: h_small = floor((h_big + 2p - k) / s) + 1; deconv (output_padding 0): h_big = (h_small - 1) s - 2p + k
  const int eh = (g->h_small - 1) * g->stride + g->kernel, ew = (g->w_small - 1) * g->stride + g->kernel;
  if (eh > g->h_big + 2 * g->padding || g->h_big + 2 * g->padding >= eh + g->stride ||
      ew > g->w_big + 2 * g->padding || g->w_big + 2 * g->padding >= ew + g->stride)
    return report_error(BFVI_ERR_ARG, "map sizes %dx%d / %dx%d do not match kernel %d stride %d padding %d", g->h_small,
                        g->w_small, g->h_big, g->w_big, g->kernel, g->stride, g->padding);
  if ((long long)g->n * g->c_big * g->h_big * g->w_big > (1ll << 40)) return report_error(BFVI_ERR_ARG, "tensor too large");
  out->N = g->n;
  out->Cs = g->c_small; out->Hs = g->h_small; out->Ws = g->w_small;
  out->Cb = g->c_big; out->Hb = g->h_big; out->Wb = g->w_big;
  out->k = g->kernel; out->s = g->stride; out->p = g->padding;
  return BFVI_OK;
}

int chunk_for(int channels, int kk, int ct) {
  int c = bfvi::conv::kMaxSmemFloats / (kk * ct);
  if (c < 1) c = 1;
  return c < channels ? c : channels;
}

template <int KT, int CT>
void launch_gather(const Geom& g, const float* big, const float* w, const float* bias, float* small, int act,
                   cudaStream_t st) {
  const int chunk = chunk_for(g.Cb, g.k * g.k, CT);
  const long long total = (long long)g.N * g.Hs * g.Ws;
  const dim3 grid((unsigned)((total + bfvi::conv::kThreads - 1) / bfvi::conv::kThreads), (unsigned)((g.Cs + CT - 1) / CT));
  auto k = bfvi::conv::conv_gather_kernel<KT, CT>;
  BFVI_LAUNCH(k, grid, dim3(bfvi::conv::kThreads), sizeof(float) * (size_t)chunk * g.k * g.k * CT, st, g, big, w, bias,
              small, chunk, act);
}

template <int KT, int ST, int CT>
void launch_scatter(const Geom& g, const float* small, const float* w, const float* bias, float* big, int act,
                    cudaStream_t st) {
  const int chunk = chunk_for(g.Cs, g.k * g.k, CT);
  const long long total = (long long)g.N * g.Hb * g.Wb;
  const dim3 grid((unsigned)((total + bfvi::conv::kThreads - 1) / bfvi::conv::kThreads), (unsigned)((g.Cb + CT - 1) / CT));
  auto k = bfvi::conv::conv_scatter_kernel<KT, ST, CT>;
  BFVI_LAUNCH(k, grid, dim3(bfvi::conv::kThreads), sizeof(float) * (size_t)chunk * g.k * g.k * CT, st, g, small, w, bias,
              big, chunk, act);
}

template <int KT, int CSR>
void launch_wgrad(const Geom& g, const float* small, const float* big, float* dw, cudaStream_t st) {
  const long long total = (long long)g.N * g.Hs * g.Ws;
  const long long base = (long long)g.Cb * ((g.Cs + CSR - 1) / CSR);
  long long splits = (1184 + base - 1) / base;                      // ~ 8 CTAs per SM of 148
  const long long max_splits = (total + 2047) / 2048;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  if (deterministic()) splits = 1;
  const long long per = (total + splits - 1) / splits;
  const dim3 grid((unsigned)g.Cb, (unsigned)((g.Cs + CSR - 1) / CSR), (unsigned)splits);
  auto k = bfvi::conv::conv_wgrad_kernel<KT, CSR>;
  BFVI_LAUNCH(k, grid, dim3(bfvi::conv::kWgradThreads), 0, st, g, small, big, dw, per);
}

struct Scratch {
  double* partial;       // [C][kMaxSplit][2]
  float* coef;           // [C][2]
};
size_t scratch_bytes(int channels) {
  return (size_t)channels * bfvi::conv::kMaxSplit * 2 * sizeof(double) + (size_t)channels * 2 * sizeof(float) + 64;
}
int carve(void* scratch, size_t bytes, int channels, Scratch* out) {
  if (scratch == nullptr || bytes < scratch_bytes(channels))
    return report_error(BFVI_ERR_ARG, "scratch too small: %zu < %zu bytes", bytes, scratch_bytes(channels));
  if ((uintptr_t)scratch % 8 != 0) return report_error(BFVI_ERR_ARG, "scratch must be 8-byte aligned");
  out->partial = (double*)scratch;
  out->coef = (float*)(out->partial + (size_t)channels * bfvi::conv::kMaxSplit * 2);
  return BFVI_OK;
}

// launches chan_reduce over (N, C, HW); returns the split count
int launch_reduce(bfvi::conv::ReduceParams p, cudaStream_t st) {
  const long long total = (long long)p.N * p.HW;
  long long nsplit = total / 4096;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > bfvi::conv::kMaxSplit) nsplit = bfvi::conv::kMaxSplit;
  p.per = (total + nsplit - 1) / nsplit;
  auto k = bfvi::conv::chan_reduce_kernel;
  BFVI_LAUNCH(k, dim3((unsigned)p.C, (unsigned)nsplit), dim3(256), 0, st, p);
  return (int)nsplit;
}

dim3 plane_grid(int N, int C, long long HW) {
  long long chunks = (HW + 1023) / 1024;
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  return dim3((unsigned)((long long)N * C), (unsigned)chunks);
}

int check_nchw(int N, int C, long long HW) {
  if (N < 1 || C < 1 || HW < 1) return report_error(BFVI_ERR_ARG, "empty tensor");
  if ((long long)N * C > 2147483647ll) return report_error(BFVI_ERR_UNSUPPORTED, "N * C up to 2^31 - 1");
  return BFVI_OK;
}

}  // namespace

extern "C" {

int bfvi_conv_gather(const bfvi_conv_geom* geom, const float* big, const float* w, const float* bias, float* small,
                     int32_t act, void* stream) {
  Geom g;
  if (int rc = check_geom(geom, &g)) return rc;
  if (!big || !w || !small) return report_error(BFVI_ERR_ARG, "null tensor");
  if (act != BFVI_ACT_NONE && act != BFVI_ACT_SIGMOID) return report_error(BFVI_ERR_ARG, "act must be none or sigmoid");
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = g.Cs > 4;
  if (g.k == 3) { if (wide) launch_gather<3, 16>(g, big, w, bias, small, act, st); else launch_gather<3, 4>(g, big, w, bias, small, act, st); }
  else if (g.k == 4) { if (wide) launch_gather<4, 16>(g, big, w, bias, small, act, st); else launch_gather<4, 4>(g, big, w, bias, small, act, st); }
  else { if (wide) launch_gather<0, 16>(g, big, w, bias, small, act, st); else launch_gather<0, 4>(g, big, w, bias, small, act, st); }
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_conv_scatter(const bfvi_conv_geom* geom, const float* small, const float* w, const float* bias, float* big,
                      int32_t act, void* stream) {
  Geom g;
  if (int rc = check_geom(geom, &g)) return rc;
  if (!big || !w || !small) return report_error(BFVI_ERR_ARG, "null tensor");
  if (act != BFVI_ACT_NONE && act != BFVI_ACT_SIGMOID) return report_error(BFVI_ERR_ARG, "act must be none or sigmoid");
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = g.Cb > 4;
  if (g.k == 3 && g.s == 2) { if (wide) launch_scatter<3, 2, 16>(g, small, w, bias, big, act, st); else launch_scatter<3, 2, 4>(g, small, w, bias, big, act, st); }
  else if (g.k == 4 && g.s == 2) { if (wide) launch_scatter<4, 2, 16>(g, small, w, bias, big, act, st); else launch_scatter<4, 2, 4>(g, small, w, bias, big, act, st); }
  else { if (wide) launch_scatter<0, 0, 16>(g, small, w, bias, big, act, st); else launch_scatter<0, 0, 4>(g, small, w, bias, big, act, st); }
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_conv_wgrad(const bfvi_conv_geom* geom, const float* small, const float* big, float* dw, void* stream) {
  Geom g;
  if (int rc = check_geom(geom, &g)) return rc;
  if (!big || !dw || !small) return report_error(BFVI_ERR_ARG, "null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (g.k == 3) launch_wgrad<3, 4>(g, small, big, dw, st);
  else if (g.k == 4) launch_wgrad<4, 4>(g, small, big, dw, st);
  else launch_wgrad<0, 1>(g, small, big, dw, st);
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

size_t bfvi_chan_scratch(int32_t channels) { return channels < 1 ? 0 : scratch_bytes(channels); }

int bfvi_chan_bias_grad(const float* dy, int32_t N, int32_t C, int64_t HW, float* db, void* scratch, size_t bytes,
                        void* stream) {
  if (int rc = check_nchw(N, C, HW)) return rc;
  if (!dy || !db) return report_error(BFVI_ERR_ARG, "null tensor");
  Scratch sc;
  if (int rc = carve(scratch, bytes, C, &sc)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  bfvi::conv::ReduceParams p;
  memset(&p, 0, sizeof(p));
  p.a = dy; p.N = N; p.C = C; p.HW = HW; p.mode = 0; p.partial = sc.partial;
  const int nsplit = launch_reduce(p, st);
  auto k = bfvi::conv::bias_grad_finish_kernel;
  BFVI_LAUNCH(k, dim3((unsigned)((C + 127) / 128)), dim3(128), 0, st, (const double*)sc.partial, nsplit, (int)C, db);
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_bn2d_fwd(const float* x, int32_t N, int32_t C, int64_t HW, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, int32_t training, float momentum, float eps, int32_t relu,
                  float* y, float* mean_rstd, void* scratch, size_t bytes, void* stream) {
  if (int rc = check_nchw(N, C, HW)) return rc;
  if (!x || !y || !mean_rstd) return report_error(BFVI_ERR_ARG, "null tensor");
  if (!training && (!running_mean || !running_var))
    return report_error(BFVI_ERR_ARG, "evaluation mode needs the running statistics");
  if (training && (long long)N * HW < 2) return report_error(BFVI_ERR_ARG, "training statistics need more than one value per channel");
  Scratch sc;
  if (int rc = carve(scratch, bytes, C, &sc)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int nsplit = 0;
  if (training) {
    bfvi::conv::ReduceParams p;
    memset(&p, 0, sizeof(p));
    p.a = x; p.N = N; p.C = C; p.HW = HW; p.mode = 0; p.partial = sc.partial;
    nsplit = launch_reduce(p, st);
  }
  auto kf = bfvi::conv::bn_stats_finish_kernel;
  BFVI_LAUNCH(kf, dim3((unsigned)((C + 127) / 128)), dim3(128), 0, st, (const double*)sc.partial, nsplit, (int)C,
              (double)N * (double)HW, eps, momentum, (int)(training != 0), running_mean, running_var, mean_rstd);
  auto ka = bfvi::conv::bn_apply_kernel;
  BFVI_LAUNCH(ka, plane_grid(N, C, HW), dim3(256), 0, st, x, (const float*)mean_rstd, gamma, beta, (int)C, (long long)HW,
              (int)(relu != 0), y);
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_bn2d_bwd(const float* dy, const float* x, const float* y, const float* mean_rstd, const float* gamma, int32_t N,
                  int32_t C, int64_t HW, int32_t training, int32_t relu, float* dx, float* d_gamma, float* d_beta,
                  void* scratch, size_t bytes, void* stream) {
  if (int rc = check_nchw(N, C, HW)) return rc;
  if (!dy || !x || !mean_rstd || !dx || (relu && !y)) return report_error(BFVI_ERR_ARG, "null tensor");
  Scratch sc;
  if (int rc = carve(scratch, bytes, C, &sc)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  bfvi::conv::ReduceParams p;
  memset(&p, 0, sizeof(p));
  p.a = x; p.b = dy; p.y = y; p.mean_rstd = mean_rstd; p.N = N; p.C = C; p.HW = HW; p.mode = 1; p.relu = relu != 0;
  p.partial = sc.partial;
  const int nsplit = launch_reduce(p, st);
  auto kf = bfvi::conv::bn_bwd_finish_kernel;
  BFVI_LAUNCH(kf, dim3((unsigned)((C + 127) / 128)), dim3(128), 0, st, (const double*)sc.partial, nsplit, (int)C,
              (double)N * (double)HW, d_gamma, d_beta, sc.coef);
  auto ka = bfvi::conv::bn_bwd_apply_kernel;
  BFVI_LAUNCH(ka, plane_grid(N, C, HW), dim3(256), 0, st, dy, x, y, mean_rstd, gamma, (const float*)sc.coef, (int)C,
              (long long)HW, (int)(relu != 0), (int)(training != 0), dx);
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_sigmoid_bwd(const float* p, const float* dp, int64_t n, float* dx, void* stream) {
  if (!p || !dp || !dx || n < 1) return report_error(BFVI_ERR_ARG, "null/empty argument");
  long long blocks = (n + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  auto k = bfvi::conv::sigmoid_bwd_kernel;
  BFVI_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, p, dp, (long long)n, dx);
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

}  // extern "C"

namespace {
void launch_dense(bfvi::conv::DenseParams p, cudaStream_t st) {
  dim3 grid((unsigned)((p.N + bfvi::conv::kDenseTile - 1) / bfvi::conv::kDenseTile),
            (unsigned)((p.M + bfvi::conv::kDenseTile - 1) / bfvi::conv::kDenseTile));
  // few output tiles and a long contraction (feat_to_z at 625 frames: 40 tiles, K = 4096): slices of K on the other SMs
  const long long tiles = (long long)grid.x * grid.y;
  if (!p.relu && tiles * 2 <= 148 && p.K >= 1024 && !deterministic()) {
    long long splits = 296 / tiles;
    if (splits > p.K / 256) splits = p.K / 256;
    if (splits > 1) {
      p.k_per_split = (int)(((p.K + splits - 1) / splits + bfvi::conv::kDenseK - 1) / bfvi::conv::kDenseK * bfvi::conv::kDenseK);
      grid.z = (unsigned)((p.K + p.k_per_split - 1) / p.k_per_split);
      if (!p.accumulate) cudaMemsetAsync(p.C, 0, sizeof(float) * (size_t)p.M * (size_t)p.ldc, st);
    }
  }
  auto k = bfvi::conv::dense_gemm_kernel;
  BFVI_LAUNCH(k, grid, dim3(256), 0, st, p);
}
int check_dense(int64_t rows, int32_t n_in, int32_t n_out) {
  if (rows < 1 || n_in < 1 || n_out < 1) return report_error(BFVI_ERR_ARG, "empty dense layer");
  if (rows > 64ll * 65535) return report_error(BFVI_ERR_UNSUPPORTED, "up to 4 193 240 rows per call");
  return BFVI_OK;
}
}  // namespace

extern "C" {

int bfvi_dense_fwd(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t n_in, int32_t n_out,
                   int32_t relu, void* stream) {
  if (int rc = check_dense(rows, n_in, n_out)) return rc;
  if (!x || !w || !y) return report_error(BFVI_ERR_ARG, "null tensor");
  bfvi::conv::DenseParams p;
  memset(&p, 0, sizeof(p));
  p.A = x; p.sai = n_in; p.sal = 1; p.a_l_contig = 1;
  p.B = w; p.sbj = n_in; p.sbl = 1; p.b_l_contig = 1;
  p.C = y; p.ldc = n_out; p.M = (int)rows; p.N = n_out; p.K = n_in; p.bias = bias; p.relu = relu != 0;
  launch_dense(p, (cudaStream_t)stream);
  BFVI_CONV_CHECK_CUDA();
  return BFVI_OK;
}

int bfvi_dense_bwd(const float* x, const float* w, const float* y, const float* dy, float* dy_masked, int64_t rows,
                   int32_t n_in, int32_t n_out, int32_t relu, float* dx, float* dw, float* db, void* scratch, size_t bytes,
                   void* stream) {
  if (int rc = check_dense(rows, n_in, n_out)) return rc;
  if (!x || !w || !dy || !dw) return report_error(BFVI_ERR_ARG, "null tensor");
  if (relu && (!y || !dy_masked)) return report_error(BFVI_ERR_ARG, "the ReLU backward needs y and the dy_masked buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (relu) {
    const long long n = (long long)rows * n_out;
    long long blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    auto k = bfvi::conv::relu_mask_kernel;
    BFVI_LAUNCH(k, dim3((unsigned)blocks), dim3(256), 0, st, dy, y, n, dy_masked);
    dy = dy_masked;
  }
  bfvi::conv::DenseParams p;
  if (dx != nullptr) {                       // dx (rows, n_in) = dy w
    memset(&p, 0, sizeof(p));
    p.A = dy; p.sai = n_out; p.sal = 1; p.a_l_contig = 1;
    p.B = w; p.sbj = 1; p.sbl = n_in; p.b_l_contig = 0;
    p.C = dx; p.ldc = n_in; p.M = (int)rows; p.N = n_in; p.K = n_out;
    launch_dense(p, st);
  }
  memset(&p, 0, sizeof(p));                  // dw (n_out, n_in) += dy^T x
  p.A = dy; p.sai = 1; p.sal = n_out; p.a_l_contig = 0;
  p.B = x; p.sbj = 1; p.sbl = n_in; p.b_l_contig = 0;
  p.C = dw; p.ldc = n_in; p.M = n_out; p.N = n_in; p.K = (int)rows; p.accumulate = 1;
  launch_dense(p, st);
  BFVI_CONV_CHECK_CUDA();
  if (db != nullptr) return bfvi_chan_bias_grad(dy, (int32_t)rows, n_out, 1, db, scratch, bytes, stream);
  return BFVI_OK;
}

}  // extern "C"
