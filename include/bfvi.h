/*
 * bfvi.h — C ABI of the B200-native BFVI training-step library (libbfvi_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of ztangent/multimodal-dmm: the
 * Backward-Forward Variational Inference ELBO forward+backward step of MultiDMM.
 * Every entry point names the reference interface it replaces (paths relative to
 * the reference repository root).  The Python host side
 * (multimodal-dmm_b200/models/*.py) binds these symbols with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types in any signature;
 *   - every data pointer is CALLER-ALLOCATED DEVICE memory (fp32 unless noted);
 *     the library never allocates or frees device memory and keeps no mutable
 *     global state, so calls on different streams may run concurrently;
 *   - every call is asynchronous on the `stream` argument (a cudaStream_t passed
 *     as void*); results are valid after that stream is synchronised;
 *   - return value: 0 = ok, <0 = error (BFVI_ERR_*); bfvi_last_error() returns a
 *     thread-local description;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails
 *     with BFVI_ERR_CUDA.
 *
 * Tensor layouts follow the reference: time first, (T, B, feature) fp32,
 * missing data = NaN (datasets/multiseq.py:340-353), sequence mask (T, B) u8.
 * Parameters live in ONE flat fp32 buffer whose layout bfvi_param_layout()
 * defines (each tensor is a PyTorch nn.Linear (out,in) row-major block, 16-byte
 * aligned) so the host can alias torch Parameters onto it and all-reduce the
 * matching flat gradient buffer with a single NCCL call.
 */
#ifndef BFVI_H_
#define BFVI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFVI_VERSION 140 /* 0.4.0: image modules (bfvi_conv_*, bfvi_bn2d_*, bfvi_sigmoid_bwd, bfvi_chan_bias_grad);
                            0.3.0: bfvi_mlp_* (encoder / decoder modules by pointer), fused transitions in bfvi_forward;
                            0.2.0: fused on-chip GTF kernels, training precision modes, batch tiles, bfvi_sizeof / bfvi_last_dispatch */

#define BFVI_MAX_MODS 16
#define BFVI_MAX_SETS (BFVI_MAX_MODS + 1)
#define BFVI_MAX_EXPERTS (BFVI_MAX_MODS + 2)

enum {
  BFVI_OK = 0,
  BFVI_ERR_ARG = -1,         /* invalid argument */
  BFVI_ERR_UNSUPPORTED = -2, /* shape / distribution not covered by a kernel */
  BFVI_ERR_CUDA = -3,        /* CUDA runtime error (or no device) */
  BFVI_ERR_WORKSPACE = -4    /* workspace too small */
};

enum { BFVI_DIST_NORMAL = 0, BFVI_DIST_BERNOULLI = 1, BFVI_DIST_CATEGORICAL = 2 };
enum { BFVI_DIR_FWD = 0, BFVI_DIR_BWD = 1 };
/* forward() modes, models/dmm.py:432-433 */
enum { BFVI_MODE_BFILTER = 0, BFVI_MODE_FFILTER = 1, BFVI_MODE_FSMOOTH = 2, BFVI_MODE_BSMOOTH = 3 };
/* expert kinds for bfvi_filter_* */
enum { BFVI_EXPERT_TENSOR = 0, BFVI_EXPERT_INV_PRIOR = 1 };

/* Static description of a MultiDMM (constructor args, models/dmm.py:29-32). */
typedef struct bfvi_model {
  int32_t n_mods;               /* M */
  int32_t z_dim;                /* latent dims */
  int32_t h_dim;                /* hidden width of every MLP / GTF */
  int32_t dims[BFVI_MAX_MODS];  /* flattened feature dim per modality */
  int32_t dists[BFVI_MAX_MODS]; /* BFVI_DIST_* */
  float min_std;                /* models/dmm.py:31 (GTF + global prior) */
} bfvi_model;

/* Offsets (in floats) of every parameter tensor inside the flat buffer.
 * Names are the reference state_dict keys. */
typedef struct bfvi_mlp_layout { /* common.GaussianMLP, models/common.py:25-41 */
  int64_t in_to_h_w, in_to_h_b;  /* in_to_h.0.{weight (H,in), bias (H)} */
  int64_t mean_w, mean_b;        /* h_to_mean.{weight (out,H), bias} */
  int64_t std_w, std_b;          /* h_to_std.0.{weight (out,H), bias} */
  int64_t begin, end;            /* [begin,end) covers the whole block */
} bfvi_mlp_layout;

typedef struct bfvi_gtf_layout { /* common.GaussianGTF, models/common.py:43-68 */
  int64_t gate0_w, gate0_b;      /* z_to_gate.0 (H,Z) */
  int64_t gate2_w, gate2_b;      /* z_to_gate.2 (Z,H) */
  int64_t lin_w, lin_b;          /* z_lin (Z,Z) */
  int64_t nonlin0_w, nonlin0_b;  /* z_nonlin.0 (H,Z) */
  int64_t nonlin2_w, nonlin2_b;  /* z_nonlin.2 (Z,H) */
  int64_t std_w, std_b;          /* z_to_std.0 (Z,Z) */
  int64_t begin, end;
} bfvi_gtf_layout;

typedef struct bfvi_layout {
  int64_t z0_mean, z0_log_std;            /* (1,Z) each, models/dmm.py:115-116 */
  bfvi_mlp_layout enc[BFVI_MAX_MODS];     /* enc.<m>  (Normal/Bernoulli MLP encoders) */
  bfvi_mlp_layout dec[BFVI_MAX_MODS];     /* dec.<m>  (Normal MLP decoders) */
  bfvi_gtf_layout trans[2];               /* [BFVI_DIR_FWD], [BFVI_DIR_BWD] */
  int64_t total;                          /* floats in the flat buffer */
} bfvi_layout;

/* One Gaussian expert stream fed to the filter (product_of_experts inputs,
 * models/dgts.py:15-51).  mean/std address = base + s*stride_s + t*stride_t +
 * b*stride_b + z; a stride of 0 broadcasts.  d_mean/d_std (nullable) receive
 * gradients with the same strides, accumulated with atomic adds. */
typedef struct bfvi_expert {
  const float* mean;
  const float* std;
  const uint8_t* mask;             /* nullable = all ones */
  int64_t stride_s, stride_t, stride_b;
  int64_t mstride_s, mstride_t, mstride_b;
  float* d_mean;
  float* d_std;
  int32_t kind;                    /* BFVI_EXPERT_* ; INV_PRIOR ignores the pointers */
  int32_t zero_mask_last_t;        /* flt_mask[-1] = 0, models/dmm.py:482 */
} bfvi_expert;

/* Reparameterisation noise source (replaces MultiDGTS._sample_gauss,
 * models/dgts.py:177-180).  Either an external tensor of N(0,1) draws laid out
 * (S, T, B, K, Z) and indexed by the time step at which the draw is consumed,
 * or (eps == NULL) a counter-based Philox4x32-10 stream keyed by
 * (seed, stream_id) and indexed by (s, t, b, k, z). */
typedef struct bfvi_noise {
  const float* eps;
  uint64_t seed;
  uint32_t stream_id;
  uint32_t b_offset;  /* global batch index of local b = 0 (data-parallel shards) */
  const uint64_t* seed_dev; /* nullable: device address the kernels read the Philox seed from at run time instead of
                               `seed` — a captured CUDA graph of the step is replayed with a fresh seed per step
                               (large-dim family only in this version) */
} bfvi_noise;

/* MultiDMM.z_filter (models/dmm.py:319-412) over S independent chain sets. */
typedef struct bfvi_filter_args {
  int32_t T, B, S;
  int32_t n_experts;
  bfvi_expert experts[BFVI_MAX_EXPERTS];
  uint32_t set_expert_bits[BFVI_MAX_SETS]; /* bit e set: chain set s uses expert e */
  int32_t direction;      /* BFVI_DIR_* */
  int32_t n_particles;    /* K */
  int32_t sample;         /* models/dmm.py:398 */
  int32_t sample_init;
  bfvi_noise noise;
  /* outputs, each (S, T, B, Z) */
  float* infer_mean; float* infer_std;
  float* prior_mean; float* prior_std;
  float* samples;         /* nullable */
  /* optional fused KL(infer || prior) term, losses.kld_gauss (models/losses.py:14-21):
   * loss_acc[0] += kl_weight * KL over (t,b) with seq_mask set */
  const uint8_t* seq_mask; /* (T,B) u8, nullable = ones */
  float kl_weight;
  double* loss_acc;        /* nullable */
  /* backward inputs (nullable), each (S, T, B, Z) */
  const float* d_infer_mean; const float* d_infer_std;
  const float* d_prior_mean; const float* d_prior_std;
  const float* d_samples;
  /* scratch of the large-dim (tcgen05) family, bfvi_filter_workspace() bytes, 256-byte
   * aligned device memory; unused (may be NULL) for the small-dim family */
  void* workspace;
  size_t workspace_bytes;
} bfvi_filter_args;

/* Arguments of one MultiDMM.step (models/dmm.py:503-554) as driven by
 * Trainer.train (trainer.py:237-243). */
typedef struct bfvi_step_args {
  int32_t T, B;
  const float* inputs[BFVI_MAX_MODS];   /* (T,B,D_m), NaN = missing */
  const float* targets[BFVI_MAX_MODS];  /* (T,B,D_m) */
  const uint8_t* seq_mask;              /* (T,B) u8 from len_to_mask */
  float kld_mult;
  float rec_mults[BFVI_MAX_MODS];
  int32_t uni_loss;
  int32_t f_mode, s_mode;               /* BFVI_MODE_* */
  float f_mult, s_mult, match_mult;
  int32_t train_particles, match_particles;
  int32_t sample, sample_init;
  /* noise: external tensors (all NULL => Philox with `seed`) */
  const float* eps_match;               /* (2, K_match, Z) */
  const float* eps_filt;                /* (S, T, B, 1, Z)   f_mode pass */
  const float* eps_sflt;                /* (S, T, B, K, Z)   s_mode filtering pass */
  const float* eps_ssmt;                /* (S, T, B, 1, Z)   s_mode smoothing pass */
  uint64_t seed;
  uint32_t b_offset;                    /* see bfvi_noise */
  float match_count;                    /* mask.sum() used by the prior-matching term
                                           (models/dmm.py:541); < 0 = count seq_mask here */
  const uint64_t* seed_dev;             /* nullable, see bfvi_noise.seed_dev (large-dim family) */
  int32_t precision;                    /* large-dim family: BFVI_PREC_TF32X3 (0) error-compensated 3xTF32 through the
                                           launch-sequence path (FP32-class everywhere); BFVI_PREC_TF32 (1) the same path with
                                           single-pass TF32 operands; BFVI_PREC_FUSED (2) the FUSED on-chip transition kernels
                                           where the shape is served (z_dim 64, h_dim a multiple of 128; otherwise mode 0):
                                           hidden activations never leave the SM, forward contractions are error-compensated
                                           FP16 hi/lo products (FP32-class: the ReLU signs match the reference's), the input
                                           gradient contracts TF32 operands and the H-wide weight gradients FP16 operands,
                                           all with FP32 accumulation */
  int32_t batch_tile;                   /* large-dim family: sequences per batch tile — the step loops over tiles with
                                           the gradient accumulated, so the workspace is O(batch_tile) and B is unbounded
                                           (SURVEY 7 "B = 65 536 at 1 GPU is a loop over batch tiles inside one step");
                                           0 = chosen by the library (bfvi_step_workspace reports the matching size) */
} bfvi_step_args;
enum { BFVI_PREC_TF32X3 = 0, BFVI_PREC_TF32 = 1, BFVI_PREC_FUSED = 2 };

int bfvi_version(void);
/* 16 hex digits of the SHA-256 over the sources this binary was compiled from (csrc/ in name order, then this header),
 * stamped by multimodal-dmm_b200/build.py; "unknown" for a build that did not stamp it.  The binary is git-ignored but
 * ships to the GPU box: multimodal-dmm_b200/_lib.py::source_id() recomputes the digest from the tree and the host-API test
 * asserts they agree, so a stale library cannot pass for the current sources. */
const char* bfvi_build_id(void);
const char* bfvi_last_error(void);

/* sizeof() of the argument structs as THIS build of the library sees them: a binding (ctypes / cgo / JNI
 * stub) asserts its own struct sizes against these at load time, so a stale stub fails loudly instead of
 * handing the library a short struct (multimodal-dmm_b200/_lib.py does; INTEGRATION.md shows it). */
enum {
  BFVI_STRUCT_MODEL = 0, BFVI_STRUCT_LAYOUT, BFVI_STRUCT_EXPERT, BFVI_STRUCT_NOISE, BFVI_STRUCT_FILTER_ARGS,
  BFVI_STRUCT_STEP_ARGS, BFVI_STRUCT_FORWARD_ARGS, BFVI_STRUCT_CONV_GEOM
};
size_t bfvi_sizeof(int32_t which);

/* Which kernel variants the last bfvi_step_fwd_bwd / bfvi_step_profile / bfvi_filter_fwd / bfvi_filter_bwd /
 * bfvi_forward call of this host thread dispatched: a ';'-separated list of distinct entries such as
 * "chain_fwd<5,20,5> lanes=5", "chain_bwd<5,20,1> K=25 lanes=5 warps=4", "segmented:cooperative seg=2",
 * "step:chunks=2", "gtf_fused_fwd<tf32>".  Thread-local, valid until the next such call.  The parity tests
 * assert on it so that a test of a throughput mapping cannot silently run the latency mapping. */
const char* bfvi_last_dispatch(void);

/* Flat parameter layout for a model (all-default MLP encoders/decoders). */
int bfvi_param_layout(const bfvi_model* model, bfvi_layout* out);

/* Which kernel family serves this model: 1 = register-resident small-dim family (every entry
 * point), 2 = tcgen05 large-dim family (bfvi_step_fwd_bwd, bfvi_forward, bfvi_filter_fwd / _bwd for
 * all-default Gaussian models of any (z_dim, h_dim); the stand-alone bfvi_encode_* / bfvi_decode_*
 * kernels exist for family 1 only and return BFVI_ERR_UNSUPPORTED otherwise), 0 = invalid model. */
int bfvi_kernel_family(const bfvi_model* model);

/* MultiDMM.encode for one default Gaussian-MLP encoder (models/dmm.py:165-173 +
 * models/common.py:38-41): NaN -> mask, zero fill, MLP.  x (n_rows, D_m). */
int bfvi_encode_fwd(const bfvi_model* model, const float* params, int32_t mod,
                    const float* x, int64_t n_rows, float* mean, float* std,
                    uint8_t* mask, void* stream);
/* Backward of the above: accumulates into grads (flat, same layout). */
int bfvi_encode_bwd(const bfvi_model* model, const float* params, float* grads,
                    int32_t mod, const float* x, int64_t n_rows,
                    const float* d_mean, const float* d_std, void* stream);

/* MultiDMM.decode for one default Gaussian-MLP decoder (models/dmm.py:207-211). */
int bfvi_decode_fwd(const bfvi_model* model, const float* params, int32_t mod,
                    const float* z, int64_t n_rows, float* mean, float* std,
                    void* stream);
/* Backward of bfvi_decode_fwd: accumulates into grads and d_z (+=, nullable). */
int bfvi_decode_bwd(const bfvi_model* model, const float* params, float* grads,
                    int32_t mod, const float* z, int64_t n_rows, const float* d_mean,
                    const float* d_std, float* d_z, void* stream);
/* Decoder + losses.nll_gauss (models/losses.py:68-89) fused, forward and
 * backward in one pass: loss_acc[0] += weight * NLL; d_z (+=, nullable) and
 * grads (nullable) receive gradients scaled by weight.  row_mask (n_rows) u8 is
 * the sequence mask broadcast to rows (nullable). */
int bfvi_decode_nll(const bfvi_model* model, const float* params, float* grads,
                    int32_t mod, const float* z, const float* target,
                    const uint8_t* row_mask, int64_t n_rows, float weight,
                    double* loss_acc, float* d_z, void* stream);

/* z_filter forward / backward (models/dmm.py:319-412).  Backward needs the
 * forward outputs in `args` plus the upstream gradients; it accumulates the
 * transition / prior gradients into `grads` and expert gradients into
 * experts[e].d_mean / d_std. */
int bfvi_filter_workspace(const bfvi_model* model, const bfvi_filter_args* args, size_t* bytes);
int bfvi_filter_fwd(const bfvi_model* model, const float* params,
                    const bfvi_filter_args* args, void* stream);
int bfvi_filter_bwd(const bfvi_model* model, const float* params, float* grads,
                    const bfvi_filter_args* args, void* stream);

/* losses.kld_gauss (models/losses.py:14-21) with an optional (n_rows) u8 mask:
 * out[0] = 0.5 * sum(...).  Backward writes the four gradients scaled by g. */
int bfvi_kld_fwd(const float* mean1, const float* std1, const float* mean2,
                 const float* std2, const uint8_t* row_mask, int64_t n_rows,
                 int32_t z_dim, double* out, void* stream);
int bfvi_kld_bwd(const float* mean1, const float* std1, const float* mean2,
                 const float* std2, const uint8_t* row_mask, int64_t n_rows,
                 int32_t z_dim, float g, float* d_mean1, float* d_std1,
                 float* d_mean2, float* d_std2, void* stream);

/* losses.nll_gauss (models/losses.py:68-89): x may hold NaN; row_mask (n_rows). */
int bfvi_nll_gauss_fwd(const float* mean, const float* std, const float* x,
                       const uint8_t* row_mask, int64_t n_rows, int32_t d,
                       double* out, void* stream);
int bfvi_nll_gauss_bwd(const float* mean, const float* std, const float* x,
                       const uint8_t* row_mask, int64_t n_rows, int32_t d, float g,
                       float* d_mean, float* d_std, void* stream);

/* losses.nll_bernoulli (models/losses.py:23-42): binary cross-entropy summed over the
 * observed (x not NaN) elements of the rows row_mask keeps; theta, x: (n_rows, d).
 * Clamps as F.binary_cross_entropy: log terms >= -100, gradient denominator >= 1e-12. */
int bfvi_nll_bernoulli_fwd(const float* theta, const float* x, const uint8_t* row_mask,
                           int64_t n_rows, int32_t d, double* out, void* stream);
int bfvi_nll_bernoulli_bwd(const float* theta, const float* x, const uint8_t* row_mask,
                           int64_t n_rows, int32_t d, float g, float* d_theta, void* stream);

/* losses.nll_categorical (models/losses.py:44-66): -sum over observed rows of
 * probs[row, (long)x[row]] (the reference feeds probabilities, not log-probabilities,
 * to F.nll_loss; kept).  probs: (n_rows, n_cat); x: (n_rows) float labels, NaN = missing. */
int bfvi_nll_categorical_fwd(const float* probs, const float* x, const uint8_t* row_mask,
                             int64_t n_rows, int32_t n_cat, double* out, void* stream);
int bfvi_nll_categorical_bwd(const float* probs, const float* x, const uint8_t* row_mask,
                             int64_t n_rows, int32_t n_cat, float g, float* d_probs, void* stream);

/* ---- batch preparation either side of the step -------------------
 * Time-first (T, B, D...) fp32 batches, missing = NaN, as datasets/multiseq.py builds them. */

/* len_to_mask (datasets/multiseq.py:321-327): mask[t, b] = t < lengths[b], (T, B) uint8. */
int bfvi_len_to_mask(const int32_t* lengths, int32_t T, int32_t B, uint8_t* mask, void* stream);

/* pad_and_merge (datasets/multiseq.py:342-353): sequences packed back to back as rows of D
 * floats; row_start (B + 1, device) = first packed row of each sequence.  out (T, B, D),
 * NaN beyond each sequence's end. */
int bfvi_pad_merge(const float* packed, const int64_t* row_start, int32_t T, int32_t B,
                   int64_t D, float* out, void* stream);

/* seq_decoll (datasets/multiseq.py:388-398): de-pad and reorder a (T, B, D) batch.  Output sequence j is
 * batch column src[j], its first (row_start[j+1] - row_start[j]) steps, written at packed row row_start[j];
 * row_start has n_out + 1 entries, total_rows = row_start[n_out] (all device pointers). */
int bfvi_unpad(const float* x, const int64_t* row_start, const int32_t* src, int32_t B, int32_t n_out,
               int64_t total_rows, int64_t D, float* packed, void* stream);

/* Per-sequence MSE of the evaluation metrics (spirals.py:105-111): out[b] = sum over unmasked t of
 * sum_m sum_d (recon_m - target_m)^2, divided by lengths[b].  recon / target: HOST arrays of n_mods device
 * pointers to (T, B, dims[m]) tensors; mask (T, B) uint8, lengths (B) float, out (B): device.
 * n_split > 1 cuts the time axis into that many slices per sequence (more blocks for few, long sequences;
 * bfvi_seq_mse_splits proposes a count) and needs scratch = B * n_split floats; the two-stage sum is
 * deterministic. */
int bfvi_seq_mse_splits(int32_t T, int32_t B);
int bfvi_seq_mse(const float* const* recon, const float* const* target, const int64_t* dims, int32_t n_mods,
                 const uint8_t* mask, const float* lengths, int32_t T, int32_t B, float* out, float* scratch,
                 int32_t n_split, void* stream);

/* eval_ssim / _ssim (utils.py:110-212): per-image SSIM (and contrast-structure term, nullable) of two
 * (N, C, H, W) fp32 batches with a separable 1-D window `win` (HOST array of win_size taps, odd, <= 15;
 * the reference default is the 11-tap Gaussian of sigma 1.5), valid padding, mean over (C, H', W').
 * scratch: bfvi_ssim_scratch(...) bytes of device memory. */
size_t bfvi_ssim_scratch(int32_t N, int32_t C, int32_t H, int32_t W, int32_t win_size);
int bfvi_ssim(const float* x, const float* y, int32_t N, int32_t C, int32_t H, int32_t W, const float* win,
              int32_t win_size, float data_range, float* ssim, float* cs, float* scratch, void* stream);

/* func_delete (datasets/multiseq.py:405-420): out = copy of x with rows (t, b) set to NaN.
 * _rows: explicit (T, B) flags (any del_func; the host replays the reference's numpy draws).
 * _spans: rows lo[b] <= t < hi[b] (burst_delete :428-434, del_segment :443-448), or with
 *   invert = 1 every row of [0, lengths[b]) OUTSIDE the span (keep_segment :436-441);
 *   lengths nullable = T. */
int bfvi_delete_rows(const float* x, const uint8_t* del_mask, int32_t T, int32_t B, int64_t D,
                     float* out, void* stream);
int bfvi_delete_spans(const float* x, const int32_t* lo, const int32_t* hi,
                      const int32_t* lengths, int32_t invert, int32_t T, int32_t B, int64_t D,
                      float* out, void* stream);

/* Device-side draw of the deleted rows (seeded Philox stream instead of numpy's global
 * generator): BFVI_DELETE_UNIFORM = rand_delete (:422-426), exactly int(frac * length)
 * steps of each sequence, uniformly without replacement; BFVI_DELETE_BURST = burst_delete
 * (:428-434).  Writes (T, B) flags for bfvi_delete_rows.  stream_id separates modalities,
 * b_offset = global index of local sequence 0 (data parallel). */
enum { BFVI_DELETE_UNIFORM = 0, BFVI_DELETE_BURST = 1 };
int bfvi_draw_deletions(const int32_t* lengths, int32_t T, int32_t B, double frac, int32_t mode,
                        uint64_t seed, uint32_t stream_id, uint32_t b_offset, uint8_t* del_mask,
                        void* stream);

/* Whole MultiDMM.step + backward (models/dmm.py:503-554, trainer.py:237-243).
 * loss_out[0] (device fp32) = un-normalised summed loss; grads (nullable: forward
 * only) = d loss / d params, OVERWRITTEN.  `launches` (nullable, host) receives
 * the number of kernels launched. */
int bfvi_step_workspace(const bfvi_model* model, const bfvi_step_args* args,
                        size_t* bytes);
int bfvi_step_fwd_bwd(const bfvi_model* model, const float* params, float* grads,
                      const bfvi_step_args* args, void* workspace,
                      size_t workspace_bytes, float* loss_out, int32_t* launches,
                      void* stream);

/* Same as bfvi_step_fwd_bwd, but brackets every phase of the step with CUDA events
 * on `stream`, SYNCHRONISES, and writes the per-phase device time in milliseconds to
 * the host array phase_ms[BFVI_N_PHASES] (measurement aid for bench.py's roofline). */
enum {
  BFVI_PHASE_MATCH = 0, BFVI_PHASE_ENCODE_FWD, BFVI_PHASE_FILTER_F_FWD,
  BFVI_PHASE_FILTER_S_FLT_FWD, BFVI_PHASE_FILTER_S_SMT_FWD, BFVI_PHASE_DECODE_NLL,
  BFVI_PHASE_FILTER_S_SMT_BWD, BFVI_PHASE_FILTER_S_FLT_BWD, BFVI_PHASE_FILTER_F_BWD,
  BFVI_PHASE_ENCODE_BWD, BFVI_PHASE_FINALIZE, BFVI_N_PHASES
};
int bfvi_step_profile(const bfvi_model* model, const float* params, float* grads,
                      const bfvi_step_args* args, void* workspace, size_t workspace_bytes,
                      float* loss_out, float* phase_ms, void* stream);

/* MultiDMM.forward (models/dmm.py:420-494) for an all-default Gaussian model of ANY
 * (z_dim, h_dim): encode -> filtering pass -> (smoothing pass) -> decode, the inference /
 * evaluation path of Trainer.evaluate (trainer.py:294-296).  This is the entry point of
 * the large-dim kernel family: every Linear layer is a tcgen05 GEMM over all rows
 * (sequences x particles) of a time step, the per-component math between them runs in
 * fused elementwise kernels, the time loop is a stream-ordered launch sequence. */
typedef struct bfvi_forward_args {
  int32_t T, B;
  const float* inputs[BFVI_MAX_MODS];  /* (T,B,D_m), NaN = missing; NULL = modality not given */
  int32_t mode;                        /* BFVI_MODE_* */
  int32_t sample, sample_init;
  int32_t flt_particles, smt_particles;
  const float* eps_flt;                /* (T,B,flt_particles,Z) injected noise or NULL (Philox) */
  const float* eps_smt;                /* (T,B,smt_particles,Z) */
  uint64_t seed;
  uint32_t b_offset;
  int32_t precision;                   /* BFVI_PREC_TF32X3 (0, default), BFVI_PREC_TF32 (1), BFVI_PREC_FUSED (2): the transitions
                                          run in the fused on-chip kernels (z_dim 64, h_dim a multiple of 128; other shapes
                                          take the 3xTF32 launch sequence) */
  float* infer_mean; float* infer_std; /* outputs (T,B,Z) */
  float* prior_mean; float* prior_std;
  float* recon_mean[BFVI_MAX_MODS];    /* outputs (T,B,D_m), nullable per modality */
  float* recon_std[BFVI_MAX_MODS];
} bfvi_forward_args;

int bfvi_forward_workspace(const bfvi_model* model, const bfvi_forward_args* args, size_t* bytes);
int bfvi_forward(const bfvi_model* model, const float* params, const bfvi_forward_args* args,
                 void* workspace, size_t workspace_bytes, void* stream);

/* y = act(x W^T + b) on the tcgen05 tensor cores (TF32 operands, rounded to nearest
 * while staged; FP32 accumulation in tensor memory): the dense layers of GaussianMLP /
 * GaussianGTF at large batch (models/common.py:38-41, 62-68).  x (n_rows, n_in) row-major
 * with leading dimension ldx; w (n_out, n_in) in nn.Linear layout, leading dimension ldw;
 * bias (n_out) nullable; y (n_rows, n_out), leading dimension ldy.  act: 0 none, 1 ReLU;
 * | 16 selects single-pass TF32 instead of the default error-compensated 3xTF32
 * (a_hi w_hi + a_lo w_hi + a_hi w_lo in one accumulator: FP32-class accuracy). */
int bfvi_linear_tf32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                     float* y, int64_t ldy, int64_t n_rows, int32_t n_in, int32_t n_out,
                     int32_t act, void* stream);

/* Weight gradient of a Linear layer on the tcgen05 tensor cores: dw (n_out, n_in) (+)= dy^T x,
 * given the TRANSPOSED activations dy_t (n_out, n_rows) and x_t (n_in, n_rows) (the contraction
 * runs over the rows, which makes both operands K-major; the producing kernels of this library
 * write the transposed copies themselves).  accumulate != 0 adds into dw.  flags: 16 =
 * single-pass TF32 (default error-compensated 3xTF32). */
int bfvi_wgrad_tf32(const float* dy_t, int64_t lddy, const float* x_t, int64_t ldx, float* dw, int64_t lddw,
                    int64_t n_rows, int32_t n_out, int32_t n_in, int32_t accumulate, int32_t flags,
                    void* stream);

/* GaussianGTF.forward (models/common.py:62-68) on n_rows latent rows (z_dim 64, h_dim a multiple of 128) through the
 * FUSED on-chip kernels — the building block of the large-dim family's mode BFVI_PREC_FUSED: per 128-row tile
 * z -> hidden -> heads runs as a chain of tcgen05 MMAs whose hidden activations stay in tensor memory; weights are
 * split into scaled FP16 hi / lo operand tiles once per call and streamed by cp.async.bulk.  Outputs are the four pre-activation heads with their biases
 * added, (n_rows, 64) each: gate_pre (before the sigmoid), nonlin, lin, std_pre (before softplus + min_std), so that
 * mean = (1 - sigmoid(gate_pre)) * lin + sigmoid(gate_pre) * nonlin and std = softplus(std_pre) + min_std.
 * keep != 0 also leaves the operands of bfvi_gtf_bwd in the workspace (bfvi_gtf_workspace(model, n_rows) bytes,
 * 256-byte aligned; 0 = shape not served). */
size_t bfvi_gtf_workspace(const bfvi_model* model, int64_t n_rows);
int bfvi_gtf_fwd(const bfvi_model* model, const float* params, int32_t direction, const float* z, int64_t n_rows,
                 float* gate_pre, float* nonlin, float* lin, float* std_pre, int32_t keep, void* workspace,
                 size_t workspace_bytes, void* stream);
/* Backward of the same rows (bfvi_gtf_fwd(keep = 1) ran on this workspace; `nonlin` is its output): from the gradients
 * at the four heads, d_z (n_rows, 64) is written and every parameter gradient of trans[direction] is ACCUMULATED into
 * `grads` (flat layout).  The hidden gradients never leave the SM; the H-wide weight gradients contract FP16-rounded
 * operand tiles with FP32 accumulation. */
int bfvi_gtf_bwd(const bfvi_model* model, const float* params, float* grads, int32_t direction, const float* z,
                 const float* nonlin, int64_t n_rows, const float* d_gate_pre, const float* d_nonlin, const float* d_lin,
                 const float* d_std_pre, float* d_z, void* workspace, size_t workspace_bytes, void* stream);

/* Measurement aid (bench.py's per-kernel roofline): average device time of ONE launch of a fused transition kernel on
 * n_rows latent rows, CUDA events on `stream` around `iters` back-to-back launches (SYNCHRONISES).  which: 0 = forward
 * (gtf_fwd_kernel), 1 = forward<keep> (the backward's recompute), 2 = input gradient (gtf_bwd_kernel), 3 = H-wide weight
 * gradients (wgrad16_kernel).  scratch_rows: 5 * n_rows * 64 floats (n_rows >= 4 * h_dim for which = 3); workspace as
 * bfvi_gtf_workspace. */
int bfvi_gtf_probe(const bfvi_model* model, const float* params, int32_t direction, int32_t which, const float* z,
                   int64_t n_rows, int32_t iters, float* scratch_rows, void* workspace, size_t workspace_bytes,
                   float* ms_per_launch, void* stream);

/* One per-modality encoder / decoder MLP of the composed path at ANY size, weights passed by pointer (the modules that
 * own them keep the reference's state_dict keys): common.GaussianMLP (models/common.py:25-41), common.CategoricalMLP
 * (models/common.py:9-23) and the categorical encoder Embedding -> ReLU -> GaussianMLP (models/dmm.py:78-82).  Dense
 * layers run on the tcgen05 tensor cores (error-compensated 3xTF32, FP32-class), everything between them in fused
 * elementwise kernels; nothing here calls a library GEMM.
 *   x        (n_rows, n_in) fp32; with `emb` set: (n_rows) class indices stored as floats (NaN counts as class 0, the
 *            reference zero-fills before .long(), models/dmm.py:166-167)
 *   out_a    (n_rows, n_out): mean (BFVI_HEAD_GAUSSIAN) or class probabilities (BFVI_HEAD_SOFTMAX)
 *   out_b    (n_rows, n_out): std = softplus(.) + min_std (Gaussian head only)
 *   mask     (n_rows) u8, nullable: 1 where the row has no NaN (written when nan_mask != 0; NaNs are zero-filled)
 * bfvi_mlp_bwd recomputes the hidden activations from x, takes the forward outputs and their gradients, ACCUMULATES the
 * parameter gradients (+=) and writes d_x (n_rows, n_in; nullable; unused with `emb`). */
enum { BFVI_HEAD_GAUSSIAN = 0, BFVI_HEAD_SOFTMAX = 1 };
typedef struct bfvi_mlp_desc {
  const float* emb;                     /* nullable: (n_classes, h_dim) embedding table (then n_in == h_dim) */
  const float* w1; const float* b1;     /* in_to_h.0: (h_dim, n_in), (h_dim) */
  const float* wa; const float* ba;     /* h_to_mean | h_to_out.0: (n_out, h_dim), (n_out) */
  const float* wb; const float* bb;     /* h_to_std.0 (Gaussian head); null for the softmax head */
  int32_t n_in, h_dim, n_out, n_classes;
  int32_t head;                         /* BFVI_HEAD_* */
  int32_t nan_mask;                     /* encoders: NaN -> mask + zero fill (models/dmm.py:165-167) */
  float min_std;                        /* 1e-3 for every GaussianMLP of the reference (models/common.py:27) */
  int32_t pad_;
} bfvi_mlp_desc;
typedef struct bfvi_mlp_grads { float* emb; float* w1; float* b1; float* wa; float* ba; float* wb; float* bb; } bfvi_mlp_grads;
int bfvi_mlp_workspace(const bfvi_mlp_desc* desc, int64_t n_rows, int32_t backward, size_t* bytes);
int bfvi_mlp_fwd(const bfvi_mlp_desc* desc, const float* x, int64_t n_rows, float* out_a, float* out_b, uint8_t* mask,
                 void* workspace, size_t workspace_bytes, void* stream);
int bfvi_mlp_bwd(const bfvi_mlp_desc* desc, const bfvi_mlp_grads* grads, const float* x, int64_t n_rows, const float* out_a,
                 const float* out_b, const float* d_out_a, const float* d_out_b, float* d_x, void* workspace,
                 size_t workspace_bytes, void* stream);

/* The optimiser step either side of the hot path (trainer.py:248-252), fused over the flat
 * buffers: optional clip_grad_norm_ (max_norm > 0; total L2 norm over the whole flat gradient,
 * coefficient max_norm / (norm + 1e-6) clamped to 1) followed by torch.optim.Adam (no amsgrad;
 * weight_decay added to the gradient like torch).  `grad_scale` multiplies the gradient first
 * (e.g. 1 / sum(lengths)).  `step` is the 1-based step count; `norm_scratch` is one device float.
 * Two launches, no host synchronisation. */
int bfvi_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                   float grad_scale, float max_norm, float* norm_scratch, void* stream);

/* ---- Image modules of the Weizmann / vidTIMIT models (models/common.py:70-175) --------------------------------------
 * common.Conv = nn.Conv2d -> BatchNorm2d -> ReLU, common.Deconv = nn.ConvTranspose2d -> BatchNorm2d -> ReLU, the
 * decoder's final nn.Sigmoid; FP32 NCHW tensors, dilation 1, groups 1, output_padding 0.  One geometry serves both layer
 * kinds: a SMALL map (Conv2d output / ConvTranspose2d input) and a BIG map (Conv2d input / ConvTranspose2d output)
 * with  h_big = h_small * stride - padding + kh, and ONE weight layout w[c_small][c_big][k][k] = nn.Conv2d.weight
 * (out, in, k, k) = nn.ConvTranspose2d.weight (in, out, k, k).
 *                               Conv2d (models/common.py:75-78)      ConvTranspose2d (models/common.py:96-99)
 *   bfvi_conv_gather            forward (bias, x -> y)               input gradient (bias = NULL, dy -> dx)
 *   bfvi_conv_scatter           input gradient (bias = NULL)         forward (bias; act = BFVI_ACT_SIGMOID fuses the
 *                                                                    decoder's final sigmoid, models/common.py:148)
 *   bfvi_conv_wgrad             dw += (small = dy, big = x)          dw += (small = x, big = dy)
 *   bfvi_chan_bias_grad         db += sum over images and pixels of dy (either layer)
 * Direct FP32 convolutions (the reference's convolutions are FP32, not TF32); kernel sizes up to 7.
 * Summation order: bfvi_conv_wgrad adds the partial sums of its pixel splits with float atomics, and bfvi_dense_* slices a
 * long contraction over idle SMs the same way when there are few output tiles, so the last bits of those results can
 * differ from run to run (like cuDNN's default backward algorithms).  BFVI_DETERMINISTIC=1 in the environment (read per
 * call) keeps one writer per output element: bit-identical runs, less parallelism on small problems. */
typedef struct bfvi_conv_geom {
  int32_t n;                          /* images (T * B frames) */
  int32_t c_small, h_small, w_small;
  int32_t c_big, h_big, w_big;
  int32_t kernel, stride, padding;
} bfvi_conv_geom;
enum { BFVI_ACT_NONE = 0, BFVI_ACT_SIGMOID = 2 };
int bfvi_conv_gather(const bfvi_conv_geom* geom, const float* big, const float* w, const float* bias, float* small,
                     int32_t act, void* stream);
int bfvi_conv_scatter(const bfvi_conv_geom* geom, const float* small, const float* w, const float* bias, float* big,
                      int32_t act, void* stream);
int bfvi_conv_wgrad(const bfvi_conv_geom* geom, const float* small, const float* big, float* dw, void* stream);
/* scratch of the per-channel reductions below (8-byte aligned device memory) */
size_t bfvi_chan_scratch(int32_t channels);
int bfvi_chan_bias_grad(const float* dy, int32_t N, int32_t C, int64_t HW, float* db, void* scratch, size_t scratch_bytes,
                        void* stream);
/* nn.BatchNorm2d (+ the ReLU that follows it in common.Conv / common.Deconv, models/common.py:79-84, 100-105) on x
 * (N, C, HW).  training != 0: batch statistics (biased variance), running_mean / running_var blended with `momentum`
 * (unbiased variance) when given; training == 0: the running statistics.  mean_rstd (C, 2) receives (mean, 1 / sqrt(var +
 * eps)) for the backward.  Deterministic: per-channel partial sums in double, fixed order.
 * bfvi_bn2d_bwd: dy is the gradient at y; relu != 0 masks it where y <= 0; d_gamma / d_beta are ACCUMULATED (nullable). */
int bfvi_bn2d_fwd(const float* x, int32_t N, int32_t C, int64_t HW, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, int32_t training, float momentum, float eps, int32_t relu,
                  float* y, float* mean_rstd, void* scratch, size_t scratch_bytes, void* stream);
int bfvi_bn2d_bwd(const float* dy, const float* x, const float* y, const float* mean_rstd, const float* gamma, int32_t N,
                  int32_t C, int64_t HW, int32_t training, int32_t relu, float* dx, float* d_gamma, float* d_beta,
                  void* scratch, size_t scratch_bytes, void* stream);
/* dx = dp * p * (1 - p): backward of the sigmoid fused into bfvi_conv_scatter */
int bfvi_sigmoid_bwd(const float* p, const float* dp, int64_t n, float* dx, void* stream);

/* nn.Linear [-> nn.ReLU] of the image modules (feat_to_z_mean / feat_to_z_std.0 / z_to_feat.0, models/common.py:127-133,
 * 146-149) in FP32 on the FFMA pipe: x (rows, n_in), w (n_out, n_in), y (rows, n_out), all dense row-major.  (The contraction
 * over feat_dim = 4096 leaves the tensor cores' 3xTF32 product at 2e-5 of the result — enough to flip BatchNorm -> ReLU masks
 * downstream; the FFMA kernel runs each GEMM in 80 - 100 us at 625 frames, 13 - 17 TFLOP/s.)
 * bfvi_dense_bwd: dy is the gradient at y; relu != 0 masks it where y <= 0 into dy_masked (rows, n_out; scratch the caller
 * provides); dx (nullable) is written, dw and db (nullable) are ACCUMULATED; scratch as bfvi_chan_scratch(n_out). */
int bfvi_dense_fwd(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t n_in, int32_t n_out,
                   int32_t relu, void* stream);
int bfvi_dense_bwd(const float* x, const float* w, const float* y, const float* dy, float* dy_masked, int64_t rows,
                   int32_t n_in, int32_t n_out, int32_t relu, float* dx, float* dw, float* db, void* scratch,
                   size_t scratch_bytes, void* stream);

/* FP32 FFMA throughput probe: `blocks` CTAs x 256 threads x iters x 16 FMAs
 * (measurement aid: the roofline denominator of the FFMA-bound small-dim path). */
int bfvi_ffma_probe(float* out, int32_t iters, int32_t blocks, void* stream);

/* Materialise the Philox stream (same generator the kernels use) as an external
 * noise tensor (S, T, B, K, Z) so the oracle can be run on identical draws. */
int bfvi_dump_noise(uint64_t seed, uint32_t stream_id, uint32_t b_offset, int32_t S,
                    int32_t T, int32_t B, int32_t K, int32_t Z, float* out,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BFVI_H_ */
