/* titanet_b200.h -- C ABI of libtitanet_sm100.so (NVIDIA B200, sm_100a).
 *
 * The drop-in boundary for the TitaNet hot path of Wadaboa/titanet: the reference is
 * pure Python on torch, so there is no FFI in it to replace; each entry point below
 * is what a ctypes binding inside the reference's own modules would call instead of
 * the torch op(s) named in its comment (paths relative to the reference checkout,
 * commit 7b77053).  INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative TN_E* code for a rejected
 *    argument (shape / alignment / unsupported value) or a positive cudaError_t.
 *    tn_last_error() returns the message for the calling thread.
 *  - All tensors are device pointers to contiguous fp32 unless stated.  Activations
 *    use the channels-last "NWC" layout [B*T, C] (row r = b*T + t); the reference's
 *    [B, C, T] tensors cross the boundary through tn_ncw_to_nwc / tn_nwc_to_ncw (or
 *    the mel / prolog kernels, which read [B, C, T] directly).
 *  - Nothing allocates, synchronises or owns memory: outputs and scratch are passed
 *    in by the caller; `stream` is a cudaStream_t.  All launches are CUDA-graph
 *    capturable.
 *  - "Lazy activation": a tensor produced by a conv in front of a BatchNorm is kept as
 *    its pre-BN values z; consumers receive (scale, shift, relu, drop_p, seed, layer)
 *    and compute a = dropout(relu(z*scale[c] + shift[c])) on load.  scale == NULL
 *    means the tensor is already an activation.  Dropout masks come from a
 *    counter-based hash (seeded multiply-xorshift) of (seed, layer, element index), so the
 *    backward pass regenerates them.
 *  - "ACCUMULATED" outputs are added to (atomically); the caller zeroes them.
 */
#ifndef TITANET_B200_H_
#define TITANET_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TN_OK 0
#define TN_EINVAL (-1)
#define TN_EUNSUPPORTED (-2)

/* epilogue flags of the conv-GEMM entry points */
#define TN_EPI_TANH 1   /* z = tanh(z)                       */
#define TN_EPI_ACCUM 2  /* z += previous contents of Z       */
#define TN_GEMM_GRAD 8  /* tensor-core GEMMs: this is a gradient GEMM (the cheaper TF32 + bf16-correction split is allowed) */

/* Scratch of the kernels that reduce across thread blocks WITHOUT floating-point atomics, so that a forward pass is
 * reproducible bit for bit:
 *   accum   BatchNorm statistics of the GEMM epilogues: every block adds its per-channel partial sums into 120-bit fixed-point
 *           accumulators (two unsigned 64-bit integer atomics per value: integer addition is associative, the order of
 *           arrival does not matter); the last block of a channel group (ticket) reads the totals.  TN_ACCUM_WORDS(C)
 *           64-bit words for C channels, ZERO on entry, returned to zero by the kernel.
 *   tickets TN_TICKETS unsigned ints, ZERO on entry and returned to zero by every kernel that uses them.
 *   parts   uninitialised fp32 workspace of parts_floats elements, 16-byte aligned: split-K partial tiles of
 *           tn_conv_gemm_simt (tn_conv_gemm_simt_scratch_floats says how many), added in split order by the last block.
 * One accum / tickets pair per stream may be shared by all calls (kernels that run concurrently must not share it). */
#define TN_TICKETS 64
#define TN_ACCUM_WORDS(C) (4 * (long long)(C) + TN_TICKETS / 2)
typedef struct tn_scratch {
  float* parts;
  long long parts_floats;
  unsigned int* tickets;
  unsigned long long* accum;
  long long accum_words;
} tn_scratch;

const char* tn_last_error(void);
int tn_version(void);
int tn_device_check(void);                       /* fails unless the current device is sm_100 */
int tn_zero(void* p, size_t bytes, void* stream);

/* ---- layout ------------------------------------------------------------------ */
int tn_ncw_to_nwc(const float* x, float* y, int B, int C, int T, void* stream);
int tn_nwc_to_ncw(const float* x, float* y, int B, int C, int T, void* stream);

/* ---- mel front end: transforms.MelSpectrogram.__call__ (src/transforms.py:158-184),
 *      batched with datasets.collate_fn zero padding (src/datasets.py:48-73) -------- */
int tn_mel_fwd(const float* wave, const int* lengths, const float* window, const float* fb, const int* band_lo,
               const int* band_hi, float* out, int B, int L_stride, int L_full, int T_out, int n_fft, int hop,
               int n_mels, int nwc, void* stream);
/* The same with SpecAugment for given per-utterance draws (src/transforms.py:168-175: torchaudio TimeStretch at a
 * random rate, 187-201: mask_along_axis over the mel axis and the time axis, value 0.0).  rates [B] fp64 (1.0 = bypass)
 * and frames [B] = ceil(T_b / rate_b), both NULL = no stretching; masks [B, n_fmask + n_tmask, 2] int32 half-open
 * (start, end) ranges, frequency masks first (NULL = none).  Frames >= frames[b] are zero filled. */
int tn_mel_specaug_fwd(const float* wave, const int* lengths, const float* window, const float* fb, const int* band_lo,
                       const int* band_hi, const double* rates, const int* frames, const int* masks, int n_fmask,
                       int n_tmask, float* out, int B, int L_stride, int L_full, int T_out, int n_fft, int hop, int n_mels,
                       int nwc, void* stream);

/* ---- convolutions as GEMMs ------------------------------------------------------
 * Z[r,co] = bias[co] + sum_{k,ci} X[r+k-K/2, ci] * W[co,ci,k]  (taps stay inside an
 * utterance).  Replaces Conv1dSamePadding.forward (src/modules.py:14-40) for dense convs,
 * the skip nn.Conv1d (src/models.py:452-455) and nn.Linear (src/models.py:549-551,
 * 510-513; src/losses.py:30,70).  stats (fp64 [2*Co], written; needs Co %% 4 == 0) receives the
 * per-channel sum and sum of squares of Z for the following BatchNorm.
 * transpose_w = 1 computes the data gradient (X := dZ, weight read as W[ci',co',K-1-k]).
 * scratch: required for statistics (accum + tickets) and when tn_conv_gemm_simt_scratch_floats(...) > 0 (a skinny problem
 * run as split-K: parts + tickets). */
long long tn_conv_gemm_simt_scratch_floats(int B, int T, int Ci, int Co, int K, int flags);
int tn_conv_gemm_simt(const float* X, const float* W, const float* bias, float* Z, double* stats, int B, int T, int Ci,
                      int Co, int K, int transpose_w, int flags, const tn_scratch* scratch, void* stream);
/* dW[co,ci,k] += sum_r dZ[r,co] X[r+k-K/2,ci];  dbias[co] += sum_r dZ[r,co]   (ACCUMULATED) */
int tn_conv_wgrad_simt(const float* dZ, const float* X, float* dW, float* dbias, int B, int T, int Ci, int Co, int K,
                       void* stream);

/* K-tap dense conv as one tensor-core GEMM (the prolog conv, src/models.py:370): tn_im2col_nwc unrolls the taps into the
 * reduction dimension, out[r, tap*Ci + ci] = x[r + tap - K/2, ci] (zero outside the utterance and from K*Ci up to Kpad, a
 * multiple of 32); tn_conv_weight_gemm reorders the weight the same way (to_gemm = 1: w[Co,Ci,K] -> w3[Co,Kpad]) and its
 * gradient back (to_gemm = 0: dw3[Co,Kpad] -> dw[Co,Ci,K]).  conv(x, w) = out w3^T then runs on tn_gemm_tc / tn_gemm_tc_bn. */
int tn_im2col_nwc(const float* x, float* out, int B, int T, int Ci, int K, int Kpad, void* stream);
int tn_conv_weight_gemm(const float* src, float* dst, int Co, int Ci, int K, int Kpad, int to_gemm, void* stream);

/* Tensor-core path for the 1x1 convs / linears (tcgen05 + TMEM + TMA, split arithmetic = fp32-equivalent):
 * Z[R,M] = bias + X[R,Kd] W[M,Kd]^T.  ws = split weights [TN_WS_PLANES, M, Kd] from tn_split_tf32
 * (transpose = 1 reads W as [Kd, M]: the data-gradient GEMM); its layout is private to the library
 * (ws[0] = tf32(W); ws[1] = tf32(W - ws[0]); ws[2] / ws[3] = the packed bf16 / scaled-fp16 correction rows).
 * nsplit: 3 = fp32-equivalent -- forward GEMMs run tf32 hi*hi + one fp16 MMA over a doubled K that carries both
 * corrections with exponent-balanced scales (operand rounding = 3xTF32's, two tensor issues instead of three), gradient
 * GEMMs (TN_GEMM_GRAD) the same with bf16 (fp32's exponent range) -- or 1 (plain TF32).  When the tile leaves TMEM columns
 * free the main products of consecutive K ranges and the corrections accumulate in separate TMEM accumulators that the epilogue adds
 * in fp32 (the tensor core's accumulate truncates).  Needs Kd %% 32 == 0 and M %% 128 == 0 (tn_gemm_tc_supported).
 * stats (fp64 [2*M], written): per-channel sum / sum of squares of Z; needs scratch (accum + tickets). */
#define TN_WS_PLANES 4
int tn_gemm_tc_supported(int R, int Kd, int M);
int tn_gemm_tc_set_trace(long long* buf);      /* debug: clock64 timeline of two CTAs (256 int64), NULL = off */
int tn_split_tf32(const float* W, float* ws, int M, int Kd, int transpose, void* stream);
int tn_gemm_tc(const float* X, const float* ws, const float* bias, float* Z, double* stats, int R, int Kd, int M, int flags,
               int nsplit, const tn_scratch* scratch, void* stream);
/* Data gradient of a depthwise-separable block in ONE kernel: du = dZ[R,Co] W[Co,C] on the tensor
 * cores (ws = tn_split_tf32(W, transpose = 1)), then, in the epilogue, the transposed depthwise conv,
 * the BN/ReLU/dropout backward of the previous layer and all per-channel reductions:
 * dzprev, dw += , dbias +=, dscale +=, dshift +=  (= tn_gemm_tc(dgrad) + tn_dw_bwd without du in HBM).
 * Autograd of DepthwiseConv1d + ConvBlock1d (src/modules.py:64-79, 119-134) under loss.backward(). */
int tn_gemm_tc_dwbwd(const float* dZ, const float* ws, const float* zprev, float* dzprev, const float* dw_w, float* g_dw,
                     float* g_dbias, float* g_dscale, float* g_dshift, const float* scale, const float* shift, int relu,
                     float drop_p, const unsigned long long* seed, unsigned int layer, int B, int T, int Co, int C, int K,
                     int nsplit, void* stream);
/* dW[Co,Ci] += dZ[R,Co]^T U[R,Ci] on the tensor cores (split-K over rows, MN-major operands; ACCUMULATED) */
int tn_wgrad_tc_supported(int R, int Ci, int Co);
int tn_wgrad_tc(const float* dZ, const float* U, float* dW, int R, int Ci, int Co, void* stream);
int tn_colsum(const float* x, float* out, int R, int C, void* stream);                /* out[c] += sum_r x[r,c] */

/* ---- train-mode BatchNorm folded INTO its producer / consumer kernels --------------------------
 * tn_bn_fold describes the nn.BatchNorm1d that follows a conv (src/modules.py:128, src/models.py:454,
 * 512).  The *_bn GEMM entry points produce the statistics of Z in their epilogue as before; the
 * last CTA of each channel group to finish (device-wide tickets of the tn_scratch) reads the fixed-point totals,
 * folds them into (scale, shift), stores (mean, invstd) for the backward pass and updates the running
 * statistics (momentum, unbiased variance) and num_batches_tracked -- i.e. tn_bn_finalize without its
 * launch.  n = samples per channel (B*T). */
typedef struct tn_bn_fold {
  const float* gamma;              /* [C] BatchNorm weight                      */
  const float* beta;               /* [C] BatchNorm bias                        */
  float* running_mean;             /* [C] or NULL (track_running_stats=False)   */
  float* running_var;              /* [C] or NULL                               */
  long long* num_batches_tracked;  /* scalar or NULL                            */
  float momentum, eps;
  double n;
  float* scale;                    /* [C] out: gamma * invstd                   */
  float* shift;                    /* [C] out: beta - mean * scale              */
  float* mean;                     /* [C] out (saved for backward)              */
  float* invstd;                   /* [C] out (saved for backward)              */
} tn_bn_fold;
int tn_gemm_tc_bn(const float* X, const float* ws, const float* bias, float* Z, double* stats, const tn_bn_fold* bn, int R,
                  int Kd, int M, int flags, int nsplit, const tn_scratch* scratch, void* stream);
/* Forward of a depthwise-separable block (+ the following train-mode BatchNorm's statistics / fold) in ONE kernel:
 * the GEMM's transform warps build the operand u = depthwise_K(act(z)) + b_dw from the raw z tile (BN/ReLU/dropout on load,
 * K-tap FIR, tf32 split) instead of reading a u tensor written by tn_dw_fwd; u_out (optional) receives u for the backward
 * weight gradient.  Z = u W^T + b_pw (3xTF32).  bn may be NULL (statistics only, or none when stats is NULL too).
 * DepthwiseConv1d.forward + the BatchNorm1d of ConvBlock1d (src/modules.py:64-93, 119-134). */
int tn_gemm_tc_dwfwd(const float* z, const float* ws, const float* dw_w, const float* dw_b, const float* scale,
                     const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer,
                     const float* pw_bias, float* u_out, float* Z, double* stats, const tn_bn_fold* bn, int B, int T, int C,
                     int Co, int K, const tn_scratch* scratch, void* stream);
int tn_conv_gemm_simt_bn(const float* X, const float* W, const float* bias, float* Z, double* stats, const tn_bn_fold* bn,
                         int B, int T, int Ci, int Co, int K, int flags, const tn_scratch* scratch, void* stream);
/* Backward of conv -> train-mode BatchNorm fold in one pass over the tensor (tn_bn_bwd_coef + tn_stats_bwd):
 * from dL/dscale, dL/dshift and the saved (mean, invstd) compute, per channel, dgamma, dbeta and the
 * statistics-path coefficients, then out = dz_direct + a[c] + b[c] * z and dbias[c] += sum_r out
 * (dbias ACCUMULATED; dz_direct may be NULL = 0; out may alias dz_direct). */
int tn_bn_stats_bwd(const float* dz_direct, const float* z, const float* dscale, const float* dshift, const float* mean,
                    const float* invstd, const float* gamma, double n, float* out, float* dbias, float* dgamma,
                    float* dbeta, int R, int C, void* stream);
/* BatchNorm backward folded into the operand load of the data-gradient GEMM (pair kernel: R >= 512, M %% 256 == 0,
 * tn_gemm_tc_bnbwd_supported).  dZ is the DIRECT gradient w.r.t. the pre-BatchNorm tensor z [R, Kd] (what the consumers of
 * the lazy activation return); the transform warps build g = dZ + a[c] + b[c] z from the dZ tile and the z tile with the
 * statistics-path coefficients a, b of tn_bn_stats_bwd (computed in the kernel from dscale, dshift, mean, invstd, gamma), feed
 * it to the tensor core and write it to g_out (the weight-gradient GEMM's operand); dgamma / dbeta are written, and dbias
 * (may be NULL) is written as ZERO: the bias of a conv in front of a train-mode BatchNorm has no gradient when dZ comes
 * from the consumers of the lazy activation (sum_r g = sum_r dZ - gamma invstd dshift = 0; the reference's value is the
 * rounding noise of that cancellation).  Replaces one tn_bn_stats_bwd launch per conv. */
typedef struct tn_bn_bwd {
  const float* z;        /* [R, Kd] pre-BatchNorm output of the conv                       */
  const float* dscale;   /* [Kd] dL/dscale of the folded BatchNorm (sum over rows)         */
  const float* dshift;   /* [Kd] dL/dshift                                                 */
  const float* mean;     /* [Kd] saved by the forward fold                                 */
  const float* invstd;   /* [Kd]                                                           */
  const float* gamma;    /* [Kd] BatchNorm weight                                          */
  double n;              /* samples per channel                                            */
  float* g_out;          /* [R, Kd] out: full gradient w.r.t. z                            */
  float* dbias;          /* [Kd] or NULL, out: conv-bias gradient (identically zero, see above) */
  float* dgamma;         /* [Kd] out                                                       */
  float* dbeta;          /* [Kd] out                                                       */
} tn_bn_bwd;
int tn_gemm_tc_bnbwd_supported(int R, int Kd, int M);
int tn_gemm_tc_bnbwd(const float* dZ, const float* ws, const tn_bn_bwd* bnb, float* dX, int R, int Kd, int M, int flags,
                     void* stream);
int tn_gemm_tc_dwbwd_bn(const float* dZ, const float* ws, const tn_bn_bwd* bnb, const float* zprev, float* dzprev,
                        const float* dw_w, float* g_dw, float* g_dbias, float* g_dscale, float* g_dshift, const float* scale,
                        const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer, int B,
                        int T, int Co, int C, int K, void* stream);
/* every weight split of a step in ONE launch: jobs (device array) = {W, ws, M, Kd, transpose, tile0} like tn_split_tf32;
 * tile0 = number of 32 x 32 tiles of the jobs before this one (a job has ceil(M / 32) * (Kd / 32)), total_tiles = their sum.
 * Only the planes the step's GEMMs read are written (forward scheme for transpose = 0, gradient scheme for transpose = 1). */
typedef struct tn_split_job {
  const float* W;
  float* ws;
  int M, Kd, transpose, tile0;
} tn_split_job;
int tn_split_tf32_batch(const tn_split_job* jobs_dev, int njobs, int total_tiles, void* stream);

/* ---- depthwise conv with fused lazy-activation prologue:
 *      DepthwiseConv1d's first conv (src/modules.py:64-75) after BN/ReLU/Dropout
 *      (src/modules.py:128-133) ---------------------------------------------------- */
int tn_dw_fwd(const float* z, float* u, const float* w, const float* bias, const float* scale, const float* shift,
              int relu, float drop_p, const unsigned long long* seed, unsigned int layer, int B, int T, int C, int K,
              void* stream);
int tn_dw_bwd(const float* du, const float* z, float* dz, const float* w, float* dw, float* dbias, float* dscale,
              float* dshift, const float* scale, const float* shift, int relu, float drop_p, const unsigned long long* seed,
              unsigned int layer, int B, int T, int C, int K, void* stream);

/* ---- BatchNorm1d folded into per-channel (scale, shift): nn.BatchNorm1d in
 *      ConvBlock1d (src/modules.py:128), skip (src/models.py:454), decoder
 *      (src/models.py:506,512) ------------------------------------------------------ */
int tn_colstats(const float* x, double* stats, int R, int C, void* stream);          /* written (fixed-order reduction) */
int tn_bn_finalize(const double* stats, double n, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, long long* num_batches_tracked, float momentum, float eps, int training,
                   float* scale, float* shift, float* mean, float* invstd, int C, void* stream);
int tn_bn_bwd_coef(const float* dscale, const float* dshift, const float* mean, const float* invstd, const float* gamma,
                   double n, int training, float* dgamma, float* dbeta, double* dstats, int C, void* stream);
/* out = (dz_direct or 0) + dstats[c] + 2 z dstats[C+c]: gradient of the statistics w.r.t. their tensor;
 * dbias (optional, ACCUMULATED) += column sums of out (the conv-bias gradient, same pass) */
int tn_stats_bwd(const float* dz_direct, const float* z, const double* dstats, float* out, float* dbias, int R, int C,
                 void* stream);
/* dropout seed stream: state = splitmix64 step, *out = this step's seed (device scalars) */
int tn_seed_next(unsigned long long* state, unsigned long long* out, void* stream);
int tn_act_fwd(const float* z, float* y, const float* scale, const float* shift, int relu, float drop_p,
               const unsigned long long* seed, unsigned int layer, int R, int C, void* stream);
int tn_act_bwd(const float* dy, const float* z, float* dz, float* dscale, float* dshift, const float* scale,
               const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer, int R, int C,
               void* stream);
/* the same for an activation with two consumers: dy2 (or NULL) is added to dy on load (replaces autograd's sum kernel) */
int tn_act_bwd2(const float* dy, const float* dy2, const float* z, float* dz, float* dscale, float* dshift, const float* scale,
                const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer, int R, int C,
                void* stream);                                                        /* dscale/dshift ACCUMULATED */
int tn_tanh_bwd(const float* dh, const float* h, float* out, long long n, void* stream);

/* ---- squeeze-excitation + mega-block tail: SqueezeExcitation.forward
 *      (src/modules.py:173-189), MegaBlock.forward (src/models.py:467-472) ----------- */
/* m[b,c] = mean_t act(z3)[b,t,c]  (written, fixed-order reduction; scale == NULL: z3 is already the activation) */
int tn_se_mean(const float* z3, float* m, const float* scale, const float* shift, int relu, float drop_p,
               const unsigned long long* seed, unsigned int layer, int B, int T, int C, void* stream);
/* tn_se_mean + tn_se_mlp_fwd in ONE launch: a thread-block cluster per utterance squeezes, exchanges the means and the hidden
 * layer's partial sums through distributed shared memory in a fixed order (reproducible, no atomics) and writes m and gate.
 * Needs tn_se_squeeze_excite_supported(C, Cr) (C <= 1024, Cr a power of two >= 8). */
int tn_se_squeeze_excite_supported(int C, int Cr);
int tn_se_squeeze_excite(const float* z3, float* m, float* gate, const float* W1, const float* W2, const float* scale,
                         const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer, int B,
                         int T, int C, int Cr, void* stream);
/* tn_tail_bwd1 + tn_se_mlp_bwd in one launch (dgate, dW1, dW2 ACCUMULATED; counters: B zeroed unsigned ints, device-wide tickets reset by the kernel) */
int tn_tail_bwd1_mlp(const float* dout, const float* out, const float* z3, float* dgate, unsigned int* counters,
                     const float* gate, const float* m, const float* W1, const float* W2, float* dm, float* dW1, float* dW2,
                     const float* scale3, const float* shift3, float drop3, unsigned int layer3, float drop_o,
                     const unsigned long long* seed, int B, int T, int C, int Cr, void* stream);
int tn_se_mlp_fwd(const float* m, const float* W1, const float* W2, float* gate, int B, int C, int Cr, void* stream);
int tn_se_mlp_bwd(const float* dgate, const float* gate, const float* m, const float* W1, const float* W2, float* dm,
                  float* dW1, float* dW2, int B, int C, int Cr, void* stream);       /* dW1/dW2 ACCUMULATED */
int tn_tail_fwd(const float* z3, const float* s, const float* gate, float* out, const float* scale3, const float* shift3,
                float drop3, unsigned int layer3, const float* scale_s, const float* shift_s, float drop_o,
                unsigned int layer_o, const unsigned long long* seed, int B, int T, int C, void* stream);
/* dgate is ACCUMULATED (the caller zeroes it) */
int tn_tail_bwd1(const float* dout, const float* out, const float* z3, float* dgate, const float* scale3,
                 const float* shift3, float drop3, unsigned int layer3, float drop_o, const unsigned long long* seed, int B, int T,
                 int C, void* stream);
/* the same for a block output with two consumers: the gradients dout and dout2 are added on load and dsum = dout + dout2 is
 * written for tn_tail_bwd2 (replaces autograd's sum kernel) */
int tn_tail_bwd1s(const float* dout, const float* dout2, float* dsum, const float* out, const float* z3, float* dgate,
                  const float* scale3, const float* shift3, float drop3, unsigned int layer3, float drop_o,
                  const unsigned long long* seed, int B, int T, int C, void* stream);
int tn_tail_bwd2(const float* dout, const float* out, const float* z3, const float* s, const float* gate, const float* dm,
                 float* dz3, float* ds, float* dsc3, float* dsh3, float* dscs, float* dshs, const float* scale3,
                 const float* shift3, float drop3, unsigned int layer3, const float* scale_s, const float* shift_s,
                 float drop_o, const unsigned long long* seed, int B, int T, int C, void* stream);
/* squeeze + excitation + tail forward in ONE cluster kernel (tn_se_squeeze_excite + tn_tail_fwd): each block keeps its
 * activated z3 tile in shared memory between the squeeze and the tail, so z3 is read once.  Needs
 * tn_se_tail_fwd_supported(T, C, Cr) (cluster plan of tn_se_squeeze_excite and a tile of at most 160 KB per block). */
int tn_se_tail_fwd_supported(int T, int C, int Cr);
int tn_se_tail_fwd(const float* z3, const float* s, float* m, float* gate, float* out, const float* W1, const float* W2,
                   const float* scale3, const float* shift3, float drop3, unsigned int layer3, const float* scale_s,
                   const float* shift_s, float drop_o, unsigned int layer_o, const unsigned long long* seed, int B, int T, int C,
                   int Cr, void* stream);

/* stand-alone SE gate multiply (SqueezeExcitation.forward outside a MegaBlock, src/modules.py:187-189):
 * out = x * gate[b,c];  backward: dx = dout * gate, dgate[b,c] += sum_t dout * x  (dgate ACCUMULATED) */
int tn_gate_mul_fwd(const float* x, const float* gate, float* out, int B, int T, int C, void* stream);
int tn_gate_mul_bwd(const float* dout, const float* x, const float* gate, float* dx, float* dgate, int B, int T, int C,
                    void* stream);
/* out[b*T+t, c] = v[b,c] * mul: backward of a mean over time (nn.AdaptiveAvgPool1d(1), src/modules.py:165, src/models.py:498) */
int tn_bcast_rows(const float* v, float* out, float mul, int B, int T, int C, void* stream);

/* ---- attentive statistics pooling: AttentiveStatsPooling.forward (src/models.py:570-584) */
int tn_asp_pool_fwd(const float* e, const float* x, float* pooled, float* aux, int B, int T, int D, float eps,
                    void* stream);
int tn_asp_pool_bwd(const float* dpooled, const float* pooled, const float* aux, const float* e, const float* x, float* de,
                    float* dx, int B, int T, int D, float eps, void* stream);

/* ---- embedding normalisation + loss heads: F.normalize (src/models.py:333,
 *      src/losses.py:43), CELoss.forward (src/losses.py:32-44),
 *      AngularMarginLoss.forward (src/losses.py:77-132) ------------------------------ */
int tn_l2norm_fwd(const float* x, float* y, float* norms, int B, int E, float eps, void* stream);
int tn_l2norm_bwd(const float* dy, const float* y, const float* norms, const float* dnorm, float* dx, int B, int E,
                  float eps, void* stream);
/* dlogits (optional) = d(mean loss)/dlogits * (*gout or 1) */
int tn_ce_fwd_bwd(const float* logits, const long long* targets, float* loss_row, float* loss, long long* preds,
                  float* dlogits, const float* gout, int B, int Cn, void* stream);
int tn_margin_fwd_bwd(const float* raw_cos, const float* norms, const long long* targets, float* loss_row, float* loss,
                      long long* preds, float* draw, float* dnorm, const float* gout, int B, int Cn, float scale,
                      int use_norm_scale, float m1, float m2, float m3, float eps, void* stream);
int tn_rownorm_inplace(float* W, int rows, int cols, float eps, void* stream);

/* ---- the step after the path: torch.optim.Adam over every parameter in ONE launch
 *      (src/train.py:130-136: Adam, lr 1e-3, weight_decay 0).  hyper (device, 10 floats) =
 *      {lr, beta1, beta2, eps, weight_decay, step, 1 - beta1^step, 1 - beta2^step, 1 - beta1, 1 - beta2}; tn_adam_tick advances
 *      step and the two bias corrections on the device, so a captured step replays correctly.
 *      g' = g + wd p;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps) */
typedef struct tn_adam_job {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
} tn_adam_job;
int tn_adam_tick(float* hyper_dev, void* stream);
int tn_adam_multi(const tn_adam_job* jobs_dev, int njobs, long long max_n, const float* hyper_dev, void* stream);

/* ---- the consumer of the eval-mode forward: trial scoring and detection metrics ----------------
 * tn_cosine_scores: scores[i*N+j] = F.cosine_similarity(e_i, e_j) (eps 1e-8) for every ORDERED pair, i.e. the
 * itertools.product order of SpeakerDataset.get_sample_pairs (src/datasets.py:165-183) as consumed by learn.test
 * (src/learn.py:428-439); labels[i*N+j] = (speakers[i] == speakers[j]) (labels / speakers may be NULL).
 * tn_det_metrics: utils.compute_error_rates / compute_mindcf / compute_eer (src/utils.py:294-367) over n trials
 * (fp32 scores, 0/1 byte labels): stable ascending sort, cumulative target / non-target counts, fnrs / fprs (optional
 * fp64 [n] outputs in sorted order), out8 (device, 8 doubles) = {min c_det (before the division by c_def + eps), EER,
 * targets, non-targets, fpr0, tpr0, fpr1, tpr1 (the ROC segment crossing tpr = 1 - fpr)}.  sorted_keys (optional u64 [n]):
 * low 32 bits = trial index at each sorted position.  workspace: 256-byte aligned, tn_det_workspace_bytes(n) bytes
 * (bytes_out is a HOST pointer).  EER is NaN when all trials carry one label. */
int tn_cosine_scores(const float* E, const long long* speakers, float* scores, unsigned char* labels, int N, int D,
                     float eps, void* stream);
int tn_det_workspace_bytes(long long n, long long* bytes_out);
int tn_det_metrics(const float* scores, const unsigned char* labels, long long n, double p_target, double c_fa,
                   double c_miss, double eps, void* workspace, long long workspace_bytes, double* out8, double* fnrs,
                   double* fprs, unsigned long long* sorted_keys, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TITANET_B200_H_ */
