// HBM-bound NWC kernels: layout transposes, lazy-activation materialise / backward,
// BatchNorm statistics -> (scale, shift) folding and its backward, small row ops.
// Reference semantics: nn.BatchNorm1d / nn.ReLU / nn.Dropout inside ConvBlock1d
// (src/modules.py:119-134), skip-connection BN (src/models.py:452-455), decoder BNs
// (src/models.py:504-513), F.normalize (src/models.py:333, src/losses.py:43).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void tn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* tn_last_error(void) { return g_err; }
extern "C" int tn_version(void) { return 100; }

static int g_num_sms = 0;
int tn_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static int g_pdl = -1;
bool tn_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("TN_PDL");
    g_pdl = (e && atoi(e) != 0) ? 1 : 0;      // measured: no gain on the graph-replayed step (the kernels cannot co-reside), so off by default
  }
  return g_pdl != 0;
}

extern "C" int tn_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  TN_CUDA(cudaGetDevice(&dev));
  TN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TN_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  TN_UNSUPPORTED(major != 10, "libtitanet_sm100 needs an sm_100a device, found sm_%d%d", major, minor);
  return TN_OK;
}

extern "C" int tn_zero(void* p, size_t bytes, void* stream) {
  TN_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
  return TN_OK;
}

// rows per block for the row-tiled kernels: aim at >= 4 waves of 8 blocks/SM
#include <stdlib.h>
static int rows_per_block(long long R) {
  if (const char* e = getenv("TN_EW_RPB")) if (atoi(e) > 0) return atoi(e);             // tuning knob
  long long target = (long long)tn_num_sms() * 2;      // few, fat blocks: every block ends in per-channel atomics (2/SM measured best)
  long long rpb = (R + target - 1) / target;
  if (rpb < 32) rpb = 32;
  if (rpb > 256) rpb = 256;
  return (int)rpb;
}

// ---------------------------------------------------------------------------
// [B, C, T] <-> [B, T, C]
// ---------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  tn_grid_dep_sync();
  // per batch item: x is [rows, cols] -> y is [cols, rows]
  __shared__ float tile[32][33];
  const float* xb = x + (size_t)blockIdx.z * rows * cols;
  float* yb = y + (size_t)blockIdx.z * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = blockIdx.y * 32 + j;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? xb[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) yb[(size_t)c2 * rows + r2] = tile[threadIdx.x][j];
  }
}

static int launch_transpose(const float* x, float* y, int B, int rows, int cols, void* stream) {
  TN_REQUIRE(B > 0 && rows > 0 && cols > 0 && B <= 65535, "transpose: bad shape B=%d rows=%d cols=%d", B, rows, cols);
  dim3 grid(tn_cdiv(cols, 32), tn_cdiv(rows, 32), B), block(32, 8);
  tn_launch(transpose_kernel, grid, block, 0, stream, x, y, rows, cols);
  TN_LAUNCH_CHECK("transpose_kernel");
  return TN_OK;
}
extern "C" int tn_ncw_to_nwc(const float* x, float* y, int B, int C, int T, void* stream) {
  return launch_transpose(x, y, B, C, T, stream);
}
extern "C" int tn_nwc_to_ncw(const float* x, float* y, int B, int C, int T, void* stream) {
  return launch_transpose(x, y, B, T, C, stream);
}

// ---------------------------------------------------------------------------
// K-tap dense conv as ONE tensor-core GEMM (the prolog ConvBlock1d(80, H, 3), src/models.py:370): the taps are unrolled
// into the reduction dimension,  X3[r, tap * Ci + ci] = x[r + tap - K/2, ci]  (zero outside the utterance, zero padding up
// to Kpad, a multiple of 32), W3[co, tap * Ci + ci] = w[co, ci, tap], so that conv(x, w) = X3 W3^T runs on tn_gemm_tc_bn
// like every 1x1 conv instead of the CUDA-core kernel (the last contraction that was left there).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_nwc_kernel(const float* __restrict__ x, float* __restrict__ out, int R, int T, int Ci,
                                                         int K, int Kpad) {
  tn_grid_dep_sync();
  const int pad = K / 2;
  const int q4 = Kpad >> 2;                                       // float4s per output row
  const size_t total = (size_t)R * q4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / q4), kk = (int)(i - (size_t)r * q4) * 4;
    const int t = r % T;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kk + j;
      const int tap = k / Ci, ci = k - tap * Ci;
      const int tt = t + tap - pad;
      v[j] = (tap < K && tt >= 0 && tt < T) ? __ldg(x + (size_t)(r + tap - pad) * Ci + ci) : 0.f;
    }
    tn_st4(out + (size_t)r * Kpad + kk, make_float4(v[0], v[1], v[2], v[3]));
  }
}
extern "C" int tn_im2col_nwc(const float* x, float* out, int B, int T, int Ci, int K, int Kpad, void* stream) {
  TN_REQUIRE(x && out && B > 0 && T > 0 && Ci > 0 && K > 0 && (K & 1) && Kpad >= K * Ci && Kpad % 4 == 0 && tn_aligned16(out),
             "im2col_nwc: bad arguments (B=%d T=%d Ci=%d K=%d Kpad=%d)", B, T, Ci, K, Kpad);
  const long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "im2col_nwc: B*T too large");
  long long blocks = (R * (Kpad / 4) + 255) / 256;
  if (blocks > (long long)tn_num_sms() * 16) blocks = (long long)tn_num_sms() * 16;
  tn_launch(im2col_nwc_kernel, (unsigned)blocks, 256, 0, stream, x, out, (int)R, T, Ci, K, Kpad);
  TN_LAUNCH_CHECK("im2col_nwc_kernel");
  return TN_OK;
}
// to_gemm = 1: w3[co, tap * Ci + ci] = w[co, ci, tap] (zeros up to Kpad);  to_gemm = 0: w[co, ci, tap] = w3[co, tap * Ci + ci]
__global__ void conv_weight_gemm_kernel(const float* __restrict__ src, float* __restrict__ dst, int Co, int Ci, int K, int Kpad, int to_gemm) {
  tn_grid_dep_sync();
  const int n = to_gemm ? Co * Kpad : Co * Ci * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (to_gemm) {
      const int co = i / Kpad, kk = i - co * Kpad;
      const int tap = kk / Ci, ci = kk - tap * Ci;
      dst[i] = tap < K ? src[((size_t)co * Ci + ci) * K + tap] : 0.f;
    } else {
      const int co = i / (Ci * K), rem = i - co * Ci * K;
      const int ci = rem / K, tap = rem - ci * K;
      dst[i] = src[(size_t)co * Kpad + tap * Ci + ci];
    }
  }
}
extern "C" int tn_conv_weight_gemm(const float* src, float* dst, int Co, int Ci, int K, int Kpad, int to_gemm, void* stream) {
  TN_REQUIRE(src && dst && Co > 0 && Ci > 0 && K > 0 && Kpad >= K * Ci, "conv_weight_gemm: bad arguments");
  const int n = to_gemm ? Co * Kpad : Co * Ci * K;
  tn_launch(conv_weight_gemm_kernel, tn_cdiv(n, 256), 256, 0, stream, src, dst, Co, Ci, K, Kpad, to_gemm);
  TN_LAUNCH_CHECK("conv_weight_gemm_kernel");
  return TN_OK;
}

// ---------------------------------------------------------------------------
// materialise a lazy activation / its backward
// ---------------------------------------------------------------------------
// Slab tiling of the streaming [R, C] kernels below.  blockIdx.y owns a slab of up to 64 channel quads (256 channels = 1 KB
// per row), the block's 256 threads are `nq` quads x `lanes` row lanes, blockIdx.x owns a contiguous range of rows that each
// thread walks FOUR rows at a time: four independent 16-byte loads per input tensor are in flight per thread (the first
// version walked one row per iteration with one load in flight, and a [R, 1536] tensor was one 256-thread pass and a
// half-empty second one per row: 2.3 - 2.7 TB/s; a [64, 3072] tensor ran on two blocks).
struct TnSlab {
  TnTile tl;
  int q;          // this thread's absolute channel quad
  int r0, r1;     // the block's rows
};
__device__ __forceinline__ TnSlab tn_slab(int R, int C, int rpb) {
  TnSlab s;
  const int Q = C >> 2, q_lo = (int)blockIdx.y * 64, nq = min(64, Q - q_lo);
  s.tl.Q = Q; s.tl.qpb = nq; s.tl.lanes = TN_EW_THREADS / nq;
  s.tl.q0 = (int)threadIdx.x % nq; s.tl.lane = (int)threadIdx.x / nq; s.tl.active = s.tl.lane < s.tl.lanes;
  s.q = q_lo + s.tl.q0;
  s.r0 = (int)blockIdx.x * rpb; s.r1 = min(R, s.r0 + rpb);
  return s;
}
// rows per block for a slab-tiled kernel: `blocks_per_sm` blocks per SM over all slabs, at least 8 rows each
static int slab_rows_per_block(long long R, int C, int blocks_per_sm) {
  const int slabs = tn_cdiv(C >> 2, 64);
  long long row_blocks = ((long long)tn_num_sms() * blocks_per_sm + slabs - 1) / slabs;
  if (row_blocks < 1) row_blocks = 1;
  long long rpb = (R + row_blocks - 1) / row_blocks;
  if (rpb < 8) rpb = 8;
  return (int)rpb;
}

__global__ void __launch_bounds__(TN_EW_THREADS) act_fwd_kernel(const float* __restrict__ z, float* __restrict__ y,
                                                                TnAct act, int R, int C, int rpb) {
  tn_grid_dep_sync();
  act = tn_act_init(act);
  const TnSlab sl = tn_slab(R, C, rpb);
  if (!sl.tl.active) return;
  const int step = sl.tl.lanes;
  for (int r = sl.r0 + sl.tl.lane; r < sl.r1; r += 4 * step) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (r + k * step < sl.r1) v[k] = tn_ld4(z + (size_t)(r + k * step) * C + 4 * sl.q);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (r + k * step < sl.r1) {
        const size_t off = (size_t)(r + k * step) * C + 4 * sl.q;
        tn_st4(y + off, tn_act4(act, v[k], 4 * sl.q, off >> 2, nullptr));
      }
  }
}

// dy2 (optional): a second gradient of the same activation (two consumers), added on load -- replaces autograd's add kernel
__global__ void __launch_bounds__(TN_EW_THREADS) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2,
                                                                const float* __restrict__ z,
                                                                float* __restrict__ dz, float* __restrict__ dscale,
                                                                float* __restrict__ dshift, TnAct act, int R, int C, int rpb) {
  tn_grid_dep_sync();
  act = tn_act_init(act);
  __shared__ float4 red[TN_EW_THREADS];
  const TnSlab sl = tn_slab(R, C, rpb);
  float4 a_sc = tn_zero4(), a_sh = tn_zero4();
  if (sl.tl.active) {
    const float4 sc = tn_ld4(act.scale + 4 * sl.q);
    const int step = sl.tl.lanes;
    for (int r = sl.r0 + sl.tl.lane; r < sl.r1; r += 4 * step) {
      float4 zz[4], g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r + k * step < sl.r1) {
          const size_t off = (size_t)(r + k * step) * C + 4 * sl.q;
          zz[k] = tn_ld4(z + off);
          g[k] = tn_ld4(dy + off);
          if (dy2) g[k] = g[k] + tn_ld4(dy2 + off);
        }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r + k * step < sl.r1) {
          const size_t off = (size_t)(r + k * step) * C + 4 * sl.q;
          float4 m;
          tn_act4(act, zz[k], 4 * sl.q, off >> 2, &m);
          const float4 gm = g[k] * m;
          a_sc = tn_fma4(gm, zz[k], a_sc);
          a_sh = a_sh + gm;
          tn_st4(dz + off, gm * sc);
        }
    }
  }
  tn_lane_reduce_atomic(sl.tl, a_sc, sl.q, dscale, red);
  tn_lane_reduce_atomic(sl.tl, a_sh, sl.q, dshift, red);
}

extern "C" int tn_act_fwd(const float* z, float* y, const float* scale, const float* shift, int relu, float drop_p,
                          const unsigned long long* seed, unsigned int layer, int R, int C, void* stream) {
  TN_REQUIRE(R > 0 && C > 0 && C % 4 == 0, "act_fwd: need C %% 4 == 0 (R=%d C=%d)", R, C);
  TN_REQUIRE(scale && shift, "act_fwd: scale/shift are required");
  TN_REQUIRE(tn_aligned16(z) && tn_aligned16(y) && tn_aligned16(scale) && tn_aligned16(shift), "act_fwd: pointers must be 16B aligned");
  const int rpb = slab_rows_per_block(R, C, 8);
  tn_launch(act_fwd_kernel, dim3(tn_cdiv(R, rpb), tn_cdiv(C >> 2, 64)), TN_EW_THREADS, 0, stream, z, y,
            tn_make_act(scale, shift, relu, drop_p, seed, layer), R, C, rpb);
  TN_LAUNCH_CHECK("act_fwd_kernel");
  return TN_OK;
}

// dz = (dy + dy2) * act'(z) * scale ; dscale += sum (dy + dy2)*act'*z ; dshift += sum (dy + dy2)*act'   (accumulating)
extern "C" int tn_act_bwd2(const float* dy, const float* dy2, const float* z, float* dz, float* dscale, float* dshift,
                           const float* scale, const float* shift, int relu, float drop_p, const unsigned long long* seed,
                           unsigned int layer, int R, int C, void* stream) {
  TN_REQUIRE(R > 0 && C > 0 && C % 4 == 0, "act_bwd: need C %% 4 == 0 (R=%d C=%d)", R, C);
  TN_REQUIRE(scale && shift && dscale && dshift, "act_bwd: scale/shift/dscale/dshift are required");
  TN_REQUIRE(tn_aligned16(z) && tn_aligned16(dy) && tn_aligned16(dy2) && tn_aligned16(dz) && tn_aligned16(scale) && tn_aligned16(shift),
             "act_bwd: pointers must be 16B aligned");
  // every block ends in two atomics per channel and same-address atomics serialise in L2: few, fat blocks -- THREE per SM over all
  // slabs: the kernel needs 74 registers, three blocks are resident per SM, and a grid of four per SM was two waves
  const int rpb = slab_rows_per_block(R, C, 3);
  tn_launch(act_bwd_kernel, dim3(tn_cdiv(R, rpb), tn_cdiv(C >> 2, 64)), TN_EW_THREADS, 0, stream, dy, dy2, z, dz, dscale, dshift,
            tn_make_act(scale, shift, relu, drop_p, seed, layer), R, C, rpb);
  TN_LAUNCH_CHECK("act_bwd_kernel");
  return TN_OK;
}
extern "C" int tn_act_bwd(const float* dy, const float* z, float* dz, float* dscale, float* dshift, const float* scale,
                          const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer, int R,
                          int C, void* stream) {
  return tn_act_bwd2(dy, nullptr, z, dz, dscale, dshift, scale, shift, relu, drop_p, seed, layer, R, C, stream);
}

// ---------------------------------------------------------------------------
// per-channel sum / sum of squares of an [R, C] tensor (fp64 accumulators)
// ---------------------------------------------------------------------------
// One block owns 64 channels (16 quads x 16 row lanes) and walks ALL rows: the lanes' partial sums are combined in a fixed
// order, so the statistics are reproducible bit for bit (no atomics, no zero-initialised output).  Used for small R (the
// decoder's BatchNorms over the batch axis, src/models.py:506,512); the conv outputs get their statistics from the GEMM
// epilogues (tn_stats_finish).
__global__ void __launch_bounds__(TN_EW_THREADS) colstats_kernel(const float* __restrict__ x, double* __restrict__ stats, int R, int C) {
  tn_grid_dep_sync();
  __shared__ double red[2][16][64];
  const int q = threadIdx.x & 15, lane = threadIdx.x >> 4;
  const int c = blockIdx.x * 64 + 4 * q;
  double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < C) {
    for (int r = lane; r < R; r += 16) {
      const float4 v = tn_ld4(x + (size_t)r * C + c);
      s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
      s2[0] += (double)v.x * v.x; s2[1] += (double)v.y * v.y; s2[2] += (double)v.z * v.z; s2[3] += (double)v.w * v.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[0][lane][4 * q + i] = s1[i]; red[1][lane][4 * q + i] = s2[i]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
    const int cc = blockIdx.x * 64 + ch;
    if (cc < C) {
      double a = 0.0;
#pragma unroll
      for (int l = 0; l < 16; ++l) a += red[which][l][ch];
      stats[(size_t)which * C + cc] = a;
    }
  }
}
extern "C" int tn_colstats(const float* x, double* stats, int R, int C, void* stream) {
  TN_REQUIRE(R > 0 && C > 0 && C % 4 == 0 && tn_aligned16(x), "colstats: need C %% 4 == 0 and aligned x (R=%d C=%d)", R, C);
  tn_launch(colstats_kernel, tn_cdiv(C, 64), TN_EW_THREADS, 0, stream, x, stats, R, C);
  TN_LAUNCH_CHECK("colstats_kernel");
  return TN_OK;
}

// per-channel column sums in fp32 (bias gradients): out[c] += sum_r x[r, c]
__global__ void __launch_bounds__(TN_EW_THREADS) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int R, int C, int rpb) {
  tn_grid_dep_sync();
  __shared__ float4 red[TN_EW_THREADS];
  const TnSlab sl = tn_slab(R, C, rpb);
  float4 s1 = tn_zero4();
  if (sl.tl.active) {
    const int step = sl.tl.lanes;
    for (int r = sl.r0 + sl.tl.lane; r < sl.r1; r += 4 * step) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (r + k * step < sl.r1) ? tn_ld4(x + (size_t)(r + k * step) * C + 4 * sl.q) : tn_zero4();
      s1 = s1 + ((v[0] + v[1]) + (v[2] + v[3]));
    }
  }
  tn_lane_reduce_atomic(sl.tl, s1, sl.q, out, red);
}
extern "C" int tn_colsum(const float* x, float* out, int R, int C, void* stream) {
  TN_REQUIRE(R > 0 && C > 0 && C % 4 == 0 && tn_aligned16(x) && out, "colsum: need C %% 4 == 0 and aligned x (R=%d C=%d)", R, C);
  const int rpb = slab_rows_per_block(R, C, 4);
  tn_launch(colsum_kernel, dim3(tn_cdiv(R, rpb), tn_cdiv(C >> 2, 64)), TN_EW_THREADS, 0, stream, x, out, R, C, rpb);
  TN_LAUNCH_CHECK("colsum_kernel");
  return TN_OK;
}

// ---------------------------------------------------------------------------
// BatchNorm folding: statistics -> (scale, shift) [+ running-stat update]
// ---------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double n, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                   long long* __restrict__ nbt, float momentum, float eps, int training,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int C) {
  tn_grid_dep_sync();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) *nbt += 1;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    double m = stats[c] / n;
    double var = stats[C + c] / n - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean) {
      double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
    }
  } else {
    mean = rmean[c];
    invstd = 1.0f / sqrtf(rvar[c] + eps);
  }
  float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
}

extern "C" int tn_bn_finalize(const double* stats, double n, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, long long* num_batches_tracked, float momentum, float eps, int training,
                              float* scale, float* shift, float* mean, float* invstd, int C, void* stream) {
  TN_REQUIRE(C > 0 && gamma && beta && scale && shift && mean && invstd, "bn_finalize: null argument");
  TN_REQUIRE(training ? (stats != nullptr && n >= 1.0) : (running_mean && running_var), "bn_finalize: missing statistics");
  tn_launch(bn_finalize_kernel, tn_cdiv(C, 128), 128, 0, stream, stats, n, gamma, beta, running_mean, running_var,
                                                                         num_batches_tracked, momentum, eps, training, scale,
                                                                         shift, mean, invstd, C);
  TN_LAUNCH_CHECK("bn_finalize_kernel");
  return TN_OK;
}

// Backward of the folding.  Given dL/dscale, dL/dshift it returns dgamma, dbeta and the
// gradient w.r.t. the statistics, dstats = [dL/dS1 | dL/dS2] (fp64, like stats), so that
// the producer of z adds dL/dz_i += dS1[c] + 2 z_i dS2[c]  (tn_stats_bwd).
__global__ void bn_bwd_coef_kernel(const float* __restrict__ dscale, const float* __restrict__ dshift,
                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                   const float* __restrict__ gamma, double n, int training, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, double* __restrict__ dstats, int C) {
  tn_grid_dep_sync();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dsc = dscale[c], dsh = dshift[c], mu = mean[c], r = invstd[c], g = gamma[c];
  double t = dsc - mu * dsh;               // dL/d(invstd) / gamma
  dgamma[c] = (float)(r * t);
  dbeta[c] = (float)dsh;
  if (training && dstats) {
    double dvar = -0.5 * g * t * r * r * r;
    double dmu = -dsh * g * r - 2.0 * mu * dvar;
    dstats[c] = dmu / n;
    dstats[C + c] = dvar / n;
  }
}
extern "C" int tn_bn_bwd_coef(const float* dscale, const float* dshift, const float* mean, const float* invstd,
                              const float* gamma, double n, int training, float* dgamma, float* dbeta, double* dstats,
                              int C, void* stream) {
  TN_REQUIRE(C > 0 && dscale && dshift && mean && invstd && gamma && dgamma && dbeta, "bn_bwd_coef: null argument");
  TN_REQUIRE(!training || dstats, "bn_bwd_coef: training mode needs dstats");
  tn_launch(bn_bwd_coef_kernel, tn_cdiv(C, 128), 128, 0, stream, dscale, dshift, mean, invstd, gamma, n, training, dgamma, dbeta, dstats, C);
  TN_LAUNCH_CHECK("bn_bwd_coef_kernel");
  return TN_OK;
}

// out = (dz_direct or 0) + dS1[c] + 2 * z * dS2[c]   -- the statistics path of the BatchNorm
// backward (and the whole backward of tn_colstats when dz_direct == NULL).  out may alias
// dz_direct.  dbias (optional, ACCUMULATED) receives the column sums of out: the gradient of the
// conv bias in front of the BatchNorm, for free in the same pass.
__global__ void __launch_bounds__(TN_EW_THREADS) stats_bwd_kernel(const float* __restrict__ dzd, const float* __restrict__ z,
                                                                  const double* __restrict__ dstats, float* __restrict__ out,
                                                                  float* __restrict__ dbias, int R, int C, int rpb) {
  tn_grid_dep_sync();
  __shared__ float4 red[TN_EW_THREADS];
  TnTile tl = tn_tile(C);
  const int r0 = blockIdx.x * rpb, r1 = min(R, r0 + rpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    float4 acc = tn_zero4();
    if (tl.active && q < tl.Q) {
      const int c = 4 * q;
      const float4 a = make_float4((float)dstats[c], (float)dstats[c + 1], (float)dstats[c + 2], (float)dstats[c + 3]);
      const float4 b = make_float4((float)(2.0 * dstats[C + c]), (float)(2.0 * dstats[C + c + 1]), (float)(2.0 * dstats[C + c + 2]),
                                   (float)(2.0 * dstats[C + c + 3]));
#pragma unroll 4
      for (int r = r0 + tl.lane; r < r1; r += tl.lanes) {
        const size_t off = (size_t)r * C + c;
        float4 v = tn_fma4(b, tn_ld4(z + off), a);
        if (dzd) v = v + tn_ld4(dzd + off);
        tn_st4(out + off, v);
        acc = acc + v;
      }
    }
    if (dbias) tn_lane_reduce_atomic(tl, acc, q, dbias, red);
  }
}
extern "C" int tn_stats_bwd(const float* dz_direct, const float* z, const double* dstats, float* out, float* dbias, int R, int C,
                            void* stream) {
  TN_REQUIRE(R > 0 && C > 0 && C % 4 == 0 && z && dstats && out && tn_aligned16(out) && tn_aligned16(z) && (!dz_direct || tn_aligned16(dz_direct)),
             "stats_bwd: need C %% 4 == 0 and aligned tensors (R=%d C=%d)", R, C);
  int rpb = rows_per_block(R);
  tn_launch(stats_bwd_kernel, tn_cdiv(R, rpb), TN_EW_THREADS, 0, stream, dz_direct, z, dstats, out, dbias, R, C, rpb);
  TN_LAUNCH_CHECK("stats_bwd_kernel");
  return TN_OK;
}

// tn_bn_bwd_coef + tn_stats_bwd in one pass: the per-channel coefficients are recomputed by every
// thread for its own four channels (a handful of fp64 operations), block 0 also writes dgamma / dbeta.
__global__ void __launch_bounds__(TN_EW_THREADS) bn_stats_bwd_kernel(const float* __restrict__ dzd, const float* __restrict__ z,
                                                                     const float* __restrict__ dscale, const float* __restrict__ dshift,
                                                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                     const float* __restrict__ gamma, double n, float* __restrict__ out,
                                                                     float* __restrict__ dbias, float* __restrict__ dgamma,
                                                                     float* __restrict__ dbeta, int R, int C, int rpb) {
  tn_grid_dep_sync();
  __shared__ float4 red[TN_EW_THREADS];
  TnTile tl = tn_tile(C);
  const int r0 = blockIdx.x * rpb, r1 = min(R, r0 + rpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    float4 acc = tn_zero4();
    if (tl.active && q < tl.Q) {
      const int c = 4 * q;
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double dsc = dscale[c + i], dsh = dshift[c + i], mu = mean[c + i], r = invstd[c + i], g = gamma[c + i];
        const double t = dsc - mu * dsh;               // dL/d(invstd) / gamma
        const double dvar = -0.5 * g * t * r * r * r;
        const double dmu = -dsh * g * r - 2.0 * mu * dvar;
        av[i] = (float)(dmu / n);
        bv[i] = (float)(2.0 * dvar / n);
        if (blockIdx.x == 0 && tl.lane == 0) {
          dgamma[c + i] = (float)(r * t);
          dbeta[c + i] = (float)dsh;
        }
      }
      const float4 a = make_float4(av[0], av[1], av[2], av[3]);
      const float4 b = make_float4(bv[0], bv[1], bv[2], bv[3]);
#pragma unroll 4
      for (int r = r0 + tl.lane; r < r1; r += tl.lanes) {
        const size_t off = (size_t)r * C + c;
        float4 v = tn_fma4(b, tn_ld4(z + off), a);
        if (dzd) v = v + tn_ld4(dzd + off);
        tn_st4(out + off, v);
        acc = acc + v;
      }
    }
    if (dbias) tn_lane_reduce_atomic(tl, acc, q, dbias, red);
  }
}
extern "C" int tn_bn_stats_bwd(const float* dz_direct, const float* z, const float* dscale, const float* dshift, const float* mean,
                               const float* invstd, const float* gamma, double n, float* out, float* dbias, float* dgamma,
                               float* dbeta, int R, int C, void* stream) {
  TN_REQUIRE(R > 0 && C > 0 && C % 4 == 0 && z && out && tn_aligned16(out) && tn_aligned16(z) && (!dz_direct || tn_aligned16(dz_direct)),
             "bn_stats_bwd: need C %% 4 == 0 and aligned tensors (R=%d C=%d)", R, C);
  TN_REQUIRE(dscale && dshift && mean && invstd && gamma && dgamma && dbeta && n >= 1.0, "bn_stats_bwd: null argument");
  int rpb = rows_per_block(R);
  tn_launch(bn_stats_bwd_kernel, tn_cdiv(R, rpb), TN_EW_THREADS, 0, stream, dz_direct, z, dscale, dshift, mean, invstd, gamma, n,
                                                                                    out, dbias, dgamma, dbeta, R, C, rpb);
  TN_LAUNCH_CHECK("bn_stats_bwd_kernel");
  return TN_OK;
}

// advance the dropout seed state (splitmix64) and publish the new step seed
__global__ void seed_next_kernel(unsigned long long* state, unsigned long long* out) {
  tn_grid_dep_sync();
  unsigned long long z = (*state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  *out = z ^ (z >> 31);
}
extern "C" int tn_seed_next(unsigned long long* state, unsigned long long* out, void* stream) {
  TN_REQUIRE(state && out, "seed_next: null argument");
  tn_launch(seed_next_kernel, 1, 1, 0, stream, state, out);
  TN_LAUNCH_CHECK("seed_next_kernel");
  return TN_OK;
}

// out = dh * (1 - h^2)
__global__ void tanh_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ h, float* __restrict__ out, size_t n) {
  tn_grid_dep_sync();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float hv = h[i];
    out[i] = dh[i] * (1.f - hv * hv);
  }
}
extern "C" int tn_tanh_bwd(const float* dh, const float* h, float* out, long long n, void* stream) {
  TN_REQUIRE(n > 0, "tanh_bwd: empty");
  int blocks = (int)((n + 255) / 256);
  int cap = tn_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  tn_launch(tanh_bwd_kernel, blocks, 256, 0, stream, dh, h, out, (size_t)n);
  TN_LAUNCH_CHECK("tanh_bwd_kernel");
  return TN_OK;
}

// ---------------------------------------------------------------------------
// row L2 normalisation:  y = x / max(||x||, eps)   (eps = 0: plain division)
// ---------------------------------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norms, int B, int E, float eps) {
  tn_grid_dep_sync();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + (size_t)row * E;
  float s = 0.f;
  for (int i = lane; i < E; i += 32) s = fmaf(xr[i], xr[i], s);
  float nrm = sqrtf(tn_warp_sum(s));
  float d = eps > 0.f ? fmaxf(nrm, eps) : nrm;
  for (int i = lane; i < E; i += 32) y[(size_t)row * E + i] = xr[i] / d;
  if (lane == 0 && norms) norms[row] = nrm;
}
// dx = (dy - y * <y, dy>) / max(norm, eps) + dnorm * y      (dnorm optional: grad w.r.t. the returned norm)
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ norms,
                                  const float* __restrict__ dnorm, float* __restrict__ dx, int B, int E, float eps) {
  tn_grid_dep_sync();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* yr = y + (size_t)row * E;
  const float* gr = dy + (size_t)row * E;
  float s = 0.f;
  for (int i = lane; i < E; i += 32) s = fmaf(yr[i], gr[i], s);
  s = tn_warp_sum(s);
  float nrm = norms[row];
  bool clamped = eps > 0.f && nrm < eps;          // F.normalize: denominator is the constant eps
  float d = eps > 0.f ? fmaxf(nrm, eps) : nrm;
  float dn = dnorm ? dnorm[row] : 0.f;
  for (int i = lane; i < E; i += 32) {
    float v = clamped ? gr[i] / d : (gr[i] - yr[i] * s) / d;
    dx[(size_t)row * E + i] = fmaf(dn, yr[i], v);
  }
}
extern "C" int tn_l2norm_fwd(const float* x, float* y, float* norms, int B, int E, float eps, void* stream) {
  TN_REQUIRE(B > 0 && E > 0 && x && y, "l2norm_fwd: bad arguments");
  tn_launch(l2norm_fwd_kernel, tn_cdiv(B, 4), 128, 0, stream, x, y, norms, B, E, eps);
  TN_LAUNCH_CHECK("l2norm_fwd_kernel");
  return TN_OK;
}
extern "C" int tn_l2norm_bwd(const float* dy, const float* y, const float* norms, const float* dnorm, float* dx, int B, int E,
                             float eps, void* stream) {
  TN_REQUIRE(B > 0 && E > 0 && dy && y && norms && dx, "l2norm_bwd: bad arguments");
  tn_launch(l2norm_bwd_kernel, tn_cdiv(B, 4), 128, 0, stream, dy, y, norms, dnorm, dx, B, E, eps);
  TN_LAUNCH_CHECK("l2norm_bwd_kernel");
  return TN_OK;
}

// ---------------------------------------------------------------------------
// Adam over all parameters in one launch (torch.optim.Adam semantics; src/train.py:130-136)
// ---------------------------------------------------------------------------
__global__ void adam_tick_kernel(float* hyper) {
  tn_grid_dep_sync();
  const float step = hyper[5] + 1.f;
  hyper[5] = step;
  hyper[6] = (float)(1.0 - pow((double)hyper[1], (double)step));      // double: 1 - 0.999^1 cancels badly in fp32
  hyper[7] = (float)(1.0 - pow((double)hyper[2], (double)step));
}
__global__ void __launch_bounds__(256) adam_multi_kernel(const tn_adam_job* __restrict__ jobs, const float* __restrict__ hyper) {
  tn_grid_dep_sync();
  const tn_adam_job j = jobs[blockIdx.y];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float step_size = lr / hyper[6], inv_sqrt_bc2 = rsqrtf(hyper[7]);
  const float omb1 = hyper[8], omb2 = hyper[9];        // 1 - beta, rounded once from double by the host (1.f - 0.999f is 1e-5 off)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += (long long)gridDim.x * blockDim.x) {
    const float p = j.p[i];
    const float g = fmaf(wd, p, j.g[i]);
    const float m = fmaf(b1, j.m[i], omb1 * g);
    const float v = fmaf(b2, j.v[i], omb2 * g * g);
    j.m[i] = m;
    j.v[i] = v;
    j.p[i] = p - step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
  }
}
extern "C" int tn_adam_tick(float* hyper_dev, void* stream) {
  TN_REQUIRE(hyper_dev, "adam_tick: null state");
  tn_launch(adam_tick_kernel, 1, 1, 0, stream, hyper_dev);
  TN_LAUNCH_CHECK("adam_tick_kernel");
  return TN_OK;
}
extern "C" int tn_adam_multi(const tn_adam_job* jobs_dev, int njobs, long long max_n, const float* hyper_dev, void* stream) {
  TN_REQUIRE(jobs_dev && hyper_dev && njobs > 0 && njobs <= 65535 && max_n > 0, "adam_multi: bad arguments");
  long long bx = (max_n + 256 * 8 - 1) / (256 * 8);
  if (bx > 64) bx = 64;
  tn_launch(adam_multi_kernel, dim3((unsigned)bx, (unsigned)njobs), 256, 0, stream, jobs_dev, hyper_dev);
  TN_LAUNCH_CHECK("adam_multi_kernel");
  return TN_OK;
}
