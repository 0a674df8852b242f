// fp32 CUDA-core conv-as-GEMM kernels on NWC activations.  These are the exact-fp32
// path for the shapes the tcgen05 kernel (gemm_tc.cu) does not take: the K-tap prolog
// conv (80 -> H, k=3), the decoder / loss-head linears (few rows, 251 classes), odd
// channel counts, and every weight-gradient that is not tile aligned.
//
//   fwd  : Z[r, co]  = bias[co] + sum_{k, ci} X[r + k - pad, ci] * W[co, ci, k]
//   dgrad: same kernel with transpose_w = 1 (reads W[ci', co', K-1-k]): dX from dZ
//   wgrad: dW[co, ci, k] += sum_r dZ[r, co] * X[r + k - pad, ci],  db[co] += sum_r dZ[r, co]
//
// Rows are b*T + t; taps never cross an utterance boundary (zero "same" padding).
// Reference: Conv1dSamePadding (src/modules.py:14-40), nn.Conv1d k=1 in the skip
// connection (src/models.py:452-455), nn.Linear in ASP / decoder / loss heads
// (src/models.py:549-551, 510-513; src/losses.py:30, 70).
#include "common.cuh"
#include <string.h>

#define GM 64   // rows per tile
#define GN 64   // output channels per tile
#define GK 16   // reduction chunk


__global__ void __launch_bounds__(256) conv_gemm_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ Z,
                                                        double* __restrict__ stats, int R, int T, int Ci, int Co, int K,
                                                        int transpose_w, int flags, int kk_per_split, tn_bn_fold bn, int has_bn,
                                                        float* __restrict__ parts, unsigned int* __restrict__ tickets,
                                                        unsigned long long* __restrict__ accum, TnFoldConst fc) {
  tn_grid_dep_sync();
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  __shared__ float red1[16][GN], red2[16][GN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  const int pad = K / 2;
  const int KK = K * Ci;                // reduction length; kk = tap * Ci + ci
  // split-K (skinny problems, e.g. the decoder's [B, 3072] x [3072, 192]): blockIdx.z owns a slice of the reduction and
  // stores its partial tile to parts[split][R][Co]; the last block of an output tile (ticket) adds the partials in split
  // order, so the result does not depend on the order the blocks ran in (flags == 0, no stats).
  const bool splitk = gridDim.z > 1;
  const int kk_begin = blockIdx.z * kk_per_split, kk_end = min(KK, kk_begin + kk_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lk = tid & 15;              // reduction index inside the chunk handled by this thread's loads
  const int lr = tid >> 4;              // row / column group
  for (int kk0 = kk_begin; kk0 < kk_end; kk0 += GK) {
    const int kk = kk0 + lk;
    const int tap = kk < kk_end ? kk / Ci : 0;
    const int ci = kk < kk_end ? kk - tap * Ci : 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      // A tile: rows r0 + lr + 16p
      const int m = lr + 16 * p;
      const int r = r0 + m;
      float v = 0.f;
      if (kk < kk_end && r < R) {
        const int t = r % T;
        const int tt = t + tap - pad;
        if (tt >= 0 && tt < T) v = __ldg(X + (size_t)(r + tap - pad) * Ci + ci);
      }
      As[lk][m] = v;
      // B tile: output channels n0 + lr + 16p
      const int n = n0 + m;
      float wv = 0.f;
      if (kk < kk_end && n < Co) {
        wv = transpose_w ? __ldg(W + ((size_t)ci * Co + n) * K + (K - 1 - tap))   // W is [Ci(red), Co(out), K]
                         : __ldg(W + ((size_t)n * Ci + ci) * K + tap);            // W is [Co(out), Ci(red), K]
      }
      Bs[lk][m] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (splitk) {
    float* mine = parts + (size_t)blockIdx.z * R * Co;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty * 4 + i;
      if (r >= R) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < Co) mine[(size_t)r * Co + n] = acc[i][j];
      }
    }
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      unsigned int* tk = tickets + blockIdx.y * gridDim.x + blockIdx.x;
      const unsigned int t = atomicAdd(tk, 1u);
      s_last = (t == gridDim.z - 1) ? 1u : 0u;
      if (s_last) *tk = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last block of the tile adds the partial tiles in split order.  All loads of a row are issued before the first
    // addition (a dependent load per split costs a global round trip each: 24 splits x 0.7 us)
    const int nz = (int)gridDim.z;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty * 4 + i;
      if (r >= R) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int n = n0 + tx * 4 + j; v[j] = (bias && n < Co) ? __ldg(bias + n) : 0.f; }
      for (int z0 = 0; z0 < nz; z0 += 8) {
        float pv[8][4];
#pragma unroll
        for (int zz = 0; zz < 8; ++zz)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            pv[zz][j] = (z0 + zz < nz && n < Co) ? __ldcg(parts + ((size_t)(z0 + zz) * R + r) * Co + n) : 0.f;
          }
#pragma unroll
        for (int zz = 0; zz < 8; ++zz)
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] += pv[zz][j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < Co) Z[(size_t)r * Co + n] = v[j];
      }
    }
    return;
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Co) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      float* zp = Z + (size_t)r * Co + n;
      if (flags & TN_EPI_TANH) v = tanhf(v);
      if (flags & TN_EPI_ACCUM) v += *zp;
      *zp = v;
      s1[j] += v;
      s2[j] = fmaf(v, v, s2[j]);
    }
  }
  if (stats) {
    // per-channel partial sums of this row tile in a fixed order, added into the fixed-point accumulators (integer atomics:
    // order-independent); the last block of this 64-channel group reads the totals and (has_bn) folds the BatchNorm
#pragma unroll
    for (int j = 0; j < 4; ++j) { red1[ty][tx * 4 + j] = s1[j]; red2[ty][tx * 4 + j] = s2[j]; }
    __syncthreads();
    unsigned int* flag = tn_fix_flag(accum, Co, (int)blockIdx.y);
    if (tid < 2 * GN) {
      const int which = tid / GN, ch = tid - which * GN;
      const int n = n0 + ch;
      if (n < Co) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < 16; ++y) a += which ? red2[y][ch] : red1[y][ch];
        tn_fix_add(accum, Co, which, n, a, flag);
      }
    }
    const int nC = min(GN, Co - n0);
    tn_stats_finish(has_bn ? &bn : nullptr, stats, accum, Co, n0, nC, tickets + blockIdx.y, flag, gridDim.x, blockIdx.y == 0, fc);
  }
}

// ---------------------------------------------------------------------------
// Skinny linear layers (few rows): the decoder's Linear(2D, E) on [B, 3072], the loss heads' Linear(E, classes) and their
// data gradients (src/models.py:510-513, src/losses.py:30, 70).  conv_gemm_kernel above walks the reduction 16 scalars at a
// time with one global round trip exposed per chunk (45 us for [64, 3072] x [3072, 192]: latency, not work).  Here a block
// owns 64 rows x 32 output channels x a slice of the reduction; chunks of 32 reduction elements are fetched with 16-byte
// loads one chunk AHEAD of the arithmetic (registers), and a thread owns 8 rows x 1 channel.  Split-K partial tiles are
// added by the last block of a tile in split order (ticket), so the result does not depend on the order the blocks ran in.
//   Z[r, n] = bias[n] + sum_k X[r, k] Wt(k, n);   TRANS = 0: W is [N, KK] (forward), TRANS = 1: W is [KK, N] (data gradient)
// ---------------------------------------------------------------------------
#define SK_ROWS 64
#define SK_COLS 32
#define SK_KC 32
template <int TRANS, int VEC>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ Z, int R, int KK, int N,
                                                          int k_per_split, float* __restrict__ parts, unsigned int* __restrict__ tickets) {
  tn_grid_dep_sync();
  __shared__ __align__(16) float Xs[SK_KC][SK_ROWS + 4];     // [k][row]
  __shared__ __align__(16) float Ws[SK_KC][SK_COLS + 4];     // [k][n]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * SK_COLS, r0 = blockIdx.z * SK_ROWS;
  const int k_begin = blockIdx.y * k_per_split, k_end = min(KK, k_begin + k_per_split);
  const int tx = tid & 31, ty = tid >> 5;                    // this thread: channel n0 + tx, rows r0 + 8 ty .. + 7
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  // loader roles.  X chunk: 64 rows x 8 quads -> two quads per thread; W chunk: 32 x 8 quads -> one per thread
  const int xr = tid >> 3, xq = tid & 7;                     // rows xr and xr + 32, reduction quad xq
  const int wa = tid >> 3, wq = tid & 7;                     // TRANS 0: channel wa, reduction quad wq; TRANS 1: reduction row wa, channel quad wq
  float4 xv[2], wv;
  auto fetch = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + xr + 32 * h, k = k0 + 4 * xq;
      float4 v = tn_zero4();
      if (r < R) {
        const float* src = X + (size_t)r * KK + k;
        if (VEC && k + 3 < k_end) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (k < k_end) v.x = __ldg(src);
          if (k + 1 < k_end) v.y = __ldg(src + 1);
          if (k + 2 < k_end) v.z = __ldg(src + 2);
          if (k + 3 < k_end) v.w = __ldg(src + 3);
        }
      }
      xv[h] = v;
    }
    float4 v = tn_zero4();
    if (TRANS == 0) {
      const int n = n0 + wa, k = k0 + 4 * wq;
      if (n < N) {
        const float* src = W + (size_t)n * KK + k;
        if (VEC && k + 3 < k_end) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (k < k_end) v.x = __ldg(src);
          if (k + 1 < k_end) v.y = __ldg(src + 1);
          if (k + 2 < k_end) v.z = __ldg(src + 2);
          if (k + 3 < k_end) v.w = __ldg(src + 3);
        }
      }
    } else {
      const int k = k0 + wa, n = n0 + 4 * wq;
      if (k < k_end) {
        const float* src = W + (size_t)k * N + n;
        if (VEC && n + 3 < N) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (n < N) v.x = __ldg(src);
          if (n + 1 < N) v.y = __ldg(src + 1);
          if (n + 2 < N) v.z = __ldg(src + 2);
          if (n + 3 < N) v.w = __ldg(src + 3);
        }
      }
    }
    wv = v;
  };
  auto stage = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = xr + 32 * h;
      Xs[4 * xq + 0][m] = xv[h].x; Xs[4 * xq + 1][m] = xv[h].y; Xs[4 * xq + 2][m] = xv[h].z; Xs[4 * xq + 3][m] = xv[h].w;
    }
    if (TRANS == 0) {
      Ws[4 * wq + 0][wa] = wv.x; Ws[4 * wq + 1][wa] = wv.y; Ws[4 * wq + 2][wa] = wv.z; Ws[4 * wq + 3][wa] = wv.w;
    } else {
      *reinterpret_cast<float4*>(&Ws[wa][4 * wq]) = wv;
    }
  };
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += SK_KC) {
    stage();
    __syncthreads();
    if (k0 + SK_KC < k_end) fetch(k0 + SK_KC);               // next chunk in flight while this one is multiplied
#pragma unroll
    for (int k = 0; k < SK_KC; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Xs[k][8 * ty]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Xs[k][8 * ty + 4]);
      const float b = Ws[k][tx];
      acc[0] = fmaf(a0.x, b, acc[0]); acc[1] = fmaf(a0.y, b, acc[1]); acc[2] = fmaf(a0.z, b, acc[2]); acc[3] = fmaf(a0.w, b, acc[3]);
      acc[4] = fmaf(a1.x, b, acc[4]); acc[5] = fmaf(a1.y, b, acc[5]); acc[6] = fmaf(a1.z, b, acc[6]); acc[7] = fmaf(a1.w, b, acc[7]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  const float bv = (bias && n < N) ? __ldg(bias + n) : 0.f;
  if (gridDim.y == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + 8 * ty + i;
      if (r < R && n < N) Z[(size_t)r * N + n] = acc[i] + bv;
    }
    return;
  }
  float* mine = parts + (size_t)blockIdx.y * R * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 8 * ty + i;
    if (r < R && n < N) mine[(size_t)r * N + n] = acc[i];
  }
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int* tk = tickets + blockIdx.z * gridDim.x + blockIdx.x;
    const unsigned int t = atomicAdd(tk, 1u);
    s_last = (t == gridDim.y - 1) ? 1u : 0u;
    if (s_last) *tk = 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last block of the tile adds the partial tiles in split order; eight splits' loads are issued before their additions
  const int nz = (int)gridDim.y;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = bv;
  for (int z0 = 0; z0 < nz; z0 += 4) {
    float pv[4][8];
#pragma unroll
    for (int zz = 0; zz < 4; ++zz)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + 8 * ty + i;
        pv[zz][i] = (z0 + zz < nz && r < R && n < N) ? __ldcg(parts + ((size_t)(z0 + zz) * R + r) * N + n) : 0.f;
      }
#pragma unroll
    for (int zz = 0; zz < 4; ++zz)
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += pv[zz][i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 8 * ty + i;
    if (r < R && n < N) Z[(size_t)r * N + n] = v[i];
  }
}

// the skinny kernel takes 1-tap layers with few rows and a plain epilogue (statistics, if wanted, come from tn_colstats)
static bool skinny_ok(long long R, int K, int flags) { return K == 1 && R <= 512 && flags == 0; }
static void skinny_plan(long long R, int KK, int N, int* splits_out, int* kps_out) {
  const long long tiles = (long long)tn_cdiv(N, SK_COLS) * tn_cdiv(R, SK_ROWS);
  int splits = (int)((2ll * tn_num_sms() + tiles - 1) / tiles);
  if (splits > KK / (4 * SK_KC)) splits = KK / (4 * SK_KC);          // at least four chunks per split
  if (tiles > TN_TICKETS) splits = 1;
  if (splits < 1) splits = 1;
  int kps = ((KK + splits - 1) / splits + SK_KC - 1) / SK_KC * SK_KC;
  *splits_out = (KK + kps - 1) / kps;
  *kps_out = kps;
}

// dW[co, ci, k] += sum_r dZ[r, co] * X[r + k - pad, ci]; rows split over blockIdx.z
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ dZ, const float* __restrict__ X,
                                                         float* __restrict__ dW, float* __restrict__ dbias, int R, int T,
                                                         int Ci, int Co, int K, int rows_per_split) {
  tn_grid_dep_sync();
  __shared__ float As[GK][GM + 4];   // As[row][co]
  __shared__ float Bs[GK][GN + 4];   // Bs[row][kk]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * GM;    // co tile
  const int n0 = blockIdx.y * GN;    // kk tile (kk = tap * Ci + ci)
  const int pad = K / 2;
  const int KK = K * Ci;
  const int ra = blockIdx.z * rows_per_split;
  const int rb = min(R, ra + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};

  const int lc = tid & 63;           // column (co or kk) handled by this thread's loads
  const int lr = tid >> 6;           // 0..3
  const int co_l = m0 + lc;
  const int kk_l = n0 + lc;
  const int tap = kk_l < KK ? kk_l / Ci : 0;
  const int ci = kk_l < KK ? kk_l - tap * Ci : 0;
  for (int rr = ra; rr < rb; rr += GK) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int k = lr + 4 * p;
      const int r = rr + k;
      float a = 0.f, b = 0.f;
      if (r < rb) {
        if (co_l < Co) a = __ldg(dZ + (size_t)r * Co + co_l);
        if (kk_l < KK) {
          const int tt = (r % T) + tap - pad;
          if (tt >= 0 && tt < T) b = __ldg(X + (size_t)(r + tap - pad) * Ci + ci);
        }
      }
      As[k][lc] = a;
      Bs[k][lc] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = m0 + ty * 4 + i;
    if (co >= Co) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = n0 + tx * 4 + j;
      if (kk >= KK) continue;
      const int tp = kk / Ci, c2 = kk - tp * Ci;
      atomicAdd(dW + ((size_t)co * Ci + c2) * K + tp, acc[i][j]);
    }
    if (dbias && blockIdx.y == 0 && tx == 0) atomicAdd(dbias + co, bsum[i]);
  }
}

// split-K plan shared by the launcher and the scratch query
static void conv_gemm_plan(long long R, int Ci, int Co, int K, int flags, int* splits_out, int* kps_out) {
  const long long tiles = (long long)tn_cdiv(R, GM) * tn_cdiv(Co, GN);
  const long long KK = (long long)K * Ci;
  int splits = 1;
  if (flags == 0 && KK >= 1024 && tiles * 4 <= tn_num_sms() && tiles <= TN_TICKETS) {
    splits = tn_num_sms() / (int)tiles;
    if (splits > KK / 128) splits = (int)(KK / 128);
    if (splits < 1) splits = 1;
  }
  int kps = (int)(((KK + splits - 1) / splits + GK - 1) / GK * GK);
  *splits_out = (int)((KK + kps - 1) / kps);
  *kps_out = kps;
}

extern "C" long long tn_conv_gemm_simt_scratch_floats(int B, int T, int Ci, int Co, int K, int flags) {
  if (B <= 0 || T <= 0 || Ci <= 0 || Co <= 0 || K <= 0) return 0;
  const long long R = (long long)B * T;
  int splits, kps;
  if (skinny_ok(R, K, flags)) skinny_plan(R, Ci, Co, &splits, &kps);
  else conv_gemm_plan(R, Ci, Co, K, flags, &splits, &kps);
  return splits > 1 ? (long long)splits * R * Co : 0;
}

static int conv_gemm_launch(const float* X, const float* W, const float* bias, float* Z, double* stats, const tn_bn_fold* bn,
                            int B, int T, int Ci, int Co, int K, int transpose_w, int flags, const tn_scratch* scratch, void* stream) {
  TN_REQUIRE(B > 0 && T > 0 && Ci > 0 && Co > 0 && K > 0 && (K & 1), "conv_gemm: bad shape B=%d T=%d Ci=%d Co=%d K=%d (odd K only)", B, T, Ci, Co, K);
  TN_REQUIRE(X && W && Z, "conv_gemm: null tensor");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31) && tn_cdiv(Co, GN) <= 65535, "conv_gemm: shape too large");
  dim3 grid(tn_cdiv(R, GM), tn_cdiv(Co, GN));
  int splits, kps;
  const bool skinny = skinny_ok(R, K, flags);
  if (skinny) skinny_plan(R, Ci, Co, &splits, &kps);
  else conv_gemm_plan(R, Ci, Co, K, flags, &splits, &kps);
  grid.z = splits;
  const long long need = tn_conv_gemm_simt_scratch_floats(B, T, Ci, Co, K, flags);
  if (need > 0)
    TN_REQUIRE(scratch && scratch->parts && scratch->tickets && scratch->parts_floats >= need,
               "conv_gemm: split-K needs a tn_scratch with %lld floats (tn_conv_gemm_simt_scratch_floats) and the ticket array", need);
  if (skinny) {
    // X rows are Ci floats apart, W rows Ci (forward) or Co (data gradient): 16-byte loads need those multiples of 4
    const bool vec = (Ci % 4 == 0) && (transpose_w ? Co % 4 == 0 : true) && tn_aligned16(X) && tn_aligned16(W);
    dim3 g(tn_cdiv(Co, SK_COLS), splits, tn_cdiv(R, SK_ROWS));
    float* parts = scratch ? scratch->parts : (float*)nullptr;
    unsigned int* tickets = scratch ? scratch->tickets : (unsigned int*)nullptr;
    if (transpose_w) {
      if (vec) tn_launch(skinny_gemm_kernel<1, 1>, g, 256, 0, stream, X, W, bias, Z, (int)R, Ci, Co, kps, parts, tickets);
      else tn_launch(skinny_gemm_kernel<1, 0>, g, 256, 0, stream, X, W, bias, Z, (int)R, Ci, Co, kps, parts, tickets);
    } else {
      if (vec) tn_launch(skinny_gemm_kernel<0, 1>, g, 256, 0, stream, X, W, bias, Z, (int)R, Ci, Co, kps, parts, tickets);
      else tn_launch(skinny_gemm_kernel<0, 0>, g, 256, 0, stream, X, W, bias, Z, (int)R, Ci, Co, kps, parts, tickets);
    }
    TN_LAUNCH_CHECK("skinny_gemm_kernel");
    if (stats) {
      int rc = tn_colstats(Z, stats, (int)R, Co, stream);
      if (rc != TN_OK) return rc;
      if (bn) return tn_bn_finalize(stats, bn->n, bn->gamma, bn->beta, bn->running_mean, bn->running_var, bn->num_batches_tracked,
                                    bn->momentum, bn->eps, 1, bn->scale, bn->shift, bn->mean, bn->invstd, Co, stream);
    }
    return TN_OK;
  }
  if (stats && splits == 1) {
    TN_REQUIRE(scratch && scratch->accum && scratch->tickets && scratch->accum_words >= TN_ACCUM_WORDS(Co),
               "conv_gemm: statistics need a tn_scratch with TN_ACCUM_WORDS(Co) zeroed accumulator words and the ticket array");
    TN_REQUIRE((int)grid.y <= TN_TICKETS, "conv_gemm: statistics of more than %d channels are not supported", TN_TICKETS * GN);
  }
  tn_bn_fold f;
  memset(&f, 0, sizeof(f));
  const int fuse_bn = (bn && splits == 1) ? 1 : 0;
  if (fuse_bn) f = *bn;
  tn_launch(conv_gemm_kernel, grid, 256, 0, stream, X, W, bias, Z, splits > 1 ? (double*)nullptr : stats, (int)R, T, Ci, Co, K,
            transpose_w, flags, kps, f, fuse_bn, scratch ? scratch->parts : (float*)nullptr, scratch ? scratch->tickets : (unsigned int*)nullptr,
            scratch ? scratch->accum : (unsigned long long*)nullptr, tn_fold_const(fuse_bn ? f.n : 1.0));
  TN_LAUNCH_CHECK("conv_gemm_kernel");
  if (splits > 1 && stats) {                       // split-K: statistics (and the fold) from the finished tensor
    int rc = tn_colstats(Z, stats, (int)R, Co, stream);
    if (rc != TN_OK) return rc;
    if (bn) return tn_bn_finalize(stats, bn->n, bn->gamma, bn->beta, bn->running_mean, bn->running_var, bn->num_batches_tracked,
                                  bn->momentum, bn->eps, 1, bn->scale, bn->shift, bn->mean, bn->invstd, Co, stream);
  }
  return TN_OK;
}

extern "C" int tn_conv_gemm_simt(const float* X, const float* W, const float* bias, float* Z, double* stats, int B, int T,
                                 int Ci, int Co, int K, int transpose_w, int flags, const tn_scratch* scratch, void* stream) {
  return conv_gemm_launch(X, W, bias, Z, stats, nullptr, B, T, Ci, Co, K, transpose_w, flags, scratch, stream);
}

extern "C" int tn_conv_gemm_simt_bn(const float* X, const float* W, const float* bias, float* Z, double* stats,
                                    const tn_bn_fold* bn, int B, int T, int Ci, int Co, int K, int flags, const tn_scratch* scratch,
                                    void* stream) {
  TN_REQUIRE(bn && stats, "conv_gemm_simt_bn: the fold needs the statistics buffer");
  TN_REQUIRE(bn->gamma && bn->beta && bn->scale && bn->shift && bn->mean && bn->invstd && bn->n >= 1.0,
             "conv_gemm_simt_bn: incomplete tn_bn_fold");
  return conv_gemm_launch(X, W, bias, Z, stats, bn, B, T, Ci, Co, K, 0, flags, scratch, stream);
}

extern "C" int tn_conv_wgrad_simt(const float* dZ, const float* X, float* dW, float* dbias, int B, int T, int Ci, int Co, int K,
                                  void* stream) {
  TN_REQUIRE(B > 0 && T > 0 && Ci > 0 && Co > 0 && K > 0 && (K & 1), "conv_wgrad: bad shape B=%d T=%d Ci=%d Co=%d K=%d (odd K only)", B, T, Ci, Co, K);
  TN_REQUIRE(dZ && X && dW, "conv_wgrad: null tensor");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "conv_wgrad: shape too large");
  int tiles = tn_cdiv(Co, GM) * tn_cdiv((long long)K * Ci, GN);
  long long want = ((long long)tn_num_sms() * 4 + tiles - 1) / tiles;     // ~4 blocks per SM in total
  long long max_split = (R + GK * 4 - 1) / (GK * 4);
  long long split = want < 1 ? 1 : (want > max_split ? max_split : want);
  if (split > 65535) split = 65535;
  int rps = (int)((R + split - 1) / split);
  rps = ((rps + GK - 1) / GK) * GK;
  split = (R + rps - 1) / rps;
  dim3 grid(tn_cdiv(Co, GM), tn_cdiv((long long)K * Ci, GN), (unsigned)split);
  TN_REQUIRE(grid.y <= 65535, "conv_wgrad: K*Ci too large");
  tn_launch(conv_wgrad_kernel, grid, 256, 0, stream, dZ, X, dW, dbias, (int)R, T, Ci, Co, K, rps);
  TN_LAUNCH_CHECK("conv_wgrad_kernel");
  return TN_OK;
}
