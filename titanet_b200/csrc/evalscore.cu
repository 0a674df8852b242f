// Speaker-verification scoring on the device: the consumer of the eval-mode forward.
//   tn_cosine_scores : F.cosine_similarity of every ordered pair of embeddings + the same-speaker labels, in the
//                      itertools.product order of SpeakerDataset.get_sample_pairs (src/datasets.py:165-183, learn.test
//                      src/learn.py:428-439)
//   tn_det_metrics   : utils.compute_error_rates / compute_mindcf / compute_eer (src/utils.py:294-367): stable ascending sort of
//                      the trial scores (bitonic network on unique 64-bit keys), prefix counts of targets / non-targets, the
//                      minimum detection cost and the crossing of the ROC polyline with the anti-diagonal.
// Integer work (sort order, counts) is exact; the rates and costs are fp64 with the reference's operation order (no FMA
// contraction), so minDCF is bit-identical to the Python loops and EER agrees with brentq to its tolerance.
#include "common.cuh"

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------------------------------
// all-pairs cosine: 32 x 32 output tile per block, D walked in chunks of 32 through shared memory
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cosine_pairs_kernel(const float* __restrict__ E, const long long* __restrict__ spk,
                                                           float* __restrict__ S, unsigned char* __restrict__ lab, int N, int D,
                                                           float eps) {
  tn_grid_dep_sync();
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty 0..7
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, na[4] = {0.f, 0.f, 0.f, 0.f}, nb = 0.f;
  for (int d0 = 0; d0 < D; d0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = ty + 8 * r, d = d0 + tx;
      sa[row][tx] = (i0 + row < N && d < D) ? E[(size_t)(i0 + row) * D + d] : 0.f;
      sb[row][tx] = (j0 + row < N && d < D) ? E[(size_t)(j0 + row) * D + d] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < 32; ++d) {
      const float b = sb[tx][d];
      nb = fmaf(b, b, nb);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float a = sa[ty + 8 * r][d];
        acc[r] = fmaf(a, b, acc[r]);
        na[r] = fmaf(a, a, na[r]);
      }
    }
    __syncthreads();
  }
  const int j = j0 + tx;
  if (j >= N) return;
  const float inv_b = 1.f / fmaxf(sqrtf(nb), eps);
  const long long sj = spk ? spk[j] : 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r;
    if (i >= N) continue;
    const float inv_a = 1.f / fmaxf(sqrtf(na[r]), eps);
    S[(size_t)i * N + j] = acc[r] * inv_a * inv_b;
    if (lab) lab[(size_t)i * N + j] = (spk[i] == sj) ? 1 : 0;
  }
}

extern "C" int tn_cosine_scores(const float* E, const long long* speakers, float* scores, unsigned char* labels, int N, int D,
                                float eps, void* stream) {
  TN_REQUIRE(E && scores, "tn_cosine_scores: null pointer");
  TN_REQUIRE(N >= 1 && D >= 1, "tn_cosine_scores: bad shape N=%d D=%d", N, D);
  TN_REQUIRE(labels == nullptr || speakers != nullptr, "tn_cosine_scores: labels need speakers");
  dim3 grid(tn_cdiv(N, 32), tn_cdiv(N, 32));
  tn_launch(cosine_pairs_kernel, grid, dim3(256), 0, stream, E, speakers, scores, labels, N, D, eps);
  TN_LAUNCH_CHECK("cosine_pairs_kernel");
  return TN_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// detection metrics
// ---------------------------------------------------------------------------------------------------------------------
#define DET_CHUNK 2048          // elements per block of the shared-memory sort / scan kernels (1024 threads x 2)

struct DetCounters {            // lives at the start of the workspace
  u64 n_pos;                    // number of target trials
  u64 min_cost_bits;            // bits of the smallest detection cost (non-negative doubles order like integers)
  unsigned int g_star;          // largest threshold-group start whose ROC point is on / above the anti-diagonal
  unsigned int e_star;          // smallest threshold-group end whose successor point is below it
};

static inline long long det_pad(long long n) {
  long long p = DET_CHUNK;
  while (p < n) p <<= 1;
  return p;
}
static inline long long det_align(long long b) { return (b + 255) & ~255ll; }
static void det_layout(long long n, long long* off_keys, long long* off_cum, long long* off_blocks, long long* total) {
  const long long np = det_pad(n);
  long long o = det_align(sizeof(DetCounters));
  *off_keys = o;   o += det_align(np * 8);
  *off_cum = o;    o += det_align(np * 4);
  *off_blocks = o; o += det_align((np / DET_CHUNK + 1) * 4);
  *total = o;
}

// key = (order-preserving image of the fp32 score) << 32 | trial index: unique, so the (unstable) bitonic network
// reproduces Python's stable sorted(..., key=score); padding keys sort last
__global__ void __launch_bounds__(256) det_keys_kernel(const float* __restrict__ scores, u64* __restrict__ keys, long long n,
                                                       long long n_pad, DetCounters* __restrict__ cnt) {
  tn_grid_dep_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { cnt->n_pos = 0; cnt->min_cost_bits = ~0ull; cnt->g_star = 0u; cnt->e_star = (unsigned int)(n - 1); }
  if (i >= n_pad) return;
  u64 k = ~0ull;
  if (i < n) {
    float s = scores[i];
    if (s == 0.f) s = 0.f;                                   // -0.0 ties with +0.0 in Python
    unsigned int u = __float_as_uint(s);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
    k = ((u64)u << 32) | (u64)i;
  }
  keys[i] = k;
}

__device__ __forceinline__ void det_cswap(u64* s, unsigned int i, unsigned int j, bool asc) {
  const u64 a = s[i], b = s[i + j];
  if ((a > b) == asc) { s[i] = b; s[i + j] = a; }
}

// FULL: sort every 2048-chunk completely (stages k = 2..2048, direction from the GLOBAL index so that later stages can merge);
// otherwise finish stage k_global: steps j = 1024..1 inside the chunk
template <bool FULL>
__global__ void __launch_bounds__(1024) det_sort_smem_kernel(u64* __restrict__ keys, u64 k_global) {
  tn_grid_dep_sync();
  __shared__ u64 s[DET_CHUNK];
  const unsigned int tid = threadIdx.x;
  const u64 base = (u64)blockIdx.x * DET_CHUNK;
  s[tid] = keys[base + tid];
  s[tid + 1024] = keys[base + tid + 1024];
  __syncthreads();
  if (FULL) {
    for (unsigned int k = 2; k <= DET_CHUNK; k <<= 1)
      for (unsigned int j = k >> 1; j > 0; j >>= 1) {
        const unsigned int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
        det_cswap(s, i, j, ((base + i) & k) == 0);
        __syncthreads();
      }
  } else {
    for (unsigned int j = DET_CHUNK / 2; j > 0; j >>= 1) {
      const unsigned int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
      det_cswap(s, i, j, ((base + i) & k_global) == 0);
      __syncthreads();
    }
  }
  keys[base + tid] = s[tid];
  keys[base + tid + 1024] = s[tid + 1024];
}

// one compare-exchange step (distance j >= 2048) of stage k over the whole array
__global__ void __launch_bounds__(256) det_sort_global_kernel(u64* __restrict__ keys, u64 j, u64 k, u64 half) {
  tn_grid_dep_sync();
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= half) return;
  const u64 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  const u64 a = keys[i], b = keys[i + j];
  if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[i + j] = a; }
}

__device__ __forceinline__ int det_label(const unsigned char* __restrict__ labels, u64 key, long long pos, long long n) {
  return pos < n ? (labels[(unsigned int)key] != 0) : 0;
}

// targets per 2048-chunk of the sorted order
__global__ void __launch_bounds__(1024) det_count_kernel(const u64* __restrict__ keys, const unsigned char* __restrict__ labels,
                                                         long long n, int* __restrict__ block_pos, DetCounters* __restrict__ cnt) {
  tn_grid_dep_sync();
  __shared__ int warp_tot[32];
  const long long p0 = (long long)blockIdx.x * DET_CHUNK + 2 * threadIdx.x;
  const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(keys + p0);
  int c = det_label(labels, kk.x, p0, n) + det_label(labels, kk.y, p0 + 1, n);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_tot[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      block_pos[blockIdx.x] = v;
      if (v) atomicAdd(&cnt->n_pos, (u64)v);
    }
  }
}

// in-place exclusive scan of the per-chunk counts (one block, running carry)
__global__ void __launch_bounds__(1024) det_scan_blocks_kernel(int* __restrict__ block_pos, int nb) {
  tn_grid_dep_sync();
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int v = b < nb ? block_pos[b] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (threadIdx.x >= o) winc += t;
      }
      warp_tot[threadIdx.x] = winc - w;                       // exclusive warp offsets
    }
    __syncthreads();
    const int excl = carry + warp_tot[threadIdx.x >> 5] + inc - v;
    if (b < nb) block_pos[b] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

// utils.compute_error_rates + the cost of compute_mindcf at every threshold, and the bracketing of the EER crossing
__global__ void __launch_bounds__(1024) det_eval_kernel(const u64* __restrict__ keys, const unsigned char* __restrict__ labels,
                                                        long long n, const int* __restrict__ block_off, int* __restrict__ cum,
                                                        DetCounters* __restrict__ cnt, double p_target, double c_fa,
                                                        double c_miss, double eps, double* __restrict__ fnrs,
                                                        double* __restrict__ fprs) {
  tn_grid_dep_sync();
  __shared__ int warp_tot[32];
  __shared__ u64 warp_min[32];
  const long long p0 = (long long)blockIdx.x * DET_CHUNK + 2 * threadIdx.x;
  const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(keys + p0);
  const int l0 = det_label(labels, kk.x, p0, n), l1 = det_label(labels, kk.y, p0 + 1, n);
  const int v = l0 + l1;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = warp_tot[threadIdx.x], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (threadIdx.x >= o) winc += t;
    }
    warp_tot[threadIdx.x] = winc - w;
  }
  __syncthreads();
  const long long before = (long long)block_off[blockIdx.x] + warp_tot[threadIdx.x >> 5] + inc - v;   // targets before p0
  const long long P = (long long)cnt->n_pos, Nn = n - P;
  const double dP = (double)P + eps, dN = (double)Nn + eps;
  const double w_miss = p_target, w_fa = 1.0 - p_target;
  const unsigned int hi_prev = p0 > 0 ? (unsigned int)(__ldg(keys + p0 - 1) >> 32) : 0u;
  const unsigned int hi_next = p0 + 2 < n ? (unsigned int)(__ldg(keys + p0 + 2) >> 32) : 0u;
  const unsigned int h0 = (unsigned int)(kk.x >> 32), h1 = (unsigned int)(kk.y >> 32);
  u64 best = ~0ull;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const long long pos = p0 + e;
    if (pos >= n) break;
    const int lab = e ? l1 : l0;
    const long long pos_incl = before + l0 + (e ? l1 : 0);        // targets among sorted[0..pos]
    const long long neg_incl = pos + 1 - pos_incl;
    cum[pos] = (int)pos_incl;
    // fnrs[i] = cum targets / (P + eps); fprs[i] = 1 - cum non-targets / (N + eps)        (src/utils.py:330-348)
    const double fnr = __ddiv_rn((double)pos_incl, dP);
    const double fpr = __dsub_rn(1.0, __ddiv_rn((double)neg_incl, dN));
    if (fnrs) { fnrs[pos] = fnr; fprs[pos] = fpr; }
    // c_det = c_miss * fnr * p_target + c_fa * fpr * (1 - p_target)                        (src/utils.py:362)
    const double c = __dadd_rn(__dmul_rn(__dmul_rn(c_miss, fnr), w_miss), __dmul_rn(__dmul_rn(c_fa, fpr), w_fa));
    const u64 cb = (u64)__double_as_longlong(c);
    best = cb < best ? cb : best;
    // ROC bracketing (sklearn.metrics.roc_curve thresholds = distinct scores, descending): the point of a threshold
    // group counts every trial from its first sorted position on; the point before it counts those after its last one
    const unsigned int me = e ? h1 : h0;
    const bool start = (pos == 0) || ((e ? h0 : hi_prev) != me);
    const bool end = (pos == n - 1) || ((e ? hi_next : h1) != me);
    if (P > 0 && Nn > 0) {
      if (start) {
        const long long tp = P - (pos_incl - lab), fp = Nn - (neg_incl - (1 - lab));
        if (fp * P + tp * Nn >= P * Nn) atomicMax(&cnt->g_star, (unsigned int)pos);
      }
      if (end) {
        const long long tp = P - pos_incl, fp = Nn - neg_incl;
        if (fp * P + tp * Nn < P * Nn) atomicMin(&cnt->e_star, (unsigned int)pos);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const u64 t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0) warp_min[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    best = warp_min[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u64 t = __shfl_xor_sync(0xffffffffu, best, o);
      best = t < best ? t : best;
    }
    if (threadIdx.x == 0 && best != ~0ull) atomicMin(&cnt->min_cost_bits, best);
  }
}

// out = {min c_det, EER, targets, non-targets, fpr0, tpr0, fpr1, tpr1}: (fpr0, tpr0) -> (fpr1, tpr1) is the ROC segment
// that crosses tpr = 1 - fpr; EER is the abscissa of the crossing (what brentq finds on interp1d(fpr, tpr), src/utils.py:298-299)
__global__ void det_final_kernel(const int* __restrict__ cum, const DetCounters* __restrict__ cnt, long long n,
                                 double* __restrict__ out) {
  tn_grid_dep_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long P = (long long)cnt->n_pos, Nn = n - P;
  out[0] = __longlong_as_double((long long)cnt->min_cost_bits);
  out[2] = (double)P;
  out[3] = (double)Nn;
  if (P == 0 || Nn == 0) {
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    out[1] = nan; out[4] = nan; out[5] = nan; out[6] = nan; out[7] = nan;
    return;
  }
  const long long g = cnt->g_star, e = cnt->e_star;
  const long long pos_before_g = g > 0 ? cum[g - 1] : 0;
  const long long tp1 = P - pos_before_g, fp1 = Nn - (g - pos_before_g);
  const long long tp0 = P - cum[e], fp0 = Nn - (e + 1 - cum[e]);
  const double x0 = (double)fp0 / (double)Nn, y0 = (double)tp0 / (double)P;
  const double x1 = (double)fp1 / (double)Nn, y1 = (double)tp1 / (double)P;
  const double t = (1.0 - x0 - y0) / ((x1 - x0) + (y1 - y0));
  out[1] = x0 + t * (x1 - x0);
  out[4] = x0; out[5] = y0; out[6] = x1; out[7] = y1;
}

extern "C" int tn_det_workspace_bytes(long long n, long long* bytes_out) {
  TN_REQUIRE(bytes_out != nullptr, "tn_det_workspace_bytes: null pointer");
  TN_REQUIRE(n >= 1 && n < (1ll << 31), "tn_det_workspace_bytes: trial count %lld outside [1, 2^31)", n);
  long long ok, oc, ob, total;
  det_layout(n, &ok, &oc, &ob, &total);
  *bytes_out = total;
  return TN_OK;
}

extern "C" int tn_det_metrics(const float* scores, const unsigned char* labels, long long n, double p_target, double c_fa,
                              double c_miss, double eps, void* workspace, long long workspace_bytes, double* out8,
                              double* fnrs, double* fprs, unsigned long long* sorted_keys, void* stream) {
  TN_REQUIRE(scores && labels && workspace && out8, "tn_det_metrics: null pointer");
  TN_REQUIRE(n >= 1 && n < (1ll << 31), "tn_det_metrics: trial count %lld outside [1, 2^31)", n);
  TN_REQUIRE((fnrs == nullptr) == (fprs == nullptr), "tn_det_metrics: fnrs and fprs go together");
  TN_REQUIRE(c_fa >= 0.0 && c_miss >= 0.0 && p_target >= 0.0 && p_target <= 1.0 && eps >= 0.0,
             "tn_det_metrics: costs / prior must be non-negative (p_target in [0, 1])");
  TN_REQUIRE((((uintptr_t)workspace) & 255u) == 0, "tn_det_metrics: workspace must be 256-byte aligned");
  long long ok, oc, ob, total;
  det_layout(n, &ok, &oc, &ob, &total);
  TN_REQUIRE(workspace_bytes >= total, "tn_det_metrics: workspace of %lld bytes, need %lld", workspace_bytes, total);
  char* ws = (char*)workspace;
  DetCounters* cnt = (DetCounters*)ws;
  u64* keys = (u64*)(ws + ok);
  int* cum = (int*)(ws + oc);
  int* blocks = (int*)(ws + ob);
  const long long np = det_pad(n);
  const int nchunks = (int)(np / DET_CHUNK);

  tn_launch(det_keys_kernel, dim3(tn_cdiv(np, 256)), dim3(256), 0, stream, scores, keys, n, np, cnt);
  TN_LAUNCH_CHECK("det_keys_kernel");
  tn_launch(det_sort_smem_kernel<true>, dim3(nchunks), dim3(1024), 0, stream, keys, (u64)0);
  TN_LAUNCH_CHECK("det_sort_smem_kernel<full>");
  for (u64 k = 2 * DET_CHUNK; k <= (u64)np; k <<= 1) {
    for (u64 j = k >> 1; j >= DET_CHUNK; j >>= 1) {
      tn_launch(det_sort_global_kernel, dim3(tn_cdiv(np / 2, 256)), dim3(256), 0, stream, keys, j, k, (u64)(np / 2));
      TN_LAUNCH_CHECK("det_sort_global_kernel");
    }
    tn_launch(det_sort_smem_kernel<false>, dim3(nchunks), dim3(1024), 0, stream, keys, k);
    TN_LAUNCH_CHECK("det_sort_smem_kernel<merge>");
  }
  tn_launch(det_count_kernel, dim3(nchunks), dim3(1024), 0, stream, (const u64*)keys, labels, n, blocks, cnt);
  TN_LAUNCH_CHECK("det_count_kernel");
  tn_launch(det_scan_blocks_kernel, dim3(1), dim3(1024), 0, stream, blocks, nchunks);
  TN_LAUNCH_CHECK("det_scan_blocks_kernel");
  tn_launch(det_eval_kernel, dim3(nchunks), dim3(1024), 0, stream, (const u64*)keys, labels, n, (const int*)blocks, cum, cnt,
            p_target, c_fa, c_miss, eps, fnrs, fprs);
  TN_LAUNCH_CHECK("det_eval_kernel");
  tn_launch(det_final_kernel, dim3(1), dim3(32), 0, stream, (const int*)cum, (const DetCounters*)cnt, n, out8);
  TN_LAUNCH_CHECK("det_final_kernel");
  if (sorted_keys)   // low 32 bits = trial index at each sorted position (the permutation Python's sorted() produces)
    TN_CUDA(cudaMemcpyAsync(sorted_keys, keys, (size_t)n * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return TN_OK;
}
