// Shared device/host helpers for libtitanet_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "titanet_b200.h"

void tn_set_error(const char* fmt, ...);

#define TN_REQUIRE(cond, ...)                      \
  do {                                             \
    if (!(cond)) {                                 \
      tn_set_error(__VA_ARGS__);                   \
      return TN_EINVAL;                            \
    }                                              \
  } while (0)

#define TN_UNSUPPORTED(cond, ...)                  \
  do {                                             \
    if (cond) {                                    \
      tn_set_error(__VA_ARGS__);                   \
      return TN_EUNSUPPORTED;                      \
    }                                              \
  } while (0)

#define TN_CUDA(expr)                                                          \
  do {                                                                         \
    cudaError_t e__ = (expr);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      tn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                   __FILE__, __LINE__);                                        \
      return (int)e__;                                                         \
    }                                                                          \
  } while (0)

#define TN_LAUNCH_CHECK(name)                                                  \
  do {                                                                         \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) {                                                  \
      tn_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));  \
      return (int)e__;                                                         \
    }                                                                          \
  } while (0)

// ---------------------------------------------------------------------------
// Launch helper.  With TN_PDL=1 in the environment every kernel of the library is launched with the
// "programmatic stream serialization" attribute (programmatic dependent launch): the grid may be scheduled while
// its predecessor in the stream is still draining, and every kernel starts with tn_grid_dep_sync()
// (griddepcontrol.wait), which blocks until the predecessor has completed and its writes are visible.  Measured on
// the graph-replayed TitaNet-S step: 10.213 ms without, 10.205 ms with (the 1-CTA/SM GEMM kernels cannot co-reside
// with their successors), so the attribute is off by default; without it griddepcontrol.* are no-ops.
// ---------------------------------------------------------------------------
bool tn_pdl_enabled();
#ifdef __CUDACC__
#include <utility>
template <typename... KArgs, typename... Args>
static inline void tn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tn_pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);     // errors surface in TN_LAUNCH_CHECK
}
// same, as thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
static inline void tn_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, unsigned cluster_x,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tn_pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
// first statement of every kernel: let the successor be scheduled, then wait for the predecessor's results
__device__ __forceinline__ void tn_grid_dep_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

static inline bool tn_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
static inline int tn_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int tn_num_sms();

// ---------------------------------------------------------------------------
// "lazy activation": a = dropout(relu(z * scale[c] + shift[c]))
// A tensor travels through the encoder as its pre-BatchNorm values z plus the
// per-channel affine (scale, shift) folded from the BatchNorm statistics; the
// consumer kernel applies affine + ReLU + dropout while loading.  scale==nullptr
// means "plain tensor" (identity).
// ---------------------------------------------------------------------------
struct TnAct {
  const float* scale;   // [C] or nullptr
  const float* shift;   // [C]
  int relu;
  float inv_keep;       // 1/(1-p), 1 when no dropout
  uint32_t thresh;      // drop when rand < thresh ; 0 => no dropout
  const unsigned long long* seed_ptr;   // device scalar: the step's dropout seed (CUDA-graph safe)
  uint32_t seed_lo, seed_hi, layer;     // seed_lo/hi are filled on the device by tn_act_init
};

static inline TnAct tn_make_act(const float* scale, const float* shift, int relu, float p,
                                const unsigned long long* seed, unsigned int layer) {
  TnAct a;
  a.scale = scale; a.shift = shift; a.relu = relu;
  if (p > 0.f) {
    // 16-bit drop threshold: one 32-bit hash decides TWO elements (its low / high half), so the drop probability is
    // p rounded to 1/65536 (|error| < 8e-6; inv_keep stays the nominal 1/(1-p))
    double t = (double)p * 65536.0 + 0.5;
    a.thresh = t >= 65535.0 ? 65535u : (uint32_t)t;
    if (a.thresh == 0) a.thresh = 1;
    a.inv_keep = 1.0f / (1.0f - p);
  } else {
    a.thresh = 0; a.inv_keep = 1.0f;
  }
  if (seed == nullptr) { a.thresh = 0; a.inv_keep = 1.0f; }   // no seed => no dropout (callers validate)
  a.seed_ptr = seed; a.seed_lo = 0; a.seed_hi = 0; a.layer = layer;
  return a;
}

#ifdef __CUDACC__
// Counter-based dropout RNG: one 32-bit hash per PAIR of consecutive elements of (step seed, layer id,
// pair index = element index >> 1) -- a seeded multiply-xorshift finaliser ("lowbias32"); the even element
// compares the low 16 bits with the threshold, the odd one the high 16 bits.  Stateless, so forward and
// backward regenerate the same mask, and any thread can ask for any element (the tensor-core epilogues own
// one channel per thread, the row-tiled kernels four = two hashes).  Measured: with one hash per element the
// hash was 30-50 % of the HBM-bound kernels' time (se_mean 10.3 us with dropout, 5.4 us without).
__device__ __forceinline__ uint32_t tn_hash_elem(uint32_t seed_lo, uint32_t seed_hi, uint32_t layer, unsigned long long idx) {
  uint32_t h = (uint32_t)idx * 0x9E3779B1u ^ seed_lo;
  h ^= (uint32_t)(idx >> 32) * 0xC2B2AE3Du + layer * 0x85EBCA77u + seed_hi;
  h ^= h >> 16; h *= 0x21F0AAADu;
  h ^= h >> 15; h *= 0x735A2D97u;
  h ^= h >> 15;
  return h;
}
// The same hash for element indices below 2^32: the high-word term is a per-launch constant, so the
// caller folds it into `key = tn_hash_key32(...)` once and pays one multiply + the finaliser per element.
__device__ __forceinline__ uint32_t tn_hash_key32(uint32_t seed_lo, uint32_t seed_hi, uint32_t layer) {
  return seed_lo ^ (layer * 0x85EBCA77u + seed_hi);
}
__device__ __forceinline__ uint32_t tn_hash_elem32(uint32_t key, uint32_t idx) {
  uint32_t h = idx * 0x9E3779B1u ^ key;
  h ^= h >> 16; h *= 0x21F0AAADu;
  h ^= h >> 15; h *= 0x735A2D97u;
  h ^= h >> 15;
  return h;
}
// keep-multiplier (0 or inv_keep) of element `idx` (= row * C + channel)
__device__ __forceinline__ float tn_drop1(const struct TnAct& a, unsigned long long idx);

// load the step seed once per thread (kernel parameters are read-only: work on a copy)
__device__ __forceinline__ TnAct tn_act_init(TnAct a) {
  if (a.thresh != 0) {
    const unsigned long long s = __ldg(a.seed_ptr);
    a.seed_lo = (uint32_t)s; a.seed_hi = (uint32_t)(s >> 32);
  }
  return a;
}

// keep-multipliers (0 or inv_keep) for the 4 consecutive elements of quad `qidx`
__device__ __forceinline__ float4 tn_drop4(const TnAct& a, unsigned long long qidx) {
  if (a.thresh == 0) return make_float4(1.f, 1.f, 1.f, 1.f);
  const uint32_t h0 = tn_hash_elem(a.seed_lo, a.seed_hi, a.layer, qidx << 1);          // elements 4q, 4q+1
  const uint32_t h1 = tn_hash_elem(a.seed_lo, a.seed_hi, a.layer, (qidx << 1) | 1ull);  // elements 4q+2, 4q+3
  return make_float4((h0 & 0xFFFFu) >= a.thresh ? a.inv_keep : 0.f, (h0 >> 16) >= a.thresh ? a.inv_keep : 0.f,
                     (h1 & 0xFFFFu) >= a.thresh ? a.inv_keep : 0.f, (h1 >> 16) >= a.thresh ? a.inv_keep : 0.f);
}
__device__ __forceinline__ float tn_drop1(const TnAct& a, unsigned long long idx) {
  if (a.thresh == 0) return 1.f;
  const uint32_t h = tn_hash_elem(a.seed_lo, a.seed_hi, a.layer, idx >> 1);
  return ((idx & 1ull) ? (h >> 16) : (h & 0xFFFFu)) >= a.thresh ? a.inv_keep : 0.f;
}
// scalar lazy activation of element (row, c): returns a, *mult = d a / d pre
__device__ __forceinline__ float tn_act1(const TnAct& a, float z, int c, unsigned long long idx, float* mult) {
  if (a.scale == nullptr) { *mult = 1.f; return z; }
  const float v = fmaf(z, __ldg(a.scale + c), __ldg(a.shift + c));
  float m = tn_drop1(a, idx);
  if (a.relu && !(v > 0.f)) m = 0.f;
  *mult = m;
  return v * m;
}

// a = act(z) for one channel quad; `mult` (optional) receives d a / d pre, i.e. the
// factor the backward pass multiplies by (before the extra *scale[c]).
__device__ __forceinline__ float4 tn_act4(const TnAct& a, float4 z, int c, unsigned long long qidx, float4* mult) {
  if (a.scale == nullptr) {
    if (mult) *mult = make_float4(1.f, 1.f, 1.f, 1.f);
    return z;
  }
  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + c));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(a.shift + c));
  float4 v = make_float4(fmaf(z.x, sc.x, sh.x), fmaf(z.y, sc.y, sh.y), fmaf(z.z, sc.z, sh.z), fmaf(z.w, sc.w, sh.w));
  float4 m = tn_drop4(a, qidx);
  if (a.relu) {
    m.x = v.x > 0.f ? m.x : 0.f; m.y = v.y > 0.f ? m.y : 0.f;
    m.z = v.z > 0.f ? m.z : 0.f; m.w = v.w > 0.f ? m.w : 0.f;
  }
  if (mult) *mult = m;
  return make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
}

__device__ __forceinline__ float tn_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float tn_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 tn_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void tn_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float b) { return make_float4(a.x * b, a.y * b, a.z * b, a.w * b); }
__device__ __forceinline__ float4 tn_fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 tn_zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ---------------------------------------------------------------------------
// Reproducible per-channel statistics + train-mode BatchNorm fold by the LAST block of a channel group.
//
// Every block of the producing kernel holds per-channel partial sums (sum z, sum z^2 over ITS rows, fp32, computed in a
// fixed order).  They are added into a 120-bit FIXED-POINT accumulator per value (units of 2^-60, two unsigned 64-bit
// atomics): integer addition is associative, so the total does not depend on the order in which the blocks arrive and two
// runs of a forward pass agree bit for bit -- without the serial tail of a last block re-reading every block's partials
// (68 KB through one SM's L2 port: +3.5 us per GEMM, measured).  Blocks that share a channel range [c0, c0 + nC) form a
// group with one device-wide ticket; the last block of a group reads the totals, returns the accumulators (and the ticket)
// to zero for the next launch / graph replay, writes stats[c] / stats[C_total + c] and, if `f` is given, folds them into
// (scale, shift), stores (mean, invstd) and updates the running statistics.  nn.BatchNorm1d semantics (biased variance for
// normalisation, unbiased for running_var).  Range: |partial| < 2^58 (larger or non-finite partials mark the channel group
// and its statistics come out NaN, as an overflowing fp32 sum would); resolution 2^-60 ~ 8.7e-19 absolute.
// accum layout: hi words [2 * C_total] | lo words [2 * C_total] | TN_TICKETS 32-bit flags (one per group); all ZERO on entry.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned int* tn_fix_flag(unsigned long long* accum, int C_total, int group) {
  return reinterpret_cast<unsigned int*>(accum + 4 * (size_t)C_total) + group;
}
__device__ __forceinline__ void tn_fix_add(unsigned long long* accum, int C_total, int which, int c, float p, unsigned int* flag) {
  if (!(fabsf(p) < 2.8e17f)) {                       // NaN, Inf or beyond the fixed-point range
    atomicAdd(flag, 1u);
    return;
  }
  const double d = (double)p;
  const double h = floor(d * 16.0);
  const double frac = d - h * 0.0625;                 // exact: in [0, 1/16)
  const unsigned long long lq = (unsigned long long)(frac * 1152921504606846976.0);     // * 2^60: exact (24-bit mantissa), < 2^56
  atomicAdd(accum + (size_t)which * C_total + c, (unsigned long long)(long long)h);
  atomicAdd(accum + (size_t)(2 + which) * C_total + c, lq);
}

// read-back + fold of channels [c0, c0 + nC) by `nthreads` threads (this one is number `tid`): the totals are final
struct TnFoldConst { double inv_n, unbias; };     // 1 / n and n / (n - 1) (1 when n == 1), computed on the host
static inline TnFoldConst tn_fold_const(double n) {
  TnFoldConst c;
  c.inv_n = n > 0.0 ? 1.0 / n : 0.0;
  c.unbias = n > 1.0 ? n / (n - 1.0) : 1.0;
  return c;
}
__device__ __forceinline__ void tn_stats_fold_channels(const tn_bn_fold* f, double* stats, unsigned long long* accum, int C_total,
                                                       int c0, int nC, unsigned int* flag, bool bump_nbt, int tid, int nthreads,
                                                       TnFoldConst fc) {
  const double inv_n = fc.inv_n, unbias = fc.unbias;
  const bool bad = __ldcg(flag) != 0u;
  for (int cb = tid; cb < nC; cb += 4 * nthreads) {
    // up to four channels per thread and pass: all sixteen loads are issued before anything is used (one global round trip
    // per pass, not per channel -- a single warp folds 128 channels in one pass)
    unsigned long long w[4][4];
    float pre[4][4];                                  // gamma, beta, running_mean, running_var: loaded up front as well (the
                                                      // stores below may alias them as far as the compiler knows)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = cb + j * nthreads;
      if (cl < nC) {
        const unsigned long long* a = accum + c0 + cl;
        w[j][0] = __ldcg(a); w[j][1] = __ldcg(a + C_total); w[j][2] = __ldcg(a + 2 * (size_t)C_total); w[j][3] = __ldcg(a + 3 * (size_t)C_total);
        if (f) {
          pre[j][0] = __ldg(f->gamma + c0 + cl); pre[j][1] = __ldg(f->beta + c0 + cl);
          if (f->running_mean) { pre[j][2] = f->running_mean[c0 + cl]; pre[j][3] = f->running_var[c0 + cl]; }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
    const int cl = cb + j * nthreads;
    if (cl >= nC) break;
    const int c = c0 + cl;
    unsigned long long* a = accum + c;
    const unsigned long long h1 = w[j][0], h2 = w[j][1], l1 = w[j][2], l2 = w[j][3];
    a[0] = 0ull; a[C_total] = 0ull; a[2 * (size_t)C_total] = 0ull; a[3 * (size_t)C_total] = 0ull;
    double s1 = (double)(long long)h1 * 0.0625 + (double)l1 * 8.673617379884035e-19;      // 2^-60
    double s2 = (double)(long long)h2 * 0.0625 + (double)l2 * 8.673617379884035e-19;
    if (bad) s1 = s2 = __longlong_as_double(0x7ff8000000000000ll);
    if (stats) { stats[c] = s1; stats[C_total + c] = s2; }
    if (f) {
      // few fp64 operations per channel (the fp64 pipe is slow): one reciprocal, and 1 / sqrt by one fp64 Newton step on the
      // fp32 rsqrt (error ~2^-44, then rounded to fp32)
      // inv_n = 1 / n and unbias = n / (n - 1) come from the host: an fp64 division per channel is most of this tail
      const double m = s1 * inv_n;
      double var = s2 * inv_n - m * m;
      if (var < 0.0) var = 0.0;
      const float mean = (float)m;
      const double ve = var + (double)f->eps;
      double r = (double)rsqrtf((float)ve);
      r = r * (1.5 - 0.5 * ve * r * r);
      const float invstd = (float)r;
      if (f->running_mean) {
        const float unbiased = (float)(var * unbias);
        f->running_mean[c] = (1.f - f->momentum) * pre[j][2] + f->momentum * mean;
        f->running_var[c] = (1.f - f->momentum) * pre[j][3] + f->momentum * unbiased;
      }
      const float sc = pre[j][0] * invstd;
      f->scale[c] = sc;
      f->shift[c] = pre[j][1] - mean * sc;
      f->mean[c] = mean;
      f->invstd[c] = invstd;
    }
    }
  }
  if (nthreads <= 32) __syncwarp(); else __syncthreads();
  if (tid == 0) {
    *flag = 0u;
    if (f && bump_nbt && f->num_batches_tracked) *f->num_batches_tracked += 1;
  }
}

__device__ __forceinline__ void tn_stats_finish(const tn_bn_fold* f, double* stats, unsigned long long* accum, int C_total,
                                                int c0, int nC, unsigned int* ticket, unsigned int* flag,
                                                unsigned int blocks_in_group, bool bump_nbt, TnFoldConst fc) {
  __shared__ unsigned int s_is_last;
  __threadfence();                                   // this thread's atomics before the block's ticket
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_is_last = (t == blocks_in_group - 1) ? 1u : 0u;
    if (s_is_last) *ticket = 0u;                     // ready for the next launch / graph replay
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  tn_stats_fold_channels(f, stats, accum, C_total, c0, nC, flag, bump_nbt, threadIdx.y * blockDim.x + threadIdx.x, blockDim.x * blockDim.y, fc);
}

// ---------------------------------------------------------------------------
// Row-tiled channel-quad work distribution shared by the HBM-bound NWC kernels.
// A block of TN_EW_THREADS threads owns `rows_per_block` consecutive rows of an
// [R, C] tensor.  Thread -> (quad q, lane l): quads are consecutive across
// threads (coalesced float4), lanes stride over rows.  Per-channel partial sums
// are combined across lanes in shared memory and leave the block as one atomic
// per channel.
// ---------------------------------------------------------------------------
#define TN_EW_THREADS 256

struct TnTile {
  int Q;        // quads per row (C/4)
  int qpb;      // quads handled per pass by the block = min(Q, TN_EW_THREADS)
  int lanes;    // row lanes = TN_EW_THREADS / qpb
  int q0;       // this thread's first quad
  int lane;     // this thread's lane
  bool active;
};

__device__ __forceinline__ TnTile tn_tile(int C) {
  TnTile t;
  t.Q = C >> 2;
  t.qpb = t.Q < TN_EW_THREADS ? t.Q : TN_EW_THREADS;
  t.lanes = TN_EW_THREADS / t.qpb;
  t.q0 = threadIdx.x % t.qpb;
  t.lane = threadIdx.x / t.qpb;
  t.active = t.lane < t.lanes;
  return t;
}

// Sum `v` (a per-thread partial for channel quad q) over the lanes of the block and
// atomically add the result to dst[4q..4q+3].  `red` is TN_EW_THREADS float4s of
// shared memory.  Must be called by all threads of the block the same number of times.
template <typename DstT>
__device__ __forceinline__ void tn_lane_reduce_atomic(const TnTile& t, float4 v, int q, DstT* dst, float4* red, float mul = 1.f) {
  // Same-line atomics serialise in the L2 slice (~27 cycles per warp-level visit of a 128-byte
  // line), so the block leaves with ONE visit per line: consecutive threads own consecutive
  // channels (32 floats = one line per warp instruction) instead of 4 strided scalars per thread.
  __syncthreads();
  red[threadIdx.x] = t.active ? v : tn_zero4();
  __syncthreads();
  const float* rf = reinterpret_cast<const float*>(red);
  const int qb = q - t.q0;                       // first quad of this pass (block-uniform)
  const int nch = min(4 * t.qpb, 4 * (t.Q - qb));
  for (int ch = threadIdx.x; ch < nch; ch += TN_EW_THREADS) {
    float s = rf[ch];
    for (int l = 1; l < t.lanes; ++l) s += rf[ch + l * 4 * t.qpb];
    atomicAdd(dst + 4 * qb + ch, (DstT)(s * mul));
  }
}
#endif  // __CUDACC__
