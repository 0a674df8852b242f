// Squeeze-and-excitation and the mega-block tail, NWC layout.
//   m[b,c]    = mean_t a3[b,t,c]                          (a3 = lazy activation of sub-block 3)
//   gate[b,:] = sigmoid(W2 relu(W1 m[b,:]))               (no biases)
//   out       = dropout(relu( (s*scale_s + shift_s) + gate[b,c] * a3 ))
// Reference: modules.SqueezeExcitation.forward (src/modules.py:173-189) fed by the
// post-ReLU/dropout output of sub-block 3 (src/models.py:435-449), and
// MegaBlock.forward (src/models.py:467-472).
#include "common.cuh"
#include <stdlib.h>

// ---------------------------------------------------------------------------
// squeeze: mean over time of the lazy activation
// ---------------------------------------------------------------------------
// A cluster of SE_SPLIT thread blocks owns one utterance and 64 channels (16 quads x 16 row lanes per block); block r of the
// cluster walks the r-th part of the T frames.  Lanes are combined in a fixed order in shared memory, the blocks' partial sums
// are combined in rank order by block 0 through distributed shared memory: m is reproducible bit for bit (no floating-point
// atomics in the forward pass), WRITTEN, not accumulated, and the grid is SE_SPLIT x larger than one block per (utterance,
// channel group) could make it (an HBM-bound reduction needs the bytes in flight).
#define SE_SPLIT 4
__global__ void __launch_bounds__(TN_EW_THREADS) se_mean_kernel(const float* __restrict__ z, float* __restrict__ m, TnAct act,
                                                                int T, int C, float inv_T) {
  tn_grid_dep_sync();
  act = tn_act_init(act);
  __shared__ float4 red[TN_EW_THREADS];
  __shared__ float4 part[16];
  const int q = threadIdx.x & 15, lane = threadIdx.x >> 4;
  const int b = blockIdx.z;
  const int c = blockIdx.y * 64 + 4 * q;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tchunk = (T + SE_SPLIT - 1) / SE_SPLIT;
  const int t0 = (int)rank * tchunk, t1 = min(T, t0 + tchunk);
  float4 s = tn_zero4();
  if (c < C) {
#pragma unroll 4
    for (int t = t0 + lane; t < t1; t += 16) {
      const size_t off = ((size_t)b * T + t) * C + c;
      s = s + tn_act4(act, tn_ld4(z + off), c, off >> 2, nullptr);
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < 16) {
    float4 a = red[q];
#pragma unroll
    for (int l = 1; l < 16; ++l) a = a + red[l * 16 + q];
    part[q] = a;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && threadIdx.x < 16 && c < C) {
    float4 a = part[q];
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&part[q]);
#pragma unroll
    for (uint32_t r = 1; r < SE_SPLIT; ++r) {
      uint32_t remote;
      float4 v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote));
      a = a + v;
    }
    tn_st4(m + (size_t)b * C + c, a * inv_T);
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' shared memory stays alive until read
}

// Squeeze AND excitation in one launch (SqueezeExcitation.forward up to the gate, src/modules.py:173-186): a cluster of
// G x S thread blocks (<= 8) owns one utterance -- G channel groups of CPC channels, S parts of the T frames.  After the
// squeeze (as above: lanes in shared memory, T parts in rank order through distributed shared memory) the G group leaders
// hold m[b, :] between them; each computes its CPC-channel share of W1 m, the shares are added in group order through
// distributed shared memory (fixed order: reproducible), and each leader finishes relu -> W2 -> sigmoid for its own
// channels.  No second launch, no atomics, no serial tail: every cluster is independent.
template <int CPC>
__global__ void __launch_bounds__(TN_EW_THREADS) se_squeeze_excite_kernel(const float* __restrict__ z, float* __restrict__ m,
                                                                          float* __restrict__ gate, const float* __restrict__ W1,
                                                                          const float* __restrict__ W2, TnAct act, int T, int C,
                                                                          int Cr, int G, int S, float inv_T) {
  tn_grid_dep_sync();
  act = tn_act_init(act);
  constexpr int Q = CPC / 4, LANES = TN_EW_THREADS / Q;
  __shared__ float4 red[TN_EW_THREADS];
  __shared__ __align__(16) float mg[CPC];            // leader: mean of this group's channels
  __shared__ float hp[256];                          // leader: this group's share of W1 m (Cr <= 256)
  __shared__ float hs[256];                          // relu(W1 m)
  const int q = threadIdx.x % Q, lane = threadIdx.x / Q;
  const int b = blockIdx.y;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int g = (int)rank / S, sp = (int)rank % S;
  const int c = g * CPC + 4 * q;
  const int tchunk = (T + S - 1) / S;
  const int t0 = sp * tchunk, t1 = min(T, t0 + tchunk);
  float4 s = tn_zero4();
  if (c < C) {
#pragma unroll 4
    for (int t = t0 + lane; t < t1; t += LANES) {
      const size_t off = ((size_t)b * T + t) * C + c;
      s = s + tn_act4(act, tn_ld4(z + off), c, off >> 2, nullptr);
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < Q) {
    float4 a = red[q];
#pragma unroll
    for (int l = 1; l < LANES; ++l) a = a + red[l * Q + q];
    red[q] = a;                                      // this block's partial (read by the group leader)
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  const bool leader = sp == 0;
  if (leader && threadIdx.x < Q) {
    float4 a = red[q];
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&red[q]);
    for (int r = 1; r < S; ++r) {                    // T parts in rank order
      uint32_t remote;
      float4 v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank + (uint32_t)r));
      asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote));
      a = a + v;
    }
    a = a * inv_T;
    *reinterpret_cast<float4*>(mg + 4 * q) = a;
    if (c < C) tn_st4(m + (size_t)b * C + c, a);
  }
  __syncthreads();
  if (leader) {
    // this group's share of h_pre[j] = sum_c W1[j, c] m[c]: TPJ threads per j, each CPC / TPJ channels, shuffle-combined
    const int TPJ = TN_EW_THREADS / Cr;              // host guarantees 1 <= TPJ <= 32, a power of two, CPC % TPJ == 0
    const int j = threadIdx.x / TPJ, part = threadIdx.x % TPJ, per = CPC / TPJ;
    float a = 0.f;
    for (int i = 0; i < per; ++i) {
      const int cc = part * per + i;
      if (g * CPC + cc < C) a = fmaf(__ldg(W1 + (size_t)j * C + g * CPC + cc), mg[cc], a);
    }
    for (int o = TPJ >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (part == 0) hp[j] = a;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (leader) {
    for (int j = threadIdx.x; j < Cr; j += TN_EW_THREADS) {
      float a = 0.f;
      const uint32_t local = (uint32_t)__cvta_generic_to_shared(&hp[j]);
      for (int gg = 0; gg < G; ++gg) {               // groups in order
        uint32_t remote;
        float v;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((uint32_t)(gg * S)));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
        a += v;
      }
      hs[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < CPC; cc += TN_EW_THREADS) {
      const int ch = g * CPC + cc;
      if (ch < C) {
        float a = 0.f;
        for (int j = 0; j < Cr; ++j) a = fmaf(__ldg(W2 + (size_t)ch * Cr + j), hs[j], a);
        gate[(size_t)b * C + ch] = 1.f / (1.f + expf(-a));
      }
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' shared memory stays alive until read
}

// Squeeze + excitation + mega-block tail in ONE launch (SqueezeExcitation.forward + the residual tail of MegaBlock.forward,
// src/modules.py:173-189, src/models.py:467-472).  Same cluster as se_squeeze_excite_kernel (G channel groups x S parts of
// T, <= 8 blocks per utterance), but every block KEEPS its activated a3 = dropout(relu(bn(z3))) tile in shared memory while it
// sums it (T/S rows x CPC channels: 38.6 KB for T = 301, C = 256), fetches the finished gate of its channel group from the
// group leader through distributed shared memory, and writes out = dropout(relu(bn(s) + gate * a3)) from the tile.  z3 is read
// once (the two-launch path read it twice and hashed its dropout mask twice) and one launch ramp / drain disappears.
template <int CPC>
// (four blocks per SM: 512 blocks of a batch of 64 must be ONE wave -- at 74 registers it was three per SM, two waves, 29 us instead of 20)
__global__ void __launch_bounds__(TN_EW_THREADS, 4) se_tail_fwd_kernel(const float* __restrict__ z, const float* __restrict__ sk,
                                                                    float* __restrict__ m, float* __restrict__ gate,
                                                                    float* __restrict__ out, const float* __restrict__ W1,
                                                                    const float* __restrict__ W2, TnAct act, TnAct act_s, TnAct act_o,
                                                                    int T, int C, int Cr, int G, int S, float inv_T) {
  tn_grid_dep_sync();
  act = tn_act_init(act);
  act_o = tn_act_init(act_o);
  constexpr int Q = CPC / 4, LANES = TN_EW_THREADS / Q;
  extern __shared__ __align__(16) float4 tile[];     // [rows of this part][Q]: the activated a3 tile
  __shared__ float4 red[TN_EW_THREADS];
  __shared__ __align__(16) float mg[CPC];            // leader: mean of this group's channels
  __shared__ __align__(16) float gs[CPC];            // leader: gate of this group's channels
  __shared__ float hp[256];                          // leader: this group's share of W1 m (Cr <= 256)
  __shared__ float hs[256];                          // relu(W1 m)
  const int q = threadIdx.x % Q, lane = threadIdx.x / Q;
  const int b = blockIdx.y;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int g = (int)rank / S, sp = (int)rank % S;
  const int c = g * CPC + 4 * q;
  const int tchunk = (T + S - 1) / S;
  const int t0 = sp * tchunk, t1 = min(T, t0 + tchunk);
  float4 s = tn_zero4();
  if (c < C) {
    // four rows per iteration in flight; the sum runs in the order of the two-launch kernel (t ascending per lane)
    for (int t = t0 + lane; t < t1; t += 4 * LANES) {
      float4 zv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t + k * LANES < t1) zv[k] = tn_ld4(z + ((size_t)b * T + t + k * LANES) * C + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t + k * LANES < t1) {
          const size_t off = ((size_t)b * T + t + k * LANES) * C + c;
          const float4 a3 = tn_act4(act, zv[k], c, off >> 2, nullptr);
          tile[(size_t)(t + k * LANES - t0) * Q + q] = a3;
          s = s + a3;
        }
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < Q) {
    float4 a = red[q];
#pragma unroll
    for (int l = 1; l < LANES; ++l) a = a + red[l * Q + q];
    red[q] = a;                                      // this block's partial (read by the group leader)
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  const bool leader = sp == 0;
  if (leader && threadIdx.x < Q) {
    float4 a = red[q];
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&red[q]);
    for (int r = 1; r < S; ++r) {                    // T parts in rank order
      uint32_t remote;
      float4 v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank + (uint32_t)r));
      asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote));
      a = a + v;
    }
    a = a * inv_T;
    *reinterpret_cast<float4*>(mg + 4 * q) = a;
    if (c < C) tn_st4(m + (size_t)b * C + c, a);
  }
  __syncthreads();
  if (leader) {
    const int TPJ = TN_EW_THREADS / Cr;              // host guarantees 1 <= TPJ <= 32, a power of two, CPC % TPJ == 0
    const int j = threadIdx.x / TPJ, part = threadIdx.x % TPJ, per = CPC / TPJ;
    float a = 0.f;
    for (int i = 0; i < per; ++i) {
      const int cc = part * per + i;
      if (g * CPC + cc < C) a = fmaf(__ldg(W1 + (size_t)j * C + g * CPC + cc), mg[cc], a);
    }
    for (int o = TPJ >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (part == 0) hp[j] = a;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (leader) {
    for (int j = threadIdx.x; j < Cr; j += TN_EW_THREADS) {
      float a = 0.f;
      const uint32_t local = (uint32_t)__cvta_generic_to_shared(&hp[j]);
      for (int gg = 0; gg < G; ++gg) {               // groups in order
        uint32_t remote;
        float v;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((uint32_t)(gg * S)));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
        a += v;
      }
      hs[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < CPC; cc += TN_EW_THREADS) {
      const int ch = g * CPC + cc;
      float gv = 0.f;
      if (ch < C) {
        float a = 0.f;
        for (int j = 0; j < Cr; ++j) a = fmaf(__ldg(W2 + (size_t)ch * Cr + j), hs[j], a);
        gv = 1.f / (1.f + expf(-a));
        gate[(size_t)b * C + ch] = gv;
      }
      gs[cc] = gv;
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");   // gates are in the leaders' shared memory
  // ===== tail: out = dropout(relu(bn_s(s) + gate * a3)) from the tile =====
  float4 gq = tn_zero4();
  {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&gs[4 * q]);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((uint32_t)(g * S)));
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(gq.x), "=f"(gq.y), "=f"(gq.z), "=f"(gq.w) : "r"(remote));
  }
  if (c < C) {
    const float4 ssc = tn_ld4(act_s.scale + c), ssh = tn_ld4(act_s.shift + c);
    for (int t = t0 + lane; t < t1; t += 4 * LANES) {
      float4 sv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t + k * LANES < t1) sv[k] = tn_ld4(sk + ((size_t)b * T + t + k * LANES) * C + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t + k * LANES < t1) {
          const size_t off = ((size_t)b * T + t + k * LANES) * C + c;
          const float4 a3 = tile[(size_t)(t + k * LANES - t0) * Q + q];
          float4 v = tn_fma4(gq, a3, tn_fma4(sv[k], ssc, ssh));
          const float4 keep = tn_drop4(act_o, off >> 2);
          v.x = v.x > 0.f ? v.x * keep.x : 0.f; v.y = v.y > 0.f ? v.y * keep.y : 0.f;
          v.z = v.z > 0.f ? v.z * keep.z : 0.f; v.w = v.w > 0.f ? v.w * keep.w : 0.f;
          tn_st4(out + off, v);
        }
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' shared memory stays alive until read
}

// excitation MLP of batch item b by one block.  sm: m[C] + h[Cr].  m is read with ld.cg: in the fused kernels it was
// just accumulated by other blocks' atomics (performed in L2).
__device__ __forceinline__ void se_mlp_fwd_block(float* sm, int b, const float* __restrict__ m, const float* __restrict__ W1,
                                                 const float* __restrict__ W2, float* __restrict__ gate, int C, int Cr) {
  float* ms = sm;
  float* hs = sm + C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) ms[c] = __ldcg(m + (size_t)b * C + c);
  __syncthreads();
  for (int j = warp; j < Cr; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(__ldg(W1 + (size_t)j * C + c), ms[c], s);
    s = tn_warp_sum(s);
    if (lane == 0) hs[j] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(__ldg(W2 + (size_t)c * Cr + j), hs[j], s);
    gate[(size_t)b * C + c] = 1.f / (1.f + expf(-s));
  }
}
__global__ void __launch_bounds__(256) se_mlp_fwd_kernel(const float* __restrict__ m, const float* __restrict__ W1,
                                                         const float* __restrict__ W2, float* __restrict__ gate, int C, int Cr) {
  tn_grid_dep_sync();
  extern __shared__ float sm[];
  se_mlp_fwd_block(sm, blockIdx.x, m, W1, W2, gate, C, Cr);
}

// device-wide "last block of batch item b" test: every thread of the block calls it after its atomics were issued;
// `counter` (zero on entry) is reset by the last block, so CUDA-graph replays start clean.
__device__ __forceinline__ bool tn_last_block_of(unsigned int* counter, unsigned int blocks) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == blocks - 1) ? 1u : 0u;
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0u;
}

// backward of the MLP for batch item b by one block: dgate -> dm, dW1 +=, dW2 +=.  sm: m[C] + h[Cr] + dh[Cr] + dp[C]
__device__ __forceinline__ void se_mlp_bwd_block(float* sm, int b, const float* __restrict__ dgate, const float* __restrict__ gate,
                                                 const float* __restrict__ m, const float* __restrict__ W1,
                                                 const float* __restrict__ W2, float* __restrict__ dm,
                                                 float* __restrict__ dW1, float* __restrict__ dW2, int C, int Cr) {
  float* ms = sm;
  float* hs = ms + C;
  float* dh = hs + Cr;
  float* dp = dh + Cr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    ms[c] = m[(size_t)b * C + c];
    float g = gate[(size_t)b * C + c];
    dp[c] = __ldcg(dgate + (size_t)b * C + c) * g * (1.f - g);
  }
  __syncthreads();
  for (int j = warp; j < Cr; j += nw) {       // recompute h and dh = W2^T dp
    float s = 0.f, d = 0.f;
    for (int c = lane; c < C; c += 32) {
      s = fmaf(__ldg(W1 + (size_t)j * C + c), ms[c], s);
      d = fmaf(__ldg(W2 + (size_t)c * Cr + j), dp[c], d);
    }
    s = tn_warp_sum(s);
    d = tn_warp_sum(d);
    if (lane == 0) { hs[j] = fmaxf(s, 0.f); dh[j] = s > 0.f ? d : 0.f; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    const float dpc = dp[c], mc = ms[c];
    for (int j = 0; j < Cr; ++j) {
      s = fmaf(__ldg(W1 + (size_t)j * C + c), dh[j], s);
      atomicAdd(dW2 + (size_t)c * Cr + j, dpc * hs[j]);
      atomicAdd(dW1 + (size_t)j * C + c, dh[j] * mc);
    }
    dm[(size_t)b * C + c] = s;
  }
}
__global__ void __launch_bounds__(256) se_mlp_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ gate,
                                                         const float* __restrict__ m, const float* __restrict__ W1,
                                                         const float* __restrict__ W2, float* __restrict__ dm,
                                                         float* __restrict__ dW1, float* __restrict__ dW2, int C, int Cr) {
  tn_grid_dep_sync();
  extern __shared__ float sm[];
  se_mlp_bwd_block(sm, blockIdx.x, dgate, gate, m, W1, W2, dm, dW1, dW2, C, Cr);
}

// ---------------------------------------------------------------------------
// tail forward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TN_EW_THREADS) tail_fwd_kernel(const float* __restrict__ z3, const float* __restrict__ s,
                                                                 const float* __restrict__ gate, float* __restrict__ out,
                                                                 TnAct act3, TnAct act_s, TnAct act_o, int R, int T, int C, int rpb) {
  tn_grid_dep_sync();
  act3 = tn_act_init(act3);
  act_o = tn_act_init(act_o);
  TnTile tl = tn_tile(C);
  const int r0 = blockIdx.x * rpb, r1 = min(R, r0 + rpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    if (!tl.active || q >= tl.Q) continue;
    const float4 ssc = tn_ld4(act_s.scale + 4 * q), ssh = tn_ld4(act_s.shift + 4 * q);
    // (four rows in flight per iteration measured slower in this kernel: 15.0 -> 16.6 us)
    for (int r = r0 + tl.lane; r < r1; r += tl.lanes) {
      const int b = r / T;
      size_t off = (size_t)r * C + 4 * q;
      float4 a3 = tn_act4(act3, tn_ld4(z3 + off), 4 * q, off >> 2, nullptr);
      float4 g = tn_ld4(gate + (size_t)b * C + 4 * q);
      float4 v = tn_fma4(g, a3, tn_fma4(tn_ld4(s + off), ssc, ssh));
      float4 keep = tn_drop4(act_o, off >> 2);
      v.x = v.x > 0.f ? v.x * keep.x : 0.f; v.y = v.y > 0.f ? v.y * keep.y : 0.f;
      v.z = v.z > 0.f ? v.z * keep.z : 0.f; v.w = v.w > 0.f ? v.w * keep.w : 0.f;
      tn_st4(out + off, v);
    }
  }
}

// g = dout * d out / d pre  (pre > 0 and kept  <=>  out > 0)
__device__ __forceinline__ float4 tail_gout(float4 dout, float4 o, float inv_keep) {
  return make_float4(o.x > 0.f ? dout.x * inv_keep : 0.f, o.y > 0.f ? dout.y * inv_keep : 0.f,
                     o.z > 0.f ? dout.z * inv_keep : 0.f, o.w > 0.f ? dout.w * inv_keep : 0.f);
}

// pass 1: dgate[b,c] += sum_t g * a3
__global__ void __launch_bounds__(TN_EW_THREADS) tail_bwd1_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                  const float* __restrict__ z3, float* __restrict__ dgate,
                                                                  TnAct act3, float inv_keep_o, int T, int C, int tpb) {
  tn_grid_dep_sync();
  act3 = tn_act_init(act3);
  __shared__ float4 red[TN_EW_THREADS];
  TnTile tl = tn_tile(C);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * tpb, t1 = min(T, t0 + tpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    float4 acc = tn_zero4();
    if (tl.active && q < tl.Q)
      for (int t = t0 + tl.lane; t < t1; t += tl.lanes) {
        size_t off = ((size_t)b * T + t) * C + 4 * q;
        float4 g = tail_gout(tn_ld4(dout + off), tn_ld4(out + off), inv_keep_o);
        float4 a3 = tn_act4(act3, tn_ld4(z3 + off), 4 * q, off >> 2, nullptr);
        acc = tn_fma4(g, a3, acc);
      }
    tn_lane_reduce_atomic(tl, acc, q, dgate + (size_t)b * C, red);
  }
}

// pass 1 for a block output with TWO consumers (the next mega-block's skip conv and first depthwise conv): the two gradients
// are added while loading and their sum is written once for pass 2 -- autograd's separate sum kernel (read 2, write 1, then
// read again here) is folded into this pass.  dsum may alias neither input.
__global__ void __launch_bounds__(TN_EW_THREADS) tail_bwd1s_kernel(const float* __restrict__ dout, const float* __restrict__ dout2,
                                                                   float* __restrict__ dsum, const float* __restrict__ out,
                                                                   const float* __restrict__ z3, float* __restrict__ dgate,
                                                                   TnAct act3, float inv_keep_o, int T, int C, int tpb) {
  tn_grid_dep_sync();
  act3 = tn_act_init(act3);
  __shared__ float4 red[TN_EW_THREADS];
  TnTile tl = tn_tile(C);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * tpb, t1 = min(T, t0 + tpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    float4 acc = tn_zero4();
    if (tl.active && q < tl.Q)
      for (int t = t0 + tl.lane; t < t1; t += tl.lanes) {
        size_t off = ((size_t)b * T + t) * C + 4 * q;
        const float4 d = tn_ld4(dout + off) + tn_ld4(dout2 + off);
        tn_st4(dsum + off, d);
        float4 g = tail_gout(d, tn_ld4(out + off), inv_keep_o);
        float4 a3 = tn_act4(act3, tn_ld4(z3 + off), 4 * q, off >> 2, nullptr);
        acc = tn_fma4(g, a3, acc);
      }
    tn_lane_reduce_atomic(tl, acc, q, dgate + (size_t)b * C, red);
  }
}

// pass 1 + the excitation MLP's backward: the last block of batch item b turns dgate[b] into dm[b], dW1 +=, dW2 +=
__global__ void __launch_bounds__(TN_EW_THREADS) tail_bwd1_mlp_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                      const float* __restrict__ z3, float* __restrict__ dgate,
                                                                      unsigned int* __restrict__ counters, const float* __restrict__ gate,
                                                                      const float* __restrict__ m, const float* __restrict__ W1,
                                                                      const float* __restrict__ W2, float* __restrict__ dm,
                                                                      float* __restrict__ dW1, float* __restrict__ dW2, TnAct act3,
                                                                      float inv_keep_o, int T, int C, int Cr, int tpb) {
  tn_grid_dep_sync();
  act3 = tn_act_init(act3);
  __shared__ float4 red[TN_EW_THREADS];
  extern __shared__ float sm[];
  TnTile tl = tn_tile(C);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * tpb, t1 = min(T, t0 + tpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    float4 acc = tn_zero4();
    if (tl.active && q < tl.Q)
      for (int t = t0 + tl.lane; t < t1; t += tl.lanes) {
        size_t off = ((size_t)b * T + t) * C + 4 * q;
        float4 g = tail_gout(tn_ld4(dout + off), tn_ld4(out + off), inv_keep_o);
        float4 a3 = tn_act4(act3, tn_ld4(z3 + off), 4 * q, off >> 2, nullptr);
        acc = tn_fma4(g, a3, acc);
      }
    tn_lane_reduce_atomic(tl, acc, q, dgate + (size_t)b * C, red);
  }
  if (tn_last_block_of(counters + b, gridDim.x)) se_mlp_bwd_block(sm, b, dgate, gate, m, W1, W2, dm, dW1, dW2, C, Cr);
}

// pass 2: dz3, ds and the four per-channel reductions
__global__ void __launch_bounds__(TN_EW_THREADS) tail_bwd2_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                  const float* __restrict__ z3, const float* __restrict__ s,
                                                                  const float* __restrict__ gate, const float* __restrict__ dm,
                                                                  float* __restrict__ dz3, float* __restrict__ ds,
                                                                  float* __restrict__ dsc3, float* __restrict__ dsh3,
                                                                  float* __restrict__ dscs, float* __restrict__ dshs,
                                                                  TnAct act3, TnAct act_s, float inv_keep_o, float inv_T, int R,
                                                                  int T, int C, int rpb) {
  tn_grid_dep_sync();
  act3 = tn_act_init(act3);
  __shared__ float4 red[TN_EW_THREADS];
  TnTile tl = tn_tile(C);
  const int r0 = blockIdx.x * rpb, r1 = min(R, r0 + rpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    float4 a1 = tn_zero4(), a2 = tn_zero4(), a3s = tn_zero4(), a4 = tn_zero4();
    if (tl.active && q < tl.Q) {
      const float4 sc3 = tn_ld4(act3.scale + 4 * q), scs = tn_ld4(act_s.scale + 4 * q);
      for (int r = r0 + tl.lane; r < r1; r += tl.lanes) {
        const int b = r / T;
        size_t off = (size_t)r * C + 4 * q;
        float4 g = tail_gout(tn_ld4(dout + off), tn_ld4(out + off), inv_keep_o);
        float4 zz = tn_ld4(z3 + off), mult;
        tn_act4(act3, zz, 4 * q, off >> 2, &mult);
        float4 gt = tn_ld4(gate + (size_t)b * C + 4 * q);
        float4 dmean = tn_ld4(dm + (size_t)b * C + 4 * q) * inv_T;
        float4 da = tn_fma4(g, gt, dmean) * mult;          // dL/d pre3
        a1 = tn_fma4(da, zz, a1);
        a2 = a2 + da;
        tn_st4(dz3 + off, da * sc3);
        float4 sv = tn_ld4(s + off);
        a3s = tn_fma4(g, sv, a3s);
        a4 = a4 + g;
        tn_st4(ds + off, g * scs);
      }
    }
    tn_lane_reduce_atomic(tl, a1, q, dsc3, red);
    tn_lane_reduce_atomic(tl, a2, q, dsh3, red);
    tn_lane_reduce_atomic(tl, a3s, q, dscs, red);
    tn_lane_reduce_atomic(tl, a4, q, dshs, red);
  }
}

static int tail_rows_per_block(long long R) {
  long long target = (long long)tn_num_sms() * 4;      // few, fat blocks: every block ends in per-channel atomics
  long long rpb = (R + target - 1) / target;
  if (rpb < 32) rpb = 32;
  if (rpb > 256) rpb = 256;
  return (int)rpb;
}
static int time_per_block(int B, int T) {
  // blocks = B * ceil(T / tpb); aim at >= 8 blocks per SM
  long long target = (long long)tn_num_sms() * 4;
  long long chunks = (target + B - 1) / B;
  if (chunks < 1) chunks = 1;
  long long tpb = (T + chunks - 1) / chunks;
  if (tpb < 32) tpb = 32;
  return (int)tpb;
}

#define SE_COMMON_CHECK(name)                                                                                   \
  TN_REQUIRE(B > 0 && T > 0 && C > 0 && C % 4 == 0 && B <= 65535, name ": need C %% 4 == 0, B <= 65535 (B=%d T=%d C=%d)", B, T, C)

extern "C" int tn_se_mean(const float* z3, float* m, const float* scale, const float* shift, int relu, float drop_p,
                          const unsigned long long* seed, unsigned int layer, int B, int T, int C, void* stream) {
  SE_COMMON_CHECK("se_mean");
  TN_REQUIRE(z3 && m && (scale == nullptr) == (shift == nullptr), "se_mean: null tensor");
  dim3 grid(SE_SPLIT, tn_cdiv(C, 64), B);              // clusters of SE_SPLIT blocks along x
  tn_launch_cluster(se_mean_kernel, grid, TN_EW_THREADS, 0, stream, SE_SPLIT, z3, m, tn_make_act(scale, shift, relu, drop_p, seed, layer), T, C,
                    1.0f / (float)T);
  TN_LAUNCH_CHECK("se_mean_kernel");
  return TN_OK;
}

// cluster shape of the fused squeeze + excitation: CPC channels per block, G = ceil(C / CPC) groups, S parts of T, G * S <= 8
static bool se_fused_plan(int C, int Cr, int* cpc, int* G, int* S) {
  if (Cr < 1 || Cr > 256 || (Cr & (Cr - 1)) != 0) return false;       // TPJ = 256 / Cr threads per hidden unit
  const int tpj = 256 / Cr;
  if (tpj > 32) return false;
  for (int c : {64, 128}) {
    const int g = (C + c - 1) / c;
    if (g <= 8 && c % tpj == 0) {
      *cpc = c; *G = g;
      int sp = 8 / g;
      *S = sp > 4 ? 4 : (sp < 1 ? 1 : sp);
      return true;
    }
  }
  return false;
}
extern "C" int tn_se_squeeze_excite_supported(int C, int Cr) {
  int a, b, c;
  return se_fused_plan(C, Cr, &a, &b, &c) ? 1 : 0;
}
extern "C" int tn_se_squeeze_excite(const float* z3, float* m, float* gate, const float* W1, const float* W2, const float* scale,
                                    const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer,
                                    int B, int T, int C, int Cr, void* stream) {
  SE_COMMON_CHECK("se_squeeze_excite");
  TN_REQUIRE(z3 && m && gate && W1 && W2 && (scale == nullptr) == (shift == nullptr), "se_squeeze_excite: null tensor");
  int cpc, G, S;
  TN_UNSUPPORTED(!se_fused_plan(C, Cr, &cpc, &G, &S), "se_squeeze_excite: unsupported channel counts C=%d Cr=%d", C, Cr);
  dim3 grid(G * S, B);
  const TnAct act = tn_make_act(scale, shift, relu, drop_p, seed, layer);
  if (cpc == 64) tn_launch_cluster(se_squeeze_excite_kernel<64>, grid, TN_EW_THREADS, 0, stream, G * S, z3, m, gate, W1, W2, act, T, C, Cr, G, S, 1.0f / (float)T);
  else tn_launch_cluster(se_squeeze_excite_kernel<128>, grid, TN_EW_THREADS, 0, stream, G * S, z3, m, gate, W1, W2, act, T, C, Cr, G, S, 1.0f / (float)T);
  TN_LAUNCH_CHECK("se_squeeze_excite_kernel");
  return TN_OK;
}

// shared memory of one block's a3 tile in the fused squeeze + excitation + tail kernel, or 0 when that kernel does not apply
static size_t se_tail_tile_bytes(int T, int C, int Cr) {
  int cpc, G, S;
  if (!se_fused_plan(C, Cr, &cpc, &G, &S)) return 0;
  const size_t bytes = (size_t)((T + S - 1) / S) * cpc * sizeof(float);
  return bytes <= 160 * 1024 ? bytes : 0;
}
extern "C" int tn_se_tail_fwd_supported(int T, int C, int Cr) { return se_tail_tile_bytes(T, C, Cr) > 0 ? 1 : 0; }
// m[B, C], gate[B, C] (saved for the backward pass) and out[B*T, C] from z3, s in one launch: tn_se_squeeze_excite + tn_tail_fwd
extern "C" int tn_se_tail_fwd(const float* z3, const float* s, float* m, float* gate, float* out, const float* W1, const float* W2,
                              const float* scale3, const float* shift3, float drop3, unsigned int layer3, const float* scale_s,
                              const float* shift_s, float drop_o, unsigned int layer_o, const unsigned long long* seed, int B, int T,
                              int C, int Cr, void* stream) {
  SE_COMMON_CHECK("se_tail_fwd");
  TN_REQUIRE(z3 && s && m && gate && out && W1 && W2 && scale3 && shift3 && scale_s && shift_s, "se_tail_fwd: null tensor");
  int cpc, G, S;
  const size_t smem = se_tail_tile_bytes(T, C, Cr);
  TN_UNSUPPORTED(smem == 0 || !se_fused_plan(C, Cr, &cpc, &G, &S), "se_tail_fwd: unsupported shape T=%d C=%d Cr=%d", T, C, Cr);
  TN_REQUIRE((long long)B * T * C < (1ll << 33), "se_tail_fwd: B*T*C too large");
  dim3 grid(G * S, B);
  const TnAct act3 = tn_make_act(scale3, shift3, 1, drop3, seed, layer3), act_s = tn_make_act(scale_s, shift_s, 0, 0.f, seed, 0),
              act_o = tn_make_act(scale_s, shift_s, 1, drop_o, seed, layer_o);
  if (cpc == 64) {
    TN_CUDA(cudaFuncSetAttribute(se_tail_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tn_launch_cluster(se_tail_fwd_kernel<64>, grid, TN_EW_THREADS, smem, stream, G * S, z3, s, m, gate, out, W1, W2, act3, act_s, act_o, T, C, Cr, G, S,
                      1.0f / (float)T);
  } else {
    TN_CUDA(cudaFuncSetAttribute(se_tail_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tn_launch_cluster(se_tail_fwd_kernel<128>, grid, TN_EW_THREADS, smem, stream, G * S, z3, s, m, gate, out, W1, W2, act3, act_s, act_o, T, C, Cr, G, S,
                      1.0f / (float)T);
  }
  TN_LAUNCH_CHECK("se_tail_fwd_kernel");
  return TN_OK;
}

extern "C" int tn_se_mlp_fwd(const float* m, const float* W1, const float* W2, float* gate, int B, int C, int Cr, void* stream) {
  TN_REQUIRE(B > 0 && C > 0 && Cr > 0 && m && W1 && W2 && gate, "se_mlp_fwd: bad arguments");
  size_t smem = sizeof(float) * (size_t)(C + Cr);
  TN_REQUIRE(smem <= 48 * 1024, "se_mlp_fwd: C too large");
  tn_launch(se_mlp_fwd_kernel, B, 256, smem, stream, m, W1, W2, gate, C, Cr);
  TN_LAUNCH_CHECK("se_mlp_fwd_kernel");
  return TN_OK;
}

extern "C" int tn_se_mlp_bwd(const float* dgate, const float* gate, const float* m, const float* W1, const float* W2, float* dm,
                             float* dW1, float* dW2, int B, int C, int Cr, void* stream) {
  TN_REQUIRE(B > 0 && C > 0 && Cr > 0 && dgate && gate && m && W1 && W2 && dm && dW1 && dW2, "se_mlp_bwd: bad arguments");
  // (an atomics-free variant -- one block per row of dW1 reducing over the batch itself -- measured 19.5 us against 10.5 us:
  //  every such block re-reads the [B, C] tensors m, gate, dgate through one SM's L2 port)
  size_t smem = sizeof(float) * (size_t)(2 * C + 2 * Cr);
  TN_REQUIRE(smem <= 48 * 1024, "se_mlp_bwd: C too large");
  tn_launch(se_mlp_bwd_kernel, B, 256, smem, stream, dgate, gate, m, W1, W2, dm, dW1, dW2, C, Cr);
  TN_LAUNCH_CHECK("se_mlp_bwd_kernel");
  return TN_OK;
}

extern "C" int tn_tail_fwd(const float* z3, const float* s, const float* gate, float* out, const float* scale3,
                           const float* shift3, float drop3, unsigned int layer3, const float* scale_s, const float* shift_s,
                           float drop_o, unsigned int layer_o, const unsigned long long* seed, int B, int T, int C, void* stream) {
  SE_COMMON_CHECK("tail_fwd");
  TN_REQUIRE(z3 && s && gate && out && scale3 && shift3 && scale_s && shift_s, "tail_fwd: null tensor");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "tail_fwd: B*T too large");
  int rpb = tail_rows_per_block(R);
  tn_launch(tail_fwd_kernel, tn_cdiv(R, rpb), TN_EW_THREADS, 0, stream, 
      z3, s, gate, out, tn_make_act(scale3, shift3, 1, drop3, seed, layer3), tn_make_act(scale_s, shift_s, 0, 0.f, seed, 0),
      tn_make_act(scale_s, shift_s, 1, drop_o, seed, layer_o), (int)R, T, C, rpb);
  TN_LAUNCH_CHECK("tail_fwd_kernel");
  return TN_OK;
}

extern "C" int tn_tail_bwd1(const float* dout, const float* out, const float* z3, float* dgate, const float* scale3,
                            const float* shift3, float drop3, unsigned int layer3, float drop_o, const unsigned long long* seed, int B,
                            int T, int C, void* stream) {
  SE_COMMON_CHECK("tail_bwd1");
  TN_REQUIRE(dout && out && z3 && dgate && scale3 && shift3, "tail_bwd1: null tensor");
  int tpb = time_per_block(B, T);
  dim3 grid(tn_cdiv(T, tpb), B);
  float inv_keep_o = drop_o > 0.f ? 1.f / (1.f - drop_o) : 1.f;
  tn_launch(tail_bwd1_kernel, grid, TN_EW_THREADS, 0, stream, dout, out, z3, dgate, tn_make_act(scale3, shift3, 1, drop3, seed, layer3), inv_keep_o, T, C, tpb);
  TN_LAUNCH_CHECK("tail_bwd1_kernel");
  return TN_OK;
}

// the same for two gradients of the block output: dsum = dout + dout2 is written for tn_tail_bwd2; dgate ACCUMULATED
extern "C" int tn_tail_bwd1s(const float* dout, const float* dout2, float* dsum, const float* out, const float* z3, float* dgate,
                             const float* scale3, const float* shift3, float drop3, unsigned int layer3, float drop_o,
                             const unsigned long long* seed, int B, int T, int C, void* stream) {
  SE_COMMON_CHECK("tail_bwd1s");
  TN_REQUIRE(dout && dout2 && dsum && out && z3 && dgate && scale3 && shift3, "tail_bwd1s: null tensor");
  TN_REQUIRE(dsum != dout && dsum != dout2, "tail_bwd1s: dsum must not alias a gradient");
  int tpb = time_per_block(B, T);
  dim3 grid(tn_cdiv(T, tpb), B);
  float inv_keep_o = drop_o > 0.f ? 1.f / (1.f - drop_o) : 1.f;
  tn_launch(tail_bwd1s_kernel, grid, TN_EW_THREADS, 0, stream, dout, dout2, dsum, out, z3, dgate, tn_make_act(scale3, shift3, 1, drop3, seed, layer3),
            inv_keep_o, T, C, tpb);
  TN_LAUNCH_CHECK("tail_bwd1s_kernel");
  return TN_OK;
}

// tn_tail_bwd1 + tn_se_mlp_bwd in one launch; dgate, dW1, dW2 ACCUMULATED, counters: B zeroed uints (self-resetting)
extern "C" int tn_tail_bwd1_mlp(const float* dout, const float* out, const float* z3, float* dgate, unsigned int* counters,
                                const float* gate, const float* m, const float* W1, const float* W2, float* dm, float* dW1,
                                float* dW2, const float* scale3, const float* shift3, float drop3, unsigned int layer3, float drop_o,
                                const unsigned long long* seed, int B, int T, int C, int Cr, void* stream) {
  SE_COMMON_CHECK("tail_bwd1_mlp");
  TN_REQUIRE(dout && out && z3 && dgate && counters && gate && m && W1 && W2 && dm && dW1 && dW2 && scale3 && shift3 && Cr > 0,
             "tail_bwd1_mlp: null tensor");
  size_t smem = sizeof(float) * (size_t)(2 * C + 2 * Cr);
  TN_REQUIRE(smem <= 40 * 1024, "tail_bwd1_mlp: C too large");
  int tpb = time_per_block(B, T);
  dim3 grid(tn_cdiv(T, tpb), B);
  float inv_keep_o = drop_o > 0.f ? 1.f / (1.f - drop_o) : 1.f;
  tn_launch(tail_bwd1_mlp_kernel, grid, TN_EW_THREADS, smem, stream, dout, out, z3, dgate, counters, gate, m, W1, W2, dm, dW1, dW2,
            tn_make_act(scale3, shift3, 1, drop3, seed, layer3), inv_keep_o, T, C, Cr, tpb);
  TN_LAUNCH_CHECK("tail_bwd1_mlp_kernel");
  return TN_OK;
}

// dsc3/dsh3/dscs/dshs are ACCUMULATED into (caller zeroes them)
extern "C" int tn_tail_bwd2(const float* dout, const float* out, const float* z3, const float* s, const float* gate,
                            const float* dm, float* dz3, float* ds, float* dsc3, float* dsh3, float* dscs, float* dshs,
                            const float* scale3, const float* shift3, float drop3, unsigned int layer3, const float* scale_s,
                            const float* shift_s, float drop_o, const unsigned long long* seed, int B, int T, int C, void* stream) {
  SE_COMMON_CHECK("tail_bwd2");
  TN_REQUIRE(dout && out && z3 && s && gate && dm && dz3 && ds && dsc3 && dsh3 && dscs && dshs && scale3 && shift3 && scale_s && shift_s,
             "tail_bwd2: null tensor");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "tail_bwd2: B*T too large");
  int rpb = tail_rows_per_block(R);
  float inv_keep_o = drop_o > 0.f ? 1.f / (1.f - drop_o) : 1.f;
  tn_launch(tail_bwd2_kernel, tn_cdiv(R, rpb), TN_EW_THREADS, 0, stream, 
      dout, out, z3, s, gate, dm, dz3, ds, dsc3, dsh3, dscs, dshs, tn_make_act(scale3, shift3, 1, drop3, seed, layer3),
      tn_make_act(scale_s, shift_s, 0, 0.f, seed, 0), inv_keep_o, 1.0f / (float)T, (int)R, T, C, rpb);
  TN_LAUNCH_CHECK("tail_bwd2_kernel");
  return TN_OK;
}

// ---------------------------------------------------------------------------
// stand-alone squeeze-excitation gate (modules.SqueezeExcitation.forward outside a MegaBlock,
// src/modules.py:173-189) and the broadcast that is the backward of a mean over time
// (nn.AdaptiveAvgPool1d(1) in SqueezeExcitation and in Decoder(simple_pool=True), src/models.py:497-502)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TN_EW_THREADS) gate_mul_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                                                     float* __restrict__ out, int R, int T, int C, int rpb) {
  tn_grid_dep_sync();
  TnTile tl = tn_tile(C);
  const int r0 = blockIdx.x * rpb, r1 = min(R, r0 + rpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    if (!tl.active || q >= tl.Q) continue;
    for (int r = r0 + tl.lane; r < r1; r += tl.lanes) {
      const size_t off = (size_t)r * C + 4 * q;
      tn_st4(out + off, tn_ld4(x + off) * tn_ld4(gate + (size_t)(r / T) * C + 4 * q));
    }
  }
}
// dx = dout * gate ; dgate[b, c] += sum_t dout * x   (one utterance per blockIdx.y so the reduction stays per batch item)
__global__ void __launch_bounds__(TN_EW_THREADS) gate_mul_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                                     const float* __restrict__ gate, float* __restrict__ dx,
                                                                     float* __restrict__ dgate, int T, int C, int tpb) {
  tn_grid_dep_sync();
  __shared__ float4 red[TN_EW_THREADS];
  TnTile tl = tn_tile(C);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * tpb, t1 = min(T, t0 + tpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    float4 s = tn_zero4();
    if (tl.active && q < tl.Q) {
      const float4 g = tn_ld4(gate + (size_t)b * C + 4 * q);
      for (int t = t0 + tl.lane; t < t1; t += tl.lanes) {
        const size_t off = ((size_t)b * T + t) * C + 4 * q;
        const float4 d = tn_ld4(dout + off);
        s = tn_fma4(d, tn_ld4(x + off), s);
        tn_st4(dx + off, d * g);
      }
    }
    tn_lane_reduce_atomic(tl, s, q, dgate + (size_t)b * C, red);
  }
}
// out[b*T + t, c] = v[b, c] * mul
__global__ void __launch_bounds__(TN_EW_THREADS) bcast_rows_kernel(const float* __restrict__ v, float* __restrict__ out, float mul,
                                                                   int R, int T, int C, int rpb) {
  tn_grid_dep_sync();
  TnTile tl = tn_tile(C);
  const int r0 = blockIdx.x * rpb, r1 = min(R, r0 + rpb);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    if (!tl.active || q >= tl.Q) continue;
    for (int r = r0 + tl.lane; r < r1; r += tl.lanes)
      tn_st4(out + (size_t)r * C + 4 * q, tn_ld4(v + (size_t)(r / T) * C + 4 * q) * mul);
  }
}

extern "C" int tn_gate_mul_fwd(const float* x, const float* gate, float* out, int B, int T, int C, void* stream) {
  SE_COMMON_CHECK("gate_mul_fwd");
  TN_REQUIRE(x && gate && out, "gate_mul_fwd: null tensor");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "gate_mul_fwd: B*T too large");
  int rpb = tail_rows_per_block(R);
  tn_launch(gate_mul_fwd_kernel, tn_cdiv(R, rpb), TN_EW_THREADS, 0, stream, x, gate, out, (int)R, T, C, rpb);
  TN_LAUNCH_CHECK("gate_mul_fwd_kernel");
  return TN_OK;
}
extern "C" int tn_gate_mul_bwd(const float* dout, const float* x, const float* gate, float* dx, float* dgate, int B, int T, int C,
                               void* stream) {
  SE_COMMON_CHECK("gate_mul_bwd");
  TN_REQUIRE(dout && x && gate && dx && dgate, "gate_mul_bwd: null tensor");
  int tpb = time_per_block(B, T);
  dim3 grid(tn_cdiv(T, tpb), B);
  tn_launch(gate_mul_bwd_kernel, grid, TN_EW_THREADS, 0, stream, dout, x, gate, dx, dgate, T, C, tpb);
  TN_LAUNCH_CHECK("gate_mul_bwd_kernel");
  return TN_OK;
}
extern "C" int tn_bcast_rows(const float* v, float* out, float mul, int B, int T, int C, void* stream) {
  SE_COMMON_CHECK("bcast_rows");
  TN_REQUIRE(v && out, "bcast_rows: null tensor");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "bcast_rows: B*T too large");
  int rpb = tail_rows_per_block(R);
  tn_launch(bcast_rows_kernel, tn_cdiv(R, rpb), TN_EW_THREADS, 0, stream, v, out, mul, (int)R, T, C, rpb);
  TN_LAUNCH_CHECK("bcast_rows_kernel");
  return TN_OK;
}
