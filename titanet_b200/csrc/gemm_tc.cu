// 1x1-convolution / linear layers on the 5th-generation tensor cores (tcgen05 + TMEM),
// operands staged by TMA, fp32-equivalent split arithmetic.
//
//   Z[r, co] = bias[co] + sum_ci X[r, ci] * W[co, ci]          X: [R, Kd]  W: [M, Kd]  Z: [R, M]
//
// Orientation: the output channels are the MMA's M (TMEM lanes), the activation rows are
// its N (TMEM columns), both operands K-major in shared memory:
//     D[co, r] (TMEM) += A[co, k] (= W, smem) * B[r, k] (= X, smem)
// so in the epilogue one thread owns one output channel: the bias is a scalar, the
// BatchNorm statistics (sum z, sum z^2 per channel) are thread-local sums, and for a
// fixed row the 32 lanes of a warp write 32 consecutive floats (one 128 B line).
//
// Precision: fp32 operands are split hi = rna_tf32(x), lo = x - hi.  One kind::tf32 MMA hi*hi plus ONE kind::f16 MMA over a
// doubled K that carries both corrections lo*hi + hi*lo (tc_store_corr): scaled fp16 in forward GEMMs (operand rounding =
// 3xTF32's), bf16 in gradient GEMMs (fp32's exponent range); the original three-MMA 3xTF32 scheme stays selectable
// (tc_corr_for).  All accumulate in fp32 TMEM tiles and are fp32-equivalent (SURVEY.md §7 hard part 1 shows plain TF32
// breaks the 1e-3 parity contract through train-mode BatchNorm).  The weight split is precomputed once per step
// (tn_split_tf32_batch); the activation tile is split in shared memory by the transform warps between the TMA arrival and
// the MMA issue.  nsplit = 1 skips the split (the tensor core truncates fp32 to tf32).
//
// Kernels: gemm_tc2_kernel (a cta_group::2 pair of CTAs per 256 channels x 2 N tiles: the main path), gemm_tc_kernel (one CTA,
// shapes the pair kernel does not take), wgrad_tc_kernel (weight gradients, MN-major operands, split-K).  Warp roles: warp 0
// TMA producer, warp 1 TMEM allocator + MMA issuer, the others operand transform, then epilogue (TMEM -> registers ->
// global).  The tcgen05.mma / TMA issue sites run in WARP-UNIFORM code and branch on one elect.sync (tc_elect_one): from
// `if (lane == 0)` code the compiler serialises every such instruction through an ELECT / BRA.U.ANY loop.
// Reference ops replaced: the pointwise Conv1dSamePadding(C, C', 1) of DepthwiseConv1d
// (src/modules.py:76-78), the skip nn.Conv1d (src/models.py:452-455), the epilog conv
// (src/models.py:384) and the ASP linears (src/models.py:549-551).
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include <stdlib.h>

#define TC_THREADS 192        // wgrad kernel: TMA + MMA + 4 epilogue warps
#define TC_GEMM_THREADS 320   // GEMM kernel: TMA + MMA + 8 transform/epilogue warps (two per TMEM lane quadrant)
#define TC_EPI_THREADS 256
// fp32 elements per K chunk: 32 = 128-byte swizzled rows, 2 stages of ~100 KB; 16 = 64-byte rows, 4 stages of ~50 KB.
// Measured at R=19264, 256x256 (tools/trace_gemm.py): BK=32 delivers a chunk (84 KB) every 1.41 us against 1.06 us of MMA,
// mainloop 12.8 us; BK=16 delivers 42 KB every 0.95 us (64-byte rows use the TMA unit worse), mainloop 15.9 us.  Both sit
// at the per-SM TMA rate (~45-60 GB/s), most of it weight tiles re-streamed by every CTA; multicasting the weights
// across a 2-CTA cluster (TN_TC_CLUSTER=1) is functional but changes nothing (L2 already de-duplicates neighbours).
#define TC_BK 32
#define TC_MAX_STAGES 6
#define TC_SMEM_LIMIT (227 * 1024)

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (spins == 1024) t0 = clock64();
    if (spins > 1024 && (spins & 1023) == 0 && clock64() - t0 > 4000000000ll) {   // ~2 s: a protocol bug, not a slow GPU
      printf("gemm_tc: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// multicast variants for a 2-CTA cluster: the tile (and its complete_tx) lands at the same offsets in every CTA of
// `mask`; the commit arrives on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// one lane of a converged warp (elect.sync): the tcgen05 issue sites branch on it from warp-uniform code
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// issue a 16-column TMEM load without waiting; tc_ld_wait() makes the registers valid.  The
// "+f" constraints on the wait tie the data dependence so the compiler cannot hoist uses above it.
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                 "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait(float* v, int n) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < n) asm volatile("" : "+f"(v[i]));
}
__device__ __forceinline__ uint32_t rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// ---------------------------------------------------------------------------
// Correction operand of the default split scheme ("TF32 + BF16 corrections", p.corr = 1; TN_TC_3XTF32=1 turns it off).
//   x.w = xh.wh + (xl.wh + xh.wl) + xl.wl,  xh = rna_tf32(x), xl = x - xh (|xl| <= 2^-11 |x|)
// The bracket is ~2^-11 of the result, so its operands only need ~8 bits: it is issued as ONE kind::f16 (bf16) MMA over a
// doubled K -- activation row [bf16(xl) x32 | bf16(xh) x32] against weight row [bf16(wh) x32 | bf16(wl) x32] -- next to
// the kind::tf32 MMA xh.wh.  A bf16 MMA of K = 16 costs what a tf32 MMA of K = 8 costs, and both advance 32 bytes of a
// 128-byte swizzled row, so a K chunk takes 2 MMA issues per 32 bytes instead of 3 (3xTF32: hi.hi + lo.hi + hi.lo): one
// third fewer tensor-pipe cycles at an operand-rounding error of ~7e-7 rms (3xTF32 8e-8, torch fp32 2e-7, plain TF32 3e-4;
// tools/split_precision.py).  The "lo" buffers keep their size and (fp32-typed) TMA maps: 64 bf16 fill the 128 bytes that
// held 32 tf32.  The environment variable switches the weight split and the GEMMs alike.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float first, float second) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(first, second);      // .x (low half, lower address) = first
  return *reinterpret_cast<const uint32_t*>(&t);
}
// Forward GEMMs (p.corr = 2): the correction operands are fp16 with exponent-balanced scales,
//   (xl * 16) . (wh / 16)  and  (xh / 256) . (wl * 256)         (TC_F16_SA = 16, TC_F16_SC = 256: powers of two, exact)
// fp16 carries 11 significand bits, so each correction product is good to 2^-12 of itself = 2^-24 of the result: the operand
// rounding of this scheme equals 3xTF32's (7.6e-8 rms, tools/split_precision.py) at two tensor-pipe issues per 32 bytes of K
// instead of three.  The scales keep all four operands in fp16's normal range for activations of typical magnitude
// 2^-6 .. 2^20 and weights 2^-10 .. 2^16 (smaller values lose relative, not absolute, precision; conversions saturate, so
// the worst case is the plain-TF32 product, never an infinity).  Gradient GEMMs stay on bf16 (p.corr = 1): gradients of
// magnitude 1e-8 are below fp16's range.
#define TC_F16_SA 16.0f
#define TC_F16_SC 256.0f
__device__ __forceinline__ uint32_t pack_f16x2(float first, float second) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));   // d.lo = second operand
  return r;
}
// v = four consecutive K elements (logical 16-byte fp32 chunk c = 0..7) of row `row` of a [rows x 32 fp32] SWIZZLE_128B
// tile, h = their tf32 parts: write the 16-bit (v - h) to K' = 4c.. and the 16-bit h to K' = 32 + 4c.. of the tile at lo_base
// (F16 = false: bf16, unscaled; F16 = true: fp16, scaled as above)
template <bool F16>
__device__ __forceinline__ void tc_store_corr_t(uint8_t* lo_base, uint32_t row, uint32_t c, float4 v, uint4 h) {
  const float hx = __uint_as_float(h.x), hy = __uint_as_float(h.y), hz = __uint_as_float(h.z), hw = __uint_as_float(h.w);
  uint2 plo, phi;
  if (F16) {
    plo = make_uint2(pack_f16x2((v.x - hx) * TC_F16_SA, (v.y - hy) * TC_F16_SA), pack_f16x2((v.z - hz) * TC_F16_SA, (v.w - hw) * TC_F16_SA));
    phi = make_uint2(pack_f16x2(hx * (1.f / TC_F16_SC), hy * (1.f / TC_F16_SC)), pack_f16x2(hz * (1.f / TC_F16_SC), hw * (1.f / TC_F16_SC)));
  } else {
    plo = make_uint2(pack_bf16x2(v.x - hx, v.y - hy), pack_bf16x2(v.z - hz, v.w - hw));
    phi = make_uint2(pack_bf16x2(hx, hy), pack_bf16x2(hz, hw));
  }
  uint8_t* r = lo_base + row * 128u + ((c & 1u) << 3);
  *reinterpret_cast<uint2*>(r + ((((c >> 1)) ^ (row & 7u)) << 4)) = plo;
  *reinterpret_cast<uint2*>(r + (((4u + (c >> 1)) ^ (row & 7u)) << 4)) = phi;
}
__device__ __forceinline__ void tc_store_corr(uint8_t* lo_base, uint32_t row, uint32_t c, float4 v, uint4 h, int corr = 1) {
  if (corr == 2) tc_store_corr_t<true>(lo_base, row, c, v, h);
  else tc_store_corr_t<false>(lo_base, row, c, v, h);
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor of the 16-bit correction MMA from the tf32 one: A / B format TF32 (2) -> BF16 (1) or F16 (0)
__device__ __forceinline__ uint32_t tc_idesc_bf16(uint32_t idesc_tf32, int corr = 1) {
  const uint32_t fmt = corr == 2 ? 0u : 1u;
  return (idesc_tf32 & ~((7u << 7) | (7u << 10))) | (fmt << 7) | (fmt << 10);
}

// shared-memory matrix descriptor, K-major operand tile of TC_BK fp32 per row:
// 128-byte rows: SWIZZLE_128B (layout 2), 8-row groups 1024 B apart; 64-byte rows: SWIZZLE_64B (layout 4), 512 B apart
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
#if TC_BK == 32
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
#else
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
#endif
}

// Epilogue of one 128-channel accumulator tile: TMEM -> registers -> (+bias, tanh, +=) -> global.
// This thread owns output channel `zp[0]`; column j of the tile is activation row j (stride ld).
// Rows >= nvalid (tail tile) are masked by predication, not branches.
// nacc > 1: the tile was accumulated in `nacc` TMEM accumulators `astride` columns apart (K-chunked main products + one
// accumulator for the small correction products, see the MMA issuer); they are added here in fp32 (round to nearest).
template <bool TANH, bool ACCUM, bool FULL>
__device__ __forceinline__ void tc_epilogue(uint32_t tbase, float* __restrict__ zp, size_t ld, int BN, int nvalid, float bv,
                                            float& s1, float& s2, int half, int nparts = 2, int nacc = 1, int astride = 0) {
  for (int c0 = 32 * half; c0 < BN; c0 += 32 * nparts) {   // the `nparts` warps of a lane quadrant alternate 32-column chunks
    float v[32];
    const bool two = c0 + 16 < BN;                   // BN is a multiple of 16
    tc_ld16_issue(tbase + c0, v);
    if (two) tc_ld16_issue(tbase + c0 + 16, v + 16);
    tc_ld_wait(v, 32);
    for (int a = 1; a < nacc; ++a) {                 // warp-uniform
      float w[32];
      tc_ld16_issue(tbase + a * astride + c0, w);
      if (two) tc_ld16_issue(tbase + a * astride + c0 + 16, w + 16);
      tc_ld_wait(w, 32);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < 16 || two) v[j] += w[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j >= 16 && !two) break;
      const bool ok = FULL || (c0 + j < nvalid);
      float x = v[j] + bv;
      if (TANH) x = tanhf(x);
      float* q = zp + (size_t)(c0 + j) * ld;
      if (ACCUM) {
        // Z += x as a fire-and-forget reduction (red.global.add.f32, one 128-byte line per warp): a load-add-store here
        // put 4-byte global loads on the eight epilogue warps' critical path (+60 us per launch, measured)
        if (ok) atomicAdd(q, x);
      } else if (ok) {
        *q = x;
        s1 += x;
        s2 = fmaf(x, x, s2);
      }
    }
  }
}

// statistics pre-pass of the plain epilogue: the same values x = acc + bias the store pass will write, summed per channel
// (sum, sum of squares) in the same column order -- nothing is stored.  See the pair kernel's epilogue for why it runs first.
template <bool FULL>
__device__ __forceinline__ void tc_epilogue_stats(uint32_t tbase, int BN, int nvalid, float bv, float& s1, float& s2, int half,
                                                  int nparts, int nacc, int astride) {
  for (int c0 = 32 * half; c0 < BN; c0 += 32 * nparts) {
    float v[32];
    const bool two = c0 + 16 < BN;
    tc_ld16_issue(tbase + c0, v);
    if (two) tc_ld16_issue(tbase + c0 + 16, v + 16);
    tc_ld_wait(v, 32);
    for (int a = 1; a < nacc; ++a) {
      float w[32];
      tc_ld16_issue(tbase + a * astride + c0, w);
      if (two) tc_ld16_issue(tbase + a * astride + c0 + 16, w + 16);
      tc_ld_wait(w, 32);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < 16 || two) v[j] += w[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j >= 16 && !two) break;
      if (FULL || (c0 + j < nvalid)) {
        const float x = v[j] + bv;
        s1 += x;
        s2 = fmaf(x, x, s2);
      }
    }
  }
}

struct TcParams {
  const float* bias;
  float* Z;
  double* stats;
  int R, Kd, M_total, BN, stages, nsplit, flags, tmem_cols;
  int corr;                     // 2: TF32 + scaled-fp16 correction (forward GEMMs), 1: TF32 + bf16 correction (gradient GEMMs), 0: 3xTF32
  int nacc;                     // TMEM accumulators per tile (>= 1), BN columns apart: nacc - 1 K-chunked main accumulators + 1 for the corrections
  unsigned long long* accum;    // fixed-point statistics accumulators [hi 2 M_total | lo 2 M_total | flags] (tn_fix_add / tn_stats_finish)
  unsigned int* tickets;        // one per channel group, zero on entry, reset by the kernel
  int cluster2;          // launched as 2-CTA clusters along x: the two CTAs share (multicast) the weight tiles
  long long* trace;      // optional timeline buffer (debug): 128 slots per traced CTA
  // fused depthwise-backward epilogue (dw_K > 0): this GEMM is the data gradient of a pointwise conv
  // whose input was depthwise_K(act(zprev)); the epilogue turns du (in TMEM) into dzprev directly.
  int dw_K, dw_T, BNo;   // taps, frames per utterance, output rows per CTA (BN = BNo + halo)
  // fused depthwise FORWARD (fdw_K > 0): the B operand is u = depthwise_K(act(z)) + b, produced by the transform warps from
  // the raw z tile TMA delivers (act = p.act); u is also written to fdw_u for the weight-gradient GEMM of the backward pass
  int fdw_K, fdw_T;
  const float* fdw_w;    // [Kd, K]
  const float* fdw_b;    // [Kd] or NULL
  float* fdw_u;          // [R, Kd] or NULL
  uint32_t par_off;      // byte offset of the per-channel parameter table [3 + K][Kd] behind the pipeline stages
  int z_early;           // z tiles may be requested while the last chunks are still in the tensor core (2 stages)
  uint32_t z_off1;       // byte offset of the second z tile in the (freed) pipeline memory when !z_early
  uint32_t red_off;      // byte offset of the per-channel reduction staging buffer (behind the pipeline stages)
  const float* dw_w;     // [C, K]
  tn_bn_fold bn;         // has_bn: the last CTA of a channel group folds the statistics into (scale, shift)
  TnFoldConst fc;
  int has_bn;
  const float* zprev;    // [R, C]
  float* dzprev;         // [R, C]
  float* g_dw;           // [C, K]   ACCUMULATED
  float* g_db;           // [C]      ACCUMULATED
  float* g_dscale;       // [C]      ACCUMULATED (act only)
  float* g_dshift;       // [C]      ACCUMULATED (act only)
  TnAct act;
  // BatchNorm backward folded into the operand load of a data-gradient GEMM (pair kernel): the B operand is
  // g = dZ + a[c] + b[c] z, built by the transform warps from the dZ tile and the z tile (delivered into the B_lo buffers);
  // g is also written to bnb.g_out for the weight-gradient GEMM and its column sums to bnb.dbias (see tn_bn_bwd)
  tn_bn_bwd bnb;
  int has_bnb;
};

// Epilogue for the fused depthwise backward.  The accumulator tile holds du[c, j] for the rows
// n0 + j (n0 = first output row - PAD: the tile carries a PAD-row halo on both sides, recomputed
// by the neighbouring CTAs).  One thread owns channel c, so the transposed depthwise conv walks
// along its own TMEM columns and every per-channel reduction (d scale, d shift, d dw weights,
// d dw bias) is a thread-local sum; z of the previous layer is read once, coalesced across lanes.
// Reference: autograd of DepthwiseConv1d's first conv + BatchNorm/ReLU/Dropout
// (src/modules.py:64-75, 128-133) as reached by loss.backward() (src/learn.py:117).
// tcgen05.ld of 2 accumulator columns, issue only (tc_ld_wait makes the registers valid)
__device__ __forceinline__ void tc_ld2_issue(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(taddr));
}
// tcgen05.ld of 4 accumulator columns, issue only (tc_ld_wait makes the registers valid)
__device__ __forceinline__ void tc_ld4_issue(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(taddr));
}

// The epilogue is a ROLLED loop over groups of four output rows with the depthwise window carried in
// registers (the first version unrolled 16 rows: 21 KB of straight-line code per chunk, and the ncu source
// view showed the eight epilogue warps stalled on instruction fetch, stall_no_inst, on nearly every line).
// Per group: four new accumulator columns (tcgen05.ld.x4, issued one group ahead), four new z rows of the
// previous layer (coalesced 128-byte lines, issued two groups ahead), their activation / dropout mask, four
// outputs.  The two warps of a lane quadrant own the two halves of the tile's rows.
template <int K>
__device__ __forceinline__ void tc_epilogue_dwbwd(uint32_t tbase, const TcParams& p, const TnAct& act, int c, int n0, int half,
                                                  const uint8_t* zs, float* red, int bn2 = 0x7fffffff, const uint8_t* zs2 = nullptr,
                                                  int nparts = 2, int nthreads = 256) {
  // `half` = which of the `nparts` warps of this lane quadrant (each owns a contiguous range of the tile's output rows);
  // `nthreads` = epilogue threads of the CTA (named barrier 1, cooperative atomics)
  // Tile geometry.  Single-CTA kernel: tile column j is TMEM column tbase + j and z row j of `zs`.  Pair kernel
  // (cta_group::2): the tile is two N tiles of bn2 columns; columns >= bn2 live at TMEM column 256 + (j - bn2) and in
  // the second z box `zs2`.  All accesses below are in aligned pairs / single rows, so they never straddle the seam.
  auto tcol = [&](int j) -> uint32_t { return tbase + (uint32_t)(j >= bn2 ? j - bn2 + 256 : j); };
  constexpr int PAD = K / 2;
  constexpr int WW = 4 + 2 * PAD;                    // window: tile columns o .. o + 3 + 2 PAD for outputs o .. o + 3
  const int C = p.M_total, T = p.dw_T, R = p.R;
  float w[K], a_w[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { w[k] = __ldg(p.dw_w + (size_t)c * K + k); a_w[k] = 0.f; }
  const bool lazy = act.scale != nullptr;
  const bool drop = act.thresh != 0;
  const float sc = lazy ? __ldg(act.scale + c) : 1.f;
  const float sh = lazy ? __ldg(act.shift + c) : 0.f;
  const uint32_t key = tn_hash_key32(act.seed_lo, act.seed_hi, act.layer);   // R * C < 2^32 (checked by the launcher)
  float a_sc = 0.f, a_sh = 0.f, a_b = 0.f;
  const int r_first = n0 + PAD;                      // global row of output 0 (tile column j <-> global row n0 + j)
  const int nout = min(p.BNo, R - r_first);
  const int hsplit = ((p.BNo + nparts - 1) / nparts + 7) & ~7;   // rows per part, a multiple of 8 (BNo is a multiple of 16)
  const int oa = half * hsplit;
  const int ob = min(min((half + 1) * hsplit, p.BNo), nout);     // this warp's outputs: [oa, ob)
  const bool traced = p.trace && threadIdx.x == 64 && blockIdx.y == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2);
  long long* tr = traced ? p.trace + (blockIdx.x == 0 ? 0 : 128) + 113 + (c >= 128 ? 4 : 0) : nullptr;
  if (tr) tr[0] = clock64();
  if (oa < ob) {
  float* dzcol = p.dzprev + c;
  // z of the previous layer for this CTA's rows was brought into shared memory by TMA (128-byte swizzled rows of
  // 32 channels; rows outside the tensor arrive as zeros): `zs` is this thread's 32-channel block, tile row j is
  // global row n0 + j.  A warp reads one 128-byte row: conflict-free.  (Per-thread global loads could not keep
  // enough bytes in flight: eight warps x 4 B left the epilogue latency-bound at ~8 GB/s per SM.)
  const uint32_t lane = threadIdx.x & 31;
  auto load_zt = [&](int j) -> float {
    const uint8_t* base = j >= bn2 ? zs2 : zs;
    const uint32_t jr = (uint32_t)(j >= bn2 ? j - bn2 : j);
    const uint32_t off = jr * 128u + ((((lane >> 2) ^ (jr & 7u)) << 4) | ((lane & 3u) << 2));
    return *reinterpret_cast<const float*>(base + off);
  };
  // activation of window row `row` from its z (0 for rows outside the tensor); m = d a / d pre
  auto actf = [&](float zv, int row, float& m) -> float {
    m = 1.f;
    if (!lazy) return zv;
    const float pre = fmaf(zv, sc, sh);
    if (drop) {
      const uint32_t idx = (uint32_t)row * (uint32_t)C + (uint32_t)c;                 // same pairing as tn_drop1 / tn_drop4
      const uint32_t h = tn_hash_elem32(key, idx >> 1);
      m = ((idx & 1u) ? (h >> 16) : (h & 0xFFFFu)) >= act.thresh ? act.inv_keep : 0.f;
    }
    if (act.relu && !(pre > 0.f)) m = 0.f;
    return (row >= 0 && row < R) ? pre * m : 0.f;
  };
  float g[WW], a[WW], zc[PAD + 4], mc[PAD + 4];
  // prime window positions [0, 2 PAD)
  if (PAD > 0) {
    float tmp[2 * PAD + 1];
#pragma unroll
    for (int jj = 0; jj < PAD; ++jj) tc_ld2_issue(tcol(oa + 2 * jj), tmp + 2 * jj);
    tc_ld_wait(tmp, 2 * PAD);
#pragma unroll
    for (int j = 0; j < 2 * PAD; ++j) {
      const int row = n0 + oa + j;
      const float zv = load_zt(oa + j);
      float m;
      g[j] = tmp[j];
      a[j] = actf(zv, row, m);
      if (j >= PAD) { zc[j - PAD] = zv; mc[j - PAD] = m; }
    }
  }
  float gq[4];
  if (tr) tr[1] = clock64();
  tc_ld2_issue(tcol(oa + 2 * PAD), gq);
  tc_ld2_issue(tcol(oa + 2 * PAD + 2), gq + 2);
  int t0 = (r_first + oa) % T;
#pragma unroll 2
  for (int o = oa; o < ob; o += 4) {
    tc_ld_wait(gq, 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) g[2 * PAD + i] = gq[i];
    if (o + 4 < ob) {                                                 // warp-uniform
      tc_ld2_issue(tcol(o + 4 + 2 * PAD), gq);
      tc_ld2_issue(tcol(o + 6 + 2 * PAD), gq + 2);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m;
      const float zv = load_zt(o + 2 * PAD + i);
      a[2 * PAD + i] = actf(zv, n0 + o + 2 * PAD + i, m);
      zc[PAD + i] = zv;
      mc[PAD + i] = m;
    }
    const int cnt = ob - o;
    float* dq = dzcol + (size_t)((uint32_t)(r_first + o) * (uint32_t)C);
    if (t0 >= PAD && t0 + 3 + PAD < T && cnt >= 4) {
      // the group and its taps lie inside one utterance and inside the tensor: no boundary tests
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float g0 = g[i + PAD];
        float da = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          da = fmaf(w[k], g[i + 2 * PAD - k], da);
          a_w[k] = fmaf(g0, a[i + k], a_w[k]);
        }
        a_b += g0;
        float out = da;
        if (lazy) {
          const float gg = da * mc[i];
          a_sc = fmaf(gg, zc[i], a_sc);
          a_sh += gg;
          out = gg * sc;
        }
        dq[(uint32_t)(i * C)] = out;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = i < cnt;
        const int t = (t0 + i) % T;
        const float g0 = g[i + PAD];
        float da = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          // forward: u[r'] += w[k] * a[r' + k - PAD]  =>  a[r] feeds u[r + PAD - k]
          const int tu = t + PAD - k;
          if (tu >= 0 && tu < T) da = fmaf(w[k], g[i + 2 * PAD - k], da);
          const int ta = t + k - PAD;
          if (ok && ta >= 0 && ta < T) a_w[k] = fmaf(g0, a[i + k], a_w[k]);
        }
        if (ok) {
          a_b += g0;
          float out = da;
          if (lazy) {
            const float gg = da * mc[i];
            a_sc = fmaf(gg, zc[i], a_sc);
            a_sh += gg;
            out = gg * sc;
          }
          dq[(uint32_t)(i * C)] = out;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2 * PAD; ++j) { g[j] = g[j + 4]; a[j] = a[j + 4]; }
#pragma unroll
    for (int j = 0; j < PAD; ++j) { zc[j] = zc[j + 4]; mc[j] = mc[j + 4]; }
    t0 += 4;
    if (t0 >= T) t0 %= T;
  }
  }
  if (tr) tr[2] = clock64();
  // Per-channel sums leave the CTA through shared memory: the two warps of a lane quadrant are combined and the
  // atomics are issued with consecutive threads on consecutive addresses, i.e. ONE visit per 128-byte line per CTA
  // (same-line atomics serialise in the L2 slice; one thread per channel with stride-K addresses made 96 line
  // visits per CTA and the backlog stalled the next channel half for ~4 us).
  constexpr int F = K + 3;
  const int chl = (int)(threadIdx.x & 127u);              // (warp & 3) * 32 + lane: channel inside this 128-channel half
  float* mine = red + ((size_t)half * 128 + chl) * F;
#pragma unroll
  for (int k = 0; k < K; ++k) mine[k] = a_w[k];
  mine[K] = a_b; mine[K + 1] = a_sc; mine[K + 2] = a_sh;
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
  const int tid = (int)threadIdx.x - 64;
  const int cbase = c - chl;
  auto rsum = [&](int ch, int f) -> float {
    float v = red[ch * F + f];
    for (int pp = 1; pp < nparts; ++pp) v += red[(pp * 128 + ch) * F + f];
    return v;
  };
  for (int L = tid; L < 128 * K; L += nthreads) {
    const int ch = L / K, k = L - ch * K;
    atomicAdd(p.g_dw + (size_t)cbase * K + L, rsum(ch, k));
  }
  if (tid < 128) {
    if (p.g_db) atomicAdd(p.g_db + cbase + tid, rsum(tid, K));
  } else if (lazy && tid < 256) {
    const int ch = tid - 128;
    atomicAdd(p.g_dscale + cbase + ch, rsum(ch, K + 1));
    atomicAdd(p.g_dshift + cbase + ch, rsum(ch, K + 2));
  }
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");   // `red` is reused by the next channel half
  if (tr) tr[3] = clock64();
}

// Fused depthwise-forward operand producer (the transform warps' mainloop when fdw_K > 0).  Per K chunk (32 channels)
// TMA delivers the RAW z tile into the B_hi region: raw row r <-> global row n0 - PAD + r (rows outside the tensor are
// zero-filled), 128-byte swizzled rows.  Thread (q = channel quad 0..7, row lane 0..31) owns up to RP consecutive output
// rows j0 .. j0+rp-1: phase 1 pulls its window of RP + 2 PAD raw float4s into registers, a named barrier lets every thread
// finish reading, phase 2 applies BN/ReLU/dropout, the K-tap FIR and the tf32 hi/lo split and writes the operand rows in
// place (hi) and to B_lo.  Per-channel parameters (scale, shift, bias, taps) sit in shared memory as [field][channel].
// Reference ops: nn.BatchNorm1d/ReLU/Dropout of the previous ConvBlock1d + DepthwiseConv1d's first conv
// (src/modules.py:64-75, 128-133).
template <int K, int RP>
__device__ __forceinline__ void tc_dw_mainloop(const TcParams& p, uint8_t* smem, uint32_t stage_bytes, uint32_t bh_off, uint32_t bl_off,
                                               uint32_t full0, uint32_t ready0, int S, int num_kc, int n0, int BN, const float* par,
                                               int tid) {
  constexpr int PAD = K / 2;
  constexpr int W = RP + 2 * PAD;
  const TnAct act = tn_act_init(p.act);
  const int C = p.Kd, T = p.fdw_T, R = p.R;
  const bool lazy = act.scale != nullptr;
  const bool drop = act.thresh != 0;
  const uint32_t key = tn_hash_key32(act.seed_lo, act.seed_hi, act.layer);
  const uint32_t q = (uint32_t)tid & 7u;
  const int rpa = (BN + 31) >> 5;                            // rows per lane actually used (<= RP)
  const int j0 = (tid >> 3) * rpa;
  const int rp = min(rpa, BN - j0);                          // may be <= 0 for the last lanes
  const int t0 = (int)(((long long)n0 + j0) % T);
  const bool interior = rp > 0 && t0 - PAD >= 0 && t0 + rp - 1 + PAD < T;   // every tap of every row inside one utterance
  // validity of the window rows (global row n0 - PAD + j0 + i inside the tensor) as a bit mask
  uint32_t vmask = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const int grow = n0 - PAD + j0 + i;
    if (grow >= 0 && grow < R) vmask |= 1u << i;
  }
  for (int kc = 0; kc < num_kc; ++kc) {
    const int s = kc % S;
    uint8_t* bh = smem + (size_t)s * stage_bytes + bh_off;
    uint8_t* bl = smem + (size_t)s * stage_bytes + bl_off;
    const int c = kc * TC_BK + 4 * (int)q;
    // per-channel parameters of this thread's four channels
    const float4 sc4 = *reinterpret_cast<const float4*>(par + c);
    const float4 sh4 = *reinterpret_cast<const float4*>(par + C + c);
    const float4 b4 = *reinterpret_cast<const float4*>(par + 2 * C + c);
    mbar_wait(full0 + 8 * s, (kc / S) & 1);
    float4 win[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const uint32_t r = (uint32_t)(j0 + i);
      if (i < rp + 2 * PAD) win[i] = *reinterpret_cast<const float4*>(bh + r * 128u + ((q ^ (r & 7u)) << 4));
      else win[i] = tn_zero4();
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");        // every thread has its window: the raw tile may be overwritten
    if (rp > 0) {
      if (lazy) {
        // 32-bit element indices (the launcher guarantees (R + 16) * C < 2^32); pair p = idx >> 1 feeds one hash for two
        // elements, exactly like tn_drop4 (common.cuh)
#pragma unroll
        for (int i = 0; i < W; ++i) {
          float4 v = make_float4(fmaf(win[i].x, sc4.x, sh4.x), fmaf(win[i].y, sc4.y, sh4.y), fmaf(win[i].z, sc4.z, sh4.z),
                                 fmaf(win[i].w, sc4.w, sh4.w));
          float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
          if (drop) {
            const uint32_t pr = ((uint32_t)(n0 - PAD + j0 + i) * (uint32_t)C + (uint32_t)c) >> 1;   // rows < 0 wrap: masked below
            const uint32_t h0 = tn_hash_elem32(key, pr), h1 = tn_hash_elem32(key, pr + 1u);
            m = make_float4((h0 & 0xFFFFu) >= act.thresh ? act.inv_keep : 0.f, (h0 >> 16) >= act.thresh ? act.inv_keep : 0.f,
                            (h1 & 0xFFFFu) >= act.thresh ? act.inv_keep : 0.f, (h1 >> 16) >= act.thresh ? act.inv_keep : 0.f);
          }
          if (act.relu) {
            m.x = v.x > 0.f ? m.x : 0.f; m.y = v.y > 0.f ? m.y : 0.f;
            m.z = v.z > 0.f ? m.z : 0.f; m.w = v.w > 0.f ? m.w : 0.f;
          }
          win[i] = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
        }
        if (vmask != (1u << W) - 1u) {                      // tensor edge: rows outside contribute nothing
#pragma unroll
          for (int i = 0; i < W; ++i)
            if (!((vmask >> i) & 1u)) win[i] = tn_zero4();
        }
      }
      float4 wk[K];
#pragma unroll
      for (int k = 0; k < K; ++k) wk[k] = *reinterpret_cast<const float4*>(par + (3 + k) * C + c);
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        if (i < rp) {
          const int j = j0 + i;
          float4 acc = b4;
          if (interior) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc = tn_fma4(wk[k], win[i + k], acc);
          } else {
            const int t = (t0 + i) % T;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const int tt = t + k - PAD;
              if (tt >= 0 && tt < T) acc = tn_fma4(wk[k], win[i + k], acc);
            }
          }
          if (p.fdw_u && n0 + j < R) tn_st4(p.fdw_u + (size_t)(n0 + j) * C + c, acc);
          uint4 h, l;
          h.x = rna_tf32(acc.x); h.y = rna_tf32(acc.y); h.z = rna_tf32(acc.z); h.w = rna_tf32(acc.w);
          l.x = rna_tf32(acc.x - __uint_as_float(h.x)); l.y = rna_tf32(acc.y - __uint_as_float(h.y));
          l.z = rna_tf32(acc.z - __uint_as_float(h.z)); l.w = rna_tf32(acc.w - __uint_as_float(h.w));
          const uint32_t o = (uint32_t)j * 128u + ((q ^ ((uint32_t)j & 7u)) << 4);
          *reinterpret_cast<uint4*>(bh + o) = h;
          if (p.corr) tc_store_corr(reinterpret_cast<uint8_t*>(bl), (uint32_t)j, (uint32_t)q, acc, h, p.corr);
          else *reinterpret_cast<uint4*>(bl + o) = l;
        }
      }
    }
    fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
    mbar_arrive(ready0 + 8 * s);
  }
}
#define TC_TRACE(slot) do { if (p.trace && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && blockIdx.y == 0) \
    p.trace[(blockIdx.x == 0 ? 0 : 128) + (slot)] = clock64(); } while (0)

// ---------------------------------------------------------------------------
// kernel: MT = number of 128-row output-channel tiles per CTA (1 or 2)
// ---------------------------------------------------------------------------
// MODE: 0 plain GEMM epilogue, 1 fused depthwise-BACKWARD epilogue (dw_K > 0), 2 fused depthwise-FORWARD operand producer
// (fdw_K > 0).  A template parameter, not a runtime branch: one instantiation carrying all three costs the plain GEMM ~1.5 us
// per launch (registers / instruction cache), measured.
template <int MT, int MODE>
__global__ void __launch_bounds__(TC_GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmZ, TcParams p) {
  tn_grid_dep_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * TC_MAX_STAGES + 1 + 2];
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { TC_TRACE(110); if (p.trace && blockIdx.y == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2)) {
      unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[(blockIdx.x == 0 ? 0 : 128) + 111] = (long long)gt; } }
  if (threadIdx.x == 0 && p.trace && blockIdx.y == 0 && blockIdx.x < 256) {
    unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[256 + 2 * blockIdx.x] = (long long)gt; }
  const int BN = p.BN, S = p.stages;
  // first activation row of this CTA's tile (fused depthwise backward: the tile starts PAD rows early;
  // TMA zero-fills rows < 0 and >= R)
  const int n0 = MODE == 1 ? blockIdx.x * p.BNo - (p.dw_K >> 1) : blockIdx.x * BN;
  const int m0 = blockIdx.y * (128 * MT);            // first output channel of this CTA
  const int num_kc = p.Kd / TC_BK;
  const bool split = p.nsplit == 3;

  const uint32_t a_tile = 128 * TC_BK * 4;           // 16 KiB per 128-channel weight tile
  const uint32_t b_tile = (uint32_t)BN * TC_BK * 4;
  // fused depthwise forward: TMA delivers the RAW z tile with 16 extra rows (halo of the K-tap FIR) into the B_hi region
  const uint32_t b_raw = b_tile + (MODE == 2 ? 16u * TC_BK * 4 : 0u);
  const uint32_t nA = (split ? 2u : 1u) * MT;
  const uint32_t stage_bytes = nA * a_tile + b_raw + (split ? b_tile : 0u);
  const uint32_t tx_bytes = nA * a_tile + b_raw;                         // B_lo is written by the transform warps, not TMA
  // stage layout: A_hi[MT] | A_lo[MT] | B_hi (raw tile) | B_lo
  auto a_hi = [&](int s, int mt) { return smem + (size_t)s * stage_bytes + mt * a_tile; };
  auto a_lo = [&](int s, int mt) { return smem + (size_t)s * stage_bytes + (MT + mt) * a_tile; };
  auto b_hi = [&](int s) { return smem + (size_t)s * stage_bytes + nA * a_tile; };
  auto b_lo = [&](int s) { return smem + (size_t)s * stage_bytes + nA * a_tile + b_raw; };
  const uint32_t full0 = smem_u32(&bars[0]), ready0 = smem_u32(&bars[TC_MAX_STAGES]), empty0 = smem_u32(&bars[2 * TC_MAX_STAGES]);
  const uint32_t accum_bar = smem_u32(&bars[3 * TC_MAX_STAGES]);
  const uint32_t z_bar0 = smem_u32(&bars[3 * TC_MAX_STAGES + 1]);     // fused depthwise backward: z tile of output-channel half mt
  // z tile mt (128 channels = 4 blocks of [BN rows x 128 B]) lands in pipeline memory the mainloop no longer needs
  const uint32_t z_blk = (uint32_t)BN * 128u;
  // z_early: the last S chunks sit in stages 0 .. S-1 in order (num_kc % S == 0), tile mt takes stages [mt S/2, (mt+1) S/2)
  auto z_off = [&](int mt) -> uint32_t {
    if (p.z_early) return (uint32_t)(mt * (S >> 1)) * stage_bytes;
    return mt ? p.z_off1 : 0u;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(ready0 + 8 * s, TC_EPI_THREADS);
      mbar_init(empty0 + 8 * s, p.cluster2 ? 2 : 1);     // cluster: both CTAs' tensor cores must be done with the stage
    }
    mbar_init(accum_bar, 1);
    mbar_init(z_bar0, 1);
    mbar_init(z_bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (p.cluster2) cluster_sync_all();        // the peer's barriers must be initialised before anything is multicast to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t cta_rank = p.cluster2 ? cluster_cta_rank() : 0u;
  if (threadIdx.x == 0) TC_TRACE(0);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % S;
        if (kc >= S) mbar_wait(empty0 + 8 * s, ((kc / S) - 1) & 1);
        const uint32_t fb = full0 + 8 * s;
        const bool skip_a = (p.flags & 1024) != 0;      // debug knob: measure the B feed alone
        TC_TRACE(1 + kc);
        mbar_expect_tx(fb, skip_a ? b_raw : tx_bytes);
        const int k0 = kc * TC_BK;
        if (MT == 2 && p.cluster2) {
          // each CTA of the pair fetches ONE of the two 128-channel weight tiles and multicasts it to both: the weights
          // are identical for every row tile, and re-streaming them per CTA made the mainloop L2-bandwidth bound
          const int mt = (int)cta_rank;
          tma_load_2d_mc(smem_u32(a_hi(s, mt)), &tmA_hi, fb, k0, m0 + mt * 128, (uint16_t)3);
          if (split) tma_load_2d_mc(smem_u32(a_lo(s, mt)), &tmA_lo, fb, k0, m0 + mt * 128, (uint16_t)3);
        } else {
#pragma unroll
          for (int mt = 0; mt < MT && !skip_a; ++mt) {
            tma_load_2d(smem_u32(a_hi(s, mt)), &tmA_hi, fb, k0, m0 + mt * 128);
            if (split) tma_load_2d(smem_u32(a_lo(s, mt)), &tmA_lo, fb, k0, m0 + mt * 128);
          }
        }
        tma_load_2d(smem_u32(b_hi(s)), &tmB, fb, k0, n0 - (MODE == 2 ? (p.fdw_K >> 1) : 0));     // fused depthwise forward: PAD rows of halo in front
      }
      if (MODE == 1) {
        // z tiles of the previous layer for the fused epilogue: wait until the tensor core has finished with the
        // pipeline stage(s) a tile overwrites (the MMA warp's commit on `empty` of the stage's last chunk)
        auto wait_stage_free = [&](int s) {
          if (s >= num_kc) return;                                   // stage never used
          const int kc_last = ((num_kc - 1 - s) / S) * S + s;         // last chunk that occupied stage s
          mbar_wait(empty0 + 8 * s, (uint32_t)(kc_last / S) & 1u);
        };
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (p.z_early) wait_stage_free((mt + 1) * (S >> 1) - 1);      // commits complete in order: the lower stages are free too
          else if (mt == 0) for (int s = 0; s < S; ++s) wait_stage_free(s);
          const uint32_t zb = z_bar0 + 8 * mt;
          mbar_expect_tx(zb, 4u * z_blk);
          const uint32_t dst = smem_u32(smem) + z_off(mt);
#pragma unroll
          for (int b = 0; b < 4; ++b) tma_load_2d(dst + b * z_blk, &tmZ, zb, m0 + mt * 128 + b * 32, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues (see gemm_tc2_kernel) =====
    {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_c = tc_idesc_bf16(idesc, p.corr);
      const int corr = p.corr;
      // Accumulator plan (p.nacc > 1 when the tile leaves TMEM columns free): the tensor core's fp32 accumulate truncates, an
      // error that grows with the number of MMAs chained into one accumulator (measured 1.8e-6 rms at K = 256, 1.1e-5 at
      // K = 1536 for 3 MMAs per 8 of K).  So the main products hi*hi of consecutive K ranges go to nacc - 1 separate
      // accumulators and the small correction products (2^-11 of the result, their own truncation is negligible) to the last
      // one; the epilogue adds them in fp32 with round-to-nearest.
      const int nacc = (split && p.nacc > 1) ? p.nacc : 1, nmain = nacc > 1 ? nacc - 1 : 1;
      int prev_am = -1;
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % S;
        const uint32_t ph = (kc / S) & 1;
        mbar_wait(full0 + 8 * s, ph);
        if (split) mbar_wait(ready0 + 8 * s, ph);
        tc_fence_after();
        const uint64_t dbh = umma_desc_k128(smem_u32(b_hi(s)));
        const uint64_t dbl = split ? umma_desc_k128(smem_u32(b_lo(s))) : 0;
        const int am = (kc * nmain) / num_kc;              // main accumulator of this K chunk
        const bool new_main = am != prev_am;
        prev_am = am;
        if (tc_elect_one()) {
        TC_TRACE(60 + kc);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint64_t dah = umma_desc_k128(smem_u32(a_hi(s, mt)));
          const uint64_t dal = split ? umma_desc_k128(smem_u32(a_lo(s, mt))) : 0;
          const uint32_t d = tmem_base + (uint32_t)(mt * 256 + am * BN);
          const uint32_t dc = tmem_base + (uint32_t)(mt * 256 + (nacc - 1) * BN);
          if (p.flags & 256) continue;                     // debug knob: no MMA issue
#pragma unroll
          for (int kk = 0; kk < TC_BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);       // 8 tf32 = 32 bytes = 2 x 16-byte units
            const uint32_t acc = (new_main && kk == 0) ? 0u : 1u;
            const uint32_t accc = (kc > 0 || kk > 0) ? 1u : 0u;
            if (!split) {
              tc_mma_tf32(d, dah + adv, dbh + adv, idesc, acc);
            } else if (nacc > 1) {
              tc_mma_tf32(d, dah + adv, dbh + adv, idesc, acc);
              if (corr) {
                tc_mma_bf16(dc, dal + adv, dbl + adv, idesc_c, accc);
              } else {
                tc_mma_tf32(dc, dal + adv, dbh + adv, idesc, accc);
                tc_mma_tf32(dc, dah + adv, dbl + adv, idesc, 1u);
              }
            } else if (corr) {
              tc_mma_tf32(d, dah + adv, dbh + adv, idesc, acc);
              tc_mma_bf16(d, dal + adv, dbl + adv, idesc_c, 1u);
            } else {
              tc_mma_tf32(d, dal + adv, dbh + adv, idesc, acc);
              tc_mma_tf32(d, dah + adv, dbl + adv, idesc, 1u);
              tc_mma_tf32(d, dah + adv, dbh + adv, idesc, 1u);
            }
          }
        }
        if (p.cluster2) tc_commit_mc(empty0 + 8 * s, (uint16_t)3);   // stage free (in both CTAs) once these MMAs have read it
        else tc_commit(empty0 + 8 * s);
        TC_TRACE(80 + kc);
        }
        __syncwarp();
      }
      if (tc_elect_one()) tc_commit(accum_bar);                  // accumulators complete
      __syncwarp();
    }
  } else {
    // ===== transform warps (split only), then epilogue =====
    const int tid = threadIdx.x - 64;        // 0..255
    if (MODE == 2) {
      // fused depthwise forward: per-channel parameters to shared memory ([scale | shift | bias | tap 0 | ... ] x Kd), then the
      // operand-producing mainloop (tc_dw_mainloop)
      float* par = reinterpret_cast<float*>(smem + p.par_off);
      const int Kd = p.Kd, Kt = p.fdw_K;
      for (int idx = tid; idx < Kd * (3 + Kt); idx += TC_EPI_THREADS) {
        const int f = idx / Kd, ch = idx - f * Kd;
        float v;
        if (f == 0) v = p.act.scale ? __ldg(p.act.scale + ch) : 1.f;
        else if (f == 1) v = p.act.scale ? __ldg(p.act.shift + ch) : 0.f;
        else if (f == 2) v = p.fdw_b ? __ldg(p.fdw_b + ch) : 0.f;
        else v = __ldg(p.fdw_w + (size_t)ch * Kt + (f - 3));
        par[idx] = v;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const uint32_t bh_off = nA * a_tile, bl_off = nA * a_tile + b_raw;
      const bool small = ((BN + 31) >> 5) <= 5;
#define TC_DW_CALL(KK, RR) tc_dw_mainloop<KK, RR>(p, smem, stage_bytes, bh_off, bl_off, full0, ready0, S, num_kc, n0, BN, par, tid)
      switch (p.fdw_K) {
        case 1: if (small) TC_DW_CALL(1, 5); else TC_DW_CALL(1, 8); break;
        case 3: if (small) TC_DW_CALL(3, 5); else TC_DW_CALL(3, 8); break;
        case 5: if (small) TC_DW_CALL(5, 5); else TC_DW_CALL(5, 8); break;
        default: if (small) TC_DW_CALL(7, 5); else TC_DW_CALL(7, 8); break;
      }
#undef TC_DW_CALL
    } else if (split) {
      const int n4 = BN * TC_BK / 4;
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % S;
        mbar_wait(full0 + 8 * s, (kc / S) & 1);
        float4* hi = reinterpret_cast<float4*>(b_hi(s));
        float4* lo = reinterpret_cast<float4*>(b_lo(s));
        for (int i = tid; i < n4; i += TC_EPI_THREADS) {
          const float4 v = hi[i];
          uint4 h, l;
          h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
          l.x = rna_tf32(v.x - __uint_as_float(h.x)); l.y = rna_tf32(v.y - __uint_as_float(h.y));
          l.z = rna_tf32(v.z - __uint_as_float(h.z)); l.w = rna_tf32(v.w - __uint_as_float(h.w));
          reinterpret_cast<uint4*>(hi)[i] = h;
          if (p.corr) tc_store_corr(reinterpret_cast<uint8_t*>(lo), (uint32_t)i >> 3, ((uint32_t)i & 7u) ^ (((uint32_t)i >> 3) & 7u), v, h, p.corr);
          else reinterpret_cast<uint4*>(lo)[i] = l;
        }
        fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(ready0 + 8 * s);
        if (tid == 0) TC_TRACE(20 + kc);
      }
    }
    mbar_wait(accum_bar, 0);
    if (tid == 0) TC_TRACE(100);
    tc_fence_after();
    const int quad = warp & 3;               // TMEM lanes 32*quad .. 32*quad+31 belong to this warp
    const int half = (warp - 2) >> 2;        // 0 / 1: which of the two warps of this quadrant
    const int nvalid = min(BN, p.R - n0);                       // rows of this tile inside the tensor
    if (MODE == 1) {
      const TnAct act = tn_act_init(p.act);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int c = m0 + mt * 128 + quad * 32 + lane;
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * 256);
        mbar_wait(z_bar0 + 8 * mt, 0);
        const uint8_t* zs = smem + z_off(mt) + (uint32_t)quad * z_blk;
        float* red = reinterpret_cast<float*>(smem + p.red_off);
        switch (p.dw_K) {
          case 1: tc_epilogue_dwbwd<1>(tbase, p, act, c, n0, half, zs, red); break;
          case 3: tc_epilogue_dwbwd<3>(tbase, p, act, c, n0, half, zs, red); break;
          case 5: tc_epilogue_dwbwd<5>(tbase, p, act, c, n0, half, zs, red); break;
          case 7: tc_epilogue_dwbwd<7>(tbase, p, act, c, n0, half, zs, red); break;
          case 9: tc_epilogue_dwbwd<9>(tbase, p, act, c, n0, half, zs, red); break;
          default: tc_epilogue_dwbwd<11>(tbase, p, act, c, n0, half, zs, red); break;
        }
      }
    } else
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int co = m0 + mt * 128 + quad * 32 + lane;
      const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
      const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * 256);
      float* zp = p.Z + (size_t)n0 * p.M_total + co;
      float s1 = 0.f, s2 = 0.f;
      const int na = (split && p.nacc > 1) ? p.nacc : 1;
      switch ((p.flags & 3) | (nvalid == BN ? 4 : 0)) {          // warp-uniform: one specialised, branch-free loop each
        case 4: tc_epilogue<false, false, true>(tbase, zp, (size_t)p.M_total, BN, nvalid, bv, s1, s2, half, 2, na, BN); break;
        case 0: tc_epilogue<false, false, false>(tbase, zp, (size_t)p.M_total, BN, nvalid, bv, s1, s2, half, 2, na, BN); break;
        case TN_EPI_TANH: case TN_EPI_TANH | 4: tc_epilogue<true, false, false>(tbase, zp, (size_t)p.M_total, BN, nvalid, bv, s1, s2, half, 2, na, BN); break;
        case TN_EPI_ACCUM | 4: tc_epilogue<false, true, true>(tbase, zp, (size_t)p.M_total, BN, nvalid, bv, s1, s2, half, 2, na, BN); break;
        case TN_EPI_ACCUM: tc_epilogue<false, true, false>(tbase, zp, (size_t)p.M_total, BN, nvalid, bv, s1, s2, half, 2, na, BN); break;
        default: tc_epilogue<true, true, false>(tbase, zp, (size_t)p.M_total, BN, nvalid, bv, s1, s2, half, 2, na, BN); break;
      }
      if (p.stats) {
        // the two warps of each lane quadrant are combined in shared memory (fixed order) and the CTA's per-channel partial
        // sums are added into the fixed-point accumulators (integer atomics: order-independent, see tn_fix_add)
        float* red = reinterpret_cast<float*>(smem + p.red_off);          // [2 halves][2 sums][128 channels]
        const int chl = (int)(threadIdx.x & 127u);
        red[(half * 2 + 0) * 128 + chl] = s1;
        red[(half * 2 + 1) * 128 + chl] = s2;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int which = tid >> 7, ch = tid & 127;                        // threads 0-127: sum, 128-255: sum of squares
        tn_fix_add(p.accum, p.M_total, which, (co - chl) + ch, red[which * 128 + ch] + red[(2 + which) * 128 + ch],
                   tn_fix_flag(p.accum, p.M_total, (int)blockIdx.y));
        asm volatile("bar.sync 1, 256;" ::: "memory");                    // `red` is reused by the next channel half
      }
    }
  }
  if (threadIdx.x == 64) TC_TRACE(101);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
  if (p.cluster2) cluster_sync_all();        // the peer may still signal this CTA's barriers until its MMAs have drained
  if (p.stats)                               // group = this CTA's 128 * MT channels; the pipeline memory is free now
    tn_stats_finish(p.has_bn ? &p.bn : nullptr, p.stats, p.accum, p.M_total, m0, 128 * MT, p.tickets + blockIdx.y,
                    tn_fix_flag(p.accum, p.M_total, (int)blockIdx.y), gridDim.x, blockIdx.y == 0, p.fc);
  if (threadIdx.x == 0 && p.trace && blockIdx.y == 0 && blockIdx.x < 256) {
    unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[257 + 2 * blockIdx.x] = (long long)gt; }
  if (threadIdx.x == 0) { TC_TRACE(102); if (p.trace && blockIdx.y == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2)) {
      unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[(blockIdx.x == 0 ? 0 : 128) + 112] = (long long)gt; } }
}

// ---------------------------------------------------------------------------
// Pair kernel: the same GEMM issued as tcgen05.mma.cta_group::2 by 2-CTA clusters.
//
// A pair of CTAs (one cluster) owns M = 256 output channels x 2 x BN2 rows.  CTA r holds HALF of every operand:
// weight rows [128 r, 128 r + 128) (hi + lo) and, of each of the two BN2-row tiles, rows [r BN2/2, (r+1) BN2/2).  One
// thread of the even ("leader") CTA issues M = 256 MMAs that read both CTAs' shared memory and write both CTAs' TMEM
// (each CTA ends up with its 128 channels x all 2 BN2 rows).  Per SM that is 50 KB of TMA traffic per K chunk instead
// of 84 KB (the single-CTA kernel re-streams the whole 64 KB weight chunk into every SM and is paced by the per-SM TMA
// rate), and the 68 KB stage leaves room for three stages instead of two.
//
// Barriers (same offsets in both CTAs): fullA (leader's is used: both CTAs' weight loads complete_tx on it through the
// .cta_group::2 TMA form), fullB (local: this CTA's activation rows, waited by the local transform warps), ready
// (leader's: one arrival per transform warp of BOTH CTAs, remote arrivals through mapa), empty and accum (local,
// signalled in both CTAs by the leader's multicast tcgen05.commit).
// ---------------------------------------------------------------------------
#define TC2_STAGES 3
// debug timeline of the pair kernel (tn_gemm_tc_set_trace): globaltimer (ns) of CTA 0 -> slots [0, 64), of CTA gridDim.x / 2 (rounded
// to its pair's leader) -> slots [64, 128)
#define TC2_TRACE(slot) do { if (p.trace && blockIdx.y == 0 && (blockIdx.x == 0 || blockIdx.x == ((gridDim.x / 2) & ~1u))) { \
    unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); \
    p.trace[(blockIdx.x == 0 ? 0 : 64) + (slot)] = (long long)gt_; } } while (0)
// per-chunk events of CTA 0 -> slots [256, 512): 16 slots per event kind (producer issue, operand arrived, transform done, MMA waits
// passed, MMAs issued)
#define TC2_TRACE2(kind, kc) do { if (p.trace && blockIdx.y == 0 && blockIdx.x == 0 && (kc) < 16) { \
    unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); \
    p.trace[256 + 16 * (kind) + (kc)] = (long long)gt_; } } while (0)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_leader, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar_leader), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_local_addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar_local_addr), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// The same without the cluster-scope release fence (MEMBAR.ALL.GPU in SASS, ~1.2 us per chunk on the mainloop's critical path,
// measured with the per-chunk timeline).  Used to forward "this CTA's operand tile is in SHARED memory": the tile's writers ran
// fence.proxy.async after their stores and arrived (release) on the CTA-local barrier this thread has just acquired, so the
// stores are complete before this arrive is issued; there is no global-memory traffic to order.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar_local_addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar_local_addr), "r"(cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// Fused depthwise-forward operand producer of the pair kernel (MODE 2, 16 transform warps = 512 threads): like
// tc_dw_mainloop, but this CTA holds HB = BN2/2 rows of each of the two N tiles.  Per K chunk TMA delivers, per tile, the raw z
// rows [first row - PAD, + HB + 8) into the B_hi buffer; thread (q = channel quad, tile t, row lane 0..31) owns RP consecutive
// rows of tile t.  Same two phases (window to registers | barrier | BN/ReLU/dropout, FIR, u side output, hi/lo split in place).
template <int K, int RP>
__device__ __forceinline__ void tc2_dw_mainloop(const TcParams& p, uint8_t* smem, uint32_t stage_bytes, uint32_t bh_off, uint32_t bh_tile,
                                                uint32_t bl_off, uint32_t bl_tile, uint32_t fullB0, uint32_t ready0, int num_kc,
                                                int row_base, int BN2, int HB, const float* par, int tid, int lane) {
  constexpr int PAD = K / 2;
  constexpr int W = RP + 2 * PAD;
  constexpr int S = TC2_STAGES;
  const TnAct act = tn_act_init(p.act);
  const int C = p.Kd, T = p.fdw_T, R = p.R;
  const bool lazy = act.scale != nullptr;
  const bool drop = act.thresh != 0;
  const uint32_t key = tn_hash_key32(act.seed_lo, act.seed_hi, act.layer);
  const uint32_t q = (uint32_t)tid & 7u;
  const int rl = tid >> 3;                                   // 0..63
  const int t = rl >> 5;                                     // N tile
  const int rpa = (HB + 31) >> 5;                            // rows per lane actually used (<= RP)
  const int j0 = (rl & 31) * rpa;
  const int rp = min(rpa, HB - j0);                          // may be <= 0 for the last lanes
  const int g0 = row_base + t * BN2 + j0;                    // global row of this thread's first output row
  const int t0 = (int)(((long long)g0 % T + T) % T);
  const bool interior = rp > 0 && g0 >= 0 && t0 - PAD >= 0 && t0 + rp - 1 + PAD < T;
  uint32_t vmask = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const int grow = g0 - PAD + i;
    if (grow >= 0 && grow < R) vmask |= 1u << i;
  }
  for (int kc = 0; kc < num_kc; ++kc) {
    const int s = kc % S;
    uint8_t* bh = smem + (size_t)s * stage_bytes + bh_off + (size_t)t * bh_tile;
    uint8_t* bl = smem + (size_t)s * stage_bytes + bl_off + (size_t)t * bl_tile;
    const int c = kc * TC_BK + 4 * (int)q;
    const float4 sc4 = *reinterpret_cast<const float4*>(par + c);
    const float4 sh4 = *reinterpret_cast<const float4*>(par + C + c);
    const float4 b4 = *reinterpret_cast<const float4*>(par + 2 * C + c);
    mbar_wait(fullB0 + 8 * s, (kc / S) & 1);
    float4 win[W], uo[RP];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const uint32_t r = (uint32_t)(j0 + i);                 // raw row r <-> global row (tile's first row) - PAD + r
      if (i < rp + 2 * PAD) win[i] = *reinterpret_cast<const float4*>(bh + r * 128u + ((q ^ (r & 7u)) << 4));
      else win[i] = tn_zero4();
    }
    asm volatile("bar.sync 2, 512;" ::: "memory");          // every thread has its window: the raw tiles may be overwritten
    if (rp > 0) {
      if (lazy) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
          float4 v = make_float4(fmaf(win[i].x, sc4.x, sh4.x), fmaf(win[i].y, sc4.y, sh4.y), fmaf(win[i].z, sc4.z, sh4.z),
                                 fmaf(win[i].w, sc4.w, sh4.w));
          float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
          if (drop) {
            const uint32_t pr = ((uint32_t)(g0 - PAD + i) * (uint32_t)C + (uint32_t)c) >> 1;   // rows < 0 wrap: masked below
            const uint32_t h0 = tn_hash_elem32(key, pr), h1 = tn_hash_elem32(key, pr + 1u);
            m = make_float4((h0 & 0xFFFFu) >= act.thresh ? act.inv_keep : 0.f, (h0 >> 16) >= act.thresh ? act.inv_keep : 0.f,
                            (h1 & 0xFFFFu) >= act.thresh ? act.inv_keep : 0.f, (h1 >> 16) >= act.thresh ? act.inv_keep : 0.f);
          }
          if (act.relu) {
            m.x = v.x > 0.f ? m.x : 0.f; m.y = v.y > 0.f ? m.y : 0.f;
            m.z = v.z > 0.f ? m.z : 0.f; m.w = v.w > 0.f ? m.w : 0.f;
          }
          win[i] = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
        }
        if (vmask != (1u << W) - 1u) {
#pragma unroll
          for (int i = 0; i < W; ++i)
            if (!((vmask >> i) & 1u)) win[i] = tn_zero4();
        }
      }
      float4 wk[K];
#pragma unroll
      for (int k = 0; k < K; ++k) wk[k] = *reinterpret_cast<const float4*>(par + (3 + k) * C + c);
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        if (i < rp) {
          const int j = j0 + i;
          float4 acc = b4;
          if (interior) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc = tn_fma4(wk[k], win[i + k], acc);
          } else {
            const int tt0 = (t0 + i) % T;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const int tt = tt0 + k - PAD;
              if (tt >= 0 && tt < T) acc = tn_fma4(wk[k], win[i + k], acc);
            }
          }
          uo[i] = acc;
          uint4 h, l;
          h.x = rna_tf32(acc.x); h.y = rna_tf32(acc.y); h.z = rna_tf32(acc.z); h.w = rna_tf32(acc.w);
          l.x = rna_tf32(acc.x - __uint_as_float(h.x)); l.y = rna_tf32(acc.y - __uint_as_float(h.y));
          l.z = rna_tf32(acc.z - __uint_as_float(h.z)); l.w = rna_tf32(acc.w - __uint_as_float(h.w));
          const uint32_t o = (uint32_t)j * 128u + ((q ^ ((uint32_t)j & 7u)) << 4);
          *reinterpret_cast<uint4*>(bh + o) = h;
          if (p.corr) tc_store_corr(reinterpret_cast<uint8_t*>(bl), (uint32_t)j, (uint32_t)q, acc, h, p.corr);
          else *reinterpret_cast<uint4*>(bl + o) = l;
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(ready0 + 8 * s);               // one arrival per warp on this CTA's barrier
    // the u side output leaves after the arrival, off the critical path of the chunk
    if (p.fdw_u && rp > 0) {
#pragma unroll
      for (int i = 0; i < RP; ++i)
        if (i < rp && g0 + i >= 0 && g0 + i < R) tn_st4(p.fdw_u + (size_t)(g0 + i) * C + c, uo[i]);
    }
  }
}

// MODE 0: plain epilogue (+ statistics / BatchNorm fold), 1: fused depthwise-backward epilogue, 2: fused depthwise-forward
// operand producer (plain epilogue).  EW = transform / epilogue warps: 8, or 12 / 16 for the fused modes with K <= 3 (their
// epilogue / producer loops are issue-bound with two warps per scheduler).
template <int MODE, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmZ,
                const __grid_constant__ CUtensorMap tmG, TcParams p) {
  tn_grid_dep_sync();
  if (threadIdx.x == 0) TC2_TRACE(0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[5 * TC2_STAGES + 1 + 2 * TC2_STAGES];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN2 = p.BN, HB = BN2 >> 1;                 // rows per N tile, rows of it held by this CTA
  constexpr int S = TC2_STAGES;
  const uint32_t rank = cluster_cta_rank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  // first row of the pair's tile; fused depthwise backward: BNo output rows per pair, the tile starts PAD rows early
  const int n0 = MODE == 1 ? pair * p.BNo - (p.dw_K >> 1) : pair * 2 * BN2;
  const int m0 = blockIdx.y * 256;                      // first output channel of the pair
  const int num_kc = p.Kd / TC_BK;
  const uint32_t a_tile = 128 * TC_BK * 4;              // 16 KiB: this CTA's 128 weight rows
  const uint32_t b_half = (uint32_t)HB * TC_BK * 4;     // this CTA's rows of one N tile
  // MODE 2: TMA delivers the RAW z rows with 8 extra rows (halo of the K-tap FIR) into the B_hi buffers
  const uint32_t b_raw = b_half + (MODE == 2 ? 8u * TC_BK * 4 : 0u);
  const uint32_t stage_bytes = 2 * a_tile + 2 * b_raw + 2 * b_half;   // A_hi | A_lo | B_hi[2] (raw) | B_lo[2]
  auto a_hi = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto a_lo = [&](int s) { return smem + (size_t)s * stage_bytes + a_tile; };
  auto b_hi = [&](int s, int t) { return smem + (size_t)s * stage_bytes + 2 * a_tile + t * b_raw; };
  auto b_lo = [&](int s, int t) { return smem + (size_t)s * stage_bytes + 2 * a_tile + 2 * b_raw + t * b_half; };
  const uint32_t fullA0 = smem_u32(&bars[0]), fullB0 = smem_u32(&bars[S]), ready0 = smem_u32(&bars[2 * S]),
                 empty0 = smem_u32(&bars[3 * S]), accum_bar = smem_u32(&bars[4 * S]), zbar0 = smem_u32(&bars[4 * S + 1]),
                 readyP0 = smem_u32(&bars[5 * S + 1]);
  // `ready` collects the LOCAL transform warps (cheap cta-scope arrivals).  The peer's otherwise idle warp 1 forwards its CTA's
  // completion with ONE remote arrival per chunk on the leader's `readyP`: a remote arrive carries a gpu-scope release fence
  // (MEMBAR.ALL.GPU in SASS, profiled as stall_membar), which must stay off the transform warps' critical path.
  // fused depthwise backward: the z tile of this CTA's 128 channels = 8 boxes [BN2 rows x 32 channels] (box b = 2 * channel
  // block + N tile), loaded into pipeline stages in the order the mainloop releases them: `zcap` boxes per stage, group g
  // of boxes -> stage (num_kc - S + g) % S, one barrier per group
  const uint32_t z_box = (uint32_t)BN2 * 128u;
  const int zcap = (int)(stage_bytes / z_box);
  auto z_stage = [&](int g) { return num_kc >= S ? (num_kc - S + g) % S : g; };
  auto z_ptr = [&](int b) { return smem + (size_t)z_stage(b / zcap) * stage_bytes + (size_t)(b % zcap) * z_box; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(fullA0 + 8 * s, 1);
      mbar_init(fullB0 + 8 * s, 1);
      mbar_init(ready0 + 8 * s, EW);                    // this CTA's transform warps
      mbar_init(readyP0 + 8 * s, 1);                    // leader only: the peer's forwarded completion
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    for (int g = 0; g < S; ++g) mbar_init(zbar0 + 8 * g, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) TC2_TRACE(1);

  if (warp == 0) {
    // ===== TMA producer (both CTAs): the whole warp runs the loops, one elected lane issues (uniform operands, see the MMA issuer) =====
    {
      uint32_t fullA_leader0;                                       // the same barrier in the even (leader) CTA of the pair
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(fullA_leader0) : "r"(fullA0), "r"(0u));
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % S;
        if (kc >= S) mbar_wait(empty0 + 8 * s, ((kc / S) - 1) & 1);
        const int k0 = kc * TC_BK;
        if (tc_elect_one()) {
        if (leader) mbar_expect_tx(fullA0 + 8 * s, 4 * a_tile);   // both CTAs' hi + lo weight tiles
        tma_load_2d_2sm(smem_u32(a_hi(s)), &tmA_hi, fullA_leader0 + 8 * s, k0, m0 + (int)rank * 128);
        tma_load_2d_2sm(smem_u32(a_lo(s)), &tmA_lo, fullA_leader0 + 8 * s, k0, m0 + (int)rank * 128);
        const int padr = MODE == 2 ? (p.fdw_K >> 1) : 0;           // fused depthwise forward: PAD rows of halo in front
        mbar_expect_tx(fullB0 + 8 * s, 2 * b_raw + (p.has_bnb ? 2 * b_half : 0u));
        tma_load_2d(smem_u32(b_hi(s, 0)), &tmB, fullB0 + 8 * s, k0, n0 + (int)rank * HB - padr);
        tma_load_2d(smem_u32(b_hi(s, 1)), &tmB, fullB0 + 8 * s, k0, n0 + BN2 + (int)rank * HB - padr);
        if (MODE != 2 && p.has_bnb) {          // BatchNorm backward as operand producer: the z rows of the same box into B_lo
          tma_load_2d(smem_u32(b_lo(s, 0)), &tmG, fullB0 + 8 * s, k0, n0 + (int)rank * HB);
          tma_load_2d(smem_u32(b_lo(s, 1)), &tmG, fullB0 + 8 * s, k0, n0 + BN2 + (int)rank * HB);
        }
        TC2_TRACE2(0, kc);
        }
        __syncwarp();
      }
      if (MODE == 1) {
        // z boxes of the previous layer for the fused epilogue, group by group as the tensor core releases the stages
        // (this CTA's `empty` barriers are signalled by the leader's multicast commit)
        const int ngroups = (8 + zcap - 1) / zcap;
        for (int g = 0; g < ngroups; ++g) {
          if (num_kc >= S) {
            const int st = z_stage(g);
            const int kc_last = ((num_kc - 1 - st) / S) * S + st;
            mbar_wait(empty0 + 8 * st, (uint32_t)(kc_last / S) & 1u);
          } else if (g == 0) {
            for (int st = 0; st < num_kc; ++st) mbar_wait(empty0 + 8 * st, 0u);
          }
          const int b0 = g * zcap, b1 = min(8, b0 + zcap);
          if (tc_elect_one()) {
            mbar_expect_tx(zbar0 + 8 * g, (uint32_t)(b1 - b0) * z_box);
            for (int b = b0; b < b1; ++b)
              tma_load_2d(smem_u32(z_ptr(b)), &tmZ, zbar0 + 8 * g, m0 + (int)rank * 128 + (b >> 1) * 32, n0 + (b & 1) * BN2);
          }
          __syncwarp();
        }
      }
    }
    if (MODE != 1 && p.stats) {
      // ===== statistics ticket + BatchNorm fold, by this otherwise idle warp, WHILE the epilogue warps store the tile =====
      // The epilogue warps first run a statistics-only pass over the accumulators, add their per-channel sums into the
      // fixed-point accumulators and arrive on named barrier 3.  Fence + ticket + (last CTA of the group:) read-back, fold and
      // stores are a chain of ~3 global round trips (~2.5 us as a serial tail of the kernel, measured); here they overlap the
      // ~2.5 us the other warps spend writing Z.
      __syncwarp();                            // lane 0 comes out of the producer loop: reconverge before the named barrier
      if (lane == 0) TC2_TRACE(10);
      asm volatile("bar.sync 3, %0;" ::"n"(32 * EW + 32) : "memory");
      if (lane == 0) TC2_TRACE(11);
      const int grp = 2 * (int)blockIdx.y + (int)rank;
      unsigned int last = 0;
      if (lane == 0) {
        __threadfence();
        TC2_TRACE(12);
        const unsigned int t = atomicAdd(p.tickets + grp, 1u);
        last = (t == (gridDim.x >> 1) - 1) ? 1u : 0u;
        if (last) p.tickets[grp] = 0u;
        TC2_TRACE(13);
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        tn_stats_fold_channels(p.has_bn ? &p.bn : nullptr, p.stats, p.accum, p.M_total, m0 + (int)rank * 128, 128, tn_fix_flag(p.accum, p.M_total, grp),
                               blockIdx.y == 0 && rank == 0, lane, 32, p.fc);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    // The WHOLE warp runs the loop (waits, descriptors: warp-uniform values in uniform registers) and one elected lane issues.
    // Issued from `if (lane == 0)` code, every tcgen05.mma was wrapped by the compiler in an ELECT / BRA.U.ANY serialisation loop
    // with its operands moved from vector to uniform registers each time: 105 cycles per M = 256, N = 144 MMA against 72 from
    // uniform code (tools/utccp_test.cu) -- the mainloop was bound by the issuing thread, not by the tensor pipe.
    if (leader) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t idesc_c = tc_idesc_bf16(idesc, p.corr);
      // accumulator plan as in gemm_tc_kernel: nacc - 1 K-chunked main accumulators + one for the corrections, BN2 columns apart
      const int nacc = (MODE != 1 && p.nacc > 1) ? p.nacc : 1, nmain = nacc > 1 ? nacc - 1 : 1;
      const int corr = p.corr;
      int prev_am = -1;
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % S;
        const uint32_t ph = (kc / S) & 1;
        mbar_wait(fullA0 + 8 * s, ph);
        mbar_wait(ready0 + 8 * s, ph);
        mbar_wait(readyP0 + 8 * s, ph);
        tc_fence_after();
        const uint64_t dah = umma_desc_k128(smem_u32(a_hi(s))), dal = umma_desc_k128(smem_u32(a_lo(s)));
        const int am = (kc * nmain) / num_kc;
        const bool new_main = am != prev_am;
        prev_am = am;
        if (tc_elect_one()) {
          TC2_TRACE2(3, kc);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint64_t dbh = umma_desc_k128(smem_u32(b_hi(s, t))), dbl = umma_desc_k128(smem_u32(b_lo(s, t)));
            const uint32_t d = tmem_base + (uint32_t)(t * 256 + am * BN2);
            const uint32_t dc = tmem_base + (uint32_t)(t * 256 + (nacc - 1) * BN2);
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 2);
              const uint32_t acc = (new_main && kk == 0) ? 0u : 1u;
              const uint32_t accc = (kc > 0 || kk > 0) ? 1u : 0u;
              if (nacc > 1) {
                tc_mma_tf32_2sm(d, dah + adv, dbh + adv, idesc, acc);
                if (corr) {
                  tc_mma_bf16_2sm(dc, dal + adv, dbl + adv, idesc_c, accc);
                } else {
                  tc_mma_tf32_2sm(dc, dal + adv, dbh + adv, idesc, accc);
                  tc_mma_tf32_2sm(dc, dah + adv, dbl + adv, idesc, 1u);
                }
              } else if (corr) {
                tc_mma_tf32_2sm(d, dah + adv, dbh + adv, idesc, acc);
                tc_mma_bf16_2sm(d, dal + adv, dbl + adv, idesc_c, 1u);
              } else {
                tc_mma_tf32_2sm(d, dal + adv, dbh + adv, idesc, acc);
                tc_mma_tf32_2sm(d, dah + adv, dbl + adv, idesc, 1u);
                tc_mma_tf32_2sm(d, dah + adv, dbh + adv, idesc, 1u);
              }
            }
          }
          tc_commit_2sm(empty0 + 8 * s, (uint16_t)3);       // stage free in both CTAs once these MMAs have read it
          TC2_TRACE2(4, kc);
        }
        __syncwarp();
      }
      if (tc_elect_one()) tc_commit_2sm(accum_bar, (uint16_t)3);              // accumulators complete (both CTAs)
      __syncwarp();
    } else if (!leader && lane == 0) {
      for (int kc = 0; kc < num_kc; ++kc) {               // forward this CTA's operand-ready events to the leader
        const int s = kc % S;
        mbar_wait(ready0 + 8 * s, (kc / S) & 1);
        mbar_arrive_cluster_relaxed(readyP0 + 8 * s, 0u);
      }
    }
  } else {
    // ===== transform warps (hi / lo split of this CTA's activation rows), then epilogue =====
    const int tid = threadIdx.x - 64;
    const int n4 = 2 * HB * TC_BK / 4;                    // both N tiles are contiguous: B_hi[0] | B_hi[1]
    if (MODE == 2) {
      // fused depthwise forward: per-channel parameters to shared memory, then the operand-producing mainloop
      float* par = reinterpret_cast<float*>(smem + p.par_off);
      const int Kd = p.Kd, Kt = p.fdw_K;
      for (int idx = tid; idx < Kd * (3 + Kt); idx += 32 * EW) {
        const int f = idx / Kd, ch = idx - f * Kd;
        float v;
        if (f == 0) v = p.act.scale ? __ldg(p.act.scale + ch) : 1.f;
        else if (f == 1) v = p.act.scale ? __ldg(p.act.shift + ch) : 0.f;
        else if (f == 2) v = p.fdw_b ? __ldg(p.fdw_b + ch) : 0.f;
        else v = __ldg(p.fdw_w + (size_t)ch * Kt + (f - 3));
        par[idx] = v;
      }
      asm volatile("bar.sync 2, 512;" ::: "memory");
      const uint32_t bh_off = 2 * a_tile, bl_off = 2 * a_tile + 2 * b_raw;
      const int row_base = n0 + (int)rank * HB;
      if (p.fdw_K == 1) tc2_dw_mainloop<1, 4>(p, smem, stage_bytes, bh_off, b_raw, bl_off, b_half, fullB0, ready0, num_kc, row_base, BN2, HB, par, tid, lane);
      else tc2_dw_mainloop<3, 4>(p, smem, stage_bytes, bh_off, b_raw, bl_off, b_half, fullB0, ready0, num_kc, row_base, BN2, HB, par, tid, lane);
    } else if (MODE != 2 && p.has_bnb) {
      // ===== BatchNorm backward as the operand producer (train-mode BN behind the conv whose data gradient this GEMM is) =====
      // g = dZ + a[c] + b[c] z with the per-channel statistics-path coefficients a, b (tn_bn_stats_bwd's arithmetic), computed
      // here from (dscale, dshift, mean, invstd, gamma); dZ arrives in B_hi, z in B_lo; g is rounded / split in place, written to
      // g_out for the weight-gradient GEMM (rows this pair owns, channel group 0 only).  Replaces one tn_bn_stats_bwd launch
      // (read dZ, z; write g) per conv.
      constexpr int NT = 32 * EW;
      constexpr int NI = (2 * 128 * 8 + NT - 1) / NT;              // float4s per thread and chunk (BN2 <= 256: HB <= 128)
      float* sa = reinterpret_cast<float*>(smem + p.par_off);
      float* sb = sa + p.Kd;
      const tn_bn_bwd& q = p.bnb;
      const bool owner = blockIdx.y == 0;
      for (int ch = tid; ch < p.Kd; ch += NT) {
        const double dsc = __ldg(q.dscale + ch), dsh = __ldg(q.dshift + ch), mu = __ldg(q.mean + ch), r = __ldg(q.invstd + ch), gm = __ldg(q.gamma + ch);
        const double t = dsc - mu * dsh;                           // dL/d(invstd) / gamma
        const double dvar = -0.5 * gm * t * r * r * r;
        const double dmu = -dsh * gm * r - 2.0 * mu * dvar;
        sa[ch] = (float)(dmu / q.n);
        sb[ch] = (float)(2.0 * dvar / q.n);
        if (owner && blockIdx.x == 0) {                            // rank 0 of pair 0
          q.dgamma[ch] = (float)(r * t);
          q.dbeta[ch] = (float)dsh;
          // The conv bias in front of a train-mode BatchNorm has NO gradient: sum_r g = sum_r dZ - gamma invstd dshift, and
          // the consumers of the lazy activation produce dZ = (dL/dpre) * scale with dshift = sum_r dL/dpre, so the two terms
          // cancel exactly (the reference's value is the rounding noise of that cancellation).  Written as 0.
          if (q.dbias) q.dbias[ch] = 0.f;
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");
      // rows this pair owns (fused depthwise backward: the tile carries a halo that the neighbouring pairs own)
      const int own_lo = MODE == 1 ? pair * p.BNo : n0;
      const int own_hi = min(p.R, MODE == 1 ? own_lo + p.BNo : n0 + 2 * BN2);
      const uint32_t cq = ((uint32_t)tid & 7u) ^ (((uint32_t)tid >> 3) & 7u);      // this thread's 16-byte chunk: fixed (NT % 64 == 0)
      for (int kc = 0; kc < num_kc; ++kc) {
        const int s = kc % S;
        mbar_wait(fullB0 + 8 * s, (kc / S) & 1);
        if (tid == 0) TC2_TRACE2(1, kc);
        float4* hi = reinterpret_cast<float4*>(b_hi(s, 0));
        float4* lo = reinterpret_cast<float4*>(b_lo(s, 0));
        float4 v[NI], zz[NI];
#pragma unroll
        for (int k = 0; k < NI; ++k) {
          const int i = tid + k * NT;
          if (i < n4) { v[k] = hi[i]; zz[k] = lo[i]; }
        }
        // B_lo is overwritten below with the packed correction rows.  A thread's stores only touch the 128-byte row(s) it read
        // from, and the eight 16-byte chunks of a row are read by eight consecutive threads of ONE warp in the same
        // iteration k: a warp-level barrier is enough (a CTA-wide one per chunk cost more than the arithmetic).
        __syncwarp();
        const int ch = kc * TC_BK + 4 * (int)cq;
        const float4 a4 = *reinterpret_cast<const float4*>(sa + ch), b4 = *reinterpret_cast<const float4*>(sb + ch);
#pragma unroll
        for (int k = 0; k < NI; ++k) {
          const int i = tid + k * NT;
          if (i < n4) {
            const int row = i >> 3;                                // 0 .. 2 HB - 1 over both N tiles
            const int t = row >= HB ? 1 : 0;
            const int gr = n0 + t * BN2 + (int)rank * HB + (row - t * HB);
            float4 g = tn_zero4();
            if (gr >= 0 && gr < p.R) g = tn_fma4(b4, zz[k], v[k] + a4);
            if (owner && gr >= own_lo && gr < own_hi) tn_st4(q.g_out + (size_t)gr * p.Kd + ch, g);
            uint4 h;
            h.x = rna_tf32(g.x); h.y = rna_tf32(g.y); h.z = rna_tf32(g.z); h.w = rna_tf32(g.w);
            reinterpret_cast<uint4*>(hi)[i] = h;
            if (p.corr) {
              tc_store_corr(reinterpret_cast<uint8_t*>(lo), (uint32_t)row, cq, g, h, p.corr);
            } else {
              uint4 l;
              l.x = rna_tf32(g.x - __uint_as_float(h.x)); l.y = rna_tf32(g.y - __uint_as_float(h.y));
              l.z = rna_tf32(g.z - __uint_as_float(h.z)); l.w = rna_tf32(g.w - __uint_as_float(h.w));
              reinterpret_cast<uint4*>(lo)[i] = l;
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(ready0 + 8 * s);
        if (tid == 0) TC2_TRACE2(2, kc);
      }
    } else
    for (int kc = 0; kc < num_kc; ++kc) {
      const int s = kc % S;
      mbar_wait(fullB0 + 8 * s, (kc / S) & 1);
      if (tid == 0) TC2_TRACE2(1, kc);
      float4* hi = reinterpret_cast<float4*>(b_hi(s, 0));
      float4* lo = reinterpret_cast<float4*>(b_lo(s, 0));
      // this thread's float4s of the chunk are all loaded before the first is split (the stage cycle TMA -> split -> MMA -> release
      // bounds the mainloop once the MMAs issue at the tensor pipe's rate; one load at a time cost ~0.3 us of it per chunk)
      constexpr int NTT = 32 * EW, NIT = (2 * 128 * 8 + NTT - 1) / NTT;       // HB <= 128
      float4 vv[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k)
        if (tid + k * NTT < n4) vv[k] = hi[tid + k * NTT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int i = tid + k * NTT;
        if (i < n4) {
          const float4 v = vv[k];
          uint4 h;
          h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
          reinterpret_cast<uint4*>(hi)[i] = h;
          if (p.corr) {
            tc_store_corr(reinterpret_cast<uint8_t*>(lo), (uint32_t)i >> 3, ((uint32_t)i & 7u) ^ (((uint32_t)i >> 3) & 7u), v, h, p.corr);
          } else {
            uint4 l;
            l.x = rna_tf32(v.x - __uint_as_float(h.x)); l.y = rna_tf32(v.y - __uint_as_float(h.y));
            l.z = rna_tf32(v.z - __uint_as_float(h.z)); l.w = rna_tf32(v.w - __uint_as_float(h.w));
            reinterpret_cast<uint4*>(lo)[i] = l;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ready0 + 8 * s);                 // one arrival per warp on this CTA's barrier
      if (tid == 0) TC2_TRACE2(2, kc);
    }
    if (tid == 0) TC2_TRACE(2);
    mbar_wait(accum_bar, 0);
    if (tid == 0) TC2_TRACE(3);
    tc_fence_after();
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int co = m0 + (int)rank * 128 + quad * 32 + lane;
    if (MODE == 1) {
      const TnAct act = tn_act_init(p.act);
      const int ba = 2 * quad, bb = 2 * quad + 1;                      // this channel block's two z boxes
      mbar_wait(zbar0 + 8 * (ba / zcap), 0);
      mbar_wait(zbar0 + 8 * (bb / zcap), 0);
      if (tid == 0) TC2_TRACE(20);
      const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16);
      float* red = reinterpret_cast<float*>(smem + p.red_off);
      constexpr int NP = EW / 4, NT = 32 * EW;
      if (EW > 12) {                                          // 113-register budget: narrow windows only (host: K <= 3)
        if (p.dw_K == 1) tc_epilogue_dwbwd<1>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT);
        else tc_epilogue_dwbwd<3>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT);
      } else if (EW > 8) {                                    // 146 registers: windows up to K = 7 (host)
        switch (p.dw_K) {
          case 1: tc_epilogue_dwbwd<1>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
          case 3: tc_epilogue_dwbwd<3>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
          case 5: tc_epilogue_dwbwd<5>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
          default: tc_epilogue_dwbwd<7>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
        }
      } else switch (p.dw_K) {
        case 1: tc_epilogue_dwbwd<1>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
        case 3: tc_epilogue_dwbwd<3>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
        case 5: tc_epilogue_dwbwd<5>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
        case 7: tc_epilogue_dwbwd<7>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
        case 9: tc_epilogue_dwbwd<9>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
        default: tc_epilogue_dwbwd<11>(tbase, p, act, co, n0, half, z_ptr(ba), red, BN2, z_ptr(bb), NP, NT); break;
      }
    } else {
    const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
    const int na = p.nacc > 1 ? p.nacc : 1;
    if (p.stats) {
      // statistics first (no stores yet), so that the ticket / fold chain of warp 0 overlaps the store pass below
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int nvalid = min(BN2, p.R - (n0 + t * BN2));
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * 256);
        if (nvalid == BN2) tc_epilogue_stats<true>(tbase, BN2, nvalid, bv, s1, s2, half, EW / 4, na, BN2);
        else if (nvalid > 0) tc_epilogue_stats<false>(tbase, BN2, nvalid, bv, s1, s2, half, EW / 4, na, BN2);
      }
      // the warps of each lane quadrant are combined in shared memory in a fixed order; the CTA's per-channel partial sums
      // are added into the fixed-point accumulators (integer atomics: order-independent, see tn_fix_add)
      if (tid == 0) TC2_TRACE(4);
      float* red = reinterpret_cast<float*>(smem + p.red_off);          // [EW/4 parts][2 sums][128 channels]
      const int chl = (int)(threadIdx.x & 127u);
      red[(half * 2 + 0) * 128 + chl] = s1;
      red[(half * 2 + 1) * 128 + chl] = s2;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      if (tid < 256) {
        const int which = tid >> 7, ch = tid & 127;
        float acc = 0.f;
#pragma unroll
        for (int pp = 0; pp < EW / 4; ++pp) acc += red[(pp * 2 + which) * 128 + ch];
        tn_fix_add(p.accum, p.M_total, which, (co - chl) + ch, acc, tn_fix_flag(p.accum, p.M_total, 2 * (int)blockIdx.y + (int)rank));
      }
      if (tid == 0) TC2_TRACE(5);
      asm volatile("bar.arrive 3, %0;" ::"n"(32 * EW + 32) : "memory");      // warp 0 takes it from here
    }
    if (tid == 0) TC2_TRACE(6);
    float d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int r0 = n0 + t * BN2;
      const int nvalid = min(BN2, p.R - r0);
      const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * 256);
      float* zp = p.Z + (size_t)r0 * p.M_total + co;
      if (nvalid == BN2) tc_epilogue<false, false, true>(tbase, zp, (size_t)p.M_total, BN2, nvalid, bv, d1, d2, half, EW / 4, na, BN2);
      else if (nvalid > 0) tc_epilogue<false, false, false>(tbase, zp, (size_t)p.M_total, BN2, nvalid, bv, d1, d2, half, EW / 4, na, BN2);
    }
    }
  }
  if (threadIdx.x == 64) TC2_TRACE(7);
  if (threadIdx.x == 0) TC2_TRACE(14);
  tc_fence_before();
  cluster_sync_all();                                     // the peer reads this CTA's smem / signals its barriers until its MMAs drained
  if (threadIdx.x == 0) TC2_TRACE(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Split scheme per GEMM.  The weight split carries every format (ws = [4, M, Kd]: tf32 hi | tf32 lo | packed bf16 correction
// rows | packed scaled-fp16 correction rows), so the launcher chooses per call:
//   forward GEMMs  -> corr = 2: TF32 hi*hi + ONE fp16 MMA over a doubled K with exponent-balanced operand scales (operand
//                     rounding 7.6e-8 rms = 3xTF32's = fp32 level, at two tensor-pipe issues per 32 bytes of K instead of
//                     three).  What reaches the train-mode BatchNorms must be fp32-equivalent: S/17 at batch 4, embeddings
//                     vs fp64 over 8 runs were 7.0e-4 .. 7.7e-4 with 3xTF32, 7.3e-4 .. 1.18e-3 with the bf16 correction
//                     (round-1 tree), 5.8e-4 for the fp32 reference itself.  TN_TC_FWD_CORR=0 selects 3xTF32, =1 bf16.
//   gradient GEMMs -> corr = 1: TF32 + one bf16 correction MMA (TN_GEMM_GRAD flag / the fused depthwise backward): 6.6e-7
//                     rms; bf16 keeps fp32's exponent range, which gradients need.  TN_TC_BWD_CORR=0 / 2 override (A/B runs).
static int tc_corr_for(bool grad) {
  static int fwd = -1, bwd = -1;
  if (fwd < 0) {
    const char* e = getenv("TN_TC_FWD_CORR");
    fwd = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
    e = getenv("TN_TC_BWD_CORR");
    bwd = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return grad ? bwd : fwd;
}

// ---------------------------------------------------------------------------
// weight split: ws[0] = rna_tf32(W), ws[1] = rna_tf32(W - ws[0]), ws[2] / ws[3] = the packed bf16 / scaled fp16 correction rows;
// optional transpose
// ---------------------------------------------------------------------------
// ws[2] holds, per row and 32-element K chunk, 64 bf16 = [bf16(hi) x32 | bf16(x - hi) x32] in the 128 bytes that hold 32 tf32
// values in the other two planes (the weight side of the bf16 correction MMA, see tc_store_corr)
// planes: bit p set = write ws[p] (the batched per-step split writes only the planes the step's GEMMs read)
__device__ __forceinline__ void split_store(float* hi, float* lo, float* cr, float* cf, size_t i, int k, float x, int planes = 15) {
  const float h = __uint_as_float(rna_tf32(x));
  hi[i] = h;
  if (planes & 2) lo[i] = __uint_as_float(rna_tf32(x - h));
  if (planes & 4) {
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(cr + (i - (size_t)(k & 31))) + (k & 31);
    c[0] = __float2bfloat16_rn(h);
    c[32] = __float2bfloat16_rn(x - h);
  }
  if (planes & 8) {
    // ws[3]: the fp16 flavour with the weight side of the exponent-balanced scales (see TC_F16_SA / TC_F16_SC)
    unsigned short* f = reinterpret_cast<unsigned short*>(cf + (i - (size_t)(k & 31))) + (k & 31);
    f[0] = (unsigned short)(pack_f16x2(h * (1.f / TC_F16_SA), 0.f) & 0xffffu);
    f[32] = (unsigned short)(pack_f16x2((x - h) * TC_F16_SC, 0.f) & 0xffffu);
  }
}
__global__ void split_tf32_kernel(const float* __restrict__ W, float* __restrict__ ws, int M, int Kd, int transpose) {
  tn_grid_dep_sync();
  // output [4, M, Kd]; input [M, Kd] or (transpose) [Kd, M]
  const size_t n = (size_t)M * Kd;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / Kd), k = (int)(i - (size_t)m * Kd);
    const float x = transpose ? W[(size_t)k * M + m] : W[i];
    split_store(ws, ws + n, ws + 2 * n, ws + 3 * n, i, k, x);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static long long* g_trace = nullptr;
// debug: device buffer of 256 int64 receiving a clock64() timeline of two CTAs of the next tn_gemm_tc launches
extern "C" int tn_gemm_tc_set_trace(long long* buf) { g_trace = buf; return TN_OK; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int get_encoder() {
  if (g_encode) return TN_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  TN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  TN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
  g_encode = (PFN_encodeTiled)fn;
  return TN_OK;
}

// 2-D fp32 row-major [rows, cols] tensor, box = [box_rows, box_cols]: 32 columns -> 128-byte swizzle, 16 -> 64-byte swizzle
static int make_map(CUtensorMap* map, const float* base, long long rows, long long cols, int box_rows, int box_cols = TC_BK) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TN_REQUIRE(box_cols == 32 || box_cols == 16, "make_map: box of %d columns", box_cols);
  TN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%lld, %lld] box %d", (int)r, rows, cols, box_rows);
  return TN_OK;
}

// ---------------------------------------------------------------------------
// weight gradient on the tensor cores:  dW[co, ci] += sum_r dZ[r, co] * U[r, ci]
// The reduction runs over the rows, so both operands are "MN-major" in shared memory
// (the contiguous global dimension is M resp. N): a 3-D TMA box {32 channels, WG_BK rows,
// channel blocks} lands as [channel block][row][32 floats] = 4 KiB slabs of 128-byte
// swizzled rows, which is the canonical MN-major SWIZZLE_128B layout (LBO = slab stride,
// one 8-row group per UMMA_K = 8).  Split-K over row ranges: each CTA reduces its rows into
// a [128*MT, NB] fp32 TMEM tile and adds it to dW with vectorised red.global.add.v4.f32.
// ---------------------------------------------------------------------------
#define WG_BK 32                         // rows per chunk
#define WG_THREADS 320                   // TMA + MMA + 8 epilogue warps (two per TMEM lane quadrant)

// MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B shared-memory layout (32-byte
// swizzle chunks, 4-row period: byte-address bits [5,7) ^= bits [7,9)); the matching TMA mode is
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = stride between 32-channel blocks, SBO = stride
// between 4-row groups along K (512 B for densely packed 128-byte rows).
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct WgParams {
  float* dW;
  int R, Co, Ci, NB, stages, rows_per_split, tmem_cols;
  long long* trace; // tn_gemm_tc_set_trace: globaltimer stamps of CTA 0 (slots 0..15) and of the last CTA (16..31)
  int dbg;          // TN_WG_DEBUG (timing experiments): 1 = no reductions, 2 = no epilogue at all, 4 = red.global.add.v4 drain instead of bulk reductions
};

#define WG_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (blockIdx.z == 0 || blockIdx.z == gridDim.z - 1)) { \
    unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); \
    p.trace[(blockIdx.z == 0 ? 0 : 16) + (slot)] = (long long)gt_; } } while (0)
template <int MT>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, WgParams p) {
  tn_grid_dep_sync();
  if (threadIdx.x == 0) WG_TRACE(0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages, NB = p.NB;
  const int co0 = blockIdx.x * (128 * MT);
  const int ci0 = blockIdx.y * NB;
  const int ra = blockIdx.z * p.rows_per_split;
  const int rb = min(p.R, ra + p.rows_per_split);
  const int num_kc = (rb - ra + WG_BK - 1) / WG_BK;
  const uint32_t slab = WG_BK * 128;                               // one 32-channel block: WG_BK rows x 128 B
  const uint32_t a_bytes = MT * 4 * slab, b_bytes = (uint32_t)(NB / 32) * slab;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_MAX_STAGES]), accum_bar = smem_u32(&bars[2 * TC_MAX_STAGES]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) WG_TRACE(1);

  if (num_kc > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int kc = 0; kc < num_kc; ++kc) {
          const int s = kc % S;
          if (kc >= S) mbar_wait(empty0 + 8 * s, ((kc / S) - 1) & 1);
          const uint32_t fb = full0 + 8 * s;
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          tma_load_3d(base, &tmA, fb, 0, ra + kc * WG_BK, co0 / 32);
          tma_load_3d(base + a_bytes, &tmB, fb, 0, ra + kc * WG_BK, ci0 / 32);
        }
        WG_TRACE(2);
      }
    } else if (warp == 1) {
      {
        // the whole warp runs the loop, one elected lane issues (see gemm_tc2_kernel)
        // tf32, fp32 accumulate, A and B MN-major (bits 15, 16), M = 128, N = NB
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
        for (int kc = 0; kc < num_kc; ++kc) {
          const int s = kc % S;
          mbar_wait(full0 + 8 * s, (kc / S) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          if (tc_elect_one()) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
              for (int k8 = 0; k8 < WG_BK / 8; ++k8) {
                const uint64_t da = umma_desc_mn128(base + mt * 4 * slab + k8 * 1024, slab);
                const uint64_t db = umma_desc_mn128(base + a_bytes + k8 * 1024, slab);
                tc_mma_tf32(tmem_base + (uint32_t)(mt * 256), da, db, idesc, (kc > 0 || k8 > 0) ? 1u : 0u);
              }
            }
            tc_commit(empty0 + 8 * s);
            if (kc == 0) WG_TRACE(3);
          }
          __syncwarp();
        }
        if (tc_elect_one()) { tc_commit(accum_bar); WG_TRACE(4); }
        __syncwarp();
      }
    } else if (!(p.dbg & 2)) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      if (threadIdx.x == 64) WG_TRACE(5);
      const int quad = warp & 3, half = (warp - 2) >> 2;
      // Epilogue through shared memory (the pipeline stages are free now): TMEM hands each thread one output channel
      // (row of dW), so a direct red.global.add.v4 would touch 32 different 128-byte lines per warp instruction; staged,
      // each warp instruction adds 512 contiguous bytes of one row.  Rows are padded to NB + 4 floats (conflict-free).
      // EIGHT epilogue warps: the two warps of a TMEM lane quadrant take alternate 64-column groups (four tcgen05.ld.x16 in
      // flight per wait) and alternate rows of the drain.  With four warps and one load per wait the epilogue was 6.0 of the
      // kernel's 12.7 us (4.5 us of it TMEM -> shared memory -> registers, 1.5 us the reductions themselves).
      float* stg = reinterpret_cast<float*>(smem) + (size_t)(quad * 32) * (NB + 4);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        for (int c0 = 64 * half; c0 < NB; c0 += 128) {
          float v[64];
          const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * 256 + c0);
          const int ng = min(4, (NB - c0) / 16);        // NB is a multiple of 16
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (g < ng) tc_ld16_issue(ta + 16 * g, v + 16 * g);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 64; ++j) asm volatile("" : "+f"(v[j]));
          float4* dst = reinterpret_cast<float4*>(stg + (size_t)lane * (NB + 4) + c0);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < 4 * ng) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        fence_proxy_async();                                                // the bulk engine (async proxy) reads what was just stored
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");       // the quadrant's two warps: staging complete
        if (threadIdx.x == 64) WG_TRACE(6);
        if (p.dbg & 4) {
          // TN_WG_DEBUG=4: the earlier drain, one red.global.add.v4 per thread and 16 bytes (A/B)
          for (int r = half; r < 32; r += 2) {
            float* grow = p.dW + (size_t)(co0 + mt * 128 + quad * 32 + r) * p.Ci + ci0;
            for (int c = 4 * lane; c < NB; c += 128) {
              const float4 x = *reinterpret_cast<const float4*>(stg + (size_t)r * (NB + 4) + c);
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(grow + c), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
            }
          }
        } else if (!(p.dbg & 1)) {
          // drain: ONE bulk reduction per staged row (NB floats, contiguous in dW): cp.reduce.async.bulk adds the row in L2
          // without a shared-memory read, an address and an instruction per 16 bytes on the SM (8192 red.v4 per CTA before)
          const int r = half + 2 * lane;
          if (r < 32) {
            float* grow = p.dW + (size_t)(co0 + mt * 128 + quad * 32 + r) * p.Ci + ci0;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                         ::"l"(grow), "r"(smem_u32(stg + (size_t)r * (NB + 4))), "r"((uint32_t)NB * 4u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");       // drained before the next tile is staged
        if (threadIdx.x == 64) WG_TRACE(7);
      }
    } else if (warp == 2) {
      mbar_wait(accum_bar, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) WG_TRACE(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// Pair (cta_group::2) variant of the weight-gradient kernel: a 2-CTA cluster reduces a row range into a [256, NB2] tile.
// CTA r loads dZ[rows, 128 r .. 128 r + 128) and U[rows, r NB2/2 .. (r+1) NB2/2): 32 KB per 32-row chunk for a 256 x 256
// tile, where the single-CTA kernel spends 32 KB per 128 x 128 tile -- half the TMA bytes per SM, which is what paces it.
// Each CTA ends with its 128 output channels x NB2 columns in TMEM; eight epilogue warps add them to dW (red.global.add.v4).
__global__ void __launch_bounds__(TC_GEMM_THREADS, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, WgParams p) {
  tn_grid_dep_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages, NB2 = p.NB, HN = NB2 >> 1;
  const uint32_t rank = cluster_cta_rank();
  const bool leader = rank == 0;
  const int co0 = blockIdx.z * 256 + (int)rank * 128;     // this CTA's output channels
  const int ci0 = blockIdx.y * NB2;                        // the pair's input-channel block
  const int split = blockIdx.x >> 1;                       // gridDim.x = 2 * splits: the cluster spans x
  const int ra = split * p.rows_per_split;
  const int rb = min(p.R, ra + p.rows_per_split);
  const int num_kc = (rb - ra + WG_BK - 1) / WG_BK;
  const uint32_t slab = WG_BK * 128;
  const uint32_t a_bytes = 4 * slab, b_bytes = (uint32_t)(HN / 32) * slab;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_MAX_STAGES]), accum_bar = smem_u32(&bars[2 * TC_MAX_STAGES]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (num_kc > 0) {
    if (warp == 0) {
      if (lane == 0) {
        uint32_t full_leader0;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(full_leader0) : "r"(full0), "r"(0u));
        for (int kc = 0; kc < num_kc; ++kc) {
          const int s = kc % S;
          if (kc >= S) mbar_wait(empty0 + 8 * s, ((kc / S) - 1) & 1);
          if (leader) mbar_expect_tx(full0 + 8 * s, 2 * stage_bytes);      // both CTAs' tiles
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(base), "l"(&tmA), "r"(full_leader0 + 8 * s), "r"(0), "r"(ra + kc * WG_BK), "r"(co0 / 32) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(base + a_bytes), "l"(&tmB), "r"(full_leader0 + 8 * s), "r"(0), "r"(ra + kc * WG_BK), "r"((ci0 + (int)rank * HN) / 32) : "memory");
        }
      }
    } else if (warp == 1) {
      if (leader) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB2 >> 3) << 17) | ((256u >> 4) << 24);
        for (int kc = 0; kc < num_kc; ++kc) {
          const int s = kc % S;
          mbar_wait(full0 + 8 * s, (kc / S) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          if (tc_elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < WG_BK / 8; ++k8) {
              const uint64_t da = umma_desc_mn128(base + k8 * 1024, slab);
              const uint64_t db = umma_desc_mn128(base + a_bytes + k8 * 1024, slab);
              tc_mma_tf32_2sm(tmem_base, da, db, idesc, (kc > 0 || k8 > 0) ? 1u : 0u);
            }
            tc_commit_2sm(empty0 + 8 * s, (uint16_t)3);
          }
          __syncwarp();
        }
        if (tc_elect_one()) tc_commit_2sm(accum_bar, (uint16_t)3);
        __syncwarp();
      }
    } else {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const int quad = warp & 3, half = (warp - 2) >> 2;
      // staged through the (free) pipeline memory so that every red.global.add.v4 instruction covers 512 contiguous bytes
      float* stg = reinterpret_cast<float*>(smem) + (size_t)(quad * 32) * (NB2 + 4);
      for (int c0 = 16 * half; c0 < NB2; c0 += 32) {       // the two warps of a lane quadrant alternate 16-column groups
        float v[16];
        tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
        float4* dst = reinterpret_cast<float4*>(stg + (size_t)lane * (NB2 + 4) + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int r = half; r < 32; r += 2) {
        float* grow = p.dW + (size_t)(co0 + quad * 32 + r) * p.Ci + ci0;
        for (int c = 4 * lane; c < NB2; c += 128) {
          const float4 x = *reinterpret_cast<const float4*>(stg + (size_t)r * (NB2 + 4) + c);
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(grow + c), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// 3-D view of a row-major fp32 [rows, cols] tensor as {32, rows, cols/32}; box {32, WG_BK, nblk}
static int make_map_mn(CUtensorMap* map, const float* base, long long rows, long long cols, int nblk) {
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)cols * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, WG_BK, (cuuint32_t)nblk};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed (%d) for [%lld, %lld] nblk %d", (int)r, rows, cols, nblk);
  return TN_OK;
}

extern "C" int tn_wgrad_tc_supported(int R, int Ci, int Co) {
  return (R >= 32 && Ci >= 32 && Ci % 32 == 0 && Co >= 128 && Co % 128 == 0) ? 1 : 0;
}

// dW[Co, Ci] += dZ[R, Co]^T U[R, Ci]   (plain TF32 operands, fp32 accumulation; ACCUMULATED)
extern "C" int tn_wgrad_tc(const float* dZ, const float* U, float* dW, int R, int Ci, int Co, void* stream) {
  TN_REQUIRE(dZ && U && dW, "wgrad_tc: null tensor");
  TN_REQUIRE(tn_wgrad_tc_supported(R, Ci, Co), "wgrad_tc: unsupported shape R=%d Ci=%d Co=%d", R, Ci, Co);
  TN_REQUIRE(tn_aligned16(dZ) && tn_aligned16(U) && tn_aligned16(dW), "wgrad_tc: operands must be 16B aligned");
  int rc = get_encoder();
  if (rc != TN_OK) return rc;
  static int wg_pair = -1;
  if (wg_pair < 0) { const char* e = getenv("TN_WG_PAIR"); wg_pair = (e && atoi(e) != 0) ? 1 : 0; }   // measured slower (16.1 vs 14.5 us): opt-in
  if (wg_pair && Co % 256 == 0 && Ci % 64 == 0 && R >= 1024) {
    int NB2 = 256;
    if (const char* e = getenv("TN_WG_NB2")) NB2 = atoi(e) >= 64 ? atoi(e) : NB2;     // tuning knob
    while (Ci % NB2 != 0) NB2 -= 64;                 // largest multiple of 64 <= 256 that divides Ci
    const int co_groups = Co / 256, ci_blocks = Ci / NB2;
    const size_t stage_bytes2 = (size_t)WG_BK * 128 * (4 + NB2 / 64);
    int stages2 = (int)((TC_SMEM_LIMIT - 2048) / stage_bytes2);
    if (stages2 > TC_MAX_STAGES) stages2 = TC_MAX_STAGES;
    const long long chunks = ((long long)R + WG_BK - 1) / WG_BK;
    long long want = (tn_num_sms() / 2) / (co_groups * ci_blocks);
    if (want < 1) want = 1;
    if (want > chunks) want = chunks;
    const long long cps = (chunks + want - 1) / want;
    const int splits = (int)((chunks + cps - 1) / cps);
    if (stages2 >= 2 && co_groups <= 65535) {
      CUtensorMap mA, mB;
      if ((rc = make_map_mn(&mA, dZ, R, Co, 4)) != TN_OK) return rc;
      if ((rc = make_map_mn(&mB, U, R, Ci, NB2 / 64)) != TN_OK) return rc;
      WgParams p;
      p.dW = dW; p.R = R; p.Co = Co; p.Ci = Ci; p.NB = NB2; p.stages = stages2; p.rows_per_split = (int)cps * WG_BK; p.dbg = 0; p.trace = nullptr;
      int cols = 32;
      while (cols < NB2) cols <<= 1;
      p.tmem_cols = cols;
      const size_t smem = stage_bytes2 * stages2 + 1024;
      // CTAs (2 s, y, z) and (2 s + 1, y, z) form the pair of row split s
      dim3 grid(2 * splits, ci_blocks, co_groups);
      TN_CUDA(cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tn_launch_cluster(wgrad_tc2_kernel, grid, TC_GEMM_THREADS, smem, stream, 2, mA, mB, p);
      TN_LAUNCH_CHECK("wgrad_tc2_kernel");
      return TN_OK;
    }
  }
  // 128 x 128 output tiles: the epilogue (TMEM -> red.global) is a fixed cost per CTA, so small tiles with long
  // row ranges win over 256 x 256 tiles with short ones (measured 14.6 us vs 25.8 us at R=19264, 256 x 256)
  // (with the epilogue staged through shared memory, 128 x 256 tiles measure best: 12.9 us against 15.3 us for 128 x 128
  //  and 19.6 us for 256 x 256 at R=19264, 256 x 256; the pair kernel above reaches 12.8 us)
  int MT = 1;
  int NB = 256;
  if (const char* e = getenv("TN_WG_MT")) MT = (atoi(e) == 1 || Co % 256 != 0) ? 1 : 2;      // tuning knobs
  if (const char* e = getenv("TN_WG_NB")) NB = atoi(e) >= 32 ? atoi(e) : NB;
  while (Ci % NB != 0) NB -= 32;                     // largest multiple of 32 <= 256 that divides Ci
  TN_REQUIRE(NB % 16 == 0 && NB >= 32, "wgrad_tc: no N tile for Ci=%d", Ci);
  const int co_groups = Co / (128 * MT), ci_blocks = Ci / NB;
  const size_t stage_bytes = (size_t)WG_BK * 128 * (MT * 4 + NB / 32);
  int stages = (int)((TC_SMEM_LIMIT - 2048) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  TN_REQUIRE(stages >= 2, "wgrad_tc: tile does not fit shared memory");
  TN_REQUIRE(stage_bytes * stages >= (size_t)128 * (NB + 4) * 4, "wgrad_tc: epilogue staging does not fit the pipeline memory");
  long long chunks = ((long long)R + WG_BK - 1) / WG_BK;
  long long want = tn_num_sms() / (co_groups * ci_blocks);
  if (want < 1) want = 1;
  if (want > chunks) want = chunks;
  long long cps = (chunks + want - 1) / want;        // chunks per split
  int splits = (int)((chunks + cps - 1) / cps);
  CUtensorMap mA, mB;
  if ((rc = make_map_mn(&mA, dZ, R, Co, MT * 4)) != TN_OK) return rc;
  if ((rc = make_map_mn(&mB, U, R, Ci, NB / 32)) != TN_OK) return rc;
  WgParams p;
  p.dW = dW; p.R = R; p.Co = Co; p.Ci = Ci; p.NB = NB; p.stages = stages; p.rows_per_split = (int)cps * WG_BK;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("TN_WG_DEBUG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  p.trace = g_trace;
  int cols = MT == 2 ? 512 : 32;
  while (MT == 1 && cols < NB) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = stage_bytes * stages + 1024;
  dim3 grid(co_groups, ci_blocks, splits);
  if (MT == 2) {
    TN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tn_launch(wgrad_tc_kernel<2>, grid, WG_THREADS, smem, stream, mA, mB, p);
  } else {
    TN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tn_launch(wgrad_tc_kernel<1>, grid, WG_THREADS, smem, stream, mA, mB, p);
  }
  TN_LAUNCH_CHECK("wgrad_tc_kernel");
  return TN_OK;
}

extern "C" int tn_gemm_tc_supported(int R, int Kd, int M) {
  return (R > 0 && Kd >= TC_BK && Kd % TC_BK == 0 && M >= 128 && M % 128 == 0) ? 1 : 0;
}

extern "C" int tn_split_tf32(const float* W, float* ws, int M, int Kd, int transpose, void* stream) {
  TN_REQUIRE(W && ws && M > 0 && Kd > 0 && Kd % 32 == 0, "split_tf32: bad arguments (Kd %% 32 == 0)");
  size_t n = (size_t)M * Kd;
  int blocks = (int)((n + 255) / 256);
  if (blocks > tn_num_sms() * 8) blocks = tn_num_sms() * 8;
  tn_launch(split_tf32_kernel, blocks, 256, 0, stream, W, ws, M, Kd, transpose);
  TN_LAUNCH_CHECK("split_tf32_kernel");
  return TN_OK;
}

// planes_fwd / planes_bwd: the planes the forward GEMMs (jobs with transpose = 0) and the gradient GEMMs (transpose = 1) read.
// 32 x 32 tiles, 256 threads: a transposed job reads W[k, m] along m and writes ws[m, k] along k through shared memory (the
// first version read the transposed weights with a stride of M floats per thread: 50 us per step for 25 MB of weights).
__global__ void __launch_bounds__(256) split_tf32_batch_kernel(const tn_split_job* __restrict__ jobs, int njobs, int total_tiles,
                                                               int tiles_per_block, int planes_fwd, int planes_bwd) {
  tn_grid_dep_sync();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // this block's contiguous range of the step's tiles (job.tile0 = index of the job's first tile): one search, then a walk
  const int t_begin = blockIdx.x * tiles_per_block, t_end = min(total_tiles, t_begin + tiles_per_block);
  if (t_begin >= t_end) return;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile0 <= t_begin) lo = mid; else hi = mid - 1;
  }
  int ji = lo;
  tn_split_job j = jobs[ji];
  int tiles_k = j.Kd / 32, ntiles = tiles_k * ((j.M + 31) / 32);       // Kd % 32 == 0 (tn_split_tf32's contract)
  for (int tg = t_begin; tg < t_end; ++tg) {
    while (tg - j.tile0 >= ntiles) {
      j = jobs[++ji];
      tiles_k = j.Kd / 32; ntiles = tiles_k * ((j.M + 31) / 32);
    }
    const int t = tg - j.tile0;
    const size_t n = (size_t)j.M * j.Kd;
    const int planes = j.transpose ? planes_bwd : planes_fwd;
    const int m0 = (t / tiles_k) * 32, k0 = (t % tiles_k) * 32;
    if (j.transpose) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty + 8 * i, m = m0 + tx;
        tile[ty + 8 * i][tx] = m < j.M ? j.W[(size_t)k * j.M + m] : 0.f;
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + 8 * i, k = k0 + tx;
      if (m < j.M) {
        const size_t idx = (size_t)m * j.Kd + k;
        const float x = j.transpose ? tile[tx][ty + 8 * i] : j.W[idx];
        split_store(j.ws, j.ws + n, j.ws + 2 * n, j.ws + 3 * n, idx, k, x, planes);
      }
    }
    if (j.transpose) __syncthreads();
  }
}
// Every weight split of a step in one launch: the 32 x 32 tiles of all jobs form one list (job.tile0 = the job's first tile,
// total_tiles = their number) that a grid of ~8 blocks per SM walks in contiguous ranges.  Only the planes this configuration's
// GEMMs read are written: tf32 hi + the forward scheme's correction plane for the [M, Kd] orientation, tf32 hi + the gradient
// scheme's for the transposed one.
extern "C" int tn_split_tf32_batch(const tn_split_job* jobs_dev, int njobs, int total_tiles, void* stream) {
  TN_REQUIRE(jobs_dev && njobs > 0 && total_tiles > 0, "split_tf32_batch: bad arguments");
  int blocks = tn_num_sms() * 8;
  if (blocks > total_tiles) blocks = total_tiles;
  const int tpb = (total_tiles + blocks - 1) / blocks;
  blocks = (total_tiles + tpb - 1) / tpb;
  const int plane_of[3] = {2, 4, 8};                 // corr 0 -> tf32 lo, 1 -> bf16 rows, 2 -> scaled fp16 rows
  tn_launch(split_tf32_batch_kernel, blocks, 256, 0, stream, jobs_dev, njobs, total_tiles, tpb, 1 | plane_of[tc_corr_for(false)],
            1 | plane_of[tc_corr_for(true)]);
  TN_LAUNCH_CHECK("split_tf32_batch_kernel");
  return TN_OK;
}

// rows of output per CTA: the multiple of 16 that minimises waves * (tile rows + fixed cost).
// `halo` extra MMA columns ride along (fused depthwise backward).
static int pick_bn(long long R, int groups, int per_row_stage_bytes, int fixed_stage_bytes, int halo, int reserve, int* stages_out) {
  const int sms = tn_num_sms();
  int best = 0;
  double best_cost = 1e30;
  for (int bn = 256 - halo; bn >= 32; bn -= 16) {
    long long stage = (long long)fixed_stage_bytes + (long long)(bn + halo) * per_row_stage_bytes;
    int stages = (int)((TC_SMEM_LIMIT - 2048 - reserve) / stage);
    if (stages < 2) continue;
    long long ctas = ((R + bn - 1) / bn) * groups;
    long long waves = (ctas + sms - 1) / sms;
    double cost = (double)waves * (bn + halo + 24);
    if (cost < best_cost) { best_cost = cost; best = bn; *stages_out = stages > TC_MAX_STAGES ? TC_MAX_STAGES : stages; }
  }
  return best;
}

template <int MT, int MODE>
static cudaError_t launch_inst(bool cluster, dim3 grid, size_t smem, void* stream, const CUtensorMap& a, const CUtensorMap& b,
                               const CUtensorMap& c, const CUtensorMap& d, const TcParams& p) {
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<MT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (cluster) tn_launch_cluster(gemm_tc_kernel<MT, MODE>, grid, TC_GEMM_THREADS, smem, stream, 2, a, b, c, d, p);
  else tn_launch(gemm_tc_kernel<MT, MODE>, grid, TC_GEMM_THREADS, smem, stream, a, b, c, d, p);
  return cudaSuccess;
}
template <int MT>
static cudaError_t launch_mode(int mode, bool cluster, dim3 grid, size_t smem, void* stream, const CUtensorMap& a, const CUtensorMap& b,
                               const CUtensorMap& c, const CUtensorMap& d, const TcParams& p) {
  switch (mode) {
    case 1: return launch_inst<MT, 1>(cluster, grid, smem, stream, a, b, c, d, p);
    case 2: return launch_inst<MT, 2>(cluster, grid, smem, stream, a, b, c, d, p);
    default: return launch_inst<MT, 0>(cluster, grid, smem, stream, a, b, c, d, p);
  }
}

// common launcher: p carries the epilogue configuration; shapes / tiles / maps are filled here
static int nacc_cap() {
  static int cap = -1;
  if (cap < 0) { const char* e = getenv("TN_TC_NACC"); cap = e ? atoi(e) : 8; if (cap < 1) cap = 1; if (cap > 8) cap = 8; }
  return cap;
}
// accumulators per tile: `cols_avail` TMEM columns per tile, `bn` columns per accumulator, at most one main chunk per K chunk
static int pick_nacc(int cols_avail, int bn, int num_kc, int nsplit) {
  if (nsplit != 3) return 1;
  int n = cols_avail / bn;
  if (n > nacc_cap()) n = nacc_cap();
  if (n > num_kc + 1) n = num_kc + 1;
  return n < 2 ? 1 : n;
}

static int launch_gemm_tc(const float* X, const float* ws, TcParams p, int R, int Kd, int M, int nsplit, const tn_scratch* scratch,
                          void* stream) {
  TN_REQUIRE(X && ws, "gemm_tc: null tensor");
  const bool grad = (p.flags & TN_GEMM_GRAD) != 0 || p.dw_K > 0;
  p.flags &= ~TN_GEMM_GRAD;
  p.corr = tc_corr_for(grad);
  const float* ws_lo = ws + (size_t)(p.corr == 2 ? 3 : p.corr ? 2 : 1) * M * Kd;
  if (p.stats) {
    TN_REQUIRE(scratch && scratch->accum && scratch->tickets && scratch->accum_words >= TN_ACCUM_WORDS(M),
               "gemm_tc: statistics need a tn_scratch with TN_ACCUM_WORDS(M) zeroed accumulator words and the ticket array");
    TN_REQUIRE(M / 128 <= TN_TICKETS, "gemm_tc: statistics of more than %d channels are not supported", TN_TICKETS * 128);
    p.accum = scratch->accum; p.tickets = scratch->tickets;
    p.fc = tn_fold_const(p.has_bn ? p.bn.n : 1.0);
  }
  TN_REQUIRE(tn_gemm_tc_supported(R, Kd, M), "gemm_tc: unsupported shape R=%d K=%d M=%d (need K %% 32 == 0, M %% 128 == 0)", R, Kd, M);
  TN_REQUIRE(nsplit == 1 || nsplit == 3, "gemm_tc: nsplit must be 1 or 3");
  TN_REQUIRE(tn_aligned16(X) && tn_aligned16(ws), "gemm_tc: operands must be 16B aligned");
  int rc = get_encoder();
  if (rc != TN_OK) return rc;
  static int use_pair = -1;
  if (use_pair < 0) { const char* e = getenv("TN_TC_PAIR"); use_pair = (e && atoi(e) == 0) ? 0 : 1; }   // measured: 9.99 -> 9.48 ms per step
  static int use_pair_dw = -1;
  if (use_pair_dw < 0) { const char* e = getenv("TN_TC_PAIR_DW"); use_pair_dw = (e && atoi(e) == 0) ? 0 : 1; }
  if (use_pair && (p.dw_K == 0 || use_pair_dw) && (p.fdw_K == 0 || p.fdw_K <= 3) && nsplit == 3 && M % 256 == 0 && !(p.flags & 3) && R >= 512) {
    // cta_group::2 pair kernel: rows per pair = 2 * BN2 (minus the halo of the fused depthwise backward), chosen like pick_bn
    // (one wave of pairs, smallest tile that achieves it)
    const int sms = tn_num_sms();
    const int halo2 = p.dw_K > 1 ? 16 : 0;
    static int ew_sel = -1;                                    // epilogue warps of the fused backward for K <= 3: 8, 12 or 16
    if (ew_sel < 0) {                                          // measured at R=19264, 256x256, K=3: 34.4 / 32.1 / 31.5 us
      const char* e = getenv("TN_TC_EW");
      const int v = e ? atoi(e) : 16;
      ew_sel = (v == 8 || v == 12) ? v : 16;
    }
    static int ew7 = -1;                                       // epilogue warps of the fused backward for 3 < K <= 7: 8 or 12
    if (ew7 < 0) { const char* e = getenv("TN_TC_EW7"); ew7 = (e && atoi(e) == 8) ? 8 : 12; }
    const int ew = (p.dw_K > 0 && p.dw_K <= 3) ? ew_sel : ((p.dw_K > 3 && p.dw_K <= 7) ? ew7 : 8);
    const bool wide = ew > 8;
    const int par2 = p.fdw_K > 0 ? (Kd * (3 + p.fdw_K) * 4 + 1023) / 1024 * 1024 : (p.has_bnb ? (Kd * 2 * 4 + 1023) / 1024 * 1024 : 0);
    const int raw2 = p.fdw_K > 0 ? 2 * 8 * TC_BK * 4 : 0;      // halo rows of the two raw tiles per stage
    const int red2 = (p.dw_K > 0 ? ((ew / 4) * 128 * (p.dw_K + 3) * 4 + 1023) / 1024 * 1024 : (p.fdw_K > 0 ? 4096 : 2048)) + par2;
    int best = 0; double best_cost = 1e30;
    for (int bn2 = 256; bn2 >= 32; bn2 -= 16) {
      const long long stage = 2ll * 128 * TC_BK * 4 + 4ll * (bn2 / 2) * TC_BK * 4 + raw2;
      if (stage * TC2_STAGES + red2 + 2048 > TC_SMEM_LIMIT) continue;
      if (p.fdw_K > 0 && bn2 > 256 - 0) continue;
      if (p.dw_K > 0) {                                  // the 8 z boxes must fit the released stages
        const long long cap = stage / ((long long)bn2 * 128);
        if (cap < 1 || (8 + cap - 1) / cap > TC2_STAGES) continue;
      }
      const long long ctas = 2 * (((long long)R + (2 * bn2 - halo2) - 1) / (2 * bn2 - halo2)) * (M / 256);
      const long long waves = (ctas + sms - 1) / sms;
      const double cost = (double)waves * (2 * bn2 + 48);
      if (cost < best_cost) { best_cost = cost; best = bn2; }
    }
    if (best > 0) {
      CUtensorMap mA_hi, mA_lo, mB;
      if ((rc = make_map(&mA_hi, ws, M, Kd, 128)) != TN_OK) return rc;
      if ((rc = make_map(&mA_lo, ws_lo, M, Kd, 128)) != TN_OK) return rc;
      if ((rc = make_map(&mB, X, R, Kd, best / 2 + (p.fdw_K > 0 ? 8 : 0))) != TN_OK) return rc;
      CUtensorMap mZ = mB, mG = mB;
      if (p.dw_K > 0 && (rc = make_map(&mZ, p.zprev, R, M, best, 32)) != TN_OK) return rc;
      if (p.has_bnb && (rc = make_map(&mG, p.bnb.z, R, Kd, best / 2)) != TN_OK) return rc;
      p.R = R; p.Kd = Kd; p.M_total = M; p.BN = best; p.BNo = 2 * best - halo2; p.nsplit = 3;
      p.nacc = p.dw_K > 0 ? 1 : pick_nacc(256, best, Kd / TC_BK, 3);
      p.trace = g_trace;
      const size_t stage_bytes = 2ull * 128 * TC_BK * 4 + 4ull * (best / 2) * TC_BK * 4 + raw2;
      p.red_off = (uint32_t)(stage_bytes * TC2_STAGES);
      p.par_off = (uint32_t)(stage_bytes * TC2_STAGES + red2 - par2);
      const size_t smem = stage_bytes * TC2_STAGES + red2 + 1024;
      dim3 grid(2 * (unsigned)tn_cdiv(R, p.BNo), M / 256);
      if (p.fdw_K > 0) {
        TN_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tn_launch_cluster(gemm_tc2_kernel<2, 16>, grid, 64 + 32 * 16, smem, stream, 2, mA_hi, mA_lo, mB, mZ, mG, p);
      } else if (wide && ew == 16) {
        TN_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tn_launch_cluster(gemm_tc2_kernel<1, 16>, grid, 64 + 32 * 16, smem, stream, 2, mA_hi, mA_lo, mB, mZ, mG, p);
      } else if (wide) {
        TN_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<1, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tn_launch_cluster(gemm_tc2_kernel<1, 12>, grid, 64 + 32 * 12, smem, stream, 2, mA_hi, mA_lo, mB, mZ, mG, p);
      } else if (p.dw_K > 0) {
        TN_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tn_launch_cluster(gemm_tc2_kernel<1, 8>, grid, TC_GEMM_THREADS, smem, stream, 2, mA_hi, mA_lo, mB, mZ, mG, p);
      } else {
        TN_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tn_launch_cluster(gemm_tc2_kernel<0, 8>, grid, TC_GEMM_THREADS, smem, stream, 2, mA_hi, mA_lo, mB, mZ, mG, p);
      }
      TN_LAUNCH_CHECK("gemm_tc2_kernel");
      return TN_OK;
    }
  }
  TN_UNSUPPORTED(p.has_bnb, "gemm_tc: the BatchNorm-backward operand producer needs the pair kernel (R >= 512, M %% 256 == 0; R=%d M=%d)", R, M);
  const int MT = (M % 256 == 0) ? 2 : 1;
  const int groups = M / (128 * MT);
  const int mult = nsplit == 3 ? 2 : 1;
  const int halo = p.dw_K > 1 ? 16 : 0;              // 2 * PAD <= 14 rows of halo, rounded to the MMA's N granularity
  int stages = 2;
  // fused depthwise backward: [2 halves][128 channels][K + 3 sums] of staging behind the pipeline stages
  const int par_bytes = p.fdw_K > 0 ? (Kd * (3 + p.fdw_K) * 4 + 1023) / 1024 * 1024 : 0;
  const int red_bytes = (p.dw_K > 0 ? (2 * 128 * (p.dw_K + 3) * 4 + 1023) / 1024 * 1024 : (p.stats ? 2048 : 0)) + par_bytes;
  const int fdw_extra = p.fdw_K > 0 ? 16 * TC_BK * 4 : 0;       // raw-tile halo rows of the fused depthwise forward
  const int bno = pick_bn(R, groups, mult * TC_BK * 4, mult * MT * 128 * TC_BK * 4 + fdw_extra, halo, red_bytes, &stages);
  TN_REQUIRE(bno >= 32, "gemm_tc: no tile configuration fits shared memory");
  const int bn = bno + halo;
  CUtensorMap mA_hi, mA_lo, mB;
  if ((rc = make_map(&mA_hi, ws, M, Kd, 128)) != TN_OK) return rc;
  if ((rc = make_map(&mA_lo, ws_lo, M, Kd, 128)) != TN_OK) return rc;
  if ((rc = make_map(&mB, X, R, Kd, bn + (p.fdw_K > 0 ? 16 : 0))) != TN_OK) return rc;
  p.R = R; p.Kd = Kd; p.M_total = M; p.BN = bn; p.BNo = bno; p.stages = stages; p.nsplit = nsplit;
  p.nacc = p.dw_K > 0 ? 1 : pick_nacc(MT == 2 ? 256 : 512, bn, Kd / TC_BK, nsplit);
  p.trace = g_trace;
  int cols = MT == 2 ? 512 : 32;
  while (MT == 1 && cols < bn * p.nacc) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage_bytes = (size_t)mult * (MT * 128 * TC_BK * 4 + (size_t)bn * TC_BK * 4) + fdw_extra;
  const size_t smem = stage_bytes * stages + red_bytes + 1024;
  p.red_off = (uint32_t)(stage_bytes * stages);
  p.par_off = (uint32_t)(stage_bytes * stages + red_bytes - par_bytes);
  CUtensorMap mZ = mB;
  if (p.dw_K > 0) {
    // z tiles of the fused epilogue reuse the pipeline memory: one tile = 128 channels x bn rows = 4 blocks of bn x 128 B
    if ((rc = make_map(&mZ, p.zprev, R, M, bn, 32)) != TN_OK) return rc;
    const size_t ztile = (size_t)4 * bn * 128;
    const int num_kc = Kd / TC_BK;
    p.z_early = (stages % 2 == 0 && num_kc % stages == 0 && ztile <= stage_bytes * (stages / 2)) ? 1 : 0;
    p.z_off1 = (uint32_t)(((ztile > stage_bytes ? ztile : stage_bytes) + 1023) / 1024 * 1024);
    TN_REQUIRE(MT == 1 || p.z_early || p.z_off1 + ztile <= stage_bytes * stages, "gemm_tc_dwbwd: z tiles do not fit the pipeline memory");
    TN_REQUIRE(ztile <= stage_bytes * stages, "gemm_tc_dwbwd: z tile does not fit the pipeline memory");
  }
  dim3 grid(tn_cdiv(R, bno), groups);
  const int mode = p.dw_K > 0 ? 1 : (p.fdw_K > 0 ? 2 : 0);
  static int use_cluster = -1;
  if (use_cluster < 0) { const char* e = getenv("TN_TC_CLUSTER"); use_cluster = (e && atoi(e) != 0) ? 1 : 0; }   // no gain measured: opt-in
  if (MT == 2 && use_cluster && grid.x >= 2) {
    // pairs of row tiles share the weight tiles by TMA multicast; an odd grid gets one idle tile (rows >= R: zero-filled loads,
    // nothing stored)
    grid.x = (grid.x + 1) & ~1u;
    p.cluster2 = 1;
    TN_CUDA(launch_mode<2>(mode, true, grid, smem, stream, mA_hi, mA_lo, mB, mZ, p));
  } else if (MT == 2) {
    TN_CUDA(launch_mode<2>(mode, false, grid, smem, stream, mA_hi, mA_lo, mB, mZ, p));
  } else {
    TN_CUDA(launch_mode<1>(mode, false, grid, smem, stream, mA_hi, mA_lo, mB, mZ, p));
  }
  TN_LAUNCH_CHECK("gemm_tc_kernel");
  return TN_OK;
}

// ws: the split weights from tn_split_tf32, [3, M, Kd].  nsplit 3 = fp32-equivalent, 1 = plain TF32.
extern "C" int tn_gemm_tc(const float* X, const float* ws, const float* bias, float* Z, double* stats, int R, int Kd, int M,
                          int flags, int nsplit, const tn_scratch* scratch, void* stream) {
  TN_REQUIRE(Z, "gemm_tc: null output");
  TN_REQUIRE(!(flags & TN_EPI_ACCUM) || !stats, "gemm_tc: statistics of an accumulated output are not available");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.Z = Z; p.stats = stats; p.flags = flags;
  return launch_gemm_tc(X, ws, p, R, Kd, M, nsplit, scratch, stream);
}

static int check_bn(const tn_bn_fold* bn, const double* stats) {
  TN_REQUIRE(bn && stats, "bn fold needs the statistics buffer");
  TN_REQUIRE(bn->gamma && bn->beta && bn->scale && bn->shift && bn->mean && bn->invstd, "bn fold: null field");
  TN_REQUIRE(bn->n >= 1.0, "bn fold: n must be >= 1");
  TN_REQUIRE((bn->running_mean == nullptr) == (bn->running_var == nullptr), "bn fold: running_mean/var must come together");
  return TN_OK;
}

extern "C" int tn_gemm_tc_bn(const float* X, const float* ws, const float* bias, float* Z, double* stats, const tn_bn_fold* bn,
                             int R, int Kd, int M, int flags, int nsplit, const tn_scratch* scratch, void* stream) {
  TN_REQUIRE(Z, "gemm_tc_bn: null output");
  int rc = check_bn(bn, stats);
  if (rc != TN_OK) return rc;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.Z = Z; p.stats = stats; p.flags = flags;
  p.bn = *bn; p.has_bn = 1;
  return launch_gemm_tc(X, ws, p, R, Kd, M, nsplit, scratch, stream);
}

static int check_bnb(const tn_bn_bwd* b) {
  TN_REQUIRE(b && b->z && b->dscale && b->dshift && b->mean && b->invstd && b->gamma && b->g_out && b->dgamma && b->dbeta && b->n >= 1.0,
             "bn backward producer: null field");
  TN_REQUIRE(tn_aligned16(b->z) && tn_aligned16(b->g_out), "bn backward producer: z and g_out must be 16B aligned");
  return TN_OK;
}
extern "C" int tn_gemm_tc_bnbwd_supported(int R, int Kd, int M) {
  return (tn_gemm_tc_supported(R, Kd, M) && R >= 512 && M % 256 == 0) ? 1 : 0;
}
// Data gradient of a conv followed by a train-mode BatchNorm with the BatchNorm backward folded into the operand load:
//   g = dZ + a[c] + b[c] z  (tn_bn_stats_bwd's arithmetic, per 32-channel chunk in shared memory), dX = g W,
//   g_out = g (for the weight-gradient GEMM), dbias += column sums of g, dgamma / dbeta written.
extern "C" int tn_gemm_tc_bnbwd(const float* dZ, const float* ws, const tn_bn_bwd* bnb, float* dX, int R, int Kd, int M, int flags,
                                void* stream) {
  TN_REQUIRE(dX, "gemm_tc_bnbwd: null output");
  int rc = check_bnb(bnb);
  if (rc != TN_OK) return rc;
  TN_UNSUPPORTED(!tn_gemm_tc_bnbwd_supported(R, Kd, M), "gemm_tc_bnbwd: unsupported shape R=%d K=%d M=%d", R, Kd, M);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.Z = dX; p.flags = (flags & TN_EPI_ACCUM) | TN_GEMM_GRAD;
  TN_UNSUPPORTED(flags & (TN_EPI_TANH | TN_EPI_ACCUM), "gemm_tc_bnbwd: epilogue flags are not supported");
  p.bnb = *bnb; p.has_bnb = 1;
  return launch_gemm_tc(dZ, ws, p, R, Kd, M, 3, nullptr, stream);
}

// Forward of a depthwise-separable conv block + train-mode BatchNorm fold in ONE kernel:
//   u = depthwise_K(act(z)) + b_dw  (transform warps, from the raw z tile)   [also written to u_out for the backward wgrad]
//   Z = u W^T + b_pw  (tensor cores, 3xTF32), statistics of Z, BatchNorm fold by the last CTA (bn may be NULL: stats only / none)
extern "C" int tn_gemm_tc_dwfwd(const float* z, const float* ws, const float* dw_w, const float* dw_b, const float* scale,
                                const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer,
                                const float* pw_bias, float* u_out, float* Z, double* stats, const tn_bn_fold* bn, int B, int T,
                                int C, int Co, int K, const tn_scratch* scratch, void* stream) {
  TN_REQUIRE(z && dw_w && Z, "gemm_tc_dwfwd: null tensor");
  TN_REQUIRE(K >= 1 && K <= 7 && (K & 1), "gemm_tc_dwfwd: unsupported depthwise kernel size %d (odd sizes 1..7; wider windows do not fit the register file)", K);
  TN_REQUIRE((scale == nullptr) == (shift == nullptr), "gemm_tc_dwfwd: scale and shift come together");
  TN_REQUIRE(drop_p <= 0.f || seed, "gemm_tc_dwfwd: dropout needs a seed");
  long long R = (long long)B * T;
  TN_REQUIRE(B > 0 && T > 0 && R < (1ll << 31), "gemm_tc_dwfwd: bad B/T");
  TN_REQUIRE(tn_aligned16(dw_b ? (const void*)dw_b : (const void*)z) && (!u_out || tn_aligned16(u_out)) && (!scale || (tn_aligned16(scale) && tn_aligned16(shift))),
             "gemm_tc_dwfwd: pointers must be 16B aligned");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.bias = pw_bias; p.Z = Z; p.stats = stats;
  if (bn) {
    int rc = check_bn(bn, stats);
    if (rc != TN_OK) return rc;
    p.bn = *bn; p.has_bn = 1;
  }
  p.fdw_K = K; p.fdw_T = T; p.fdw_w = dw_w; p.fdw_b = dw_b; p.fdw_u = u_out;
  p.act = tn_make_act(scale, shift, relu, drop_p, seed, layer);
  return launch_gemm_tc(z, ws, p, (int)R, C, Co, 3, scratch, stream);
}

// Data gradient of a depthwise-separable conv block in one kernel:
//   du = dZ[R, Co] W[Co, C]            (tensor cores; ws = tn_split_tf32(W, transpose = 1), [2, C, Co])
//   dzprev = act'(zprev) * scale * depthwise_K^T(du),   dw += ..., dbias += ..., dscale += ..., dshift += ...
// i.e. tn_gemm_tc(dgrad) followed by tn_dw_bwd, without du ever leaving the SM.  scale == NULL: zprev is
// already the activation (first sub-block of a mega-block) and dzprev is the gradient w.r.t. it.
static int dwbwd_impl(const float* dZ, const float* ws, const float* zprev, float* dzprev, const float* dw_w,
                      float* g_dw, float* g_dbias, float* g_dscale, float* g_dshift, const float* scale,
                      const float* shift, int relu, float drop_p, const unsigned long long* seed,
                      unsigned int layer, int B, int T, int Co, int C, int K, int nsplit, const tn_bn_bwd* bnb, void* stream) {
  TN_REQUIRE(zprev && dzprev && dw_w && g_dw, "gemm_tc_dwbwd: null tensor");
  TN_REQUIRE(K >= 1 && K <= 11 && (K & 1), "gemm_tc_dwbwd: unsupported depthwise kernel size %d (odd sizes 1..11)", K);
  TN_REQUIRE(!scale || (shift && g_dscale && g_dshift), "gemm_tc_dwbwd: scale given without shift/dscale/dshift");
  TN_REQUIRE(drop_p <= 0.f || seed, "gemm_tc_dwbwd: dropout needs a seed");
  long long R = (long long)B * T;
  TN_REQUIRE(B > 0 && T > 0 && R < (1ll << 31), "gemm_tc_dwbwd: bad B/T");
  TN_REQUIRE((R + 16) * C < (1ll << 32), "gemm_tc_dwbwd: R * C must stay below 2^32 (32-bit element offsets)");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.dw_K = K; p.dw_T = T; p.dw_w = dw_w; p.zprev = zprev; p.dzprev = dzprev;
  p.g_dw = g_dw; p.g_db = g_dbias; p.g_dscale = g_dscale; p.g_dshift = g_dshift;
  p.act = tn_make_act(scale, shift, relu, drop_p, seed, layer);
  if (bnb) {
    int rc = check_bnb(bnb);
    if (rc != TN_OK) return rc;
    TN_UNSUPPORTED(!tn_gemm_tc_bnbwd_supported((int)R, Co, C) || nsplit != 3, "gemm_tc_dwbwd_bn: unsupported shape R=%lld K=%d M=%d", R, Co, C);
    p.bnb = *bnb; p.has_bnb = 1;
  }
  return launch_gemm_tc(dZ, ws, p, (int)R, Co, C, nsplit, nullptr, stream);
}
extern "C" int tn_gemm_tc_dwbwd(const float* dZ, const float* ws, const float* zprev, float* dzprev, const float* dw_w,
                                float* g_dw, float* g_dbias, float* g_dscale, float* g_dshift, const float* scale,
                                const float* shift, int relu, float drop_p, const unsigned long long* seed,
                                unsigned int layer, int B, int T, int Co, int C, int K, int nsplit, void* stream) {
  return dwbwd_impl(dZ, ws, zprev, dzprev, dw_w, g_dw, g_dbias, g_dscale, g_dshift, scale, shift, relu, drop_p, seed, layer, B, T, Co, C, K,
                    nsplit, nullptr, stream);
}
// the same with the BatchNorm backward of the conv's OUTPUT folded into the operand load (dZ is the direct gradient w.r.t.
// the pre-BatchNorm tensor z; see tn_gemm_tc_bnbwd)
extern "C" int tn_gemm_tc_dwbwd_bn(const float* dZ, const float* ws, const tn_bn_bwd* bnb, const float* zprev, float* dzprev,
                                   const float* dw_w, float* g_dw, float* g_dbias, float* g_dscale, float* g_dshift, const float* scale,
                                   const float* shift, int relu, float drop_p, const unsigned long long* seed, unsigned int layer,
                                   int B, int T, int Co, int C, int K, void* stream) {
  TN_REQUIRE(bnb, "gemm_tc_dwbwd_bn: null tn_bn_bwd");
  return dwbwd_impl(dZ, ws, zprev, dzprev, dw_w, g_dw, g_dbias, g_dscale, g_dshift, scale, shift, relu, drop_p, seed, layer, B, T, Co, C, K,
                    3, bnb, stream);
}
