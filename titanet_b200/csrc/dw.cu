// Depthwise K-tap convolution along time in NWC layout, with the producer's
// BatchNorm-affine + ReLU + dropout applied while loading ("lazy activation").
// Reference: modules.DepthwiseConv1d's first conv, Conv1dSamePadding(C, C, K, groups=C,
// bias=True) with zero "same" padding (K-1)//2 (src/modules.py:30-40, 64-75).
#include "common.cuh"

#define DW_THREADS TN_EW_THREADS

template <int K>
__global__ void __launch_bounds__(DW_THREADS) dw_fwd_kernel(const float* __restrict__ z, float* __restrict__ u,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            TnAct act, int R, int C, int T, int rpb) {
  tn_grid_dep_sync();
  constexpr int PAD = K / 2;
  act = tn_act_init(act);
  TnTile tl = tn_tile(C);
  const int run = (rpb + tl.lanes - 1) / tl.lanes;
  const int blk0 = blockIdx.x * rpb;
  const int ra = blk0 + tl.lane * run;
  const int rb = min(min(R, blk0 + rpb), ra + run);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    if (!tl.active || q >= tl.Q || ra >= rb) continue;
    const int c = 4 * q;
    float4 wk[K];
#pragma unroll
    for (int j = 0; j < K; ++j)
      wk[j] = make_float4(__ldg(w + (c + 0) * K + j), __ldg(w + (c + 1) * K + j), __ldg(w + (c + 2) * K + j), __ldg(w + (c + 3) * K + j));
    const float4 b4 = bias ? tn_ld4(bias + c) : tn_zero4();
    auto load = [&](int s) -> float4 {
      if (s < 0 || s >= R) return tn_zero4();
      size_t off = (size_t)s * C + c;
      return tn_act4(act, tn_ld4(z + off), c, off >> 2, nullptr);
    };
    float4 win[K];
#pragma unroll
    for (int j = 0; j < K - 1; ++j) win[j] = load(ra - PAD + j);
    int t = ra % T;
    // rows in batches of four: the four raw loads are issued back to back before any of them is used (the ncu capture showed
    // the loop latency-bound: long-scoreboard stalls 6 per issued instruction at 33 % occupancy with one load in flight)
    for (int r4 = ra; r4 < rb; r4 += 4) {
      float4 raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int srow = r4 + i + PAD;
        raw[i] = (r4 + i < rb && srow >= 0 && srow < R) ? tn_ld4(z + (size_t)srow * C + c) : tn_zero4();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r4 + i;
        if (r < rb) {
          const int srow = r + PAD;
          const size_t off = (size_t)srow * C + c;
          win[K - 1] = (srow >= 0 && srow < R) ? tn_act4(act, raw[i], c, off >> 2, nullptr) : tn_zero4();
          float4 acc = b4;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            int tt = t + j - PAD;
            if (tt >= 0 && tt < T) acc = tn_fma4(wk[j], win[j], acc);
          }
          tn_st4(u + (size_t)r * C + c, acc);
#pragma unroll
          for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
          if (++t == T) t = 0;
        }
      }
    }
  }
}

// backward: da = dw^T(du); dz = da * act'(z) * scale; accumulates dscale, dshift, dw, dbias
template <int K>
__global__ void __launch_bounds__(DW_THREADS, (K <= 5 ? 2 : 1)) dw_bwd_kernel(const float* __restrict__ du, const float* __restrict__ z,
                                                            float* __restrict__ dz, const float* __restrict__ w,
                                                            float* __restrict__ dw, float* __restrict__ dbias,
                                                            float* __restrict__ dscale, float* __restrict__ dshift,
                                                            TnAct act, int R, int C, int T, int rpb) {
  tn_grid_dep_sync();
  constexpr int PAD = K / 2;
  act = tn_act_init(act);
  __shared__ float4 red[DW_THREADS];
  TnTile tl = tn_tile(C);
  const int run = (rpb + tl.lanes - 1) / tl.lanes;
  const int blk0 = blockIdx.x * rpb;
  const int ra = blk0 + tl.lane * run;
  const int rb = min(min(R, blk0 + rpb), ra + run);
  for (int qb = 0; qb < tl.Q; qb += tl.qpb) {
    const int q = qb + tl.q0;
    const int c = 4 * q;
    float4 a_sc = tn_zero4(), a_sh = tn_zero4(), a_b = tn_zero4();
    float4 a_w[K];
#pragma unroll
    for (int j = 0; j < K; ++j) a_w[j] = tn_zero4();
    if (tl.active && q < tl.Q && ra < rb) {
      float4 wk[K];
#pragma unroll
      for (int j = 0; j < K; ++j)
        wk[j] = make_float4(__ldg(w + (c + 0) * K + j), __ldg(w + (c + 1) * K + j), __ldg(w + (c + 2) * K + j), __ldg(w + (c + 3) * K + j));
      const float4 sc = act.scale ? tn_ld4(act.scale + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      auto load_a = [&](int s) -> float4 {
        if (s < 0 || s >= R) return tn_zero4();
        size_t off = (size_t)s * C + c;
        return tn_act4(act, tn_ld4(z + off), c, off >> 2, nullptr);
      };
      auto load_g = [&](int s) -> float4 {
        if (s < 0 || s >= R) return tn_zero4();
        return tn_ld4(du + (size_t)s * C + c);
      };
      float4 aw[K], gw[K];
#pragma unroll
      for (int j = 0; j < K - 1; ++j) { aw[j] = load_a(ra - PAD + j); gw[j] = load_g(ra - PAD + j); }
      int t = ra % T;
#pragma unroll 2
      for (int r = ra; r < rb; ++r) {
        aw[K - 1] = load_a(r + PAD);
        gw[K - 1] = load_g(r + PAD);
        const float4 g0 = gw[PAD];
        float4 da = tn_zero4();
#pragma unroll
        for (int j = 0; j < K; ++j) {
          // forward: u[r'] += w[j] * a[r' + j - PAD]  =>  a[r] feeds u[r + PAD - j]
          int tu = t + PAD - j;
          if (tu >= 0 && tu < T) da = tn_fma4(wk[j], gw[K - 1 - j], da);
          int ta = t + j - PAD;
          if (ta >= 0 && ta < T) a_w[j] = tn_fma4(g0, aw[j], a_w[j]);
        }
        a_b = a_b + g0;
        size_t off = (size_t)r * C + c;
        if (act.scale) {
          float4 zz = tn_ld4(z + off), m;
          tn_act4(act, zz, c, off >> 2, &m);
          float4 g = da * m;
          a_sc = tn_fma4(g, zz, a_sc);
          a_sh = a_sh + g;
          tn_st4(dz + off, g * sc);
        } else {
          tn_st4(dz + off, da);
        }
#pragma unroll
        for (int j = 0; j < K - 1; ++j) { aw[j] = aw[j + 1]; gw[j] = gw[j + 1]; }
        if (++t == T) t = 0;
      }
    }
    if (dbias) tn_lane_reduce_atomic(tl, a_b, q, dbias, red);
    if (act.scale) {
      tn_lane_reduce_atomic(tl, a_sc, q, dscale, red);
      tn_lane_reduce_atomic(tl, a_sh, q, dshift, red);
    }
    // dw is [C, K]: one lane-reduction per tap (consecutive threads -> consecutive channels)
#pragma unroll
    for (int j = 0; j < K; ++j) {
      __syncthreads();
      red[threadIdx.x] = tl.active ? a_w[j] : tn_zero4();
      __syncthreads();
      const float* rf = reinterpret_cast<const float*>(red);
      const int nch = min(4 * tl.qpb, 4 * (tl.Q - qb));
      for (int ch = threadIdx.x; ch < nch; ch += DW_THREADS) {
        float s = rf[ch];
        for (int l = 1; l < tl.lanes; ++l) s += rf[ch + l * 4 * tl.qpb];
        atomicAdd(dw + (size_t)(4 * qb + ch) * K + j, s);
      }
    }
  }
}

#include <stdlib.h>
static int dw_rows_per_block(int C, int K, long long R) {
  int Q = C / 4, qpb = Q < DW_THREADS ? Q : DW_THREADS, lanes = DW_THREADS / qpb;
  int run = K <= 5 ? 8 : 16;
  // one wave: at most 3 blocks per SM (602 blocks of 8-row runs on 148 SMs x 4 resident blocks left a 10-block second wave;
  // measured 12.3 -> 11.2 us at R=19264, C=256)
  const long long one_wave = (R + (long long)lanes * 3 * tn_num_sms() - 1) / ((long long)lanes * 3 * tn_num_sms());
  if (one_wave > run) run = (int)(one_wave > 64 ? 64 : one_wave);
  if (const char* e = getenv("TN_DW_RUN")) run = atoi(e) > 0 ? atoi(e) : run;      // tuning knob
  return lanes * run;
}

#define DW_DISPATCH(K_, CALL)                        \
  switch (K_) {                                      \
    case 1: { constexpr int KK = 1; CALL; } break;   \
    case 3: { constexpr int KK = 3; CALL; } break;   \
    case 5: { constexpr int KK = 5; CALL; } break;   \
    case 7: { constexpr int KK = 7; CALL; } break;   \
    case 9: { constexpr int KK = 9; CALL; } break;   \
    case 11: { constexpr int KK = 11; CALL; } break; \
    case 13: { constexpr int KK = 13; CALL; } break; \
    case 15: { constexpr int KK = 15; CALL; } break; \
    default:                                         \
      tn_set_error("depthwise conv: unsupported kernel size %d (odd sizes 1..15)", K_); \
      return TN_EUNSUPPORTED;                        \
  }

// u[R,C] = bias + depthwise_K(act(z))   (act = identity when scale == NULL)
extern "C" int tn_dw_fwd(const float* z, float* u, const float* w, const float* bias, const float* scale, const float* shift,
                         int relu, float drop_p, const unsigned long long* seed, unsigned int layer, int B, int T, int C, int K,
                         void* stream) {
  TN_REQUIRE(B > 0 && T > 0 && C > 0 && C % 4 == 0, "dw_fwd: need C %% 4 == 0 (B=%d T=%d C=%d)", B, T, C);
  TN_REQUIRE(z && u && w, "dw_fwd: null tensor");
  TN_REQUIRE(tn_aligned16(z) && tn_aligned16(u) && (!bias || tn_aligned16(bias)) && (!scale || (tn_aligned16(scale) && tn_aligned16(shift))),
             "dw_fwd: pointers must be 16B aligned");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "dw_fwd: B*T too large");
  int rpb = dw_rows_per_block(C, K, R);
  TnAct act = tn_make_act(scale, shift, relu, drop_p, seed, layer);
  DW_DISPATCH(K, (tn_launch(dw_fwd_kernel<KK>, tn_cdiv(R, rpb), DW_THREADS, 0, stream, z, u, w, bias, act, (int)R, C, T, rpb)));
  TN_LAUNCH_CHECK("dw_fwd_kernel");
  return TN_OK;
}

// dz = d/dz of the above given du; dw[C,K], dbias[C], dscale[C], dshift[C] are ACCUMULATED into.
extern "C" int tn_dw_bwd(const float* du, const float* z, float* dz, const float* w, float* dw, float* dbias, float* dscale,
                         float* dshift, const float* scale, const float* shift, int relu, float drop_p,
                         const unsigned long long* seed, unsigned int layer, int B, int T, int C, int K, void* stream) {
  TN_REQUIRE(B > 0 && T > 0 && C > 0 && C % 4 == 0, "dw_bwd: need C %% 4 == 0 (B=%d T=%d C=%d)", B, T, C);
  TN_REQUIRE(du && z && dz && w && dw, "dw_bwd: null tensor");
  TN_REQUIRE(!scale || (shift && dscale && dshift), "dw_bwd: scale given without shift/dscale/dshift");
  TN_REQUIRE(tn_aligned16(du) && tn_aligned16(z) && tn_aligned16(dz) && (!scale || (tn_aligned16(scale) && tn_aligned16(shift))),
             "dw_bwd: pointers must be 16B aligned");
  long long R = (long long)B * T;
  TN_REQUIRE(R < (1ll << 31), "dw_bwd: B*T too large");
  int rpb = dw_rows_per_block(C, K, R);
  TnAct act = tn_make_act(scale, shift, relu, drop_p, seed, layer);
  DW_DISPATCH(K, (tn_launch(dw_bwd_kernel<KK>, tn_cdiv(R, rpb), DW_THREADS, 0, stream, du, z, dz, w, dw, dbias, dscale, dshift, act, (int)R, C, T, rpb)));
  TN_LAUNCH_CHECK("dw_bwd_kernel");
  return TN_OK;
}
