// Attentive statistics pooling: softmax over time per (utterance, channel) of the
// attention energies e, then the attention-weighted mean and standard deviation of x.
//   alpha = softmax_t(e);  mu = sum_t alpha x;  sigma = sqrt(clamp(sum_t alpha x^2 - mu^2, eps))
// Reference: models.AttentiveStatsPooling.forward (src/models.py:570-584).  The two
// linears in front (in_linear + tanh, out_linear; lines 564-567) are conv-GEMMs.
// NWC layout: e, x are [B, T, D]; one thread owns one (b, channel) column and walks T
// with an online (running-max) softmax, so e / alpha are never re-read.
#include "common.cuh"

// aux[b, 0, c] = log-sum-exp of e over t ; aux[b, 1, c] = sum_t alpha x^2
__global__ void __launch_bounds__(128) asp_pool_fwd_kernel(const float* __restrict__ e, const float* __restrict__ x,
                                                           float* __restrict__ pooled, float* __restrict__ aux, int T, int D, float eps) {
  tn_grid_dep_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= D) return;
  const float* ep = e + (size_t)b * T * D + c;
  const float* xp = x + (size_t)b * T * D + c;
  float M = -INFINITY, S = 0.f, A = 0.f, Q = 0.f;
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    const float ev = __ldg(ep + (size_t)t * D), xv = __ldg(xp + (size_t)t * D);
    if (ev > M) {
      const float f = expf(M - ev);     // exp(-inf) = 0 on the first step
      S *= f; A *= f; Q *= f;
      M = ev;
    }
    const float w = expf(ev - M);
    S += w;
    A = fmaf(w, xv, A);
    Q = fmaf(w * xv, xv, Q);
  }
  const float mu = A / S, q = Q / S;
  const float resid = q - mu * mu;
  pooled[(size_t)b * 2 * D + c] = mu;
  pooled[(size_t)b * 2 * D + D + c] = sqrtf(fmaxf(resid, eps));
  aux[((size_t)b * 2 + 0) * D + c] = M + logf(S);
  aux[((size_t)b * 2 + 1) * D + c] = q;
}

__global__ void __launch_bounds__(128) asp_pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ pooled,
                                                           const float* __restrict__ aux, const float* __restrict__ e,
                                                           const float* __restrict__ x, float* __restrict__ de,
                                                           float* __restrict__ dx, int T, int D, float eps) {
  tn_grid_dep_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= D) return;
  const float mu = pooled[(size_t)b * 2 * D + c], sigma = pooled[(size_t)b * 2 * D + D + c];
  const float lse = aux[((size_t)b * 2 + 0) * D + c], q = aux[((size_t)b * 2 + 1) * D + c];
  const float dmu = dpooled[(size_t)b * 2 * D + c], dsig = dpooled[(size_t)b * 2 * D + D + c];
  const float dq = (q - mu * mu >= eps) ? dsig / (2.f * sigma) : 0.f;   // clamp(min=eps) passes gradient when resid >= eps
  const float dmt = dmu - 2.f * mu * dq;
  const float kconst = mu * dmt + q * dq;                               // sum_t alpha_t dalpha_t
  size_t base = (size_t)b * T * D + c;
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    const size_t off = base + (size_t)t * D;
    const float ev = __ldg(e + off), xv = __ldg(x + off);
    const float a = expf(ev - lse);
    const float dalpha = xv * (dmt + xv * dq);
    de[off] = a * (dalpha - kconst);
    dx[off] = a * (dmt + 2.f * xv * dq);
  }
}

extern "C" int tn_asp_pool_fwd(const float* e, const float* x, float* pooled, float* aux, int B, int T, int D, float eps, void* stream) {
  TN_REQUIRE(e && x && pooled && aux && B > 0 && B <= 65535 && T > 0 && D > 0, "asp_pool_fwd: bad arguments (B=%d T=%d D=%d)", B, T, D);
  dim3 grid(tn_cdiv(D, 128), B);
  tn_launch(asp_pool_fwd_kernel, grid, 128, 0, stream, e, x, pooled, aux, T, D, eps);
  TN_LAUNCH_CHECK("asp_pool_fwd_kernel");
  return TN_OK;
}

extern "C" int tn_asp_pool_bwd(const float* dpooled, const float* pooled, const float* aux, const float* e, const float* x,
                               float* de, float* dx, int B, int T, int D, float eps, void* stream) {
  TN_REQUIRE(dpooled && pooled && aux && e && x && de && dx && B > 0 && B <= 65535 && T > 0 && D > 0, "asp_pool_bwd: bad arguments (B=%d T=%d D=%d)", B, T, D);
  dim3 grid(tn_cdiv(D, 128), B);
  tn_launch(asp_pool_bwd_kernel, grid, 128, 0, stream, dpooled, pooled, aux, e, x, de, dx, T, D, eps);
  TN_LAUNCH_CHECK("asp_pool_bwd_kernel");
  return TN_OK;
}
