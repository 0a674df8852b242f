// Loss heads: row-wise softmax cross-entropy and the angular-margin family
// (SphereFace / CosFace / ArcFace), forward + gradient w.r.t. the logits in one kernel,
// plus the in-place row normalisation of the class weights.
// Reference: losses.CELoss.forward (src/losses.py:32-44) and
// losses.AngularMarginLoss.forward (src/losses.py:77-132).
#include "common.cuh"

// first-occurrence argmax across a warp
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// one warp per row: loss_row[b] = lse - logit[y]; dlogits = (softmax - onehot) * inv_B
__global__ void __launch_bounds__(128) ce_kernel(const float* __restrict__ logits, const long long* __restrict__ targets,
                                                 float* __restrict__ loss_row, long long* __restrict__ preds,
                                                 float* __restrict__ dlogits, const float* __restrict__ gout, int B, int Cn,
                                                 float inv_B) {
  tn_grid_dep_sync();
  if (gout) inv_B *= __ldg(gout);
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* l = logits + (size_t)row * Cn;
  float mv = -INFINITY;
  int mi = 0x7fffffff;
  for (int j = lane; j < Cn; j += 32) {
    float v = l[j];
    if (v > mv) { mv = v; mi = j; }
  }
  warp_argmax(mv, mi);
  float s = 0.f;
  for (int j = lane; j < Cn; j += 32) s += expf(l[j] - mv);
  s = tn_warp_sum(s);
  const float lse = mv + logf(s);
  // a label outside [0, Cn) never indexes memory: its row's loss is NaN (F.cross_entropy raises; here the reference's own
  // non-finite-loss guard, src/learn.py:110-112, stops the run)
  const long long yt = targets[row];
  const bool bad = yt < 0 || yt >= Cn;
  const int y = bad ? 0 : (int)yt;
  if (dlogits)
    for (int j = lane; j < Cn; j += 32)
      dlogits[(size_t)row * Cn + j] = bad ? __int_as_float(0x7fc00000) : (expf(l[j] - lse) - (j == y ? 1.f : 0.f)) * inv_B;
  if (lane == 0) {
    loss_row[row] = bad ? __int_as_float(0x7fc00000) : lse - l[y];
    preds[row] = mi;
  }
}

// one warp per row of raw cosines (x_hat . w_hat, before the clamp)
__global__ void __launch_bounds__(128) margin_kernel(const float* __restrict__ raw, const float* __restrict__ norms,
                                                     const long long* __restrict__ targets, float* __restrict__ loss_row,
                                                     long long* __restrict__ preds, float* __restrict__ draw,
                                                     float* __restrict__ dnorm, const float* __restrict__ gout, int B, int Cn,
                                                     float scale, int use_norm_scale, float m1, float m2, float m3, float eps,
                                                     float inv_B) {
  tn_grid_dep_sync();
  if (gout) inv_B *= __ldg(gout);
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* rw = raw + (size_t)row * Cn;
  const float s = use_norm_scale ? norms[row] : scale;
  const long long yt = targets[row];
  const bool bad = yt < 0 || yt >= Cn;           // out-of-range label: NaN loss for the row, no out-of-bounds access
  const int y = bad ? 0 : (int)yt;
  float mv = -INFINITY;
  int mi = 0x7fffffff;
  float others = 0.f, cw = 0.f;       // sum_{j != y} exp(s c_j) and sum_{j != y} c_j exp(s c_j)
  for (int j = lane; j < Cn; j += 32) {
    const float c = fminf(fmaxf(rw[j], -1.f), 1.f);
    if (c > mv) { mv = c; mi = j; }
    if (j != y) {
      const float ex = expf(s * c);
      others += ex;
      cw = fmaf(c, ex, cw);
    }
  }
  warp_argmax(mv, mi);
  others = tn_warp_sum(others);
  cw = tn_warp_sum(cw);
  const float cy = fminf(fmaxf(rw[y], -1.f), 1.f);
  const float theta = acosf(cy);
  const float phi = m1 * theta + m2;
  const float cphi = cosf(phi), sphi = sinf(phi);
  const float num = s * (cphi - m3);
  const float enum_ = expf(num);
  const float den = enum_ + others + eps;
  const float dnum = enum_ / den - 1.f;                          // d loss_row / d num
  if (draw) {
    for (int j = lane; j < Cn; j += 32) {
      const float r = rw[j];
      const bool inside = r >= -1.f && r <= 1.f;                 // clamp passes gradient on the closed interval
      float g;
      if (j == y) {
        g = dnum * s * m1 * sphi / sqrtf(1.f - cy * cy);
      } else {
        const float c = fminf(fmaxf(r, -1.f), 1.f);
        g = s * expf(s * c) / den;
      }
      draw[(size_t)row * Cn + j] = bad ? __int_as_float(0x7fc00000) : (inside ? g * inv_B : 0.f);
    }
  }
  if (lane == 0) {
    loss_row[row] = bad ? __int_as_float(0x7fc00000) : logf(den) - num;
    preds[row] = mi;
    if (dnorm) dnorm[row] = use_norm_scale ? (dnum * (cphi - m3) + cw / den) * inv_B : 0.f;
  }
}

// deterministic mean of the per-row losses
__global__ void loss_mean_kernel(const float* __restrict__ loss_row, float* __restrict__ loss, int B) {
  tn_grid_dep_sync();
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += loss_row[i];
  s = tn_warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = tn_warp_sum(v);
    if (threadIdx.x == 0) *loss = v / (float)B;
  }
}

// W[row, :] /= max(||W[row, :]||, eps)   in place (F.normalize(dim=1))
__global__ void __launch_bounds__(128) rownorm_kernel(float* __restrict__ W, int rows, int cols, float eps) {
  tn_grid_dep_sync();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* w = W + (size_t)row * cols;
  float s = 0.f;
  for (int i = lane; i < cols; i += 32) s = fmaf(w[i], w[i], s);
  const float d = fmaxf(sqrtf(tn_warp_sum(s)), eps);
  for (int i = lane; i < cols; i += 32) w[i] = w[i] / d;
}

extern "C" int tn_ce_fwd_bwd(const float* logits, const long long* targets, float* loss_row, float* loss, long long* preds,
                             float* dlogits, const float* gout, int B, int Cn, void* stream) {
  TN_REQUIRE(logits && targets && loss_row && loss && preds && B > 0 && Cn > 0, "ce_fwd_bwd: bad arguments");
  tn_launch(ce_kernel, tn_cdiv(B, 4), 128, 0, stream, logits, targets, loss_row, preds, dlogits, gout, B, Cn, 1.0f / (float)B);
  TN_LAUNCH_CHECK("ce_kernel");
  tn_launch(loss_mean_kernel, 1, 256, 0, stream, loss_row, loss, B);
  TN_LAUNCH_CHECK("loss_mean_kernel");
  return TN_OK;
}

extern "C" int tn_margin_fwd_bwd(const float* raw_cos, const float* norms, const long long* targets, float* loss_row,
                                 float* loss, long long* preds, float* draw, float* dnorm, const float* gout, int B, int Cn,
                                 float scale, int use_norm_scale, float m1, float m2, float m3, float eps, void* stream) {
  TN_REQUIRE(raw_cos && targets && loss_row && loss && preds && B > 0 && Cn > 0, "margin_fwd_bwd: bad arguments");
  TN_REQUIRE(!use_norm_scale || norms, "margin_fwd_bwd: scale=None needs the input norms");
  tn_launch(margin_kernel, tn_cdiv(B, 4), 128, 0, stream, raw_cos, norms, targets, loss_row, preds, draw, dnorm, gout, B,
                                                                  Cn, scale, use_norm_scale, m1, m2, m3, eps, 1.0f / (float)B);
  TN_LAUNCH_CHECK("margin_kernel");
  tn_launch(loss_mean_kernel, 1, 256, 0, stream, loss_row, loss, B);
  TN_LAUNCH_CHECK("loss_mean_kernel");
  return TN_OK;
}

extern "C" int tn_rownorm_inplace(float* W, int rows, int cols, float eps, void* stream) {
  TN_REQUIRE(W && rows > 0 && cols > 0, "rownorm_inplace: bad arguments");
  tn_launch(rownorm_kernel, tn_cdiv(rows, 4), 128, 0, stream, W, rows, cols, eps);
  TN_LAUNCH_CHECK("rownorm_kernel");
  return TN_OK;
}
