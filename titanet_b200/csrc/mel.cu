// Waveform -> normalised log-mel spectrogram, one kernel.
//   STFT (center=True reflect padding n_fft/2, window zero-padded to n_fft, one-sided,
//   un-normalised) -> |.|^2 -> mel filterbank -> 10*log10(clamp(., 1e-10)) ->
//   L2 normalisation over the mel axis (eps 1e-12).
// Reference: transforms.MelSpectrogram.__call__ (src/transforms.py:158-184) with
// torchaudio Spectrogram(power=None) / MelScale / AmplitudeToDB (lines 134-144); the
// batched form reproduces datasets.collate_fn's zero padding of frames past each
// utterance's own length (src/datasets.py:48-73).
//
// One block transforms a tile of MEL_FT (a launch parameter) consecutive frames of one utterance.  Two real
// frames share one complex radix-2 FFT in shared memory (frame A in the real lane,
// frame B in the imaginary lane), the spectra are separated with the conjugate-symmetry
// identity, and the tile is written out in one coalesced pass in either layout.
#include "common.cuh"
#include <stdlib.h>

#define MEL_FT_MAX 32     // frames per block (even, chosen by the launcher so that the grid fills whole waves)
#define MEL_THREADS 256

// SpecAugment (src/transforms.py:168-175, 187-201) for given per-utterance draws; all pointers NULL = off.
//  - time stretch (torchaudio TimeStretch = phase vocoder, then .abs().pow(2): the accumulated phase never reaches the
//    power spectrogram, so only its magnitude interpolation alpha |X[i+1]| + (1 - alpha) |X[i]| at the fractional frame
//    positions arange(0, T, rate) is computed); an output frame then needs TWO source frames, which ride in the real
//    and imaginary lanes of one FFT.  rate == 1 is torchaudio's bypass.
//  - masks: n_fmask ranges over the mel axis then n_tmask ranges over the (stretched) time axis, value 0.0.
struct TnSpecAug {
  const double* rates;       // [B]
  const int* frames;         // [B] stretched frame count ceil(T_b / rate_b)
  const int* masks;          // [B, n_fmask + n_tmask, 2] half-open (start, end)
  int n_fmask, n_tmask;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(MEL_THREADS) mel_kernel(const float* __restrict__ wave, const int* __restrict__ lengths,
                                                          const float* __restrict__ window, const float* __restrict__ fb,
                                                          const int* __restrict__ band_lo, const int* __restrict__ band_hi,
                                                          float* __restrict__ out, int L_stride, int L_full, int T_out,
                                                          int N, int log2N, int hop, int n_mels, int nwc, TnSpecAug aug, int MEL_FT) {
  tn_grid_dep_sync();
  extern __shared__ float smem[];
  const int NF = N / 2 + 1;
  float2* buf = reinterpret_cast<float2*>(smem);             // N complex
  float2* tw = buf + N;                                      // N/2 twiddles
  float* win = reinterpret_cast<float*>(tw + N / 2);         // N
  float* pw = win + N;                                       // 2 * NF
  float* melv = pw + 2 * NF;                                 // MEL_FT * (n_mels + 1)
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * MEL_FT;
  // per-utterance lengths are device data: never trust them with an address.  A length beyond the row is clipped to it; a
  // length the reflect padding cannot serve (<= n_fft / 2, where torch.stft raises) yields an all-zero spectrogram.
  int L = lengths ? lengths[b] : L_full;
  if (L > L_stride) L = L_stride;
  const bool bad_len = L <= N / 2;
  if (bad_len) L = N / 2 + 1;
  const int T_src = 1 + L / hop;                             // frames of the (unstretched) spectrogram
  const double rate = aug.rates ? aug.rates[b] : 1.0;
  const bool stretch = rate != 1.0;
  const int T_valid = bad_len ? 0 : (stretch ? aug.frames[b] : T_src);   // frames this utterance produces; later ones are zero
  const int step = stretch ? 1 : 2;                          // output frames per FFT
  const float* x = wave + (size_t)b * L_stride;
  const int half = N / 2;

  for (int i = tid; i < half; i += MEL_THREADS) {
    float s, c;
    sincospif(-(float)i / (float)half, &s, &c);
    tw[i] = make_float2(c, s);
  }
  for (int i = tid; i < N; i += MEL_THREADS) win[i] = window[i];
  __syncthreads();

  for (int jp = 0; jp < MEL_FT; jp += step) {
    int ta = t0 + jp, tb = ta + 1;                            // source frames in the two FFT lanes
    bool va = ta < T_valid && ta < T_out, vb = tb < T_valid && tb < T_out;
    float alpha = 0.f;
    if (!va && !vb) {
      for (int i = tid; i < step * n_mels; i += MEL_THREADS) melv[(jp + i / n_mels) * (n_mels + 1) + i % n_mels] = 0.f;
      continue;        // uniform across the block
    }
    if (stretch) {     // output frame t0 + jp sits at time_steps[t0 + jp] = fp32(rate * index), like torch.arange
      const float pos = (float)(rate * (double)(t0 + jp));
      ta = (int)pos; tb = (int)(pos + 1.0f);
      alpha = pos - floorf(pos);
      va = ta < T_src; vb = tb < T_src;                       // frames past the end are the vocoder's zero padding
    }
    // windowed frames, bit-reversed order
    for (int n = tid; n < N; n += MEL_THREADS) {
      float w = win[n];
      int sa = ta * hop + n - half, sb = sa + hop;
      sa = sa < 0 ? -sa : (sa >= L ? 2 * (L - 1) - sa : sa);
      sb = sb < 0 ? -sb : (sb >= L ? 2 * (L - 1) - sb : sb);
      float xa = (va && w != 0.f) ? x[sa] * w : 0.f;
      float xb = (vb && w != 0.f) ? x[sb] * w : 0.f;
      buf[__brev((unsigned)n) >> (32 - log2N)] = make_float2(xa, xb);
    }
    __syncthreads();
    for (int s = 0; s < log2N; ++s) {
      const int h = 1 << s;
      for (int i = tid; i < half; i += MEL_THREADS) {
        const int pos = i & (h - 1);
        const int i0 = ((i >> s) << (s + 1)) + pos, i1 = i0 + h;
        const float2 w = tw[pos * (half >> s)];
        const float2 a = buf[i0], t = cmul(w, buf[i1]);
        buf[i0] = make_float2(a.x + t.x, a.y + t.y);
        buf[i1] = make_float2(a.x - t.x, a.y - t.y);
      }
      __syncthreads();
    }
    // split the two real spectra, take |.|^2
    for (int k = tid; k < NF; k += MEL_THREADS) {
      const float2 zk = buf[k], zc = buf[(N - k) & (N - 1)];
      const float ar = 0.5f * (zk.x + zc.x), ai = 0.5f * (zk.y - zc.y);     // X_A = (Z[k] + conj Z[N-k]) / 2
      const float br = 0.5f * (zk.y + zc.y), bi = -0.5f * (zk.x - zc.x);    // X_B = (Z[k] - conj Z[N-k]) / (2i)
      const float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
      if (stretch) {
        const float mg = alpha * sqrtf(pb) + (1.f - alpha) * sqrtf(pa);
        pw[k] = mg * mg;
      } else {
        pw[k] = pa;
        pw[NF + k] = pb;
      }
    }
    __syncthreads();
    for (int i = tid; i < step * n_mels; i += MEL_THREADS) {
      const int j = i / n_mels, m = i - j * n_mels;
      const bool valid = stretch ? true : (j == 0 ? va : vb);
      float acc = 0.f;
      const float* p = pw + j * NF;
      const int lo = band_lo[m], hi = band_hi[m];
      for (int f = lo; f < hi; ++f) acc = fmaf(p[f], __ldg(fb + (size_t)f * n_mels + m), acc);
      melv[(jp + j) * (n_mels + 1) + m] = valid ? 10.0f * log10f(fmaxf(acc, 1e-10f)) : 0.f;
    }
    __syncthreads();
  }
  __syncthreads();
  // L2-normalise every frame of the tile over the mel axis (warp per frame)
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp; j < MEL_FT; j += MEL_THREADS / 32) {
    float* row = melv + j * (n_mels + 1);
    float s = 0.f;
    for (int m = lane; m < n_mels; m += 32) s = fmaf(row[m], row[m], s);
    s = tn_warp_sum(s);
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    for (int m = lane; m < n_mels; m += 32) row[m] *= inv;
  }
  __syncthreads();
  if (aug.masks) {
    const int* mk = aug.masks + (size_t)b * (aug.n_fmask + aug.n_tmask) * 2;
    for (int i = tid; i < MEL_FT * n_mels; i += MEL_THREADS) {
      const int j = i / n_mels, m = i - j * n_mels, t = t0 + j;
      bool hit = false;
      for (int q = 0; q < aug.n_fmask; ++q) hit |= (m >= mk[2 * q] && m < mk[2 * q + 1]);
      for (int q = aug.n_fmask; q < aug.n_fmask + aug.n_tmask; ++q) hit |= (t >= mk[2 * q] && t < mk[2 * q + 1]);
      if (hit) melv[j * (n_mels + 1) + m] = 0.f;
    }
    __syncthreads();
  }
  const int total = MEL_FT * n_mels;
  if (nwc) {
    for (int i = tid; i < total; i += MEL_THREADS) {
      const int j = i / n_mels, m = i - j * n_mels;
      if (t0 + j < T_out) out[((size_t)b * T_out + t0 + j) * n_mels + m] = melv[j * (n_mels + 1) + m];
    }
  } else {
    for (int i = tid; i < total; i += MEL_THREADS) {
      const int m = i / MEL_FT, j = i - m * MEL_FT;
      if (t0 + j < T_out) out[((size_t)b * n_mels + m) * T_out + t0 + j] = melv[j * (n_mels + 1) + m];
    }
  }
}

// wave [B, L_stride] (utterance b uses its first lengths[b] samples, or L_full when
// lengths == NULL) -> out [B, n_mels, T_out] (nwc = 0) or [B, T_out, n_mels] (nwc = 1).
// window: [n_fft] (already zero-padded); fb: [n_fft/2+1, n_mels]; band_lo/hi: [n_mels]
// half-open ranges of non-zero filterbank rows.  Frames >= 1 + L_b/hop are zero filled.
static int mel_launch(const float* wave, const int* lengths, const float* window, const float* fb, const int* band_lo,
                      const int* band_hi, float* out, int B, int L_stride, int L_full, int T_out, int n_fft, int hop, int n_mels,
                      int nwc, TnSpecAug aug, void* stream) {
  TN_REQUIRE(wave && window && fb && band_lo && band_hi && out, "mel_fwd: null tensor");
  TN_REQUIRE(B > 0 && B <= 65535 && T_out > 0 && hop > 0 && n_mels > 0 && n_mels <= 256, "mel_fwd: bad shape B=%d T_out=%d hop=%d n_mels=%d", B, T_out, hop, n_mels);
  int log2N = 0;
  while ((1 << log2N) < n_fft) ++log2N;
  TN_UNSUPPORTED((1 << log2N) != n_fft || n_fft < 64 || n_fft > 4096, "mel_fwd: n_fft=%d must be a power of two in [64, 4096]", n_fft);
  TN_REQUIRE(lengths || L_full > n_fft / 2, "mel_fwd: reflect padding needs more than n_fft/2 samples (L=%d)", L_full);
  TN_REQUIRE(L_full <= L_stride, "mel_fwd: L_full > L_stride");
  const int NF = n_fft / 2 + 1;
  auto smem_for = [&](int ft) { return sizeof(float) * ((size_t)2 * n_fft + n_fft + n_fft + 2 * NF + (size_t)ft * (n_mels + 1)); };
  TN_CUDA(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(MEL_FT_MAX)));
  // frames per block: the even count in [8, 32] that minimises waves x frames.  B = 64, T = 301 with 16 frames per block is
  // 1216 blocks on 148 x 8 resident ones -- a second wave of 32 blocks that doubled the kernel's time (127 us).
  static int per_sm = 0;
  if (per_sm == 0) {
    TN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mel_kernel, MEL_THREADS, smem_for(MEL_FT_MAX)));
    if (per_sm < 1) per_sm = 1;
  }
  const long long cap = (long long)per_sm * tn_num_sms();
  int ft = 16;
  long long best = -1;
  for (int f = 8; f <= MEL_FT_MAX; f += 2) {
    const long long blocks = (long long)tn_cdiv(T_out, f) * B, cost = ((blocks + cap - 1) / cap) * (f + 2);     // + 2: the block's twiddle / window set-up
    if (best < 0 || cost < best || (cost == best && f == 16)) { best = cost; ft = f; }
  }
  if (const char* e = getenv("TN_MEL_FT")) { const int v = atoi(e); if (v >= 2 && v <= MEL_FT_MAX && v % 2 == 0) ft = v; }    // tuning knob
  const size_t smem = smem_for(ft);
  dim3 grid(tn_cdiv(T_out, ft), B);
  tn_launch(mel_kernel, grid, MEL_THREADS, smem, stream, wave, lengths, window, fb, band_lo, band_hi, out, L_stride,
                                                                 L_full, T_out, n_fft, log2N, hop, n_mels, nwc, aug, ft);
  TN_LAUNCH_CHECK("mel_kernel");
  return TN_OK;
}

// wave [B, L_stride] (utterance b uses its first lengths[b] samples, or L_full when
// lengths == NULL) -> out [B, n_mels, T_out] (nwc = 0) or [B, T_out, n_mels] (nwc = 1).
// window: [n_fft] (already zero-padded); fb: [n_fft/2+1, n_mels]; band_lo/hi: [n_mels]
// half-open ranges of non-zero filterbank rows.  Frames >= 1 + L_b/hop are zero filled.
extern "C" int tn_mel_fwd(const float* wave, const int* lengths, const float* window, const float* fb, const int* band_lo,
                          const int* band_hi, float* out, int B, int L_stride, int L_full, int T_out, int n_fft, int hop,
                          int n_mels, int nwc, void* stream) {
  TnSpecAug aug = {nullptr, nullptr, nullptr, 0, 0};
  return mel_launch(wave, lengths, window, fb, band_lo, band_hi, out, B, L_stride, L_full, T_out, n_fft, hop, n_mels, nwc, aug,
                    stream);
}

// The same with SpecAugment for given draws: rates [B] fp64 (1.0 = not stretched) and frames [B] = ceil(T_b / rate_b)
// (both NULL = no stretching); masks [B, n_fmask + n_tmask, 2] int32 half-open ranges (NULL = no masks; an empty range
// masks nothing).  Frames >= frames[b] are zero filled.
extern "C" int tn_mel_specaug_fwd(const float* wave, const int* lengths, const float* window, const float* fb,
                                  const int* band_lo, const int* band_hi, const double* rates, const int* frames,
                                  const int* masks, int n_fmask, int n_tmask, float* out, int B, int L_stride, int L_full,
                                  int T_out, int n_fft, int hop, int n_mels, int nwc, void* stream) {
  TN_REQUIRE((rates == nullptr) == (frames == nullptr), "mel_specaug_fwd: rates and frames go together");
  TN_REQUIRE(n_fmask >= 0 && n_tmask >= 0 && n_fmask + n_tmask <= 64, "mel_specaug_fwd: bad mask counts %d / %d", n_fmask, n_tmask);
  TN_REQUIRE(masks != nullptr || n_fmask + n_tmask == 0, "mel_specaug_fwd: mask counts without masks");
  TnSpecAug aug = {rates, frames, (n_fmask + n_tmask) ? masks : nullptr, n_fmask, n_tmask};
  return mel_launch(wave, lengths, window, fb, band_lo, band_hi, out, B, L_stride, L_full, T_out, n_fft, hop, n_mels, nwc, aug,
                    stream);
}
