// Probe: the A operand of tcgen05.mma from TENSOR MEMORY, staged there by tcgen05.cp (smem -> TMEM), against the usual
// shared-memory A operand.  One CTA, cta_group::1, M = 128, N = 64, K = 32 fp32 (four tf32 MMAs of K = 8), both operands
// K-major SWIZZLE_128B.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/utccp_test tools/utccp_test.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

#define NN 64
__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int mode, int kind16) {
  __shared__ __align__(1024) uint8_t sa[128 * 128];
  __shared__ __align__(1024) uint8_t sb[NN * 128];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // K-major SWIZZLE_128B: element (r, k) at r * 128 + ((k / 4) ^ (r % 8)) * 16 + (k % 4) * 4
  for (int i = tid; i < 128 * 32; i += 128) {
    const int r = i >> 5, k = i & 31;
    *reinterpret_cast<float*>(sa + r * 128 + (((k >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2)) = A[i];
  }
  for (int i = tid; i < NN * 32; i += 128) {
    const int r = i >> 5, k = i & 31;
    *reinterpret_cast<float*>(sb + r * 128 + (((k >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2)) = B[i];
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  const uint32_t acol = tbase + 128;                     // A staging columns behind the accumulator
  if (tid == 0) {
    const uint32_t fmt = kind16 ? 0u : 2u;               // F16 : TF32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NN >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = desc_k128(s32(sa)), db = desc_k128(s32(sb));
    if (mode == 1) {
      for (int kk = 0; kk < 4; ++kk)
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(acol + 8 * kk), "l"(da + (uint64_t)(2 * kk)) : "memory");
    }
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t acc = kk > 0 ? 1u : 0u;
      if (mode == 1) {
        if (kind16)
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                       ::"r"(tbase), "r"(acol + 8 * kk), "l"(db + (uint64_t)(2 * kk)), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                       ::"r"(tbase), "r"(acol + 8 * kk), "l"(db + (uint64_t)(2 * kk)), "r"(idesc), "r"(acc) : "memory");
      } else {
        if (kind16)
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tbase), "l"(da + (uint64_t)(2 * kk)), "l"(db + (uint64_t)(2 * kk)), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tbase), "l"(da + (uint64_t)(2 * kk)), "l"(db + (uint64_t)(2 * kk)), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  mbar_wait(s32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // thread = lane (row of D) tid; columns 0 .. NN-1
  for (int c0 = 0; c0 < NN; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)tid * NN + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u) : "memory");
}


// timing: `reps` back-to-back tf32 MMAs (M = 128, N = n, K = 8) from one thread, A from shared memory (mode 0) or tensor memory (mode 1)
__global__ void __launch_bounds__(128, 1) rate(long long* out, int mode, int n, int reps) {
  extern __shared__ __align__(1024) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* sm = (uint8_t*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = sm; uint8_t* sb = sm + 128 * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 32; i += 128) reinterpret_cast<float*>(sm)[i] = 0.001f * (float)(i % 97);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot, acol = tbase + 384;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = desc_k128(s32(sa)), db = desc_k128(s32(sb));
    for (int kk = 0; kk < 4; ++kk)
      asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(acol + 8 * kk), "l"(da + (uint64_t)(2 * kk)) : "memory");
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int kk = i & 3;
      if (mode == 1)
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                     ::"r"(tbase), "r"(acol + 8 * kk), "l"(db + (uint64_t)(2 * kk)), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tbase), "l"(da + (uint64_t)(2 * kk)), "l"(db + (uint64_t)(2 * kk)), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    mbar_wait(s32(&bar), 0);
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

// the same MMA stream issued from WARP-UNIFORM code: all 32 lanes of warp 0 run the loop and compute the (uniform) descriptors,
// one elected lane issues.  Under `if (tid == 0)` the compiler wraps every tcgen05.mma in an ELECT / BRA.U.ANY serialisation loop
// and moves its operands from vector to uniform registers (R2UR) each time.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__global__ void __launch_bounds__(128, 1) rate_uniform(long long* out, int n, int reps) {
  extern __shared__ __align__(1024) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* sm = (uint8_t*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = sm; uint8_t* sb = sm + 128 * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 32; i += 128) reinterpret_cast<float*>(sm)[i] = 0.001f * (float)(i % 97);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = desc_k128(s32(sa)), db = desc_k128(s32(sb));
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tbase), "l"(da + (uint64_t)(2 * kk)), "l"(db + (uint64_t)(2 * kk)), "r"(idesc), "r"(1u) : "memory");
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(s32(&bar), 0);
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int main() {
  float hA[128 * 32], hB[NN * 32], hD[2][128 * NN];
  srand(1);
  for (auto& v : hA) v = (float)(rand() % 17 - 8) / 8.f;      // exactly representable in tf32: both paths must agree bit for bit
  for (auto& v : hB) v = (float)(rand() % 13 - 6) / 4.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD[0]));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dD, 0, sizeof(hD[0]));
    probe<<<1, 128>>>(dA, dB, dD, mode, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD[mode], dD, sizeof(hD[0]), cudaMemcpyDeviceToHost);
  }
  double e0 = 0, e1 = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < NN; ++n) {
      double ref = 0;
      for (int k = 0; k < 32; ++k) ref += (double)hA[m * 32 + k] * hB[n * 32 + k];
      e0 = fmax(e0, fabs(hD[0][m * NN + n] - ref));
      e1 = fmax(e1, fabs(hD[1][m * NN + n] - ref));
    }
  printf("tf32: max |err| A from shared memory %.3g, A from tensor memory (tcgen05.cp 128x256b) %.3g\n", e0, e1);
  if (e1 > 1e-3) {
    printf("row 0 / 1 / 9, first 6 columns, smem-A vs tmem-A:\n");
    for (int m : {0, 1, 9}) { for (int n = 0; n < 6; ++n) printf(" %8.3f/%8.3f", hD[0][m * NN + n], hD[1][m * NN + n]); printf("\n"); }
  }
  long long* dT; cudaMalloc(&dT, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int ctas : {1, 148})
    for (int n : {64, 144, 256})
      for (int mode = 0; mode < 2; ++mode) {
        long long hT[148];
        rate<<<ctas, 128, 64 * 1024>>>(dT, mode, n, 512);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("rate failed\n"); return 1; }
        rate<<<ctas, 128, 64 * 1024>>>(dT, mode, n, 512);
        cudaDeviceSynchronize();
        cudaMemcpy(hT, dT, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < ctas; ++i) mx = hT[i] > mx ? hT[i] : mx;
        printf("%3d CTA(s), M=128 N=%3d K=8 tf32, A from %s: %6.1f clk per MMA (floor %d)\n", ctas, n, mode ? "tensor memory" : "shared memory", mx / 512.0, n / 2);
        if (mode == 0) {
          cudaFuncSetAttribute(rate_uniform, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
          rate_uniform<<<ctas, 128, 64 * 1024>>>(dT, n, 512); cudaDeviceSynchronize();
          rate_uniform<<<ctas, 128, 64 * 1024>>>(dT, n, 512);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("rate_uniform failed\n"); return 1; }
          cudaMemcpy(hT, dT, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
          mx = 0; for (int i = 0; i < ctas; ++i) mx = hT[i] > mx ? hT[i] : mx;
          printf("%3d CTA(s), M=128 N=%3d K=8 tf32, A from shared memory, warp-uniform issue + elect.sync: %6.1f clk per MMA\n", ctas, n, mx / 512.0);
        }
      }
  return 0;
}
