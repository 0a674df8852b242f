// L2 -> SM feed rate microbenchmark (build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l2bw tools/l2bw.cu).
// Question: what bounds the operand feed of the tensor-core kernels, whose CTAs all re-stream the SAME weight tiles from L2 --
// the L2 slices' output, or each SM's ingest?  And does TMA multicast across a cluster lift it?
// Every CTA pulls `iters` chunks of `chunk` bytes into shared memory with cp.async.bulk (4 in flight):
//   mode 0: every CTA reads the same addresses (a weight tile)        mode 1: every CTA reads its own addresses (activations)
//   mode 2: like 0, but one CTA of each cluster fetches 1/CS of the chunk and multicasts it to the CS CTAs of the cluster
// Prints bytes DELIVERED to shared memory per second over all SMs.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#define NST 4
template <int MODE>
__global__ void __launch_bounds__(128, 1) feed_kernel(const uint8_t* __restrict__ buf, size_t region, uint32_t chunk, int iters, int cs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[NST], empty[NST];
  const uint32_t rank = MODE == 2 ? cta_rank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(s32(&full[s]), 1); mbar_init(s32(&empty[s]), MODE == 2 ? cs : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == 2) cluster_sync(); else __syncthreads();
  if (threadIdx.x == 0) {
    // producer: mode 0 / 2 walk the same addresses in every CTA, mode 1 a CTA-private region
    const size_t base = MODE == 1 ? ((size_t)blockIdx.x * (size_t)chunk * 16) % region : 0;
    for (int it = 0; it < iters; ++it) {
      const int s = it % NST;
      if (it >= NST) mbar_wait(s32(&empty[s]), ((it / NST) - 1) & 1);
      const size_t off = (base + (size_t)(it % 16) * chunk) % region;
      mbar_expect(s32(&full[s]), chunk);
      if (MODE == 2) {
        const uint32_t part = chunk / cs;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(s32(smem + (size_t)s * chunk + rank * part)), "l"(buf + off + rank * part), "r"(part), "r"(s32(&full[s])), "h"((uint16_t)((1u << cs) - 1)) : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(smem + (size_t)s * chunk)), "l"(buf + off), "r"(chunk), "r"(s32(&full[s])) : "memory");
      }
    }
  } else if (threadIdx.x == 32) {
    // consumer: releases the stage as soon as it is full (in every CTA of the cluster for the multicast mode)
    for (int it = 0; it < iters; ++it) {
      const int s = it % NST;
      mbar_wait(s32(&full[s]), (it / NST) & 1);
      if (MODE == 2) {
        for (int c = 0; c < cs; ++c) {
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(s32(&empty[s])), "r"(c));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
      }
    }
  }
  if (MODE == 2) cluster_sync(); else __syncthreads();
}


// mode 3: the same pipeline fed by 2-D TENSOR copies (cp.async.bulk.tensor.2d, SWIZZLE_128B) of boxes [rows x 32 floats] out
// of a row-major [*, 256] fp32 tensor -- the GEMM kernels' operand loads: every box row is a separate 128-byte piece
__global__ void __launch_bounds__(128, 1) feed_tensor_kernel(const __grid_constant__ CUtensorMap map, uint32_t box_rows, int boxes, int iters, int same) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[NST], empty[NST];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t box_bytes = box_rows * 128u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(s32(&full[s]), 1); mbar_init(s32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int row_base = same ? 0 : (int)blockIdx.x * (int)box_rows * boxes;
    for (int it = 0; it < iters; ++it) {
      const int s = it % NST;
      if (it >= NST) mbar_wait(s32(&empty[s]), ((it / NST) - 1) & 1);
      mbar_expect(s32(&full[s]), box_bytes * boxes);
      for (int b = 0; b < boxes; ++b)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s32(smem + ((size_t)s * boxes + b) * box_bytes)), "l"(&map), "r"(s32(&full[s])), "r"((it % 8) * 32), "r"(row_base + b * (int)box_rows) : "memory");
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % NST;
      mbar_wait(s32(&full[s]), (it / NST) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
  }
  __syncthreads();
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static float run_tensor(const uint8_t* buf, uint32_t box_rows, int boxes, int iters, int ctas, int same) {
  static EncodeFn enc = nullptr;
  if (!enc) { cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q); }
  CUtensorMap map;
  const cuuint64_t rows = (32u << 20) / 1024;
  cuuint64_t dims[2] = {256, rows}; cuuint64_t strides[1] = {1024}; cuuint32_t box[2] = {32, box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return -1.f; }
  const size_t smem = (size_t)NST * boxes * box_rows * 128 + 1024;
  cudaFuncSetAttribute(feed_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    feed_tensor_kernel<<<ctas, 128, smem>>>(map, box_rows, boxes, iters, same);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (cudaGetLastError() != cudaSuccess) { printf("tensor launch failed\n"); return -1.f; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

template <int MODE>
static float run(const uint8_t* buf, size_t region, uint32_t chunk, int iters, int ctas, int cs) {
  const size_t smem = (size_t)NST * chunk;
  cudaFuncSetAttribute(feed_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = MODE == 2 ? 1 : 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, feed_kernel<MODE>, buf, region, chunk, iters, cs);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(err)); return -1.f; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t region = 32u << 20;                  // 32 MB: L2-resident
  uint8_t* buf; cudaMalloc(&buf, region + (1u << 20)); cudaMemset(buf, 1, region + (1u << 20));   // slack: a chunk may start just below `region`
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 512;
  printf("SMs %d; %d chunks per CTA, %d in flight\n", sms, iters, NST);
  for (uint32_t chunk : {16384u, 32768u, 49152u}) {
    for (int ctas : {128, sms}) {
      const double bytes = (double)ctas * iters * chunk;
      float t0 = run<0>(buf, region, chunk, iters, ctas, 1), t1 = run<1>(buf, region, chunk, iters, ctas, 1);
      printf("chunk %5u B, %3d CTAs: same addresses %7.2f TB/s   private addresses %7.2f TB/s", chunk, ctas, bytes / t0 / 1e9, bytes / t1 / 1e9);
      if (ctas == 128)
        for (int cs : {2, 4, 8}) {
          float t2 = run<2>(buf, region, chunk, iters, ctas, cs);
          printf("   multicast x%d %7.2f TB/s", cs, t2 > 0 ? bytes / t2 / 1e9 : 0.0);
        }
      printf("\n");
    }
  }
  for (int boxes : {1, 2, 3})
    for (uint32_t box_rows : {64u, 128u}) {
      const double bytes = (double)sms * iters * boxes * box_rows * 128.0;
      float ta = run_tensor(buf, box_rows, boxes, iters, sms, 1), tb = run_tensor(buf, box_rows, boxes, iters, sms, 0);
      printf("tensor 2-D, %d box(es) of [%3u rows x 128 B] per stage, %d CTAs: same tile %6.2f TB/s (%5.1f GB/s per SM)   private tiles %6.2f TB/s (%5.1f GB/s per SM)\n",
             boxes, box_rows, sms, bytes / ta / 1e9, bytes / ta / 1e6 / sms, bytes / tb / 1e9, bytes / tb / 1e6 / sms);
    }
  // occupancy of clusters: how many clusters of 2 / 4 / 8 CTAs with 200 KB of shared memory can be co-resident
  for (int cs : {2, 4, 8}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaFuncSetAttribute(feed_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int n = 0;
    cudaOccupancyMaxActiveClusters(&n, feed_kernel<2>, &cfg);
    printf("max co-resident clusters of %d CTAs (200 KB smem each): %d = %d CTAs\n", cs, n, n * cs);
  }
  return 0;
}
