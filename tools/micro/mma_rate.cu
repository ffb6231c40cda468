// Micro-benchmark: issue cost and throughput of tcgen05.mma (cta_group::1, kind::f16, SS mode) for several N,
// K-major / MN-major B.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../speechmix_b200/csrc mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_prims.cuh"
using namespace smx;

template <int N, bool B_MN>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  fence_proxy_async_smem();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, B_MN);
    const uint32_t sb = smem_u32(smem);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = umma_smem_desc(sb + kk * 32, 16, 1024, kLayoutSW128);
        const uint64_t bd = B_MN ? umma_smem_desc(sb + 16384 + kk * 2048, 8192, 1024, kLayoutSW128)
                                 : umma_smem_desc(sb + 16384 + kk * 32, 16, 1024, kLayoutSW128);
        umma_ss(tm + (r & 1) * 256, ad, bd, idesc, kk > 0 ? 1u : 0u);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, bool B_MN>
__global__ void __launch_bounds__(128, 1) kts(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  fence_proxy_async_smem();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, B_MN);
    const uint32_t sb = smem_u32(smem);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t bd = B_MN ? umma_smem_desc(sb + 16384 + kk * 2048, 8192, 1024, kLayoutSW128)
                                 : umma_smem_desc(sb + 16384 + kk * 32, 16, 1024, kLayoutSW128);
        umma_ts(tm + (r & 1) * 128, tm + 384 + kk * 8, bd, idesc, kk > 0 ? 1u : 0u);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, bool B_MN>
void runts(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(kts<N, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 256;
  for (int w = 0; w < 2; ++w) kts<N, B_MN><<<1, 128, 64 * 1024>>>(d, reps);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s issue %.1f clk/MMA   complete %.1f clk/MMA  (err %s)\n", name, h[0] / (4.0 * reps), h[1] / (4.0 * reps),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

template <int N, bool B_MN>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<N, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 256;
  for (int w = 0; w < 2; ++w) k<N, B_MN><<<1, 128, 64 * 1024>>>(d, reps);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s issue %.1f clk/MMA   complete %.1f clk/MMA  (err %s)\n", name, h[0] / (4.0 * reps), h[1] / (4.0 * reps),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<64, false>("128x64x16  B K-major");
  run<64, true>("128x64x16  B MN-major");
  run<128, false>("128x128x16 B K-major");
  run<128, true>("128x128x16 B MN-major");
  run<256, false>("128x256x16 B K-major");
  runts<64, false>("TS 128x64x16  B K-major");
  runts<64, true>("TS 128x64x16  B MN-major");
  runts<128, false>("TS 128x128x16 B K-major");
  runts<128, true>("TS 128x128x16 B MN-major");
  return 0;
}
