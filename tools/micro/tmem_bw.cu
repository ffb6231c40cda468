// Micro-benchmark: TMEM -> register bandwidth (tcgen05.ld 32x32b.x32) with 4 / 8 warps, alone and while one thread
// streams tcgen05.mma (TS mode, N=64) -- is the attention backward bound by TMEM reads?
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_prims.cuh"
using namespace smx;

__global__ void __launch_bounds__(320, 1) k(long long* out, int reps, int ld_warps, int with_mma, int batch) {
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
  const int warp = threadIdx.x >> 5;
  if (warp == 1 && (threadIdx.x & 31) == 0 && with_mma) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, false, false);
    const uint32_t sb = smem_u32(smem);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t bd = umma_smem_desc(sb + 16384 + kk * 32, 16, 1024, kLayoutSW128);
        umma_ts(tm + 384 + (r & 1) * 64, tm + 320 + kk * 8, bd, idesc, kk > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[2] = clock64() - t0;
  }
  if (warp >= 2 && warp < 2 + ld_warps) {
    const uint32_t t_lane = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      uint32_t v[32], w[32];
      tmem_ld_x32(t_lane + ((r * 64) & 255), v);
      if (batch == 2) tmem_ld_x32(t_lane + ((r * 64 + 32) & 255), w);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
      if (batch == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= w[i];
      }
    }
    long long t1 = clock64();
    if (acc == 0x12345678) out[7] = acc;
    if (threadIdx.x == 64) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 2048;
  for (int with_mma = 0; with_mma < 2; ++with_mma)
    for (int ldw : {4, 8})
      for (int batch : {1, 2}) {
        cudaMemset(d, 0, 64);
        k<<<1, 320, 64 * 1024>>>(d, reps, ldw, with_mma, batch);
        long long h[8];
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        const double bytes = (double)reps * batch * 32 * 4 * 32 * ldw;   // all warps
        printf("ld_warps %d batch %d mma %d : %.1f B/clk TMEM->RF   (%.1f clk per x32 ld per warp)   mma %.1f clk/MMA  (%s)\n",
               ldw, batch, with_mma, bytes / h[0], (double)h[0] / (reps * batch), with_mma ? h[2] / (4.0 * reps) : 0.0,
               cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
