// Micro-benchmark: per-SM throughput of MUFU.EX2 / MUFU.TANH, of a degree-3 polynomial exp2 on the FMA pipe (FFMA2),
// and of both issued together -- the numbers the attention softmax roofline is built on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../speechmix_b200/csrc mufu_rate.cu -o mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_prims.cuh"
using namespace smx;

// exp2 of x <= 0 on the FMA pipe: round-to-integer by the magic-number add, cubic on the fraction, exponent by integer add
__device__ __forceinline__ f32x2 poly_ex2_pair(f32x2 x) {
  const f32x2 magic = f2_rep(12582912.0f);                 // 1.5 * 2^23
  const f32x2 t = f2_add(x, magic);                        // integer part in the low mantissa bits
  const f32x2 fl = f2_add(t, f2_rep(-12582912.0f));
  const f32x2 f = f2_add(x, f2_mul(fl, f2_rep(-1.0f)));    // fraction in [-0.5, 0.5]
  f32x2 p = f2_fma(f, f2_rep(0.0555041f), f2_rep(0.2402265f));
  p = f2_fma(p, f, f2_rep(0.6931472f));
  p = f2_fma(p, f, f2_rep(1.0f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
  return f2_pack(p0, p1);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* clk, int reps) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i + 1);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2_approx(a[i]) - 1.0f;
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = tanh_approx(a[i]) - 1.0f;
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        f32x2 v = f2_add(poly_ex2_pair(f2_pack(a[i], a[i + 1])), f2_rep(-1.0f));
        f2_unpack(v, a[i], a[i + 1]);
      }
    } else if (MODE == 3) {  // six MUFU + one polynomial pair per 8 values (25 % on the FMA pipe)
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] = ex2_approx(a[i]) - 1.0f;
      f32x2 v = f2_add(poly_ex2_pair(f2_pack(a[6], a[7])), f2_rep(-1.0f));
      f2_unpack(v, a[6], a[7]);
    } else if (MODE == 4) {  // fp32 pair -> packed bf16 conversions only (8 values = 4 F2FP)
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const uint32_t u = pack_bf16x2(a[i], a[i + 1]);
        a[i] = __uint_as_float(u << 16) - 1.0f, a[i + 1] = __uint_as_float(u & 0xffff0000u) - 1.0f;
      }
    } else {  // 5: the softmax mix -- 8 exponentials + 4 packs
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float e0 = ex2_approx(a[i]), e1 = ex2_approx(a[i + 1]);
        const uint32_t u = pack_bf16x2(e0, e1);
        a[i] = __uint_as_float(u << 16) - 1.0f, a[i + 1] = __uint_as_float(u & 0xffff0000u) - 1.0f;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int blocks_per_sm) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm, reps = 4096;
  float* out;
  long long* clk;
  cudaMalloc(&out, sizeof(float) * blocks * 256);
  cudaMalloc(&clk, sizeof(long long) * blocks);
  k<MODE><<<blocks, 256>>>(out, clk, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, clk, reps);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[8];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const double ops_per_sm = 8.0 * reps * 256 * blocks_per_sm;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  // block 0's own clock count is only meaningful with one block per SM; the event time covers every case
  printf("{\"micro\": \"%s\", \"blocks_per_sm\": %d, \"values_per_clk_per_sm_block0\": %.2f, \"ms\": %.3f, "
         "\"values_per_clk_per_sm_at_max_clock\": %.2f, \"err\": \"%s\"}\n", name, blocks_per_sm,
         8.0 * reps * 256 / (double)h[0], ms, ops_per_sm / (ms * 1e-3 * khz * 1e3), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out), cudaFree(clk);
}

int main() {
  for (int b : {1, 4}) {
    run<0>("mufu.ex2", b);
    run<1>("mufu.tanh", b);
    run<2>("poly exp2 (FFMA2)", b);
    run<3>("6 mufu + 2 poly per 8", b);
    run<4>("f2fp bf16x2 pack", b);
    run<5>("8 ex2 + 4 f2fp", b);
  }
  return 0;
}
