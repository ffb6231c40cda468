// sm_100a device primitives shared by every kernel in libspeechmix_sm100:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld),
// shared-memory matrix descriptors and the bf16 instruction descriptor.
//
// Everything here is inline PTX; nothing is borrowed from CUTLASS at build time.
// Bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace smx {

using bf16 = __nv_bfloat16;

// ---------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug traps (-> a CUDA error on the host)
// after ~4 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fff) == 0 && globaltimer_ns() - t0 > 4000000000ull) {
      printf("smx: mbarrier watchdog block %d thread %d bar %u parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------
// proxy / tcgen05 fences
// ---------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1,
                                             int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// 1-D bulk copy global -> shared (no tensor map), completes on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- CTA-pair (cta_group::2) variants: loads of both CTAs signal the LEADER's mbarrier ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, 256 rows over the CTA pair] (+)= A * B; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma2_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this smem offset in every CTA of `mask` once all prior MMAs have retired
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ---------------------------------------------------------------------------
// TMEM allocation (one full warp; power-of-two columns >= 32)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------------------
// UMMA descriptors
// ---------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset>>4 [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0           [61,64) layout: 0 none, 2 = 128B swizzle
constexpr uint64_t kLayoutNone = 0, kLayoutSW128 = 2;
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint64_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}
// Instruction descriptor, kind::f16, bf16 x bf16 -> f32:
//   [4,6) D fmt = 1 (f32)  [7,10) A fmt = 1 (bf16)  [10,13) B fmt = 1 (bf16)
//   [15] A major (1 = MN)  [16] B major (1 = MN)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every MMA previously issued by this thread has retired
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------
// TMEM -> registers.  32x32b: lane i of the warp reads TMEM lane (base_lane+i),
// N consecutive 32-bit columns.  A warp may only touch lanes 32*(warp%4)..+31.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------
// math helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// three-input maximum (FMNMX3, sm_100): NaN-free inputs assumed
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf-GELU (torch.nn.functional.gelu, approximate="none") evaluated as x * sigmoid(2 u(x)) with
// u = x (c0 + c1 x^2 + c2 x^4), |x| clamped to 7 for u; coefficients are a minimax fit of
// atanh(erf(x / sqrt 2)):  max |gelu_fit - gelu_erf| = 2.6e-5, max |gelu'_fit - gelu'_erf| = 1.1e-4
// over the reals -- far below the bf16 rounding of the stored activation.
// ~10 issue slots (2 MUFU) instead of ~30 for erff, which is what lets the GEMM epilogue keep
// pace with the tensor pipe.
// sigmoid(2u) = 0.5 + 0.5 tanh(u): ONE MUFU (tanh.approx.f32, max rel. error 2^-11) instead of ex2 + rcp; the
// resulting absolute error of gelu / gelu' (< 1e-3 at |x| ~ 3) stays below the bf16 rounding of the stored value.
//   u(x)  = x (C0 + C1 x^2 + C2 x^4)                 (minimax fit of atanh(erf(x / sqrt 2)), |x| clamped to 7)
//   t(x)  = 2 x u'(x) = x (D0 + D1 x^2 + D2 x^4),    D_i = 2 (2i+1) C_i
constexpr float kGC0 = 7.97507884e-01f, kGC1 = 3.70056460e-02f, kGC2 = -3.51516790e-04f;
constexpr float kGD0 = 2.0f * kGC0, kGD1 = 6.0f * kGC1, kGD2 = 10.0f * kGC2;
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float xc = fminf(fmaxf(x, -7.0f), 7.0f);
  const float x2 = xc * xc;
  const float s = fmaf(tanh_approx(xc * fmaf(x2, fmaf(x2, kGC2, kGC1), kGC0)), 0.5f, 0.5f);
  return x * s;  // x * sigmoid(2u)
}
// derivative of the fit itself: s + x s (1 - s) 2u'(x), s = sigmoid(2u).  Outside the clamp s (1 - s) < 1e-21,
// so the clamped x can stand in for x in the second term.
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float xc = fminf(fmaxf(x, -7.0f), 7.0f);
  const float x2 = xc * xc;
  const float s = fmaf(tanh_approx(xc * fmaf(x2, fmaf(x2, kGC2, kGC1), kGC0)), 0.5f, 0.5f);
  const float t = xc * fmaf(x2, fmaf(x2, kGD2, kGD1), kGD0);
  return fmaf(fmaf(-s, s, s), t, s);
}
// gelu and its derivative from ONE sigmoid (forward epilogue that stores gelu'(pre) for the backward pass:
// the data-gradient GEMM's epilogue then is a single multiply).  14 issue slots, 1 MUFU per element.
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  const float xc = fminf(fmaxf(x, -7.0f), 7.0f);
  const float x2 = xc * xc;
  const float s = fmaf(tanh_approx(xc * fmaf(x2, fmaf(x2, kGC2, kGC1), kGC0)), 0.5f, 0.5f);
  const float t = xc * fmaf(x2, fmaf(x2, kGD2, kGD1), kGD0);
  y = x * s;
  dy = fmaf(fmaf(-s, s, s), t, s);
}
// ---------------------------------------------------------------------------
// packed fp32 pairs: sm_100 executes fma/mul/add .f32x2 as ONE instruction (FFMA2 / FMUL2 / FADD2) on an aligned
// register pair -- the same IEEE round-to-nearest results as two scalar operations in half the issue slots.
// Issue-bound epilogues and row kernels (GELU, LayerNorm) use them for everything but min/max, MUFU and converts.
// ---------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 f2_rep(float a) { return f2_pack(a, a); }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// pair versions of the GELU fits above: bit-identical to the scalar functions (same operations, same rounding)
__device__ __forceinline__ void gelu_erf_both2(float& x0, float& x1, float& d0, float& d1) {
  const f32x2 xc = f2_pack(fminf(fmaxf(x0, -7.0f), 7.0f), fminf(fmaxf(x1, -7.0f), 7.0f));
  const f32x2 x2 = f2_mul(xc, xc);
  float u0, u1;
  f2_unpack(f2_mul(xc, f2_fma(x2, f2_fma(x2, f2_rep(kGC2), f2_rep(kGC1)), f2_rep(kGC0))), u0, u1);
  const f32x2 s = f2_fma(f2_pack(tanh_approx(u0), tanh_approx(u1)), f2_rep(0.5f), f2_rep(0.5f));
  const f32x2 t = f2_mul(xc, f2_fma(x2, f2_fma(x2, f2_rep(kGD2), f2_rep(kGD1)), f2_rep(kGD0)));
  const f32x2 y = f2_mul(f2_pack(x0, x1), s);
  const f32x2 dy = f2_fma(f2_fma(f2_mul(s, f2_rep(-1.0f)), s, s), t, s);
  f2_unpack(y, x0, x1);
  f2_unpack(dy, d0, d1);
}
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const f32x2 xc = f2_pack(fminf(fmaxf(x0, -7.0f), 7.0f), fminf(fmaxf(x1, -7.0f), 7.0f));
  const f32x2 x2 = f2_mul(xc, xc);
  float u0, u1;
  f2_unpack(f2_mul(xc, f2_fma(x2, f2_fma(x2, f2_rep(kGC2), f2_rep(kGC1)), f2_rep(kGC0))), u0, u1);
  const f32x2 s = f2_fma(f2_pack(tanh_approx(u0), tanh_approx(u1)), f2_rep(0.5f), f2_rep(0.5f));
  f2_unpack(f2_mul(f2_pack(x0, x1), s), x0, x1);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ---------------------------------------------------------------------------
// Counter-based dropout (the train-mode dropout of the HF modules the reference runs, ref:speechmix/hf_model.py:397,
// 357-365).  A keep decision is a pure function of (seed, step, call, element), so the backward kernels REGENERATE
// the mask instead of storing it, and a replayed CUDA graph draws new masks because `step` lives in device memory
// (state[0] = seed, state[1] = step counter, advanced on the device once per training step).
// One 32-bit hash serves the two elements of an aligned pair, 16 bits each: keep <=> bits >= round(p * 65536).
// ---------------------------------------------------------------------------
struct DropKey {
  uint32_t key, thresh;   // thresh = 0 <=> dropout off
  float scale;            // 1 / (1 - p)
};
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ DropKey drop_key(const unsigned long long* state, uint32_t call, float p) {
  DropKey k;
  k.thresh = 0u, k.key = 0u, k.scale = 1.0f;
  if (state != nullptr && p > 0.0f) {
    const unsigned long long seed = state[0], step = state[1];
    const uint32_t a = mix32(static_cast<uint32_t>(seed) ^ 0x9e3779b9u) + mix32(static_cast<uint32_t>(seed >> 32) + 0x7f4a7c15u);
    const uint32_t b = mix32(static_cast<uint32_t>(step) * 0x85ebca6bu + static_cast<uint32_t>(step >> 32) + 0x165667b1u);
    k.key = mix32(a ^ b ^ mix32(call * 0xc2b2ae35u + 0x27d4eb2fu));
    k.thresh = static_cast<uint32_t>(p * 65536.0f + 0.5f);
    k.scale = 1.0f / (1.0f - p);
  }
  return k;
}
// 2 x 16 random bits for the element pair `pair` (elements 2 pair, 2 pair + 1 of the tensor's own numbering)
__device__ __forceinline__ uint32_t drop_bits(const DropKey& k, uint32_t pair) { return mix32(pair * 0x9e3779b1u + k.key); }
__device__ __forceinline__ bool drop_keep_lo(const DropKey& k, uint32_t bits) { return (bits & 0xffffu) >= k.thresh; }
__device__ __forceinline__ bool drop_keep_hi(const DropKey& k, uint32_t bits) { return (bits >> 16) >= k.thresh; }

// fire-and-forget fp32 x4 reduction into global memory (address 16-byte aligned)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every hot kernel is launched with the programmatic-stream-serialisation attribute
// (launch_pdl below; captured into the whole-step CUDA graph as a programmatic edge): its CTAs may become resident while
// the previous kernel of the stream is still draining, so launch latency, block scheduling and the on-chip prologue
// (barrier init, TMEM allocation, descriptor prefetch) overlap the predecessor's tail.  Contract, kept by every kernel
// launched this way:
//   * pdl_trigger() first (lets the NEXT kernel's CTAs be scheduled once all of ours have started -- they can never
//     displace our own pending CTAs);
//   * pdl_wait() executed by EVERY thread before its first global-memory access of any kind (reads of the producer's
//     output, and writes -- the predecessor may still be reading what we overwrite).  It returns when the prerequisite
//     grid has completed and its memory is visible; since every kernel in the chain waits, ordering is transitive.
//   Kernel parameters and __grid_constant__ tensor maps live in constant / parameter space and may be touched earlier.
// Launched WITHOUT the attribute (SMX_PDL=0, or by anyone else) both instructions are no-ops.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // host_common.cu: SMX_PDL != "0" (read once)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace smx
