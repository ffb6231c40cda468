// Flash-style attention backward for head_dim 64 on sm_100a (no atomics, deterministic):
//   1. delta[b,h,q] = rowsum(dO * O)
//   2. dKV kernel: one CTA per 128-row key/value tile, loops over query tiles
//        S^T = K.Q^T, dP^T = V.dO^T (TMEM) -> P^T, dS^T (bf16, smem) -> dV += P^T.dO, dK += dS^T.Q (TMEM)
//   3. dQ kernel: one CTA per 128-row query tile, loops over key tiles
//        S = Q.K^T, dP = dO.V^T -> dS -> dQ += dS.K
// P is recomputed from the forward LSE; scores never reach HBM.  The operands that
// are contracted over their row index (dO, Q, K as "B") are read MN-major straight
// from the TMA-loaded tiles, so nothing is transposed in memory.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

#include <string.h>

namespace smx {
namespace attn {

int make_head_map(CUtensorMap* m, const void* ptr, int t, int heads, int batch, long long row_stride,
                  long long batch_stride);

constexpr int D = 64;
constexpr int TILE = 128 * D * 2;   // 16 KiB
constexpr int SQ = 128 * 128 * 2;   // 32 KiB: a [128 x 128] bf16 tile (two K-blocks)
constexpr float kLog2e = 1.4426950408889634f;

struct BwdParams {
  const float* lse;
  const float* delta;
  const float* bias;
  float* dbias;      // [heads, tq, tk] fp32, accumulated with atomics over the batch (T5 relative position bias)
  float inv_scale;
  bf16 *dq, *dk, *dv;
  long long dq_row_stride, dq_batch_stride, dk_row_stride, dk_batch_stride, dv_row_stride, dv_batch_stride;
  int batch, heads, tq, tk, causal;
  float scale, scale_log2;
};

// ------------------------------------------------------------------ delta = rowsum(dO * O)
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta,
                                  long long o_rs, long long o_bs, long long do_rs, long long do_bs, int batch,
                                  int heads, int tq) {
  const long long n = (long long)batch * heads * tq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % tq);
    const int h = (int)((i / tq) % heads);
    const int b = (int)(i / ((long long)tq * heads));
    const uint4* po = reinterpret_cast<const uint4*>(o + b * o_bs + q * o_rs + h * D);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * do_bs + q * do_rs + h * D);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint4 a = po[k], c = pd[k];
      acc += bf16_lo(a.x) * bf16_lo(c.x) + bf16_hi(a.x) * bf16_hi(c.x) + bf16_lo(a.y) * bf16_lo(c.y) +
             bf16_hi(a.y) * bf16_hi(c.y) + bf16_lo(a.z) * bf16_lo(c.z) + bf16_hi(a.z) * bf16_hi(c.z) +
             bf16_lo(a.w) * bf16_lo(c.w) + bf16_hi(a.w) * bf16_hi(c.w);
    }
    delta[i] = acc;  // [b][h][q]
  }
}

// write 32 fp32 values as bf16 into row r of a K-major 128B-swizzled [128 x 128] tile, columns c32*32..+31
__device__ __forceinline__ void store_row_chunk(uint8_t* tile, int r, int c32, const float (&f)[32]) {
  uint8_t* blk = tile + (c32 >> 1) * (128 * 128) + r * 128;
  const int sw = r & 7;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4 u;
    u.x = pack_bf16x2(f[k * 8 + 0], f[k * 8 + 1]);
    u.y = pack_bf16x2(f[k * 8 + 2], f[k * 8 + 3]);
    u.z = pack_bf16x2(f[k * 8 + 4], f[k * 8 + 5]);
    u.w = pack_bf16x2(f[k * 8 + 6], f[k * 8 + 7]);
    const int chunk = ((c32 & 1) * 4 + k) ^ sw;
    *reinterpret_cast<uint4*>(blk + chunk * 16) = u;
  }
}

__device__ __forceinline__ void store_out_row(bf16* dst, uint32_t taddr, bool valid) {
#pragma unroll
  for (int c = 0; c < D / 32; ++c) {
    uint32_t v[32];
    tmem_ld_x32(taddr + c * 32, v);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        u.y = pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        u.z = pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
        u.w = pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
        *reinterpret_cast<uint4*>(dst + c * 32 + i) = u;
      }
    }
    __syncwarp();
  }
}

// A (K-major, [128 x 128] two K-blocks) x B (MN-major [128 rows(K) x 64]) -> D[128 x 64], 8 K-steps
__device__ __forceinline__ void mma_sq_times_tile(uint32_t d_tmem, uint32_t a_sq, uint32_t b_tile, uint32_t idesc,
                                                  bool accumulate) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint64_t ad = umma_smem_desc(a_sq + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024, kLayoutSW128);
    const uint64_t bd = umma_smem_desc(b_tile + kk * 2048, TILE, 1024, kLayoutSW128);
    umma_ss(d_tmem, ad, bd, idesc, (accumulate || kk > 0) ? 1u : 0u);
  }
}
// A (K-major [128 x 64]) x B (K-major [128 x 64]) -> D[128 x 128], 4 K-steps
__device__ __forceinline__ void mma_tile_times_tileT(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile, uint32_t idesc) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint64_t ad = umma_smem_desc(a_tile + kk * 32, 16, 1024, kLayoutSW128);
    const uint64_t bd = umma_smem_desc(b_tile + kk * 32, 16, 1024, kLayoutSW128);
    umma_ss(d_tmem, ad, bd, idesc, kk > 0 ? 1u : 0u);
  }
}

constexpr int BWD_THREADS = 320;  // producer warp, MMA warp, 8 compute warps

// =====================================================================================
// dK / dV
// =====================================================================================
namespace kv {
constexpr int OFF_K = 0, OFF_V = TILE, OFF_Q = 2 * TILE, OFF_DO = 4 * TILE, OFF_PT = 6 * TILE, OFF_DST = OFF_PT + SQ;
constexpr int OFF_STAT = OFF_DST + SQ;          // [2][2][128] floats
constexpr int OFF_BAR = OFF_STAT + 2 * 2 * 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int COL_ST = 0, COL_DPT = 128, COL_DV = 256, COL_DK = 320;
enum { B_KV = 0, B_QFULL = 1, B_QEMPTY = 3, B_STFULL = 5, B_STEMPTY = 6, B_PDSFULL = 7, B_PDSEMPTY = 8, B_DONE = 9,
       B_COUNT = 10 };
}  // namespace kv

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                    const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo,
                    const BwdParams p) {
  using namespace kv;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
  float* stat = reinterpret_cast<float*>(smem + OFF_STAT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int head = blockIdx.y, b = blockIdx.z;
  const int nq_tiles = (p.tq + 127) / 128;
  int i_start = 0;
  if (p.causal) {  // first query row that can see key kv0:  q >= kv0 - (tk - tq)
    int qmin = kv0 - (p.tk - p.tq);
    if (qmin < 0) qmin = 0;
    i_start = qmin / 128;
    if (i_start > nq_tiles) i_start = nq_tiles;
  }
  const int n_iter = nq_tiles - i_start;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(&bars[B_KV], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[B_QFULL + s], 1);
      mbar_init(&bars[B_QEMPTY + s], 1);
    }
    mbar_init(&bars[B_STFULL], 1);
    mbar_init(&bars[B_STEMPTY], 256);
    mbar_init(&bars[B_PDSFULL], 256);
    mbar_init(&bars[B_PDSEMPTY], 1);
    mbar_init(&bars[B_DONE], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars[B_KV], 2 * TILE);
      tma_load_4d(smem + OFF_K, &mk, &bars[B_KV], 0, kv0, head, b);
      tma_load_4d(smem + OFF_V, &mv, &bars[B_KV], 0, kv0, head, b);
      for (int it = 0; it < n_iter; ++it) {
        const int slot = it & 1;
        const uint32_t par = (it >> 1) & 1;
        const int qr = (i_start + it) * 128;
        mbar_wait(&bars[B_QEMPTY + slot], par ^ 1);
        mbar_expect_tx(&bars[B_QFULL + slot], 2 * TILE);
        tma_load_4d(smem + OFF_Q + slot * TILE, &mq, &bars[B_QFULL + slot], 0, qr, head, b);
        tma_load_4d(smem + OFF_DO + slot * TILE, &mdo, &bars[B_QFULL + slot], 0, qr, head, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);
      mbar_wait(&bars[B_KV], 0);
      for (int it = 0; it < n_iter; ++it) {
        const int slot = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&bars[B_QFULL + slot], par);
        if (it > 0) mbar_wait(&bars[B_STEMPTY], (it - 1) & 1);
        tc_fence_after_sync();
        mma_tile_times_tileT(tmem_base + COL_ST, sbase + OFF_K, sbase + OFF_Q + slot * TILE, idesc_s);
        mma_tile_times_tileT(tmem_base + COL_DPT, sbase + OFF_V, sbase + OFF_DO + slot * TILE, idesc_s);
        umma_commit(&bars[B_STFULL]);
        mbar_wait(&bars[B_PDSFULL], it & 1);
        tc_fence_after_sync();
        mma_sq_times_tile(tmem_base + COL_DV, sbase + OFF_PT, sbase + OFF_DO + slot * TILE, idesc_o, it > 0);
        mma_sq_times_tile(tmem_base + COL_DK, sbase + OFF_DST, sbase + OFF_Q + slot * TILE, idesc_o, it > 0);
        umma_commit(&bars[B_PDSEMPTY]);
        umma_commit(&bars[B_QEMPTY + slot]);
      }
      umma_commit(&bars[B_DONE]);
    }
  } else {
    const int cw = warp - 2;        // 0..7
    const int q = warp & 3;         // TMEM lane quarter
    const int half = cw >> 2;       // which 64 of the 128 columns
    const int r = q * 32 + lane;    // key row inside the tile
    const int kvi = kv0 + r;
    const int ct = threadIdx.x - 64;  // 0..255
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int it = 0; it < n_iter; ++it) {
      const int q0 = (i_start + it) * 128;
      float* st = stat + (it & 1) * 256;
      {  // stage lse (log2 domain) and delta for this query tile
        const int qi = q0 + (ct & 127);
        float val = 0.f;
        if (qi < p.tq) {
          const long long idx = ((long long)b * p.heads + head) * p.tq + qi;
          val = ct < 128 ? p.lse[idx] * kLog2e : p.delta[idx] * p.scale;
        }
        st[ct] = val;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&bars[B_STFULL], it & 1);
      tc_fence_after_sync();
      mbar_wait(&bars[B_PDSEMPTY], (it & 1) ^ 1);  // previous dV/dK MMAs finished reading P^T / dS^T
      // Out-of-range queries / keys need no masking here: their Q / dO / K / V rows are zero-filled by
      // TMA, so every product they enter vanishes as long as P and dS stay finite (they do).  Only the
      // causal diagonal and an additive bias need the per-element path.
      const bool lean = !p.causal && p.bias == nullptr;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c32 = half * 2 + cc;
        uint32_t sv[32], dv[32];
        tmem_ld_x32(t_lane + COL_ST + c32 * 32, sv);
        tmem_ld_x32(t_lane + COL_DPT + c32 * 32, dv);
        tmem_ld_wait();
        float pt[32], ds[32];
        if (lean) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(st + c32 * 32 + i);
            const float4 d4 = *reinterpret_cast<const float4*>(st + 128 + c32 * 32 + i);
            const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float e = ex2_approx(fmaf(__uint_as_float(sv[i + k]), p.scale_log2, -lv[k]));
              pt[i + k] = e;
              ds[i + k] = e * fmaf(__uint_as_float(dv[i + k]), p.scale, -dl[k]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c32 * 32 + i;
            const int qi = q0 + col;
            float s = __uint_as_float(sv[i]) * p.scale_log2;
            if (p.bias && qi < p.tq && kvi < p.tk) s += p.bias[((long long)head * p.tq + qi) * p.tk + kvi] * kLog2e;
            const bool ok = (qi < p.tq) && (kvi < p.tk) && (!p.causal || kvi <= qi + (p.tk - p.tq));
            const float e = ok ? ex2_approx(s - st[col]) : 0.f;
            pt[i] = e;
            ds[i] = e * fmaf(__uint_as_float(dv[i]), p.scale, -st[128 + col]);
          }
        }
        store_row_chunk(smem + OFF_PT, r, c32, pt);
        store_row_chunk(smem + OFF_DST, r, c32, ds);
      }
      tc_fence_before_sync();
      mbar_arrive(&bars[B_STEMPTY]);
      fence_proxy_async_smem();
      mbar_arrive(&bars[B_PDSFULL]);
    }
    mbar_wait(&bars[B_DONE], 0);
    tc_fence_after_sync();
    const bool valid = kvi < p.tk;
    if (half == 0) {
      bf16* dst = p.dv + (long long)b * p.dv_batch_stride + (long long)kvi * p.dv_row_stride + head * D;
      if (n_iter == 0) {
        if (valid)
          for (int i = 0; i < D; i += 8) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
      } else {
        store_out_row(dst, t_lane + COL_DV, valid);
      }
    } else {
      bf16* dst = p.dk + (long long)b * p.dk_batch_stride + (long long)kvi * p.dk_row_stride + head * D;
      if (n_iter == 0) {
        if (valid)
          for (int i = 0; i < D; i += 8) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
      } else {
        store_out_row(dst, t_lane + COL_DK, valid);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================
// dQ
// =====================================================================================
namespace dq {
constexpr int OFF_Q = 0, OFF_DO = TILE, OFF_K = 2 * TILE, OFF_V = 4 * TILE, OFF_DS = 6 * TILE;
constexpr int OFF_BAR = OFF_DS + SQ;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int COL_S = 0, COL_DP = 128, COL_DQ = 256;
enum { B_Q = 0, B_KFULL = 1, B_KEMPTY = 3, B_SFULL = 5, B_SEMPTY = 6, B_DSFULL = 7, B_DSEMPTY = 8, B_DONE = 9,
       B_COUNT = 10 };
}  // namespace dq

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                   const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo,
                   const BwdParams p) {
  using namespace dq;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y, b = blockIdx.z;
  int n_iter = (p.tk + 127) / 128;
  if (p.causal) {
    const int last_col = q0 + 127 + (p.tk - p.tq);
    int nc = last_col / 128 + 1;
    if (nc < 1) nc = 1;
    if (nc < n_iter) n_iter = nc;
  }

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(&bars[B_Q], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[B_KFULL + s], 1);
      mbar_init(&bars[B_KEMPTY + s], 1);
    }
    mbar_init(&bars[B_SFULL], 1);
    mbar_init(&bars[B_SEMPTY], 256);
    mbar_init(&bars[B_DSFULL], 256);
    mbar_init(&bars[B_DSEMPTY], 1);
    mbar_init(&bars[B_DONE], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars[B_Q], 2 * TILE);
      tma_load_4d(smem + OFF_Q, &mq, &bars[B_Q], 0, q0, head, b);
      tma_load_4d(smem + OFF_DO, &mdo, &bars[B_Q], 0, q0, head, b);
      for (int it = 0; it < n_iter; ++it) {
        const int slot = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&bars[B_KEMPTY + slot], par ^ 1);
        mbar_expect_tx(&bars[B_KFULL + slot], 2 * TILE);
        tma_load_4d(smem + OFF_K + slot * TILE, &mk, &bars[B_KFULL + slot], 0, it * 128, head, b);
        tma_load_4d(smem + OFF_V + slot * TILE, &mv, &bars[B_KFULL + slot], 0, it * 128, head, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);
      mbar_wait(&bars[B_Q], 0);
      for (int it = 0; it < n_iter; ++it) {
        const int slot = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&bars[B_KFULL + slot], par);
        if (it > 0) mbar_wait(&bars[B_SEMPTY], (it - 1) & 1);
        tc_fence_after_sync();
        mma_tile_times_tileT(tmem_base + COL_S, sbase + OFF_Q, sbase + OFF_K + slot * TILE, idesc_s);
        mma_tile_times_tileT(tmem_base + COL_DP, sbase + OFF_DO, sbase + OFF_V + slot * TILE, idesc_s);
        umma_commit(&bars[B_SFULL]);
        mbar_wait(&bars[B_DSFULL], it & 1);
        tc_fence_after_sync();
        mma_sq_times_tile(tmem_base + COL_DQ, sbase + OFF_DS, sbase + OFF_K + slot * TILE, idesc_o, it > 0);
        umma_commit(&bars[B_DSEMPTY]);
        umma_commit(&bars[B_KEMPTY + slot]);
      }
      umma_commit(&bars[B_DONE]);
    }
  } else {
    const int cw = warp - 2;
    const int q = warp & 3;
    const int half = cw >> 2;
    const int r = q * 32 + lane;
    const int qi = q0 + r;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float lse2 = 0.f, delta = 0.f;
    if (qi < p.tq) {
      const long long idx = ((long long)b * p.heads + head) * p.tq + qi;
      lse2 = p.lse[idx] * kLog2e;
      delta = p.delta[idx] * p.scale;
    }
    const bool lean = !p.causal && p.bias == nullptr;
    const int causal_lim = p.causal ? qi + (p.tk - p.tq) : 0x7fffffff;
    for (int it = 0; it < n_iter; ++it) {
      const int k0 = it * 128;
      mbar_wait(&bars[B_SFULL], it & 1);
      tc_fence_after_sync();
      mbar_wait(&bars[B_DSEMPTY], (it & 1) ^ 1);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c32 = half * 2 + cc;
        uint32_t sv[32], dv[32];
        tmem_ld_x32(t_lane + COL_S + c32 * 32, sv);
        tmem_ld_x32(t_lane + COL_DP + c32 * 32, dv);
        tmem_ld_wait();
        float ds[32];
        if (lean) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(sv[i]), p.scale_log2, -lse2));
            ds[i] = e * fmaf(__uint_as_float(dv[i]), p.scale, -delta);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kvi = k0 + c32 * 32 + i;
            float s = __uint_as_float(sv[i]) * p.scale_log2;
            if (p.bias && qi < p.tq && kvi < p.tk) s += p.bias[((long long)head * p.tq + qi) * p.tk + kvi] * kLog2e;
            const bool ok = (qi < p.tq) && (kvi < p.tk) && (kvi <= causal_lim);
            const float e = ok ? ex2_approx(s - lse2) : 0.f;
            ds[i] = e * fmaf(__uint_as_float(dv[i]), p.scale, -delta);
            if (p.dbias && ok) atomicAdd(p.dbias + ((long long)head * p.tq + qi) * p.tk + kvi, ds[i] * p.inv_scale);
          }
        }
        store_row_chunk(smem + OFF_DS, r, c32, ds);
      }
      tc_fence_before_sync();
      mbar_arrive(&bars[B_SEMPTY]);
      fence_proxy_async_smem();
      mbar_arrive(&bars[B_DSFULL]);
    }
    mbar_wait(&bars[B_DONE], 0);
    tc_fence_after_sync();
    if (half == 0) {
      bf16* dst = p.dq + (long long)b * p.dq_batch_stride + (long long)qi * p.dq_row_stride + head * D;
      store_out_row(dst, t_lane + COL_DQ, qi < p.tq);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace attn
}  // namespace smx

extern "C" int smx_attn_bwd(const SmxAttn* a, void* stream) {
  using namespace smx;
  using namespace smx::attn;
  SMX_REQUIRE(a && a->q && a->k && a->v && a->o && a->d_o && a->dq && a->dk && a->dv && a->lse && a->delta,
              "attn_bwd: null pointer");
  SMX_REQUIRE(a->dbias == nullptr || a->bias != nullptr, "attn_bwd: dbias needs the additive bias it differentiates");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mq, mk, mv, mdo;
  if (make_head_map(&mq, a->q, a->tq, a->heads, a->batch, a->q_row_stride, a->q_batch_stride)) return -1;
  if (make_head_map(&mk, a->k, a->tk, a->heads, a->batch, a->k_row_stride, a->k_batch_stride)) return -1;
  if (make_head_map(&mv, a->v, a->tk, a->heads, a->batch, a->v_row_stride, a->v_batch_stride)) return -1;
  if (make_head_map(&mdo, a->d_o, a->tq, a->heads, a->batch, a->do_row_stride, a->do_batch_stride)) return -1;
  {
    const long long n = (long long)a->batch * a->heads * a->tq;
    long long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    attn_delta_kernel<<<(int)g, 256, 0, st>>>((const bf16*)a->o, (const bf16*)a->d_o, a->delta, a->o_row_stride,
                                             a->o_batch_stride, a->do_row_stride, a->do_batch_stride, a->batch,
                                             a->heads, a->tq);
    SMX_CHECK_CUDA(cudaGetLastError());
  }
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.lse = a->lse, p.delta = a->delta, p.bias = a->bias;
  p.dbias = a->dbias, p.inv_scale = 1.0f / a->scale;
  p.dq = (bf16*)a->dq, p.dk = (bf16*)a->dk, p.dv = (bf16*)a->dv;
  p.dq_row_stride = a->dq_row_stride, p.dq_batch_stride = a->dq_batch_stride;
  p.dk_row_stride = a->dk_row_stride, p.dk_batch_stride = a->dk_batch_stride;
  p.dv_row_stride = a->dv_row_stride, p.dv_batch_stride = a->dv_batch_stride;
  p.batch = a->batch, p.heads = a->heads, p.tq = a->tq, p.tk = a->tk, p.causal = a->causal;
  p.scale = a->scale;
  p.scale_log2 = a->scale * kLog2e;
  static bool attr_set = false;
  if (!attr_set) {
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kv::SMEM_BYTES));
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::SMEM_BYTES));
    attr_set = true;
  }
  dim3 gkv((a->tk + 127) / 128, a->heads, a->batch);
  attn_bwd_dkv_kernel<<<gkv, BWD_THREADS, kv::SMEM_BYTES, st>>>(mq, mk, mv, mdo, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  dim3 gq((a->tq + 127) / 128, a->heads, a->batch);
  attn_bwd_dq_kernel<<<gq, BWD_THREADS, dq::SMEM_BYTES, st>>>(mq, mk, mv, mdo, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
