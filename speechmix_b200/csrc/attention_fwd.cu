// Flash-style attention forward for head_dim 64 on sm_100a.
//
// One CTA = one 128-row query tile of one (batch, head); two CTAs co-reside per SM.
//   warp 0      TMA producer: Q once, K/V tiles double-buffered (128-byte swizzle)
//   warp 1      tcgen05.mma issuer: S = Q.K^T  (128x128x64, TMEM cols 0..127)
//                                   PV = P.V   (128x64x128, TMEM cols 128..255, double-buffered)
//               issue order S_0, [S_{j+1}, PV_j]...: the next score tile goes out as soon as the softmax warps
//               have released the score buffer, BEFORE the P.V product of the current tile
//   warps 2..5  softmax: one query row per thread; S row read from TMEM (loads one chunk ahead of the
//               arithmetic), online max / sum in fp32 (packed fp32 pairs), P written to smem as bf16 in the
//               UMMA K-major swizzled layout, running O kept in registers and rescaled as PV tiles arrive.
// Scores and probabilities never reach HBM; LSE is written for the backward pass.  Optional per-sample key
// counts (kv_len): tiles past the count are neither loaded nor computed, the tile it cuts is masked per chunk.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

#include <stdlib.h>
#include <string.h>

namespace smx {
namespace attn {

constexpr int BQ = 128, BKV = 128, D = 64;
constexpr int TILE_BYTES = 128 * D * 2;       // 16 KiB: a [128 x 64] bf16 tile
constexpr int P_BYTES = BQ * BKV * 2;         // 32 KiB
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;     // 2 slots
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES; // 2 slots
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;      // 114688
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_PV = 128;

struct FwdParams {
  bf16* o;
  float* lse;
  const float* bias;
  long long o_row_stride, o_batch_stride;
  int batch, heads, tq, tk, causal;
  float scale_log2;  // scale * log2(e)
  const int* kv_len; // optional [batch]: keys at or past kv_len[b] are masked (never with causal)
  const unsigned long long* drop_state;   // dropout on the probabilities (general path only): {seed, step}, call, p
  uint32_t drop_call;
  float drop_p;
};

// barrier indices
enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 3, B_VFULL = 5, B_VEMPTY = 7, B_SFULL = 9, B_SEMPTY = 10, B_PFULL = 11,
       B_PVFULL = 12, B_PVEMPTY = 14, B_COUNT = 16 };

__device__ __forceinline__ int num_kv_tiles(const FwdParams& p, int q0, int tk) {
  int n = (tk + BKV - 1) / BKV;
  if (p.causal) {
    const int last_col = q0 + BQ - 1 + (p.tk - p.tq);  // last visible key for the last row of the tile
    int nc = last_col / BKV + 1;
    if (nc < 1) nc = 1;
    if (nc < n) n = nc;
  }
  return n;
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tq_map, const __grid_constant__ CUtensorMap tk_map,
                const __grid_constant__ CUtensorMap tv_map, const FwdParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y, b = blockIdx.z;
  // effective key count of this sample: tiles past it are never loaded, the tile it cuts takes the "cut" path
  int tk = p.tk;
  if (p.kv_len) {
    const int l = p.kv_len[b];
    tk = l < 1 ? 1 : (l < p.tk ? l : p.tk);
  }
  const int n_tiles = num_kv_tiles(p, q0, tk);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("smx attn_fwd: dynamic smem not 1024-aligned\n");
      __trap();
    }
    mbar_init(&bars[B_QFULL], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_KFULL + i], 1);
      mbar_init(&bars[B_KEMPTY + i], 1);
      mbar_init(&bars[B_VFULL + i], 1);
      mbar_init(&bars[B_VEMPTY + i], 1);
      mbar_init(&bars[B_PVFULL + i], 1);
      mbar_init(&bars[B_PVEMPTY + i], 128);
    }
    mbar_init(&bars[B_SFULL], 1);
    mbar_init(&bars[B_SEMPTY], 128);
    mbar_init(&bars[B_PFULL], 128);
    fence_barrier_init();
    tma_prefetch_desc(&tq_map);
    tma_prefetch_desc(&tk_map);
    tma_prefetch_desc(&tv_map);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars[B_QFULL], TILE_BYTES);
      tma_load_4d(smem + OFF_Q, &tq_map, &bars[B_QFULL], 0, q0, head, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int slot = j & 1;
        const uint32_t par = (j >> 1) & 1;
        mbar_wait(&bars[B_KEMPTY + slot], par ^ 1);
        mbar_expect_tx(&bars[B_KFULL + slot], TILE_BYTES);
        tma_load_4d(smem + OFF_K + slot * TILE_BYTES, &tk_map, &bars[B_KFULL + slot], 0, j * BKV, head, b);
        mbar_wait(&bars[B_VEMPTY + slot], par ^ 1);
        mbar_expect_tx(&bars[B_VFULL + slot], TILE_BYTES);
        tma_load_4d(smem + OFF_V + slot * TILE_BYTES, &tv_map, &bars[B_VFULL + slot], 0, j * BKV, head, b);
      }
    }
  } else if (warp == 1) {
    // The WHOLE warp runs the loop (warp-uniform waits and address arithmetic -> uniform-register operands); one elected
    // lane issues.  Under `if (lane == 0)` every MMA cost ~80 clk of issue against 46-64 clk of tensor-pipe time.
    constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, false, true);
    const uint32_t sbase = smem_u32(smem);
    mbar_wait(&bars[B_QFULL], 0);
    // S_{j+1} is issued BEFORE PV_j: the softmax warps get the next score tile after one 128x128x64 MMA instead of
    // waiting behind the P.V product as well, and PV_j runs on the tensor pipe while they take the row maxima.
    auto issue_s = [&](int j) {
      const int slot = j & 1;
      mbar_wait(&bars[B_KFULL + slot], (j >> 1) & 1);
      if (j > 0) mbar_wait(&bars[B_SEMPTY], (j - 1) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t ad = umma_smem_desc(sbase + OFF_Q + kk * 32, 16, 1024, kLayoutSW128);
          const uint64_t bd = umma_smem_desc(sbase + OFF_K + slot * TILE_BYTES + kk * 32, 16, 1024, kLayoutSW128);
          umma_ss(tmem_base + COL_S, ad, bd, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&bars[B_SFULL]);
        umma_commit(&bars[B_KEMPTY + slot]);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int slot = j & 1;
      const uint32_t par = (j >> 1) & 1;
      // ---- PV_j = P_j . V_j needs P_j; by then the softmax warps have also released the score buffer
      mbar_wait(&bars[B_PFULL], j & 1);
      if (j + 1 < n_tiles) issue_s(j + 1);
      mbar_wait(&bars[B_VFULL + slot], par);
      mbar_wait(&bars[B_PVEMPTY + (j & 1)], par ^ 1);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint32_t pa = sbase + OFF_P + (kk >> 2) * (BQ * 128) + (kk & 3) * 32;
          const uint64_t ad = umma_smem_desc(pa, 16, 1024, kLayoutSW128);
          const uint64_t bd = umma_smem_desc(sbase + OFF_V + slot * TILE_BYTES + kk * 2048, TILE_BYTES, 1024, kLayoutSW128);
          umma_ss(tmem_base + COL_PV + (j & 1) * D, ad, bd, idesc_pv, kk > 0 ? 1u : 0u);
        }
        umma_commit(&bars[B_PVFULL + (j & 1)]);
        umma_commit(&bars[B_VEMPTY + slot]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;            // row inside the tile == TMEM lane
    const int row = q0 + r;                 // query index
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int causal_lim = p.causal ? row + (p.tk - p.tq) : 0x7fffffff;
    const float* bias_row = p.bias ? p.bias + ((long long)head * p.tq + (row < p.tq ? row : 0)) * p.tk : nullptr;
    uint8_t* p_row = smem + OFF_P + r * 128;
    const int sw = r & 7;

    float m = -INFINITY, l = 0.f, alpha_pending = 0.f;
    const DropKey dkey = drop_key(p.drop_state, p.drop_call, p.drop_p);   // thresh 0: every element is kept
    const uint32_t drop_row = static_cast<uint32_t>(((long long)b * p.heads + head) * p.tq + row) * static_cast<uint32_t>((p.tk + 1) >> 1);
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&bars[B_SFULL], j & 1);
      tc_fence_after_sync();
      const int c_base = j * BKV;
      // Three flavours of tile.  "lean": no bias, not cut by the causal diagonal -- the common case; its chunks of 32
      // columns are either full, cut by the key length (one chunk of the last tile: compile-time column index against a
      // warp-uniform count) or empty (no exponentials at all).  Everything else takes the general per-element path.
      // Lean arithmetic per score: a third of an FMNMX3 in pass 1; half an FFMA2 + MUFU.EX2 + half an FADD2 + half a
      // pack in pass 2.  TMEM loads are issued one chunk ahead of the arithmetic (tcgen05.wait::ld is the only stall).
      const bool plain = (bias_row == nullptr) && dkey.thresh == 0u && (!p.causal || c_base + BKV - 1 <= q0 + (p.tk - p.tq));
      const int n_valid = tk - c_base;   // >= 1; columns of this tile inside the key length (may exceed BKV)
      const bool lean = plain && n_valid >= BKV;   // full tile
      const bool cut = plain && n_valid < BKV;     // last tile of a ragged key length
      float tmax = -INFINITY;
      if (lean) {
        uint32_t va[32], vb[32];
        tmem_ld_x32(t_lane + COL_S, va);
        tmem_ld_x32(t_lane + COL_S + 32, vb);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            tmax = fmax3(tmax, __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
            tmax = fmax3(tmax, __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
          }
          if (h == 0) {
            tmem_ld_x32(t_lane + COL_S + 64, va);
            tmem_ld_x32(t_lane + COL_S + 96, vb);
          }
        }
        tmax *= p.scale_log2;   // scale > 0
      } else if (cut) {
#pragma unroll 1
        for (int c = 0; c * 32 < n_valid; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_lane + COL_S + c * 32, v);
          tmem_ld_wait();
          const int nv = n_valid - c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, i < nv ? __uint_as_float(v[i]) : -INFINITY);
        }
        tmax *= p.scale_log2;
      } else {
#pragma unroll 1
        for (int c = 0; c < BKV / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_lane + COL_S + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c_base + c * 32 + i;
            float s = __uint_as_float(v[i]) * p.scale_log2;
            if (bias_row && col < tk) s += bias_row[col] * 1.4426950408889634f;
            s = (col < tk && col <= causal_lim) ? s : -INFINITY;
            tmax = fmaxf(tmax, s);
          }
        }
      }
      const float m_new = fmaxf(m, tmax);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_approx(m - m_use);
      // ---- pass 2: probabilities -> smem (bf16, swizzled K-major), row sum
      float lt = 0.f;
      if (lean) {
        const f32x2 sc2 = f2_rep(p.scale_log2), nm2 = f2_rep(-m_use);
        f32x2 lt2 = f2_rep(0.f);
        uint32_t va[16], vb[16];   // 16-column sub-chunks, one in flight while the other is consumed
        tmem_ld_x16(t_lane + COL_S, va);
        if (j > 0) {  // PV_{j-1} finished: its result is readable and the P buffer is free again
          mbar_wait(&bars[B_PVFULL + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
          tc_fence_after_sync();
        }
#pragma unroll
        for (int c = 0; c < BKV / 16; ++c) {
          uint32_t (&v)[16] = (c & 1) ? vb : va;
          tmem_ld_wait();
          if (c + 1 < BKV / 16) tmem_ld_x16(t_lane + COL_S + (c + 1) * 16, (c & 1) ? va : vb);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float e0, e1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), e0, e1);
            e0 = ex2_approx(e0), e1 = ex2_approx(e1);
            lt2 = f2_add(lt2, f2_pack(e0, e1));
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
          // 16 columns = 2 chunks of 16 bytes inside K-block (c>>2), chunk index (c&3)*2 + k
          uint8_t* blk = p_row + (c >> 2) * (BQ * 128);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int chunk = ((c & 3) * 2 + k) ^ sw;
            *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          }
        }
        float l0, l1;
        f2_unpack(lt2, l0, l1);
        lt = l0 + l1;
      } else if (cut) {
        if (j > 0) {
          mbar_wait(&bars[B_PVFULL + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
          tc_fence_after_sync();
        }
#pragma unroll 1
        for (int c = 0; c < BKV / 32; ++c) {
          const int nv = n_valid - c * 32;   // warp-uniform; <= 0: the chunk is all padding, no exponentials
          uint32_t pk[16];
          if (nv > 0) {
            uint32_t v[32];
            tmem_ld_x32(t_lane + COL_S + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float e0 = fmaf(__uint_as_float(v[i]), p.scale_log2, -m_use);
              float e1 = fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_use);
              e0 = i < nv ? ex2_approx(e0) : 0.f, e1 = i + 1 < nv ? ex2_approx(e1) : 0.f;
              lt += e0 + e1;
              pk[i >> 1] = pack_bf16x2(e0, e1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          }
          uint8_t* blk = p_row + (c >> 1) * (BQ * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int chunk = ((c & 1) * 4 + k) ^ sw;
            *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          }
        }
      } else {
        if (j > 0) {
          mbar_wait(&bars[B_PVFULL + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
          tc_fence_after_sync();
        }
#pragma unroll 1
        for (int c = 0; c < BKV / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_lane + COL_S + c * 32, v);
          tmem_ld_wait();
          float pr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c_base + c * 32 + i;
            float s = __uint_as_float(v[i]) * p.scale_log2;
            if (bias_row && col < tk) s += bias_row[col] * 1.4426950408889634f;
            const float e = (col < tk && col <= causal_lim) ? ex2_approx(s - m_use) : 0.f;
            lt += e;
            // dropout: the row sum stays that of the full probabilities, P.V sees the kept ones (scaled at the end)
            const uint32_t bits = drop_bits(dkey, drop_row + static_cast<uint32_t>(col >> 1));
            pr[i] = ((col & 1) ? drop_keep_hi(dkey, bits) : drop_keep_lo(dkey, bits)) ? e : 0.f;
          }
          uint8_t* blk = p_row + (c >> 1) * (BQ * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 u;
            u.x = pack_bf16x2(pr[k * 8 + 0], pr[k * 8 + 1]);
            u.y = pack_bf16x2(pr[k * 8 + 2], pr[k * 8 + 3]);
            u.z = pack_bf16x2(pr[k * 8 + 4], pr[k * 8 + 5]);
            u.w = pack_bf16x2(pr[k * 8 + 6], pr[k * 8 + 7]);
            const int chunk = ((c & 1) * 4 + k) ^ sw;
            *reinterpret_cast<uint4*>(blk + chunk * 16) = u;
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars[B_SEMPTY]);
      fence_proxy_async_smem();
      mbar_arrive(&bars[B_PFULL]);
      l = l * alpha + lt;
      m = m_new;
      if (j > 0) {
        const uint32_t ta = t_lane + COL_PV + ((j - 1) & 1) * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(ta + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            f2_unpack(f2_fma(f2_pack(o[c * 32 + i], o[c * 32 + i + 1]), f2_rep(alpha_pending),
                             f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1]))), o[c * 32 + i], o[c * 32 + i + 1]);
        }
        tc_fence_before_sync();
        mbar_arrive(&bars[B_PVEMPTY + ((j - 1) & 1)]);
      }
      alpha_pending = alpha;
    }
    {  // last PV tile
      const int j = n_tiles - 1;
      mbar_wait(&bars[B_PVFULL + (j & 1)], (j >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t ta = t_lane + COL_PV + (j & 1) * D;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        tmem_ld_x32(ta + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2)
            f2_unpack(f2_fma(f2_pack(o[c * 32 + i], o[c * 32 + i + 1]), f2_rep(alpha_pending),
                             f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1]))), o[c * 32 + i], o[c * 32 + i + 1]);
      }
    }
    if (row < p.tq) {
      const float inv = l > 0.f ? dkey.scale / l : 0.f;
      bf16* op = p.o + (long long)b * p.o_batch_stride + (long long)row * p.o_row_stride + head * D;
#pragma unroll
      for (int i = 0; i < D; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv);
        u.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        u.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
        u.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(op + i) = u;
      }
      if (p.lse) p.lse[((long long)b * p.heads + head) * p.tq + row] = (m + log2f(l)) * 0.6931471805599453f;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int make_head_map(CUtensorMap* m, const void* ptr, int t, int heads, int batch, long long row_stride,
                  long long batch_stride) {
  const uint64_t dims[4] = {(uint64_t)D, (uint64_t)t, (uint64_t)heads, (uint64_t)batch};
  const uint64_t str[3] = {(uint64_t)row_stride, (uint64_t)D, (uint64_t)(batch > 1 ? batch_stride : row_stride * t)};
  const uint32_t box[4] = {64, 128, 1, 1};
  return encode_tmap_bf16(m, ptr, 4, dims, str, box, true);
}

int make_head_map_rows(CUtensorMap* m, const void* ptr, int t, int heads, int batch, long long row_stride,
                       long long batch_stride, int box_rows) {
  const uint64_t dims[4] = {(uint64_t)D, (uint64_t)t, (uint64_t)heads, (uint64_t)batch};
  const uint64_t str[3] = {(uint64_t)row_stride, (uint64_t)D, (uint64_t)(batch > 1 ? batch_stride : row_stride * t)};
  const uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
  return encode_tmap_bf16(m, ptr, 4, dims, str, box, true);
}

int launch_fwd2(const SmxAttn* a, cudaStream_t stream);   // attention_fwd2.cu: two query tiles per CTA, P / O in TMEM

}  // namespace attn
}  // namespace smx

extern "C" int smx_attn_fwd(const SmxAttn* a, void* stream) {
  using namespace smx;
  using namespace smx::attn;
  SMX_REQUIRE(a && a->q && a->k && a->v && a->o, "attn_fwd: null pointer");
  SMX_REQUIRE(a->batch > 0 && a->heads > 0 && a->tq > 0 && a->tk > 0, "attn_fwd: empty problem");
  SMX_REQUIRE(a->o_row_stride % 8 == 0 && a->o_batch_stride % 8 == 0, "attn_fwd: o strides must be multiples of 8");
  SMX_REQUIRE(a->kv_len == nullptr || !a->causal, "attn_fwd: kv_len is not combined with causal masking");
  SMX_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "attn_fwd: dropout p outside [0, 1)");
  SMX_REQUIRE(a->dropout_state == nullptr || a->dropout_p == 0.0f ||
                  (long long)a->batch * a->heads * a->tq * ((a->tk + 1) / 2) < (1ll << 32),
              "attn_fwd: problem too large for the 32-bit dropout pair index");
  // plain attention over more than one query tile (speech / text encoder self-attention): the ping-pong kernel;
  // causal, biased (T5) and short-query (decoder, incremental decoding) problems stay on the general kernel
  static const bool force_v1 = getenv("SMX_ATTN_FWD_V1") != nullptr;
  if (!force_v1 && !a->causal && a->bias == nullptr && a->tq > BQ) return launch_fwd2(a, (cudaStream_t)stream);
  CUtensorMap mq, mk, mv;
  if (make_head_map(&mq, a->q, a->tq, a->heads, a->batch, a->q_row_stride, a->q_batch_stride)) return -1;
  if (make_head_map(&mk, a->k, a->tk, a->heads, a->batch, a->k_row_stride, a->k_batch_stride)) return -1;
  if (make_head_map(&mv, a->v, a->tk, a->heads, a->batch, a->v_row_stride, a->v_batch_stride)) return -1;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.o = reinterpret_cast<bf16*>(a->o);
  p.lse = a->lse;
  p.bias = a->bias;
  p.o_row_stride = a->o_row_stride;
  p.o_batch_stride = a->o_batch_stride;
  p.batch = a->batch, p.heads = a->heads, p.tq = a->tq, p.tk = a->tk, p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.kv_len = a->kv_len;
  p.drop_state = reinterpret_cast<const unsigned long long*>(a->dropout_state);
  p.drop_call = a->dropout_call;
  p.drop_p = a->dropout_state ? a->dropout_p : 0.0f;
  static bool attr_set = false;
  if (!attr_set) {
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a->tq + BQ - 1) / BQ, a->heads, a->batch);
  launch_pdl(attn_fwd_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, (cudaStream_t)stream, mq, mk, mv, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
