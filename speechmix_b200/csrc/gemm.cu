// Persistent, warp-specialised bf16 GEMM for sm_100a on CTA PAIRS (tcgen05 cta_group::2):
// a cluster of two CTAs (two SMs of one TPC) owns a 256 x 256 output tile.  Each CTA loads its 128
// rows of A and HALF of B (128 of the 256 columns) per 64-deep k-block, the leader CTA's single MMA
// thread issues 256x256x16 MMAs that read both CTAs' shared memory, and each CTA keeps its 128 rows
// of the accumulator in its own TMEM.  Halving B per SM is what lifts the shared-memory bandwidth
// bound of the single-CTA 128x256 tile (TMA fill + MMA operand reads ~192 B/clk vs 128 B/clk).
//   warp 0     : TMA producer (lanes issue the boxes of a stage in parallel) -> 6-stage smem ring
//   warp 1     : tcgen05.mma issuer (leader CTA only), accumulators in TMEM (2 x 256 columns,
//                so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..9 : epilogue, TMEM -> registers -> fused bias/activation/residual -> swizzled smem -> TMA store
// 128-byte swizzled operands, K-major or MN-major via the shared-memory descriptors (no transpose
// copies for dgrad / wgrad).
//
// Modes (see include/speechmix_sm100.h): NT forward (+ implicit-GEMM conv taps),
// NN data gradient, TN weight gradient (fp32 out, split-K with atomics).
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

#include <string.h>

namespace smx {

namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64;   // per CTA: 128 rows; the pair covers PAIR_M = 256 rows
constexpr int PAIR_M = 2 * BM;
constexpr int BN_HALF = BN / 2;              // columns of B each CTA of the pair loads
// Ring depth / staging banks are compile-time knobs.  Measured on B200 (gpurun_out/r01s_gemm_*.log, tools/probe_gemm.py):
// 5 stages + two staging tiles per epilogue warp (bulk store of one tile overlapping the fill of the other) is
// within noise of 6 stages + one tile on the K = 768 shapes and no better on the step, so the deeper ring stays.
#ifndef SMX_GEMM_STAGES
#define SMX_GEMM_STAGES 6
#endif
#ifndef SMX_GEMM_STG_BANKS
#define SMX_GEMM_STG_BANKS 1
#endif
constexpr int STAGES = SMX_GEMM_STAGES;
constexpr int STG_BANKS = SMX_GEMM_STG_BANKS;  // staging tiles per epilogue warp
constexpr int A_BYTES = BM * BK * 2;       // 16 KiB
constexpr int B_BYTES = BN_HALF * BK * 2;  // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int PANEL_BYTES = 64 * BK * 2;  // one 64(MN) x 64(K) MN-major panel, 8 KiB
constexpr int BAR_BYTES = 256;
constexpr int STG_BYTES = 32 * 128;  // per epilogue warp: 32 rows x 64 bf16, 128B-swizzled, TMA-store staging
constexpr int OFF_STG = STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_STG + 8 * STG_BANKS * STG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;
constexpr int NUM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;

struct Params {
  int mode, out_f32;
  long long m, n;
  int batches;
  int kblocks, kb_per_batch;
  int m_tiles_per_batch, m_tiles, n_tiles, split_k, n_fastest;
  int nseg, seg_len;
  int a_row_off[SMX_MAX_SEG], a_col_off[SMX_MAX_SEG];
  int b_row_off[SMX_MAX_SEG], b_col_off[SMX_MAX_SEG];
  int b_inner_oob;
  void* c;
  long long c_row_stride, c_batch_stride;
  int act, atomic;
  float alpha;
  const float* bias;
  const bf16* residual;
  long long res_row_stride, res_batch_stride;
  bf16* aux_out;
  const bf16* aux_in;
  // special epilogues (LM head): 0 = regular, 1 = per-tile softmax statistics, 2 = dlogits
  int epi, accumulate_f32;
  float4* lm_partial;        // [rows][n_tiles] {max, sumexp, best, best_idx}
  float* lm_label_logit;     // [rows]
  const long long* lm_labels;
  const float* lm_lse;       // [rows]
  const float* lm_coef;      // [rows]
  long long lm_label_off;    // vocab index of column 0 of this launch
  int tma_store;             // bf16 C (and aux_out) leave through smem staging + TMA stores
  int tma_in;                // 1: aux_in, 2: residual arrives through a TMA load into the staging tile
};

struct TileCoord {
  int m_blk, n_blk, split, kb_begin, kb_end;
};

__device__ __forceinline__ TileCoord decode_tile(const Params& p, int tile) {
  TileCoord t;
  if (p.n_fastest) {
    // the n-tiles of one row block run concurrently: the (large) A operand is fetched from DRAM once
    // and re-served from L2, the (small) weight operand stays L2-resident across waves
    t.n_blk = tile % p.n_tiles;
    const int rest = tile / p.n_tiles;
    t.m_blk = rest % p.m_tiles;
    t.split = rest / p.m_tiles;
  } else {
    t.m_blk = tile % p.m_tiles;
    const int rest = tile / p.m_tiles;
    t.n_blk = rest % p.n_tiles;
    t.split = rest / p.n_tiles;
  }
  t.kb_begin = (int)(((long long)p.kblocks * t.split) / p.split_k);
  t.kb_end = (int)(((long long)p.kblocks * (t.split + 1)) / p.split_k);
  return t;
}

// ---------------------------------------------------------------------------------------------
// Epilogue helpers.  The common case (a full 32-column chunk with 16-byte aligned rows) is a compact
// straight-line path; ragged / unaligned chunks go through a small out-of-line scalar routine so the
// hot loop stays inside the instruction cache (an earlier fully-unrolled version was I$-bound).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x), f[1] = bf16_hi(u.x), f[2] = bf16_lo(u.y), f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z), f[5] = bf16_hi(u.z), f[6] = bf16_lo(u.w), f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]), u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]), u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// v * act'(x) for the three backward epilogue flavours (x = stored pre-activation, or the stored derivative)
__device__ __forceinline__ float dact_apply(int act, float v, float x) {
  if (act == SMX_ACT_MULAUX) return v * x;
  if (act == SMX_ACT_DGELU) return v * gelu_erf_grad(x);
  return x > 0.0f ? v : 0.0f;
}

// eight elements at a time; the multiply / add flavours run as packed fp32 pairs
__device__ __forceinline__ void dact_apply8(int act, float* f, const float* x) {
  if (act == SMX_ACT_MULAUX) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) f2_unpack(f2_mul(f2_pack(f[i], f[i + 1]), f2_pack(x[i], x[i + 1])), f[i], f[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = dact_apply(act, f[i], x[i]);
  }
}
__device__ __forceinline__ void add8(float* f, const float* x) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) f2_unpack(f2_add(f2_pack(f[i], f[i + 1]), f2_pack(x[i], x[i + 1])), f[i], f[i + 1]);
}

// generic path: any n, any alignment (per-element predicates; only instantiated in the <FAST = false> kernels)
__device__ __forceinline__ void epilogue_chunk_generic(const Params& p, float (&f)[32], long long c_off,
                                                       long long res_off, int col0) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int col = col0 + j;
    if (col < p.n) {
      float v = p.alpha * f[j];
      if (p.bias) v += __ldg(p.bias + col);
      if (p.act == SMX_ACT_GELU_G) {
        float y, dy;
        gelu_erf_both(v, y, dy);
        p.aux_out[c_off + col] = __float2bfloat16(dy);
        v = y;
      } else if (p.aux_out) {
        p.aux_out[c_off + col] = __float2bfloat16(v);
      }
      if (p.act == SMX_ACT_GELU) v = gelu_erf(v);
      else if (p.act == SMX_ACT_MULAUX) v *= __bfloat162float(p.aux_in[c_off + col]);
      else if (p.act == SMX_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == SMX_ACT_DGELU) v *= gelu_erf_grad(__bfloat162float(p.aux_in[c_off + col]));
      else if (p.act == SMX_ACT_DRELU) v = __bfloat162float(p.aux_in[c_off + col]) > 0.f ? v : 0.f;
      if (p.residual) v += __bfloat162float(p.residual[res_off + col]);
      if (p.out_f32) {
        float* cp = reinterpret_cast<float*>(p.c) + c_off + col;
        if (p.accumulate_f32 && p.atomic) atomicAdd(cp, v);          // split contraction: several CTAs add into this tile
        else *cp = p.accumulate_f32 ? *cp + v : v;
      } else {
        reinterpret_cast<bf16*>(p.c)[c_off + col] = __float2bfloat16(v);
      }
    }
  }
}

__device__ __forceinline__ void epilogue_chunk_fast(const Params& p, float (&f)[32], long long c_off,
                                                    long long res_off, int col0) {
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
      f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
    }
  }
  if (p.act == SMX_ACT_GELU_G) {
    bf16* ap = p.aux_out + c_off + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) gelu_erf_both(f[j + i], f[j + i], d[i]);
      *reinterpret_cast<uint4*>(ap + j) = pack8(d);
    }
  } else if (p.aux_out) {
    bf16* ap = p.aux_out + c_off + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(ap + j) = pack8(f + j);
  }
  const bool has_dact = p.act == SMX_ACT_DGELU || p.act == SMX_ACT_DRELU || p.act == SMX_ACT_MULAUX;
  if (p.act == SMX_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
  } else if (p.act == SMX_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (has_dact) {
    const bf16* xp = p.aux_in + c_off + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(xp + j)), x);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[j + i] = dact_apply(p.act, f[j + i], x[i]);
    }
  }
  if (p.residual) {
    const bf16* rp = p.residual + res_off + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(rp + j)), x);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[j + i] += x[i];
    }
  }
  if (p.out_f32) {
    float* cp = reinterpret_cast<float*>(p.c) + c_off + col0;
    if (p.accumulate_f32 && p.atomic) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) red_add_v4(cp + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else if (p.accumulate_f32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = *reinterpret_cast<float4*>(cp + j);
        o.x += f[j], o.y += f[j + 1], o.z += f[j + 2], o.w += f[j + 3];
        *reinterpret_cast<float4*>(cp + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
  } else {
    bf16* cp = reinterpret_cast<bf16*>(p.c) + c_off + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(cp + j) = pack8(f + j);
  }
}

// ---- pieces of the fast epilogue used by the TMA-store variant (64 columns per step) ----
__device__ __forceinline__ void epi_scale_bias(const Params& p, float (&f)[64], int col0) {
  const float a = p.alpha;
  if (p.bias) {  // one packed FMA per element pair
    const f32x2 a2 = f2_rep(a);
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
      f2_unpack(f2_fma(f2_pack(f[j], f[j + 1]), a2, f2_pack(b4.x, b4.y)), f[j], f[j + 1]);
      f2_unpack(f2_fma(f2_pack(f[j + 2], f[j + 3]), a2, f2_pack(b4.z, b4.w)), f[j + 2], f[j + 3]);
    }
  } else if (a != 1.0f) {
#pragma unroll
    for (int j = 0; j < 64; ++j) f[j] *= a;
  }
}
__device__ __forceinline__ void epi_act_res(const Params& p, float (&f)[64], long long c_off, long long res_off,
                                            int col0) {
  if (p.act == SMX_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 64; j += 2) gelu_erf2(f[j], f[j + 1]);
  } else if (p.act == SMX_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 64; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (p.act == SMX_ACT_DGELU || p.act == SMX_ACT_DRELU || p.act == SMX_ACT_MULAUX) {
    const bf16* xp = p.aux_in + c_off + col0;
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(xp + j)), x);
      dact_apply8(p.act, f + j, x);
    }
  }
  if (p.residual) {
    const bf16* rp = p.residual + res_off + col0;
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(rp + j)), x);
      add8(f + j, x);
    }
  }
}
// same, with the 64 bf16 of aux_in (tma_in == 1) or residual (tma_in == 2) already in registers
__device__ __forceinline__ void epi_act_res_pre(const Params& p, float (&f)[64], const uint4 (&pre)[8], long long c_off,
                                                long long res_off, int col0) {
  if (p.act == SMX_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 64; j += 2) gelu_erf2(f[j], f[j + 1]);
  } else if (p.act == SMX_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 64; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (p.act == SMX_ACT_DGELU || p.act == SMX_ACT_DRELU || p.act == SMX_ACT_MULAUX) {
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      float x[8];
      unpack8(pre[j >> 3], x);
      dact_apply8(p.act, f + j, x);
    }
  }
  if (p.residual) {
    const bf16* rp = p.residual + res_off + col0;
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      float x[8];
      unpack8(p.tma_in == 2 ? pre[j >> 3] : __ldg(reinterpret_cast<const uint4*>(rp + j)), x);
      add8(f + j, x);
    }
  }
}
// one row (64 bf16 = 128 B) of the staging tile, 16-byte chunks XOR-swizzled with the row index
__device__ __forceinline__ void stage_row(uint8_t* stg, int r, const float (&f)[64]) {
  uint8_t* row = stg + r * 128;
  const int sw = r & 7;
#pragma unroll
  for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(row + ((k ^ sw) << 4)) = pack8(f + k * 8);
}

// the same from eight already packed 16-byte chunks: all arithmetic is done BEFORE the caller waits for the previous
// bulk store to release the staging tile, so that wait overlaps the math instead of following it
__device__ __forceinline__ void stage_row_packed(uint8_t* stg, int r, const uint4 (&q)[8]) {
  uint8_t* row = stg + r * 128;
  const int sw = r & 7;
#pragma unroll
  for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(row + ((k ^ sw) << 4)) = q[k];
}
__device__ __forceinline__ void pack_row(const float (&f)[64], uint4 (&q)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) q[k] = pack8(f + k * 8);
}
// SMX_ACT_GELU_G: q = packed gelu'(f) (the auxiliary output), f replaced by gelu(f) in the same sweep
__device__ __forceinline__ void gelu_both_row(float (&f)[64], uint4 (&q)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) gelu_erf_both2(f[k * 8 + i], f[k * 8 + i + 1], d[i], d[i + 1]);
    q[k] = pack8(d);
  }
}

// EPI: 0 regular fused epilogue, 1 LM-head statistics, 2 LM-head dlogits.  FAST: n % 32 == 0 and every
// pointer / stride 16-byte aligned, so the epilogue is pure vector code (smaller I$ footprint).
template <int MODE, int EPI, bool FAST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_aux,
            const __grid_constant__ CUtensorMap tma_in, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* inbar = tempty + 2;  // one per epilogue warp: staged epilogue-input tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbar + EPI_WARPS);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();  // 0 = leader of the pair
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);    // leader only: one arrive.expect_tx covering both CTAs' loads
      mbar_init(&empty[i], 1);   // multicast tcgen05.commit, one arrival per CTA
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * EPI_WARPS);  // leader only: epilogue warps of BOTH CTAs
    }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&inbar[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1) tmem_alloc2(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above is on-chip setup; global memory is touched only below

  const int total_tiles = p.m_tiles * p.n_tiles * p.split_k;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // Lane 0 owns the ring (waits for a free slot, arms the barrier); the boxes of a stage are then
    // issued by different lanes in parallel, each with coordinates it tracks incrementally, so no
    // integer division or serial descriptor math sits on the per-k-block critical path.
    constexpr int N_BOXES = MODE == SMX_GEMM_NT ? 2 : (MODE == SMX_GEMM_NN ? 1 + BN_HALF / 64 : BM / 64 + BN_HALF / 64);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const TileCoord t = decode_tile(p, tile);
      const int n0 = t.n_blk * BN + (int)cta_rank * BN_HALF;  // this CTA's half of the B columns
      // per-lane constants of this tile
      int c0_fixed = 0, row_fixed = 0, batch = 0, smem_off = 0;
      bool is_a = false, active = lane < N_BOXES;
      if (MODE == SMX_GEMM_TN) {
        const int m0 = t.m_blk * PAIR_M + (int)cta_rank * BM;
        if (lane < BM / 64) {
          is_a = true;
          c0_fixed = p.a_col_off[0] + m0 + 64 * lane;
          row_fixed = p.a_row_off[0];
          smem_off = lane * PANEL_BYTES;
        } else if (active) {
          const int pn = lane - BM / 64;
          const int nn = n0 + 64 * pn;
          c0_fixed = p.b_inner_oob;
          if (nn < p.n) {
            const int s = nn / p.seg_len;
            c0_fixed = p.b_col_off[s] + (nn - s * p.seg_len);
            row_fixed = p.b_row_off[s];
          }
          smem_off = A_BYTES + pn * PANEL_BYTES;
        }
      } else {
        batch = t.m_blk / p.m_tiles_per_batch;
        row_fixed = (t.m_blk % p.m_tiles_per_batch) * PAIR_M + (int)cta_rank * BM;
        if (lane == 0) {
          is_a = true;
        } else if (active) {
          smem_off = A_BYTES + (lane - 1) * PANEL_BYTES;
          c0_fixed = MODE == SMX_GEMM_NT ? n0 : n0 + 64 * (lane - 1);
        }
      }
      // running contraction position
      int seg = 0, kin = 0, cb = 0, cr0 = 0;
      if (MODE == SMX_GEMM_TN) {
        cb = t.kb_begin / p.kb_per_batch;
        cr0 = (t.kb_begin % p.kb_per_batch) * BK;
      } else {
        const int k0 = t.kb_begin * BK;
        seg = k0 / p.seg_len;
        kin = k0 - seg * p.seg_len;
      }
      for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
        if (lane == 0) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (cta_rank == 0) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);  // bytes of both CTAs
        }
        __syncwarp();
        if (active) {
          uint8_t* dst = smem + stage * STAGE_BYTES + smem_off;
          if (MODE == SMX_GEMM_TN) {
            tma2_load_3d(dst, is_a ? &tma_a : &tma_b, &full[stage], c0_fixed, cr0 + row_fixed, cb);
          } else if (is_a) {
            tma2_load_3d(dst, &tma_a, &full[stage], p.a_col_off[seg] + kin, row_fixed + p.a_row_off[seg], batch);
          } else if (MODE == SMX_GEMM_NT) {
            tma2_load_2d(dst, &tma_b, &full[stage], p.b_col_off[seg] + kin, c0_fixed);
          } else {
            tma2_load_2d(dst, &tma_b, &full[stage], p.b_col_off[seg] + c0_fixed, p.b_row_off[seg] + kin);
          }
        }
        if (MODE == SMX_GEMM_TN) {
          cr0 += BK;
          if (cr0 >= p.kb_per_batch * BK) cr0 = 0, ++cb;
        } else {
          kin += BK;
          if (kin >= p.seg_len) kin = 0, ++seg;
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && cta_rank == 0) {
      constexpr bool A_MN = (MODE == SMX_GEMM_TN);
      constexpr bool B_MN = (MODE != SMX_GEMM_NT);
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR_M, BN, A_MN, B_MN);
      // K-major SW128: 8-row groups 1024 B apart, 16 K-elements = +32 B inside the swizzle atom.
      // MN-major SW128: 64-wide panels LBO = 8 KiB apart, 8-row K groups 1024 B apart, 16 K = +2 KiB.
      constexpr uint32_t a_lbo = A_MN ? PANEL_BYTES : 16, a_kstep = A_MN ? 2048 : 32;
      constexpr uint32_t b_lbo = B_MN ? PANEL_BYTES : 16, b_kstep = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // The issuing thread is on the critical path (one MMA per 128 tensor-pipe cycles), so the
      // descriptors are reduced to one 32-bit add per operand per MMA: the high word (SBO, version,
      // swizzle mode) is constant and the low word is (address >> 4) | LBO << 16.
      constexpr uint32_t desc_hi = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (static_cast<uint32_t>(kLayoutSW128) << 29);
      const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((a_lbo >> 4) << 16);
      const uint32_t b_lo0 = ((smem_u32(smem) + A_BYTES) >> 4) | ((b_lbo >> 4) << 16);
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const TileCoord t = decode_tile(p, tile);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accum = 0;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_lo = a_lo0 + stage * (STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (STAGE_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adesc = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + kk * (a_kstep >> 4));
            const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + kk * (b_kstep >> 4));
            umma2_ss(d_tmem, adesc, bdesc, idesc, accum);
            accum = 1;
          }
          umma2_commit_mc(&empty[stage], 3);  // frees the stage in both CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_mc(&tfull[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // two warps share a lane quarter: columns [0,128) / [128,256)
    const int c_begin = half * (BN / 64), c_end = c_begin + BN / 64;
    const int row_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t in_phase = 0;
    const bool vec_ok = (p.c_row_stride % 8 == 0) && (p.c_batch_stride % 8 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0) &&
                        (p.residual == nullptr || ((p.res_row_stride % 8 == 0) && (p.res_batch_stride % 8 == 0) &&
                                                   ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0))) &&
                        ((p.aux_out == nullptr) || ((reinterpret_cast<uintptr_t>(p.aux_out) & 15) == 0)) &&
                        ((p.aux_in == nullptr) || ((reinterpret_cast<uintptr_t>(p.aux_in) & 15) == 0)) &&
                        (p.bias == nullptr || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
    int bank = 0;   // two-bank mode: toggles after EVERY bulk store of this warp, across tiles (store k uses bank k & 1,
                    // and waiting until at most one store is pending means store k - 2 has released that bank)
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const TileCoord t = decode_tile(p, tile);
      const int n0 = t.n_blk * BN;
      long long row, c_off, res_off = 0;
      if (MODE == SMX_GEMM_TN) {
        row = (long long)t.m_blk * PAIR_M + cta_rank * BM + row_in_tile;
        c_off = row * p.c_row_stride;
      } else {
        const int bidx = t.m_blk / p.m_tiles_per_batch;
        row = (long long)(t.m_blk % p.m_tiles_per_batch) * PAIR_M + cta_rank * BM + row_in_tile;
        c_off = (long long)bidx * p.c_batch_stride + row * p.c_row_stride;
        res_off = (long long)bidx * p.res_batch_stride + row * p.res_row_stride;
      }
      const bool row_ok = row < p.m;

      const bool use_in = FAST && EPI == 0 && MODE != SMX_GEMM_TN && p.tma_store && p.tma_in != 0;
      uint8_t* stg = smem + OFF_STG + (warp - 2) * STG_BANKS * STG_BYTES;
      int ep_bidx = 0, ep_row0 = 0;
      if (MODE != SMX_GEMM_TN) {
        ep_bidx = t.m_blk / p.m_tiles_per_batch;
        ep_row0 = (t.m_blk % p.m_tiles_per_batch) * PAIR_M + (int)cta_rank * BM + q * 32;
      }
      if (use_in) {  // prefetch the first 64-column input tile while the accumulator is still being produced
        const int col0 = n0 + c_begin * 32;
        if (col0 + 32 < p.n && lane == 0) {
          tma_store_wait_read<0>();
          mbar_expect_tx(&inbar[warp - 2], STG_BYTES);
          tma_load_3d(stg, &tma_in, &inbar[warp - 2], col0, ep_row0, ep_bidx);
        }
      }

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

      if (MODE == SMX_GEMM_NT && EPI == 1) {
        // ---- LM head: online softmax statistics of this half of a 256-column vocabulary tile
        float mx = -INFINITY, se = 0.f, best = -INFINITY;
        int best_idx = 0x7fffffff;
        const long long label = row_ok ? p.lm_labels[row] : -1;
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.n) break;
          uint32_t v[32];
          tmem_ld_x32(t_row + c * 32, v);
          tmem_ld_wait();
          float f[32];
          float cmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            float x = p.alpha * __uint_as_float(v[j]);
            if (col < p.n) {
              if (p.bias) x += __ldg(p.bias + col);
            } else {
              x = -INFINITY;
            }
            f[j] = x;
            cmax = fmaxf(cmax, x);
            if (x > best) best = x, best_idx = col;
          }
          const float m_new = fmaxf(mx, cmax);
          float acc_e = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc_e += __expf(f[j] - m_new);
          se = se * __expf(mx - m_new) + acc_e;
          mx = m_new;
          if (label >= col0 && label < col0 + 32) {
            float lv = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) lv = (label - col0 == j) ? f[j] : lv;
            if (row_ok) p.lm_label_logit[row] = lv;
          }
          __syncwarp();
        }
        if (row_ok)
          p.lm_partial[(row * p.n_tiles + t.n_blk) * 2 + half] = make_float4(mx, se, best, __int_as_float(best_idx));
      } else if (MODE == SMX_GEMM_NT && EPI == 2) {
        // ---- LM head backward: dlogits = (softmax - onehot) * coef, bf16
        const float lse = row_ok ? p.lm_lse[row] : 0.f;
        const float coef = row_ok ? p.lm_coef[row] : 0.f;
        const long long label = row_ok ? p.lm_labels[row] - p.lm_label_off : -1;
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.n) break;
          uint32_t v[32];
          tmem_ld_x32(t_row + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              float x = p.alpha * __uint_as_float(v[j]);
              if (p.bias && col < p.n) x += __ldg(p.bias + col);
              float g = __expf(x - lse) * coef;
              if (col == label) g -= coef;
              f[j] = g;
            }
            bf16* cp = reinterpret_cast<bf16*>(p.c) + c_off + col0;
            if (col0 + 32 <= p.n && vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(cp + j) = pack8(f + j);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n) cp[j] = __float2bfloat16(f[j]);
            }
          }
          __syncwarp();
        }
      } else if (MODE == SMX_GEMM_TN) {
        // ---- weight gradient: fp32, atomics when the contraction is split
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.n) break;
          uint32_t v[32];
          tmem_ld_x32(t_row + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float* cp = reinterpret_cast<float*>(p.c) + c_off + col0;
            const bool full_chunk = col0 + 32 <= p.n;
            if (p.atomic) {
              if (full_chunk && (p.c_row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0)) {
                // 16-byte vector reductions: 8 instead of 32 L2 atomics per row chunk (split-K partial sums)
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  red_add_v4(cp + j, p.alpha * __uint_as_float(v[j]), p.alpha * __uint_as_float(v[j + 1]),
                             p.alpha * __uint_as_float(v[j + 2]), p.alpha * __uint_as_float(v[j + 3]));
              } else if (full_chunk) {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(cp + j, p.alpha * __uint_as_float(v[j]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.n) atomicAdd(cp + j, p.alpha * __uint_as_float(v[j]));
              }
            } else if (full_chunk && (p.c_row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(cp + j) =
                    make_float4(p.alpha * __uint_as_float(v[j]), p.alpha * __uint_as_float(v[j + 1]),
                                p.alpha * __uint_as_float(v[j + 2]), p.alpha * __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n) cp[j] = p.alpha * __uint_as_float(v[j]);
            }
          }
          __syncwarp();
        }
      } else {
        // ---- regular fused epilogue
        if (FAST && p.tma_store) {
          // 64 columns per step: TMEM -> registers -> fused math -> swizzled smem tile -> one TMA store of
          // full 128-byte row segments (coalesced; rows / columns past the edge are clipped by the map).
          const int bidx = ep_bidx, row0 = ep_row0;
#pragma unroll 1
          for (int pr = 0; pr < BN / 128; ++pr) {
            const int c = c_begin + 2 * pr;
            const int col0 = n0 + c * 32;
            if (col0 >= p.n) break;  // warp-uniform
            float f[64];
            {
              // both 32-column loads in flight before the wait (c is even, so chunk c + 1 is always inside the
              // accumulator; past the matrix edge it holds padding that is never stored)
              uint32_t v[32], w[32];
              tmem_ld_x32(t_row + c * 32, v);
              tmem_ld_x32(t_row + (c + 1) * 32, w);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]), f[32 + j] = __uint_as_float(w[j]);
            }
            const bool two = col0 + 32 < p.n;  // n % 32 == 0 here, so the second chunk is all-or-nothing
            if (row_ok) {
              if (two) {
                epi_scale_bias(p, f, col0);
              } else {
                float g[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) g[j] = f[j];
                epilogue_chunk_fast(p, g, c_off, res_off, col0);  // lone trailing chunk: direct path
              }
            }
            if (two) {
              if (p.aux_out) {
                uint4 q[8];
                if (p.act == SMX_ACT_GELU_G) gelu_both_row(f, q);
                else pack_row(f, q);
                // two banks (and no TMA-loaded input sharing the tile): only the store BEFORE the previous one has to
                // have released its tile, so the wait no longer serialises behind the store just issued
                uint8_t* sa = stg;
                if (STG_BANKS == 2 && !use_in) {
                  sa = stg + bank * STG_BYTES;
                  bank ^= 1;
                  if (lane == 0) tma_store_wait_read<1>();
                } else {
                  if (lane == 0) tma_store_wait_read<0>();
                }
                __syncwarp();
                stage_row_packed(sa, lane, q);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_3d(&tma_aux, sa, col0, row0, bidx);
                  tma_store_commit();
                }
              }
              uint8_t* sc = stg;
              if (use_in) {
                mbar_wait(&inbar[warp - 2], in_phase);
                in_phase ^= 1;
                uint4 pre[8];
                const uint8_t* rowp = stg + lane * 128;
#pragma unroll
                for (int k = 0; k < 8; ++k) pre[k] = *reinterpret_cast<const uint4*>(rowp + ((k ^ (lane & 7)) << 4));
                __syncwarp();  // every lane has its input row before the tile is overwritten with the output
                if (row_ok) epi_act_res_pre(p, f, pre, c_off, res_off, col0);
                stage_row(stg, lane, f);
              } else {
                if (row_ok) epi_act_res(p, f, c_off, res_off, col0);
                uint4 q[8];
                pack_row(f, q);
                if (STG_BANKS == 2) {
                  sc = stg + bank * STG_BYTES;
                  bank ^= 1;
                  if (lane == 0) tma_store_wait_read<1>();
                } else {
                  if (lane == 0) tma_store_wait_read<0>();
                }
                __syncwarp();
                stage_row_packed(sc, lane, q);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&tma_c, sc, col0, row0, bidx);
                tma_store_commit();
                if (use_in && pr + 1 < BN / 128 && col0 + 64 + 32 < p.n) {  // next input tile of this warp
                  tma_store_wait_read<0>();
                  mbar_expect_tx(&inbar[warp - 2], STG_BYTES);
                  tma_load_3d(stg, &tma_in, &inbar[warp - 2], col0 + 64, row0, bidx);
                }
              }
            }
            __syncwarp();
          }
        } else if (FAST) {
          for (int c = c_begin; c < c_end; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= p.n) break;  // warp-uniform
            uint32_t v[32];
            tmem_ld_x32(t_row + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              epilogue_chunk_fast(p, f, c_off, res_off, col0);
            }
            __syncwarp();
          }
        } else {
          for (int c = c_begin; c < c_end; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= p.n) break;  // warp-uniform
            uint32_t v[32];
            tmem_ld_x32(t_row + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              epilogue_chunk_generic(p, f, c_off, res_off, col0);
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);  // the leader's MMA thread owns both accumulators
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait_read<0>();  // staging smem must outlive the last bulk store's read
  tc_fence_before_sync();
  cluster_sync_all();  // the peer's smem / TMEM must stay alive until the leader's last MMA has retired
  if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

template <int MODE, int EPI, bool FAST>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx,
                  const CUtensorMap& ti, const Params& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<MODE, EPI, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * (total < pairs ? total : pairs));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see sm100_prims.cuh (pdl_trigger / pdl_wait)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SMX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<MODE, EPI, FAST>, ta, tb, tc, tx, ti, p));
  return 0;
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// all chunks full and every epilogue operand vectorisable?
static bool fast_epilogue_ok(const Params& p) {
  return p.n % 32 == 0 && p.c_row_stride % 8 == 0 && p.c_batch_stride % 8 == 0 && aligned16(p.c) &&
         (!p.residual || (p.res_row_stride % 8 == 0 && p.res_batch_stride % 8 == 0 && aligned16(p.residual))) &&
         aligned16(p.aux_out) && aligned16(p.aux_in) && aligned16(p.bias);
}

}  // namespace gemm
}  // namespace smx

namespace smx {
namespace gemm {
struct LmExtra {
  int epi;
  float4* partial;
  float* label_logit;
  const long long* labels;
  const float* lse;
  const float* coef;
  long long label_off;
};
int run(const SmxGemm* g, const LmExtra* lm, void* stream);
}  // namespace gemm
}  // namespace smx

extern "C" int smx_gemm(const SmxGemm* g, void* stream) { return smx::gemm::run(g, nullptr, stream); }

// Tensor maps for the epilogue's TMA stores of C (and aux_out): bf16 [batches][m][n] with C's strides,
// box 64 columns x 32 rows, 128-byte swizzle.  Falls back to direct stores for fp32 / accumulate outputs.
static int make_store_maps(const SmxGemm* g, smx::gemm::Params* p, CUtensorMap* tc, CUtensorMap* tx, CUtensorMap* ti) {
  using namespace smx;
  p->tma_store = 0;
  p->tma_in = 0;
  *tc = CUtensorMap();
  *tx = CUtensorMap();
  *ti = CUtensorMap();
  if (p->out_f32 || p->n < 64) return 0;
  const uint64_t dims[3] = {(uint64_t)g->n, (uint64_t)g->m, (uint64_t)g->batches};
  const uint64_t str[2] = {(uint64_t)g->c_row_stride,
                           (uint64_t)(g->batches > 1 ? g->c_batch_stride : g->c_row_stride * g->m)};
  const uint32_t box[3] = {64, 32, 1};
  if (encode_tmap_bf16(tc, g->c, 3, dims, str, box, true)) return -1;
  if (g->aux_out) {
    if (encode_tmap_bf16(tx, g->aux_out, 3, dims, str, box, true)) return -1;
  } else {
    *tx = *tc;
  }
  *ti = *tc;
  if (!g->aux_out) {  // the staging tile is free for an input: activation-gradient operand first, else the residual
    const bool dact = g->act == SMX_ACT_DGELU || g->act == SMX_ACT_DRELU || g->act == SMX_ACT_MULAUX;
    if (dact) {
      if (encode_tmap_bf16(ti, g->aux_in, 3, dims, str, box, true)) return -1;
      p->tma_in = 1;
    } else if (g->residual) {
      const uint64_t rstr[2] = {(uint64_t)g->res_row_stride,
                                (uint64_t)(g->batches > 1 ? g->res_batch_stride : g->res_row_stride * g->m)};
      if (encode_tmap_bf16(ti, g->residual, 3, dims, rstr, box, true)) return -1;
      p->tma_in = 2;
    }
  }
  p->tma_store = 1;
  return 0;
}

int smx::gemm::run(const SmxGemm* g, const LmExtra* lm, void* stream) {
  using namespace smx;
  using namespace smx::gemm;
  SMX_REQUIRE(g != nullptr, "smx_gemm: null descriptor");
  SMX_REQUIRE(g->mode >= 0 && g->mode <= 2, "smx_gemm: bad mode %d", g->mode);
  SMX_REQUIRE(g->m > 0 && g->n > 0 && g->k > 0 && g->batches > 0, "smx_gemm: empty problem m=%lld n=%lld k=%lld b=%lld",
              (long long)g->m, (long long)g->n, (long long)g->k, (long long)g->batches);
  SMX_REQUIRE(g->nseg >= 1 && g->nseg <= SMX_MAX_SEG, "smx_gemm: nseg %d out of range", g->nseg);
  SMX_REQUIRE(g->nseg == 1 || g->seg_len % 64 == 0, "smx_gemm: seg_len %d must be a multiple of 64", g->seg_len);
  SMX_REQUIRE(g->a.ptr && g->b.ptr && g->c, "smx_gemm: null operand");

  Params p;
  memset(&p, 0, sizeof(p));
  p.mode = g->mode;
  p.m = g->m;
  p.n = g->n;
  p.batches = (int)g->batches;
  p.nseg = g->nseg;
  p.seg_len = g->seg_len;
  for (int i = 0; i < SMX_MAX_SEG; ++i) {
    p.a_row_off[i] = g->a_row_off[i];
    p.a_col_off[i] = g->a_col_off[i];
    p.b_row_off[i] = g->b_row_off[i];
    p.b_col_off[i] = g->b_col_off[i];
  }
  p.c = g->c;
  p.c_row_stride = g->c_row_stride;
  p.c_batch_stride = g->c_batch_stride;
  p.act = g->act;
  p.alpha = g->alpha;
  p.bias = g->bias;
  p.residual = reinterpret_cast<const bf16*>(g->residual);
  p.res_row_stride = g->res_row_stride;
  p.res_batch_stride = g->res_batch_stride;
  p.aux_out = reinterpret_cast<bf16*>(g->aux_out);
  p.aux_in = reinterpret_cast<const bf16*>(g->aux_in);
  p.n_tiles = (int)ceil_div(g->n, BN);
  p.split_k = 1;
  p.accumulate_f32 = (g->mode != SMX_GEMM_TN && g->accumulate) ? 1 : 0;
  if (lm) {
    p.epi = lm->epi;
    p.lm_partial = lm->partial;
    p.lm_label_logit = lm->label_logit;
    p.lm_labels = lm->labels;
    p.lm_lse = lm->lse;
    p.lm_coef = lm->coef;
    p.lm_label_off = lm->label_off;
  }

  CUtensorMap ta, tb, tc, tx, ti;
  const uint64_t a_dims[3] = {(uint64_t)g->a.inner, (uint64_t)g->a.rows, (uint64_t)g->a.batches};
  const uint64_t a_str[2] = {(uint64_t)g->a.row_stride,
                             (uint64_t)(g->a.batches > 1 ? g->a.batch_stride : g->a.row_stride * g->a.rows)};
  const uint64_t b_dims[3] = {(uint64_t)g->b.inner, (uint64_t)g->b.rows, (uint64_t)g->b.batches};
  const uint64_t b_str[2] = {(uint64_t)g->b.row_stride,
                             (uint64_t)(g->b.batches > 1 ? g->b.batch_stride : g->b.row_stride * g->b.rows)};

  if (g->mode == SMX_GEMM_TN) {
    SMX_REQUIRE(g->n == (int64_t)g->nseg * g->seg_len, "smx_gemm TN: n %lld != nseg*seg_len", (long long)g->n);
    p.out_f32 = 1;
    p.m_tiles_per_batch = 0;
    p.m_tiles = (int)ceil_div(g->m, PAIR_M);
    p.kb_per_batch = (int)ceil_div(g->k, BK);
    p.kblocks = p.kb_per_batch * (int)g->batches;
    p.split_k = g->split_k > 1 ? g->split_k : 1;
    if (p.split_k > p.kblocks) p.split_k = p.kblocks;
    p.atomic = (p.split_k > 1 || g->accumulate) ? 1 : 0;
    p.b_inner_oob = (int)g->b.inner + 64;
    const uint32_t box[3] = {64, 64, 1};
    if (encode_tmap_bf16(&ta, g->a.ptr, 3, a_dims, a_str, box, true)) return -1;
    if (encode_tmap_bf16(&tb, g->b.ptr, 3, b_dims, b_str, box, true)) return -1;
    return launch<SMX_GEMM_TN, 0, false>(ta, tb, ta, ta, ta, p, (cudaStream_t)stream);
  }

  SMX_REQUIRE(g->k == (int64_t)g->nseg * g->seg_len, "smx_gemm: k %lld != nseg*seg_len", (long long)g->k);
  SMX_REQUIRE((g->act != SMX_ACT_DGELU && g->act != SMX_ACT_DRELU && g->act != SMX_ACT_MULAUX) || g->aux_in,
              "smx_gemm: DGELU/DRELU/MULAUX need aux_in");
  SMX_REQUIRE(g->act != SMX_ACT_GELU_G || g->aux_out, "smx_gemm: GELU_G needs aux_out");
  p.out_f32 = g->out_dtype == SMX_OUT_F32;
  p.m_tiles_per_batch = (int)ceil_div(g->m, PAIR_M);
  p.m_tiles = p.m_tiles_per_batch * (int)g->batches;
  p.kblocks = (int)ceil_div(g->k, BK);
  p.kb_per_batch = p.kblocks;
  // split contraction for fp32-ACCUMULATING outputs (the LM-head data gradient: 24 output tiles, K = 8192 per vocabulary
  // chunk): partial sums are added with fp32 reductions, which the accumulate contract (C += A.B) already allows
  if (p.out_f32 && p.accumulate_f32 && g->split_k > 1 && g->nseg == 1 && !g->bias && !g->residual && g->act == SMX_ACT_NONE) {
    p.split_k = g->split_k < p.kblocks ? g->split_k : p.kblocks;
    p.atomic = 1;
  }
  // raster order: the LARGER operand should stream from DRAM once (see decode_tile)
  p.n_fastest = (g->m * (int64_t)g->batches > g->n) ? 1 : 0;
  const uint32_t a_box[3] = {64, 128, 1};
  if (encode_tmap_bf16(&ta, g->a.ptr, 3, a_dims, a_str, a_box, true)) return -1;
  if (g->mode == SMX_GEMM_NT) {
    const uint32_t b_box[2] = {64, 128};  // each CTA of the pair loads half of the 256 B rows
    if (encode_tmap_bf16(&tb, g->b.ptr, 2, b_dims, b_str, b_box, true)) return -1;
    if (p.epi == 1) return launch<SMX_GEMM_NT, 1, false>(ta, tb, ta, ta, ta, p, (cudaStream_t)stream);
    if (p.epi == 2) return launch<SMX_GEMM_NT, 2, false>(ta, tb, ta, ta, ta, p, (cudaStream_t)stream);
    if (!fast_epilogue_ok(p)) return launch<SMX_GEMM_NT, 0, false>(ta, tb, ta, ta, ta, p, (cudaStream_t)stream);
    if (make_store_maps(g, &p, &tc, &tx, &ti)) return -1;
    return launch<SMX_GEMM_NT, 0, true>(ta, tb, tc, tx, ti, p, (cudaStream_t)stream);
  }
  const uint32_t b_box[2] = {64, 64};
  if (encode_tmap_bf16(&tb, g->b.ptr, 2, b_dims, b_str, b_box, true)) return -1;
  if (!fast_epilogue_ok(p)) return launch<SMX_GEMM_NN, 0, false>(ta, tb, ta, ta, ta, p, (cudaStream_t)stream);
  if (make_store_maps(g, &p, &tc, &tx, &ti)) return -1;
  return launch<SMX_GEMM_NN, 0, true>(ta, tb, tc, tx, ti, p, (cudaStream_t)stream);
}
