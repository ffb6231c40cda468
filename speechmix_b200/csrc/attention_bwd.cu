// Flash-style attention backward for head_dim 64 on sm_100a (no atomics, deterministic):
//   1. dQ kernel: one CTA per 128-row query tile, loops over key sub-tiles; also writes delta[b,h,q] = rowsum(dO * O)
//        S = Q.K^T, dP = dO.V^T -> dS -> dQ += dS.K
//   2. dKV kernel: one CTA per 128-row key/value tile, loops over query sub-tiles
//        S^T = K.Q^T, dP^T = V.dO^T (TMEM) -> P^T, dS^T (bf16, TMEM) -> dV += P^T.dO, dK += dS^T.Q (TMEM)
// P is recomputed from the forward LSE; scores never reach HBM.  The operands that
// are contracted over their row index (dO, Q, K as "B") are read MN-major straight
// from the TMA-loaded tiles, so nothing is transposed in memory.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

#include <stdlib.h>
#include <string.h>

namespace smx {
namespace attn {

int make_head_map(CUtensorMap* m, const void* ptr, int t, int heads, int batch, long long row_stride,
                  long long batch_stride);
int make_head_map_rows(CUtensorMap* m, const void* ptr, int t, int heads, int batch, long long row_stride,
                       long long batch_stride, int box_rows);

constexpr int D = 64;
constexpr int TILE = 128 * D * 2;   // 16 KiB
constexpr float kLog2e = 1.4426950408889634f;

struct BwdParams {
  const float* lse;
  const float* delta;
  const float* bias;
  float* dbias;      // [heads, tq, tk] fp32, accumulated with atomics over the batch (T5 relative position bias)
  float inv_scale;
  const bf16* o;      // forward output, for delta = rowsum(dO * O) inside the dQ kernel
  long long o_row_stride, o_batch_stride;
  float* delta_out;   // written by the dQ kernel (read by the dK/dV kernel that runs after it)
  bf16 *dq, *dk, *dv;
  long long dq_row_stride, dq_batch_stride, dk_row_stride, dk_batch_stride, dv_row_stride, dv_batch_stride;
  int batch, heads, tq, tk, causal;
  float scale, scale_log2;
  const int* kv_len;  // optional [batch] key counts (k / v rows past it hold zeros, see the header)
  const unsigned long long* drop_state;   // dropout on the probabilities, regenerated here: {seed, step}, call, p
  uint32_t drop_call;
  float drop_p;
};

__device__ __forceinline__ int effective_tk(const BwdParams& p, int b) {
  if (p.kv_len == nullptr) return p.tk;
  const int l = p.kv_len[b];
  return l < 1 ? 1 : (l < p.tk ? l : p.tk);
}

// live == false: the row exists in memory but is masked out (key past kv_len[b]) -> its gradient is zero
__device__ __forceinline__ void store_out_row(bf16* dst, uint32_t taddr, bool valid, bool live = true) {
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
        *reinterpret_cast<uint4*>(dst + c * 32 + i) = live ? u : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    __syncwarp();
  }
}

// 32 fp32 accumulator columns of this thread's TMEM lane -> 32 bf16 of one output row
__device__ __forceinline__ void store_out_half(bf16* dst, uint32_t taddr, bool valid, bool live = true) {
  uint32_t v[32];
  tmem_ld_x32(taddr, v);
  tmem_ld_wait();
  if (valid) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 u;
      u.x = pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      u.y = pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
      u.z = pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
      u.w = pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
      *reinterpret_cast<uint4*>(dst + i) = live ? u : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// =====================================================================================
// Pipelined TMEM-operand kernels.  (A first generation with 128-wide single-buffered score tiles and every operand in
// shared memory measured 114 clk per 128x64x16 MMA inside the kernel against 45 clk in isolation,
// tools/micro/mma_rate.cu; it was removed in round 2.)
// Here the stationary tile (K,V for dK/dV; Q,dO for dQ) is copied ONCE into TMEM and used as the A operand
// (tcgen05.mma with A in TMEM), the streamed dimension is walked in 64-wide sub-tiles whose score / dP
// accumulators are double-buffered in TMEM, P^T / dS^T are written back to TMEM as bf16 (aliasing the fp32
// scores they came from, FA4-style) and consumed from there as A operands of the gradient MMAs, and two groups
// of four compute warps alternate on the sub-tiles.  Shared memory only carries the streamed 64-row tiles.
// TMEM columns: [0,32) [32,64) stationary A operands | buffer b: scores [64+128b, +64), dP [128+128b, +64)
//               | accumulators from 320.
// =====================================================================================
constexpr int SUB = 64;                  // streamed rows per sub-tile
constexpr int SUBTILE = SUB * D * 2;     // 8 KiB: a [64 x 64] bf16 tile
constexpr int NST = 4;                   // TMA ring depth
constexpr int COL_A0 = 0, COL_A1 = 32, COL_BUF = 64, COL_ACC = 320;

__device__ __forceinline__ void tmem_st_x32_raw(uint32_t taddr, const uint32_t (&v)[32]) { tmem_st_x32(taddr, v); }

// Lean descriptor form (the issuing thread shares its scheduler with MUFU-bound compute warps, so every
// instruction in front of a tcgen05.mma costs issue slots): the high descriptor word (SBO = 1024 B, version,
// 128B swizzle) is a constant, the low word is (address >> 4) | (LBO >> 4) << 16 plus a per-K-step immediate.
constexpr uint32_t kDescHi = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (static_cast<uint32_t>(kLayoutSW128) << 29);
__device__ __forceinline__ uint64_t lean_desc(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }
// D[128 x 64] = A (TMEM, bf16 [128 x 64] in 32 columns) x B^T, B = K-major [64 x 64] smem sub-tile; 4 K-steps
__device__ __forceinline__ void mma_tA_x_subT(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_sub, uint32_t idesc) {
  const uint32_t lo = (b_sub >> 4) | ((16u >> 4) << 16);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ts(d_tmem, a_tmem + kk * 8, lean_desc(lo + kk * 2), idesc, kk > 0 ? 1u : 0u);
}
// D[128 x 64] (+)= A (TMEM, bf16 [128 x 64]) x B, B = MN-major [64 rows(K) x 64] smem sub-tile; 4 K-steps
__device__ __forceinline__ void mma_tA_x_sub(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_sub, uint32_t idesc,
                                             bool accumulate) {
  const uint32_t lo = (b_sub >> 4) | ((static_cast<uint32_t>(SUBTILE) >> 4) << 16);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    umma_ts(d_tmem, a_tmem + kk * 8, lean_desc(lo + kk * 128), idesc, (accumulate || kk > 0) ? 1u : 0u);
}
// the same with the bf16 A operand stored as two K halves of 32 elements: K steps 0,1 at a_tmem + {0, 8}, K steps 2,3
// at a_tmem + 32 + {0, 8} (each half over the first 16 columns of the 32 fp32 columns it was computed from)
__device__ __forceinline__ void mma_tA2_x_sub(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_sub, uint32_t idesc,
                                              bool accumulate) {
  const uint32_t lo = (b_sub >> 4) | ((static_cast<uint32_t>(SUBTILE) >> 4) << 16);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    umma_ts(d_tmem, a_tmem + (kk >> 1) * 32 + (kk & 1) * 8, lean_desc(lo + kk * 128), idesc, (accumulate || kk > 0) ? 1u : 0u);
}
// copy row r of a K-major 128B-swizzled [128 x 64] bf16 smem tile into 32 TMEM columns of this thread's lane
__device__ __forceinline__ void smem_row_to_tmem(const uint8_t* tile, int r, uint32_t taddr) {
  uint32_t v[32];
  const uint8_t* row = tile + r * 128;
  const int sw = r & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + ((c ^ sw) << 4));
    v[c * 4 + 0] = u.x, v[c * 4 + 1] = u.y, v[c * 4 + 2] = u.z, v[c * 4 + 3] = u.w;
  }
  tmem_st_x32(taddr, v);
}

// ---- co-resident variants: 320 threads (8 compute warps, producer, MMA issuer), 256 TMEM columns and ~97 KiB of
// shared memory per CTA, so TWO CTAs share an SM.  A row of a 64-wide sub-tile is split between TWO threads (32
// columns each; warps 0-3 / 4-7), so four compute warps share every scheduler: with one thread per row the MUFU / FMA
// chains of two warps per scheduler left ~65 % of the issue slots empty (round-1 ncu: issue active 27 %).
// Two CTAs per SM: while one CTA's compute warps turn a sub-tile into P^T / dS^T
// (MUFU bound) the other CTA's MMAs run, and the prologue / epilogue of one CTA (TMEM alloc, first TMA round trip,
// gradient store) hides behind the main loop of the other -- with one 512-column CTA per SM those fixed costs
// were ~6k of ~25k cycles per 128-row tile.
constexpr int BWD2_THREADS = 320;   // 8 compute warps (lane quarter x column half), TMA producer, MMA issuer

// D[128 x 64] = A (smem, K-major [128 x 64]) x B^T, B = K-major [64 x 64] smem sub-tile; 4 K-steps
__device__ __forceinline__ void mma_sA_x_subT(uint32_t d_tmem, uint32_t a_tile, uint32_t b_sub, uint32_t idesc) {
  const uint32_t alo = (a_tile >> 4) | ((16u >> 4) << 16);
  const uint32_t blo = (b_sub >> 4) | ((16u >> 4) << 16);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(d_tmem, lean_desc(alo + kk * 2), lean_desc(blo + kk * 2), idesc, kk > 0 ? 1u : 0u);
}

namespace kv2 {
constexpr int OFF_K = 0, OFF_V = TILE, OFF_Q = 2 * TILE, OFF_DO = OFF_Q + NST * SUBTILE,
              OFF_STAT = OFF_DO + NST * SUBTILE;  // stat: [buf][128] floats (lse | delta of 64 queries)
constexpr int OFF_BAR = OFF_STAT + 2 * 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int COL_ST = 0, COL_DPT = 64, COL_DV = 128, COL_DK = 192;   // P^T aliases S^T, dS^T aliases dP^T
enum { B_KV = 0, B_QFULL = 1, B_QEMPTY = B_QFULL + NST, B_STFULL = B_QEMPTY + NST, B_PDSFULL = B_STFULL + 1,
       B_DONE = B_PDSFULL + 1, B_COUNT = B_DONE + 1 };
}  // namespace kv2

// DROP = true adds a packed-pair path for plain attention WITH dropout on the probabilities (mask regenerated from the
// counter-based generator); the general path (causal / bias, with or without dropout) is shared.  A separate
// instantiation, so the kernel of the deterministic step is compiled exactly as before the path existed.
template <bool DROP>
__global__ void __launch_bounds__(BWD2_THREADS, 2)
attn_bwd_dkv2_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                     const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo,
                     const BwdParams p) {
  pdl_trigger();
  pdl_wait();
  using namespace kv2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int head = blockIdx.y, b = blockIdx.z;
  const int nq_sub = (p.tq + SUB - 1) / SUB;
  const int tk = effective_tk(p, b);
  int i_start = 0;
  if (p.causal) {  // first query row that can see key kv0:  q >= kv0 - (tk - tq)
    int qmin = kv0 - (p.tk - p.tq);
    if (qmin < 0) qmin = 0;
    i_start = qmin / SUB;
    if (i_start > nq_sub) i_start = nq_sub;
  }
  const int n_iter = kv0 >= tk ? 0 : nq_sub - i_start;   // a tile of masked keys only writes zeros

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(&bars[B_KV], 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars[B_QFULL + s], 1);
      mbar_init(&bars[B_QEMPTY + s], 1);
    }
    mbar_init(&bars[B_STFULL], 1);
    mbar_init(&bars[B_PDSFULL], 256);
    mbar_init(&bars[B_DONE], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 8 && lane == 0) {
    // the loads of this CTA start BEFORE its TMEM allocation: when the co-resident CTA still owns the other half
    // of TMEM (or a previous CTA has not released its columns yet) the first tiles are already in flight
    mbar_expect_tx(&bars[B_KV], 2 * TILE);
    tma_load_4d(smem + OFF_K, &mk, &bars[B_KV], 0, kv0, head, b);
    tma_load_4d(smem + OFF_V, &mv, &bars[B_KV], 0, kv0, head, b);
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 8) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % NST;
        const uint32_t par = (it / NST) & 1;
        const int qr = (i_start + it) * SUB;
        mbar_wait(&bars[B_QEMPTY + st], par ^ 1);
        mbar_expect_tx(&bars[B_QFULL + st], 2 * SUBTILE);
        tma_load_4d(smem + OFF_Q + st * SUBTILE, &mq, &bars[B_QFULL + st], 0, qr, head, b);
        tma_load_4d(smem + OFF_DO + st * SUBTILE, &mdo, &bars[B_QFULL + st], 0, qr, head, b);
      }
    }
  } else if (warp == 9) {
    // The WHOLE warp runs the loop (waits and address arithmetic are warp-uniform, so the tcgen05.mma operands sit in
    // uniform registers); one elected lane issues.  Under `if (lane == 0)` every MMA cost ~80 clk of issue (a
    // register -> uniform-register broadcast loop per descriptor) against ~45 clk of tensor-pipe time.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, SUB, false, false);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);
    mbar_wait(&bars[B_KV], 0);
    // in-order tensor pipe: the scores of sub-tile it+1 may overwrite P^T / dS^T of sub-tile it without a
    // barrier because the gradient MMAs that read them are issued first
    for (int it = 0; it < n_iter; ++it) {
      const int st = it % NST;
      mbar_wait(&bars[B_QFULL + st], (it / NST) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_sA_x_subT(tmem_base + COL_ST, sbase + OFF_K, sbase + OFF_Q + st * SUBTILE, idesc_s);
        mma_sA_x_subT(tmem_base + COL_DPT, sbase + OFF_V, sbase + OFF_DO + st * SUBTILE, idesc_s);
        umma_commit(&bars[B_STFULL]);
      }
      __syncwarp();
      mbar_wait(&bars[B_PDSFULL], it & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_tA2_x_sub(tmem_base + COL_DV, tmem_base + COL_ST, sbase + OFF_DO + st * SUBTILE, idesc_o, it > 0);   // P^T . dO
        mma_tA2_x_sub(tmem_base + COL_DK, tmem_base + COL_DPT, sbase + OFF_Q + st * SUBTILE, idesc_o, it > 0);   // dS^T . Q
        umma_commit(&bars[B_QEMPTY + st]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bars[B_DONE]);
    __syncwarp();
  } else {
    const int q = warp & 3;         // TMEM lane quarter
    const int hf = warp >> 2;       // column half of the 64-wide sub-tile this thread owns: queries [32 hf, 32 hf + 32)
    const int r = q * 32 + lane;    // key row inside the tile
    const int kvi = kv0 + r;
    const int gt = warp * 32 + lane;  // 0..255; the first 128 threads stage lse / delta
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stat = reinterpret_cast<float*>(smem + OFF_STAT);
    const DropKey dkey = drop_key(p.drop_state, p.drop_call, p.drop_p);   // thresh 0: off
    const uint32_t drop_hp = static_cast<uint32_t>((p.tk + 1) >> 1);
    const uint32_t drop_bh = static_cast<uint32_t>((long long)b * p.heads + head) * static_cast<uint32_t>(p.tq);
    const bool plain = !p.causal && p.bias == nullptr;
    const bool lean = plain && dkey.thresh == 0u;           // packed-pair path of the deterministic step
    const bool lean_drop = DROP && plain && !lean;          // the same plus the regenerated mask
    // DROP, lean: this thread owns ONE key, so which 16 bits of a pair's hash are its own is a per-thread constant, and
    // the hash input ((bh + q) * hp + (key >> 1)) * C + seed-key is linear in the query index: one IMAD per element
    const uint32_t drop_sh = (kvi & 1) ? 0u : 16u;
    const uint32_t drop_thi = dkey.thresh << 16;
    const uint32_t drop_step = drop_hp * 0x9e3779b1u;
    const uint32_t drop_in0 = (drop_bh * drop_hp + static_cast<uint32_t>(kvi >> 1)) * 0x9e3779b1u + dkey.key;
    // lse (log2 domain) / delta*scale of the 64 queries of a sub-tile: thread gt < 64 owns lse[gt], the others
    // delta[gt - 64]; the global load for the NEXT sub-tile is issued one sub-tile ahead.
    const float* stat_src = (gt < 64 ? p.lse : p.delta) + ((long long)b * p.heads + head) * p.tq;
    const float stat_mul = gt < 64 ? -kLog2e : -p.scale;   // stored NEGATED: both are subtracted through an FMA
    auto load_stat = [&](int it) -> float {
      const int qi = (i_start + it) * SUB + (gt & 63);
      return (gt < 128 && it < n_iter && qi < p.tq) ? stat_src[qi] * stat_mul : 0.f;
    };
    float stat_next = load_stat(0);
    for (int it = 0; it < n_iter; ++it) {
      const int q0 = (i_start + it) * SUB;
      float* st = stat + (it & 1) * 128;
      if (gt < 128) st[gt] = stat_next;
      stat_next = load_stat(it + 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&bars[B_STFULL], it & 1);
      tc_fence_after_sync();
      uint32_t pk[16], dk[16];   // P^T and dS^T of this thread's 32 queries, bf16 pairs
      if (lean_drop) {
        // the plain path below plus the regenerated mask: dS = P (keep / (1-p) dP - delta), P_drop = keep ? P / (1-p) : 0
        const f32x2 sl2 = f2_rep(p.scale_log2);
        const float mk = p.scale * dkey.scale, ps = dkey.scale;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int c = 2 * hf + c2;
          uint32_t sv[16], dv[16];
          tmem_ld_x16(t_row + COL_ST + c * 16, sv);
          tmem_ld_x16(t_row + COL_DPT + c * 16, dv);
          tmem_ld_wait();
          const uint32_t in_c = drop_in0 + static_cast<uint32_t>(q0 + c * 16) * drop_step;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(st + c * 16 + i);        // -lse (log2 domain)
            const float4 d4 = *reinterpret_cast<const float4*>(st + 64 + c * 16 + i);   // -delta * scale
            float e0, e1, e2, e3;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2, f2_pack(l4.x, l4.y)), e0, e1);
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2, f2_pack(l4.z, l4.w)), e2, e3);
            e0 = ex2_approx(e0), e1 = ex2_approx(e1), e2 = ex2_approx(e2), e3 = ex2_approx(e3);
            const bool k0 = (mix32(in_c + static_cast<uint32_t>(i) * drop_step) << drop_sh) >= drop_thi;
            const bool k1 = (mix32(in_c + static_cast<uint32_t>(i + 1) * drop_step) << drop_sh) >= drop_thi;
            const bool k2 = (mix32(in_c + static_cast<uint32_t>(i + 2) * drop_step) << drop_sh) >= drop_thi;
            const bool k3 = (mix32(in_c + static_cast<uint32_t>(i + 3) * drop_step) << drop_sh) >= drop_thi;
            float g0, g1, g2, g3;
            f2_unpack(f2_mul(f2_pack(e0, e1), f2_fma(f2_pack(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])),
                                                     f2_pack(k0 ? mk : 0.f, k1 ? mk : 0.f), f2_pack(d4.x, d4.y))), g0, g1);
            f2_unpack(f2_mul(f2_pack(e2, e3), f2_fma(f2_pack(__uint_as_float(dv[i + 2]), __uint_as_float(dv[i + 3])),
                                                     f2_pack(k2 ? mk : 0.f, k3 ? mk : 0.f), f2_pack(d4.z, d4.w))), g2, g3);
            f2_unpack(f2_mul(f2_pack(e0, e1), f2_pack(k0 ? ps : 0.f, k1 ? ps : 0.f)), e0, e1);
            f2_unpack(f2_mul(f2_pack(e2, e3), f2_pack(k2 ? ps : 0.f, k3 ? ps : 0.f)), e2, e3);
            pk[c2 * 8 + (i >> 1)] = pack_bf16x2(e0, e1);
            pk[c2 * 8 + (i >> 1) + 1] = pack_bf16x2(e2, e3);
            dk[c2 * 8 + (i >> 1)] = pack_bf16x2(g0, g1);
            dk[c2 * 8 + (i >> 1) + 1] = pack_bf16x2(g2, g3);
          }
        }
      } else if (lean) {
        // two 16-column chunks; fp32 pairs
        const f32x2 sl2 = f2_rep(p.scale_log2), sc2 = f2_rep(p.scale);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int c = 2 * hf + c2;
          uint32_t sv[16], dv[16];
          tmem_ld_x16(t_row + COL_ST + c * 16, sv);
          tmem_ld_x16(t_row + COL_DPT + c * 16, dv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(st + c * 16 + i);        // -lse (log2 domain)
            const float4 d4 = *reinterpret_cast<const float4*>(st + 64 + c * 16 + i);   // -delta * scale
            float e0, e1, e2, e3;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2, f2_pack(l4.x, l4.y)), e0, e1);
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2, f2_pack(l4.z, l4.w)), e2, e3);
            e0 = ex2_approx(e0), e1 = ex2_approx(e1), e2 = ex2_approx(e2), e3 = ex2_approx(e3);
            float g0, g1, g2, g3;
            f2_unpack(f2_mul(f2_pack(e0, e1), f2_fma(f2_pack(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), sc2,
                                                     f2_pack(d4.x, d4.y))), g0, g1);
            f2_unpack(f2_mul(f2_pack(e2, e3), f2_fma(f2_pack(__uint_as_float(dv[i + 2]), __uint_as_float(dv[i + 3])), sc2,
                                                     f2_pack(d4.z, d4.w))), g2, g3);
            pk[c2 * 8 + (i >> 1)] = pack_bf16x2(e0, e1);
            pk[c2 * 8 + (i >> 1) + 1] = pack_bf16x2(e2, e3);
            dk[c2 * 8 + (i >> 1)] = pack_bf16x2(g0, g1);
            dk[c2 * 8 + (i >> 1) + 1] = pack_bf16x2(g2, g3);
          }
        }
      } else {
        {
          const int cc = hf;
          uint32_t sv[32], dv[32];
          tmem_ld_x32(t_row + COL_ST + cc * 32, sv);
          tmem_ld_x32(t_row + COL_DPT + cc * 32, dv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e[2], d[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int col = cc * 32 + i + k;
              const int qi = q0 + col;
              float s = __uint_as_float(sv[i + k]) * p.scale_log2;
              if (p.bias && qi < p.tq && kvi < tk) s += p.bias[((long long)head * p.tq + qi) * p.tk + kvi] * kLog2e;
              const bool ok = (qi < p.tq) && (kvi < tk) && (!p.causal || kvi <= qi + (p.tk - p.tq));
              e[k] = ok ? ex2_approx(s + st[col]) : 0.f;
              // dropout (mask regenerated): dP_eff = keep ? dP / (1 - p) : 0, and dV sees P_drop = keep ? P / (1 - p) : 0
              const uint32_t bits = drop_bits(dkey, (drop_bh + static_cast<uint32_t>(qi)) * drop_hp + static_cast<uint32_t>(kvi >> 1));
              const bool keep = (kvi & 1) ? drop_keep_hi(dkey, bits) : drop_keep_lo(dkey, bits);
              d[k] = e[k] * fmaf(__uint_as_float(dv[i + k]), keep ? p.scale * dkey.scale : 0.f, st[64 + col]);
              e[k] = keep ? e[k] * dkey.scale : 0.f;
            }
            pk[(i >> 1)] = pack_bf16x2(e[0], e[1]);
            dk[(i >> 1)] = pack_bf16x2(d[0], d[1]);
          }
        }
      }
      // each half writes its 32 bf16 (16 columns) over the first columns of ITS OWN 32 fp32 scores, so the two threads
      // of a row never touch each other's columns; the gradient MMAs address the two K halves separately
      tmem_st_x16(t_row + COL_ST + 32 * hf, pk);     // P^T  of queries [32 hf, 32 hf + 32)
      tmem_st_x16(t_row + COL_DPT + 32 * hf, dk);    // dS^T of the same queries
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(&bars[B_PDSFULL]);
    }
    mbar_wait(&bars[B_DONE], 0);
    tc_fence_after_sync();
    const bool valid = kvi < p.tk;
    // the two halves share the epilogue: half 0 stores the dV row, half 1 the dK row
    bf16* dst = (hf == 0 ? p.dv + (long long)b * p.dv_batch_stride + (long long)kvi * p.dv_row_stride
                         : p.dk + (long long)b * p.dk_batch_stride + (long long)kvi * p.dk_row_stride) + head * D;
    if (n_iter == 0) {
      if (valid)
        for (int i = 0; i < D; i += 8) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
    } else {
      store_out_row(dst, t_row + (hf == 0 ? COL_DV : COL_DK), valid, kvi < tk);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 256);
}

namespace dq2 {
constexpr int OFF_Q = 0, OFF_DO = TILE, OFF_K = 2 * TILE, OFF_V = OFF_K + NST * SUBTILE;
constexpr int OFF_BAR = OFF_V + NST * SUBTILE;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int COL_QA = 0, COL_DOA = 32, COL_S = 64, COL_DP = 128, COL_DQ = 192;   // dS aliases S
enum { B_Q = 0, B_QT = 1, B_KFULL = 2, B_KEMPTY = B_KFULL + NST, B_SFULL = B_KEMPTY + NST, B_DSFULL = B_SFULL + 1,
       B_DONE = B_DSFULL + 1, B_COUNT = B_DONE + 1 };
}  // namespace dq2

template <bool DROP>
__global__ void __launch_bounds__(BWD2_THREADS, 2)
attn_bwd_dq2_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                    const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mdo,
                    const BwdParams p) {
  pdl_trigger();
  pdl_wait();
  using namespace dq2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y, b = blockIdx.z;
  const int tk = effective_tk(p, b);
  int n_iter = (tk + SUB - 1) / SUB;
  if (p.causal) {
    const int last_col = q0 + 127 + (p.tk - p.tq);
    int nc = last_col / SUB + 1;
    if (nc < 1) nc = 1;
    if (nc < n_iter) n_iter = nc;
  }

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(&bars[B_Q], 1);
    mbar_init(&bars[B_QT], 256);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars[B_KFULL + s], 1);
      mbar_init(&bars[B_KEMPTY + s], 1);
    }
    mbar_init(&bars[B_SFULL], 1);
    mbar_init(&bars[B_DSFULL], 256);
    mbar_init(&bars[B_DONE], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 8 && lane == 0) {   // first loads go out before the TMEM allocation (see the dK/dV kernel)
    mbar_expect_tx(&bars[B_Q], 2 * TILE);
    tma_load_4d(smem + OFF_Q, &mq, &bars[B_Q], 0, q0, head, b);
    tma_load_4d(smem + OFF_DO, &mdo, &bars[B_Q], 0, q0, head, b);
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 8) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % NST;
        const uint32_t par = (it / NST) & 1;
        mbar_wait(&bars[B_KEMPTY + st], par ^ 1);
        mbar_expect_tx(&bars[B_KFULL + st], 2 * SUBTILE);
        tma_load_4d(smem + OFF_K + st * SUBTILE, &mk, &bars[B_KFULL + st], 0, it * SUB, head, b);
        tma_load_4d(smem + OFF_V + st * SUBTILE, &mv, &bars[B_KFULL + st], 0, it * SUB, head, b);
      }
    }
  } else if (warp == 9) {
    // warp-uniform loop, one elected lane issues (see the dK/dV kernel)
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, SUB, false, false);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);
    mbar_wait(&bars[B_QT], 0);    // Q, dO sit in TMEM as A operands
    tc_fence_after_sync();
    for (int it = 0; it < n_iter; ++it) {
      const int st = it % NST;
      mbar_wait(&bars[B_KFULL + st], (it / NST) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_tA_x_subT(tmem_base + COL_S, tmem_base + COL_QA, sbase + OFF_K + st * SUBTILE, idesc_s);
        mma_tA_x_subT(tmem_base + COL_DP, tmem_base + COL_DOA, sbase + OFF_V + st * SUBTILE, idesc_s);
        umma_commit(&bars[B_SFULL]);
      }
      __syncwarp();
      mbar_wait(&bars[B_DSFULL], it & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_tA2_x_sub(tmem_base + COL_DQ, tmem_base + COL_S, sbase + OFF_K + st * SUBTILE, idesc_o, it > 0);
        umma_commit(&bars[B_KEMPTY + st]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bars[B_DONE]);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int hf = warp >> 2;       // column half of the 64-wide sub-tile this thread owns: keys [32 hf, 32 hf + 32)
    const int r = q * 32 + lane;
    const int qi = q0 + r;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float lse2 = 0.f, delta = 0.f;
    uint4 orow[8];   // this thread's row of the forward output O (for delta = rowsum(dO * O))
    if (qi < p.tq) {
      const long long idx = ((long long)b * p.heads + head) * p.tq + qi;
      lse2 = p.lse[idx] * kLog2e;
      const uint4* po = reinterpret_cast<const uint4*>(p.o + (long long)b * p.o_batch_stride + (long long)qi * p.o_row_stride + head * D);
#pragma unroll
      for (int c = 0; c < 8; ++c) orow[c] = po[c];
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) orow[c] = make_uint4(0u, 0u, 0u, 0u);
    }
    const DropKey dkey = drop_key(p.drop_state, p.drop_call, p.drop_p);   // thresh 0: off
    const uint32_t drop_row = static_cast<uint32_t>(((long long)b * p.heads + head) * p.tq + qi) * static_cast<uint32_t>((p.tk + 1) >> 1);
    const bool plain = !p.causal && p.bias == nullptr;
    const bool lean = plain && dkey.thresh == 0u;           // packed-pair path of the deterministic step
    const bool lean_drop = DROP && plain && !lean;          // the same plus the regenerated mask
    const uint32_t drop_thi = dkey.thresh << 16;
    const int causal_lim = p.causal ? qi + (p.tk - p.tq) : 0x7fffffff;
    // stationary operands -> TMEM
    mbar_wait(&bars[B_Q], 0);
    {  // delta from the dO row sitting in shared memory (128B-swizzled tile) and the O row in registers
      const uint8_t* drow = smem + OFF_DO + r * 128;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 dd = *reinterpret_cast<const uint4*>(drow + ((c ^ (r & 7)) << 4));
        const uint4 oo = orow[c];
        acc += bf16_lo(dd.x) * bf16_lo(oo.x) + bf16_hi(dd.x) * bf16_hi(oo.x) + bf16_lo(dd.y) * bf16_lo(oo.y) +
               bf16_hi(dd.y) * bf16_hi(oo.y) + bf16_lo(dd.z) * bf16_lo(oo.z) + bf16_hi(dd.z) * bf16_hi(oo.z) +
               bf16_lo(dd.w) * bf16_lo(oo.w) + bf16_hi(dd.w) * bf16_hi(oo.w);
      }
      if (hf == 0 && qi < p.tq) p.delta_out[((long long)b * p.heads + head) * p.tq + qi] = acc;
      delta = acc * p.scale;
    }
    // the two threads of a row share the copy of the stationary operands into TMEM
    if (hf == 0) smem_row_to_tmem(smem + OFF_Q, r, t_row + COL_QA);
    else smem_row_to_tmem(smem + OFF_DO, r, t_row + COL_DOA);
    tmem_st_wait();
    tc_fence_before_sync();
    mbar_arrive(&bars[B_QT]);
    for (int it = 0; it < n_iter; ++it) {
      const int k0 = it * SUB;
      mbar_wait(&bars[B_SFULL], it & 1);
      tc_fence_after_sync();
      uint32_t dk[16];   // dS of this thread's 32 keys, bf16 pairs
      if (lean_drop) {
        // the plain path below plus the regenerated mask; this thread owns ONE query, so the two keys of a column pair are
        // the two 16-bit halves of one hash
        const f32x2 sl2 = f2_rep(p.scale_log2), nl2 = f2_rep(-lse2), nd2 = f2_rep(-delta);
        const float mk = p.scale * dkey.scale;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int c = 2 * hf + c2;
          uint32_t sv[16], dv[16];
          tmem_ld_x16(t_row + COL_S + c * 16, sv);
          tmem_ld_x16(t_row + COL_DP + c * 16, dv);
          tmem_ld_wait();
          const uint32_t pair0 = drop_row + static_cast<uint32_t>((k0 + c * 16) >> 1);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const uint32_t bits = drop_bits(dkey, pair0 + static_cast<uint32_t>(i >> 1));
            const float m0 = ((bits << 16) >= drop_thi) ? mk : 0.f;   // even key: low half of the hash
            const float m1 = (bits >= drop_thi) ? mk : 0.f;           // odd key: high half
            float e0, e1, g0, g1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2, nl2), e0, e1);
            e0 = ex2_approx(e0), e1 = ex2_approx(e1);
            f2_unpack(f2_mul(f2_pack(e0, e1), f2_fma(f2_pack(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])),
                                                     f2_pack(m0, m1), nd2)), g0, g1);
            dk[c2 * 8 + (i >> 1)] = pack_bf16x2(g0, g1);
          }
        }
      } else if (lean) {
        const f32x2 sl2 = f2_rep(p.scale_log2), sc2 = f2_rep(p.scale), nl2 = f2_rep(-lse2), nd2 = f2_rep(-delta);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int c = 2 * hf + c2;
          uint32_t sv[16], dv[16];
          tmem_ld_x16(t_row + COL_S + c * 16, sv);
          tmem_ld_x16(t_row + COL_DP + c * 16, dv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float e0, e1, g0, g1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2, nl2), e0, e1);
            e0 = ex2_approx(e0), e1 = ex2_approx(e1);
            f2_unpack(f2_mul(f2_pack(e0, e1), f2_fma(f2_pack(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), sc2, nd2)),
                      g0, g1);
            dk[c2 * 8 + (i >> 1)] = pack_bf16x2(g0, g1);
          }
        }
      } else {
        {
          const int cc = hf;
          uint32_t sv[32], dv[32];
          tmem_ld_x32(t_row + COL_S + cc * 32, sv);
          tmem_ld_x32(t_row + COL_DP + cc * 32, dv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float d[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int kvi = k0 + cc * 32 + i + k;
              float s = __uint_as_float(sv[i + k]) * p.scale_log2;
              if (p.bias && qi < p.tq && kvi < tk) s += p.bias[((long long)head * p.tq + qi) * p.tk + kvi] * kLog2e;
              const bool ok = (qi < p.tq) && (kvi < tk) && (kvi <= causal_lim);
              const float e = ok ? ex2_approx(s - lse2) : 0.f;
              const uint32_t bits = drop_bits(dkey, drop_row + static_cast<uint32_t>(kvi >> 1));
              const bool keep = (kvi & 1) ? drop_keep_hi(dkey, bits) : drop_keep_lo(dkey, bits);
              d[k] = e * fmaf(__uint_as_float(dv[i + k]), keep ? p.scale * dkey.scale : 0.f, -delta);
              if (p.dbias && ok) atomicAdd(p.dbias + ((long long)head * p.tq + qi) * p.tk + kvi, d[k] * p.inv_scale);
            }
            dk[(i >> 1)] = pack_bf16x2(d[0], d[1]);
          }
        }
      }
      tmem_st_x16(t_row + COL_S + 32 * hf, dk);   // dS of keys [32 hf, 32 hf + 32) over the first columns of their own scores
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(&bars[B_DSFULL]);
    }
    mbar_wait(&bars[B_DONE], 0);
    tc_fence_after_sync();
    bf16* dst = p.dq + (long long)b * p.dq_batch_stride + (long long)qi * p.dq_row_stride + head * D + 32 * hf;
    store_out_half(dst, t_row + COL_DQ + 32 * hf, qi < p.tq);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 256);
}

}  // namespace attn
}  // namespace smx

extern "C" int smx_attn_bwd(const SmxAttn* a, void* stream) {
  using namespace smx;
  using namespace smx::attn;
  SMX_REQUIRE(a && a->q && a->k && a->v && a->o && a->d_o && a->dq && a->dk && a->dv && a->lse && a->delta,
              "attn_bwd: null pointer");
  SMX_REQUIRE(a->dbias == nullptr || a->bias != nullptr, "attn_bwd: dbias needs the additive bias it differentiates");
  SMX_REQUIRE(a->kv_len == nullptr || !a->causal, "attn_bwd: kv_len is not combined with causal masking");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mq, mk, mv, mdo;
  if (make_head_map(&mq, a->q, a->tq, a->heads, a->batch, a->q_row_stride, a->q_batch_stride)) return -1;
  if (make_head_map(&mk, a->k, a->tk, a->heads, a->batch, a->k_row_stride, a->k_batch_stride)) return -1;
  if (make_head_map(&mv, a->v, a->tk, a->heads, a->batch, a->v_row_stride, a->v_batch_stride)) return -1;
  if (make_head_map(&mdo, a->d_o, a->tq, a->heads, a->batch, a->do_row_stride, a->do_batch_stride)) return -1;
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
  p.kv_len = a->kv_len;
  p.drop_state = reinterpret_cast<const unsigned long long*>(a->dropout_state);
  p.drop_call = a->dropout_call;
  p.drop_p = a->dropout_state ? a->dropout_p : 0.0f;
  static bool attr_set = false;
  if (!attr_set) {
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kv2::SMEM_BYTES));
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq2::SMEM_BYTES));
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kv2::SMEM_BYTES));
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq2::SMEM_BYTES));
    attr_set = true;
  }
  dim3 gkv((a->tk + 127) / 128, a->heads, a->batch);
  dim3 gq((a->tq + 127) / 128, a->heads, a->batch);
  p.o = reinterpret_cast<const bf16*>(a->o);
  p.o_row_stride = a->o_row_stride, p.o_batch_stride = a->o_batch_stride;
  p.delta_out = a->delta;
  // the streamed operands arrive as 64-row sub-tiles
  CUtensorMap sq, sk, sv, sdo;
  if (make_head_map_rows(&sq, a->q, a->tq, a->heads, a->batch, a->q_row_stride, a->q_batch_stride, SUB)) return -1;
  if (make_head_map_rows(&sdo, a->d_o, a->tq, a->heads, a->batch, a->do_row_stride, a->do_batch_stride, SUB)) return -1;
  if (make_head_map_rows(&sk, a->k, a->tk, a->heads, a->batch, a->k_row_stride, a->k_batch_stride, SUB)) return -1;
  if (make_head_map_rows(&sv, a->v, a->tk, a->heads, a->batch, a->v_row_stride, a->v_batch_stride, SUB)) return -1;
  // the dQ kernel also produces delta = rowsum(dO * O) (its dO tile is already in shared memory), so it runs first
  const bool drop = p.drop_state != nullptr && p.drop_p > 0.0f;
  if (drop) launch_pdl(attn_bwd_dq2_kernel<true>, dim3(gq), dim3(BWD2_THREADS), dq2::SMEM_BYTES, st, mq, sk, sv, mdo, p);
  else launch_pdl(attn_bwd_dq2_kernel<false>, dim3(gq), dim3(BWD2_THREADS), dq2::SMEM_BYTES, st, mq, sk, sv, mdo, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  if (drop) launch_pdl(attn_bwd_dkv2_kernel<true>, dim3(gkv), dim3(BWD2_THREADS), kv2::SMEM_BYTES, st, sq, mk, mv, sdo, p);
  else launch_pdl(attn_bwd_dkv2_kernel<false>, dim3(gkv), dim3(BWD2_THREADS), kv2::SMEM_BYTES, st, sq, mk, mv, sdo, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
