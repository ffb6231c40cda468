// Flash-style attention forward, head_dim 64, plain (non-causal, no additive bias) case -- the speech-encoder and
// text-encoder self-attention of the SpeechMix path
// (hf:models/wav2vec2/modeling_wav2vec2.py:466-549 called from ref:speechmix/hf_model.py:397).
//
// One CTA = one 128-row query tile of one (batch, head); TWO CTAs co-reside per SM (256 TMEM columns, ~84 KiB of
// shared memory, 320 threads each), so the load / allocate / drain phases of one CTA and the waits of its softmax
// warps for the tensor pipe hide behind the other CTA's exponentials.
//   warps 0..7    softmax: two groups of 128 threads = key half h of the 128-key tile.  A thread owns one row of one
//                 64-key half: it pulls its 64 scores into registers with one batch of tcgen05.ld (single pass: row
//                 maximum, then exponentials) and writes P back to TMEM as bf16 OVER the scores it came from
//                 (FA4-style aliasing) -- probabilities never touch shared memory.  The two halves of a row keep
//                 SEPARATE running maxima, row sums and output accumulators (like a split-KV decode) and are merged once
//                 at the end, so nothing is exchanged between threads inside the key loop and four softmax warps
//                 share every scheduler (the single-warp-per-scheduler versions were latency-, not MUFU-bound).
//   warp 8        TMA producer: Q once, K / V tiles through a 2-stage ring (128-byte swizzle)
//   warp 9        tcgen05.mma issuer:  S = Q.K^T (128x128x64 -> 128 TMEM columns),
//                 O_h += P_h.V[64 h ..]  (128x64x64, P read from TMEM as the A operand).  The whole warp runs the loop
//                 (uniform registers feed the descriptors); one elected lane issues -- with the loop under
//                 `if (lane == 0)` every MMA cost ~80 clk of issue (measured with smx_debug_attn_trace).
// The running outputs O_h stay in TMEM across key tiles (accumulating MMAs).  Their rescale is LAZY: the reference
// maximum used for the exponentials is only moved when the true maximum grew by more than 2^8 since the last move, so
// O is touched by the softmax warps on a handful of tiles instead of all of them (probabilities stay <= 256, exact in
// the fp32 row sum and harmless in bf16); LSE = m + log2(l) is exact whatever the staleness.
// Issue order per key tile j:  P_0.V, P_1.V, S(j+1) -- the in-order tensor pipe makes S(j+1) overwrite P(j) only after
// both products have consumed it.
// A share of the exponentials can run on the FMA pipe (Cody-Waite + cubic, ~1 bf16 ulp): at head_dim 64 the
// 16 ex2/clk/SM of the MUFU unit bound the arithmetic (profiles/r01q_micro_mufu.txt).
// Optional per-sample key counts (kv_len): tiles past the count are never loaded, the tile it cuts is masked.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

#include <stdlib.h>
#include <string.h>

namespace smx {
namespace attn {

int make_head_map(CUtensorMap* m, const void* ptr, int t, int heads, int batch, long long row_stride,
                  long long batch_stride);

namespace f2 {

constexpr int BQ = 128, BKV = 128, D = 64, HALF = 64;
constexpr int TILE_BYTES = 128 * D * 2;   // 16 KiB: a [128 x 64] bf16 tile
constexpr int NST = 2;                    // K / V ring depth
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;
constexpr int OFF_V = OFF_K + NST * TILE_BYTES;
constexpr int OFF_ML = OFF_V + NST * TILE_BYTES;    // [2 halves][128 rows] {m, l}: merge of the key halves
constexpr int OFF_BAR = OFF_ML + 2 * 128 * 8;       // 80 KiB + 2 KiB
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int NUM_THREADS = 320;          // 8 softmax warps, producer, MMA
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0;      // scores (fp32, 128 columns); P of key half h (bf16) over columns [64 h, 64 h + 32)
constexpr int COL_O = 128;    // + 64 h : running output over key half h (fp32)
constexpr float kRescaleThreshold = 8.0f;   // log2 domain: move the reference maximum when it grew by > 2^8

enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = B_KFULL + NST, B_VFULL = B_KEMPTY + NST, B_VEMPTY = B_VFULL + NST,
       B_SFULL = B_VEMPTY + NST, B_PFULL = B_SFULL + 1, B_ODONE = B_PFULL + 2, B_COUNT = B_ODONE + 2 };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier block");

struct Params {
  bf16* o;
  float* lse;
  long long o_row_stride, o_batch_stride;
  int batch, heads, tq, tk;
  float scale_log2;  // scale * log2(e)
  const int* kv_len;
  const unsigned long long* drop_state;   // dropout on the probabilities: {seed, step} on the device, call index, p
  uint32_t drop_call;
  float drop_p;
  unsigned long long* trace;   // debug: SM clock stamps of CTA (0,0,0) -- [role][64] (smx_debug_attn_trace)
};

// role 0: MMA warp, 1: softmax group of key half 0, 2: softmax group of key half 1
#define F2_TRACE(role, cond)                                                                  \
  do {                                                                                        \
    if (p.trace != nullptr && (cond) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tr_n < 64) \
      p.trace[(role) * 64 + tr_n++] = clock64();                                              \
  } while (0)

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): x = n + f with n = floor(x), 2^f by a cubic (max rel. error
// 8.8e-5, below the bf16 rounding of the probability it feeds), 2^n through the exponent field.  Packed pairs.
__device__ __forceinline__ void ex2_poly2(float& a, float& b) {
  a = fmaxf(a, -126.0f), b = fmaxf(b, -126.0f);
  const f32x2 x = f2_pack(a, b);
  f32x2 n;   // floor(x): add the 1.5 * 2^23 magic constant with round-towards-minus-infinity, subtract it back
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(n) : "l"(x), "l"(f2_rep(12582912.0f)));
  float n0, n1;
  f2_unpack(n, n0, n1);
  const f32x2 nf = f2_add(n, f2_rep(-12582912.0f));
  const f32x2 f = f2_fma(nf, f2_rep(-1.0f), x);   // in [0, 1)
  f32x2 p = f2_fma(f, f2_rep(0.07711909f), f2_rep(0.22756439f));
  p = f2_fma(p, f, f2_rep(0.69514614f));
  p = f2_fma(p, f, f2_rep(1.0f));
  float p0, p1;
  f2_unpack(p, p0, p1);
  // the low mantissa bits of the magic sum hold n (two's complement): shift them into the exponent field
  a = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(n0) << 23));
  b = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(n1) << 23));
}

// Lean shared-memory descriptors (128-byte swizzle, SBO 1024 B): the high word is a constant, the low word is
// (address >> 4) | (LBO >> 4) << 16 -- one 32-bit add per K step on the issuing warp's (uniform) datapath.
constexpr uint32_t kDescHi = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (static_cast<uint32_t>(kLayoutSW128) << 29);
__device__ __forceinline__ uint64_t lean_desc(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

// DROP: dropout on the probabilities -- P_drop = keep ? P / (1 - p) : 0 feeds P.V, the row sum stays that of P
template <int POLY_EVERY, bool DROP>   // every POLY_EVERY-th pair of exponentials is evaluated on the FMA pipe (0: none)
__global__ void __launch_bounds__(NUM_THREADS, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tq_map, const __grid_constant__ CUtensorMap tk_map,
                 const __grid_constant__ CUtensorMap tv_map, const Params p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y, b = blockIdx.z;
  int tk = p.tk;
  if (p.kv_len) {
    const int l = p.kv_len[b];
    tk = l < 1 ? 1 : (l < p.tk ? l : p.tk);
  }
  const int n_tiles = (tk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("smx attn_fwd2: dynamic smem not 1024-aligned\n");
      __trap();
    }
    mbar_init(&bars[B_QFULL], 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars[B_KFULL + s], 1);
      mbar_init(&bars[B_KEMPTY + s], 1);
      mbar_init(&bars[B_VFULL + s], 1);
      mbar_init(&bars[B_VEMPTY + s], 1);
    }
    mbar_init(&bars[B_SFULL], 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars[B_PFULL + g], 128);
      mbar_init(&bars[B_ODONE + g], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tq_map);
    tma_prefetch_desc(&tk_map);
    tma_prefetch_desc(&tv_map);
  }
  __syncthreads();
  if (warp == 8 && lane == 0) {   // the first loads go out before the TMEM allocation (the co-resident CTA may still hold columns)
    mbar_expect_tx(&bars[B_QFULL], TILE_BYTES);
    tma_load_4d(smem + OFF_Q, &tq_map, &bars[B_QFULL], 0, q0, head, b);
    mbar_expect_tx(&bars[B_KFULL], TILE_BYTES);
    tma_load_4d(smem + OFF_K, &tk_map, &bars[B_KFULL], 0, 0, head, b);
    mbar_expect_tx(&bars[B_VFULL], TILE_BYTES);
    tma_load_4d(smem + OFF_V, &tv_map, &bars[B_VFULL], 0, 0, head, b);
  }
  if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      for (int j = 1; j < n_tiles; ++j) {
        const int slot = j % NST;
        const uint32_t par = ((j / NST) & 1) ^ 1;
        mbar_wait(&bars[B_KEMPTY + slot], par);
        mbar_expect_tx(&bars[B_KFULL + slot], TILE_BYTES);
        tma_load_4d(smem + OFF_K + slot * TILE_BYTES, &tk_map, &bars[B_KFULL + slot], 0, j * BKV, head, b);
        mbar_wait(&bars[B_VEMPTY + slot], par);
        mbar_expect_tx(&bars[B_VFULL + slot], TILE_BYTES);
        tma_load_4d(smem + OFF_V + slot * TILE_BYTES, &tv_map, &bars[B_VFULL + slot], 0, j * BKV, head, b);
      }
    }
  } else if (warp == 9) {
    // The WHOLE warp runs this loop (waits, address arithmetic: warp-uniform, so the operands of tcgen05.mma sit in
    // uniform registers); only the issue itself is predicated on one elected lane.
    constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, false, true);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t q_lo = ((sbase + OFF_Q) >> 4) | (1u << 16);                                  // K-major, LBO 16 B
    const uint32_t k_lo = ((sbase + OFF_K) >> 4) | (1u << 16);
    const uint32_t v_lo = ((sbase + OFF_V) >> 4) | ((static_cast<uint32_t>(TILE_BYTES) >> 4) << 16);   // MN-major
    int tr_n = 0;
    auto issue_s = [&](int slot) {
      if (elect_one()) {
        const uint32_t bb = k_lo + slot * (TILE_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_ss(tmem_base + COL_S, lean_desc(q_lo + kk * 2), lean_desc(bb + kk * 2), idesc_qk, kk > 0 ? 1u : 0u);
        umma_commit(&bars[B_SFULL]);
        umma_commit(&bars[B_KEMPTY + slot]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int h, int slot, uint32_t acc) {   // O_h (+)= P_h . V[64 h .. 64 h + 63]
      if (elect_one()) {
        const uint32_t bb = v_lo + slot * (TILE_BYTES >> 4) + h * (4 * 2048 >> 4);
        const uint32_t d = tmem_base + COL_O + h * D, a = tmem_base + COL_S + h * HALF;
        umma_ts(d, a, lean_desc(bb), idesc_pv, acc);
#pragma unroll
        for (int kk = 1; kk < HALF / 16; ++kk) umma_ts(d, a + kk * 8, lean_desc(bb + kk * 128), idesc_pv, 1u);
        umma_commit(&bars[B_ODONE + h]);
      }
      __syncwarp();
    };
    mbar_wait(&bars[B_QFULL], 0);
    mbar_wait(&bars[B_KFULL + 0], 0);
    tc_fence_after_sync();
    F2_TRACE(0, lane == 0);
    issue_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int slot = j % NST, nslot = (j + 1) % NST;
      const bool more = j + 1 < n_tiles;
      const uint32_t acc = j > 0 ? 1u : 0u;
      mbar_wait(&bars[B_VFULL + slot], (j / NST) & 1);
      F2_TRACE(0, lane == 0);      // start waiting for P_0(j)
      mbar_wait(&bars[B_PFULL + 0], j & 1);
      tc_fence_after_sync();
      F2_TRACE(0, lane == 0);      // P_0(j) seen
      issue_pv(0, slot, acc);
      mbar_wait(&bars[B_PFULL + 1], j & 1);
      tc_fence_after_sync();
      issue_pv(1, slot, acc);
      if (elect_one()) umma_commit(&bars[B_VEMPTY + slot]);
      __syncwarp();
      F2_TRACE(0, lane == 0);      // both P.V issued
      if (more) {
        mbar_wait(&bars[B_KFULL + nslot], ((j + 1) / NST) & 1);
        tc_fence_after_sync();
        issue_s(nslot);
        F2_TRACE(0, lane == 0);    // S(j+1) issued
      }
    }
  } else {
    const int h = warp >> 2;                // key half of this softmax group
    const int q = warp & 3;
    const int r = q * 32 + lane;            // row inside the tile == TMEM lane
    const int row = q0 + r;                 // query index
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_s = t_lane + COL_S + h * HALF;
    const uint32_t t_o = t_lane + COL_O + h * D;
    const f32x2 sc2 = f2_rep(p.scale_log2);
    float m = -INFINITY, l = 0.f;
    int tr_n = 0;
    const bool tr_on = (warp & 3) == 0 && lane == 0;
    const int tr_role = 1 + h;
    const DropKey dkey = DROP ? drop_key(p.drop_state, p.drop_call, p.drop_p) : DropKey{0u, 0u, 1.0f};
    // pairs of keys are numbered per query row: row * ceil(tk / 2) + (k >> 1)  (rows past tq are never stored)
    const uint32_t drop_row = DROP ? static_cast<uint32_t>(((long long)b * p.heads + head) * p.tq + row) * static_cast<uint32_t>((p.tk + 1) >> 1) : 0u;

    for (int j = 0; j < n_tiles; ++j) {
      F2_TRACE(tr_role, tr_on);   // start waiting for S(j)
      mbar_wait(&bars[B_SFULL], j & 1);
      tc_fence_after_sync();
      F2_TRACE(tr_role, tr_on);   // S(j) seen
      uint32_t s[64];
      tmem_ld_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      const int n_valid = tk - j * BKV - h * HALF;   // columns of this half inside the key count (may be <= 0)
      if (n_valid < HALF) {
#pragma unroll
        for (int c = 0; c < HALF; ++c)
          if (c >= n_valid) s[c] = 0xff800000u;   // -inf: exp2 -> 0
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < HALF; c += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[c + 2]), __uint_as_float(s[c + 3]));
        mx2 = fmax3(mx2, __uint_as_float(s[c + 4]), __uint_as_float(s[c + 5]));
        mx3 = fmax3(mx3, __uint_as_float(s[c + 6]), __uint_as_float(s[c + 7]));
      }
      const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;   // scale > 0
      if (j == 0) {
        m = fmaxf(tmax, -1e30f);   // a half without any key (kv_len cut): finite sentinel, its weight in the merge is 0
      } else {
        const bool grow = tmax > m + kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {
          // rare: move the reference maximum of the rows that need it and rescale their running output in TMEM
          const float m_new = grow ? tmax : m;
          const float f = ex2_approx(m - m_new);   // 1 for the rows that keep their reference
          mbar_wait(&bars[B_ODONE + h], (j - 1) & 1);   // P_h.V(j-1) has landed (it precedes S(j) in the pipe)
          tc_fence_after_sync();
#pragma unroll
          for (int c = 0; c < D / 16; ++c) {
            uint32_t v[16];
            tmem_ld_x16(t_o + c * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * f);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                "%14, %15, %16};" ::"r"(t_o + c * 16),
                "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                : "memory");
          }
          l *= f;
          m = m_new;
        }
      }
      // probabilities (<= 2^8), their row sum, bf16 packing -- 64 independent elements: the scheduler is free to
      // interleave the FFMA2 / MUFU / FADD2 / F2FP streams
      const f32x2 nm2 = f2_rep(-m);
      f32x2 la = f2_rep(0.f), lb = f2_rep(0.f);
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < HALF; c += 8) {
        float e[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          f2_unpack(f2_fma(f2_pack(__uint_as_float(s[c + k]), __uint_as_float(s[c + k + 1])), sc2, nm2), e[k], e[k + 1]);
          if (POLY_EVERY > 0 && ((c + k) >> 1) % (POLY_EVERY > 0 ? POLY_EVERY : 1) == POLY_EVERY - 1) {
            ex2_poly2(e[k], e[k + 1]);
          } else {
            e[k] = ex2_approx(e[k]);
            e[k + 1] = ex2_approx(e[k + 1]);
          }
        }
        la = f2_add(la, f2_add(f2_pack(e[0], e[1]), f2_pack(e[2], e[3])));
        lb = f2_add(lb, f2_add(f2_pack(e[4], e[5]), f2_pack(e[6], e[7])));
        if (DROP) {
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            const uint32_t bits = drop_bits(dkey, drop_row + static_cast<uint32_t>((j * BKV + h * HALF + c + k) >> 1));
            e[k] = drop_keep_lo(dkey, bits) ? e[k] : 0.f;
            e[k + 1] = drop_keep_hi(dkey, bits) ? e[k + 1] : 0.f;
          }
        }
        pk[(c >> 1) + 0] = pack_bf16x2(e[0], e[1]);
        pk[(c >> 1) + 1] = pack_bf16x2(e[2], e[3]);
        pk[(c >> 1) + 2] = pack_bf16x2(e[4], e[5]);
        pk[(c >> 1) + 3] = pack_bf16x2(e[6], e[7]);
      }
      F2_TRACE(tr_role, tr_on);   // exponentials done
      tmem_st_x32(t_s, pk);
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(&bars[B_PFULL + h]);
      F2_TRACE(tr_role, tr_on);   // P published
      float l0, l1;
      f2_unpack(f2_add(la, lb), l0, l1);
      l += l0 + l1;
    }
    // ---- merge the two key halves of this row (different reference maxima), normalise, store
    float2* ml = reinterpret_cast<float2*>(smem + OFF_ML);
    ml[h * 128 + r] = make_float2(m, l);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 other = ml[(h ^ 1) * 128 + r];
    const float mm = fmaxf(m, other.x);
    const float w_self = ex2_approx(m - mm), w_other = ex2_approx(other.x - mm);
    const float lt = l * w_self + other.y * w_other;
    const float inv = lt > 0.f ? dkey.scale / lt : 0.f;
    const float wa = (h == 0 ? w_self : w_other) * inv, wb = (h == 0 ? w_other : w_self) * inv;   // weights of O_0, O_1
    mbar_wait(&bars[B_ODONE + 0], (n_tiles - 1) & 1);
    mbar_wait(&bars[B_ODONE + 1], (n_tiles - 1) & 1);
    tc_fence_after_sync();
    uint32_t va[32], vb[32];   // this thread normalises output columns [32 h, 32 h + 32) of its row
    tmem_ld_x32(t_lane + COL_O + 32 * h, va);
    tmem_ld_x32(t_lane + COL_O + D + 32 * h, vb);
    tmem_ld_wait();
    if (row < p.tq) {
      bf16* op = p.o + (long long)b * p.o_batch_stride + (long long)row * p.o_row_stride + head * D + 32 * h;
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(va[k + e]) * wa + __uint_as_float(vb[k + e]) * wb;
        uint4 u;
        u.x = pack_bf16x2(f[0], f[1]), u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]), u.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(op + k) = u;
      }
      if (h == 0 && p.lse) p.lse[((long long)b * p.heads + head) * p.tq + row] = (mm + log2f(lt)) * 0.6931471805599453f;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace f2

static unsigned long long* g_trace = nullptr;
void set_trace(unsigned long long* t) { g_trace = t; }

// host side: launched by smx_attn_fwd (attention_fwd.cu) for the plain case
int launch_fwd2(const SmxAttn* a, cudaStream_t stream) {
  using namespace f2;
  CUtensorMap mq, mk, mv;
  if (make_head_map(&mq, a->q, a->tq, a->heads, a->batch, a->q_row_stride, a->q_batch_stride)) return -1;
  if (make_head_map(&mk, a->k, a->tk, a->heads, a->batch, a->k_row_stride, a->k_batch_stride)) return -1;
  if (make_head_map(&mv, a->v, a->tk, a->heads, a->batch, a->v_row_stride, a->v_batch_stride)) return -1;
  Params p;
  memset(&p, 0, sizeof(p));
  p.o = reinterpret_cast<bf16*>(a->o);
  p.lse = a->lse;
  p.o_row_stride = a->o_row_stride;
  p.o_batch_stride = a->o_batch_stride;
  p.batch = a->batch, p.heads = a->heads, p.tq = a->tq, p.tk = a->tk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.kv_len = a->kv_len;
  p.trace = g_trace;
  const bool drop = a->dropout_state != nullptr && a->dropout_p > 0.0f;
  p.drop_state = reinterpret_cast<const unsigned long long*>(a->dropout_state);
  p.drop_call = a->dropout_call;
  p.drop_p = a->dropout_p;

  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("SMX_ATTN_POLY");
    poly = e ? atoi(e) : 0;   // measured: the kernel is not MUFU-bound yet (profiles/r02_attn.txt), so MUFU-only is the default
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  dim3 grid((a->tq + BQ - 1) / BQ, a->heads, a->batch);
  if (drop)
    launch_pdl(attn_fwd2_kernel<0, true>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, mq, mk, mv, p);
  else if (poly == 0)
    launch_pdl(attn_fwd2_kernel<0, false>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, mq, mk, mv, p);
  else
    launch_pdl(attn_fwd2_kernel<4, false>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, mq, mk, mv, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace attn
}  // namespace smx

// development hook: device buffer of 3 x 64 uint64 that receives SM-clock stamps of CTA (0,0,0) of the next
// attn_fwd2 launches (role-major: MMA thread, softmax group 0, softmax group 1); NULL switches tracing off
extern "C" int smx_debug_attn_trace(void* device_buffer) {
  smx::attn::set_trace(reinterpret_cast<unsigned long long*>(device_buffer));
  return 0;
}
