// Positional convolution embedding of wav2vec2 / HuBERT on tensor cores:
//   grouped Conv1d(H -> H, k = 128, padding 64, groups = 16), drop the last frame, GELU
//   hf:models/wav2vec2/modeling_wav2vec2.py:326-379
//
// Forward / data-gradient kernel: one CTA owns 256 frames of one (batch, group).  The
// input slab (256 + 127 frames x cg channels) is loaded ONCE into shared memory in a
// "channel-chunk major" layout ([cg/8][row][8 channels]); in the un-swizzled UMMA
// K-major layout a window that starts `tap` frames later is just the same descriptor
// with its start address advanced by tap*16 bytes, so the 128 taps are 128 shifted
// MMAs (M=128, N=cg, K=cg) accumulating into one TMEM tile while the per-tap weights
// stream through a 4-stage bulk-copy ring.  Zero padding comes from TMA out-of-bounds
// fill.  Epilogue fuses bias, GELU and the residual add.
//
// Weight-gradient kernel: contraction over frames.  dpre (as A) and the shifted x
// windows (as B) are both read MN-major from the same chunk-major slabs; 8 taps
// accumulate side by side in TMEM, partial sums leave through fp32 atomics.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

#include <string.h>

namespace smx {
namespace posconv {

constexpr int SLAB_ROWS = 384;  // 256 output frames + 127 halo (+1)
constexpr int W_STAGES = 4;
constexpr int NUM_THREADS = 192;

struct FwdParams {
  const bf16* w_packed;  // [G][ksize][cg/8][cg(n)][8]
  const float* bias;
  const bf16* x_res;
  bf16* y;
  bf16* pre_out;
  int t, hidden, cg, ksize, pad, apply_gelu, taps_per_stage;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
posconv_fwd_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap xmap_tail,
                   const FwdParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int cg = p.cg, chunks = cg / 8;
  const int slab_bytes = chunks * SLAB_ROWS * 16;
  const int tap_bytes = cg * cg * 2;
  const int stage_bytes = p.taps_per_stage * tap_bytes;
  uint8_t* slab = smem;
  uint8_t* wst = smem + slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + W_STAGES * stage_bytes);
  uint64_t* slab_full = bars;
  uint64_t* wfull = bars + 1;
  uint64_t* wempty = wfull + W_STAGES;
  uint64_t* acc_full = wempty + W_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * 256, g = blockIdx.y, b = blockIdx.z;
  const int n_stages_total = p.ksize / p.taps_per_stage;

  if (threadIdx.x == 0) {
    mbar_init(slab_full, 1);
    for (int i = 0; i < W_STAGES; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(slab_full, slab_bytes);
      for (int c = 0; c < chunks; ++c) {
        tma_load_4d(slab + c * SLAB_ROWS * 16, &xmap, slab_full, 0, t0 - p.pad, g * chunks + c, b);
        tma_load_4d(slab + c * SLAB_ROWS * 16 + 256 * 16, &xmap_tail, slab_full, 0, t0 - p.pad + 256, g * chunks + c, b);
      }
      const bf16* wg = p.w_packed + (long long)g * p.ksize * cg * cg;
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_stages_total; ++s) {
        mbar_wait(&wempty[stage], phase ^ 1);
        mbar_expect_tx(&wfull[stage], stage_bytes);
        bulk_load_1d(wst + stage * stage_bytes, wg + (long long)s * p.taps_per_stage * cg * cg, stage_bytes,
                     &wfull[stage]);
        if (++stage == W_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, cg, false, false);
      const uint32_t slab_a = smem_u32(slab);
      const uint32_t w_a = smem_u32(wst);
      mbar_wait(slab_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_stages_total; ++s) {
        mbar_wait(&wfull[stage], phase);
        tc_fence_after_sync();
        for (int tl = 0; tl < p.taps_per_stage; ++tl) {
          const int tap = s * p.taps_per_stage + tl;
          for (int kk = 0; kk < cg / 16; ++kk) {
            const uint32_t a0 = slab_a + (2 * kk) * (SLAB_ROWS * 16) + tap * 16;
            const uint32_t bb = w_a + stage * stage_bytes + tl * tap_bytes + (2 * kk) * (cg * 16);
            const uint64_t bd = umma_smem_desc(bb, cg * 16, 128, kLayoutNone);
            const uint32_t accum = (tap > 0 || kk > 0) ? 1u : 0u;
            umma_ss(tmem_base, umma_smem_desc(a0, SLAB_ROWS * 16, 128, kLayoutNone), bd, idesc, accum);
            umma_ss(tmem_base + 64, umma_smem_desc(a0 + 128 * 16, SLAB_ROWS * 16, 128, kLayoutNone), bd, idesc, accum);
          }
        }
        umma_commit(&wempty[stage]);
        if (++stage == W_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    for (int half = 0; half < 2; ++half) {
      const int t = t0 + half * 128 + r;
      const bool ok = t < p.t;
      const long long off = ((long long)b * p.t + t) * p.hidden + g * cg;
      for (int c16 = 0; c16 < cg / 16; ++c16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 64 + c16 * 16, v);
        tmem_ld_wait();
        if (ok) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + (p.bias ? __ldg(p.bias + g * cg + c16 * 16 + i) : 0.f);
          if (p.pre_out) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              uint4 u;
              u.x = pack_bf16x2(f[i], f[i + 1]), u.y = pack_bf16x2(f[i + 2], f[i + 3]);
              u.z = pack_bf16x2(f[i + 4], f[i + 5]), u.w = pack_bf16x2(f[i + 6], f[i + 7]);
              *reinterpret_cast<uint4*>(p.pre_out + off + c16 * 16 + i) = u;
            }
          }
          if (p.apply_gelu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = gelu_erf(f[i]);
          }
          if (p.x_res) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              const uint4 u = *reinterpret_cast<const uint4*>(p.x_res + off + c16 * 16 + i);
              f[i] += bf16_lo(u.x), f[i + 1] += bf16_hi(u.x), f[i + 2] += bf16_lo(u.y), f[i + 3] += bf16_hi(u.y);
              f[i + 4] += bf16_lo(u.z), f[i + 5] += bf16_hi(u.z), f[i + 6] += bf16_lo(u.w), f[i + 7] += bf16_hi(u.w);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            uint4 u;
            u.x = pack_bf16x2(f[i], f[i + 1]), u.y = pack_bf16x2(f[i + 2], f[i + 3]);
            u.z = pack_bf16x2(f[i + 4], f[i + 5]), u.w = pack_bf16x2(f[i + 6], f[i + 7]);
            *reinterpret_cast<uint4*>(p.y + off + c16 * 16 + i) = u;
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------ weight gradient
// dW[g][tap][o][c] = sum_{b,t} dpre[b, t, g cg + o] * x[b, t + tap - pad, g cg + c]     (contraction over frames)
//
// One CTA owns (group g, ONE 8-channel chunk j of the group's input channels, up to 64 taps) and a share of the
// 128-frame tiles.  Both operands are MN-major slabs in the chunk-major layout ([chunk][frame][8 channels], 16 B per
// frame): A = dpre^T (M = output channels, chunks SBO = 2 KiB apart), B = the x chunk.  Because the B slab holds a
// single chunk, "the window of the next tap" is the same slab 16 bytes further on -- so the N dimension of ONE MMA can
// run over (tap, channel-in-chunk) with a stride-dimension offset of 16 B (overlapping core matrices): N = 32 taps x 8
// channels = 256, the full-rate shape, instead of one N = cg MMA per tap (which sat on the ~45-clk floor of a small
// MMA and re-read the 4 KiB A tile for every tap: shared-memory bound at 11 % of the tensor peak).  Two tap blocks
// share the A tile (2 x 256 TMEM columns).  Rows of D past cg are products with the neighbouring groups' channels (A is
// over-read to M = 128) and are never read back.  Partial sums leave through 16-byte fp32 reductions.
constexpr int WG_STAGES = 4;
constexpr int WG_XROWS = 128 + 64;   // 128 frames + up to 64 taps of halo

struct WgParams {
  float* dw;  // [G][ksize][cg(o)][cg(c)]
  int t, cg, ksize, pad, batch, splits;
  int tpm, nblk;   // taps per MMA (16 or 32), tap blocks per CTA (1 or 2)
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
posconv_wgrad_kernel(const __grid_constant__ CUtensorMap dmap, const __grid_constant__ CUtensorMap xmap,
                     const WgParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int cg = p.cg, chunks = cg / 8;
  const int d_bytes = chunks * 128 * 16;
  const int x_bytes = WG_XROWS * 16;
  uint8_t* dsl = smem;                                 // WG_STAGES dpre slabs first (A over-reads stay in smem)
  uint8_t* xsl = smem + WG_STAGES * d_bytes;
  const int bar_off = ((WG_STAGES * (d_bytes + x_bytes) + 1023) / 1024) * 1024;
  const int bar_off2 = bar_off > (WG_STAGES - 1) * d_bytes + 32768 ? bar_off : (WG_STAGES - 1) * d_bytes + 32768;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + bar_off2);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* done = empty + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps_cta = p.tpm * p.nblk;
  const int tap0 = blockIdx.x * taps_cta, g = blockIdx.y / chunks, j = blockIdx.y % chunks, split = blockIdx.z;
  const int tiles_per_batch = (p.t + 127) / 128;
  const int n_tiles = tiles_per_batch * p.batch;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = split; tile < n_tiles; tile += p.splits) {
        const int bb = tile / tiles_per_batch, tt = (tile % tiles_per_batch) * 128;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], d_bytes + x_bytes);
        tma_load_4d(dsl + stage * d_bytes, &dmap, &full[stage], 0, tt, g * chunks, bb);
        tma_load_4d(xsl + stage * x_bytes, &xmap, &full[stage], 0, tt + tap0 - p.pad, g * chunks + j, bb);
        if (++stage == WG_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 8 * p.tpm, true, true);
      const uint32_t d_a = smem_u32(dsl), x_a = smem_u32(xsl);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int tile = split; tile < n_tiles; tile += p.splits) {
        mbar_wait(&full[stage], phase);
        tc_fence_after_sync();
        for (int blk = 0; blk < p.nblk; ++blk) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            // A: MN-major, no swizzle: 8-channel chunks SBO = 2 KiB apart, 8-frame groups LBO = 128 B apart.
            // B: same frame stepping; the "next chunk" of N is the next TAP = the same chunk one frame (16 B) later.
            const uint64_t ad = umma_smem_desc(d_a + stage * d_bytes + kk * 256, 128, 128 * 16, kLayoutNone);
            const uint64_t bd = umma_smem_desc(x_a + stage * x_bytes + (kk * 16 + blk * p.tpm) * 16, 128, 16, kLayoutNone);
            umma_ss(tmem_base + blk * 256, ad, bd, idesc, (!first || kk > 0) ? 1u : 0u);
          }
        }
        first = false;
        umma_commit(&empty[stage]);
        if (++stage == WG_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(done);
    }
  } else {
    const int q = warp & 3;
    const int o = q * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after_sync();
    const bool any = split < n_tiles;
    if (q * 32 < cg) {
      for (int blk = 0; blk < p.nblk; ++blk) {
        for (int t2 = 0; t2 < p.tpm; t2 += 2) {     // 16 columns = 2 taps x 8 input channels
          uint32_t v[16];
          tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * 256 + t2 * 8, v);
          tmem_ld_wait();
          if (o < cg && any) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float* dst = p.dw + (((long long)g * p.ksize + tap0 + blk * p.tpm + t2 + h) * cg + o) * cg + j * 8;
              red_add_v4(dst, __uint_as_float(v[8 * h]), __uint_as_float(v[8 * h + 1]), __uint_as_float(v[8 * h + 2]),
                         __uint_as_float(v[8 * h + 3]));
              red_add_v4(dst + 4, __uint_as_float(v[8 * h + 4]), __uint_as_float(v[8 * h + 5]),
                         __uint_as_float(v[8 * h + 6]), __uint_as_float(v[8 * h + 7]));
            }
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int make_chunk_map(CUtensorMap* m, const void* ptr, int64_t batch, int64_t t, int hidden, int rows_box,
                          int chunks_box) {
  const uint64_t dims[4] = {8, (uint64_t)t, (uint64_t)(hidden / 8), (uint64_t)batch};
  const uint64_t str[3] = {(uint64_t)hidden, 8, (uint64_t)(t * hidden)};
  const uint32_t box[4] = {8, (uint32_t)rows_box, (uint32_t)chunks_box, 1};
  return encode_tmap_bf16(m, ptr, 4, dims, str, box, false);
}

static int fwd_like(const void* x, const void* w_packed, const float* bias, void* y, void* pre_out, int64_t batch,
                    int64_t t, int hidden, int groups, int ksize, int pad, int apply_gelu, const void* residual,
                    cudaStream_t st) {
  const int cg = hidden / groups;
  SMX_REQUIRE(hidden % groups == 0 && cg % 16 == 0 && cg <= 64, "posconv: channels/group %d must be 16..64, multiple of 16", cg);
  SMX_REQUIRE(ksize % 16 == 0 && ksize <= 128, "posconv: kernel size %d unsupported", ksize);
  CUtensorMap xmap, xmap_tail;
  // two row pieces per channel chunk: 256 + 128 rows (TMA boxes are limited to 256 rows)
  if (make_chunk_map(&xmap, x, batch, t, hidden, 256, 1)) return -1;
  if (make_chunk_map(&xmap_tail, x, batch, t, hidden, 128, 1)) return -1;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.w_packed = (const bf16*)w_packed;
  p.bias = bias;
  p.x_res = (const bf16*)residual;
  p.y = (bf16*)y;
  p.pre_out = (bf16*)pre_out;
  p.t = (int)t, p.hidden = hidden, p.cg = cg, p.ksize = ksize, p.pad = pad, p.apply_gelu = apply_gelu;
  int tps = 32768 / (cg * cg * 2);
  int pw = 1;
  while (pw * 2 <= tps && pw * 2 <= 16) pw *= 2;
  while (ksize % pw) pw /= 2;
  p.taps_per_stage = pw;
  const int slab_bytes = (cg / 8) * SLAB_ROWS * 16;
  const int smem_bytes = slab_bytes + W_STAGES * pw * cg * cg * 2 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    SMX_CHECK_CUDA(cudaFuncSetAttribute(posconv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  SMX_REQUIRE(smem_bytes <= 200 * 1024, "posconv: smem budget exceeded");
  dim3 grid((unsigned)ceil_div(t, 256), groups, (unsigned)batch);
  launch_pdl(posconv_fwd_kernel, dim3(grid), dim3(NUM_THREADS), smem_bytes, st, xmap, xmap_tail, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace posconv
}  // namespace smx

extern "C" {

int smx_posconv_fwd(const void* x, const void* w_packed, const float* bias, void* y, void* pre_out, int64_t batch,
                    int64_t t, int hidden, int groups, int ksize, int add_input, void* stream) {
  return smx::posconv::fwd_like(x, w_packed, bias, y, pre_out, batch, t, hidden, groups, ksize, ksize / 2, 1,
                                add_input ? x : nullptr, (cudaStream_t)stream);
}

int smx_posconv_dgrad(const void* dpre, const void* w_packed_t, const void* residual, void* dx, int64_t batch,
                      int64_t t, int hidden, int groups, int ksize, void* stream) {
  return smx::posconv::fwd_like(dpre, w_packed_t, nullptr, dx, nullptr, batch, t, hidden, groups, ksize,
                                ksize / 2 - 1, 0, residual, (cudaStream_t)stream);
}

int smx_posconv_wgrad(const void* dpre, const void* x, float* dw, int64_t batch, int64_t t, int hidden, int groups,
                      int ksize, void* stream) {
  using namespace smx;
  using namespace smx::posconv;
  const int cg = hidden / groups;
  SMX_REQUIRE(hidden % groups == 0 && cg % 16 == 0 && cg <= 64, "posconv_wgrad: channels/group %d unsupported", cg);
  SMX_REQUIRE(ksize % 16 == 0 && ksize <= 128, "posconv_wgrad: kernel size %d unsupported", ksize);
  CUtensorMap dmap, xmap;
  if (make_chunk_map(&dmap, dpre, batch, t, hidden, 128, cg / 8)) return -1;
  if (make_chunk_map(&xmap, x, batch, t, hidden, WG_XROWS, 1)) return -1;
  WgParams p;
  p.dw = dw;
  p.t = (int)t, p.cg = cg, p.ksize = ksize, p.pad = ksize / 2, p.batch = (int)batch;
  p.tpm = (ksize % 32 == 0) ? 32 : 16;
  p.nblk = ((ksize / p.tpm) % 2 == 0) ? 2 : 1;
  const int chunks = cg / 8;
  const int n_tiles = (int)(ceil_div(t, 128) * batch);
  const int base = (ksize / (p.tpm * p.nblk)) * groups * chunks;
  // contraction splits: whole waves of one CTA per SM, per-CTA cost = its tiles + ~3 tiles of prologue / reduction
  int splits = 1;
  double best = 1e30;
  for (int sp = 1; sp <= 16 && sp <= n_tiles; ++sp) {
    const double waves = (double)ceil_div((int64_t)base * sp, num_sms());
    const double cost = waves * ((double)ceil_div(n_tiles, sp) + 3.0);
    if (cost < best - 1e-9) best = cost, splits = sp;
  }
  p.splits = splits;
  const int d_bytes = chunks * 128 * 16, x_bytes = WG_XROWS * 16;
  int bar_off = ((WG_STAGES * (d_bytes + x_bytes) + 1023) / 1024) * 1024;
  const int min_off = (WG_STAGES - 1) * d_bytes + 32768;
  if (bar_off < min_off) bar_off = min_off;
  const int smem_bytes = bar_off + 256;
  SMX_CHECK_CUDA(cudaFuncSetAttribute(posconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid(ksize / (p.tpm * p.nblk), groups * chunks, splits);
  launch_pdl(posconv_wgrad_kernel, dim3(grid), dim3(NUM_THREADS), smem_bytes, (cudaStream_t)stream, dmap, xmap, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}
