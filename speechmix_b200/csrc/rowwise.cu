// HBM-bound row-wise / element-wise kernels: LayerNorm (+RMSNorm) forward and
// backward, bias-gradient column sums, embedding gather/scatter, weighted layer
// sum, small packing helpers.  All accesses are 16-byte vectorised and coalesced
// along the channel dimension; reductions use warp shuffles.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

namespace smx {
namespace rw {

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(u.x), f[1] = bf16_hi(u.x), f[2] = bf16_lo(u.y), f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z), f[5] = bf16_hi(u.z), f[6] = bf16_lo(u.w), f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]), u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]), u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x), f[1] = bf16_hi(u.x), f[2] = bf16_lo(u.y), f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z), f[5] = bf16_hi(u.z), f[6] = bf16_lo(u.w), f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]), u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]), u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void loadf8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}

// ------------------------------------------------------------------ LayerNorm forward (register-prefetch variant)
// Used when there is no residual input (forward) / no residual-branch gradient (backward): measured with ncu at
// [23968, 768] (profiles/r01q_ln_variants.txt) these win for one / two input streams (23.2 vs 24.6 us forward,
// 36.8 vs 41.0 us backward with column sums), the shared-memory-ring kernels further down win with one more stream
// (30.0 vs 34.5 us forward with residual, 39.3 vs 46.9 us backward with a residual-branch gradient).
// one warp per row; VPL = 16-byte vectors per lane (cols <= VPL*256)
template <int VPL>
__global__ void __launch_bounds__(256) ln_fwd_reg_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     bf16* __restrict__ y, bf16* __restrict__ sum_out,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     long long rows, int cols, float eps, int rms_only, int act) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  // the next row's vectors are requested before the current row is reduced (latency-bound stream otherwise)
  uint4 nx[VPL];
  if (warp_global < rows) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) nx[i] = *reinterpret_cast<const uint4*>(x + warp_global * cols + c);
    }
  }
  for (long long row = warp_global; row < rows; row += nwarps) {
    float v[VPL][8];
    float s = 0.f;
    uint4 cur[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) cur[i] = nx[i];
    if (row + nwarps < rows) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < cols) nx[i] = *reinterpret_cast<const uint4*>(x + (row + nwarps) * cols + c);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        unpack_bf16x8(cur[i], v[i]);
        if (res) {
          float r[8];
          load8(res + row * cols + c, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += r[j];
        }
        if (sum_out) {
          store8(sum_out + row * cols + c, v[i]);
          // keep the statistics consistent with what backward will re-read
          float t[8];
          load8(sum_out + row * cols + c, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = t[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
    }
    float mean = 0.f;
    if (!rms_only) mean = warp_sum(s) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        float g[8], b[8], o[8];
        loadf8(gamma + c, g);
        if (beta) loadf8(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + (beta ? b[j] : 0.f);
        if (act == SMX_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = gelu_erf(o[j]);
        }
        store8(y + row * cols + c, o);
      }
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward (register variant)
// One warp per row, rows strided over a persistent grid; parameter gradients (and, optionally, the column sums
// of dx = the bias gradient of the linear layer in front of a post-LN block) accumulate in registers and are
// reduced once per block.  Between the statistics pass and the dx pass a row is held as the PACKED bf16 it was
// loaded as (2 x 4 registers per 8 elements) instead of two fp32 copies, which is what keeps the kernel at
// two resident 256-thread blocks per SM -- an HBM-bound row kernel needs the
// warps to cover the memory latency (the first version: 159 registers, 8 warps per SM, 29 % of HBM peak).
template <int VPL, bool WITH_CS>
__global__ void __launch_bounds__(256, (VPL <= 3 ? 2 : 1))
ln_bwd_reg_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dx_colsum, long long rows, int cols, int rms_only, int act) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][32 * 8 + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;
  float ag[VPL][8], ab[VPL][8], ac[WITH_CS ? VPL : 1][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ag[i][j] = 0.f, ab[i][j] = 0.f;
      if (WITH_CS) ac[i][j] = 0.f;
    }

  for (long long row = warp_global; row < rows; row += nwarps) {
    const float mean = rms_only ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    uint4 dp[VPL], xp[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        dp[i] = *reinterpret_cast<const uint4*>(dy + row * cols + c);
        xp[i] = *reinterpret_cast<const uint4*>(x + row * cols + c);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        float d[8], xh[8], gm[8];
        unpack_bf16x8(dp[i], d);
        unpack_bf16x8(xp[i], xh);
        loadf8(gamma + c, gm);
        if (act == SMX_ACT_GELU) {  // y = gelu(z), z = xhat*gamma + beta: fold gelu'(z) into dy first
          float bt[8];
          loadf8(beta + c, bt);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] *= gelu_erf_grad(fmaf((xh[j] - mean) * rstd, gm[j], bt[j]));
          dp[i] = pack_bf16x8(d);   // the dx pass re-reads the folded gradient
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = (xh[j] - mean) * rstd;
          ag[i][j] = fmaf(d[j], h, ag[i][j]);
          ab[i][j] += d[j];
          const float gj = d[j] * gm[j];
          s1 += gj;
          s2 = fmaf(gj, h, s2);
        }
      }
    }
    s1 = rms_only ? 0.f : warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        float d[8], xh[8], gm[8], o[8];
        unpack_bf16x8(dp[i], d);
        unpack_bf16x8(xp[i], xh);
        loadf8(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (d[j] * gm[j] - s1 - (xh[j] - mean) * rstd * s2);
        if (dres) {
          float r[8];
          load8(dres + row * cols + c, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        store8(dx + row * cols + c, o);
        if (WITH_CS) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ac[i][j] += o[j];
        }
      }
    }
  }
  // block reduction of the column accumulators, then one atomic per column per block
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int pass = 0; pass < (WITH_CS ? 3 : 2); ++pass) {
      float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dx_colsum);
      if (dst == nullptr) continue;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? ag[i][j] : (pass == 1 ? ab[i][j] : ac[WITH_CS ? i : 0][j]);
      __syncthreads();
      const int col_local = threadIdx.x;  // 256 threads, 256 columns of this vector slot
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][col_local];
      const int c = i * 256 + col_local;
      if (c < cols) atomicAdd(dst + c, t);
    }
  }
}

// ------------------------------------------------------------------ post-LN block tail WITH output dropout (training)
// forward:  s = (keep ? x / (1 - p) : 0) + res,  y = LN(s)    -- x = the sub-block's output; replaces the elementwise dropout
//           launch (110 MB at [23968, 768]) + the LayerNorm launch; s is stored, and the statistics are those of the stored
//           (bf16-rounded) values, exactly as the two separate kernels produce them.
// backward: ds = LN'(dy) (the residual-branch gradient),  dsd = keep ? ds / (1 - p) : 0 (the sub-block's output gradient),
//           column sums of dsd (bias gradient of the linear layer in front) -- replaces LayerNorm backward + the dropout
//           launch on ds + the separate column-sum launch.  Same mask numbering as dropout_kernel: pair = element >> 1.
// One warp per row, register variants (see ln_fwd_reg_kernel / ln_bwd_reg_kernel).
template <int VPL>
__global__ void __launch_bounds__(256, (VPL <= 3 ? 3 : 1)) ln_fwd_drop_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      bf16* __restrict__ y, bf16* __restrict__ sum_out,
                                                      float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                      long long rows, int cols, float eps,
                                                      const unsigned long long* __restrict__ state, uint32_t call, float p) {
  pdl_trigger();
  pdl_wait();
  const DropKey key = drop_key(state, call, p);
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  uint4 nx[VPL];
  if (warp_global < rows) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) nx[i] = *reinterpret_cast<const uint4*>(x + warp_global * cols + c);
    }
  }
  for (long long row = warp_global; row < rows; row += nwarps) {
    float v[VPL][8];
    float s = 0.f;
    uint4 cur[VPL], rr[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      cur[i] = nx[i];
      const int c = (i * 32 + lane) * 8;
      if (c < cols) rr[i] = *reinterpret_cast<const uint4*>(res + row * cols + c);   // in flight during the mask arithmetic
    }
    if (row + nwarps < rows) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < cols) nx[i] = *reinterpret_cast<const uint4*>(x + (row + nwarps) * cols + c);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        unpack_bf16x8(cur[i], v[i]);
        const uint32_t pair0 = static_cast<uint32_t>((row * cols + c) >> 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t bits = drop_bits(key, pair0 + j);
          v[i][2 * j] = drop_keep_lo(key, bits) ? v[i][2 * j] * key.scale : 0.f;
          v[i][2 * j + 1] = drop_keep_hi(key, bits) ? v[i][2 * j + 1] * key.scale : 0.f;
        }
        float r[8];
        unpack_bf16x8(rr[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += r[j];
        const uint4 pk = pack_bf16x8(v[i]);
        *reinterpret_cast<uint4*>(sum_out + row * cols + c) = pk;
        unpack_bf16x8(pk, v[i]);   // statistics of what backward will re-read
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
    }
    const float mean = warp_sum(s) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        float g[8], b[8], o[8];
        loadf8(gamma + c, g);
        if (beta) loadf8(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + (beta ? b[j] : 0.f);
        store8(y + row * cols + c, o);
      }
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(256, (VPL <= 3 ? 2 : 1))
ln_bwd_drop_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in, bf16* __restrict__ dx,
                   bf16* __restrict__ dx_drop, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ dxd_colsum, long long rows, int cols, const unsigned long long* __restrict__ state,
                   uint32_t call, float p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][32 * 8 + 1];
  const DropKey key = drop_key(state, call, p);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;
  float ag[VPL][8], ab[VPL][8], ac[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = 0.f, ab[i][j] = 0.f, ac[i][j] = 0.f;

  for (long long row = warp_global; row < rows; row += nwarps) {
    const float mean = mean_in[row];
    const float rstd = rstd_in[row];
    uint4 dp[VPL], xp[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        dp[i] = *reinterpret_cast<const uint4*>(dy + row * cols + c);
        xp[i] = *reinterpret_cast<const uint4*>(x + row * cols + c);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        float d[8], xh[8], gm[8];
        unpack_bf16x8(dp[i], d);
        unpack_bf16x8(xp[i], xh);
        loadf8(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = (xh[j] - mean) * rstd;
          ag[i][j] = fmaf(d[j], h, ag[i][j]);
          ab[i][j] += d[j];
          const float gj = d[j] * gm[j];
          s1 += gj;
          s2 = fmaf(gj, h, s2);
        }
      }
    }
    s1 = warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        float d[8], xh[8], gm[8], o[8];
        unpack_bf16x8(dp[i], d);
        unpack_bf16x8(xp[i], xh);
        loadf8(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (d[j] * gm[j] - s1 - (xh[j] - mean) * rstd * s2);
        const uint4 po = pack_bf16x8(o);
        *reinterpret_cast<uint4*>(dx + row * cols + c) = po;
        unpack_bf16x8(po, o);     // the mask acts on the stored (bf16) gradient, as the separate dropout launch did
        const uint32_t pair0 = static_cast<uint32_t>((row * cols + c) >> 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t bits = drop_bits(key, pair0 + j);
          o[2 * j] = drop_keep_lo(key, bits) ? o[2 * j] * key.scale : 0.f;
          o[2 * j + 1] = drop_keep_hi(key, bits) ? o[2 * j + 1] * key.scale : 0.f;
        }
        const uint4 pd = pack_bf16x8(o);
        *reinterpret_cast<uint4*>(dx_drop + row * cols + c) = pd;
        unpack_bf16x8(pd, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) ac[i][j] += o[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxd_colsum);
      if (dst == nullptr) continue;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? ag[i][j] : (pass == 1 ? ab[i][j] : ac[i][j]);
      __syncthreads();
      const int col_local = threadIdx.x;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][col_local];
      const int c = i * 256 + col_local;
      if (c < cols) atomicAdd(dst + c, t);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm forward / backward (ring variants)
// Row kernels below run their fp32 arithmetic on PACKED PAIRS (FFMA2 / FMUL2 / FADD2, sm100_prims.cuh): the first
// versions were instruction-issue-bound (ln_fwd ~21, ln_bwd ~28 warp instructions per element and lane at ~50 % of
// the HBM roofline); pairs halve every add / multiply / FMA, and the remaining per-element work is the bf16 unpack.
__device__ __forceinline__ void unpack_pairs(const uint4& u, f32x2 (&p)[4]) {
  p[0] = f2_pack(bf16_lo(u.x), bf16_hi(u.x)), p[1] = f2_pack(bf16_lo(u.y), bf16_hi(u.y));
  p[2] = f2_pack(bf16_lo(u.z), bf16_hi(u.z)), p[3] = f2_pack(bf16_lo(u.w), bf16_hi(u.w));
}
__device__ __forceinline__ uint32_t pack_pair(f32x2 p) {
  float lo, hi;
  f2_unpack(p, lo, hi);
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ uint4 pack_pairs(const f32x2 (&p)[4]) {
  return make_uint4(pack_pair(p[0]), pack_pair(p[1]), pack_pair(p[2]), pack_pair(p[3]));
}
__device__ __forceinline__ void loadf_pairs(const float* q, f32x2 (&p)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(q));
  const float4 b = __ldg(reinterpret_cast<const float4*>(q + 4));
  p[0] = f2_pack(a.x, a.y), p[1] = f2_pack(a.z, a.w), p[2] = f2_pack(b.x, b.y), p[3] = f2_pack(b.z, b.w);
}
__device__ __forceinline__ float pair_sum(f32x2 p) {
  float lo, hi;
  f2_unpack(p, lo, hi);
  return lo + hi;
}

// Input rows reach the warps through a per-warp ring of shared-memory slots filled by 1-D bulk copies
// (cp.async.bulk + mbarrier): lane 0 requests the rows `stages` iterations ahead, so ~100 KB per block are in flight
// without holding a single register -- the register-prefetch versions had one 1.5 KB row in flight per warp
// (~24 KB per SM), which is what capped them near 3 TB/s (Little's law at ~1 us of loaded DRAM latency).
struct RowRing {
  uint8_t* slots;      // this warp's slots
  uint64_t* bars;      // this warp's barriers, one per slot
  uint32_t row_bytes;  // bytes of one row of one input
  uint32_t slot_bytes; // n_inputs * row_bytes
  int stages, s;
  uint32_t parity;
};
__device__ __forceinline__ RowRing ring_setup(uint8_t* smem, int warp, int n_warps, int stages, int n_inputs, int cols) {
  RowRing r;
  r.row_bytes = (uint32_t)cols * 2u;
  r.slot_bytes = r.row_bytes * n_inputs;
  r.slots = smem + (size_t)warp * stages * r.slot_bytes;
  r.bars = reinterpret_cast<uint64_t*>(smem + (size_t)n_warps * stages * r.slot_bytes) + warp * stages;
  r.stages = stages, r.s = 0, r.parity = 0;
  return r;
}
// lane 0 only: request one row of up to three inputs into slot s
__device__ __forceinline__ void ring_issue(const RowRing& r, int s, long long row, int cols, const bf16* a, const bf16* b,
                                           const bf16* c) {
  uint8_t* dst = r.slots + (size_t)s * r.slot_bytes;
  mbar_expect_tx(&r.bars[s], r.slot_bytes);
  bulk_load_1d(dst, a + row * cols, r.row_bytes, &r.bars[s]);
  if (b) bulk_load_1d(dst + r.row_bytes, b + row * cols, r.row_bytes, &r.bars[s]);
  if (c) bulk_load_1d(dst + (b ? 2 : 1) * r.row_bytes, c + row * cols, r.row_bytes, &r.bars[s]);
}
__device__ __forceinline__ void ring_prime(const RowRing& r, int lane, long long first_row, long long row_step,
                                           long long rows, int cols, const bf16* a, const bf16* b, const bf16* c) {
  if (lane == 0) {
    for (int s = 0; s < r.stages; ++s) mbar_init(&r.bars[s], 1);
    fence_barrier_init();
    for (int s = 0; s < r.stages; ++s) {
      const long long row = first_row + s * row_step;
      if (row < rows) ring_issue(r, s, row, cols, a, b, c);
    }
  }
  __syncwarp();
}
__device__ __forceinline__ const uint8_t* ring_wait(const RowRing& r) {
  mbar_wait(&r.bars[r.s], r.parity);
  return r.slots + (size_t)r.s * r.slot_bytes;
}
// every lane has finished reading the current slot: refill it with the row `stages` iterations ahead
__device__ __forceinline__ void ring_release(RowRing& r, int lane, long long row, long long row_step, long long rows,
                                             int cols, const bf16* a, const bf16* b, const bf16* c) {
  __syncwarp();
  if (lane == 0) {
    const long long next = row + (long long)r.stages * row_step;
    if (next < rows) {
      fence_proxy_async_smem();   // generic-proxy accesses of the slot are ordered before the async-proxy refill
      ring_issue(r, r.s, next, cols, a, b, c);
    }
  }
  if (++r.s == r.stages) r.s = 0, r.parity ^= 1u;
}

// one warp per row; VPL = 16-byte vectors per lane (cols <= VPL*256)
template <int VPL>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     bf16* __restrict__ y, bf16* __restrict__ sum_out,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     long long rows, int cols, float eps, int rms_only, int act,
                                                     int stages) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;
  const float inv_cols = 1.0f / cols;
  RowRing ring = ring_setup(ring_smem, warp, 8, stages, res ? 2 : 1, cols);
  ring_prime(ring, lane, warp_global, nwarps, rows, cols, x, res, nullptr);

  for (long long row = warp_global; row < rows; row += nwarps) {
    const uint8_t* slot = ring_wait(ring);
    f32x2 v[VPL][4];
    f32x2 s2 = f2_rep(0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        unpack_pairs(*reinterpret_cast<const uint4*>(slot + c * 2), v[i]);
        if (res) {
          f32x2 r[4];
          unpack_pairs(*reinterpret_cast<const uint4*>(slot + ring.row_bytes + c * 2), r);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[i][j] = f2_add(v[i][j], r[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = f2_rep(0.f);
      }
    }
    ring_release(ring, lane, row, nwarps, rows, cols, x, res, nullptr);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        if (sum_out) {
          // keep the statistics consistent with what backward will re-read: round to bf16 first
          const uint4 u = pack_pairs(v[i]);
          *reinterpret_cast<uint4*>(sum_out + row * cols + c) = u;
          unpack_pairs(u, v[i]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) s2 = f2_add(s2, v[i][j]);
      }
    }
    float mean = 0.f;
    if (!rms_only) mean = warp_sum(pair_sum(s2)) * inv_cols;
    const f32x2 nmean = f2_rep(-mean);
    f32x2 sq2 = f2_rep(0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[i][j] = f2_add(v[i][j], nmean);   // v now holds x - mean
          sq2 = f2_fma(v[i][j], v[i][j], sq2);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(pair_sum(sq2)) * inv_cols + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    const f32x2 rstd2 = f2_rep(rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        f32x2 g[4], b[4], o[4];
        loadf_pairs(gamma + c, g);
        if (beta) {
          loadf_pairs(beta + c, b);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = f2_rep(0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = f2_fma(f2_mul(v[i][j], rstd2), g[j], b[j]);
        if (act == SMX_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo, hi;
            f2_unpack(o[j], lo, hi);
            gelu_erf2(lo, hi);
            o[j] = f2_pack(lo, hi);
          }
        }
        *reinterpret_cast<uint4*>(y + row * cols + c) = pack_pairs(o);
      }
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// One warp per row, rows strided over a persistent grid; parameter gradients (and, optionally, the column sums
// of dx = the bias gradient of the linear layer in front of a post-LN block) accumulate in registers and are
// reduced once per block.  dy / x (/ the residual-branch gradient) of a row stay in the warp's ring slot between the
// statistics pass and the dx pass, so neither pass holds a register copy of the row.
template <int VPL, bool WITH_CS>
__global__ void __launch_bounds__(256, (VPL <= 3 ? 2 : 1))
ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dx_colsum, long long rows, int cols, int rms_only, int act, int stages) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t ring_smem[];
  __shared__ float red[8][32 * 8 + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;
  const float inv_cols = 1.0f / cols;
  RowRing ring = ring_setup(ring_smem, warp, 8, stages, dres ? 3 : 2, cols);
  ring_prime(ring, lane, warp_global, nwarps, rows, cols, dy, x, dres);
  f32x2 ag[VPL][4], ab[VPL][4], ac[WITH_CS ? VPL : 1][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ag[i][j] = f2_rep(0.f), ab[i][j] = f2_rep(0.f);
      if (WITH_CS) ac[i][j] = f2_rep(0.f);
    }

  for (long long row = warp_global; row < rows; row += nwarps) {
    const float mean = rms_only ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    const f32x2 rstd2 = f2_rep(rstd), nmr = f2_rep(-mean * rstd);   // xhat = x * rstd - mean * rstd
    uint8_t* slot = const_cast<uint8_t*>(ring_wait(ring));
    f32x2 s1p = f2_rep(0.f), s2p = f2_rep(0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        f32x2 d[4], h[4], gm[4];
        unpack_pairs(*reinterpret_cast<const uint4*>(slot + c * 2), d);
        unpack_pairs(*reinterpret_cast<const uint4*>(slot + ring.row_bytes + c * 2), h);
        loadf_pairs(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = f2_fma(h[j], rstd2, nmr);
        if (act == SMX_ACT_GELU) {  // y = gelu(z), z = xhat*gamma + beta: fold gelu'(z) into dy first
          f32x2 bt[4];
          loadf_pairs(beta + c, bt);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float z0, z1;
            f2_unpack(f2_fma(h[j], gm[j], bt[j]), z0, z1);
            d[j] = f2_mul(d[j], f2_pack(gelu_erf_grad(z0), gelu_erf_grad(z1)));
          }
          *reinterpret_cast<uint4*>(slot + c * 2) = pack_pairs(d);   // the dx pass re-reads the folded gradient
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ag[i][j] = f2_fma(d[j], h[j], ag[i][j]);
          ab[i][j] = f2_add(ab[i][j], d[j]);
          const f32x2 gj = f2_mul(d[j], gm[j]);
          s1p = f2_add(s1p, gj);
          s2p = f2_fma(gj, h[j], s2p);
        }
      }
    }
    const float s1 = rms_only ? 0.f : warp_sum(pair_sum(s1p)) * inv_cols;
    const float s2 = warp_sum(pair_sum(s2p)) * inv_cols;
    // dx = rstd * (dy*gamma - s1 - xhat*s2) = (dy*gamma + xhat*(-s2)) * rstd + (-s1*rstd)
    const f32x2 ns2 = f2_rep(-s2), ns1r = f2_rep(-s1 * rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < cols) {
        f32x2 d[4], h[4], gm[4], o[4];
        unpack_pairs(*reinterpret_cast<const uint4*>(slot + c * 2), d);   // each lane re-reads its own columns
        unpack_pairs(*reinterpret_cast<const uint4*>(slot + ring.row_bytes + c * 2), h);
        loadf_pairs(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h[j] = f2_fma(h[j], rstd2, nmr);
          o[j] = f2_fma(f2_fma(h[j], ns2, f2_mul(d[j], gm[j])), rstd2, ns1r);
        }
        if (dres) {
          f32x2 r[4];
          unpack_pairs(*reinterpret_cast<const uint4*>(slot + 2 * ring.row_bytes + c * 2), r);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = f2_add(o[j], r[j]);
        }
        *reinterpret_cast<uint4*>(dx + row * cols + c) = pack_pairs(o);
        if (WITH_CS) {
#pragma unroll
          for (int j = 0; j < 4; ++j) ac[i][j] = f2_add(ac[i][j], o[j]);
        }
      }
    }
    ring_release(ring, lane, row, nwarps, rows, cols, dy, x, dres);
  }
  // block reduction of the column accumulators, then one atomic per column per block
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int pass = 0; pass < (WITH_CS ? 3 : 2); ++pass) {
      float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dx_colsum);
      if (dst == nullptr) continue;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo, hi;
        f2_unpack(pass == 0 ? ag[i][j] : (pass == 1 ? ab[i][j] : ac[WITH_CS ? i : 0][j]), lo, hi);
        red[warp][lane * 8 + 2 * j] = lo, red[warp][lane * 8 + 2 * j + 1] = hi;
      }
      __syncthreads();
      const int col_local = threadIdx.x;  // 256 threads, 256 columns of this vector slot
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][col_local];
      const int c = i * 256 + col_local;
      if (c < cols) atomicAdd(dst + c, t);
    }
  }
}

// ------------------------------------------------------------------ column sums (bias grads)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out,
                                                     long long rows, int cols, long long row_stride) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][32 * 8 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < cols) {
    // four independent 16-byte loads in flight per thread (a latency-bound stream otherwise)
    const long long step = (long long)gridDim.y * 8;
    long long r = (long long)blockIdx.y * 8 + warp;
    for (; r + 3 * step < rows; r += 4 * step) {
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] = *reinterpret_cast<const uint4*>(x + (r + i * step) * row_stride + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v[8];
        unpack_bf16x8(u[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
    for (; r < rows; r += step) {
      float v[8];
      load8(x + r * row_stride + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) atomicAdd(out + cc, t);
}

// ------------------------------------------------------------------ padded-row masking
// x[b][t][col_begin : col_begin + col_count] = 0 for t >= len[b]; one 16-byte vector per thread
__global__ void mask_rows_kernel(bf16* __restrict__ x, const int* __restrict__ len, long long batch, long long t,
                                 long long row_stride, long long batch_stride, int col_begin, int vec_per_row) {
  pdl_trigger();
  pdl_wait();
  const long long total = batch * t * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int vc = (int)(i % vec_per_row);
    const long long rt = i / vec_per_row;
    const long long tt = rt % t, b = rt / t;
    if (tt >= len[b])
      *reinterpret_cast<uint4*>(x + b * batch_stride + tt * row_stride + col_begin + vc * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------ dropout (elementwise positions of the HF modules)
// out = residual + keep ? x / (1 - p) : 0          (residual optional; backward: the same kernel on dy, no residual)
// aux_mode 1: aux_out = keep ? aux_in / (1 - p) : 0       (aux_in = act'(pre): the multiplier of the MULAUX epilogue)
// aux_mode 2: aux_out = (keep && aux_in > 0) ? 1 / (1 - p) : 0   (ReLU: aux_in = pre-activation)
// One 16-byte vector (8 elements = 4 hashed pairs) per thread; element numbering = linear index of the tensor.
__global__ void dropout_kernel(const bf16* __restrict__ x, const bf16* __restrict__ residual, bf16* __restrict__ out,
                               const bf16* __restrict__ aux_in, bf16* __restrict__ aux_out, int aux_mode, long long n_vec,
                               const unsigned long long* __restrict__ state, uint32_t call, float p) {
  pdl_trigger();
  pdl_wait();
  const DropKey key = drop_key(state, call, p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    float f[8], r[8], a[8];
    load8(x + i * 8, f);
    if (residual) load8(residual + i * 8, r);
    if (aux_mode) load8(aux_in + i * 8, a);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t bits = drop_bits(key, static_cast<uint32_t>(i) * 4u + j);
      const bool k0 = drop_keep_lo(key, bits), k1 = drop_keep_hi(key, bits);
      f[2 * j] = k0 ? f[2 * j] * key.scale : 0.f;
      f[2 * j + 1] = k1 ? f[2 * j + 1] * key.scale : 0.f;
      if (aux_mode == 1) {
        a[2 * j] = k0 ? a[2 * j] * key.scale : 0.f;
        a[2 * j + 1] = k1 ? a[2 * j + 1] * key.scale : 0.f;
      } else if (aux_mode == 2) {
        a[2 * j] = (k0 && a[2 * j] > 0.f) ? key.scale : 0.f;
        a[2 * j + 1] = (k1 && a[2 * j + 1] > 0.f) ? key.scale : 0.f;
      }
    }
    if (residual) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    store8(out + i * 8, f);
    if (aux_mode) store8(aux_out + i * 8, a);
  }
}
// keep mask as bytes (tests: the same decisions fed to the CPU oracle).  mode 0: linear numbering (elementwise
// dropout); mode 1: attention probabilities [rows][tk] -- pairs are numbered per row, row * ceil(tk / 2) + (k >> 1)
__global__ void dropout_mask_kernel(unsigned char* __restrict__ mask, long long n, long long tk, int mode,
                                    const unsigned long long* __restrict__ state, uint32_t call, float p) {
  pdl_trigger();
  pdl_wait();
  const DropKey key = drop_key(state, call, p);
  const long long hp = (tk + 1) / 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    uint32_t pair;
    bool hi;
    if (mode == 0) {
      pair = static_cast<uint32_t>(e >> 1), hi = (e & 1) != 0;
    } else {
      const long long row = e / tk, k = e - row * tk;
      pair = static_cast<uint32_t>(row * hp + (k >> 1)), hi = (k & 1) != 0;
    }
    const uint32_t bits = drop_bits(key, pair);
    mask[e] = (hi ? drop_keep_hi(key, bits) : drop_keep_lo(key, bits)) ? 1 : 0;
  }
}

// ------------------------------------------------------------------ SpecAugment (hf:...wav2vec2.py:1280-1324)
// y[b,t,:] = time_mask[b,t] ? embed : x[b,t,:];  then y[b,t,c] = 0 where feat_mask[b,c].  One 16-byte vector per thread.
__global__ void spec_augment_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                        const unsigned char* __restrict__ time_mask,
                                        const unsigned char* __restrict__ feat_mask, const float* __restrict__ embed,
                                        long long rows, long long t, int hidden) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = hidden / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vec_per_row) * 8;
    const long long r = i / vec_per_row;
    float f[8];
    if (time_mask && time_mask[r]) loadf8(embed + c, f);
    else load8(x + r * hidden + c, f);
    if (feat_mask) {
      const unsigned char* fm = feat_mask + (r / t) * hidden + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fm[j] ? 0.f : f[j];
    }
    store8(y + r * hidden + c, f);
  }
}
// dx = dy outside the masks, 0 inside
__global__ void spec_augment_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                        const unsigned char* __restrict__ time_mask,
                                        const unsigned char* __restrict__ feat_mask, long long rows, long long t,
                                        int hidden) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = hidden / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vec_per_row) * 8;
    const long long r = i / vec_per_row;
    float f[8];
    load8(dy + r * hidden + c, f);
    const bool tm = time_mask && time_mask[r];
    const unsigned char* fm = feat_mask ? feat_mask + (r / t) * hidden + c : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (tm || (fm && fm[j])) ? 0.f : f[j];
    store8(dx + r * hidden + c, f);
  }
}
// dembed[c] += sum over time-masked rows of dy[r][c] (zero where the feature mask cleared the value afterwards)
__global__ void __launch_bounds__(256) spec_augment_dembed_kernel(const bf16* __restrict__ dy, float* __restrict__ dembed,
                                                                  const unsigned char* __restrict__ time_mask,
                                                                  const unsigned char* __restrict__ feat_mask,
                                                                  long long rows, long long t, int hidden) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= hidden) return;
  float acc = 0.f;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    if (!time_mask[r]) continue;   // block-uniform
    if (feat_mask && feat_mask[(r / t) * hidden + c]) continue;
    acc += __bfloat162float(dy[r * hidden + c]);
  }
  if (acc != 0.f) atomicAdd(dembed + c, acc);
}

// ------------------------------------------------------------------ elementwise
__global__ void cast_kernel(const float* __restrict__ s, bf16* __restrict__ d, long long n) {
  pdl_trigger();
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float f[8];
    loadf8(s + i, f);
    store8(d + i, f);
  }
  if (i < n && i + 8 > n)
    for (long long j = i; j < n; ++j) d[j] = __float2bfloat16(s[j]);
}
// one block per 4096-element chunk of one table entry (binary search block -> entry)
__global__ void __launch_bounds__(256) multi_cast_kernel(const SmxCastEntry* __restrict__ table, int n_entries) {
  pdl_trigger();
  pdl_wait();
  int lo = 0, hi = n_entries - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].first_chunk <= b) lo = mid; else hi = mid - 1;
  }
  const SmxCastEntry e = table[lo];
  const long long base = (long long)(b - e.first_chunk) * SMX_CAST_CHUNK;
  const long long end = (base + SMX_CAST_CHUNK < e.n) ? base + SMX_CAST_CHUNK : e.n;
  const float* s = e.src;
  if (e.dst_f32) {
    float* d = reinterpret_cast<float*>(e.dst);
    for (long long i = base + threadIdx.x; i < end; i += 256) d[i] = s[i];
    return;
  }
  bf16* d = reinterpret_cast<bf16*>(e.dst);
  const bool vec = ((reinterpret_cast<uintptr_t>(s) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d) & 15) == 0);
  if (vec && end - base == SMX_CAST_CHUNK) {
#pragma unroll
    for (int j = 0; j < SMX_CAST_CHUNK / (256 * 8); ++j) {
      const long long i = base + (long long)(j * 256 + threadIdx.x) * 8;
      float f[8];
      loadf8(s + i, f);
      store8(d + i, f);
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += 256) d[i] = __float2bfloat16(s[i]);
  }
}
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, long long n) {
  pdl_trigger();
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float x[8], y[8];
    load8(a + i, x);
    load8(b + i, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    store8(o + i, x);
  }
  if (i < n && i + 8 > n)
    for (long long j = i; j < n; ++j) o[j] = __float2bfloat16(__bfloat162float(a[j]) + __bfloat162float(b[j]));
}
// o = a * b (bf16): the gate product of the gated feed-forward (T5 v1.1 / mT5: act(x W0) * (x W1)) and its gradients
__global__ void mul_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, long long n) {
  pdl_trigger();
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float x[8], y[8];
    load8(a + i, x);
    load8(b + i, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= y[j];
    store8(o + i, x);
  }
  if (i < n && i + 8 > n)
    for (long long j = i; j < n; ++j) o[j] = __float2bfloat16(__bfloat162float(a[j]) * __bfloat162float(b[j]));
}
__global__ void act_kernel(const bf16* __restrict__ a, bf16* __restrict__ o, long long n, int act) {
  pdl_trigger();
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float x[8];
    load8(a + i, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = act == SMX_ACT_GELU ? gelu_erf(x[j]) : fmaxf(x[j], 0.f);
    store8(o + i, x);
  }
}
// out = dy * act'(pre)
__global__ void dact_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pre, bf16* __restrict__ o, long long n,
                            int act) {
  pdl_trigger();
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float d[8], x[8];
    load8(dy + i, d);
    load8(pre + i, x);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      d[j] = act == SMX_ACT_MULAUX ? d[j] * x[j] : (act == SMX_ACT_GELU ? d[j] * gelu_erf_grad(x[j]) : (x[j] > 0.f ? d[j] : 0.f));
    store8(o + i, d);
  }
}
__global__ void pack_conv_w_kernel(const float* __restrict__ s, bf16* __restrict__ d, long long cout, long long cin,
                                   long long k) {
  pdl_trigger();
  pdl_wait();
  const long long n = cout * cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / (cin * k), r = i % (cin * k), t = r / cin, c = r % cin;
    d[i] = __float2bfloat16(s[(o * cin + c) * k + t]);
  }
}
__global__ void unpack_conv_g_kernel(const float* __restrict__ s, float* __restrict__ d, long long cout, long long cin,
                                     long long k) {
  pdl_trigger();
  pdl_wait();
  const long long n = cout * cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / (cin * k), r = i % (cin * k), c = r / k, t = r % k;
    d[i] = s[o * cin * k + t * cin + c];
  }
}

// ------------------------------------------------------------------ embeddings
// out[b,t,:] = (ids ? tok[ids[b,t]]*scale : 0) + (x_in ? x_in[b,t,:] : 0) + (pos ? pos[t+t_start+off] : 0)
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                                 const float* __restrict__ pos, const bf16* __restrict__ x_in, bf16* __restrict__ out,
                                 long long rows, long long t_len, int dim, float scale, long long pos_off) {
  pdl_trigger();
  pdl_wait();
  const int vecs = dim / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vecs;
    const int c = (int)(i % vecs) * 8;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ids) {
      float e[8];
      loadf8(tok + ids[row] * dim + c, e);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = e[j] * scale;
    }
    if (x_in) {
      float e[8];
      load8(x_in + row * dim + c, e);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += e[j];
    }
    if (pos) {
      float e[8];
      loadf8(pos + ((row % t_len) + pos_off) * dim + c, e);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += e[j];
    }
    store8(out + row * dim + c, f);
  }
}
__global__ void embed_bwd_kernel(const long long* __restrict__ ids, const bf16* __restrict__ dout,
                                 float* __restrict__ dtok, float* __restrict__ dpos, long long rows, long long t_len,
                                 int dim, float scale, long long pos_off) {
  pdl_trigger();
  pdl_wait();
  const int vecs = dim / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vecs;
    const int c = (int)(i % vecs) * 8;
    float f[8];
    load8(dout + row * dim + c, f);
    if (dtok && ids) {
      float* p = dtok + ids[row] * dim + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(p + j, f[j] * scale);
    }
    if (dpos) {
      float* p = dpos + ((row % t_len) + pos_off) * dim + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(p + j, f[j]);
    }
  }
}

// ------------------------------------------------------------------ weighted layer sum
constexpr int kMaxLayers = 32;
struct PtrPack {
  const bf16* p[kMaxLayers];
};
__global__ void wsum_fwd_kernel(PtrPack xs, const float* __restrict__ w, bf16* __restrict__ out, int nl, long long n) {
  pdl_trigger();
  pdl_wait();
  float wl[kMaxLayers];
  for (int l = 0; l < nl; ++l) wl[l] = w[l];
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = 0; l < nl; ++l) {
      float v[8];
      load8(xs.p[l] + i, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += wl[l] * v[j];
    }
    store8(out + i, acc);
  }
}
__global__ void __launch_bounds__(256) wsum_bwd_w_kernel(PtrPack xs, const bf16* __restrict__ dout,
                                                         float* __restrict__ dw, int nl, long long n) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (int l = 0; l < nl; ++l) {
    float acc = 0.f;
    for (long long i = i0; i + 8 <= n; i += stride) {
      float v[8], d[8];
      load8(xs.p[l] + i, v);
      load8(dout + i, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j] * d[j];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += red[k];
      atomicAdd(dw + l, t);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ weight norm over the last dim's complement
// w[r][k] = g[k] * v[r][k] / ||v[:, k]||   (torch weight_norm(dim=2) of the positional conv: one norm per tap,
// hf:models/wav2vec2/modeling_wav2vec2.py:341-355).  v viewed as [rows = out*in/groups][k].
__global__ void __launch_bounds__(256) wn_colstat_kernel(const float* __restrict__ a, const float* __restrict__ b2,
                                                         float* __restrict__ out, long long rows, int k) {
  pdl_trigger();
  pdl_wait();
  // out[j] += sum_r a[r][j] * (b2 ? b2[r][j] : a[r][j]);  blockDim.x = 256 = 8 row lanes x 32 columns
  __shared__ float red[8][33];
  const int cj = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cj;
  float acc = 0.f;
  if (j < k)
    for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += (long long)gridDim.y * 8) {
      const float x = a[r * k + j];
      acc = fmaf(x, b2 ? b2[r * k + j] : x, acc);
    }
  red[rl][cj] = acc;
  __syncthreads();
  if (rl == 0 && j < k) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][cj];
    atomicAdd(out + j, t);
  }
}
__global__ void wn_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ sq,
                              float* __restrict__ w, long long n, int k) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % k);
    w[i] = v[i] * g[j] * rsqrtf(sq[j]);
  }
}
// dv = g/||v|| * (dw - v * dot/||v||^2),  dg = dot / ||v||   with dot[j] = sum_r dw[r][j] v[r][j]
__global__ void wn_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ sq,
                              const float* __restrict__ dot, const float* __restrict__ dw, float* __restrict__ dv,
                              float* __restrict__ dg, long long n, int k) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % k);
    const float inv = rsqrtf(sq[j]);
    dv[i] = g[j] * inv * (dw[i] - v[i] * dot[j] * inv * inv);
    if (i < k) dg[i] = dot[i] * inv;
  }
}

static int grid_for(long long work_items, int block, int max_waves = 8) {
  long long g = (work_items + block - 1) / block;
  const long long cap = (long long)num_sms() * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// Shared-memory ring of the LayerNorm kernels: as many row slots per warp as fit ~96 KB per block (2..8), two blocks
// per SM -> up to ~190 KB of rows in flight per SM.  max_blocks_per_sm == 0: free grid (fwd), otherwise a persistent
// wave of that many resident blocks (bwd keeps column accumulators in registers).
struct RingPlan {
  int grid, stages;
  size_t smem;
};
static RingPlan ring_plan(long long rows, long long cols, int n_inputs, int max_blocks_per_sm) {
  const size_t slot = (size_t)cols * 2 * n_inputs;       // one row of every input
  // forward (58 registers): four blocks of ~48 KB per SM; backward (128 registers): two blocks of ~96 KB
  const size_t budget = (max_blocks_per_sm == 0 ? 48 : 96) * 1024;
  int stages = (int)(budget / (8 * slot));
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  RingPlan rp;
  rp.stages = stages;
  rp.smem = 8 * stages * slot + 8 * stages * sizeof(uint64_t);
  int per_sm = (int)((227 * 1024) / (rp.smem + 10 * 1024));   // + the static reduction buffer and the 1 KB reserve
  if (per_sm < 1) per_sm = 1;
  const int cap = max_blocks_per_sm > 0 ? max_blocks_per_sm : 4;
  if (per_sm > cap) per_sm = cap;
  rp.grid = grid_for(rows, 8, per_sm);
  return rp;
}
// opt in to > 48 KB of dynamic shared memory once per kernel instantiation (one caller thread per process)
template <typename K>
static cudaError_t ring_attr(K kernel) {
  static const void* done[64];
  static int n_done = 0;
  for (int i = 0; i < n_done; ++i)
    if (done[i] == reinterpret_cast<const void*>(kernel)) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess && n_done < 64) done[n_done++] = reinterpret_cast<const void*>(kernel);
  return e;
}

}  // namespace rw
}  // namespace smx

using namespace smx;
using namespace smx::rw;

#define LN_DISPATCH(VPLV, KERNEL, ...)           \
  switch (VPLV) {                                \
    case 1: KERNEL<1> __VA_ARGS__; break;        \
    case 2: KERNEL<2> __VA_ARGS__; break;        \
    case 3: KERNEL<3> __VA_ARGS__; break;        \
    case 4: KERNEL<4> __VA_ARGS__; break;        \
    case 5: case 6: KERNEL<6> __VA_ARGS__; break;\
    default: KERNEL<8> __VA_ARGS__; break;       \
  }

extern "C" {

int smx_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, void* sum_out,
                      float* mean, float* rstd, int64_t rows, int64_t cols, float eps, int rms_only, int act,
                      void* stream) {
  SMX_REQUIRE(act == SMX_ACT_NONE || act == SMX_ACT_GELU, "layernorm: unsupported fused activation %d", act);
  SMX_REQUIRE(cols % 8 == 0 && cols <= 2048 && cols > 0, "layernorm: cols %lld must be a multiple of 8 and <= 2048",
              (long long)cols);
  if (rows == 0) return 0;
  SMX_REQUIRE(aligned16(x) && (res == nullptr || aligned16(res)), "layernorm: inputs must be 16-byte aligned");
  const int vpl = (int)ceil_div(cols, 256);
  RingPlan rp = ring_plan(rows, cols, res ? 2 : 1, 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int reg_grid = grid_for(rows, 8);
#define LN_FWD_LAUNCH(V)                                                                                         \
  do {                                                                                                           \
    if (res == nullptr) {                                                                                        \
      launch_pdl(ln_fwd_reg_kernel<V>, dim3(reg_grid), dim3(256), 0, st, (const bf16*)x, (const bf16*)res, gamma, beta, (bf16*)y,    \
                                                     (bf16*)sum_out, mean, rstd, rows, (int)cols, eps, rms_only, \
                                                     act);                                                       \
    } else {                                                                                                     \
      SMX_CHECK_CUDA(ring_attr(ln_fwd_kernel<V>));                                                               \
      launch_pdl(ln_fwd_kernel<V>, dim3(rp.grid), dim3(256), rp.smem, st, (const bf16*)x, (const bf16*)res, gamma, beta, (bf16*)y,   \
                                                      (bf16*)sum_out, mean, rstd, rows, (int)cols, eps, rms_only, \
                                                      act, rp.stages);                                           \
    }                                                                                                            \
  } while (0)
  switch (vpl) {
    case 1: LN_FWD_LAUNCH(1); break;
    case 2: LN_FWD_LAUNCH(2); break;
    case 3: LN_FWD_LAUNCH(3); break;
    case 4: LN_FWD_LAUNCH(4); break;
    case 5: case 6: LN_FWD_LAUNCH(6); break;
    default: LN_FWD_LAUNCH(8); break;
  }
#undef LN_FWD_LAUNCH
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, const void* dres_in, void* dx, float* dgamma, float* dbeta, float* dx_colsum,
                      int64_t rows, int64_t cols, int rms_only, int act, void* stream) {
  SMX_REQUIRE(act == SMX_ACT_NONE || (act == SMX_ACT_GELU && beta != nullptr), "layernorm_bwd: bad fused activation");
  SMX_REQUIRE(cols % 8 == 0 && cols <= 2048 && cols > 0, "layernorm_bwd: cols %lld unsupported", (long long)cols);
  SMX_REQUIRE(dgamma != nullptr, "layernorm_bwd: dgamma required");
  if (rows == 0) return 0;
  SMX_REQUIRE(aligned16(dy) && aligned16(x) && (dres_in == nullptr || aligned16(dres_in)),
              "layernorm_bwd: inputs must be 16-byte aligned");
  const int vpl = (int)ceil_div(cols, 256);
  const bool cs = dx_colsum != nullptr;
  // the column accumulators live in registers for the whole kernel: one persistent wave of resident blocks
  RingPlan rp = ring_plan(rows, cols, dres_in ? 3 : 2, vpl <= 3 ? 2 : 1);
  cudaStream_t st = (cudaStream_t)stream;
  const int reg_grid = grid_for(rows, 8, vpl <= 3 ? 2 : 1);
#define LN_BWD_LAUNCH(V, C)                                                                                        \
  do {                                                                                                             \
    if (dres_in == nullptr) {                                                                                      \
      launch_pdl(ln_bwd_reg_kernel<V, C>, dim3(reg_grid), dim3(256), 0, st, (const bf16*)dy, (const bf16*)x, gamma, beta, mean, rstd,  \
                                                        (const bf16*)dres_in, (bf16*)dx, dgamma, dbeta, dx_colsum, \
                                                        rows, (int)cols, rms_only, act);                           \
    } else {                                                                                                       \
      SMX_CHECK_CUDA(ring_attr(ln_bwd_kernel<V, C>));                                                              \
      launch_pdl(ln_bwd_kernel<V, C>, dim3(rp.grid), dim3(256), rp.smem, st, (const bf16*)dy, (const bf16*)x, gamma, beta, mean, rstd, \
                                                         (const bf16*)dres_in, (bf16*)dx, dgamma, dbeta, dx_colsum, \
                                                         rows, (int)cols, rms_only, act, rp.stages);               \
    }                                                                                                              \
  } while (0)
#define LN_BWD_CASE(V) if (cs) LN_BWD_LAUNCH(V, true); else LN_BWD_LAUNCH(V, false); break
  switch (vpl) {
    case 1: LN_BWD_CASE(1);
    case 2: LN_BWD_CASE(2);
    case 3: LN_BWD_CASE(3);
    case 4: LN_BWD_CASE(4);
    case 5: case 6: LN_BWD_CASE(6);
    default: LN_BWD_CASE(8);
  }
#undef LN_BWD_CASE
#undef LN_BWD_LAUNCH
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_mask_rows(void* x, const int32_t* len, int64_t batch, int64_t t, int64_t row_stride, int64_t batch_stride,
                  int64_t col_begin, int64_t col_count, void* stream) {
  SMX_REQUIRE(x && len, "mask_rows: null pointer");
  SMX_REQUIRE(col_begin % 8 == 0 && col_count % 8 == 0 && row_stride % 8 == 0 && batch_stride % 8 == 0 && aligned16(x),
              "mask_rows: columns / strides must be multiples of 8 elements and x 16-byte aligned");
  if (batch == 0 || t == 0 || col_count == 0) return 0;
  const int vec = (int)(col_count / 8);
  launch_pdl(mask_rows_kernel, dim3(grid_for(batch * t * vec, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (bf16*)x, len, batch, t, row_stride, batch_stride, (int)col_begin, vec);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_dropout(const void* x, const void* residual, void* out, const void* aux_in, void* aux_out, int aux_mode, int64_t n,
                const uint64_t* state, uint32_t call, float p, void* stream) {
  SMX_REQUIRE(x && out && state, "dropout: null pointer");
  SMX_REQUIRE(p >= 0.0f && p < 1.0f, "dropout: p = %g outside [0, 1)", (double)p);
  SMX_REQUIRE(n % 8 == 0 && n / 2 < (1ll << 32), "dropout: n = %lld must be a multiple of 8 below 2^33", (long long)n);
  SMX_REQUIRE(aux_mode == 0 || (aux_mode >= 1 && aux_mode <= 2 && aux_in && aux_out), "dropout: bad auxiliary operand");
  SMX_REQUIRE(aligned16(x) && aligned16(out) && aligned16(residual) && aligned16(aux_in) && aligned16(aux_out),
              "dropout: operands must be 16-byte aligned");
  if (n == 0) return 0;
  launch_pdl(dropout_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)residual, (bf16*)out,
                                                                       (const bf16*)aux_in, (bf16*)aux_out, aux_mode, n / 8,
                                                                       (const unsigned long long*)state, call, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_layernorm_dropout_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, void* sum_out,
                              float* mean, float* rstd, int64_t rows, int64_t cols, float eps, const uint64_t* state,
                              uint32_t call, float p, void* stream) {
  SMX_REQUIRE(x && res && gamma && y && sum_out && mean && rstd && state, "layernorm_dropout_fwd: null pointer");
  SMX_REQUIRE(p > 0.0f && p < 1.0f, "layernorm_dropout_fwd: p = %g outside (0, 1)", (double)p);
  SMX_REQUIRE(cols % 8 == 0 && cols <= 2048 && cols > 0, "layernorm_dropout_fwd: cols %lld unsupported", (long long)cols);
  SMX_REQUIRE(rows * cols / 2 < (1ll << 32), "layernorm_dropout_fwd: more than 2^33 elements");
  if (rows == 0) return 0;
  SMX_REQUIRE(aligned16(x) && aligned16(res) && aligned16(y) && aligned16(sum_out), "layernorm_dropout_fwd: 16-byte alignment");
  const int vpl = (int)ceil_div(cols, 256);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(rows, 8);
#define LN_FWD_DROP_LAUNCH(V)                                                                                              \
  launch_pdl(ln_fwd_drop_kernel<V>, dim3(grid), dim3(256), 0, st, (const bf16*)x, (const bf16*)res, gamma, beta, (bf16*)y, \
             (bf16*)sum_out, mean, rstd, (long long)rows, (int)cols, eps, (const unsigned long long*)state, call, p)
  switch (vpl) {
    case 1: LN_FWD_DROP_LAUNCH(1); break;
    case 2: LN_FWD_DROP_LAUNCH(2); break;
    case 3: LN_FWD_DROP_LAUNCH(3); break;
    case 4: LN_FWD_DROP_LAUNCH(4); break;
    case 5: case 6: LN_FWD_DROP_LAUNCH(6); break;
    default: LN_FWD_DROP_LAUNCH(8); break;
  }
#undef LN_FWD_DROP_LAUNCH
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_layernorm_dropout_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                              void* dx_drop, float* dgamma, float* dbeta, float* dxd_colsum, int64_t rows, int64_t cols,
                              const uint64_t* state, uint32_t call, float p, void* stream) {
  SMX_REQUIRE(dy && x && gamma && mean && rstd && dx && dx_drop && dgamma && state, "layernorm_dropout_bwd: null pointer");
  SMX_REQUIRE(p > 0.0f && p < 1.0f, "layernorm_dropout_bwd: p = %g outside (0, 1)", (double)p);
  SMX_REQUIRE(cols % 8 == 0 && cols <= 2048 && cols > 0, "layernorm_dropout_bwd: cols %lld unsupported", (long long)cols);
  SMX_REQUIRE(rows * cols / 2 < (1ll << 32), "layernorm_dropout_bwd: more than 2^33 elements");
  if (rows == 0) return 0;
  SMX_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(dx_drop), "layernorm_dropout_bwd: 16-byte alignment");
  const int vpl = (int)ceil_div(cols, 256);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(rows, 8, vpl <= 3 ? 2 : 1);   // one persistent wave: the column accumulators live in registers
#define LN_BWD_DROP_LAUNCH(V)                                                                                             \
  launch_pdl(ln_bwd_drop_kernel<V>, dim3(grid), dim3(256), 0, st, (const bf16*)dy, (const bf16*)x, gamma, mean, rstd,     \
             (bf16*)dx, (bf16*)dx_drop, dgamma, dbeta, dxd_colsum, (long long)rows, (int)cols,                            \
             (const unsigned long long*)state, call, p)
  switch (vpl) {
    case 1: LN_BWD_DROP_LAUNCH(1); break;
    case 2: LN_BWD_DROP_LAUNCH(2); break;
    case 3: LN_BWD_DROP_LAUNCH(3); break;
    case 4: LN_BWD_DROP_LAUNCH(4); break;
    case 5: case 6: LN_BWD_DROP_LAUNCH(6); break;
    default: LN_BWD_DROP_LAUNCH(8); break;
  }
#undef LN_BWD_DROP_LAUNCH
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_dropout_mask(uint8_t* mask, int64_t n, int64_t tk, int mode, const uint64_t* state, uint32_t call, float p,
                     void* stream) {
  SMX_REQUIRE(mask && state && (mode == 0 || (mode == 1 && tk > 0)), "dropout_mask: bad arguments");
  if (n == 0) return 0;
  launch_pdl(dropout_mask_kernel, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, mask, n, tk, mode, (const unsigned long long*)state,
                                                                         call, p);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_spec_augment_fwd(const void* x, void* y, const uint8_t* time_mask, const uint8_t* feat_mask, const float* embed,
                         int64_t batch, int64_t t, int64_t hidden, void* stream) {
  SMX_REQUIRE(x && y && hidden % 8 == 0 && aligned16(x) && aligned16(y), "spec_augment: bad arguments");
  SMX_REQUIRE(time_mask == nullptr || (embed != nullptr && aligned16(embed)), "spec_augment: time mask needs the embedding");
  if (batch * t == 0) return 0;
  launch_pdl(spec_augment_fwd_kernel, dim3(grid_for(batch * t * (hidden / 8), 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const bf16*)x, (bf16*)y, time_mask, feat_mask, embed, batch * t, t, (int)hidden);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_spec_augment_bwd(const void* dy, void* dx, float* dembed, const uint8_t* time_mask, const uint8_t* feat_mask,
                         int64_t batch, int64_t t, int64_t hidden, void* stream) {
  SMX_REQUIRE(dy && dx && hidden % 8 == 0 && aligned16(dy) && aligned16(dx), "spec_augment_bwd: bad arguments");
  if (batch * t == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl(spec_augment_bwd_kernel, dim3(grid_for(batch * t * (hidden / 8), 256)), dim3(256), 0, st, 
      (const bf16*)dy, (bf16*)dx, time_mask, feat_mask, batch * t, t, (int)hidden);
  SMX_CHECK_CUDA(cudaGetLastError());
  if (dembed && time_mask) {   // dembed is zero-initialised by the caller
    long long gy = batch * t < 592 ? batch * t : 592;
    launch_pdl(spec_augment_dembed_kernel, dim3(dim3((unsigned)ceil_div(hidden, 256), (unsigned)gy)), dim3(256), 0, st, 
        (const bf16*)dy, dembed, time_mask, feat_mask, batch * t, t, (int)hidden);
    SMX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int smx_colsum(const void* x, float* out, int64_t rows, int64_t cols, int64_t row_stride, void* stream) {
  SMX_REQUIRE(cols % 8 == 0 && row_stride % 8 == 0, "colsum: cols/stride must be multiples of 8");
  if (rows == 0) return 0;
  const int gx = (int)ceil_div(cols, 256);
  long long gy = ceil_div(rows, 8 * 16);
  const long long cap = (long long)num_sms() * 4 / gx + 1;
  if (gy > cap) gy = cap;
  launch_pdl(colsum_kernel, dim3(dim3(gx, (unsigned)gy)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, out, rows, (int)cols,
                                                                          row_stride);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n == 0) return 0;
  SMX_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "cast: pointers must be 16-byte aligned");
  launch_pdl(cast_kernel, dim3(grid_for(ceil_div(n, 8), 256)), dim3(256), 0, (cudaStream_t)stream, src, (bf16*)dst, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_multi_cast(const SmxCastEntry* table, int32_t n_entries, int32_t total_chunks, void* stream) {
  if (n_entries <= 0 || total_chunks <= 0) return 0;
  launch_pdl(multi_cast_kernel, dim3(total_chunks), dim3(256), 0, (cudaStream_t)stream, table, n_entries);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_weightnorm_fwd(const float* v, const float* g, float* sq_ws, float* w, int64_t rows, int64_t k, void* stream) {
  SMX_REQUIRE(v && g && sq_ws && w, "weightnorm_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SMX_CHECK_CUDA(cudaMemsetAsync(sq_ws, 0, sizeof(float) * k, st));
  long long gy = ceil_div(rows, 8 * 32);
  if (gy > 512) gy = 512;
  launch_pdl(wn_colstat_kernel, dim3(dim3((unsigned)ceil_div(k, 32), (unsigned)gy)), dim3(256), 0, st, v, nullptr, sq_ws, rows, (int)k);
  SMX_CHECK_CUDA(cudaGetLastError());
  launch_pdl(wn_fwd_kernel, dim3(grid_for(rows * k, 256)), dim3(256), 0, st, v, g, sq_ws, w, rows * k, (int)k);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_weightnorm_bwd(const float* v, const float* g, const float* sq, const float* dw, float* dot_ws, float* dv,
                       float* dg, int64_t rows, int64_t k, void* stream) {
  SMX_REQUIRE(v && g && sq && dw && dot_ws && dv && dg, "weightnorm_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SMX_CHECK_CUDA(cudaMemsetAsync(dot_ws, 0, sizeof(float) * k, st));
  long long gy = ceil_div(rows, 8 * 32);
  if (gy > 512) gy = 512;
  launch_pdl(wn_colstat_kernel, dim3(dim3((unsigned)ceil_div(k, 32), (unsigned)gy)), dim3(256), 0, st, dw, v, dot_ws, rows, (int)k);
  SMX_CHECK_CUDA(cudaGetLastError());
  launch_pdl(wn_bwd_kernel, dim3(grid_for(rows * k, 256)), dim3(256), 0, st, v, g, sq, dot_ws, dw, dv, dg, rows * k, (int)k);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  launch_pdl(add_kernel, dim3(grid_for(ceil_div(n, 8), 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)a, (const bf16*)b,
                                                                              (bf16*)out, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_mul_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  launch_pdl(mul_kernel, dim3(grid_for(ceil_div(n, 8), 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)a, (const bf16*)b,
             (bf16*)out, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_act_bf16(const void* x, void* y, int64_t n, int act, void* stream) {
  if (n == 0) return 0;
  SMX_REQUIRE(n % 8 == 0, "act: n must be a multiple of 8");
  launch_pdl(act_kernel, dim3(grid_for(ceil_div(n, 8), 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, n, act);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_dact_bf16(const void* dy, const void* pre, void* out, int64_t n, int act, void* stream) {
  if (n == 0) return 0;
  SMX_REQUIRE(n % 8 == 0, "dact: n must be a multiple of 8");
  launch_pdl(dact_kernel, dim3(grid_for(ceil_div(n, 8), 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dy, (const bf16*)pre,
                                                                               (bf16*)out, n, act);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_pack_conv_weight(const float* src, void* dst, int64_t cout, int64_t cin, int64_t k, void* stream) {
  launch_pdl(pack_conv_w_kernel, dim3(grid_for(cout * cin * k, 256)), dim3(256), 0, (cudaStream_t)stream, src, (bf16*)dst, cout, cin, k);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_unpack_conv_wgrad(const float* src, float* dst, int64_t cout, int64_t cin, int64_t k, void* stream) {
  launch_pdl(unpack_conv_g_kernel, dim3(grid_for(cout * cin * k, 256)), dim3(256), 0, (cudaStream_t)stream, src, dst, cout, cin, k);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_embed_fwd(const int64_t* ids, const float* tok_emb, const float* pos_emb, const void* x_in, void* out,
                  int64_t batch, int64_t t, int64_t dim, float scale, int64_t pos_offset, int64_t t_start,
                  void* stream) {
  SMX_REQUIRE(dim % 8 == 0, "embed: dim must be a multiple of 8");
  const long long rows = batch * t;
  if (rows == 0) return 0;
  launch_pdl(embed_fwd_kernel, dim3(grid_for(rows * (dim / 8), 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const long long*)ids, tok_emb, pos_emb, (const bf16*)x_in, (bf16*)out, rows, t, (int)dim, scale,
      pos_offset + t_start);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_embed_bwd(const int64_t* ids, const void* dout, float* d_tok_emb, float* d_pos_emb, int64_t batch, int64_t t,
                  int64_t dim, float scale, int64_t pos_offset, void* stream) {
  SMX_REQUIRE(dim % 8 == 0, "embed: dim must be a multiple of 8");
  const long long rows = batch * t;
  if (rows == 0) return 0;
  launch_pdl(embed_bwd_kernel, dim3(grid_for(rows * (dim / 8), 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const long long*)ids, (const bf16*)dout, d_tok_emb, d_pos_emb, rows, t, (int)dim, scale, pos_offset);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_weighted_sum_fwd(const void* const* xs, const float* w, void* out, int n_layers, int64_t n, void* stream) {
  SMX_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, "weighted_sum: %d layers unsupported", n_layers);
  SMX_REQUIRE(n % 8 == 0, "weighted_sum: n must be a multiple of 8");
  PtrPack pk;
  for (int l = 0; l < n_layers; ++l) pk.p[l] = (const bf16*)xs[l];
  launch_pdl(wsum_fwd_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, (cudaStream_t)stream, pk, w, (bf16*)out, n_layers, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_weighted_sum_bwd_w(const void* const* xs, const void* dout, float* dw, int n_layers, int64_t n,
                           void* stream) {
  SMX_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, "weighted_sum: %d layers unsupported", n_layers);
  SMX_REQUIRE(n % 8 == 0, "weighted_sum: n must be a multiple of 8");
  PtrPack pk;
  for (int l = 0; l < n_layers; ++l) pk.p[l] = (const bf16*)xs[l];
  launch_pdl(wsum_bwd_w_kernel, dim3(grid_for(n / 8, 256, 2)), dim3(256), 0, (cudaStream_t)stream, pk, (const bf16*)dout, dw, n_layers, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
