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

// ------------------------------------------------------------------ LayerNorm forward
// one warp per row; VPL = 16-byte vectors per lane (cols <= VPL*256)
template <int VPL>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     bf16* __restrict__ y, bf16* __restrict__ sum_out,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     long long rows, int cols, float eps, int rms_only, int act) {
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

// ------------------------------------------------------------------ LayerNorm backward
// One warp per row, rows strided over a persistent grid; parameter gradients (and, optionally, the column sums
// of dx = the bias gradient of the linear layer in front of a post-LN block) accumulate in registers and are
// reduced once per block.  Between the statistics pass and the dx pass a row is held as the PACKED bf16 it was
// loaded as (2 x 4 registers per 8 elements) instead of two fp32 copies, which is what keeps the kernel at
// two resident 256-thread blocks per SM -- an HBM-bound row kernel needs the
// warps to cover the memory latency (the first version: 159 registers, 8 warps per SM, 29 % of HBM peak).
template <int VPL, bool WITH_CS>
__global__ void __launch_bounds__(256, (VPL <= 3 ? 2 : 1))
ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dx_colsum, long long rows, int cols, int rms_only, int act) {
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

// ------------------------------------------------------------------ column sums (bias grads)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out,
                                                     long long rows, int cols, long long row_stride) {
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

// ------------------------------------------------------------------ elementwise
__global__ void cast_kernel(const float* __restrict__ s, bf16* __restrict__ d, long long n) {
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
__global__ void act_kernel(const bf16* __restrict__ a, bf16* __restrict__ o, long long n, int act) {
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
  const long long n = cout * cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / (cin * k), r = i % (cin * k), t = r / cin, c = r % cin;
    d[i] = __float2bfloat16(s[(o * cin + c) * k + t]);
  }
}
__global__ void unpack_conv_g_kernel(const float* __restrict__ s, float* __restrict__ d, long long cout, long long cin,
                                     long long k) {
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
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % k);
    w[i] = v[i] * g[j] * rsqrtf(sq[j]);
  }
}
// dv = g/||v|| * (dw - v * dot/||v||^2),  dg = dot / ||v||   with dot[j] = sum_r dw[r][j] v[r][j]
__global__ void wn_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ sq,
                              const float* __restrict__ dot, const float* __restrict__ dw, float* __restrict__ dv,
                              float* __restrict__ dg, long long n, int k) {
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
  const int vpl = (int)ceil_div(cols, 256);
  const int grid = grid_for(rows, 8);
  cudaStream_t st = (cudaStream_t)stream;
  LN_DISPATCH(vpl, ln_fwd_kernel,
              <<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)res, gamma, beta, (bf16*)y, (bf16*)sum_out, mean,
                                     rstd, rows, (int)cols, eps, rms_only, act));
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
  const int vpl = (int)ceil_div(cols, 256);
  const bool cs = dx_colsum != nullptr;
  int grid = grid_for(rows, 8, vpl <= 3 ? 2 : 1);   // resident blocks per SM: one persistent wave
  cudaStream_t st = (cudaStream_t)stream;
#define LN_BWD_ARGS                                                                                              \
  <<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, gamma, beta, mean, rstd, (const bf16*)dres_in, (bf16*)dx, \
                         dgamma, dbeta, dx_colsum, rows, (int)cols, rms_only, act)
  switch (vpl) {
    case 1: if (cs) ln_bwd_kernel<1, true> LN_BWD_ARGS; else ln_bwd_kernel<1, false> LN_BWD_ARGS; break;
    case 2: if (cs) ln_bwd_kernel<2, true> LN_BWD_ARGS; else ln_bwd_kernel<2, false> LN_BWD_ARGS; break;
    case 3: if (cs) ln_bwd_kernel<3, true> LN_BWD_ARGS; else ln_bwd_kernel<3, false> LN_BWD_ARGS; break;
    case 4: if (cs) ln_bwd_kernel<4, true> LN_BWD_ARGS; else ln_bwd_kernel<4, false> LN_BWD_ARGS; break;
    case 5: case 6: if (cs) ln_bwd_kernel<6, true> LN_BWD_ARGS; else ln_bwd_kernel<6, false> LN_BWD_ARGS; break;
    default: if (cs) ln_bwd_kernel<8, true> LN_BWD_ARGS; else ln_bwd_kernel<8, false> LN_BWD_ARGS; break;
  }
#undef LN_BWD_ARGS
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_colsum(const void* x, float* out, int64_t rows, int64_t cols, int64_t row_stride, void* stream) {
  SMX_REQUIRE(cols % 8 == 0 && row_stride % 8 == 0, "colsum: cols/stride must be multiples of 8");
  if (rows == 0) return 0;
  const int gx = (int)ceil_div(cols, 256);
  long long gy = ceil_div(rows, 8 * 16);
  const long long cap = (long long)num_sms() * 4 / gx + 1;
  if (gy > cap) gy = cap;
  colsum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, out, rows, (int)cols,
                                                                          row_stride);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n == 0) return 0;
  SMX_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "cast: pointers must be 16-byte aligned");
  cast_kernel<<<grid_for(ceil_div(n, 8), 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_multi_cast(const SmxCastEntry* table, int32_t n_entries, int32_t total_chunks, void* stream) {
  if (n_entries <= 0 || total_chunks <= 0) return 0;
  multi_cast_kernel<<<total_chunks, 256, 0, (cudaStream_t)stream>>>(table, n_entries);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_weightnorm_fwd(const float* v, const float* g, float* sq_ws, float* w, int64_t rows, int64_t k, void* stream) {
  SMX_REQUIRE(v && g && sq_ws && w, "weightnorm_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SMX_CHECK_CUDA(cudaMemsetAsync(sq_ws, 0, sizeof(float) * k, st));
  long long gy = ceil_div(rows, 8 * 32);
  if (gy > 512) gy = 512;
  wn_colstat_kernel<<<dim3((unsigned)ceil_div(k, 32), (unsigned)gy), 256, 0, st>>>(v, nullptr, sq_ws, rows, (int)k);
  SMX_CHECK_CUDA(cudaGetLastError());
  wn_fwd_kernel<<<grid_for(rows * k, 256), 256, 0, st>>>(v, g, sq_ws, w, rows * k, (int)k);
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
  wn_colstat_kernel<<<dim3((unsigned)ceil_div(k, 32), (unsigned)gy), 256, 0, st>>>(dw, v, dot_ws, rows, (int)k);
  SMX_CHECK_CUDA(cudaGetLastError());
  wn_bwd_kernel<<<grid_for(rows * k, 256), 256, 0, st>>>(v, g, sq, dot_ws, dw, dv, dg, rows * k, (int)k);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  add_kernel<<<grid_for(ceil_div(n, 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b,
                                                                              (bf16*)out, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_act_bf16(const void* x, void* y, int64_t n, int act, void* stream) {
  if (n == 0) return 0;
  SMX_REQUIRE(n % 8 == 0, "act: n must be a multiple of 8");
  act_kernel<<<grid_for(ceil_div(n, 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n, act);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_dact_bf16(const void* dy, const void* pre, void* out, int64_t n, int act, void* stream) {
  if (n == 0) return 0;
  SMX_REQUIRE(n % 8 == 0, "dact: n must be a multiple of 8");
  dact_kernel<<<grid_for(ceil_div(n, 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)pre,
                                                                               (bf16*)out, n, act);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_pack_conv_weight(const float* src, void* dst, int64_t cout, int64_t cin, int64_t k, void* stream) {
  pack_conv_w_kernel<<<grid_for(cout * cin * k, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, cout, cin, k);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_unpack_conv_wgrad(const float* src, float* dst, int64_t cout, int64_t cin, int64_t k, void* stream) {
  unpack_conv_g_kernel<<<grid_for(cout * cin * k, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, cout, cin, k);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_embed_fwd(const int64_t* ids, const float* tok_emb, const float* pos_emb, const void* x_in, void* out,
                  int64_t batch, int64_t t, int64_t dim, float scale, int64_t pos_offset, int64_t t_start,
                  void* stream) {
  SMX_REQUIRE(dim % 8 == 0, "embed: dim must be a multiple of 8");
  const long long rows = batch * t;
  if (rows == 0) return 0;
  embed_fwd_kernel<<<grid_for(rows * (dim / 8), 256), 256, 0, (cudaStream_t)stream>>>(
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
  embed_bwd_kernel<<<grid_for(rows * (dim / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)ids, (const bf16*)dout, d_tok_emb, d_pos_emb, rows, t, (int)dim, scale, pos_offset);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_weighted_sum_fwd(const void* const* xs, const float* w, void* out, int n_layers, int64_t n, void* stream) {
  SMX_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, "weighted_sum: %d layers unsupported", n_layers);
  SMX_REQUIRE(n % 8 == 0, "weighted_sum: n must be a multiple of 8");
  PtrPack pk;
  for (int l = 0; l < n_layers; ++l) pk.p[l] = (const bf16*)xs[l];
  wsum_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(pk, w, (bf16*)out, n_layers, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_weighted_sum_bwd_w(const void* const* xs, const void* dout, float* dw, int n_layers, int64_t n,
                           void* stream) {
  SMX_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, "weighted_sum: %d layers unsupported", n_layers);
  SMX_REQUIRE(n % 8 == 0, "weighted_sum: n must be a multiple of 8");
  PtrPack pk;
  for (int l = 0; l < n_layers; ++l) pk.p[l] = (const bf16*)xs[l];
  wsum_bwd_w_kernel<<<grid_for(n / 8, 256, 2), 256, 0, (cudaStream_t)stream>>>(pk, (const bf16*)dout, dw, n_layers, n);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
