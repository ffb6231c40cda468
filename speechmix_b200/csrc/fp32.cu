// fp32 verification path (inference only): the same forward graph evaluated with fp32 activations and
// fp32 CUDA-core arithmetic, so that greedy-decoded ids can be compared bit for bit with the reference's fp32
// CPU run (BASELINE.json north_star: "Greedy-decoded token ids must be bit-exact on fp32 verification runs").
// Nothing here is on the training path and nothing is tuned beyond "finishes quickly": a 128x64x16 register-
// tiled SGEMM with a strided-A view (which makes the k-tap stride-2 convolutions of the channels-last feature
// encoder plain GEMMs -- k consecutive frames are contiguous in memory), row-wise LayerNorm / RMSNorm, a direct
// grouped positional convolution, a one-row-per-warp attention and an argmax sweep over logit chunks.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

namespace smx {
namespace f32 {

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == SMX_ACT_GELU) return gelu_exact(v);
  if (act == SMX_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// ------------------------------------------------------------------ SGEMM  C = act(alpha * A W^T + bias) + residual
constexpr int BM = 128, BN = 64, BK = 16;
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, long long lda, long long a_bs,
                                                       const float* __restrict__ W, const float* __restrict__ bias,
                                                       const float* __restrict__ R, long long ldr, long long r_bs,
                                                       float* __restrict__ C, long long ldc, long long c_bs, int M, int N,
                                                       int K, int act, float alpha) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int bz = blockIdx.z;
  A += (long long)bz * a_bs;
  C += (long long)bz * c_bs;
  if (R) R += (long long)bz * r_bs;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 8 x 4 outputs each
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      const int r = i / BK, c = i % BK;
      const int m = m0 + r, k = k0 + c;
      As[c][r] = (m < M && k < K) ? A[(long long)m * lda + k] : 0.f;
    }
    for (int i = threadIdx.x; i < BN * BK; i += 256) {
      const int r = i / BK, c = i % BK;
      const int n = n0 + r, k = k0 + c;
      Ws[c][r] = (n < N && k < K) ? W[(long long)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], w[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[n];
      v = apply_act(v, act);
      if (R) v += R[(long long)m * ldr + n];
      C[(long long)m * ldc + n] = v;
    }
  }
}

// ------------------------------------------------------------------ LayerNorm / RMSNorm (+ GELU), one warp per row
__global__ void __launch_bounds__(256) ln32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float* __restrict__ y, long long rows,
                                                   int cols, float eps, int rms_only, int act) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c];
  const float mean = rms_only ? 0.f : warp_sum(s) / cols;
  float v = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float d = xr[c] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / cols + eps);
  float* yr = y + row * cols;
  for (int c = lane; c < cols; c += 32) {
    float o = (xr[c] - mean) * rstd * gamma[c] + (beta ? beta[c] : 0.f);
    yr[c] = apply_act(o, act);
  }
}

// ------------------------------------------------------------------ GroupNorm(num_groups == channels) over time + GELU
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, long long t,
                                                       int channels, int t_chunk) {
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * t_chunk;
  const long long t1 = t0 + t_chunk < t ? t0 + t_chunk : t;
  for (int c = threadIdx.x; c < channels; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (long long i = t0; i < t1; ++i) {
      const double v = x[((long long)b * t + i) * channels + c];
      s += v;
      q += v * v;
    }
    atomicAdd(stats + ((long long)b * channels + c) * 2, s);
    atomicAdd(stats + ((long long)b * channels + c) * 2 + 1, q);
  }
}
__global__ void gn_apply_kernel(float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, long long t, int channels, long long total, float eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % channels);
    const long long b = i / ((long long)t * channels);
    const double mean = stats[(b * channels + c) * 2] / (double)t;
    double var = stats[(b * channels + c) * 2 + 1] / (double)t - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    x[i] = gelu_exact((x[i] - (float)mean) * rstd * gamma[c] + beta[c]);
  }
}

// ------------------------------------------------------------------ grouped positional convolution (direct)
// y[b,t,o] = x[b,t,o] * add_input + gelu(bias[o] + sum_{c,tap} w[o][c][tap] * x[b, t + tap - pad, g*cg + c])
constexpr int PC_T = 16;
__global__ void __launch_bounds__(256) posconv32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int t_len,
                                                        int hidden, int groups, int ksize, int add_input) {
  extern __shared__ float xs[];  // [(PC_T + ksize - 1)][cg]
  const int cg = hidden / groups, pad = ksize / 2;
  const int g = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * PC_T;
  const int rows = PC_T + ksize - 1;
  for (int i = threadIdx.x; i < rows * cg; i += blockDim.x) {
    const int r = i / cg, c = i % cg;
    const int t = t0 + r - pad;
    xs[i] = (t >= 0 && t < t_len) ? x[((long long)b * t_len + t) * hidden + g * cg + c] : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < PC_T * cg; idx += blockDim.x) {
    const int tt = idx / cg, o = idx % cg;
    const int t = t0 + tt;
    if (t >= t_len) continue;
    const float* wo = w + (long long)(g * cg + o) * cg * ksize;
    float acc = 0.f;
    for (int c = 0; c < cg; ++c) {
      const float* wc = wo + c * ksize;
      for (int tap = 0; tap < ksize; ++tap) acc = fmaf(wc[tap], xs[(tt + tap) * cg + c], acc);
    }
    const long long off = ((long long)b * t_len + t) * hidden + g * cg + o;
    const float v = gelu_exact(acc + bias[g * cg + o]);
    y[off] = add_input ? x[off] + v : v;
  }
}

// ------------------------------------------------------------------ attention, one warp per (b, h, q) row; head_dim 64
__global__ void __launch_bounds__(128) attn32_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                     const float* __restrict__ v, float* __restrict__ o, long long q_rs,
                                                     long long q_bs, long long k_rs, long long k_bs, long long v_rs,
                                                     long long v_bs, long long o_rs, long long o_bs, int heads, int tq, int tk,
                                                     int causal, float scale, const float* __restrict__ bias) {
  extern __shared__ float sm[];  // per warp: scores [tk] + q [64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 4 + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (qi >= tq) return;
  float* sc = sm + warp * (tk + 64);
  float* qs = sc + tk;
  const float* qp = q + b * q_bs + (long long)qi * q_rs + h * 64;
  qs[lane] = qp[lane];
  qs[lane + 32] = qp[lane + 32];
  __syncwarp();
  const int lim = causal ? qi + (tk - tq) : tk - 1;
  float mx = -INFINITY;
  for (int j = lane; j < tk; j += 32) {
    float s = -INFINITY;
    if (j <= lim) {
      const float* kp = k + b * k_bs + (long long)j * k_rs + h * 64;
      float acc = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) acc = fmaf(qs[d], kp[d], acc);
      s = acc * scale;
      if (bias) s += bias[((long long)h * tq + qi) * tk + j];
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float se = 0.f;
  for (int j = lane; j < tk; j += 32) {
    const float e = sc[j] == -INFINITY ? 0.f : expf(sc[j] - mx);
    sc[j] = e;
    se += e;
  }
  se = warp_sum(se);
  __syncwarp();
  const float inv = 1.f / se;
  float a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < tk; ++j) {
    const float pj = sc[j];
    const float* vp = v + b * v_bs + (long long)j * v_rs + h * 64;
    a0 = fmaf(pj, vp[lane], a0);
    a1 = fmaf(pj, vp[lane + 32], a1);
  }
  float* op = o + b * o_bs + (long long)qi * o_rs + h * 64;
  op[lane] = a0 * inv;
  op[lane + 32] = a1 * inv;
}

// ------------------------------------------------------------------ embeddings, argmax over logit chunks, axpy
__global__ void embed32_kernel(const long long* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                               const float* __restrict__ x_in, float* __restrict__ out, long long rows, int t_len, int dim,
                               float scale, long long pos_offset) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * dim; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / dim;
    const int c = (int)(i % dim);
    const int t = (int)(row % t_len);
    float v = 0.f;
    if (ids) v = tok[ids[row] * dim + c] * scale;
    if (x_in) v += x_in[i];
    if (pos) v += pos[(t + pos_offset) * dim + c];
    out[i] = v;
  }
}
__global__ void __launch_bounds__(256) argmax_chunk_kernel(const float* __restrict__ logits, long long ld, long long rows, int vn,
                                                           long long v0, float* __restrict__ best, long long* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* lr = logits + row * ld;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < vn; j += 32) {
    const float x = lr[j];
    if (x > bv) bv = x, bi = j;  // ascending j per lane: first occurrence wins inside a lane
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
  }
  if (lane == 0 && (v0 == 0 || bv > best[row])) {  // chunks arrive in ascending vocabulary order: strict > keeps the lowest index
    best[row] = bv;
    idx[row] = v0 + bi;
  }
}
__global__ void axpy32_kernel(const float* __restrict__ x, const float* __restrict__ w, int wi, float* __restrict__ y, long long n,
                              int first) {
  const float a = w[wi];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = first ? a * x[i] : fmaf(a, x[i], y[i]);
}

static int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace f32
}  // namespace smx

using namespace smx;
using namespace smx::f32;

extern "C" {

int smx_f32_gemm_nt(const float* a, int64_t lda, int64_t a_batch_stride, const float* w, const float* bias,
                    const float* residual, int64_t ldr, int64_t r_batch_stride, float* c, int64_t ldc,
                    int64_t c_batch_stride, int64_t m, int64_t n, int64_t k, int64_t batches, int act, float alpha,
                    void* stream) {
  SMX_REQUIRE(a && w && c && m > 0 && n > 0 && k > 0 && batches > 0, "f32_gemm_nt: bad arguments");
  dim3 grid((unsigned)ceil_div(n, BN), (unsigned)ceil_div(m, BM), (unsigned)batches);
  sgemm_nt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, lda, a_batch_stride, w, bias, residual, ldr, r_batch_stride, c,
                                                         ldc, c_batch_stride, (int)m, (int)n, (int)k, act, alpha);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_f32_layernorm(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int64_t cols,
                      float eps, int rms_only, int act, void* stream) {
  SMX_REQUIRE(x && gamma && y, "f32_layernorm: null pointer");
  ln32_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, rows, (int)cols, eps,
                                                                            rms_only, act);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_f32_groupnorm_gelu(float* x, double* stats_ws, const float* gamma, const float* beta, int64_t batch, int64_t t,
                           int64_t channels, float eps, void* stream) {
  SMX_REQUIRE(x && stats_ws && gamma && beta, "f32_groupnorm_gelu: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SMX_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, (size_t)batch * channels * 2 * sizeof(double), st));
  const int t_chunk = 512;
  dim3 g1((unsigned)ceil_div(t, t_chunk), (unsigned)batch);
  gn_stats_kernel<<<g1, 256, 0, st>>>(x, stats_ws, t, (int)channels, t_chunk);
  SMX_CHECK_CUDA(cudaGetLastError());
  const long long total = batch * t * channels;
  gn_apply_kernel<<<grid_for(total), 256, 0, st>>>(x, stats_ws, gamma, beta, t, (int)channels, total, eps);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_f32_posconv(const float* x, const float* w, const float* bias, float* y, int64_t batch, int64_t t,
                    int64_t hidden, int64_t groups, int64_t ksize, int add_input, void* stream) {
  SMX_REQUIRE(x && w && bias && y && hidden % groups == 0, "f32_posconv: bad arguments");
  const int cg = (int)(hidden / groups);
  const size_t smem = (size_t)(PC_T + ksize - 1) * cg * 4;
  SMX_REQUIRE(smem <= 48 * 1024, "f32_posconv: slab too large");
  dim3 grid((unsigned)ceil_div(t, PC_T), (unsigned)groups, (unsigned)batch);
  posconv32_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, bias, y, (int)t, (int)hidden, (int)groups, (int)ksize,
                                                             add_input);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_f32_attn(const float* q, const float* k, const float* v, float* o, int64_t q_rs, int64_t q_bs, int64_t k_rs,
                 int64_t k_bs, int64_t v_rs, int64_t v_bs, int64_t o_rs, int64_t o_bs, int64_t batch, int64_t heads,
                 int64_t tq, int64_t tk, int causal, float scale, const float* bias, void* stream) {
  SMX_REQUIRE(q && k && v && o, "f32_attn: null pointer");
  const size_t smem = (size_t)4 * (tk + 64) * 4;
  SMX_REQUIRE(smem <= 200 * 1024, "f32_attn: tk too large");
  static bool attr = false;
  if (!attr) {
    SMX_CHECK_CUDA(cudaFuncSetAttribute(attn32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid((unsigned)ceil_div(tq, 4), (unsigned)heads, (unsigned)batch);
  attn32_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(q, k, v, o, q_rs, q_bs, k_rs, k_bs, v_rs, v_bs, o_rs, o_bs,
                                                          (int)heads, (int)tq, (int)tk, causal, scale, bias);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_f32_embed(const int64_t* ids, const float* tok, const float* pos, const float* x_in, float* out, int64_t batch,
                  int64_t t, int64_t dim, float scale, int64_t pos_offset, void* stream) {
  SMX_REQUIRE(out && (ids == nullptr || tok != nullptr), "f32_embed: bad arguments");
  embed32_kernel<<<grid_for(batch * t * dim), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(ids), tok, pos, x_in, out, batch * t, (int)t, (int)dim, scale, pos_offset);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_f32_argmax_chunk(const float* logits, int64_t ld, int64_t rows, int64_t vn, int64_t v0, float* best, int64_t* idx,
                         void* stream) {
  SMX_REQUIRE(logits && best && idx, "f32_argmax_chunk: null pointer");
  argmax_chunk_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(logits, ld, rows, (int)vn, v0, best,
                                                                                    reinterpret_cast<long long*>(idx));
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_f32_axpy(const float* x, const float* w, int32_t wi, float* y, int64_t n, int first, void* stream) {
  axpy32_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, w, wi, y, n, first);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}
