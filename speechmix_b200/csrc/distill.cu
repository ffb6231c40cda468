// SpeechMixSelf auxiliary losses (ref:speechmix/hf_model.py:551-581) and the T5 relative position bias
// (hf:models/t5/modeling_t5.py:188-247).  All CUDA-core, HBM/L2-bound; the contractions that feed them
// (logit chunks) come from the tcgen05 GEMM.
//
//  * KL(softmax(teacher logits) || softmax(student logits)), reduction "batchmean", evaluated per
//    vocabulary chunk on fp32 logit chunks that stay L2-resident between the GEMM that wrote them and
//    the kernels here:      KL_row = sum_v p_t(v) (t_v - s_v) - lse_t + lse_s
//    backward chunk:        dlogit_s = coef_ce (p_s - onehot) + coef_kl (p_s - p_t)
//  * attention-projection MSE:  A = softmax(T . view(S, [D, Ts]) / sqrt(D)),  P = A . S,  mse(P, T),
//    where view() is the reference's MEMORY REINTERPRETATION of the [Ts, D] matrix (not a transpose).
//  * relative position bias gather / scatter through a host-built bucket table.
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

namespace smx {
namespace distill {

// ------------------------------------------------------------------ KL over a logit chunk
__global__ void __launch_bounds__(256) kl_chunk_fwd_kernel(const float* __restrict__ s, const float* __restrict__ t,
                                                           long long ld, long long rows, int vn,
                                                           const float* __restrict__ lse_t, float* __restrict__ cross) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * ld;
  const float* tr = t + row * ld;
  const float lt = lse_t[row];
  float acc = 0.f;
  for (int v = lane * 4; v + 4 <= vn; v += 128) {
    const float4 a = *reinterpret_cast<const float4*>(sr + v);
    const float4 b = *reinterpret_cast<const float4*>(tr + v);
    acc = fmaf(__expf(b.x - lt), b.x - a.x, acc);
    acc = fmaf(__expf(b.y - lt), b.y - a.y, acc);
    acc = fmaf(__expf(b.z - lt), b.z - a.z, acc);
    acc = fmaf(__expf(b.w - lt), b.w - a.w, acc);
  }
  for (int v = (vn & ~3) + lane; v < vn; v += 32) acc = fmaf(__expf(tr[v] - lt), tr[v] - sr[v], acc);
  acc = warp_sum(acc);
  if (lane == 0) cross[row] += acc;
}

__global__ void kl_finalize_kernel(const float* __restrict__ cross, const float* __restrict__ lse_s,
                                   const float* __restrict__ lse_t, long long rows, float inv_batch,
                                   float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  float acc = 0.f;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) acc += cross[r] - lse_t[r] + lse_s[r];
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v * inv_batch;
  }
}

__global__ void __launch_bounds__(256) kl_chunk_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t,
                                                           long long ld, long long rows, int vn, long long v0,
                                                           const long long* __restrict__ labels,
                                                           const float* __restrict__ lse_s, const float* __restrict__ lse_t,
                                                           const float* __restrict__ coef_ce, const float* __restrict__ coef_kl,
                                                           bf16* __restrict__ dl, long long ld_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * ld;
  const float* tr = t + row * ld;
  bf16* dr = dl + row * ld_out;
  const float ls = lse_s[row], lt = lse_t[row], cc = coef_ce[row], ck = coef_kl[0];
  const long long lab = labels[row] - v0;
  for (int v = lane; v < vn; v += 32) {
    const float ps = __expf(sr[v] - ls);
    const float pt = __expf(tr[v] - lt);
    float g = cc * ps + ck * (ps - pt);
    if (v == lab) g -= cc;
    dr[v] = __float2bfloat16(g);
  }
}

// ------------------------------------------------------------------ attention-projection MSE
// forward: one CTA per (batch, text row i).  smem: T row [D], scores / probabilities [Ts]
__global__ void __launch_bounds__(256) mse_fwd_kernel(const bf16* __restrict__ T, const bf16* __restrict__ S,
                                                      float* __restrict__ A, float* __restrict__ diff,
                                                      float* __restrict__ loss, int Tt, int Ts, int D, float inv_sqrt_d,
                                                      float inv_n) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* trow = sm;          // [D]
  float* prob = sm + D;      // [Ts]
  __shared__ float red[32];
  const int b = blockIdx.y, i = blockIdx.x;
  const bf16* Tb = T + ((long long)b * Tt + i) * D;
  const bf16* Sb = S + (long long)b * Ts * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) trow[d] = __bfloat162float(Tb[d]);
  __syncthreads();
  // scores[j] = sum_d T[i][d] * Sflat[d*Ts + j]
  for (int j = threadIdx.x; j < Ts; j += blockDim.x) {
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(trow[d], __bfloat162float(Sb[(long long)d * Ts + j]), acc);
    prob[j] = acc * inv_sqrt_d;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < Ts; j += blockDim.x) mx = fmaxf(mx, prob[j]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float se = 0.f;
  for (int j = threadIdx.x; j < Ts; j += blockDim.x) {
    const float e = __expf(prob[j] - mx);
    prob[j] = e;
    se += e;
  }
  se = warp_sum(se);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
  __syncthreads();
  se = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) se += red[w];
  const float inv = 1.f / se;
  __syncthreads();
  float* Ar = A + ((long long)b * Tt + i) * Ts;
  for (int j = threadIdx.x; j < Ts; j += blockDim.x) {
    prob[j] *= inv;
    Ar[j] = prob[j];
  }
  __syncthreads();
  float sq = 0.f;
  float* dr = diff + ((long long)b * Tt + i) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Ts; ++j) acc = fmaf(prob[j], __bfloat162float(Sb[(long long)j * D + c]), acc);
    const float df = acc - trow[c];
    dr[c] = df;
    sq = fmaf(df, df, sq);
  }
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) v += red[w];
    atomicAdd(loss, v * inv_n);
  }
}

// backward part 1: dscores[b][i][:] (pre-softmax score gradient incl. the 1/sqrt(D)), one CTA per (b, i)
__global__ void __launch_bounds__(256) mse_bwd_scores_kernel(const bf16* __restrict__ S, const float* __restrict__ A,
                                                             const float* __restrict__ diff, float* __restrict__ dsc,
                                                             int Tt, int Ts, int D, float inv_sqrt_d) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* drow = sm;       // [D] diff row
  float* da = sm + D;     // [Ts]
  __shared__ float red[32];
  const int b = blockIdx.y, i = blockIdx.x;
  const bf16* Sb = S + (long long)b * Ts * D;
  const float* dr = diff + ((long long)b * Tt + i) * D;
  const float* Ar = A + ((long long)b * Tt + i) * Ts;
  for (int d = threadIdx.x; d < D; d += blockDim.x) drow[d] = dr[d];
  __syncthreads();
  // dA[j] = sum_c diff[c] * S[j][c]   (warp per j: coalesced over c)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < Ts; j += nw) {
    float acc = 0.f;
    for (int c = lane; c < D; c += 32) acc = fmaf(drow[c], __bfloat162float(Sb[(long long)j * D + c]), acc);
    acc = warp_sum(acc);
    if (lane == 0) da[j] = acc;
  }
  __syncthreads();
  float dot = 0.f;
  for (int j = threadIdx.x; j < Ts; j += blockDim.x) dot = fmaf(Ar[j], da[j], dot);
  dot = warp_sum(dot);
  if (lane == 0) red[warp] = dot;
  __syncthreads();
  dot = 0.f;
  for (int w = 0; w < nw; ++w) dot += red[w];
  float* out = dsc + ((long long)b * Tt + i) * Ts;
  for (int j = threadIdx.x; j < Ts; j += blockDim.x) out[j] = Ar[j] * (da[j] - dot) * inv_sqrt_d;
}

// backward part 2: dS[b][f] for the flat index f of the [Ts, D] matrix:
//   through P = A.S       : f = j*D + c  -> sum_i A[i][j] * diff[i][c]
//   through view(S,[D,Ts]) : f = d*Ts + j -> sum_i T[i][d] * dsc[i][j]
__global__ void __launch_bounds__(256) mse_bwd_ds_kernel(const bf16* __restrict__ T, const float* __restrict__ A,
                                                         const float* __restrict__ diff, const float* __restrict__ dsc,
                                                         bf16* __restrict__ dS, int Tt, int Ts, int D,
                                                         const float* __restrict__ gscale) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= (long long)Ts * D) return;
  const int j1 = (int)(f / D), c = (int)(f % D);
  const int d = (int)(f / Ts), j2 = (int)(f % Ts);
  const float* Ab = A + (long long)b * Tt * Ts;
  const float* db = diff + (long long)b * Tt * D;
  const float* sb = dsc + (long long)b * Tt * Ts;
  const bf16* Tb = T + (long long)b * Tt * D;
  float acc = 0.f;
  for (int i = 0; i < Tt; ++i) {
    acc = fmaf(Ab[(long long)i * Ts + j1], db[(long long)i * D + c], acc);
    acc = fmaf(__bfloat162float(Tb[(long long)i * D + d]), sb[(long long)i * Ts + j2], acc);
  }
  dS[(long long)b * Ts * D + f] = __float2bfloat16(acc * gscale[0]);
}

// ------------------------------------------------------------------ relative position bias
__global__ void relpos_fwd_kernel(const float* __restrict__ w, const int* __restrict__ table, float* __restrict__ bias,
                                  int heads, int tq, int tk, int q_offset) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)heads * tq * tk;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % tk);
    const int i = (int)((e / tk) % tq);
    const int h = (int)(e / ((long long)tk * tq));
    const int bucket = table[j - (i + q_offset) + (tq + q_offset - 1)];
    bias[e] = w[bucket * heads + h];
  }
}
__global__ void __launch_bounds__(256) relpos_bwd_kernel(const float* __restrict__ dbias, const int* __restrict__ table,
                                                         float* __restrict__ dw, int heads, int tq, int tk, int q_offset,
                                                         int n_buckets) {
  pdl_trigger();
  pdl_wait();
  __shared__ float hist[256];
  const int h = blockIdx.y;
  for (int i = threadIdx.x; i < n_buckets; i += blockDim.x) hist[i] = 0.f;
  __syncthreads();
  const long long n = (long long)tq * tk;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % tk), i = (int)(e / tk);
    atomicAdd(&hist[table[j - (i + q_offset) + (tq + q_offset - 1)]], dbias[(long long)h * n + e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_buckets; i += blockDim.x) atomicAdd(dw + i * heads + h, hist[i]);
}

}  // namespace distill
}  // namespace smx

using namespace smx;
using namespace smx::distill;


namespace smx {
namespace gan {
__device__ __forceinline__ void load8(const bf16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) f[2 * j] = bf16_lo(w[j]), f[2 * j + 1] = bf16_hi(w[j]);
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                            pack_bf16x2(f[6], f[7]));
}
// SpeechMixGAN discriminator input (ref:speechmix/hf_model.py:637-686): the reference flattens the Gram-like matrix
//   G_b = X_b.view(D, T) . X_b.view(T, D)        (a MEMORY REINTERPRETATION of the [T, D] states, not a transpose)
// into D*D features and applies Linear(D*D, 1).  With W = weight.view(D, D):
//   logit_b = <W, G_b> = sum_{t,i} Xflat_b[i T + t] * Z_b[t, i],     Z = X W^T   (one GEMM; G never exists).
// These kernels are that last contraction and its two gradients.
__global__ void __launch_bounds__(256) gram_dot_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ z,
                                                           float* __restrict__ out, int t, int d) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  const int b = blockIdx.y;
  const long long n = (long long)t * d;
  const bf16* xb = x + b * n;
  const bf16* zb = z + b * n;
  float acc = 0.f;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; e < n; e += (long long)gridDim.x * blockDim.x * 8) {
    const int tt = (int)(e / d), i0 = (int)(e % d);          // 8 consecutive i of one frame of Z
    float zv[8];
    load8(zb + e, zv);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(zv[k], __bfloat162float(xb[(long long)(i0 + k) * t + tt]), acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(out + b, s);
  }
}
// dz[b, t, i] = g[b] * Xflat_b[i T + t];   dxflat[b, f] = g[b] * Z_b[f % T, f / T]
__global__ void __launch_bounds__(256) gram_dot_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ z,
                                                           const float* __restrict__ g, bf16* __restrict__ dx,
                                                           bf16* __restrict__ dz, int t, int d) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const long long n = (long long)t * d;
  const bf16* xb = x + b * n;
  const bf16* zb = z + b * n;
  const float gb = g[b];
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; e < n; e += (long long)gridDim.x * blockDim.x * 8) {
    const int tt = (int)(e / d), i0 = (int)(e % d);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = gb * __bfloat162float(xb[(long long)(i0 + k) * t + tt]);
    store8(dz + b * n + e, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long f = e + k;                              // flat index of X: i = f / T, frame = f % T
      o[k] = gb * __bfloat162float(zb[(f % t) * d + f / t]);
    }
    store8(dx + b * n + e, o);
  }
}
}  // namespace gan
}  // namespace smx

extern "C" {

int smx_kl_chunk_fwd(const float* s, const float* t, int64_t ld, int64_t rows, int64_t vn, const float* lse_t,
                     float* cross, void* stream) {
  SMX_REQUIRE(s && t && lse_t && cross && ld % 4 == 0, "kl_chunk_fwd: bad arguments");
  launch_pdl(kl_chunk_fwd_kernel, dim3((unsigned)ceil_div(rows, 8)), dim3(256), 0, (cudaStream_t)stream, s, t, ld, rows, (int)vn, lse_t, cross);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_kl_finalize(const float* cross, const float* lse_s, const float* lse_t, int64_t rows, float inv_batch,
                    float* out, void* stream) {
  launch_pdl(kl_finalize_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, cross, lse_s, lse_t, rows, inv_batch, out);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_kl_chunk_bwd(const float* s, const float* t, int64_t ld, int64_t rows, int64_t vn, int64_t v0,
                     const int64_t* labels, const float* lse_s, const float* lse_t, const float* coef_ce,
                     const float* coef_kl, void* dlogits, int64_t ld_out, void* stream) {
  SMX_REQUIRE(s && t && labels && lse_s && lse_t && coef_ce && coef_kl && dlogits, "kl_chunk_bwd: null pointer");
  launch_pdl(kl_chunk_bwd_kernel, dim3((unsigned)ceil_div(rows, 8)), dim3(256), 0, (cudaStream_t)stream, 
      s, t, ld, rows, (int)vn, v0, reinterpret_cast<const long long*>(labels), lse_s, lse_t, coef_ce, coef_kl,
      reinterpret_cast<bf16*>(dlogits), ld_out);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_self_mse_fwd(const void* text_h, const void* speech_h, float* attn, float* diff, float* loss, int64_t batch,
                     int64_t tt, int64_t ts, int64_t dim, void* stream) {
  SMX_REQUIRE(text_h && speech_h && attn && diff && loss, "self_mse_fwd: null pointer");
  const size_t smem = (size_t)(dim + ts) * 4;
  SMX_REQUIRE(smem <= 48 * 1024, "self_mse_fwd: dim + ts too large (%lld)", (long long)(dim + ts));
  dim3 grid((unsigned)tt, (unsigned)batch);
  launch_pdl(mse_fwd_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, 
      reinterpret_cast<const bf16*>(text_h), reinterpret_cast<const bf16*>(speech_h), attn, diff, loss, (int)tt, (int)ts,
      (int)dim, 1.0f / sqrtf((float)dim), 1.0f / (float)(batch * tt * dim));
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_self_mse_bwd(const void* text_h, const void* speech_h, const float* attn, const float* diff, float* dscores,
                     const float* gscale, void* d_speech_h, int64_t batch, int64_t tt, int64_t ts, int64_t dim,
                     void* stream) {
  SMX_REQUIRE(text_h && speech_h && attn && diff && dscores && gscale && d_speech_h, "self_mse_bwd: null pointer");
  const size_t smem = (size_t)(dim + ts) * 4;
  SMX_REQUIRE(smem <= 48 * 1024, "self_mse_bwd: dim + ts too large");
  dim3 grid((unsigned)tt, (unsigned)batch);
  launch_pdl(mse_bwd_scores_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, reinterpret_cast<const bf16*>(speech_h), attn, diff,
                                                                  dscores, (int)tt, (int)ts, (int)dim,
                                                                  1.0f / sqrtf((float)dim));
  SMX_CHECK_CUDA(cudaGetLastError());
  dim3 g2((unsigned)ceil_div(ts * dim, 256), (unsigned)batch);
  launch_pdl(mse_bwd_ds_kernel, dim3(g2), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const bf16*>(text_h), attn, diff, dscores,
                                                         reinterpret_cast<bf16*>(d_speech_h), (int)tt, (int)ts, (int)dim,
                                                         gscale);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_relpos_bias_fwd(const float* weight, const int32_t* table, float* bias, int64_t heads, int64_t tq, int64_t tk,
                        int64_t q_offset, void* stream) {
  SMX_REQUIRE(weight && table && bias, "relpos_bias_fwd: null pointer");
  const long long n = heads * tq * tk;
  long long g = ceil_div(n, 256);
  if (g > 148 * 16) g = 148 * 16;
  launch_pdl(relpos_fwd_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, weight, table, bias, (int)heads, (int)tq, (int)tk,
                                                                  (int)q_offset);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_relpos_bias_bwd(const float* dbias, const int32_t* table, float* dweight, int64_t heads, int64_t tq, int64_t tk,
                        int64_t q_offset, int64_t n_buckets, void* stream) {
  SMX_REQUIRE(dbias && table && dweight && n_buckets <= 256, "relpos_bias_bwd: bad arguments");
  long long g = ceil_div(tq * tk, 256 * 8);
  if (g < 1) g = 1;
  if (g > 64) g = 64;
  dim3 grid((unsigned)g, (unsigned)heads);
  launch_pdl(relpos_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, dbias, table, dweight, (int)heads, (int)tq, (int)tk,
                                                           (int)q_offset, (int)n_buckets);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_gram_dot_fwd(const void* x, const void* z, float* out, int64_t batch, int64_t t, int64_t dim, void* stream) {
  using namespace smx;
  SMX_REQUIRE(x && z && out && dim % 8 == 0 && batch > 0 && t > 0, "gram_dot_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SMX_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * batch, st));
  long long gx = ceil_div(t * dim, 256 * 8);
  if (gx > 64) gx = 64;
  launch_pdl(gan::gram_dot_fwd_kernel, dim3((unsigned)gx, (unsigned)batch), dim3(256), 0, st, (const bf16*)x, (const bf16*)z, out,
             (int)t, (int)dim);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int smx_gram_dot_bwd(const void* x, const void* z, const float* g, void* dx, void* dz, int64_t batch, int64_t t, int64_t dim,
                     void* stream) {
  using namespace smx;
  SMX_REQUIRE(x && z && g && dx && dz && dim % 8 == 0 && batch > 0 && t > 0, "gram_dot_bwd: bad arguments");
  long long gx = ceil_div(t * dim, 256 * 8);
  if (gx > 64) gx = 64;
  launch_pdl(gan::gram_dot_bwd_kernel, dim3((unsigned)gx, (unsigned)batch), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x,
             (const bf16*)z, g, (bf16*)dx, (bf16*)dz, (int)t, (int)dim);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}
