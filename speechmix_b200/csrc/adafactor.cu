// Multi-tensor Adafactor step (the optimizer of the reference recipe: ref:train.py:298 `optim="adafactor"`, i.e.
// transformers.optimization.Adafactor with relative_step=False, scale_parameter=False, beta1=None as the HF Trainer
// configures it).  All parameters of the model are updated by SIX launches (four over a tile table, two block-per-slice)
// instead of the ~15 small kernels per parameter of the eager implementation (~7000 launches per step for
// wav2vec2-base + bart-base):
//   1 stats     per 64 x 256 tile: row / column sums of g^2 -> fp32 atomics into per-tensor accumulators
//   2 finalize  per (tensor, leading index): EMA of the factored second moments, mean of the row moments
//   3 sumsq     per tile: u = g * rsqrt(row / mean(row)) * rsqrt(col)  (or g * rsqrt(v) for vectors), sum u^2
//   4 apply     per tile: p <- p (1 - wd lr) - lr u / max(1, rms(u) / clip)
// Factored slices of at most 16384 elements (the [512, 512, 3] / [512, 512, 2] / [512, 1, 10] convolution weights of the
// feature encoder and the [768, 48, 128] positional convolution: one leading index = one tiny [rows][cols] matrix) do
// not go through the 64 x 256 tile table -- a [512][3] slice would be eight tiles of 192 useful elements each, and
// those near-empty tiles were 64 % of all tiles of the wav2vec2-base + bart-base model.  They take two block-per-slice
// kernels instead: `small_moments` (slice in shared memory: row / column sums, both EMAs, mean of the row moments and
// the slice's share of sum(u^2) -- steps 1-3 with ONE read of g and no atomics but the last) and `small_apply`.
// HBM-bound: g is read three times and p once read / once written (20 bytes per parameter); a tensor with
// len(shape) >= 2 is factored over its last two dims exactly like the reference (leading dims = independent slices).
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

namespace smx {
namespace adafactor {

constexpr int TILE_R = 64, TILE_C = 256;   // 8 warps x 8 rows, 32 lanes x 8 columns (lane + 32 j: coalesced)

struct Hyper {
  float beta2t, eps1, lr, clip, weight_decay;
  // capturable mode (whole-step CUDA graph): the step count lives on the device and beta2(t) = 1 - t^decay_rate is
  // evaluated by the kernels, so a replayed graph advances the schedule (a host-computed beta2t would be frozen)
  const long long* step_dev;
  float decay_rate;
};
__device__ __forceinline__ Hyper resolve(Hyper h) {
  if (h.step_dev) h.beta2t = 1.0f - powf((float)(*h.step_dev), h.decay_rate);
  return h;
}
__global__ void bump_step_kernel(long long* step) {
  pdl_trigger();
  pdl_wait();
  *step += 1;
}

// element (r, c) of slice b of a tensor; vectors are viewed as [ceil(n / 256)][256] with a ragged last row
__device__ __forceinline__ bool in_range(const SmxAdafactorTensor& t, long long r, long long c) {
  return t.factored ? (r < t.rows && c < t.cols) : (r * TILE_C + c < t.numel);
}
__device__ __forceinline__ long long offset(const SmxAdafactorTensor& t, int b, long long r, long long c) {
  return t.factored ? ((long long)b * t.rows + r) * t.cols + c : r * TILE_C + c;
}

// Fast path of a tile: factored tensor, cols % 4 == 0 -> every lane owns two float4 per row (columns 4 lane + 128 j),
// a warp reads 512 contiguous bytes per request, and the 16 loads of a thread (8 rows x 2) are all in flight before the
// first use.  (The first version issued 64 scalar loads per thread, each behind a 64-bit index computation, and the
// update kernel interleaved them with stores: ~1 TB/s, profiles/r01z_adafactor.txt.)
__device__ __forceinline__ bool vec_tile(const SmxAdafactorTensor& t) { return t.factored && (t.cols & 3) == 0; }

__global__ void __launch_bounds__(256) stats_kernel(const SmxAdafactorTensor* __restrict__ tensors,
                                                    const SmxAdafactorTile* __restrict__ tiles) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][TILE_C + 1];
  const SmxAdafactorTile tl = tiles[blockIdx.x];
  const SmxAdafactorTensor t = tensors[tl.tensor];
  if (!t.factored) return;   // vectors need no factored statistics (their tiles are skipped uniformly)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float colp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (vec_tile(t)) {
    const float* __restrict__ gp = t.g + ((long long)tl.b * t.rows) * t.cols;
    float4 v[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long r = tl.r0 + warp + 8 * k;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const long long c = tl.c0 + 4 * lane + 128 * j;
        v[k][j] = (r < t.rows && c < t.cols) ? __ldg(reinterpret_cast<const float4*>(gp + r * t.cols + c))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long r = tl.r0 + warp + 8 * k;
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 g = v[k][j];
        colp[4 * j + 0] = fmaf(g.x, g.x, colp[4 * j + 0]);
        colp[4 * j + 1] = fmaf(g.y, g.y, colp[4 * j + 1]);
        colp[4 * j + 2] = fmaf(g.z, g.z, colp[4 * j + 2]);
        colp[4 * j + 3] = fmaf(g.w, g.w, colp[4 * j + 3]);
        rs += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
      }
      rs = warp_sum(rs);
      if (lane == 0 && r < t.rows) atomicAdd(t.row_acc + (long long)tl.b * t.rows + r, rs);
    }
    // column c0 + 4 lane + 128 j + e lives in colp[4 j + e]
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[warp][4 * lane + 128 * j + e] = colp[4 * j + e];
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long r = tl.r0 + warp + 8 * k;
      float rs = 0.f;
      if (r < t.rows) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const long long c = tl.c0 + lane + 32 * j;
          if (c < t.cols) {
            const float g = t.g[offset(t, tl.b, r, c)];
            rs = fmaf(g, g, rs);
            colp[j] = fmaf(g, g, colp[j]);
          }
        }
      }
      rs = warp_sum(rs);
      if (lane == 0 && r < t.rows) atomicAdd(t.row_acc + (long long)tl.b * t.rows + r, rs);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane + 32 * j] = colp[j];
  }
  __syncthreads();
  float cs = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) cs += red[w][threadIdx.x];
  const long long c = tl.c0 + threadIdx.x;
  if (c < t.cols) atomicAdd(t.col_acc + (long long)tl.b * t.cols + c, cs);
}

// one block per (factored tensor, leading index): exp_avg_sq_row / exp_avg_sq_col EMAs and the mean of the row moments
__global__ void __launch_bounds__(256) finalize_kernel(const SmxAdafactorTensor* __restrict__ tensors,
                                                       const SmxAdafactorSlice* __restrict__ slices, Hyper hp) {
  pdl_trigger();
  pdl_wait();
  const Hyper h = resolve(hp);
  __shared__ float red[8];
  const SmxAdafactorSlice s = slices[blockIdx.x];
  const SmxAdafactorTensor t = tensors[s.tensor];
  float* row = t.row + (long long)s.b * t.rows;
  float* col = t.col + (long long)s.b * t.cols;
  const float* racc = t.row_acc + (long long)s.b * t.rows;
  const float* cacc = t.col_acc + (long long)s.b * t.cols;
  const float inv_c = 1.0f / (float)t.cols, inv_r = 1.0f / (float)t.rows, om = 1.0f - h.beta2t;
  float sum = 0.f;
  for (long long r = threadIdx.x; r < t.rows; r += blockDim.x) {
    const float v = h.beta2t * row[r] + om * (racc[r] * inv_c + h.eps1);   // mean(g^2 + eps1) over the last dim
    row[r] = v;
    sum += v;
  }
  for (long long c = threadIdx.x; c < t.cols; c += blockDim.x)
    col[c] = h.beta2t * col[c] + om * (cacc[c] * inv_r + h.eps1);          // mean over dim -2
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    t.rmean[s.b] = tot * inv_r;
  }
}

// APPLY = false: accumulate sum(u^2) per tensor (and update exp_avg_sq of vectors); APPLY = true: write the parameters
template <bool APPLY>
__global__ void __launch_bounds__(256) update_kernel(const SmxAdafactorTensor* __restrict__ tensors,
                                                     const SmxAdafactorTile* __restrict__ tiles, Hyper hp) {
  pdl_trigger();
  pdl_wait();
  const Hyper h = resolve(hp);
  __shared__ float red[8];
  const SmxAdafactorTile tl = tiles[blockIdx.x];
  const SmxAdafactorTensor t = tensors[tl.tensor];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float cf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long c = tl.c0 + lane + 32 * j;
    cf[j] = (t.factored && !vec_tile(t) && c < t.cols) ? 1.0f / sqrtf(t.col[(long long)tl.b * t.cols + c]) : 0.f;
  }
  float scale = 0.f, decay = 1.f;
  if (APPLY) {
    const float rms = sqrtf(*t.sumsq / (float)t.numel);
    scale = h.lr / fmaxf(1.0f, rms / h.clip);
    decay = 1.0f - h.weight_decay * h.lr;
  }
  const float inv_rmean = t.factored ? 1.0f / t.rmean[tl.b] : 0.f;
  const float om = 1.0f - h.beta2t;
  float ss = 0.f;
  if (vec_tile(t)) {
    // all loads of the tile first (g, and p when applying), then the arithmetic, then the stores
    const float* __restrict__ gp = t.g + ((long long)tl.b * t.rows) * t.cols;
    float* __restrict__ pp = t.p + ((long long)tl.b * t.rows) * t.cols;
    float4 cf4[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long c = tl.c0 + 4 * lane + 128 * j;
      if (c < t.cols) {
        const float4 cv = *reinterpret_cast<const float4*>(t.col + (long long)tl.b * t.cols + c);
        cf4[j] = make_float4(rsqrtf(cv.x), rsqrtf(cv.y), rsqrtf(cv.z), rsqrtf(cv.w));
      } else {
        cf4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // two halves of four rows: 8 (+8) float4 loads in flight per thread at ~90 registers, so two to three blocks share
    // an SM (all sixteen rows at once needed 154 registers = one block per SM)
#pragma unroll 1
    for (int kh = 0; kh < 2; ++kh) {
      float4 gv[4][2], pv[4][2];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long r = tl.r0 + warp + 8 * (4 * kh + k);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const long long c = tl.c0 + 4 * lane + 128 * j;
          const bool ok = r < t.rows && c < t.cols;
          gv[k][j] = ok ? __ldg(reinterpret_cast<const float4*>(gp + r * t.cols + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (APPLY) pv[k][j] = ok ? *reinterpret_cast<const float4*>(pp + r * t.cols + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long r = tl.r0 + warp + 8 * (4 * kh + k);
        const float rf = r < t.rows ? rsqrtf(t.row[(long long)tl.b * t.rows + r] * inv_rmean) : 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const long long c = tl.c0 + 4 * lane + 128 * j;
          float4 u = gv[k][j];
          u.x *= rf * cf4[j].x, u.y *= rf * cf4[j].y, u.z *= rf * cf4[j].z, u.w *= rf * cf4[j].w;
          if (APPLY) {
            if (r < t.rows && c < t.cols) {
              float4 q = pv[k][j];
              q.x = q.x * decay - scale * u.x, q.y = q.y * decay - scale * u.y;
              q.z = q.z * decay - scale * u.z, q.w = q.w * decay - scale * u.w;
              *reinterpret_cast<float4*>(pp + r * t.cols + c) = q;
            }
          } else {
            ss += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
          }
        }
      }
    }
  } else
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long r = tl.r0 + warp + 8 * k;
    float rf = 0.f;
    if (t.factored && r < t.rows) rf = 1.0f / sqrtf(t.row[(long long)tl.b * t.rows + r] * inv_rmean);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long c = tl.c0 + lane + 32 * j;
      if (!in_range(t, r, c)) continue;
      const long long o = offset(t, tl.b, r, c);
      const float g = t.g[o];
      float u;
      if (t.factored) {
        u = g * (rf * cf[j]);
      } else {
        float v = t.row[o];                       // exp_avg_sq of a vector lives in `row`
        if (!APPLY) {
          v = h.beta2t * v + om * (g * g + h.eps1);
          t.row[o] = v;
        }
        u = g / sqrtf(v);
      }
      if (APPLY) t.p[o] = t.p[o] * decay - scale * u;
      else ss = fmaf(u, u, ss);
    }
  }
  if (!APPLY) {
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      atomicAdd(t.sumsq, tot);
    }
  }
}


// ---- block-per-slice path for small factored slices (rows * cols <= SMALL_ELEMS, rows + cols <= SMALL_RC)
constexpr int SMALL_ELEMS = 16384, SMALL_RC = 4096;

__global__ void __launch_bounds__(256) small_moments_kernel(const SmxAdafactorTensor* __restrict__ tensors,
                                                            const SmxAdafactorSlice* __restrict__ slices, Hyper hp) {
  pdl_trigger();
  pdl_wait();
  const Hyper h = resolve(hp);
  extern __shared__ float sm[];                 // g [rows * cols] | rowv [rows] | colv [cols] | red [8]
  const SmxAdafactorSlice s = slices[blockIdx.x];
  const SmxAdafactorTensor t = tensors[s.tensor];
  const int rows = (int)t.rows, cols = (int)t.cols, n = rows * cols;
  float* gs = sm;
  float* rowv = sm + n;
  float* colv = rowv + rows;
  float* red = colv + cols;
  const float* __restrict__ g = t.g + (long long)s.b * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
    for (int i = threadIdx.x * 4; i < n; i += 1024) *reinterpret_cast<float4*>(gs + i) = __ldg(reinterpret_cast<const float4*>(g + i));
  } else {
    for (int i = threadIdx.x; i < n; i += 256) gs[i] = __ldg(g + i);
  }
  __syncthreads();
  const float om = 1.0f - h.beta2t, inv_c = 1.0f / (float)cols, inv_r = 1.0f / (float)rows;
  float* row = t.row + (long long)s.b * rows;
  float* col = t.col + (long long)s.b * cols;
  // row moments: narrow rows (cols <= 8) one thread per row, otherwise one warp per row
  float rsum = 0.f;
  if (cols <= 8) {
    for (int r = threadIdx.x; r < rows; r += 256) {
      float a = 0.f;
      for (int c = 0; c < cols; ++c) a = fmaf(gs[r * cols + c], gs[r * cols + c], a);
      const float v = h.beta2t * row[r] + om * (a * inv_c + h.eps1);
      row[r] = v, rowv[r] = v, rsum += v;
    }
  } else {
    for (int r = warp; r < rows; r += 8) {
      float a = 0.f;
      for (int c = lane; c < cols; c += 32) a = fmaf(gs[r * cols + c], gs[r * cols + c], a);
      a = warp_sum(a);
      if (lane == 0) {
        const float v = h.beta2t * row[r] + om * (a * inv_c + h.eps1);
        row[r] = v, rowv[r] = v, rsum += v;
      }
    }
  }
  // column moments: wide slices one thread per column; narrow ones one warp per column
  if (cols >= 32) {
    for (int c = threadIdx.x; c < cols; c += 256) {
      float a = 0.f;
      for (int r = 0; r < rows; ++r) a = fmaf(gs[r * cols + c], gs[r * cols + c], a);
      const float v = h.beta2t * col[c] + om * (a * inv_r + h.eps1);
      col[c] = v, colv[c] = v;
    }
  } else {
    for (int c = warp; c < cols; c += 8) {
      float a = 0.f;
      for (int r = lane; r < rows; r += 32) a = fmaf(gs[r * cols + c], gs[r * cols + c], a);
      a = warp_sum(a);
      if (lane == 0) {
        const float v = h.beta2t * col[c] + om * (a * inv_r + h.eps1);
        col[c] = v, colv[c] = v;
      }
    }
  }
  rsum = warp_sum(rsum);
  if (lane == 0) red[warp] = rsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float rmean = tot * inv_r;
  if (threadIdx.x == 0) t.rmean[s.b] = rmean;
  __syncthreads();                                // red is reused below
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int r = i / cols, c = i - r * cols;
    const float u = gs[i] * rsqrtf(rowv[r] / rmean) * rsqrtf(colv[c]);
    ss = fmaf(u, u, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w];
    atomicAdd(t.sumsq, a);
  }
}

__global__ void __launch_bounds__(256) small_apply_kernel(const SmxAdafactorTensor* __restrict__ tensors,
                                                          const SmxAdafactorSlice* __restrict__ slices, Hyper hp) {
  pdl_trigger();
  pdl_wait();
  const Hyper h = resolve(hp);
  const SmxAdafactorSlice s = slices[blockIdx.x];
  const SmxAdafactorTensor t = tensors[s.tensor];
  const int rows = (int)t.rows, cols = (int)t.cols, n = rows * cols;
  const float rms = sqrtf(*t.sumsq / (float)t.numel);
  const float scale = h.lr / fmaxf(1.0f, rms / h.clip), decay = 1.0f - h.weight_decay * h.lr;
  const float inv_rmean = 1.0f / t.rmean[s.b];
  const float* __restrict__ g = t.g + (long long)s.b * n;
  float* __restrict__ pp = t.p + (long long)s.b * n;
  const float* __restrict__ row = t.row + (long long)s.b * rows;
  const float* __restrict__ col = t.col + (long long)s.b * cols;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int r = i / cols, c = i - r * cols;
    const float u = __ldg(g + i) * rsqrtf(row[r] * inv_rmean) * rsqrtf(col[c]);
    pp[i] = pp[i] * decay - scale * u;
  }
}

}  // namespace adafactor
}  // namespace smx

extern "C" int smx_adafactor_step(const SmxAdafactorTensor* tensors, int32_t n_tensors, const SmxAdafactorTile* tiles,
                                  int32_t n_tiles, const SmxAdafactorSlice* slices, int32_t n_slices,
                                  const SmxAdafactorSlice* small_slices, int32_t n_small, int32_t small_smem_floats,
                                  void* scratch, int64_t scratch_bytes, float beta2t, float eps1, float lr,
                                  float clip_threshold, float weight_decay, int64_t* step_dev, float decay_rate,
                                  void* stream) {
  using namespace smx;
  using namespace smx::adafactor;
  SMX_REQUIRE(tensors && (n_tiles == 0 || tiles) && (n_slices == 0 || slices) && (n_small == 0 || small_slices) && scratch,
              "adafactor: null table");
  if (n_tensors <= 0 || (n_tiles <= 0 && n_small <= 0)) return 0;
  const size_t small_smem = ((size_t)small_smem_floats + 8) * sizeof(float);
  SMX_REQUIRE(n_small == 0 || (small_smem_floats > 0 && small_smem_floats <= SMALL_ELEMS + SMALL_RC),
              "adafactor: small-slice shared memory %d floats out of range", small_smem_floats);
  cudaStream_t st = (cudaStream_t)stream;
  // row_acc / col_acc / sumsq of every tensor live in one scratch block: one memset per step
  SMX_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)scratch_bytes, st));
  Hyper h{beta2t, eps1, lr, clip_threshold, weight_decay, reinterpret_cast<const long long*>(step_dev), decay_rate};
  if (step_dev) {
    launch_pdl(bump_step_kernel, dim3(1), dim3(1), 0, st, reinterpret_cast<long long*>(step_dev));
    SMX_CHECK_CUDA(cudaGetLastError());
  }
  if (n_small > 0) {
    static bool attr_set = false;
    if (!attr_set) {
      SMX_CHECK_CUDA(cudaFuncSetAttribute(small_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (SMALL_ELEMS + SMALL_RC + 8) * (int)sizeof(float)));
      attr_set = true;
    }
    launch_pdl(small_moments_kernel, dim3(n_small), dim3(256), small_smem, st, tensors, small_slices, h);
    SMX_CHECK_CUDA(cudaGetLastError());
  }
  if (n_tiles > 0) {
    launch_pdl(stats_kernel, dim3(n_tiles), dim3(256), 0, st, tensors, tiles);
    SMX_CHECK_CUDA(cudaGetLastError());
    if (n_slices > 0) {
      launch_pdl(finalize_kernel, dim3(n_slices), dim3(256), 0, st, tensors, slices, h);
      SMX_CHECK_CUDA(cudaGetLastError());
    }
    launch_pdl(update_kernel<false>, dim3(n_tiles), dim3(256), 0, st, tensors, tiles, h);
    SMX_CHECK_CUDA(cudaGetLastError());
    launch_pdl(update_kernel<true>, dim3(n_tiles), dim3(256), 0, st, tensors, tiles, h);
    SMX_CHECK_CUDA(cudaGetLastError());
  }
  if (n_small > 0) {
    launch_pdl(small_apply_kernel, dim3(n_small), dim3(256), 0, st, tensors, small_slices, h);
    SMX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
