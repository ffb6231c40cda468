// LM head + cross-entropy without materialising [rows, vocab] logits.
//   hf:models/bart/modeling_bart.py:940-947 (lm_head + final_logits_bias + CrossEntropyLoss),
//   ref:speechmix/hf_model.py:446 (argmax of the logits).
// Forward: the tcgen05 GEMM's epilogue reduces every 128 x 256 logit tile to per-row
// {max, sum-exp, best value, best index} while it is still in TMEM/registers; a small
// second kernel merges the tiles of a row (lowest index wins argmax ties, like torch).
// Backward: logits are recomputed tile by tile and turned into (softmax - onehot) * coef
// for one vocabulary chunk at a time (see smx_lmhead_dlogits).
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

#include <string.h>

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

namespace lmhead {

__global__ void __launch_bounds__(256) combine_kernel(const float4* __restrict__ partial,
                                                      const float* __restrict__ label_logit,
                                                      const long long* __restrict__ labels, float* __restrict__ lse_out,
                                                      long long* __restrict__ argmax_out, float* __restrict__ row_loss,
                                                      float* __restrict__ loss_sum, float* __restrict__ count,
                                                      long long rows, int n_tiles, long long ignore_index) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float mx = -INFINITY, se = 0.f, best = -INFINITY;
  int best_idx = 0x7fffffff;
  for (int t = lane; t < n_tiles; t += 32) {
    const float4 v = partial[row * n_tiles + t];
    const float m_new = fmaxf(mx, v.x);
    se = se * __expf(mx - m_new) + v.y * __expf(v.x - m_new);
    mx = m_new;
    const int idx = __float_as_int(v.w);
    if (v.z > best || (v.z == best && idx < best_idx)) best = v.z, best_idx = idx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const float ose = __shfl_xor_sync(0xffffffffu, se, o);
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
    const float m_new = fmaxf(mx, omx);
    const float a = (mx == -INFINITY) ? 0.f : se * __expf(mx - m_new);
    const float b = (omx == -INFINITY) ? 0.f : ose * __expf(omx - m_new);
    se = a + b;
    mx = m_new;
    if (ob > best || (ob == best && oi < best_idx)) best = ob, best_idx = oi;
  }
  if (lane == 0) {
    const float lse = mx + logf(se);
    lse_out[row] = lse;
    if (argmax_out) argmax_out[row] = best_idx;
    if (labels) {
      const long long lab = labels[row];
      float loss = 0.f;
      if (lab != ignore_index) {
        loss = lse - label_logit[row];
        atomicAdd(loss_sum, loss);
        atomicAdd(count, 1.0f);
      }
      if (row_loss) row_loss[row] = loss;
    }
  }
}

static void fill_gemm(SmxGemm* g, const void* h, const void* emb, int64_t rows, int64_t dim, int64_t n, float scale) {
  memset(g, 0, sizeof(*g));
  g->mode = SMX_GEMM_NT;
  g->out_dtype = SMX_OUT_BF16;
  g->a.ptr = h, g->a.inner = dim, g->a.rows = rows, g->a.batches = 1, g->a.row_stride = dim, g->a.batch_stride = rows * dim;
  g->b.ptr = emb, g->b.inner = dim, g->b.rows = n, g->b.batches = 1, g->b.row_stride = dim, g->b.batch_stride = n * dim;
  g->m = rows, g->n = n, g->k = dim, g->batches = 1;
  g->nseg = 1, g->seg_len = (int)dim;
  g->alpha = scale;
}

}  // namespace lmhead
}  // namespace smx

extern "C" {

size_t smx_lmhead_ws_bytes(int64_t rows, int64_t vocab) {
  const int64_t n_parts = 2 * ((vocab + 255) / 256);  // two column halves per 256-wide vocabulary tile
  return (size_t)(rows * n_parts * 16 + rows * 4 + 256);
}

int smx_lmhead_ce_fwd(const void* h, const void* emb, const float* bias, const int64_t* labels, float* lse,
                      int64_t* argmax, float* row_loss, float* loss_sum, float* count, void* workspace, int64_t rows,
                      int64_t dim, int64_t vocab, float logit_scale, int64_t ignore_index, void* stream) {
  using namespace smx;
  SMX_REQUIRE(h && emb && lse && workspace, "lmhead_ce_fwd: null pointer");
  SMX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "lmhead_ce_fwd: workspace must be 16-byte aligned");
  const int64_t n_tiles = 2 * ((vocab + 255) / 256);
  float4* partial = reinterpret_cast<float4*>(workspace);
  float* label_logit = reinterpret_cast<float*>(partial + rows * n_tiles);
  cudaStream_t st = (cudaStream_t)stream;
  // labels == NULL (generation): still need a label array for the epilogue -> use label -1 via a zeroed dummy
  SmxGemm g;
  lmhead::fill_gemm(&g, h, emb, rows, dim, vocab, logit_scale);
  g.bias = bias;
  g.c = workspace;  // unused by this epilogue
  g.c_row_stride = 0;
  gemm::LmExtra ex;
  memset(&ex, 0, sizeof(ex));
  ex.epi = 1;
  ex.partial = partial;
  ex.label_logit = label_logit;
  ex.labels = reinterpret_cast<const long long*>(labels);
  SMX_REQUIRE(labels != nullptr, "lmhead_ce_fwd: labels required (pass -100 rows for pure argmax)");
  if (gemm::run(&g, &ex, stream)) return -1;
  launch_pdl(lmhead::combine_kernel, dim3((unsigned)ceil_div(rows, 8)), dim3(256), 0, st, 
      partial, label_logit, reinterpret_cast<const long long*>(labels), lse, reinterpret_cast<long long*>(argmax),
      row_loss, loss_sum, count, rows, (int)n_tiles, ignore_index);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_lmhead_dlogits(const void* h, const void* emb, const float* bias, const int64_t* labels, const float* lse,
                       const float* coef, void* dlogits, int64_t ld, int64_t rows, int64_t dim, int64_t vocab,
                       int64_t v0, int64_t vn, float logit_scale, void* stream) {
  using namespace smx;
  SMX_REQUIRE(h && emb && labels && lse && coef && dlogits, "lmhead_dlogits: null pointer");
  SMX_REQUIRE(v0 >= 0 && vn > 0 && v0 + vn <= vocab && ld >= vn && ld % 8 == 0, "lmhead_dlogits: bad chunk");
  SmxGemm g;
  lmhead::fill_gemm(&g, h, reinterpret_cast<const __nv_bfloat16*>(emb) + v0 * dim, rows, dim, vn, logit_scale);
  g.bias = bias ? bias + v0 : nullptr;
  g.c = dlogits;
  g.c_row_stride = ld;
  g.c_batch_stride = rows * ld;
  gemm::LmExtra ex;
  memset(&ex, 0, sizeof(ex));
  ex.epi = 2;
  ex.labels = reinterpret_cast<const long long*>(labels);
  ex.lse = lse;
  ex.coef = coef;
  ex.label_off = v0;
  return gemm::run(&g, &ex, stream);
}
}
