/* libspeechmix_sm100.so -- C ABI of the B200-native SpeechMix hot path.
 *
 * The reference (voidful/SpeechMix) has no FFI of its own: its hot path is
 * Python glue (ref:speechmix/hf_model.py:185-447) over `transformers` modules.
 * The drop-in seam is therefore the Python class API (speechmix_b200.SpeechMixEED
 * etc.); THIS header is the boundary between that Python host layer and the
 * hand-written sm_100a kernels.  Every entry point names the reference /
 * transformers code whose arithmetic it replaces.
 *
 * Conventions
 *  - plain C: raw device pointers, sizes, a cudaStream_t passed as void*.
 *  - pointers are BORROWED; the library never allocates or frees device memory
 *    and keeps no global state besides lazily-set kernel attributes.
 *  - every call only ENQUEUES work on the given stream (no host sync) and is
 *    CUDA-graph capturable.
 *  - return 0 on success, negative on error; smx_last_error() gives the message
 *    (thread-local).
 *  - activations: bf16, row-major, channels-last ([batch][time][channel]);
 *    statistics / losses / master gradients: fp32; token ids: int64.
 */
#ifndef SPEECHMIX_SM100_H_
#define SPEECHMIX_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMX_ABI_VERSION 3

const char* smx_last_error(void);
int smx_abi_version(void);
/* 1 if the current device is sm_100 (B200), 0 otherwise, <0 on CUDA error. */
int smx_device_ok(void);

/* ------------------------------------------------------------------------
 * Tensor-core GEMM family (tcgen05.mma + TMEM accumulators, TMA-fed).
 *
 * A bf16 operand is a 3-D view  [batches][rows][inner]  with `inner`
 * contiguous; strides are in ELEMENTS and must be multiples of 8.
 * ------------------------------------------------------------------------ */
typedef struct SmxView3 {
  const void* ptr;
  int64_t inner, rows, batches;
  int64_t row_stride, batch_stride;
} SmxView3;

#define SMX_MAX_SEG 4

enum { SMX_GEMM_NT = 0, SMX_GEMM_NN = 1, SMX_GEMM_TN = 2 };
enum { SMX_ACT_NONE = 0, SMX_ACT_GELU = 1, SMX_ACT_RELU = 2, SMX_ACT_DGELU = 3, SMX_ACT_DRELU = 4,
       /* forward: c = gelu(v) and aux_out = gelu'(v) (instead of v), so that the backward epilogue is ... */
       SMX_ACT_GELU_G = 5,
       /* ... c = v * aux_in (aux_in = the stored derivative) */
       SMX_ACT_MULAUX = 6 };
enum { SMX_OUT_BF16 = 0, SMX_OUT_F32 = 1 };

/* One descriptor covers every contraction on the path:
 *
 *  NT  C[b,r,n] = sum_k A[b, r+a_row_off[s], a_col_off[s]+kin] * B[n, b_col_off[s]+kin]
 *      (k = s*seg_len + kin).  Linear layers (hf:models/wav2vec2/modeling_wav2vec2.py:466-573,
 *      hf:models/bart/modeling_bart.py:143-391) use one segment; the strided
 *      Conv1d layers of the feature encoder (hf:...wav2vec2.py:254-323) and the
 *      k=2,s=2 length adapters (ref:speechmix/hf_model.py:253-266) use one
 *      segment per tap over a frame-pair view of the channels-last input, so the
 *      convolution runs as an implicit GEMM with no im2col buffer.
 *  NN  C[b,r,n] = sum_k A[b, r+a_row_off[s], a_col_off[s]+kin] * B[b_row_off[s]+kin, b_col_off[s]+n]
 *      (data gradients: B is the forward weight read MN-major, no transpose copy).
 *  TN  C[m,n]  (+)= sum_{b,r} A[b, r+a_row_off[0], a_col_off[0]+m] * B[b, r+b_row_off[s], b_col_off[s]+nin]
 *      (n = s*seg_len + nin; weight gradients, contraction over rows, fp32 output,
 *       optional split over the contraction with atomic accumulation).
 *
 * Epilogue (NT/NN):  v = alpha*acc (+bias[n]);  aux_out <- v (optional);
 *   v = act(v) | v*gelu'(aux_in) | v*[aux_in>0];  v += residual;  C <- v.
 */
typedef struct SmxGemm {
  int32_t mode;
  int32_t out_dtype;    /* SMX_OUT_* (TN: always fp32) */
  SmxView3 a, b;
  int64_t m;            /* NT/NN: output rows per batch.  TN: output rows (= extent of A.inner used) */
  int64_t n;            /* output columns */
  int64_t k;            /* NT/NN: contraction length (= nseg*seg_len).  TN: contraction rows per batch */
  int64_t batches;
  int32_t nseg;
  int32_t seg_len;
  int32_t a_row_off[SMX_MAX_SEG], a_col_off[SMX_MAX_SEG];
  int32_t b_row_off[SMX_MAX_SEG], b_col_off[SMX_MAX_SEG];
  void* c;
  int64_t c_row_stride, c_batch_stride; /* elements */
  int32_t act;
  int32_t split_k;      /* TN only; >1 => atomic fp32 accumulation into a zeroed or live C */
  int32_t accumulate;   /* TN: add into C even when split_k == 1.  NT/NN with fp32 output: C += v */
  float alpha;
  const float* bias;    /* [n] fp32 or NULL */
  const void* residual; /* bf16, same shape/strides as C, or NULL */
  int64_t res_row_stride, res_batch_stride;
  void* aux_out;        /* bf16 pre-activation copy, C's strides, or NULL */
  const void* aux_in;   /* bf16 pre-activation for DGELU/DRELU, C's strides */
} SmxGemm;

int smx_gemm(const SmxGemm* g, void* stream);

/* ------------------------------------------------------------------------
 * Row-wise kernels (HBM-bound)
 * ------------------------------------------------------------------------ */
/* y = act(LayerNorm(x (+ res)) * gamma + beta) over the last dim (cols), eps as
 * given (act: SMX_ACT_NONE or SMX_ACT_GELU -- the "layer" feature-encoder convs,
 * hf:...wav2vec2.py:275-299); optionally also writes sum = x + res (bf16) and per-row mean / rstd.
 * hf:...wav2vec2.py:422-434 (feature projection), :576-655 (encoder layers),
 * hf:...bart.py:261-391.  rms_only=1 gives T5 RMSNorm (hf:models/t5/modeling_t5.py:46-69). */
int smx_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y,
                      void* sum_out, float* mean, float* rstd, int64_t rows, int64_t cols, float eps,
                      int rms_only, int act, void* stream);
/* dx (+= dres_in) for the op above; dgamma/dbeta and the optional dx_colsum[c] = sum_rows dx[r, c] (the bias
 * gradient of the linear layer feeding a post-LN block) are ACCUMULATED (fp32 atomics) into zero-initialised
 * buffers. */
int smx_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, const void* dres_in, void* dx, float* dgamma, float* dbeta, float* dx_colsum,
                      int64_t rows, int64_t cols, int rms_only, int act, void* stream);

/* out[n] (+)= sum_r x[r, n]   (bias gradients). fp32 accumulate via atomics into a zeroed buffer. */
int smx_colsum(const void* x, float* out, int64_t rows, int64_t cols, int64_t row_stride, void* stream);

/* elementwise helpers */
int smx_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* Weight norm of the positional conv (torch weight_norm(dim=2), hf:...wav2vec2.py:341-355): v [rows][k] fp32
 * (rows = out * in/groups), g [k]; w[r][j] = g[j] v[r][j] / ||v[:, j]||.  sq_ws [k] receives ||v[:, j]||^2 (kept
 * for the backward); bwd: dv, dg from dw (dot_ws [k] workspace). */
int smx_weightnorm_fwd(const float* v, const float* g, float* sq_ws, float* w, int64_t rows, int64_t k, void* stream);
int smx_weightnorm_bwd(const float* v, const float* g, const float* sq, const float* dw, float* dot_ws, float* dv,
                       float* dg, int64_t rows, int64_t k, void* stream);
/* Multi-tensor refresh of the bf16 (or fp32) working copies of all parameters in ONE launch: entry i copies
 * src[i][0..n[i]) (fp32) to dst[i] as bf16 (dst_f32[i]=0) or fp32 (=1).  `table` is a DEVICE array of
 * SmxCastEntry sorted by first_chunk (chunk = SMX_CAST_CHUNK elements; first_chunk = running chunk count);
 * total_chunks = sum of ceil(n/chunk).  Replaces the per-parameter `.to(bf16)` an autocast reference run performs. */
#define SMX_CAST_CHUNK 4096
typedef struct {
  const float* src;
  void* dst;
  int64_t n;
  int32_t dst_f32;
  int32_t first_chunk;
} SmxCastEntry;
int smx_multi_cast(const SmxCastEntry* table, int32_t n_entries, int32_t total_chunks, void* stream);
/* ------------------------------------------------------------------------
 * Multi-tensor Adafactor step -- the optimizer of the reference recipe (ref:train.py:298 optim="adafactor" =
 * transformers.optimization.Adafactor.step with relative_step=False, scale_parameter=False, beta1=None, as the HF
 * Trainer configures it).  One call updates every listed fp32 parameter:
 *   tensors with len(shape) >= 2 are factored over their last two dims [rows][cols], leading dims = `batch`
 *   independent slices: row = exp_avg_sq_row [batch][rows], col = exp_avg_sq_col [batch][cols];
 *   vectors (factored = 0): row = exp_avg_sq [numel], col unused, batch = rows = 1, cols = numel.
 * row_acc / col_acc / sumsq are per-tensor scratch carved out of ONE block `scratch` (zeroed by the call);
 * rmean is [batch] scratch.  tiles: 64 x 256 element tiles covering every slice once (vectors are viewed as
 * [ceil(numel/256)][256]); slices: one (tensor, b) pair per factored slice.  All three tables are DEVICE arrays.
 * beta2t = 1 - step^decay_rate is computed by the caller; eps1 = eps[0].  Capturable mode: step_dev != NULL is a
 * DEVICE int64 step counter -- the call increments it and the kernels evaluate beta2t from it (decay_rate), so the
 * launch sequence can be replayed from a CUDA graph; beta2t is then ignored.
 * ------------------------------------------------------------------------ */
typedef struct {
  float* p;
  const float* g;
  float *row, *col, *row_acc, *col_acc, *rmean, *sumsq;
  int64_t batch, rows, cols, numel;
  int32_t factored, pad_;
} SmxAdafactorTensor;
typedef struct {
  int32_t tensor, b, r0, c0;
} SmxAdafactorTile;
typedef struct {
  int32_t tensor, b;
} SmxAdafactorSlice;
/* small_slices: factored slices of <= 16384 elements (rows + cols <= 4096) that are NOT in `tiles` / `slices` and take
 * the block-per-slice kernels; small_smem_floats = max over them of rows*cols + rows + cols. */
int smx_adafactor_step(const SmxAdafactorTensor* tensors, int32_t n_tensors, const SmxAdafactorTile* tiles,
                       int32_t n_tiles, const SmxAdafactorSlice* slices, int32_t n_slices,
                       const SmxAdafactorSlice* small_slices, int32_t n_small, int32_t small_smem_floats,
                       void* scratch, int64_t scratch_bytes, float beta2t, float eps1, float lr, float clip_threshold,
                       float weight_decay, int64_t* step_dev, float decay_rate, void* stream);

int smx_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);
/* out = a * b: gate product of the gated feed-forward of T5 v1.1 / mT5 (hf:models/t5/modeling_t5.py T5DenseGatedActDense) */
int smx_mul_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);
int smx_act_bf16(const void* x, void* y, int64_t n, int act, void* stream);
/* out = dy * act'(pre), act in {SMX_ACT_GELU, SMX_ACT_RELU} */
int smx_dact_bf16(const void* dy, const void* pre, void* out, int64_t n, int act, void* stream);
/* dst[o, t*cin + c] = src[o, c, t] : Conv1d weight [out][in][k] -> packed [out][k*in] bf16 */
int smx_pack_conv_weight(const float* src, void* dst, int64_t cout, int64_t cin, int64_t k, void* stream);
/* inverse of the above for fp32 gradients: dst[o, c, t] = src[o, t*cin + c] */
int smx_unpack_conv_wgrad(const float* src, float* dst, int64_t cout, int64_t cin, int64_t k, void* stream);

/* ------------------------------------------------------------------------
 * conv0 of the feature encoder: Conv1d(1->C,k,s, no bias) + GroupNorm(C groups)
 * + GELU, fused; the un-normalised conv output never exists in HBM.
 * hf:...wav2vec2.py:302-323 (Wav2Vec2GroupNormConvLayer).
 * ------------------------------------------------------------------------ */
/* moments: workspace [batch][110] fp32 (window sums + second-moment matrix of the raw audio);
 * stats[b][c] = {mean, rstd} of the conv output over time, derived from the moments
 * (exact algebra, one pass over the waveform). */
#define SMX_CONV0_MOMENT_BLOCKS 32
/* partial_ws: batch * SMX_CONV0_MOMENT_BLOCKS * 65 floats (per-block partial moments, reduced in a fixed order so
 * that the forward pass is bit-reproducible) */
int smx_conv0_stats(const float* audio, const float* w, float* moments, float* stats, float* partial_ws,
                    int64_t batch, int64_t n_samples, int64_t t_out, int channels, int ksize, int stride, float eps,
                    void* stream);
/* y = gelu(z) (bf16 [batch][t_out][channels]); gprime (optional, training) = gelu'(z) in the same layout, so
 * that the backward pass is a pure stream over dy and gprime (no convolution / activation recompute). */
int smx_conv0_gn_gelu_fwd(const float* audio, const float* w, const float* gamma, const float* beta,
                          const float* stats, void* y, void* gprime, int64_t batch, int64_t n_samples,
                          int64_t t_out, int channels, int ksize, int stride, void* stream);
/* one pass over dy and gprime; partial: workspace [batch][channels][ksize+2] fp32; writes dw [C][k], dgamma, dbeta */
int smx_conv0_gn_gelu_bwd(const float* audio, const float* w, const float* gamma, const float* beta,
                          const float* stats, const float* moments, const void* dy, const void* gprime,
                          float* partial, float* dw, float* dgamma, float* dbeta, int64_t batch, int64_t n_samples,
                          int64_t t_out, int channels, int ksize, int stride, void* stream);

/* feat_extract_norm="layer" variant of layer 0 (HuBERT-large / wav2vec2-large-lv60,
 * hf:...wav2vec2.py:275-299): Conv1d(1->C,k,s,bias) -> LayerNorm over channels -> GELU, one warp per frame. */
int smx_conv0_ln_gelu_fwd(const float* audio, const float* w, const float* conv_bias, const float* gamma,
                          const float* beta, void* y, int64_t batch, int64_t n_samples, int64_t t_out,
                          int channels, int ksize, int stride, float eps, void* stream);
/* dconv = gradient w.r.t. the (biased) conv output, bf16 [B,T,C]; dgamma/dbeta accumulated into zeroed buffers */
int smx_conv0_ln_gelu_bwd(const float* audio, const float* w, const float* conv_bias, const float* gamma,
                          const float* beta, const void* dy, void* dconv, float* dgamma, float* dbeta,
                          int64_t batch, int64_t n_samples, int64_t t_out, int channels, int ksize, int stride,
                          float eps, void* stream);
/* dw[C][k], dbias[C] (zeroed, accumulated) = correlation of dconv with the waveform windows */
int smx_conv0_wgrad(const float* audio, const void* dconv, float* dw, float* dbias, int64_t batch,
                    int64_t n_samples, int64_t t_out, int channels, int ksize, int stride, void* stream);

/* ------------------------------------------------------------------------
 * Positional conv embedding: grouped Conv1d(H->H, k=128, pad=64, groups=16),
 * drop last frame, GELU.  hf:...wav2vec2.py:326-379.  Tensor-core implicit
 * GEMM over shifted windows of one smem-resident input slab.
 * w_packed: [groups][k][cg_out][cg_in] bf16.  y may carry the fused residual:
 * y = x + gelu(conv(x) + bias)   when add_input != 0.
 * ------------------------------------------------------------------------ */
int smx_posconv_fwd(const void* x, const void* w_packed, const float* bias, void* y, void* pre_out,
                    int64_t batch, int64_t t, int hidden, int groups, int ksize, int add_input, void* stream);
/* dx = conv^T(dpre) (+ residual) with w_packed_t = taps flipped, in/out swapped */
int smx_posconv_dgrad(const void* dpre, const void* w_packed_t, const void* residual, void* dx, int64_t batch,
                      int64_t t, int hidden, int groups, int ksize, void* stream);
/* dw[g][k][o][c] (fp32, zero-initialised) += sum_{b,t} dpre[b,t,g*cg+o] * x[b,t+k-pad,g*cg+c] */
int smx_posconv_wgrad(const void* dpre, const void* x, float* dw, int64_t batch, int64_t t, int hidden,
                      int groups, int ksize, void* stream);

/* ------------------------------------------------------------------------
 * Attention (head_dim 64), flash-style: scores never reach HBM.
 * q/k/v/o are bf16 views [batch][time][heads*64] with arbitrary row strides
 * (so a fused QKV buffer can be addressed in place).
 * hf:...wav2vec2.py:438-549, hf:...bart.py:143-258, hf:models/t5/modeling_t5.py:153-345.
 * causal: 0/1.  scale: multiplies q.k (T5: 1.0).  bias: optional additive
 * [heads][tq][tk] fp32 (T5 relative position bias) or NULL.
 * lse: [batch][heads][tq] fp32 (natural log), written by fwd, read by bwd.
 * kv_len: optional [batch] int32, 1 <= kv_len[b] <= tk: keys/values at or past kv_len[b] are masked out for
 *   every query of sample b (key-padding mask of a padded batch, hf:...wav2vec2.py:1026-1044 turned into
 *   per-sample frame counts).  Not combined with causal.  Contract for smx_attn_bwd: the k / v rows at or past
 *   kv_len[b] hold ZEROS (smx_mask_rows does that on the projection output), and dk / dv come back zero there.
 * ------------------------------------------------------------------------ */
typedef struct SmxAttn {
  const void *q, *k, *v;
  void* o;
  float* lse;
  int64_t q_row_stride, k_row_stride, v_row_stride, o_row_stride;       /* elements */
  int64_t q_batch_stride, k_batch_stride, v_batch_stride, o_batch_stride;
  int32_t batch, heads, tq, tk, causal;
  float scale;
  const float* bias;
  /* backward only */
  const void* d_o;
  void *dq, *dk, *dv;
  float* delta; /* [batch][heads][tq] workspace: rowsum(dO * O) */
  float* dbias; /* optional [heads][tq][tk] fp32, accumulated */
  int64_t do_row_stride, do_batch_stride;
  int64_t dq_row_stride, dk_row_stride, dv_row_stride;
  int64_t dq_batch_stride, dk_batch_stride, dv_batch_stride;
  const int32_t* kv_len; /* optional per-sample key count (see above); NULL = all tk keys */
  /* dropout on the attention probabilities (hf eager_attention_forward: F.dropout(attn_weights, p)); the mask is a
   * function of (state, call, b, h, q, k) and is regenerated by the backward kernels.  NULL / p = 0: off. */
  const uint64_t* dropout_state; /* device: {seed, step} (see smx_dropout) */
  uint32_t dropout_call;
  float dropout_p;
} SmxAttn;
int smx_attn_fwd(const SmxAttn* a, void* stream);
int smx_attn_bwd(const SmxAttn* a, void* stream);
/* Development hook (tools/probe_attn.py trace): a device buffer of 3 x 64 uint64 receives SM-clock stamps of CTA (0,0,0)
 * of the following plain-attention forward launches -- [MMA thread | softmax group 0 | softmax group 1][event];
 * NULL switches tracing off.  Not used by the product path. */
int smx_debug_attn_trace(void* device_buffer);

/* Zero the columns [col_begin, col_begin + col_count) of every row t >= len[b] of a bf16 [batch][t][...] buffer
 * (row / batch strides in elements; col_begin, col_count multiples of 8).  Used for "padded frames output 0"
 * (hf:...wav2vec2.py:672-675) and for the k | v part of a fused QKV projection under a key-padding mask. */
int smx_mask_rows(void* x, const int32_t* len, int64_t batch, int64_t t, int64_t row_stride, int64_t batch_stride,
                  int64_t col_begin, int64_t col_count, void* stream);

/* Counter-based dropout at the elementwise positions of the HF modules (hidden / activation / embedding dropout):
 *   out = residual + (keep ? x / (1 - p) : 0)     bf16, n elements (multiple of 8), residual optional
 *   aux_mode 1: aux_out = keep ? aux_in / (1 - p) : 0;   2: aux_out = (keep && aux_in > 0) ? 1 / (1 - p) : 0
 * keep is a pure function of (state[0] = seed, state[1] = step, call, element index): the backward pass calls the same
 * function on the gradient, and `state` lives in device memory so that a replayed CUDA graph draws new masks.
 * smx_dropout_mask writes the keep decisions as bytes (mode 0: the elementwise numbering; mode 1: attention
 * probabilities [rows][tk] as numbered by the attention kernels) -- used by the parity tests. */
int smx_dropout(const void* x, const void* residual, void* out, const void* aux_in, void* aux_out, int aux_mode, int64_t n,
                const uint64_t* state, uint32_t call, float p, void* stream);
int smx_dropout_mask(uint8_t* mask, int64_t n, int64_t tk, int mode, const uint64_t* state, uint32_t call, float p,
                     void* stream);

/* Tail of a post-LN block with output dropout (hf:...wav2vec2.py:590-602, hf:...bart.py:280-309: dropout -> residual add ->
 * LayerNorm), as ONE launch each way instead of dropout + LayerNorm (forward) and LayerNorm backward + dropout + column
 * sums (backward).  Same mask as smx_dropout on the same [rows][cols] tensor with the same (state, call, p).
 *   fwd: sum_out = (keep ? x / (1 - p) : 0) + res;  y = LN(sum_out) * gamma + beta;  mean / rstd of the stored sum
 *   bwd: dx = LN'(dy) (the gradient of the residual branch);  dx_drop = keep ? dx / (1 - p) : 0 (the gradient of x);
 *        dgamma / dbeta (dbeta may be NULL) and dxd_colsum (may be NULL) += column sums -- all three zeroed by the caller.
 * x here is the stored sum (the LayerNorm input) with its saved mean / rstd. */
int smx_layernorm_dropout_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, void* sum_out,
                              float* mean, float* rstd, int64_t rows, int64_t cols, float eps, const uint64_t* state,
                              uint32_t call, float p, void* stream);
int smx_layernorm_dropout_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                              void* dx_drop, float* dgamma, float* dbeta, float* dxd_colsum, int64_t rows, int64_t cols,
                              const uint64_t* state, uint32_t call, float p, void* stream);

/* SpecAugment on the projected features (hf:models/wav2vec2/modeling_wav2vec2.py:1280-1324; the mask INDICES come
 * from the host, drawn exactly like the reference's _compute_mask_indices):
 *   y[b,t,:] = time_mask[b,t] ? embed : x[b,t,:];  y[b,t,c] = 0 where feat_mask[b,c]   (either mask may be NULL)
 * time_mask: uint8 [batch][t]; feat_mask: uint8 [batch][hidden]; embed: fp32 [hidden] (masked_spec_embed).
 * bwd: dx = dy outside the masks and 0 inside; dembed (fp32 [hidden], zero-initialised by the caller, may be NULL)
 * += the column sums of dy over the time-masked rows. */
int smx_spec_augment_fwd(const void* x, void* y, const uint8_t* time_mask, const uint8_t* feat_mask, const float* embed,
                         int64_t batch, int64_t t, int64_t hidden, void* stream);
int smx_spec_augment_bwd(const void* dy, void* dx, float* dembed, const uint8_t* time_mask, const uint8_t* feat_mask,
                         int64_t batch, int64_t t, int64_t hidden, void* stream);

/* ------------------------------------------------------------------------
 * Input embeddings:  out[b,t,:] = tok_emb[ids[b,t]]*scale (if ids) + x_in[b,t,:] (if x_in)
 *                                 + pos_emb[t + t_start + pos_offset] (if pos_emb)      (bf16 out)
 * hf:...bart.py:74-111 (learned positions, offset 2), :508-530 (encoder on inputs_embeds),
 * :620-640 (decoder).  bwd scatters fp32 into the (tied) table gradients.
 * ------------------------------------------------------------------------ */
int smx_embed_fwd(const int64_t* ids, const float* tok_emb, const float* pos_emb, const void* x_in, void* out,
                  int64_t batch, int64_t t, int64_t dim, float scale, int64_t pos_offset, int64_t t_start,
                  void* stream);
int smx_embed_bwd(const int64_t* ids, const void* dout, float* d_tok_emb, float* d_pos_emb, int64_t batch,
                  int64_t t, int64_t dim, float scale, int64_t pos_offset, void* stream);

/* ------------------------------------------------------------------------
 * LM head + cross-entropy without materialising [rows, vocab] logits.
 * hf:...bart.py:940-947 ; ref:speechmix/hf_model.py:446 (argmax).
 * fwd: per row online logsumexp, label logit, argmax (lowest index wins ties)
 *      over vocab tiles of  h . E^T * logit_scale + bias.
 * partial: workspace of smx_lmhead_ws_bytes(rows, vocab) bytes.
 * loss_sum[0] += sum of per-row NLL over labels != ignore_index; count[0] += #rows counted.
 * ------------------------------------------------------------------------ */
size_t smx_lmhead_ws_bytes(int64_t rows, int64_t vocab);
int smx_lmhead_ce_fwd(const void* h, const void* emb, const float* bias, const int64_t* labels, float* lse,
                      int64_t* argmax, float* row_loss, float* loss_sum, float* count, void* workspace,
                      int64_t rows, int64_t dim, int64_t vocab, float logit_scale, int64_t ignore_index,
                      void* stream);
/* dlogits[rows][0:vn) (row stride ld, bf16) = (softmax(logits)[v0:v0+vn) - onehot(label)) * coef[row]
 * for the chunked backward: the chunk stays L2-resident between the two GEMMs that consume it
 * (dh += dlogits . E[v0:v0+vn), dE[v0:v0+vn) = dlogits^T . h).  coef[row] = dloss/count or 0. */
int smx_lmhead_dlogits(const void* h, const void* emb, const float* bias, const int64_t* labels,
                       const float* lse, const float* coef, void* dlogits, int64_t ld, int64_t rows,
                       int64_t dim, int64_t vocab, int64_t v0, int64_t vn, float logit_scale, void* stream);

/* ------------------------------------------------------------------------
 * Weighted layer sum (ref:speechmix/hf_model.py:411-423): out = sum_l w[l]*x_l
 * ------------------------------------------------------------------------ */
int smx_weighted_sum_fwd(const void* const* xs, const float* w, void* out, int n_layers, int64_t n,
                         void* stream);
/* dw[l] += sum(dout * x_l) */
int smx_weighted_sum_bwd_w(const void* const* xs, const void* dout, float* dw, int n_layers, int64_t n,
                           void* stream);

/* ------------------------------------------------------------------------
 * SpeechMixSelf auxiliary losses (ref:speechmix/hf_model.py:551-581, HFSpeechMixSelf.cal_loss).
 * KL(softmax(teacher) || softmax(student)) with reduction "batchmean" is evaluated per vocabulary chunk on
 * fp32 logit chunks s / t [rows][vn] (row stride ld) written by smx_gemm:
 *   cross[row] += sum_v exp(t_v - lse_t) (t_v - s_v);   kld = sum_row (cross - lse_t + lse_s) * inv_batch
 *   dlogits    = coef_ce[row] (softmax(s) - onehot(label)) + coef_kl[0] (softmax(s) - softmax(t))    (bf16)
 * ------------------------------------------------------------------------ */
int smx_kl_chunk_fwd(const float* s, const float* t, int64_t ld, int64_t rows, int64_t vn, const float* lse_t,
                     float* cross, void* stream);
int smx_kl_finalize(const float* cross, const float* lse_s, const float* lse_t, int64_t rows, float inv_batch,
                    float* out, void* stream);
int smx_kl_chunk_bwd(const float* s, const float* t, int64_t ld, int64_t rows, int64_t vn, int64_t v0,
                     const int64_t* labels, const float* lse_s, const float* lse_t, const float* coef_ce,
                     const float* coef_kl, void* dlogits, int64_t ld_out, void* stream);
/* attention-projection MSE (ref :561-570): A = softmax(T . view(S,[dim,ts]) / sqrt(dim)) -- view() is the
 * reference's memory reinterpretation of the [ts, dim] matrix --, P = A . S, loss[0] += mean((P - T)^2).
 * text_h [batch,tt,dim], speech_h [batch,ts,dim] bf16; attn [batch,tt,ts], diff [batch,tt,dim] fp32 are
 * kept for the backward, which writes d_speech_h = gscale[0] * dLoss/dS (gscale = upstream grad * 2/N). */
int smx_self_mse_fwd(const void* text_h, const void* speech_h, float* attn, float* diff, float* loss, int64_t batch,
                     int64_t tt, int64_t ts, int64_t dim, void* stream);
int smx_self_mse_bwd(const void* text_h, const void* speech_h, const float* attn, const float* diff, float* dscores,
                     const float* gscale, void* d_speech_h, int64_t batch, int64_t tt, int64_t ts, int64_t dim,
                     void* stream);

/* ------------------------------------------------------------------------
 * SpeechMixGAN discriminator features (ref:speechmix/hf_model.py:637-686): the reference feeds Linear(dim*dim, 1) with
 * flatten(X.view(dim, t) . X.view(t, dim)) -- a memory reinterpretation of the [t, dim] states.  With W = weight.view(dim,
 * dim) and Z = X W^T (a GEMM), logit[b] = sum_{t,i} Xflat[b][i t_len + t] * Z[b][t][i]: fwd accumulates that into out[b]
 * (zeroed by the call); bwd writes dz[b][t][i] = g[b] Xflat[b][i t_len + t] and dx[b][f] = g[b] Z[b][f % t_len][f / t_len].
 * x, z, dx, dz: bf16 [batch][t][dim] contiguous; out, g: fp32 [batch].
 * ------------------------------------------------------------------------ */
int smx_gram_dot_fwd(const void* x, const void* z, float* out, int64_t batch, int64_t t, int64_t dim, void* stream);
int smx_gram_dot_bwd(const void* x, const void* z, const float* g, void* dx, void* dz, int64_t batch, int64_t t, int64_t dim,
                     void* stream);

/* ------------------------------------------------------------------------
 * T5 relative position bias (hf:models/t5/modeling_t5.py:188-247): bias[h][i][j] = weight[bucket][h] with
 * bucket = table[(j - (i + q_offset)) + (tq + q_offset - 1)]; the bucket table (tq + q_offset + tk - 1 int32)
 * is built on the host with the reference's own formula.  bwd accumulates into a zeroed dweight.
 * ------------------------------------------------------------------------ */
int smx_relpos_bias_fwd(const float* weight, const int32_t* table, float* bias, int64_t heads, int64_t tq, int64_t tk,
                        int64_t q_offset, void* stream);
int smx_relpos_bias_bwd(const float* dbias, const int32_t* table, float* dweight, int64_t heads, int64_t tq, int64_t tk,
                        int64_t q_offset, int64_t n_buckets, void* stream);

/* ------------------------------------------------------------------------
 * fp32 verification path (inference only; csrc/fp32.cu): the forward graph of ref:speechmix/hf_model.py:378-447
 * with fp32 activations and fp32 CUDA-core arithmetic, used to compare greedy-decoded ids bit for bit with the
 * reference's fp32 run.  c[m,n] = act(alpha * sum_k a[m,k] w[n,k] + bias[n]) + residual[m,n]; `lda` may be
 * smaller than k (overlapping rows): a k-tap stride-s convolution over channels-last frames is this GEMM with
 * lda = s*C, k = taps*C and w = weight[out][tap][in].  act: SMX_ACT_NONE / GELU (exact erf) / RELU.
 * ------------------------------------------------------------------------ */
int smx_f32_gemm_nt(const float* a, int64_t lda, int64_t a_batch_stride, const float* w, const float* bias,
                    const float* residual, int64_t ldr, int64_t r_batch_stride, float* c, int64_t ldc,
                    int64_t c_batch_stride, int64_t m, int64_t n, int64_t k, int64_t batches, int act, float alpha,
                    void* stream);
int smx_f32_layernorm(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int64_t cols,
                      float eps, int rms_only, int act, void* stream);
/* in place: x[b,t,c] = gelu(GroupNorm_{groups == channels}(x)); stats_ws: batch*channels*2 doubles */
int smx_f32_groupnorm_gelu(float* x, double* stats_ws, const float* gamma, const float* beta, int64_t batch, int64_t t,
                           int64_t channels, float eps, void* stream);
/* y = x*add_input + gelu(grouped_conv(x) + bias), weight [hidden][hidden/groups][ksize], padding ksize/2 */
int smx_f32_posconv(const float* x, const float* w, const float* bias, float* y, int64_t batch, int64_t t,
                    int64_t hidden, int64_t groups, int64_t ksize, int add_input, void* stream);
int smx_f32_attn(const float* q, const float* k, const float* v, float* o, int64_t q_rs, int64_t q_bs, int64_t k_rs,
                 int64_t k_bs, int64_t v_rs, int64_t v_bs, int64_t o_rs, int64_t o_bs, int64_t batch, int64_t heads,
                 int64_t tq, int64_t tk, int causal, float scale, const float* bias, void* stream);
int smx_f32_embed(const int64_t* ids, const float* tok, const float* pos, const float* x_in, float* out, int64_t batch,
                  int64_t t, int64_t dim, float scale, int64_t pos_offset, void* stream);
/* running argmax over ascending vocabulary chunks (lowest index wins ties, like torch.argmax) */
int smx_f32_argmax_chunk(const float* logits, int64_t ld, int64_t rows, int64_t vn, int64_t v0, float* best, int64_t* idx,
                         void* stream);
/* y = (first ? 0 : y) + w[wi] * x */
int smx_f32_axpy(const float* x, const float* w, int32_t wi, float* y, int64_t n, int first, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPEECHMIX_SM100_H_ */
