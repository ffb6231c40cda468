// conv0 of the wav2vec2 / HuBERT feature encoder, fused:
//   Conv1d(1 -> C, k = 10, stride 5, no bias)  ->  GroupNorm(C groups == per (batch, channel) over time)
//   -> GELU                                     hf:models/wav2vec2/modeling_wav2vec2.py:302-323
//
// The convolution is linear in a 10-sample window, so the GroupNorm statistics
// follow exactly from the window moments of the raw waveform:
//   mean_c = w_c . E[win],   E[y_c^2] = w_c^T E[win win^T] w_c.
// Forward therefore makes ONE pass that reads the waveform (tiny) and writes the
// normalised, activated [B, T, C] bf16 output (the largest activation of the whole
// model) exactly once; the un-normalised conv output never exists in HBM.
// Backward makes one pass over dY, recomputing the pre-activation from the waveform,
// and reduces everything the weight / affine gradients need to 12 numbers per (b, c).
#include "../../include/speechmix_sm100.h"
#include "host_common.h"
#include "sm100_prims.cuh"

namespace smx {
namespace conv0 {

constexpr int K = 10, S = 5;
constexpr int NMOM = K + K * K;  // window sums + full second-moment matrix
constexpr int NPART = K + K * (K + 1) / 2;  // what one block accumulates: window sums + upper triangle

// ------------------------------------------------------------------ window moments
__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ audio, float* __restrict__ partial,
                                                      long long n_samples, long long t_out) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const float* x = audio + (long long)b * n_samples;
  float m[K], r[K * (K + 1) / 2];
#pragma unroll
  for (int i = 0; i < K; ++i) m[i] = 0.f;
#pragma unroll
  for (int i = 0; i < K * (K + 1) / 2; ++i) r[i] = 0.f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < t_out; t += (long long)gridDim.x * blockDim.x) {
    float w[K];
#pragma unroll
    for (int i = 0; i < K; ++i) w[i] = x[t * S + i];
    int idx = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      m[i] += w[i];
#pragma unroll
      for (int j = i; j < K; ++j) r[idx++] += w[i] * w[j];
    }
  }
  __shared__ float red[8][K + K * (K + 1) / 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const float v = warp_sum(m[i]);
    if (lane == 0) red[warp][i] = v;
  }
#pragma unroll
  for (int i = 0; i < K * (K + 1) / 2; ++i) {
    const float v = warp_sum(r[i]);
    if (lane == 0) red[warp][K + i] = v;
  }
  __syncthreads();
  if (threadIdx.x < K + K * (K + 1) / 2) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    // per-block partial (no atomics: the forward pass stays bit-reproducible from run to run)
    partial[((long long)b * gridDim.x + blockIdx.x) * NPART + threadIdx.x] = v;
  }
}

// moments[b] = sum over the blocks' partials in a fixed order; expands the upper triangle to the full matrix
__global__ void moments_reduce_kernel(const float* __restrict__ partial, float* __restrict__ moments, int blocks) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  if (threadIdx.x >= NPART) return;
  float v = 0.f;
  for (int k = 0; k < blocks; ++k) v += partial[((long long)b * blocks + k) * NPART + threadIdx.x];
  float* mo = moments + (long long)b * NMOM;
  if (threadIdx.x < K) {
    mo[threadIdx.x] = v;
  } else {
    int rem = threadIdx.x - K, i = 0;
    while (rem >= K - i) {
      rem -= K - i;
      ++i;
    }
    const int j = i + rem;
    mo[K + i * K + j] = v;
    mo[K + j * K + i] = v;
  }
}

__global__ void stats_kernel(const float* __restrict__ w, const float* __restrict__ moments, float* __restrict__ stats,
                             int batch, int channels, long long t_out, float eps) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * channels) return;
  const int b = i / channels, c = i % channels;
  const float* mo = moments + (long long)b * NMOM;
  const float* wc = w + c * K;
  const double inv_t = 1.0 / (double)t_out;
  double mean = 0.0, ey2 = 0.0;
  for (int j = 0; j < K; ++j) {
    mean += (double)wc[j] * mo[j];
    double row = 0.0;
    for (int l = 0; l < K; ++l) row += (double)wc[l] * mo[K + j * K + l];
    ey2 += (double)wc[j] * row;
  }
  mean *= inv_t;
  ey2 *= inv_t;
  double var = ey2 - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// ------------------------------------------------------------------ forward
// block: 128 channel quads x 2 frame lanes; FR frames per block staged through smem
constexpr int FWD_FR = 256;
__global__ void __launch_bounds__(256, 2) fwd_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  const float* __restrict__ stats, bf16* __restrict__ y,
                                                  bf16* __restrict__ gprime, long long n_samples, long long t_out,
                                                  int channels) {
  pdl_trigger();
  pdl_wait();
  // The kernel is issue-bound (ncu: 33 instructions per element, 74 % issue-active, DRAM at 39 %), so the window
  // arithmetic runs on packed fp32 pairs: z = b + sum over tap PAIRS of (w[2k], w[2k+1]) * (x[2k], x[2k+1]) = 5 FFMA2 +
  // 1 add instead of 10 FFMA, the window pairs are 8-byte shared-memory loads (S = 5 is odd: a second copy of the
  // slab shifted by one float keeps odd frames aligned), and GELU / GELU' are evaluated for two channels at once.
  static_assert(K % 2 == 0 && (S & 1) == 1, "pair loads assume an even window and an odd stride");
  constexpr int XS = FWD_FR * S + K + 2;
  __shared__ __align__(16) float xs[2][XS];       // xs[1][i] = xs[0][i + 1]
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * FWD_FR;
  const long long nfr = (t_out - t0) < FWD_FR ? (t_out - t0) : FWD_FR;
  const float* x = audio + (long long)b * n_samples + t0 * S;
  const int nload = (int)nfr * S + (K - S);
  for (int i = threadIdx.x; i < nload; i += blockDim.x) {
    const float v = x[i];
    xs[0][i] = v;
    if (i > 0) xs[1][i - 1] = v;
  }
  __syncthreads();
  // 4 channels per thread, 128 channel quads x 2 frame lanes per block; a warp writes 256 contiguous bytes per frame
  // and output tensor
  const int quads = channels / 4;
  const int lanes = blockDim.x / 128;
  for (int qg = threadIdx.x % 128; qg < quads; qg += 128) {
    const int c0 = qg * 4;
    f32x2 wf[4][K / 2];
    float bf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      const float mean = stats[2 * ((long long)b * channels + c)];
      const float rstd = stats[2 * ((long long)b * channels + c) + 1];
      const float g = gamma[c] * rstd;
#pragma unroll
      for (int k = 0; k < K / 2; ++k) wf[j][k] = f2_pack(w[c * K + 2 * k] * g, w[c * K + 2 * k + 1] * g);
      bf[j] = beta[c] - mean * g;
    }
    const long long base = ((long long)b * t_out + t0) * channels + c0;
#pragma unroll 2
    for (int f = threadIdx.x / 128; f < nfr; f += lanes) {
      const int odd = f & 1;
      const f32x2* wp = reinterpret_cast<const f32x2*>(&xs[odd][f * S - odd]);
      f32x2 win[K / 2];
#pragma unroll
      for (int k = 0; k < K / 2; ++k) win[k] = wp[k];
      float o[4], d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f32x2 a = f2_mul(wf[j][0], win[0]);
#pragma unroll
        for (int k = 1; k < K / 2; ++k) a = f2_fma(wf[j][k], win[k], a);
        float lo, hi;
        f2_unpack(a, lo, hi);
        o[j] = (lo + hi) + bf[j];
      }
      if (gprime) {   // training: also keep gelu'(z) so backward never recomputes the conv
        gelu_erf_both2(o[0], o[1], d[0], d[1]);
        gelu_erf_both2(o[2], o[3], d[2], d[3]);
      } else {
        gelu_erf2(o[0], o[1]);
        gelu_erf2(o[2], o[3]);
      }
      *reinterpret_cast<uint2*>(y + base + (long long)f * channels) =
          make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
      if (gprime)
        *reinterpret_cast<uint2*>(gprime + base + (long long)f * channels) =
            make_uint2(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]));
    }
  }
}

// ------------------------------------------------------------------ backward (single pass over dy and gelu'(z))
// dz = dy * gelu'(z) with gelu'(z) stored by the forward pass (bf16): no convolution / activation recompute, the
// kernel is two streaming reads plus 11 FMAs per element.
// partial[b][c][0] = sum dz, [2+k] = sum_t dz * x[S t + k]; [1] (= sum dz*xhat) follows algebraically in finalize.
constexpr int BWD_FR = 1024;
// The first version prefetched into registers: 64 bytes per thread in flight = 32 KiB per SM, which by Little's law is
// ~3.2 TB/s at the ~1.5 us loaded HBM latency -- exactly what it measured (0.98 ms) -- and it was issue-heavy (ten scalar
// FMAs per element).  A block's dy / gelu' tile is ONE contiguous chunk of memory per stream (channels-last, consecutive
// frames), so thread 0 streams it through a 5-stage shared-memory ring of bulk copies (160 KiB in flight per SM, no
// registers spent on prefetching) and every thread reads its 8-byte quads back from shared memory.  The ten window FMAs
// of a channel run as FIVE packed-pair FMAs (dz replicated x (win[k], win[k+1])); S = 5 is odd, so a second copy of the
// waveform slab shifted by one float keeps the 8-byte window-pair loads of odd frames aligned.  channels <= 512.
constexpr int BWD_STAGES = 5;
constexpr int BWD_STAGE_BYTES = 16384;                      // per stream and stage
constexpr int BWD_XS = BWD_FR * S + K + 2;                  // floats per waveform copy
constexpr int BWD_XS_BYTES = ((2 * BWD_XS * 4 + 127) / 128) * 128;
constexpr int BWD_RING_SMEM = BWD_XS_BYTES + BWD_STAGES * 2 * BWD_STAGE_BYTES + 128;
static_assert(4 * 512 * (K + 2) * 4 <= BWD_STAGES * 2 * BWD_STAGE_BYTES, "lane reduction reuses the ring");

__global__ void __launch_bounds__(512, 1) bwd_ring_kernel(const float* __restrict__ audio, const bf16* __restrict__ dy,
                                                          const bf16* __restrict__ gprime, float* __restrict__ partial,
                                                          long long n_samples, long long t_out, int channels) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t sm[];
  float* xs0 = reinterpret_cast<float*>(sm);                 // xs1[i] = xs0[i + 1]
  float* xs1 = xs0 + BWD_XS;
  uint8_t* ring = sm + BWD_XS_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + BWD_STAGES * 2 * BWD_STAGE_BYTES);
  uint64_t* empty = full + BWD_STAGES;
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * BWD_FR;
  const int nfr = (int)((t_out - t0) < BWD_FR ? (t_out - t0) : BWD_FR);
  const int row_bytes = channels * 2;
  const int F = BWD_STAGE_BYTES / row_bytes;                 // frames per stage (>= 16)
  const int nst = (nfr + F - 1) / F;
  const uint8_t* gdy = reinterpret_cast<const uint8_t*>(dy) + ((long long)b * t_out + t0) * row_bytes;
  const uint8_t* ggp = reinterpret_cast<const uint8_t*>(gprime) + ((long long)b * t_out + t0) * row_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < BWD_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], blockDim.x >> 5);               // one arrival per warp
    }
    fence_barrier_init();
  }
  __syncthreads();
  // thread 0 refills a slot with ONE bulk copy per stream (<= 16 KiB each).  Splitting a refill into eight 4 KiB pieces
  // issued by eight lanes was measured slower: 908 vs 779 us (profiles/r02x_ncu_rowwise.txt, r02fin).
  auto issue = [&](int st) {                                 // thread 0 only
    const int slot = st % BWD_STAGES;
    const int f0 = st * F;
    const uint32_t bytes = (uint32_t)((nfr - f0 < F ? nfr - f0 : F) * row_bytes);
    uint8_t* dst = ring + slot * 2 * BWD_STAGE_BYTES;
    mbar_expect_tx(&full[slot], 2 * bytes);
    bulk_load_1d(dst, gdy + (long long)f0 * row_bytes, bytes, &full[slot]);
    bulk_load_1d(dst + BWD_STAGE_BYTES, ggp + (long long)f0 * row_bytes, bytes, &full[slot]);
  };
  if (threadIdx.x == 0)
    for (int st = 0; st < nst && st < BWD_STAGES; ++st) issue(st);
  const float* x = audio + (long long)b * n_samples + t0 * S;
  const int nload = nfr * S + (K - S);
  for (int i = threadIdx.x; i < nload; i += blockDim.x) {
    const float v = x[i];
    xs0[i] = v;
    if (i > 0) xs1[i - 1] = v;
  }
  __syncthreads();
  const int quads = channels / 4;
  const int qg = threadIdx.x & 127, fl = threadIdx.x >> 7;   // 128 channel quads x 4 frame lanes
  const int lanes = blockDim.x >> 7;
  const bool active = qg < quads;
  f32x2 acc[4][K / 2];
  float acc0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc0[j] = 0.f;
#pragma unroll
    for (int k = 0; k < K / 2; ++k) acc[j][k] = f2_rep(0.f);
  }
  for (int st = 0; st < nst; ++st) {
    const int slot = st % BWD_STAGES;
    // refill the slot consumed one iteration ago (no block-wide barrier: the warps drift up to a ring apart; thread 0
    // only waits for the LAST stage's readers, who are at most one stage behind it)
    if (threadIdx.x == 0 && st >= 1 && st - 1 + BWD_STAGES < nst) {
      mbar_wait(&empty[(st - 1) % BWD_STAGES], (uint32_t)(((st - 1) / BWD_STAGES) & 1));
      fence_proxy_async_smem();                            // generic-proxy reads before the async-proxy refill
      issue(st - 1 + BWD_STAGES);
    }
    mbar_wait(&full[slot], (uint32_t)((st / BWD_STAGES) & 1));
    const uint8_t* sdy = ring + slot * 2 * BWD_STAGE_BYTES + qg * 8;
    const uint8_t* sgp = sdy + BWD_STAGE_BYTES;
    const int f0 = st * F;
    const int nf = nfr - f0 < F ? nfr - f0 : F;
    if (active) {
#pragma unroll 4
      for (int fi = fl; fi < nf; fi += lanes) {
        const uint2 u = *reinterpret_cast<const uint2*>(sdy + fi * row_bytes);
        const uint2 g = *reinterpret_cast<const uint2*>(sgp + fi * row_bytes);
        const int fs = f0 + fi;
        const int odd = fs & 1;
        const f32x2* wp = reinterpret_cast<const f32x2*>((odd ? xs1 : xs0) + fs * S - odd);
        f32x2 win[K / 2];
#pragma unroll
        for (int k = 0; k < K / 2; ++k) win[k] = wp[k];
        const float dz[4] = {bf16_lo(u.x) * bf16_lo(g.x), bf16_hi(u.x) * bf16_hi(g.x),
                             bf16_lo(u.y) * bf16_lo(g.y), bf16_hi(u.y) * bf16_hi(g.y)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc0[j] += dz[j];
          const f32x2 d2 = f2_rep(dz[j]);
#pragma unroll
          for (int k = 0; k < K / 2; ++k) acc[j][k] = f2_fma(d2, win[k], acc[j][k]);
        }
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);  // this warp is done with the slot
  }
  // Reduce the frame lanes through shared memory (the ring is idle now) and leave with COALESCED reductions: consecutive
  // threads add consecutive floats of partial[b][c][0..K+1] (4 sectors per warp request).  One atomic per thread and
  // accumulator straight from the registers was 44 requests of 32 sectors each per thread -- ncu showed the kernel's
  // first stall reason to be lg_throttle (profiles/r02w_conv0.txt).
  __syncthreads();
  float* red = reinterpret_cast<float*>(ring);               // [lanes][channels][K + 2]
  const int per_lane = channels * (K + 2);
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* pp = red + fl * per_lane + (qg * 4 + j) * (K + 2);
      pp[0] = acc0[j];
      pp[1] = 0.f;
#pragma unroll
      for (int k = 0; k < K / 2; ++k) {
        float lo, hi;
        f2_unpack(acc[j][k], lo, hi);
        pp[2 + 2 * k] = lo;
        pp[3 + 2 * k] = hi;
      }
    }
  }
  __syncthreads();
  float* gp = partial + (long long)b * per_lane;
  for (int i = threadIdx.x; i < per_lane; i += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[l * per_lane + i];
    if ((i % (K + 2)) != 1) atomicAdd(gp + i, t);
  }
}

// dw[c][j], dgamma[c], dbeta[c] from the per-(b,c) partial sums.  One WARP per output, lanes over the batch (the first
// version looped over the batch in one thread: 32 rounds of dependent L2-latency loads = 59 us for 6144 numbers).
__global__ void __launch_bounds__(256) bwd_finalize_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                                           const float* __restrict__ stats, const float* __restrict__ moments,
                                                           const float* __restrict__ partial, float* __restrict__ dw,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int batch,
                                                           int channels, long long t_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= channels * (K + 2)) return;
  const int c = i / (K + 2), j = i % (K + 2);
  const double inv_t = 1.0 / (double)t_out;
  double out = 0.0;
  for (int b = lane; b < batch; b += 32) {
    const float* pp = partial + ((long long)b * channels + c) * (K + 2);
    // sum_t dz*xhat = rstd * (sum_k w_k sum_t dz x[St+k] - mean * sum_t dz)   (xhat = (w.win - mean) * rstd)
    double pp1 = 0.0;
#pragma unroll
    for (int l = 0; l < K; ++l) pp1 += (double)w[c * K + l] * (double)pp[2 + l];
    const double mean = stats[2 * ((long long)b * channels + c)];
    const double rstd = stats[2 * ((long long)b * channels + c) + 1];
    pp1 = (pp1 - mean * (double)pp[0]) * rstd;
    if (j == K) {
      out += pp1;  // dgamma
    } else if (j == K + 1) {
      out += pp[0];  // dbeta
    } else {
      const float* mo = moments + (long long)b * NMOM;
      double wr = 0.0;
#pragma unroll
      for (int l = 0; l < K; ++l) wr += (double)w[c * K + l] * mo[K + l * K + j];
      const double sum_xhat_x = rstd * (wr - mean * mo[j]);
      out += (double)gamma[c] * rstd * ((double)pp[2 + j] - (double)pp[0] * inv_t * mo[j] - pp1 * inv_t * sum_xhat_x);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) out += __shfl_xor_sync(0xffffffffu, out, o);
  if (lane == 0) {
    if (j == K) dgamma[c] = (float)out;
    else if (j == K + 1) dbeta[c] = (float)out;
    else dw[c * K + j] = (float)out;
  }
}


// =====================================================================================
// feat_extract_norm = "layer" variant (HuBERT-large / wav2vec2-large-lv60):
//   Conv1d(1 -> C, k = 10, s = 5, bias) -> LayerNorm over the C channels of each frame -> GELU
//   hf:models/wav2vec2/modeling_wav2vec2.py:275-299 (Wav2Vec2LayerNormConvLayer), layer 0.
// One warp per frame, each lane owns C/32 channels; the conv is recomputed from the waveform in
// backward, so only the activated output is ever stored.
// =====================================================================================
constexpr int LN_MAXC = 16;  // channels per lane (C <= 512)

template <bool BWD>
__global__ void __launch_bounds__(256) ln_variant_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                                         const float* __restrict__ cbias, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, bf16* __restrict__ y,
                                                         const bf16* __restrict__ dy, bf16* __restrict__ dconv,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         long long n_samples, long long t_out, long long total_frames,
                                                         int channels, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int cpl = channels / 32;  // channels per lane, contiguous: [lane*cpl, lane*cpl + cpl)
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float wl[LN_MAXC][K], bl[LN_MAXC], gl[LN_MAXC], btl[LN_MAXC];
  float ag[LN_MAXC], ab[LN_MAXC];
#pragma unroll
  for (int j = 0; j < LN_MAXC; ++j) {
    ag[j] = ab[j] = 0.f;
    if (j < cpl) {
      const int c = lane * cpl + j;
#pragma unroll
      for (int k = 0; k < K; ++k) wl[j][k] = w[c * K + k];
      bl[j] = cbias ? cbias[c] : 0.f;
      gl[j] = gamma[c];
      btl[j] = beta[c];
    }
  }
  for (long long fr = warp_global; fr < total_frames; fr += nwarps) {
    const long long b = fr / t_out, t = fr % t_out;
    const float* x = audio + b * n_samples + t * S;
    float win[K];
#pragma unroll
    for (int k = 0; k < K; ++k) win[k] = __ldg(x + k);
    float v[LN_MAXC];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXC; ++j) {
      if (j < cpl) {
        float a = bl[j];
#pragma unroll
        for (int k = 0; k < K; ++k) a = fmaf(wl[j][k], win[k], a);
        v[j] = a;
        s += a;
      }
    }
    const float mean = warp_sum(s) / channels;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXC; ++j)
      if (j < cpl) sq += (v[j] - mean) * (v[j] - mean);
    const float rstd = rsqrtf(warp_sum(sq) / channels + eps);
    const long long off = fr * channels + lane * cpl;
    if (!BWD) {
      float o[LN_MAXC];
#pragma unroll
      for (int j = 0; j < LN_MAXC; ++j)
        if (j < cpl) o[j] = gelu_erf(fmaf((v[j] - mean) * rstd, gl[j], btl[j]));
#pragma unroll
      for (int j = 0; j < LN_MAXC; j += 2)
        if (j < cpl) *reinterpret_cast<uint32_t*>(y + off + j) = pack_bf16x2(o[j], o[j + 1]);
    } else {
      float g[LN_MAXC], xh[LN_MAXC];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < LN_MAXC; j += 2) {
        if (j < cpl) {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(dy + off + j);
          g[j] = bf16_lo(u);
          g[j + 1] = bf16_hi(u);
        }
      }
#pragma unroll
      for (int j = 0; j < LN_MAXC; ++j) {
        if (j < cpl) {
          xh[j] = (v[j] - mean) * rstd;
          const float dz = g[j] * gelu_erf_grad(fmaf(xh[j], gl[j], btl[j]));
          ag[j] += dz * xh[j];
          ab[j] += dz;
          g[j] = dz * gl[j];
          s1 += g[j];
          s2 += g[j] * xh[j];
        }
      }
      s1 = warp_sum(s1) / channels;
      s2 = warp_sum(s2) / channels;
      float o[LN_MAXC];
#pragma unroll
      for (int j = 0; j < LN_MAXC; ++j)
        if (j < cpl) o[j] = rstd * (g[j] - s1 - xh[j] * s2);
#pragma unroll
      for (int j = 0; j < LN_MAXC; j += 2)
        if (j < cpl) *reinterpret_cast<uint32_t*>(dconv + off + j) = pack_bf16x2(o[j], o[j + 1]);
    }
  }
  if (BWD) {
#pragma unroll
    for (int j = 0; j < LN_MAXC; ++j) {
      if (j < cpl) {
        atomicAdd(dgamma + lane * cpl + j, ag[j]);
        atomicAdd(dbeta + lane * cpl + j, ab[j]);
      }
    }
  }
}

// dw[c][k] += sum_{b,t} dconv[b,t,c] * x[b, S t + k],  dbias[c] += sum dconv   (fp32 atomics into zeroed buffers)
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ audio, const bf16* __restrict__ dconv,
                                                    float* __restrict__ dw, float* __restrict__ dbias,
                                                    long long n_samples, long long t_out, int channels) {
  pdl_trigger();
  pdl_wait();
  __shared__ float xs[BWD_FR * S + K];
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * BWD_FR;
  const long long nfr = (t_out - t0) < BWD_FR ? (t_out - t0) : BWD_FR;
  const float* x = audio + (long long)b * n_samples + t0 * S;
  const int nload = (int)nfr * S + (K - S);
  for (int i = threadIdx.x; i < nload; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  const int quads = channels / 4;
  const int lanes = blockDim.x / 128;
  for (int qg = threadIdx.x % 128; qg < quads; qg += 128) {
    const int c0 = qg * 4;
    float acc[4][K + 1];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < K + 1; ++k) acc[j][k] = 0.f;
    for (int f = threadIdx.x / 128; f < nfr; f += lanes) {
      const uint2 u = *reinterpret_cast<const uint2*>(dconv + ((long long)b * t_out + t0 + f) * channels + c0);
      const float d[4] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][K] += d[j];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[j][k] = fmaf(d[j], xs[f * S + k], acc[j][k]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int k = 0; k < K; ++k) atomicAdd(dw + (c0 + j) * K + k, acc[j][k]);
      if (dbias) atomicAdd(dbias + c0 + j, acc[j][K]);
    }
  }
}

}  // namespace conv0
}  // namespace smx

using namespace smx;
using namespace smx::conv0;

extern "C" {

int smx_conv0_stats(const float* audio, const float* w, float* moments, float* stats, float* partial_ws, int64_t batch,
                    int64_t n_samples, int64_t t_out, int channels, int ksize, int stride, float eps, void* stream) {
  SMX_REQUIRE(ksize == K && stride == S, "conv0: only kernel 10 / stride 5 is supported (got %d/%d)", ksize, stride);
  SMX_REQUIRE(t_out == (n_samples - K) / S + 1 && t_out > 0, "conv0: inconsistent t_out");
  SMX_REQUIRE(partial_ws != nullptr, "conv0_stats: partial_ws (batch * SMX_CONV0_MOMENT_BLOCKS * 65 floats) required");
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl(moments_kernel, dim3(dim3(SMX_CONV0_MOMENT_BLOCKS, (unsigned)batch)), dim3(256), 0, st, audio, partial_ws, n_samples, t_out);
  SMX_CHECK_CUDA(cudaGetLastError());
  launch_pdl(moments_reduce_kernel, dim3((unsigned)batch), dim3(128), 0, st, partial_ws, moments, SMX_CONV0_MOMENT_BLOCKS);
  SMX_CHECK_CUDA(cudaGetLastError());
  launch_pdl(stats_kernel, dim3((int)ceil_div(batch * channels, 256)), dim3(256), 0, st, w, moments, stats, (int)batch, channels, t_out, eps);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_conv0_gn_gelu_fwd(const float* audio, const float* w, const float* gamma, const float* beta,
                          const float* stats, void* y, void* gprime, int64_t batch, int64_t n_samples, int64_t t_out,
                          int channels, int ksize, int stride, void* stream) {
  SMX_REQUIRE(ksize == K && stride == S, "conv0: only kernel 10 / stride 5 is supported");
  SMX_REQUIRE(channels % 8 == 0, "conv0: channels must be a multiple of 8");
  launch_pdl(fwd_kernel, dim3(dim3((unsigned)ceil_div(t_out, FWD_FR), (unsigned)batch)), dim3(256), 0, (cudaStream_t)stream, 
      audio, w, gamma, beta, stats, (bf16*)y, (bf16*)gprime, n_samples, t_out, channels);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int smx_conv0_gn_gelu_bwd(const float* audio, const float* w, const float* gamma, const float* beta,
                          const float* stats, const float* moments, const void* dy, const void* gprime,
                          float* partial, float* dw, float* dgamma, float* dbeta, int64_t batch, int64_t n_samples,
                          int64_t t_out, int channels, int ksize, int stride, void* stream) {
  SMX_REQUIRE(gprime != nullptr, "conv0 bwd: needs the gelu'(z) tensor the training forward stored");
  SMX_REQUIRE(ksize == K && stride == S, "conv0: only kernel 10 / stride 5 is supported");
  SMX_REQUIRE(channels % 4 == 0, "conv0: channels must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  SMX_CHECK_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * (K + 2) * batch * channels, st));
  const dim3 grid((unsigned)ceil_div(t_out, BWD_FR), (unsigned)batch);
  SMX_REQUIRE(channels <= 512 && channels % 8 == 0, "conv0 bwd: channels %d unsupported (multiple of 8, <= 512)", channels);
  SMX_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(gprime)) & 15) == 0,
              "conv0 bwd: dy / gprime must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    SMX_CHECK_CUDA(cudaFuncSetAttribute(bwd_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_RING_SMEM));
    attr_set = true;
  }
  launch_pdl(bwd_ring_kernel, grid, dim3(512), BWD_RING_SMEM, st, audio, (const bf16*)dy, (const bf16*)gprime, partial,
             n_samples, t_out, channels);
  SMX_CHECK_CUDA(cudaGetLastError());
  launch_pdl(bwd_finalize_kernel, dim3((int)ceil_div(channels * (K + 2), 8)), dim3(256), 0, st, w, gamma, stats, moments, partial, dw,
                                                                             dgamma, dbeta, (int)batch, channels, t_out);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}

extern "C" {

int smx_conv0_ln_gelu_fwd(const float* audio, const float* w, const float* conv_bias, const float* gamma,
                          const float* beta, void* y, int64_t batch, int64_t n_samples, int64_t t_out, int channels,
                          int ksize, int stride, float eps, void* stream) {
  SMX_REQUIRE(ksize == K && stride == S, "conv0: only kernel 10 / stride 5 is supported");
  SMX_REQUIRE(channels % 64 == 0 && channels <= 32 * LN_MAXC, "conv0 (layer norm): channels must be a multiple of 64, <= 512");
  const long long frames = batch * t_out;
  long long grid = ceil_div(frames, 8 * 16);
  if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
  launch_pdl(ln_variant_kernel<false>, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, 
      audio, w, conv_bias, gamma, beta, (bf16*)y, nullptr, nullptr, nullptr, nullptr, n_samples, t_out, frames,
      channels, eps);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

/* dconv (bf16 [B,T,C], gradient w.r.t. the biased conv output) + dgamma/dbeta (zero-initialised, accumulated) */
int smx_conv0_ln_gelu_bwd(const float* audio, const float* w, const float* conv_bias, const float* gamma,
                          const float* beta, const void* dy, void* dconv, float* dgamma, float* dbeta, int64_t batch,
                          int64_t n_samples, int64_t t_out, int channels, int ksize, int stride, float eps,
                          void* stream) {
  SMX_REQUIRE(ksize == K && stride == S, "conv0: only kernel 10 / stride 5 is supported");
  SMX_REQUIRE(channels % 64 == 0 && channels <= 32 * LN_MAXC, "conv0 (layer norm): channels must be a multiple of 64, <= 512");
  const long long frames = batch * t_out;
  long long grid = ceil_div(frames, 8 * 16);
  if (grid > (long long)num_sms() * 4) grid = (long long)num_sms() * 4;
  launch_pdl(ln_variant_kernel<true>, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, 
      audio, w, conv_bias, gamma, beta, nullptr, (const bf16*)dy, (bf16*)dconv, dgamma, dbeta, n_samples, t_out,
      frames, channels, eps);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

/* dw [C][k] and dbias [C] (zero-initialised, accumulated) from dconv */
int smx_conv0_wgrad(const float* audio, const void* dconv, float* dw, float* dbias, int64_t batch, int64_t n_samples,
                    int64_t t_out, int channels, int ksize, int stride, void* stream) {
  SMX_REQUIRE(ksize == K && stride == S, "conv0: only kernel 10 / stride 5 is supported");
  SMX_REQUIRE(channels % 4 == 0, "conv0: channels must be a multiple of 4");
  launch_pdl(wgrad_kernel, dim3(dim3((unsigned)ceil_div(t_out, BWD_FR), (unsigned)batch)), dim3(256), 0, (cudaStream_t)stream, 
      audio, (const bf16*)dconv, dw, dbias, n_samples, t_out, channels);
  SMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}
