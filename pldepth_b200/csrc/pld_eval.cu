// Evaluation metrics that sit right of the training step: same
// gather-compare-reduce pattern as the loss.  One CTA per image.
#include "pld_common.cuh"
#include "pld_score.cuh"

namespace pld {

// ordinal_error (pldepth/active_learning/metrics.py:60-70): fraction of the `num` fixed pixel pairs
// whose predicted order (>) disagrees with the ground-truth order.
__global__ void __launch_bounds__(256) ordinal_error_kernel(const float* __restrict__ op, const float* __restrict__ gt,
                                                            const int32_t* __restrict__ idx0,
                                                            const int32_t* __restrict__ idx1, int HW, int num,
                                                            float* __restrict__ err, int* status) {
  const float* o = op + (size_t)blockIdx.x * HW;
  const float* g = gt + (size_t)blockIdx.x * HW;
  int agree = 0, bad = 0;
  for (int i = threadIdx.x; i < num; i += 256) {
    int a = __ldg(idx0 + i), b = __ldg(idx1 + i);
    if (a < 0 || a >= HW || b < 0 || b >= HW) { bad = PLD_ST_BAD_INDEX; a = b = 0; }
    const bool oo = __ldg(o + a) > __ldg(o + b);
    const bool go = __ldg(g + a) > __ldg(g + b);
    agree += (oo == go) ? 1 : 0;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) agree += __shfl_xor_sync(0xffffffffu, agree, s);
  __shared__ int sw[8];
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = agree;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += sw[i];
    // 1 - accuracy, accuracy = agree / num in float64 like NumPy, rounded once to float32 on store
    err[blockIdx.x] = (float)(1.0 - (double)t / (double)num);
  }
  if (bad) atomicOr(status, bad);
}

// calc_d (metrics.py:92-110): min-max normalise the prediction to [0,1], take the `n` fixed sample
// pixels, sort prediction and ground truth samples ascending, rel = 1/(x+1),
// DCG = sum rel_i / log2(i+2); returns DCG(pred) / DCG(gt).  n <= 1024 (bitonic sort in shared memory).
__device__ void smem_bitonic_asc(float* v, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const float a = v[i], b = v[ixj];
          if ((a > b) == up) { v[i] = b; v[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) ndcg_kernel(const float* __restrict__ op, const float* __restrict__ gt,
                                                   const int32_t* __restrict__ ids, int HW, int n, int n_pow2,
                                                   float* __restrict__ out, int* status) {
  extern __shared__ float sm[];
  float* so = sm;            // [n_pow2] prediction samples
  float* sg = sm + n_pow2;   // [n_pow2] ground-truth samples
  __shared__ float s_red[2][8];
  __shared__ double s_sum[2][8];
  const float* o = op + (size_t)blockIdx.x * HW;
  const float* g = gt + (size_t)blockIdx.x * HW;
  float mn = 3.402823466e38f, mx = -3.402823466e38f;
  for (int i = threadIdx.x; i < HW; i += 256) {
    const float v = __ldg(o + i);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = mn; s_red[1][threadIdx.x >> 5] = mx; }
  __syncthreads();
  for (int i = 0; i < 8; ++i) { mn = fminf(mn, s_red[0][i]); mx = fmaxf(mx, s_red[1][i]); }
  const double scale = (mx > mn) ? 1.0 / ((double)mx - (double)mn) : 0.0;
  int bad = 0;
  for (int i = threadIdx.x; i < n_pow2; i += 256) {
    float a = 3.402823466e38f, b = 3.402823466e38f;   // pads sort to the end
    if (i < n) {
      int id = __ldg(ids + i);
      if (id < 0 || id >= HW) { bad = PLD_ST_BAD_INDEX; id = 0; }
      a = (float)(((double)__ldg(o + id) - (double)mn) * scale);
      b = __ldg(g + id);
    }
    so[i] = a;
    sg[i] = b;
  }
  __syncthreads();
  smem_bitonic_asc(so, n_pow2);
  smem_bitonic_asc(sg, n_pow2);
  double dp = 0.0, dg = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double w = 1.0 / log2((double)i + 2.0);
    dp += (1.0 / ((double)so[i] + 1.0)) * w;
    dg += (1.0 / ((double)sg[i] + 1.0)) * w;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    dp += __shfl_xor_sync(0xffffffffu, dp, s);
    dg += __shfl_xor_sync(0xffffffffu, dg, s);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[0][threadIdx.x >> 5] = dp; s_sum[1][threadIdx.x >> 5] = dg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += s_sum[0][i]; b += s_sum[1][i]; }
    out[blockIdx.x] = (float)(a / b);
  }
  if (bad) atomicOr(status, bad);
}

// prepare_fully_fledged_loss_input (depth_utils.py:39-61) as a standalone op: batched gather of the
// predictions at the flat indices of the rankings + de-interleaved labels.
__global__ void __launch_bounds__(256) gather_predictions_kernel(const float2* __restrict__ rankings,
                                                                const float* __restrict__ pred, size_t per_image,
                                                                int HW, size_t total, float* __restrict__ selected,
                                                                float* __restrict__ labels, int* status) {
  int bad = 0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const float2 v = __ldg(rankings + i);
    const size_t b = i / per_image;
    int q = (int)v.x;
    if (q < 0 || q >= HW) { bad = PLD_ST_BAD_INDEX; q = 0; }
    selected[i] = __ldg(pred + b * (size_t)HW + q);
    if (labels != nullptr) labels[i] = v.y;
  }
  if (bad) atomicOr(status, bad);
}

// ------------------------------------------------------------------------------------------------------------------
// Evaluation list generators (pldepth/data/providers/generic_ranking_provider.py:80-111, 180-215), MT19937-compatible.
//
// generate_ordinal_pairs draws, per pair, x0 = randint(H), y0 = randint(W), x1 = randint(H), y1 = randint(W) from the
// global NumPy stream: masked rejection with a bound that ALTERNATES between H and W.  Whether a raw word is accepted
// therefore depends on the parity of the draws accepted before it -- a two-state automaton.  One warp walks the
// stream 32 words at a time and composes the per-word transition functions with a shuffle scan
// (state -> (next state, accepted count)), so every lane knows the bound that applies to its word and the rank of its
// draw.  Evaluation lists are generated once per dataset (the reference caches them to .npy), so one warp is enough:
// about 1 M draws per 4 ms.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t np_mask_eval(uint32_t M) {  // smallest 2^k - 1 >= M - 1
  const uint32_t r = M - 1u;
  return r == 0 ? 0u : (0xFFFFFFFFu >> __clz(r));
}

__global__ void __launch_bounds__(32) mt_alternating_draws_kernel(const uint32_t* __restrict__ raw, long long n_raw,
                                                                  long long* __restrict__ consumed_io, uint32_t M0,
                                                                  uint32_t M1, long long need, int32_t* __restrict__ draws,
                                                                  int* status) {
  const int lane = threadIdx.x;
  const uint32_t m0 = np_mask_eval(M0), m1 = np_mask_eval(M1);
  long long p = *consumed_io, rank = 0, last = -1;
  int S = 0;
  while (rank < need && p < n_raw) {
    const long long idx = p + lane;
    const bool valid = idx < n_raw;
    const uint32_t w = valid ? __ldg(raw + idx) : 0u;
    const int a0 = (valid && (w & m0) <= M0 - 1u) ? 1 : 0, a1 = (valid && (w & m1) <= M1 - 1u) ? 1 : 0;
    // transition function of this word: from state s go to s ^ a_s, having accepted a_s draws
    int e0 = a0, e1 = 1 ^ a1, n0 = a0, n1 = a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int pe0 = __shfl_up_sync(0xffffffffu, e0, o), pe1 = __shfl_up_sync(0xffffffffu, e1, o);
      const int pn0 = __shfl_up_sync(0xffffffffu, n0, o), pn1 = __shfl_up_sync(0xffffffffu, n1, o);
      if (lane >= o) {   // compose: the earlier segment first, then this one
        const int ne0 = pe0 ? e1 : e0, ne1 = pe1 ? e1 : e0;
        const int nn0 = pn0 + (pe0 ? n1 : n0), nn1 = pn1 + (pe1 ? n1 : n0);
        e0 = ne0; e1 = ne1; n0 = nn0; n1 = nn1;
      }
    }
    // exclusive prefix = inclusive of the lane before (identity for lane 0)
    int xe0 = __shfl_up_sync(0xffffffffu, e0, 1), xe1 = __shfl_up_sync(0xffffffffu, e1, 1);
    int xn0 = __shfl_up_sync(0xffffffffu, n0, 1), xn1 = __shfl_up_sync(0xffffffffu, n1, 1);
    if (lane == 0) { xe0 = 0; xe1 = 1; xn0 = 0; xn1 = 0; }
    const int st = S ? xe1 : xe0;
    const long long r = rank + (S ? xn1 : xn0);
    const bool acc = st ? (a1 != 0) : (a0 != 0);
    if (acc && r < need) {
      draws[r] = (int32_t)(w & (st ? m1 : m0));
      if (r == need - 1) last = idx + 1;
    }
    const int te0 = __shfl_sync(0xffffffffu, e0, 31), te1 = __shfl_sync(0xffffffffu, e1, 31);
    const int tn0 = __shfl_sync(0xffffffffu, n0, 31), tn1 = __shfl_sync(0xffffffffu, n1, 31);
    rank += S ? tn1 : tn0;
    S = S ? te1 : te0;
    p += 32;
  }
  // the lane that produced the last needed draw knows where the stream stands
  long long end = last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long t = __shfl_xor_sync(0xffffffffu, end, o);
    end = t > end ? t : end;
  }
  if (lane == 0) {
    if (need > 0 && end < 0) { atomicOr(status, PLD_ST_MT_EXHAUSTED); end = n_raw; }
    if (need > 0) *consumed_io = end;
  }
}

// get_depth_relation (depth_utils.py:5-21) with NumPy's scalar promotion: +1 / -1 / 0
template <typename T>
__device__ __forceinline__ int depth_relation(float z0, float z1, bool thresholded, const ScoreCfg& C) {
  if (!thresholded) return z0 > z1 ? 1 : (z0 < z1 ? -1 : 0);
  if (sizeof(T) == 4) {
    const float r = __fdiv_rn(__fadd_rn(z0, 1e-10f), __fadd_rn(z1, 1e-10f));
    return r >= C.thr_hi_f ? 1 : (r <= C.thr_lo_f ? -1 : 0);
  }
  const double r = __ddiv_rn(__dadd_rn((double)z0, 1e-10), __dadd_rn((double)z1, 1e-10));
  return r >= C.thr_hi ? 1 : (r <= C.thr_lo ? -1 : 0);
}

__global__ void __launch_bounds__(256) ordinal_pairs_kernel(const float* __restrict__ gt, const int32_t* __restrict__ draws,
                                                            int W, int HW, long long per_image, long long total,
                                                            int thresholded, int invert, ScoreCfg C,
                                                            float* __restrict__ out) {
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < total; q += (long long)gridDim.x * 256) {
    const long long b = q / per_image;
    const int x0 = draws[4 * q + 0], y0 = draws[4 * q + 1], x1 = draws[4 * q + 2], y1 = draws[4 * q + 3];
    const int p0 = x0 * W + y0, p1 = x1 * W + y1;
    const float z0 = __ldg(gt + b * HW + p0), z1 = __ldg(gt + b * HW + p1);
    int rel = (C.promotion == PLD_PROMOTION_NEP50) ? depth_relation<float>(z0, z1, thresholded != 0, C)
                                                   : depth_relation<double>(z0, z1, thresholded != 0, C);
    if (invert) rel = -rel;
    float* o = out + 5 * q;
    o[0] = (float)p0; o[1] = (float)p1; o[2] = (float)rel; o[3] = z0; o[4] = z1;
  }
}

// generate_rankings with invert_relation_sign (generic_ranking_provider.py:200-206): lists ordered by ORIGINAL depth
// ascending (= the descending order reversed, ties: earlier draw first) and depths stored as 1 / (d + 1) (float64
// arithmetic, rounded once to float32).  In place, one thread per (list, mirrored pair of positions).
__global__ void __launch_bounds__(256) invert_rankings_kernel(float2* __restrict__ rk, long long n_lists, int K) {
  const int half = (K + 1) / 2;
  const long long total = n_lists * half;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const long long l = t / half;
    const int i = (int)(t - l * half), j = K - 1 - i;
    float2* row = rk + l * K;
    const float2 a = row[i], c = row[j];
    const float2 ai = make_float2(a.x, (float)(1.0 / ((double)a.y + 1.0)));
    const float2 ci = make_float2(c.x, (float)(1.0 / ((double)c.y + 1.0)));
    row[j] = ai;
    if (i != j) row[i] = ci;
  }
}

}  // namespace pld

using namespace pld;

extern "C" {

int pld_ordinal_error(pld_ctx* ctx, const float* pred, const float* gt, const int32_t* idx0, const int32_t* idx1,
                      int N, int HW, int num, float* err, void* stream) {
  PLD_REQUIRE(ctx && pred && gt && idx0 && idx1 && err, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(N > 0 && HW > 0 && num > 0, "bad shape");
  ordinal_error_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(pred, gt, idx0, idx1, HW, num, err, ctx->d_status);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_gather_predictions(pld_ctx* ctx, const float* rankings, const float* pred, int B, int R, int K, int HW,
                           float* selected, float* labels, void* stream) {
  PLD_REQUIRE(ctx && rankings && pred && selected, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(B > 0 && R >= 0 && K >= 1 && HW > 0, "bad shape");
  const size_t per_image = (size_t)R * K, total = per_image * B;
  if (total == 0) return PLD_OK;
  int gx = (int)((total + 255) / 256);
  if (gx > ctx->num_sms * 16) gx = ctx->num_sms * 16;
  gather_predictions_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(rankings), pred,
                                                                  per_image, HW, total, selected, labels, ctx->d_status);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_ndcg(pld_ctx* ctx, const float* pred, const float* gt, const int32_t* ids, int N, int HW, int n, float* out,
             void* stream) {
  PLD_REQUIRE(ctx && pred && gt && ids && out, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(N > 0 && HW > 0 && n > 0 && n <= 1024, "list_size must be in [1, 1024]");
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  ndcg_kernel<<<N, 256, sizeof(float) * 2 * (size_t)p2, (cudaStream_t)stream>>>(pred, gt, ids, HW, n, p2, out,
                                                                              ctx->d_status);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_eval_ordinal_pairs_mt(pld_ctx* ctx, const float* gt, int N, int H, int W, int n_pairs, double threshold,
                              int invert_sign, int promotion, const uint32_t* raw, int64_t n_raw, int64_t* consumed_io,
                              float* pairs_out, void* stream) {
  PLD_REQUIRE(ctx && gt && raw && consumed_io && pairs_out, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(N > 0 && n_pairs >= 0 && n_raw >= 0, "bad shape");
  PLD_REQUIRE(H > 1 && W > 1 && (long long)H * W <= PLD_MAX_PIXELS, "maps must be at least 2 x 2 (randint(1) consumes no word)");
  PLD_REQUIRE(promotion == PLD_PROMOTION_NEP50 || promotion == PLD_PROMOTION_LEGACY, "bad promotion");
  const long long total = (long long)N * n_pairs;
  if (total == 0) return PLD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ctx->ensure_scratch(sizeof(int32_t) * 4 * (size_t)total);
  if (rc) return rc;
  int32_t* draws = (int32_t*)ctx->d_scratch;
  mt_alternating_draws_kernel<<<1, 32, 0, st>>>(raw, (long long)n_raw, (long long*)consumed_io, (uint32_t)H, (uint32_t)W,
                                                4 * total, draws, ctx->d_status);
  PLD_CHECK_LAUNCH();
  const bool thresholded = threshold >= 0.0;          // negative = the reference's threshold=None (plain comparison)
  const ScoreCfg C = make_score_cfg(nullptr, PLD_STRATEGY_THRESHOLDED, thresholded ? threshold : 0.0, 0.0, promotion);
  int gx = (int)((total + 255) / 256);
  if (gx > ctx->num_sms * 8) gx = ctx->num_sms * 8;
  ordinal_pairs_kernel<<<gx, 256, 0, st>>>(gt, draws, W, H * W, (long long)n_pairs, total, thresholded ? 1 : 0,
                                           invert_sign ? 1 : 0, C, pairs_out);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_eval_invert_rankings(pld_ctx* ctx, float* rankings, int64_t n_lists, int K, void* stream) {
  PLD_REQUIRE(ctx && rankings, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(n_lists >= 0 && K >= 1, "bad shape");
  if (n_lists == 0) return PLD_OK;
  const long long total = (long long)n_lists * ((K + 1) / 2);
  int gx = (int)((total + 255) / 256);
  if (gx > ctx->num_sms * 8) gx = ctx->num_sms * 8;
  invert_rankings_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(rankings), (long long)n_lists, K);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

}  // extern "C"
