// Evaluation metrics that sit right of the training step: same
// gather-compare-reduce pattern as the loss.  One CTA per image.
#include "pld_common.cuh"

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

}  // extern "C"
