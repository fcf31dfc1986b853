// C ABI of the list stages: sampling (Philox / fed / MT stream), ListMLE fwd+bwd, fused step.
#include "pld_lists.cuh"

namespace pld {
int launch_lists_small(const ListParams& P, int src, bool loss, int num_sms, cudaStream_t st);
int launch_lists_large(const ListParams& P, int src, bool loss, int num_sms, cudaStream_t st);
int launch_acc_finalize(pld_ctx* ctx, float* grad, size_t n, float scale, int accumulate, cudaStream_t st);
int launch_offset_advance(pld_ctx* ctx, cudaStream_t st);
int mt_compact_images(pld_ctx* ctx, const int32_t* n_valid, int B, int need_per_image, const uint32_t* raw,
                      int64_t n_raw, int64_t* consumed_io, int32_t* sel_out, cudaStream_t st);

static int run_lists(pld_ctx* ctx, ListParams& P, int src, bool loss, int accumulate, cudaStream_t st) {
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(P.B > 0 && P.B <= 65535, "B out of range");
  PLD_REQUIRE(P.n >= 0, "negative list count");
  PLD_REQUIRE(P.K >= 1 && P.K <= PLD_MAX_RANKING_SIZE, "ranking_size must be in [1, 512]");
  PLD_REQUIRE(P.HW > 0 && P.HW <= PLD_MAX_PIXELS, "H*W out of range");
  PLD_REQUIRE((long long)P.B * P.n < (1ll << 31), "too many lists");
  P.status = ctx->d_status;
  P.ticket = ctx->d_ticket;
  if (loss) {
    // upper bound of the grid either launcher picks
    const int per_image_cap = lists_per_image_cap(ctx->num_sms, P.B);
    int rc = ctx->ensure_partials(per_image_cap * P.B + P.B);
    if (rc) return rc;
    P.partials = ctx->d_partials;
    const size_t gn = (size_t)P.B * (size_t)P.HW;
    if (P.grad != nullptr && ctx->deterministic) {
      int rc2 = ctx->ensure_acc(gn);
      if (rc2) return rc2;
      P.acc = ctx->d_acc;
      PLD_CUDA(cudaMemsetAsync(P.acc, 0, sizeof(long long) * gn, st));
    } else if (P.grad != nullptr && !accumulate) {
      PLD_CUDA(cudaMemsetAsync(P.grad, 0, sizeof(float) * gn, st));
    }
  }
  if (P.n == 0) {
    if (loss && P.grad != nullptr && P.acc != nullptr && !accumulate)   // deterministic mode: nothing to convert
      PLD_CUDA(cudaMemsetAsync(P.grad, 0, sizeof(float) * (size_t)P.B * (size_t)P.HW, st));
    if (loss) {
      if (P.loss) PLD_CUDA(cudaMemsetAsync(P.loss, 0, sizeof(float), st));
      if (P.loss_sum) PLD_CUDA(cudaMemsetAsync(P.loss_sum, 0, sizeof(double), st));
    }
    return PLD_OK;
  }
  const bool dev_off = ctx->use_device_offset && src == SRC_PHILOX;
  if (dev_off) P.offset_dev = ctx->d_offset;
  ctx->time_begin(st);
  int rc = (P.K <= 16) ? launch_lists_small(P, src, loss, ctx->num_sms, st)
                       : launch_lists_large(P, src, loss, ctx->num_sms, st);
  ctx->time_end(st);
  if (rc == PLD_OK && P.acc != nullptr)
    rc = launch_acc_finalize(ctx, P.grad, (size_t)P.B * (size_t)P.HW, P.scale, accumulate, st);
  if (rc == PLD_OK && dev_off) rc = launch_offset_advance(ctx, st);
  return rc;
}

static void set_rng(ListParams& P, uint64_t seed, uint64_t offset, int image_base) {
  P.seed_lo = (uint32_t)seed;
  P.seed_hi = (uint32_t)(seed >> 32);
  philox_round_keys(P.seed_lo, P.seed_hi, P.rk0, P.rk1);
  P.off_lo = (uint32_t)offset;
  P.off_hi16 = (uint32_t)((offset >> 32) & 0xFFFFu) << 16;
  P.image_base = image_base;
}
}  // namespace pld

using namespace pld;

extern "C" {

int pld_sample_lists_philox(pld_ctx* ctx, const float* gt, const int32_t* valid_flat, const int32_t* n_valid,
                            int B, int HW, int valid_stride, int K, int n, uint64_t seed, uint64_t offset,
                            int image_base, float* rankings, int32_t* sel_out, void* stream) {
  PLD_REQUIRE(ctx && gt && valid_flat && n_valid, "null argument");
  PLD_REQUIRE(rankings || sel_out, "no output requested");
  PLD_REQUIRE((offset >> 48) == 0, "offset must fit in 48 bits");
  ListParams P = {};
  P.gt = gt; P.valid_flat = valid_flat; P.n_valid = n_valid; P.pred = gt;
  P.rank_out = rankings; P.sel_out = sel_out;
  P.B = B; P.HW = HW; P.valid_stride = valid_stride; P.n = n; P.K = K; P.scale = 0.f;
  set_rng(P, seed, offset, image_base);
  return run_lists(ctx, P, SRC_PHILOX, false, 0, (cudaStream_t)stream);
}

int pld_sample_lists_fed(pld_ctx* ctx, const float* gt, const int32_t* valid_flat, const int32_t* n_valid,
                         int B, int HW, int valid_stride, int K, int n, const int32_t* sel, float* rankings,
                         void* stream) {
  PLD_REQUIRE(ctx && gt && valid_flat && n_valid && sel && rankings, "null argument");
  ListParams P = {};
  P.gt = gt; P.valid_flat = valid_flat; P.n_valid = n_valid; P.pred = gt;
  P.sel_in = sel; P.rank_out = rankings;
  P.B = B; P.HW = HW; P.valid_stride = valid_stride; P.n = n; P.K = K;
  return run_lists(ctx, P, SRC_FED_SEL, false, 0, (cudaStream_t)stream);
}

int pld_sample_lists_mt(pld_ctx* ctx, const float* gt, const int32_t* valid_flat, const int32_t* n_valid,
                        int B, int HW, int valid_stride, int K, int n, const uint32_t* raw, int64_t n_raw,
                        int64_t* consumed_io, float* rankings, int32_t* sel_out, void* stream) {
  PLD_REQUIRE(ctx && gt && valid_flat && n_valid && raw && consumed_io && rankings && sel_out, "null argument");
  PLD_REQUIRE(n_raw >= 0, "negative stream length");
  PLD_REQUIRE(B > 0 && n >= 0 && K >= 1 && K <= PLD_MAX_RANKING_SIZE, "bad shape");
  PLD_REQUIRE((long long)n * K < (1ll << 30), "too many draws per image");
  if (n == 0) return PLD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = mt_compact_images(ctx, n_valid, B, n * K, raw, n_raw, consumed_io, sel_out, st);
  if (rc) return rc;
  return pld_sample_lists_fed(ctx, gt, valid_flat, n_valid, B, HW, valid_stride, K, n, sel_out, rankings, stream);
}

int pld_listmle_fwd_bwd(pld_ctx* ctx, const float* rankings, const float* pred, int B, int R, int K, int HW,
                        float scale, float* loss, double* loss_sum, float* per_list, float* grad,
                        int accumulate, void* stream) {
  PLD_REQUIRE(ctx && rankings && pred && loss, "null argument");
  ListParams P = {};
  P.gt = pred; P.pred = pred; P.rank_in = rankings;
  P.per_list = per_list; P.grad = grad; P.loss = loss; P.loss_sum = loss_sum;
  P.B = B; P.HW = HW; P.n = R; P.K = K; P.scale = scale;
  return run_lists(ctx, P, SRC_FED_RANK, true, accumulate, (cudaStream_t)stream);
}

int pld_fused_sample_loss_bwd(pld_ctx* ctx, const float* gt, const int32_t* valid_flat, const int32_t* n_valid,
                              const float* pred, int B, int HW, int valid_stride, int K, int n, uint64_t seed,
                              uint64_t offset, int image_base, float scale, float* rankings, float* loss,
                              double* loss_sum, float* per_list, float* grad, int accumulate, void* stream) {
  PLD_REQUIRE(ctx && gt && valid_flat && n_valid && pred && loss, "null argument");
  PLD_REQUIRE((offset >> 48) == 0, "offset must fit in 48 bits");
  ListParams P = {};
  P.gt = gt; P.valid_flat = valid_flat; P.n_valid = n_valid; P.pred = pred;
  P.rank_out = rankings; P.per_list = per_list; P.grad = grad; P.loss = loss; P.loss_sum = loss_sum;
  P.B = B; P.HW = HW; P.valid_stride = valid_stride; P.n = n; P.K = K; P.scale = scale;
  set_rng(P, seed, offset, image_base);
  return run_lists(ctx, P, SRC_PHILOX, true, accumulate, (cudaStream_t)stream);
}

}  // extern "C"
