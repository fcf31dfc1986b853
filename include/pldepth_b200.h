/*
 * pldepth_b200 -- C ABI of the B200-native PLDepth hot path
 * (ranking sampling -> gather -> ListMLE / Plackett-Luce NLL forward + backward).
 *
 * This is the drop-in boundary: plain pointers and sizes, no framework types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree,
 * praneeth-b/PLDepth).  The reference has no FFI of its own (it is pure Python on top of
 * NumPy / TensorFlow / TF-Ranking); the Python host side in pldepth_b200/ binds these symbols
 * with ctypes and mirrors the reference's classes.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - all functions return 0 on success, a negative PLD_E* code on failure and never throw;
 *     pld_last_error() returns a thread-local message for the last failure on this thread;
 *   - all array arguments are DEVICE pointers unless the name ends in _host; tensors are dense
 *     row-major, float32 / int32 / uint32 / float64 as declared, 16-byte aligned;
 *   - the caller allocates every output; the library only owns the context's private scratch;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = legacy default
 *     stream) and the call returns without synchronising, except where stated;
 *   - a context (scratch buffers, loss partials, device status word) must not be used from
 *     two threads / streams at once: create one per worker thread.
 *
 * Shapes: B images, H x W pixels (HW), mask Hm x Wm, K = ranking_size (1..512),
 * n = lists drawn per image, R = lists kept per image, L = B*R lists.
 * A "ranking" is K (flat_index, depth) float32 pairs, depth-descending: float32[..., K, 2],
 * the reference's on-the-wire format (pldepth/data/sampling.py:142-143).
 */
#ifndef PLDEPTH_B200_H
#define PLDEPTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLD_OK 0
#define PLD_EINVAL (-1)   /* bad argument (null pointer, K out of range, ...) */
#define PLD_ECUDA (-2)    /* a CUDA runtime call failed */
#define PLD_ENOMEM (-3)   /* scratch allocation failed */
#define PLD_ESTATE (-4)   /* device status word reports a data error (see pld_ctx_status) */

#define PLD_MAX_RANKING_SIZE 512
#define PLD_MAX_PIXELS (1 << 23) /* flat indices are packed into 23 bits inside the sort key */

/* bits of the device status word (pld_ctx_status) */
#define PLD_ST_EMPTY_MASK 1   /* an image has no valid mask pixel (reference: randint(0) raises) */
#define PLD_ST_BAD_INDEX 2    /* a fed flat index / selection is outside its map */
#define PLD_ST_MT_EXHAUSTED 4 /* the fed MT19937 word stream ended before all draws were accepted */
#define PLD_ST_INTERNAL 8     /* a self-check of the library failed (never expected; results of the call are void) */

/* strategies of pldepth/data/sampling.py */
#define PLD_STRATEGY_PURELY 0      /* PurelyMaskedRandomSamplingStrategy, sampling.py:106-150 */
#define PLD_STRATEGY_MASKED 1      /* MaskedRandomSamplingStrategy, sampling.py:153-169 */
#define PLD_STRATEGY_THRESHOLDED 2 /* ThresholdedMaskedRandomSamplingStrategy, sampling.py:172-208 */
#define PLD_STRATEGY_INFORMATION 3 /* InformationScoreBasedSampling, sampling.py:211-239 */

/* NumPy scalar-promotion flavour of the score arithmetic (DESIGN.md "score arithmetic") */
#define PLD_PROMOTION_NEP50 0  /* NumPy >= 2: float32 accumulators */
#define PLD_PROMOTION_LEGACY 1 /* NumPy 1.x (the reference's pinned 1.19.5): float64 */

typedef struct pld_ctx pld_ctx;

/* ---- library / context ------------------------------------------------------------------- */
int pld_version(void);
const char* pld_last_error(void);
/* Number of kernels this library has launched in this process (for bench accounting). */
uint64_t pld_launch_count(void);

int pld_ctx_create(int device, pld_ctx** out);
int pld_ctx_destroy(pld_ctx* ctx);
/* Synchronises `stream`, returns the OR of PLD_ST_* bits raised since the last call and clears
 * them.  This is how data errors that the reference reports as Python exceptions surface. */
int pld_ctx_status(pld_ctx* ctx, void* stream, int* status_host);

/* Deterministic mode (off by default): gradient contributions are accumulated with 64-bit fixed-point
 * (2^-32) integer atomics in context scratch and converted once, so the dense gradient is
 * bit-reproducible run to run and independent of launch geometry (float atomics are not).  Costs an
 * extra 8 B/pixel scratch, a memset and a conversion pass. */
int pld_ctx_set_deterministic(pld_ctx* ctx, int on);

/* Device-resident Philox offset (off by default).  When enabled, every Philox entry point ignores its by-value
 * `offset` argument, reads the offset from a counter in device memory instead (set to `start` here) and
 * advances it by one at the end of the call.  The whole step then has no per-step host arguments, so it can be
 * captured once in a CUDA graph and replayed with fresh random numbers each time.  (This call synchronises.) */
int pld_ctx_device_offset(pld_ctx* ctx, int enable, uint64_t start);

/* Measurement hook: with slots > 0 the library records a CUDA event pair on the launch stream
 * around every list-kernel launch (the dominant kernel of each entry point below) into a ring of
 * `slots` pairs; pld_ctx_kernel_times waits for them, returns the durations in ms and resets the
 * ring.  slots == 0 switches it off (default). */
int pld_ctx_kernel_timing(pld_ctx* ctx, int slots);
int pld_ctx_kernel_times(pld_ctx* ctx, float* ms_host, int capacity, int* count_host);

/* ---- stage 1a: valid-pixel table ----------------------------------------------------------
 * Replaces `mask_points = np.where(mask > 0)` (sampling.py:135), determine_x_y_scales
 * (sampling.py:124-129) and the per-draw `int(rows[sel]*x_scale) * W + int(cols[sel]*y_scale)`
 * (sampling.py:115-119): valid_flat[b][j] = flat image index of the j-th valid mask pixel in
 * row-major order, n_valid[b] = their number.
 * Identity shortcut: when the mask has the image's resolution and every pixel of image b is
 * valid, the table would be valid_flat[b][j] == j; the row is then NOT written and n_valid[b] is
 * stored NEGATED (-Hm*Wm).  Every consumer below takes M = |n_valid[b]| and skips the lookup
 * for negative counts.
 *   mask f32[B,Hm,Wm] -> valid_flat i32[B,Hm*Wm], n_valid i32[B] */
int pld_mask_compact(pld_ctx* ctx, const float* mask, int B, int Hm, int Wm, int H, int W,
                     int32_t* valid_flat, int32_t* n_valid, void* stream);

/* ---- stage 1b: draw + order lists ---------------------------------------------------------
 * Replaces PurelyMaskedRandomSamplingStrategy.sample_masked_rankings /
 * sample_single_masked_ranking (sampling.py:110-145): per list K draws sel in [0, n_valid[b]),
 * p = valid_flat[b][sel], g = gt[b][p]; the list is ordered by g descending (ties: later draw
 * first) and written as (float(p), g) pairs.
 *   gt f32[B,HW], valid_flat i32[B,valid_stride], n_valid i32[B]
 *   -> rankings f32[B,n,K,2], sel_out i32[B,n,K] (nullable; the draws, in draw order)
 *
 * Philox mode (throughput): Philox4x32-10, key = seed, counter = (list, image_base + b, word
 * block, offset); unbiased Lemire mapping to [0, M).  Same (seed, offset, image, list, draw)
 * always gives the same selection, independent of launch geometry and of sharding. */
int pld_sample_lists_philox(pld_ctx* ctx, const float* gt, const int32_t* valid_flat,
                            const int32_t* n_valid, int B, int HW, int valid_stride, int K, int n,
                            uint64_t seed, uint64_t offset, int image_base, float* rankings,
                            int32_t* sel_out, void* stream);

/* Fed-selection mode: sel i32[B,n,K] are the draws (e.g. from np.random.randint) -- the "same
 * fed indices" parity mode. */
int pld_sample_lists_fed(pld_ctx* ctx, const float* gt, const int32_t* valid_flat,
                         const int32_t* n_valid, int B, int HW, int valid_stride, int K, int n,
                         const int32_t* sel, float* rankings, void* stream);

/* MT19937-stream mode (bit-exact with `np.random.randint(M)`, sampling.py:113): consumes raw
 * 32-bit MT19937 outputs `raw[n_raw]` with NumPy's masked rejection (mask = 2^ceil(log2 M) - 1,
 * reject > M-1; M == 1 consumes nothing), image after image, starting at word *consumed_io and
 * leaving the next unconsumed position there (device int64).  Raises PLD_ST_MT_EXHAUSTED in the
 * status word if raw is too short.  sel_out i32[B,n,K] is required (used as scratch). */
int pld_sample_lists_mt(pld_ctx* ctx, const float* gt, const int32_t* valid_flat,
                        const int32_t* n_valid, int B, int HW, int valid_stride, int K, int n,
                        const uint32_t* raw, int64_t n_raw, int64_t* consumed_io,
                        float* rankings, int32_t* sel_out, void* stream);

/* MT19937 generator on the device (np.random.seed(int) + raw 32-bit outputs).
 * state u32[624] + pos i32[1] live in caller memory (device); init fills them from a seed like
 * numpy's mt19937_seed, generate writes n tempered words and advances (state, pos). */
int pld_mt19937_init(pld_ctx* ctx, uint32_t seed, uint32_t* state, int32_t* pos, void* stream);
int pld_mt19937_generate(pld_ctx* ctx, uint32_t* state, int32_t* pos, uint32_t* out, int64_t n,
                         void* stream);

/* ---- stage 1c: score candidates, keep the best R ------------------------------------------
 * Replaces the tails of Masked / Thresholded / InformationScore sample_masked_point_batch
 * (sampling.py:161-169, 194-208, 219-239) and get_depth_relation (depth_utils.py:5-21).
 *   gt_minmax f32[B,2] (min, max of each gt map; INFORMATION only, see pld_gt_minmax)
 *   rankings f32[B,n,K,2] -> scores f64[B,n] */
int pld_gt_minmax(pld_ctx* ctx, const float* gt, int B, int HW, float* gt_minmax, void* stream);
int pld_score_lists(pld_ctx* ctx, const float* rankings, const float* gt_minmax, int B, int n,
                    int K, int strategy, double threshold, double equality_penalty, int promotion,
                    double* scores, void* stream);
/* `result[np.argsort(scores)[::-1]][:R]` per image (ties: larger candidate index first).
 *   -> rankings_out f32[B,R,K,2], order_out i32[B,R] (nullable) */
int pld_select_top(pld_ctx* ctx, const double* scores, const float* rankings, int B, int n, int K,
                   int R, float* rankings_out, int32_t* order_out, void* stream);

/* ---- stages 2+3: gather + ListMLE NLL forward / backward ----------------------------------
 * Replaces prepare_fully_fledged_loss_input (depth_utils.py:39-61), TF-Ranking 0.3.1
 * ListMLELoss.compute_unreduced_loss behind FullyFledgedMetaBatchListMLELoss
 * (losses/nll_loss.py:43-62), the Keras reduction of HourglassNegativeLogLikelihood
 * (nll_loss.py:32-40) and the autodiff backward (gather gradient = scatter-add).
 *   rankings f32[B,R,K,2], pred f32[B,HW]
 *   -> loss f32[1]  = scale * sum of per-list NLL      (overwritten)
 *      loss_sum f64[1] = unscaled sum (nullable, overwritten; multi-GPU all-reduces this)
 *      per_list f32[B*R] (nullable) unscaled per-list NLL
 *      grad f32[B,HW] (nullable = forward only): scale * d(sum NLL)/d pred, zeroed by the
 *      call unless accumulate != 0, duplicates accumulate. */
int pld_listmle_fwd_bwd(pld_ctx* ctx, const float* rankings, const float* pred, int B, int R,
                        int K, int HW, float scale, float* loss, double* loss_sum,
                        float* per_list, float* grad, int accumulate, void* stream);

/* ---- fused step: stages 1b + 2 + 3 in one launch ------------------------------------------
 * Philox draws -> gt gather -> order -> (optionally emit rankings) -> pred gather -> NLL ->
 * gradient scatter-add.  Same outputs as pld_sample_lists_philox followed by
 * pld_listmle_fwd_bwd on its rankings (n == R: the reference's core sampler,
 * sample_masked_rankings, feeding the loss).  rankings may be NULL. */
int pld_fused_sample_loss_bwd(pld_ctx* ctx, const float* gt, const int32_t* valid_flat,
                              const int32_t* n_valid, const float* pred, int B, int HW,
                              int valid_stride, int K, int n, uint64_t seed, uint64_t offset,
                              int image_base, float scale, float* rankings, float* loss,
                              double* loss_sum, float* per_list, float* grad, int accumulate,
                              void* stream);

/* Stage 2 alone: prepare_fully_fledged_loss_input (depth_utils.py:39-61) -- selected[b*R*K + t] =
 * pred[b][int32(rankings[b][t].index)], labels = rankings[...].depth; both shaped [B*R, K].
 * labels may be NULL.  (The fused kernels do this gather in registers; this entry point exists for callers
 * that want the reference's intermediate tensors.) */
int pld_gather_predictions(pld_ctx* ctx, const float* rankings, const float* pred, int B, int R, int K, int HW,
                           float* selected, float* labels, void* stream);

/* ---- whole step in one call: stages 1a + 1b + 2 + 3 -------------------------------------------
 * What one training step of the reference does between the data pipeline and the decoder's
 * backward: np.where(mask) (sampling.py:135) + sample_masked_rankings (sampling.py:131-145, n == R
 * lists per image) + prepare_fully_fledged_loss_input (depth_utils.py:39-61) + ListMLE loss and
 * its gradient (nll_loss.py:32-62).  Three launches: mask analysis (+ zeroing of grad), per-image
 * 8-byte lookup tables in context scratch, fused list kernel.  Outputs are identical to
 * pld_mask_compact + pld_fused_sample_loss_bwd with the same (seed, offset, image_base).
 * With rankings == NULL holed masks take the valid-index mode (one table gather per draw instead of two; the
 * gradient is accumulated per valid pixel and expanded once), results unchanged.
 *   mask f32[B,Hm,Wm], gt f32[B,H*W], pred f32[B,H*W]
 *   -> n_valid i32[B] (nullable; negative = identity table, see pld_mask_compact),
 *      rankings f32[B,n,K,2] (nullable), loss f32[1], loss_sum f64[1] (nullable),
 *      per_list f32[B*n] (nullable), grad f32[B,H*W] (nullable = forward only; overwritten) */
int pld_fused_step(pld_ctx* ctx, const float* mask, const float* gt, const float* pred, int B, int Hm,
                   int Wm, int H, int W, int K, int n, uint64_t seed, uint64_t offset, int image_base,
                   float scale, int32_t* n_valid, float* rankings, float* loss, double* loss_sum,
                   float* per_list, float* grad, void* stream);

/* The same call for a uint8 mask (nonzero = valid pixel), the form the reference's consistency masks have on disk
 * (HR-WSI valid_masks are 8-bit PNGs, pldepth/data/dao/hr_wsi.py:55-83, cast to float by the tf.data pipeline): a host
 * caller ships a quarter of the mask bytes, everything else is identical to pld_fused_step.
 *   mask u8[B,Hm,Wm] */
int pld_fused_step_m8(pld_ctx* ctx, const uint8_t* mask, const float* gt, const float* pred, int B, int Hm,
                      int Wm, int H, int W, int K, int n, uint64_t seed, uint64_t offset, int image_base,
                      float scale, int32_t* n_valid, float* rankings, float* loss, double* loss_sum,
                      float* per_list, float* grad, void* stream);

/* Same step for the score-based strategies (Masked / Thresholded / InformationScore,
 * sampling.py:157-169, 190-208, 218-239): n = int(R * factor) Philox candidate lists per image are drawn
 * and scored (only the 8-byte ordered score is stored), the best R are ordered by score descending (ties:
 * larger candidate index first) -- one shared-memory sort per image up to 8192 candidates, radix top-R
 * selection + segmented radix sort above that -- and the kept lists are
 * REDRAWN from their Philox list ids inside the fused kernel that emits rankings / loss / gradient.
 * Results equal pld_sample_lists_philox(n) -> pld_score_lists -> pld_select_top(R) -> pld_listmle_fwd_bwd.
 * pred / loss / grad may be NULL (sampler only: rankings and order_out).  ranking_size 1..512: thread-per-list
 * kernels up to 16, above that the group-per-list kernel scores the ordered depths (per-position terms in parallel,
 * NumPy's pairwise / sequential summation order by one lane of the group).
 * With rankings == NULL nobody observes the order of the kept lists (the loss is invariant to it): the sort is
 * replaced by an exact selection of the same SET and order_out lists the kept candidates in unspecified order.
 * Above 8192 candidates per image that selection is the sampled-window one (csrc/pld_pilot.cu: thresholds from the
 * first 8192 candidates, which are an i.i.d. sample; the scoring pass stores only candidates above / inside the
 * window; exact cut inside it; exact fallback over all candidates whenever a window misses).  In deterministic mode
 * (pld_ctx_set_deterministic) the order-preserving radix selection is used instead (ascending candidate order).
 *   -> order_out i32[B,R] (nullable): candidate index of every kept list */
int pld_fused_step_scored(pld_ctx* ctx, const float* mask, const float* gt, const float* pred, int B, int Hm,
                          int Wm, int H, int W, int K, int n, int R, int strategy, double threshold,
                          double equality_penalty, int promotion, uint64_t seed, uint64_t offset,
                          int image_base, float scale, int32_t* n_valid, int32_t* order_out, float* rankings,
                          float* loss, double* loss_sum, float* per_list, float* grad, void* stream);

/* ---- evaluation metrics next to the training step ------------------------------
 * ordinal_error (pldepth/active_learning/metrics.py:60-70): pred, gt f32[N,HW]; idx0, idx1 i32[num] are
 * the fixed pixel pairs (np.random.seed(10); np.random.choice(HW, 2*num, replace=False), split in two);
 * err[i] = 1 - (# pairs with (pred[a] > pred[b]) == (gt[a] > gt[b])) / num.
 * calc_d / nDCG (metrics.py:92-110): ids i32[n] the fixed sample (seed 69), n <= 1024;
 * out[i] = DCG(sorted 1/(minmax(pred)[ids]+1)) / DCG(sorted 1/(gt[ids]+1)). */
int pld_ordinal_error(pld_ctx* ctx, const float* pred, const float* gt, const int32_t* idx0, const int32_t* idx1,
                      int N, int HW, int num, float* err, void* stream);
int pld_ndcg(pld_ctx* ctx, const float* pred, const float* gt, const int32_t* ids, int N, int HW, int n, float* out,
             void* stream);

/* ---- evaluation list generators, bit-compatible with the NumPy global stream ---------------
 * GenericHourglassPairRelationDataProvider.generate_ordinal_pairs (pldepth/data/providers/generic_ranking_provider.py:
 * 80-111): per image n_pairs pairs; each pair draws x0 = randint(H), y0 = randint(W), x1 = randint(H), y1 = randint(W)
 * (masked rejection on raw MT19937 words with ALTERNATING bounds, consumed from word *consumed_io on, which is
 * advanced); row = (x0*W + y0, x1*W + y1, get_depth_relation(z0, z1, threshold) [negated if invert_sign], z0, z1) as
 * float32.  threshold < 0 = the reference's threshold=None (plain comparison).  Raises PLD_ST_MT_EXHAUSTED when the
 * stream is too short.
 *   gt f32[N,H*W], raw u32[n_raw], consumed_io i64[1] (device, in/out) -> pairs_out f32[N,n_pairs,5] */
int pld_eval_ordinal_pairs_mt(pld_ctx* ctx, const float* gt, int N, int H, int W, int n_pairs, double threshold,
                              int invert_sign, int promotion, const uint32_t* raw, int64_t n_raw, int64_t* consumed_io,
                              float* pairs_out, void* stream);
/* GenericHourglassRankingDataProvider.generate_rankings (generic_ranking_provider.py:180-215) draws and orders its
 * lists exactly like the core sampler on a full mask (pld_sample_lists_mt with the identity table); with
 * invert_relation_sign it stores them by original depth ASCENDING with depth -> 1 / (depth + 1): this call applies
 * that transformation in place to depth-descending lists.
 *   rankings f32[n_lists,K,2] in/out */
int pld_eval_invert_rankings(pld_ctx* ctx, float* rankings, int64_t n_lists, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLDEPTH_B200_H */
