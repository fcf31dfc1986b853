// Thread-per-list SCORING pass of the score-based strategies for ranking_size 17..64 (float32 / NEP-50 arithmetic):
// Philox draws -> one table gather per draw -> depths ordered in registers -> the candidate's score key.
//
// The group-per-list kernel (pld_lists_tab.cu) spends ~280 lane-instructions per point on moving a list across eight
// lanes; that is hidden behind the reductions of the loss pass, but the scoring pass has no reductions and was
// issue-bound (profiles/r02b_tabscore_summary.txt: 55 % issue slots, L1TEX 46 %, 0.24 sectors per SM and clock).  A score
// depends on the ORDERED DEPTHS only -- no pixel, no prediction, ties interchangeable -- so one thread can hold a whole
// list as 32-bit order-preserving depth images, sort them with Batcher's odd-even merge network (explicit comparator
// list, pld_oem_networks.cuh: two integer min / max per comparator, no shuffle, no shared memory) and evaluate the score
// in NumPy's order of operations from registers.  (The same organisation was measured and rejected for the LOSS pass,
// whose parked payloads, 168 registers and 200 KB of straight-line code left ten latency-bound warps per SM:
// profiles/r02_reg_kernel_rejected_summary.txt.  Here a thread carries 4 bytes per entry and nothing else.)
//
// Arithmetic: identical to score_list / score_regs (pld_score.cuh), i.e. to NumPy's -- sampling.py:161-167 (masked),
// 194-205 (thresholded), 219-237 (information), get_depth_relation depth_utils.py:5-21; keys equal those of the staged
// pld_score_lists and of lists_tab_kernel<.., SCORE> bit for bit (tests: scored step == staged pipeline).
#include "pld_lists.cuh"
#include "pld_oem_networks.cuh"

namespace pld {

constexpr int SREG_THREADS = 64;

#ifndef PLD_SREG_MINBLOCKS
#define PLD_SREG_MINBLOCKS 6
#endif

// the rare redraw path (Lemire rejection, P < M / 2^32 per draw) out of line: inlined it would put 56 copies of Philox
// into the unrolled draw loop
static __device__ __noinline__ uint32_t lemire_redraw(uint32_t word, uint32_t M, uint32_t thresh, DrawStream ds, uint32_t k) {
  return lemire_bounded(word, M, thresh, ds, k);
}

// INFO: information strategy (chi-square against the ladder); else masked / thresholded (sum of adjacent differences)
template <int NP, bool INFO>
__global__ void __launch_bounds__(SREG_THREADS, (NP <= 32) ? 12 : PLD_SREG_MINBLOCKS) score_reg_kernel(const ListParams P) {
  __shared__ float s_lad[64];        // information strategy: the image's ladder of expected depths
  __shared__ uint32_t s_key[(NP + 1) * SREG_THREADS];   // ordered depths, [position][thread]
  pdl_sync();
  constexpr int T = SREG_THREADS;
  const ScoreCfg& C = P.score_cfg;
  const int K = P.K;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  constexpr bool info = INFO;
  const bool thr = C.strategy == PLD_STRATEGY_THRESHOLDED;
  if (info && tid < K) fill_ladder<float>(C, b, K, tid, s_lad);
  __syncthreads();
  uint32_t off_lo, off_hi16;
  launch_offset(P, off_lo, off_hi16);
  const int mraw = P.n_valid[b];
  const int m = mraw < 0 ? -mraw : mraw;
  if (m == 0) return;                // empty mask: the redraw pass raises PLD_ST_EMPTY_MASK
  const uint32_t M = (uint32_t)m, thresh = (0u - M) % M;
  // identity table: (gt, pred); holed mask: (bits of the pixel, gt)   (prep_build_kernel, pld_step.cu)
  const float* __restrict__ depth = reinterpret_cast<const float*>(P.table + (size_t)b * P.table_stride) + (mraw < 0 ? 0 : 1);
  const uint32_t image = (uint32_t)(P.image_base + b);
  const int n8 = K & ~7;

  for (int l = blockIdx.x * T + tid; l < P.n; l += gridDim.x * T) {
    // ---- draws; every gather of the list is issued before the first one is used -----------------------------------
    float g[NP];
    {
      const DrawStream ds{(uint32_t)l, image, off_lo, off_hi16, P.seed_lo, P.seed_hi};
#pragma unroll
      for (int q = 0; q < NP / 4; ++q) {
        if (q * 4 < K) {   // uniform
          const Philox4 r = philox4x32_10_rk((uint32_t)l, image, (uint32_t)q | off_hi16, off_lo, P.rk0, P.rk1);
          const uint32_t w[4] = {r.x, r.y, r.z, r.w};
          uint32_t sel[4];
          bool rej = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) sel[j] = lemire_try(w[j], M, thresh, rej);
          if (rej) {  // rare (P < 4 M / 2^32): the redraw stream (identical for the words that were not rejected)
#pragma unroll
#ifdef PLD_SREG_INLINE
            for (int j = 0; j < 4; ++j) sel[j] = lemire_bounded(w[j], M, thresh, ds, (uint32_t)(q * 4 + j));
#else
            for (int j = 0; j < 4; ++j) sel[j] = lemire_redraw(w[j], M, thresh, ds, (uint32_t)(q * 4 + j));
#endif
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) g[q * 4 + j] = (q * 4 + j < K) ? __ldg(depth + 2 * (size_t)sel[j]) : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) g[q * 4 + j] = 0.f;
        }
      }
    }

    // ---- order (descending): order-preserving images of the depths; pads = 0 sort last ---------------------------------
    uint32_t key[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) key[i] = (i < K) ? float_to_ordered(g[i]) : 0u;
    OemNetwork<NP>::sort_desc(key);

    // ---- score, NumPy's order of operations ----------------------------------------------------------------------------
    // The ordered depths go to the thread's column of shared memory and are scored by ROLLED loops: unrolled over the
    // registers the scoring code was 100 KB of straight-line SASS per list, three times the instruction cache, and a third
    // of the stall samples were `no_instructions` (profiles/r02b_scorereg_summary.txt).
#pragma unroll
    for (int i = 0; i < NP; ++i) s_key[i * T + tid] = key[i];
    const size_t list_id = (size_t)b * (size_t)P.n + (size_t)l;
    float gk = ordered_to_float(s_key[tid]);
    if constexpr (info) {
      // chi_k = (g_k - e_k)^2 / e_k; NumPy's pairwise summation of 8 <= n <= 128 terms: eight strided accumulators
      // r[j] = chi_j + chi_{8+j} + ..., a fixed tree, a sequential tail; "equal" relations counted on the way
      int cnt = 0;
      float r[8];
      auto term = [&](int k) {
        const float gn = ordered_to_float(s_key[(k + 1) * T + tid]);      // row K exists (NP + 1 rows); unused then
        if (k + 1 < K && relation_equal<float>(gk, gn, C)) ++cnt;
        const float e = s_lad[k];
        const float d = __fsub_rn(gk, e);
        gk = gn;
        return __fdiv_rn(__fmul_rn(d, d), e);
      };
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.f;
      for (int k0 = 0; k0 < n8; k0 += 8) {       // one copy of the term code: r[j] = a[j] on the first round, not 0 + a[j]
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = term(k0 + j);
          r[j] = k0 == 0 ? t : __fadd_rn(r[j], t);
        }
      }
      float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                            __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
      for (int k = n8; k < K; ++k) res = __fadd_rn(res, term(k));
      double sc = (double)(-res);
      for (int c = 0; c < cnt; ++c) sc = __dadd_rn(sc, C.penalty);
      P.score_keys[list_id] = score_key(sc);
    } else {
      const float pen = (float)C.penalty;
      float acc = 0.f;
#pragma unroll 4
      for (int k = 0; k + 1 < K; ++k) {
        const float gn = ordered_to_float(s_key[(k + 1) * T + tid]);
        if (thr && relation_equal<float>(gk, gn, C)) acc = __fadd_rn(acc, pen);
        acc = __fadd_rn(acc, fabsf(__fsub_rn(gk, gn)));
        gk = gn;
      }
      P.score_keys[list_id] = score_key_f32(acc);
    }
  }
}

template <int NP>
static int launch_score_reg_cfg(const ListParams& P, int num_sms, cudaStream_t st) {
  // enough CTAs that the tail evens out (64 threads each: eight times the CTAs of the 256-thread kernels)
  const int per_image_cap = lists_per_image_cap(num_sms, P.B) * 8;
  int gx = (P.n + SREG_THREADS - 1) / SREG_THREADS;
  if (gx > per_image_cap) gx = per_image_cap;
  if (gx < 1) gx = 1;
  if (P.score_cfg.strategy == PLD_STRATEGY_INFORMATION)
    PLD_CUDA(launch_pdl(score_reg_kernel<NP, true>, dim3((unsigned)gx, (unsigned)P.B), dim3(SREG_THREADS), 0, st, P));
  else
    PLD_CUDA(launch_pdl(score_reg_kernel<NP, false>, dim3((unsigned)gx, (unsigned)P.B), dim3(SREG_THREADS), 0, st, P));
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

// float32 (NEP-50) scores of lists of 17..64 entries
bool score_reg_fits(const ListParams& P) {
  static const bool off = getenv("PLD_NO_SCORE_REG") != nullptr && getenv("PLD_NO_SCORE_REG")[0] == '1';
  return !off && P.K >= 17 && P.K <= 64 && P.score_cfg.promotion == PLD_PROMOTION_NEP50;
}

int launch_score_reg(const ListParams& P, int num_sms, cudaStream_t st) {
  const int K = P.K;
  if (K <= 24) return launch_score_reg_cfg<24>(P, num_sms, st);
  if (K <= 32) return launch_score_reg_cfg<32>(P, num_sms, st);
  if (K <= 40) return launch_score_reg_cfg<40>(P, num_sms, st);
  if (K <= 48) return launch_score_reg_cfg<48>(P, num_sms, st);
  if (K <= 56) return launch_score_reg_cfg<56>(P, num_sms, st);
  if (K <= 64) return launch_score_reg_cfg<64>(P, num_sms, st);
  set_error("score_reg: K=%d out of range", K);
  return PLD_EINVAL;
}

}  // namespace pld
