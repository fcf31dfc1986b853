// Top-R selection of the score-based strategies when nobody sees the ORDER of the kept lists (fused training step,
// rankings not materialised): pilot sample -> window -> scoring pass that stores only what can still matter -> exact cut.
// Replaces, on that path, the key array + three MSB histogram passes + compaction + refinement passes of pld_step.cu
// (about 25 launches and 4 full passes over 8 B per candidate) by a handful of small launches and no pass over the
// candidates besides the scoring pass itself.
//
// Candidate lists are i.i.d. (every list has its own Philox counter), so the first S candidates of an image are a
// uniform sample of its n candidates:
//   1. pilot_score_kernel    scores candidates 0..S-1 of every image (S = 8192) into a small key array
//      pilot_rank_kernel     one CTA per image: brackets the sample keys of rank mu -+ z sigma (mu = S R / n,
//                            sigma^2 = S p (1 - p)) by two 11-bit MSB radix passes below the common prefix of the
//                            sample = thresholds t_hi >= t_lo (bucket edges, so never tighter than the exact ranks)
//   2. score_select_kernel   scores all n candidates (software-pipelined gathers); key > t_hi: kept for sure;
//                            t_lo <= key <= t_hi: band; below t_lo: dropped.  Every CTA appends to its OWN segment of
//                            the sure / band buffers through shared-memory counters (a single global counter per
//                            image was measured first: the returning atomics serialise with the gathers, 110 -> 170 us)
//                            and histograms its band entries over 2048 linear bins of [t_lo, t_hi]
//   3. gather_kernel         one CTA per segment: the bin holding the (R - #sure)-th best band entry is read off the
//                            histogram; sure entries and band entries in higher bins go to the output (one global atomic
//                            per CTA), entries of the boundary bin -- a few dozen per image -- to a small list
//      boundary_kernel       one CTA per image: exact cut inside the boundary bin by (key desc, id desc) (MSB radix
//                            selection); heavy ties overflow the small list and are resolved over the band itself
//   4. kernels 2-3 again for images whose window missed (#sure > R or #sure + #band < R: probability
//      < 1e-11 per image for z = 7 by the binomial tail, but always possible): window (0, +inf), i.e. every
//      candidate goes through the exact selection.  Normally these launches exit at once.
// The kept SET is exactly that of np.argsort(scores)[::-1][:R] with the stable tie rule (larger candidate id first,
// DESIGN.md "Oracle"); its order in `order` is unspecified.
//
// Reference semantics: sample_masked_point_batch of the score-based strategies (pldepth/data/sampling.py:157-169,
// 190-208, 218-239).
#include <math.h>
#include <stdlib.h>

#include "pld_lists.cuh"

namespace pld {

constexpr int RS_BITS = 11;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int BD_CAP = 4096;       // boundary-bin entries kept in the small list (per image)

struct PilotParams {
  uint64_t* pilot_keys; // [B, pilot_stride] sample keys: the first S of every image
  size_t pilot_stride;  // S_pad for the pilot's own array; n when the sample is the head of a stored key array
  uint64_t* t_hi;       // [B] keys above are kept for sure
  uint64_t* t_lo;       // [B] keys below are dropped
  int* flags;           // [B] 1 = window missed, image is redone with the trivial window
  int* tot;             // [B, 4] #sure, #band, #written to the output so far, #boundary entries
  unsigned int* hist;   // [B, RS_BINS] band entries per linear bin of [t_lo, t_hi]
  // Every WARP of the scoring pass appends to its own sub-segment (8 per CTA, sub-segment s = cta * 8 + warp at
  // s * sub_cap): its counters are warp-uniform registers -- no atomic, no broadcast on the critical path (the
  // shared-memory counters of the first version were 14 % of the pass's stall samples).
  int* cnt_sure;        // [B, nsub] entries in every sub-segment
  int* cnt_band;        // [B, nsub]
  uint32_t* sure_v;     // [B, nsub * sub_cap] candidate ids kept for sure
  uint64_t* band_k;     // [B, nsub * sub_cap]
  uint32_t* band_v;     // [B, nsub * sub_cap]
  uint64_t* bd_k;       // [B, BD_CAP] entries of the boundary bin
  uint32_t* bd_v;       // [B, BD_CAP]
  uint32_t* order;      // [B, R] kept candidate ids
  int32_t* order_out;   // [B, R] nullable copy for the caller
  int R, S, S_pad, i_hi, i_lo, low_bits_zero, nseg, seg_cap, nsub, sub_cap;
};

// linear bin of a band key: (key - t_lo) >> bin_shift, bin_shift chosen so that t_hi lands below RS_BINS
__device__ __forceinline__ int bin_shift(uint64_t t_hi, uint64_t t_lo) {
  const uint64_t width = t_hi - t_lo;
  const int bits = width == 0ull ? 0 : 64 - __clzll((long long)width);
  return bits > RS_BITS ? bits - RS_BITS : 0;
}

// ordered score key of one candidate from its K depths (any order); identical to the scoring branch of
// lists_small_kernel, so both selections see the same keys
template <int K>
__device__ __forceinline__ uint64_t candidate_key(float (&gs)[K], const ScoreCfg& C, int b, const float* lad_f,
                                                  const double* lad) {
  sort_desc_floats<K>(gs);
  const double sc = (C.promotion == PLD_PROMOTION_NEP50) ? score_regs<float, K>(gs, C, b, lad_f)
                                                         : score_regs<double, K>(gs, C, b, lad);
  const bool f32_exact = C.promotion == PLD_PROMOTION_NEP50 && C.strategy != PLD_STRATEGY_INFORMATION;
  return f32_exact ? score_key_f32((float)sc) : score_key(sc);
}

// the image's ladder of expected depths in shared memory (float or double by promotion); needs a barrier afterwards
template <int K>
__device__ __forceinline__ void ladder_to_shared(const ScoreCfg& C, int b, double* s_lad) {
  if (C.strategy == PLD_STRATEGY_INFORMATION && threadIdx.x < K) {
    if (C.promotion == PLD_PROMOTION_NEP50) fill_ladder<float>(C, b, K, (int)threadIdx.x, reinterpret_cast<float*>(s_lad));
    else fill_ladder<double>(C, b, K, (int)threadIdx.x, s_lad);
  }
}

struct ImageDraw {
  uint32_t M, thresh;
  const float* depth;   // depth of table entry j at depth[2 * j]
};
__device__ __forceinline__ bool image_draw(const ListParams& P, int b, ImageDraw& D) {
  const int mraw = P.n_valid[b];
  const int m = mraw < 0 ? -mraw : mraw;
  if (m == 0) return false;
  D.M = (uint32_t)m;
  D.thresh = (0u - D.M) % D.M;
  // identity table: (gt, pred); holed mask: (bits of the pixel, gt)   (prep_build_kernel, pld_step.cu)
  D.depth = reinterpret_cast<const float*>(P.table + (size_t)b * P.table_stride) + (mraw < 0 ? 0 : 1);
  return true;
}

template <int K>
__device__ __forceinline__ void issue_depths(const ListParams& P, const ImageDraw& D, uint32_t off_lo, uint32_t off_hi16,
                                             int b, int l, float (&g)[K]) {
  int sel[K];
  draw_philox<K, true>(P, off_lo, off_hi16, b, l, D.M, D.thresh, sel);
#pragma unroll
  for (int k = 0; k < K; ++k) g[k] = __ldg(D.depth + 2 * (size_t)sel[k]);
}

// CTA-wide (1024 threads): the bin holding the rem-th largest element of s_hist (bins scanned from the top); returns
// bin, the number of elements in higher bins and the bin's own count through s_res
__device__ __forceinline__ void find_bin_desc(unsigned int* s_hist, unsigned int* s_wsum, int* s_res, unsigned int rem) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int hi = RS_BINS - 1 - 2 * tid;            // thread 0 owns the two top bins
  const unsigned int h0 = s_hist[hi], h1 = s_hist[hi - 1];
  const unsigned int tot = h0 + h1;
  unsigned int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_wsum[wid] = incl;
  __syncthreads();
  for (int w = 0; w < wid; ++w) incl += s_wsum[w];
  const unsigned int excl = incl - tot;
  if (excl < rem && rem <= incl) {
    if (excl + h0 >= rem) { s_res[0] = hi; s_res[1] = (int)excl; s_res[2] = (int)h0; }
    else { s_res[0] = hi - 1; s_res[1] = (int)(excl + h0); s_res[2] = (int)h1; }
  }
  __syncthreads();
}

// ---- 1. pilot -------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) pilot_score_kernel(const ListParams P, const PilotParams Q) {
  __shared__ double s_lad[16];
  pdl_sync();
  const int b = blockIdx.y;
  const int l = blockIdx.x * 256 + threadIdx.x;
  ladder_to_shared<K>(P.score_cfg, b, s_lad);
  __syncthreads();
  if (l >= Q.S_pad) return;
  ImageDraw D;
  uint64_t key = 0ull;
  if (l < Q.S && image_draw(P, b, D)) {
    uint32_t off_lo, off_hi16;
    launch_offset(P, off_lo, off_hi16);
    float g[K];
    issue_depths<K>(P, D, off_lo, off_hi16, b, l, g);
    key = candidate_key<K>(g, P.score_cfg, b, reinterpret_cast<const float*>(s_lad), s_lad);
  }
  Q.pilot_keys[(size_t)b * Q.pilot_stride + l] = key;
}

// Bucket of descending rank r (0-based) among s_keys[0, N) after `passes` 11-bit digits below bit `top` (all keys agree
// above it): returns the decided prefix and its mask; every thread gets the same values
__device__ void bracket_rank_desc(const uint64_t* s_keys, int N, int r, int top, int passes, uint64_t common,
                                  unsigned int* s_hist, unsigned int* s_wsum, int* s_res, uint64_t& pre_val,
                                  uint64_t& pre_mask) {
  pre_mask = top >= 64 ? 0ull : (~0ull << top);
  pre_val = common & pre_mask;
  unsigned int rem = (unsigned int)r + 1u;
  int hi = top;
  for (int p = 0; p < passes && hi > 0; ++p) {
    const int w = hi < RS_BITS ? hi : RS_BITS;
    const int sh = hi - w;
    for (int i = threadIdx.x; i < RS_BINS; i += 1024) s_hist[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += 1024) {
      const uint64_t k = s_keys[i];
      if ((k & pre_mask) == pre_val) atomicAdd(&s_hist[(unsigned int)(k >> sh) & ((1u << w) - 1u)], 1u);
    }
    __syncthreads();
    find_bin_desc(s_hist, s_wsum, s_res, rem);
    const unsigned int bin = (unsigned int)s_res[0], above = (unsigned int)s_res[1];
    __syncthreads();
    rem -= above;
    pre_mask |= ((uint64_t)((1u << w) - 1u)) << sh;
    pre_val |= (uint64_t)bin << sh;
    hi = sh;
  }
}

__global__ void __launch_bounds__(1024) pilot_rank_kernel(const PilotParams Q, const int32_t* __restrict__ n_valid) {
  extern __shared__ __align__(16) unsigned char pilot_smem[];
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(pilot_smem);
  __shared__ unsigned int s_hist[RS_BINS];
  __shared__ unsigned int s_wsum[32];
  __shared__ int s_res[3];
  __shared__ unsigned long long s_or, s_and;
  pdl_sync();
  const int b = blockIdx.x, tid = threadIdx.x;
  const uint64_t* __restrict__ pk = Q.pilot_keys + (size_t)b * Q.pilot_stride;
  if (tid == 0) { s_or = 0ull; s_and = ~0ull; }
  __syncthreads();
  unsigned long long vo = 0ull, va = ~0ull;
  for (int i = tid; i < Q.S_pad; i += 1024) {
    const uint64_t k = pk[i];
    s_keys[i] = k;
    if (i < Q.S) { vo |= k; va &= k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vo |= __shfl_xor_sync(0xffffffffu, vo, o);
    va &= __shfl_xor_sync(0xffffffffu, va, o);
  }
  if ((tid & 31) == 0) { atomicOr(&s_or, vo); atomicAnd(&s_and, va); }
  __syncthreads();
  const uint64_t varying = s_or ^ s_and;                    // bits in which the sample keys differ
  const int top = varying == 0ull ? 0 : 64 - __clzll((long long)varying);
  const uint64_t common = s_and;
  const bool ok = n_valid[b] != 0;
  uint64_t hi_key = ~0ull, lo_key = 0ull;
  if (ok && Q.i_hi >= 0) {
    uint64_t v, m;
    bracket_rank_desc(s_keys, Q.S, Q.i_hi, top, 2, common, s_hist, s_wsum, s_res, v, m);
    hi_key = v | ~m;          // upper edge of the bucket
  }
  if (ok && Q.i_lo < Q.S) {
    uint64_t v, m;
    bracket_rank_desc(s_keys, Q.S, Q.i_lo, top, 2, common, s_hist, s_wsum, s_res, v, m);
    lo_key = v;               // lower edge of the bucket
  }
  if (tid == 0) {
    Q.t_hi[b] = hi_key;
    Q.t_lo[b] = lo_key;
    Q.flags[b] = 0;
  }
  if (tid < 4) Q.tot[b * 4 + tid] = 0;
  for (int i = tid; i < RS_BINS; i += 1024) Q.hist[(size_t)b * RS_BINS + i] = 0u;
}

// ---- 2. scoring pass with the window test ------------------------------------------------------------------------------
#ifndef PLD_SCORESEL_MINBLOCKS
#define PLD_SCORESEL_MINBLOCKS 3
#endif
#ifndef PLD_SCORESEL_LPT
#define PLD_SCORESEL_LPT 1
#endif
template <int K>
__global__ void __launch_bounds__(256, (K <= 5) ? 4 : ((K <= 8) ? PLD_SCORESEL_MINBLOCKS : 2)) score_select_kernel(const ListParams P,
                                                                                                 const PilotParams Q,
                                                                                                 int only_flagged) {
  __shared__ unsigned int s_hist[RS_BINS];
  __shared__ double s_lad[16];
  __shared__ int s_tot[2];
  pdl_sync();
  const int b = blockIdx.y;
  if (only_flagged && Q.flags[b] == 0) return;
  ladder_to_shared<K>(P.score_cfg, b, s_lad);
  for (int i = threadIdx.x; i < RS_BINS; i += 256) s_hist[i] = 0u;
  if (threadIdx.x < 2) s_tot[threadIdx.x] = 0;
  __syncthreads();
  // short lists keep the float32 ladder in registers: the shared-memory read per term showed up as 10 % of the stall
  // samples (short scoreboard) of the information pass
  float lad_r[K <= 8 ? K : 1];
  if (K <= 8) {
#pragma unroll
    for (int k = 0; k < (K <= 8 ? K : 1); ++k) lad_r[k] = reinterpret_cast<const float*>(s_lad)[k];
  }
  const float* lad_f = K <= 8 ? lad_r : reinterpret_cast<const float*>(s_lad);
  ImageDraw D;
  const bool ok = image_draw(P, b, D);      // empty mask: nothing is appended; the redraw pass raises PLD_ST_EMPTY_MASK
  uint32_t off_lo, off_hi16;
  launch_offset(P, off_lo, off_hi16);
  const uint64_t t_hi = Q.t_hi[b], t_lo = Q.t_lo[b];
  const int bsh = bin_shift(t_hi, t_lo);
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int n = P.n;
  const size_t subseg = (size_t)b * Q.nsub + (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  uint32_t* sv = Q.sure_v + subseg * (size_t)Q.sub_cap;
  uint64_t* bk = Q.band_k + subseg * (size_t)Q.sub_cap;
  uint32_t* bv = Q.band_v + subseg * (size_t)Q.sub_cap;
  int cnt_s = 0, cnt_b = 0;                 // this warp's appends so far (warp-uniform)
  constexpr int LPT = PLD_SCORESEL_LPT;     // lists per thread and iteration: LPT * K gathers in flight per thread
  const int stride = gridDim.x * 256 * LPT;
  float pre[LPT][K];
  int base = blockIdx.x * 256 * LPT;
  if (ok && base < n) {
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
      const int l = base + j * 256 + threadIdx.x;
      issue_depths<K>(P, D, off_lo, off_hi16, b, l < n ? l : n - 1, pre[j]);
    }
  }
  for (; ok && base < n; base += stride) {
    float gs[LPT][K];
#pragma unroll
    for (int j = 0; j < LPT; ++j)
#pragma unroll
      for (int k = 0; k < K; ++k) gs[j][k] = pre[j][k];
    if (base + stride < n) {   // the next lists' draws and gathers go out before these are scored
#pragma unroll
      for (int j = 0; j < LPT; ++j) {
        const int ln = base + stride + j * 256 + threadIdx.x;
        issue_depths<K>(P, D, off_lo, off_hi16, b, ln < n ? ln : n - 1, pre[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
      const int l = base + j * 256 + threadIdx.x;
      const bool active = l < n;
      const uint64_t key = candidate_key<K>(gs[j], P.score_cfg, b, lad_f, s_lad);
      const bool sure = active && key > t_hi;
      const bool band = active && !sure && key >= t_lo;
      const unsigned ms = __ballot_sync(0xffffffffu, sure), mb = __ballot_sync(0xffffffffu, band);
      if (sure) sv[cnt_s + __popc(ms & lt)] = (uint32_t)l;
      if (band) {
        const int pos = cnt_b + __popc(mb & lt);
        bk[pos] = key;
        bv[pos] = (uint32_t)l;
        atomicAdd(&s_hist[(unsigned int)((key - t_lo) >> bsh)], 1u);
      }
      cnt_s += __popc(ms);
      cnt_b += __popc(mb);
    }
  }
  if (lane == 0) {
    Q.cnt_sure[subseg] = cnt_s;
    Q.cnt_band[subseg] = cnt_b;
    if (cnt_s) atomicAdd(&s_tot[0], cnt_s);
    if (cnt_b) atomicAdd(&s_tot[1], cnt_b);
  }
  const bool any_band = __syncthreads_or(cnt_b != 0);
  if (threadIdx.x == 0) {   // one global atomic per CTA and counter: hundreds of warps on two addresses would queue up
    if (s_tot[0]) atomicAdd(Q.tot + b * 4 + 0, s_tot[0]);
    if (s_tot[1]) atomicAdd(Q.tot + b * 4 + 1, s_tot[1]);
  }
  if (any_band) {
    unsigned int* h = Q.hist + (size_t)b * RS_BINS;
    for (int i = threadIdx.x; i < RS_BINS; i += 256)
      if (s_hist[i]) atomicAdd(h + i, s_hist[i]);
  }
}

// ---- 2'. the same window test over STORED keys (ranking_size > 16: the scoring pass of the long-list kernels leaves a key
// array, pld_score_reg.cu / pld_lists_tab.cu); same segments, counters and histogram as score_select_kernel
__global__ void __launch_bounds__(256) classify_keys_kernel(const uint64_t* __restrict__ keys, int n, const PilotParams Q,
                                                            const int32_t* __restrict__ n_valid, int only_flagged) {
  __shared__ unsigned int s_hist[RS_BINS];
  __shared__ int s_tot[2];
  pdl_sync();
  const int b = blockIdx.y;
  if (only_flagged && Q.flags[b] == 0) return;
  if (n_valid[b] == 0) return;
  for (int i = threadIdx.x; i < RS_BINS; i += 256) s_hist[i] = 0u;
  if (threadIdx.x < 2) s_tot[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t t_hi = Q.t_hi[b], t_lo = Q.t_lo[b];
  const int bsh = bin_shift(t_hi, t_lo);
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const size_t subseg = (size_t)b * Q.nsub + (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  uint32_t* sv = Q.sure_v + subseg * (size_t)Q.sub_cap;
  uint64_t* bk = Q.band_k + subseg * (size_t)Q.sub_cap;
  uint32_t* bv = Q.band_v + subseg * (size_t)Q.sub_cap;
  int cnt_s = 0, cnt_b = 0;
  const uint64_t* __restrict__ kb = keys + (size_t)b * (size_t)n;
  for (int base = blockIdx.x * 256; base < n; base += gridDim.x * 256) {
    const int l = base + threadIdx.x;
    const bool active = l < n;
    const uint64_t key = active ? kb[l] : 0ull;
    const bool sure = active && key > t_hi;
    const bool band = active && !sure && key >= t_lo;
    const unsigned ms = __ballot_sync(0xffffffffu, sure), mb = __ballot_sync(0xffffffffu, band);
    if (sure) sv[cnt_s + __popc(ms & lt)] = (uint32_t)l;
    if (band) {
      const int pos = cnt_b + __popc(mb & lt);
      bk[pos] = key;
      bv[pos] = (uint32_t)l;
      atomicAdd(&s_hist[(unsigned int)((key - t_lo) >> bsh)], 1u);
    }
    cnt_s += __popc(ms);
    cnt_b += __popc(mb);
  }
  if (lane == 0) {
    Q.cnt_sure[subseg] = cnt_s;
    Q.cnt_band[subseg] = cnt_b;
    if (cnt_s) atomicAdd(&s_tot[0], cnt_s);
    if (cnt_b) atomicAdd(&s_tot[1], cnt_b);
  }
  const bool any_band = __syncthreads_or(cnt_b != 0);
  if (threadIdx.x == 0) {   // one global atomic per CTA and counter: hundreds of warps on two addresses would queue up
    if (s_tot[0]) atomicAdd(Q.tot + b * 4 + 0, s_tot[0]);
    if (s_tot[1]) atomicAdd(Q.tot + b * 4 + 1, s_tot[1]);
  }
  if (any_band) {
    unsigned int* h = Q.hist + (size_t)b * RS_BINS;
    for (int i = threadIdx.x; i < RS_BINS; i += 256)
      if (s_hist[i]) atomicAdd(h + i, s_hist[i]);
  }
}

// ---- 3a. sure entries + band entries above the boundary bin -> output; boundary bin -> small list --------------------------
// grid (nseg, B), 256 threads.  The boundary bin t* is the largest bin with  #(band entries in bins >= t*) >= need.
__global__ void __launch_bounds__(256) gather_kernel(const PilotParams Q, const int32_t* __restrict__ n_valid,
                                                     int only_flagged) {
  __shared__ unsigned int s_part[256];
  __shared__ int s_tstar, s_base[2], s_k[2];
  pdl_sync();
  const int b = blockIdx.y, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  if (only_flagged && Q.flags[b] == 0) return;
  if (n_valid[b] == 0) return;
  const int R = Q.R;
  const int ns = Q.tot[b * 4 + 0], nb = Q.tot[b * 4 + 1];
  if (ns > R || ns + nb < R) return;          // missed window: boundary_kernel flags the image
  const int need = R - ns;
  // the CTA's eight warp sub-segments, addressed as ONE virtual array: entry v of the concatenation lives at
  // sub-segment u (the last one whose prefix count is <= v), offset v - prefix[u]
  int ps_s[9], ps_b[9];
  ps_s[0] = 0; ps_b[0] = 0;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    ps_s[u + 1] = ps_s[u] + Q.cnt_sure[(size_t)b * Q.nsub + c * 8 + u];
    ps_b[u + 1] = ps_b[u] + Q.cnt_band[(size_t)b * Q.nsub + c * 8 + u];
  }
  const int cnt_s = ps_s[8], cnt_b = ps_b[8];
  const int sub_cap = Q.sub_cap;
  auto vmap = [&](const int (&ps)[9], int v) {
    int base = 0, off = v;
#pragma unroll
    for (int u = 1; u < 8; ++u) {
      const bool ge = v >= ps[u];
      base = ge ? u * sub_cap : base;
      off = ge ? v - ps[u] : off;
    }
    return base + off;
  };
  if (cnt_s == 0 && cnt_b == 0) return;
  // boundary bin: thread t owns bins [8t, 8t + 8), scanned from the top
  int tstar = RS_BINS;                        // need == 0: nothing from the band
  if (need > 0 && cnt_b > 0) {
    const unsigned int* __restrict__ h = Q.hist + (size_t)b * RS_BINS;
    unsigned int loc[8], sum = 0;
    const int hi = RS_BINS - 8 * tid;         // thread 0 owns the eight top bins
#pragma unroll
    for (int i = 0; i < 8; ++i) { loc[i] = h[hi - 1 - i]; sum += loc[i]; }
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) s_tstar = 0;
    unsigned int excl = 0;
    for (int t = 0; t < tid; ++t) excl += s_part[t];
    __syncthreads();
    if (excl < (unsigned int)need && (unsigned int)need <= excl + sum) {
      unsigned int above = excl;
      int i = 0;
      while (i < 7 && above + loc[i] < (unsigned int)need) { above += loc[i]; ++i; }
      s_tstar = hi - 1 - i;
    }
    __syncthreads();
    tstar = s_tstar;
  }
  const uint64_t t_hi = Q.t_hi[b], t_lo = Q.t_lo[b];
  const int bsh = bin_shift(t_hi, t_lo);
  const size_t seg = ((size_t)b * Q.nsub + (size_t)c * 8) * (size_t)sub_cap;
  const uint64_t* __restrict__ bk = Q.band_k + seg;
  const uint32_t* __restrict__ bv = Q.band_v + seg;
  // pass 1: how many of this segment's band entries lie above / inside the boundary bin
  int keep = 0, bd = 0;
  if (need > 0) {
    for (int i = tid; i < cnt_b; i += 256) {
      const int bin = (int)((bk[vmap(ps_b, i)] - t_lo) >> bsh);
      keep += bin > tstar;
      bd += bin == tstar;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    keep += __shfl_xor_sync(0xffffffffu, keep, o);
    bd += __shfl_xor_sync(0xffffffffu, bd, o);
  }
  __syncthreads();
  if (tid == 0) { s_k[0] = 0; s_k[1] = 0; }
  __syncthreads();
  if (lane == 0) { if (keep) atomicAdd(&s_k[0], keep); if (bd) atomicAdd(&s_k[1], bd); }
  __syncthreads();
  if (tid == 0) {
    const int k_all = cnt_s + s_k[0];
    s_base[0] = k_all ? atomicAdd(Q.tot + b * 4 + 2, k_all) : 0;
    s_base[1] = s_k[1] ? atomicAdd(Q.tot + b * 4 + 3, s_k[1]) : 0;
    s_k[0] = 0; s_k[1] = 0;
  }
  __syncthreads();
  uint32_t* order = Q.order + (size_t)b * R;
  int32_t* order_out = Q.order_out != nullptr ? Q.order_out + (size_t)b * R : nullptr;
  const int obase = s_base[0], bbase = s_base[1];
  // sure entries
  const uint32_t* __restrict__ sv = Q.sure_v + seg;
  for (int i = tid; i < cnt_s; i += 256) {
    const int pos = obase + i;
    if (pos < R) {
      const uint32_t v = sv[vmap(ps_s, i)];
      order[pos] = v;
      if (order_out != nullptr) order_out[pos] = (int32_t)v;
    }
  }
  // pass 2 over the band entries
  if (need > 0) {
    const unsigned lt = (1u << lane) - 1u;
    uint64_t* bdk = Q.bd_k + (size_t)b * BD_CAP;
    uint32_t* bdv = Q.bd_v + (size_t)b * BD_CAP;
    for (int i0 = 0; i0 < cnt_b; i0 += 256) {
      const int i = i0 + tid;
      uint64_t k = 0ull;
      uint32_t v = 0u;
      int bin = -1;
      if (i < cnt_b) { const int a = vmap(ps_b, i); k = bk[a]; v = bv[a]; bin = (int)((k - t_lo) >> bsh); }
      const bool kp = bin > tstar, isbd = bin == tstar;
      const unsigned mk = __ballot_sync(0xffffffffu, kp), mbd = __ballot_sync(0xffffffffu, isbd);
      int pk = 0, pb = 0;
      if (lane == 0) { if (mk) pk = atomicAdd(&s_k[0], __popc(mk)); if (mbd) pb = atomicAdd(&s_k[1], __popc(mbd)); }
      pk = __shfl_sync(0xffffffffu, pk, 0);
      pb = __shfl_sync(0xffffffffu, pb, 0);
      if (kp) {
        const int pos = obase + cnt_s + pk + __popc(mk & lt);
        if (pos < R) {
          order[pos] = v;
          if (order_out != nullptr) order_out[pos] = (int32_t)v;
        }
      }
      if (isbd) {
        const int pos = bbase + pb + __popc(mbd & lt);
        if (pos < BD_CAP) { bdk[pos] = k; bdv[pos] = v; }     // beyond: boundary_kernel walks the band instead
      }
    }
  }
}

// ---- 3b. exact cut inside the boundary bin -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) boundary_kernel(const PilotParams Q, const int32_t* __restrict__ n_valid,
                                                        int only_flagged, int* __restrict__ status) {
  __shared__ unsigned int s_hist[RS_BINS];
  __shared__ unsigned int s_wsum[32];
  __shared__ int s_res[3];
  __shared__ int s_count;
  pdl_sync();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (only_flagged && Q.flags[b] == 0) return;
  if (n_valid[b] == 0) return;
  const int R = Q.R, nseg = Q.nsub;          // the band is walked sub-segment by sub-segment
  const int ns = Q.tot[b * 4 + 0], nb = Q.tot[b * 4 + 1], nout = Q.tot[b * 4 + 2], nbd = Q.tot[b * 4 + 3];
  const uint64_t t_hi = Q.t_hi[b], t_lo = Q.t_lo[b];
  __syncthreads();     // everybody has read the image's state before it may be reset
  if (ns > R || ns + nb < R) {
    // the window missed (only possible in the first round: the trivial window has ns = 0, nb = n >= R)
    if (tid == 0) {
      if (only_flagged) atomicOr(status, PLD_ST_INTERNAL);   // cannot happen; never fail silently
      Q.flags[b] = 1; Q.t_hi[b] = ~0ull; Q.t_lo[b] = 0ull;
    }
    if (tid < 4) Q.tot[b * 4 + tid] = 0;
    for (int i = tid; i < RS_BINS; i += 1024) Q.hist[(size_t)b * RS_BINS + i] = 0u;
    return;
  }
  if (tid == 0 && only_flagged) Q.flags[b] = 0;
  const int rem0 = R - nout;            // entries still to take, all from the boundary bin
  if (rem0 == 0) return;
  if (rem0 < 0 || rem0 > nbd) {
    if (tid == 0) atomicOr(status, PLD_ST_INTERNAL);
    return;
  }
  const int bsh = bin_shift(t_hi, t_lo);
  const bool overflow = nbd > BD_CAP;
  // the boundary bin again (only the overflow path needs it, to filter the band)
  int tstar = 0;
  if (overflow) {
    for (int i = tid; i < RS_BINS; i += 1024) s_hist[i] = Q.hist[(size_t)b * RS_BINS + i];
    __syncthreads();
    find_bin_desc(s_hist, s_wsum, s_res, (unsigned int)(R - ns));
    tstar = s_res[0];
    __syncthreads();
  }
  const uint64_t* __restrict__ bdk = Q.bd_k + (size_t)b * BD_CAP;
  const uint32_t* __restrict__ bdv = Q.bd_v + (size_t)b * BD_CAP;
  const size_t seg0 = (size_t)b * nseg * (size_t)Q.sub_cap;
  const uint64_t* __restrict__ bk = Q.band_k + seg0;
  const uint32_t* __restrict__ bv = Q.band_v + seg0;
  const int* __restrict__ cb = Q.cnt_band + (size_t)b * nseg;
  const size_t cap = (size_t)Q.sub_cap;
  // visit every entry of the boundary bin as (k_, v_): the small list, or -- overflow -- the band filtered by bin
#define PLD_FOR_BD(BODY)                                                   \
  if (!overflow) {                                                         \
    for (int i_ = tid; i_ < nbd; i_ += 1024) {                             \
      const uint64_t k_ = bdk[i_];                                         \
      const uint32_t v_ = bdv[i_];                                         \
      BODY                                                                 \
    }                                                                      \
  } else {                                                                 \
    /* a warp walks the eight sub-segments of one scoring CTA as one virtual array (full lanes) */ \
    for (int g_ = wid; g_ < nseg / 8; g_ += 32) {                          \
      int ps_[9];                                                          \
      ps_[0] = 0;                                                          \
      _Pragma("unroll") for (int u_ = 0; u_ < 8; ++u_) ps_[u_ + 1] = ps_[u_] + cb[g_ * 8 + u_]; \
      for (int i_ = lane; i_ < ps_[8]; i_ += 32) {                         \
        int base_ = 0, off_ = i_;                                          \
        _Pragma("unroll") for (int u_ = 1; u_ < 8; ++u_) {                 \
          const bool ge_ = i_ >= ps_[u_];                                  \
          base_ = ge_ ? u_ * (int)cap : base_;                             \
          off_ = ge_ ? i_ - ps_[u_] : off_;                                \
        }                                                                  \
        const size_t a_ = (size_t)g_ * 8 * cap + (size_t)(base_ + off_);   \
        const uint64_t k_ = bk[a_];                                        \
        if ((int)((k_ - t_lo) >> bsh) != tstar) continue;                  \
        const uint32_t v_ = bv[a_];                                        \
        BODY                                                               \
      }                                                                    \
    }                                                                      \
  }
  uint64_t key_cut = 0ull;       // keep (key, id) >= (key_cut, idx_cut) lexicographically
  uint32_t idx_cut = 0u;
  if (rem0 < nbd) {
    // offsets d = key - t_lo of the bin agree above bit bsh: MSB radix selection on the bits below
    uint64_t pre_mask = 0ull, pre_val = 0ull;
    const int lowest = Q.low_bits_zero ? 32 : 0;
    unsigned int rem = (unsigned int)rem0;
    bool done = false;
    int hi = bsh > lowest ? bsh : lowest;
    while (hi > lowest && !done) {
      const int w = (hi - lowest) < RS_BITS ? (hi - lowest) : RS_BITS;
      const int sh = hi - w;
      for (int i = tid; i < RS_BINS; i += 1024) s_hist[i] = 0u;
      __syncthreads();
      PLD_FOR_BD({
        const uint64_t d = k_ - t_lo;
        if ((d & pre_mask) == pre_val) atomicAdd(&s_hist[(unsigned int)(d >> sh) & ((1u << w) - 1u)], 1u);
      })
      __syncthreads();
      find_bin_desc(s_hist, s_wsum, s_res, rem);
      const unsigned int bin = (unsigned int)s_res[0], above = (unsigned int)s_res[1], cnt = (unsigned int)s_res[2];
      __syncthreads();
      rem -= above;
      pre_mask |= ((uint64_t)((1u << w) - 1u)) << sh;
      pre_val |= (uint64_t)bin << sh;
      hi = sh;
      if (cnt == rem) done = true;     // the boundary bucket is taken whole
    }
    // keys of the bin share (k - t_lo) >> bsh, so k >= cut  <=>  (d & low) >= pre_val with low = bits below bsh
    const uint64_t lowmask = bsh >= 64 ? ~0ull : ((1ull << bsh) - 1ull);
    key_cut = pre_val;                 // compared against d & lowmask below
    if (!done) {
      // more candidates share the boundary key than are needed: larger candidate id first
      uint32_t imask = 0u, ival = 0u;
      int ihi = 23;                    // n <= 2^23 candidates per image
      while (ihi > 0 && !done) {
        const int w = ihi < RS_BITS ? ihi : RS_BITS;   // 23 bits = digits of 11, 11 and 1
        const int sh = ihi - w;
        for (int i = tid; i < RS_BINS; i += 1024) s_hist[i] = 0u;
        __syncthreads();
        PLD_FOR_BD({
          if (((k_ - t_lo) & lowmask) == key_cut && (v_ & imask) == ival)
            atomicAdd(&s_hist[(v_ >> sh) & ((1u << w) - 1u)], 1u);
        })
        __syncthreads();
        find_bin_desc(s_hist, s_wsum, s_res, rem);
        const unsigned int bin = (unsigned int)s_res[0], above = (unsigned int)s_res[1], cnt = (unsigned int)s_res[2];
        __syncthreads();
        rem -= above;
        imask |= ((1u << w) - 1u) << sh;
        ival |= bin << sh;
        ihi = sh;
        if (cnt == rem) done = true;
      }
      idx_cut = ival;
    }
    // append the keepers of the boundary bin
    if (tid == 0) s_count = 0;
    __syncthreads();
    uint32_t* order = Q.order + (size_t)b * R + nout;
    int32_t* order_out = Q.order_out != nullptr ? Q.order_out + (size_t)b * R + nout : nullptr;
    PLD_FOR_BD({
      const uint64_t dl = (k_ - t_lo) & lowmask;
      if (dl > key_cut || (dl == key_cut && v_ >= idx_cut)) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < rem0) {
          order[pos] = v_;
          if (order_out != nullptr) order_out[pos] = (int32_t)v_;
        }
      }
    })
    __syncthreads();
    if (tid == 0 && s_count != rem0) atomicOr(status, PLD_ST_INTERNAL);   // never expected
  } else {
    // the whole boundary bin is kept
    if (tid == 0) s_count = 0;
    __syncthreads();
    uint32_t* order = Q.order + (size_t)b * R + nout;
    int32_t* order_out = Q.order_out != nullptr ? Q.order_out + (size_t)b * R + nout : nullptr;
    PLD_FOR_BD({
      (void)k_;
      const int pos = atomicAdd(&s_count, 1);
      if (pos < rem0) {
        order[pos] = v_;
        if (order_out != nullptr) order_out[pos] = (int32_t)v_;
      }
    })
    __syncthreads();
    if (tid == 0 && s_count != rem0) atomicOr(status, PLD_ST_INTERNAL);
  }
#undef PLD_FOR_BD
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static double pilot_z() {
  const char* e = getenv("PLD_PILOT_Z");   // test hook: a tiny z forces the redo path
  if (e != nullptr) {
    const double z = atof(e);
    if (z >= 0.0 && z < 100.0) return z;
  }
  return 7.0;
}

constexpr int PILOT_SAMPLE = 8192;

bool pilot_select_fits(int n) { return n > PILOT_SAMPLE; }

static void pilot_geometry(int B, int n, int num_sms, int* nseg, int* seg_cap) {
  const int per_image_cap = lists_per_image_cap(num_sms, B);
  const int per_cta = 256 * PLD_SCORESEL_LPT;
  int gx = (n + per_cta - 1) / per_cta;
  if (gx > per_image_cap) gx = per_image_cap;
  if (gx < 1) gx = 1;
  const int stride = gx * per_cta;
  *nseg = gx;
  *seg_cap = ((n + stride - 1) / stride) * per_cta;      // lists one CTA of the scoring pass can see
}

constexpr int PILOT_NOFFS = 14;

size_t pilot_select_bytes(int B, int n, int R, int num_sms, size_t* offs) {
  int nseg, seg_cap;
  pilot_geometry(B, n, num_sms, &nseg, &seg_cap);
  const size_t ent = (size_t)B * nseg * (size_t)seg_cap;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += al(bytes); return o; };
  offs[0] = take(sizeof(uint64_t) * (size_t)B * PILOT_SAMPLE);   // pilot keys
  offs[1] = take(sizeof(uint64_t) * B);            // t_hi
  offs[2] = take(sizeof(uint64_t) * B);            // t_lo
  offs[3] = take(sizeof(int) * B);                 // flags
  offs[4] = take(sizeof(int) * 4 * (size_t)B);     // tot
  offs[5] = take(sizeof(unsigned int) * (size_t)B * RS_BINS);   // hist
  offs[6] = take(sizeof(int) * (size_t)B * nseg * 8);  // cnt_sure (per warp sub-segment)
  offs[7] = take(sizeof(int) * (size_t)B * nseg * 8);  // cnt_band
  offs[8] = take(sizeof(uint32_t) * ent);          // sure ids
  offs[9] = take(sizeof(uint64_t) * ent);          // band keys
  offs[10] = take(sizeof(uint32_t) * ent);         // band ids
  offs[11] = take(sizeof(uint64_t) * (size_t)B * BD_CAP);   // boundary keys
  offs[12] = take(sizeof(uint32_t) * (size_t)B * BD_CAP);   // boundary ids
  offs[13] = take(sizeof(uint32_t) * (size_t)B * R);        // order
  return off;
}

template <int K>
static int pilot_select_k(const ListParams& P, const PilotParams& Q, cudaStream_t st) {
  PLD_CUDA(launch_pdl(pilot_score_kernel<K>, dim3((unsigned)((Q.S_pad + 255) / 256), (unsigned)P.B), dim3(256), 0, st, P, Q));
  PLD_CHECK_LAUNCH();
  PLD_CUDA(launch_pdl(pilot_rank_kernel, dim3((unsigned)P.B), dim3(1024), sizeof(uint64_t) * (size_t)Q.S_pad, st, Q,
                      (const int32_t*)P.n_valid));
  PLD_CHECK_LAUNCH();
  const dim3 grid((unsigned)Q.nseg, (unsigned)P.B);
  for (int round = 0; round < 2; ++round) {
    PLD_CUDA(launch_pdl(score_select_kernel<K>, grid, dim3(256), 0, st, P, Q, round));
    PLD_CHECK_LAUNCH();
    PLD_CUDA(launch_pdl(gather_kernel, grid, dim3(256), 0, st, Q, (const int32_t*)P.n_valid, round));
    PLD_CHECK_LAUNCH();
    PLD_CUDA(launch_pdl(boundary_kernel, dim3((unsigned)P.B), dim3(1024), 0, st, Q, (const int32_t*)P.n_valid, round, P.status));
    PLD_CHECK_LAUNCH();
  }
  return PLD_OK;
}

// raises the dynamic shared-memory limit of pilot_rank_kernel on the current device (called by pld_ctx_create)
int pilot_select_init() {
  PLD_CUDA(cudaFuncSetAttribute(pilot_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(uint64_t) * PILOT_SAMPLE)));
  return PLD_OK;
}

static void pilot_params(int B, int n, int R, int low_bits_zero, void* scratch, int32_t* order_out, int num_sms,
                         PilotParams& Q) {
  size_t offs[PILOT_NOFFS];
  pilot_select_bytes(B, n, R, num_sms, offs);
  char* sb = (char*)scratch;
  Q.pilot_keys = (uint64_t*)(sb + offs[0]);
  Q.t_hi = (uint64_t*)(sb + offs[1]); Q.t_lo = (uint64_t*)(sb + offs[2]);
  Q.flags = (int*)(sb + offs[3]); Q.tot = (int*)(sb + offs[4]); Q.hist = (unsigned int*)(sb + offs[5]);
  Q.cnt_sure = (int*)(sb + offs[6]); Q.cnt_band = (int*)(sb + offs[7]);
  Q.sure_v = (uint32_t*)(sb + offs[8]);
  Q.band_k = (uint64_t*)(sb + offs[9]); Q.band_v = (uint32_t*)(sb + offs[10]);
  Q.bd_k = (uint64_t*)(sb + offs[11]); Q.bd_v = (uint32_t*)(sb + offs[12]);
  Q.order = (uint32_t*)(sb + offs[13]); Q.order_out = order_out;
  pilot_geometry(B, n, num_sms, &Q.nseg, &Q.seg_cap);
  Q.nsub = Q.nseg * 8;
  Q.sub_cap = Q.seg_cap / 8;
  Q.R = R;
  Q.S = n < PILOT_SAMPLE ? n : PILOT_SAMPLE;
  int pad = 256;
  while (pad < Q.S) pad <<= 1;
  Q.S_pad = pad;
  Q.pilot_stride = (size_t)pad;
  const double p = (double)R / (double)n, mu = Q.S * p, sigma = sqrt(Q.S * p * (1.0 - p)), z = pilot_z();
  Q.i_hi = (int)floor(mu - z * sigma) - 1;
  Q.i_lo = (int)ceil(mu + z * sigma) + 1;
  Q.low_bits_zero = low_bits_zero;
}

// P: the scoring-pass parameters of pld_fused_step_scored (table, n candidates, score_cfg, Philox stream);
// scratch: pilot_select_bytes() bytes.  Leaves the R kept candidate ids of every image in *order_dev (unordered).
int pilot_select(const ListParams& P, int R, int low_bits_zero, void* scratch, int32_t* order_out, int num_sms,
                 uint32_t** order_dev, cudaStream_t st) {
  PilotParams Q;
  pilot_params(P.B, P.n, R, low_bits_zero, scratch, order_out, num_sms, Q);
  *order_dev = Q.order;
  switch (P.K) {
#define PLD_CASE(KK) case KK: return pilot_select_k<KK>(P, Q, st);
    PLD_CASE(1) PLD_CASE(2) PLD_CASE(3) PLD_CASE(4) PLD_CASE(5) PLD_CASE(6) PLD_CASE(7) PLD_CASE(8)
    PLD_CASE(9) PLD_CASE(10) PLD_CASE(11) PLD_CASE(12) PLD_CASE(13) PLD_CASE(14) PLD_CASE(15) PLD_CASE(16)
#undef PLD_CASE
    default:
      set_error("pilot_select: K=%d out of range", P.K);
      return PLD_EINVAL;
  }
}

// The same selection over a stored key array keys[B, n] (n > PILOT_SAMPLE): the i.i.d. candidates' first 8192 keys are
// the sample, one cheap pass classifies the rest.  scratch: pilot_select_bytes() bytes.
int pilot_select_keys(const uint64_t* keys, int B, int n, const int32_t* n_valid, int* status, int R, int low_bits_zero,
                      void* scratch, int32_t* order_out, int num_sms, uint32_t** order_dev, cudaStream_t st) {
  PilotParams Q;
  pilot_params(B, n, R, low_bits_zero, scratch, order_out, num_sms, Q);
  Q.pilot_keys = const_cast<uint64_t*>(keys);     // read-only here
  Q.pilot_stride = (size_t)n;
  *order_dev = Q.order;
  PLD_CUDA(launch_pdl(pilot_rank_kernel, dim3((unsigned)B), dim3(1024), sizeof(uint64_t) * (size_t)Q.S_pad, st, Q, n_valid));
  PLD_CHECK_LAUNCH();
  const dim3 grid((unsigned)Q.nseg, (unsigned)B);
  for (int round = 0; round < 2; ++round) {
    PLD_CUDA(launch_pdl(classify_keys_kernel, grid, dim3(256), 0, st, keys, n, Q, n_valid, round));
    PLD_CHECK_LAUNCH();
    PLD_CUDA(launch_pdl(gather_kernel, grid, dim3(256), 0, st, Q, n_valid, round));
    PLD_CHECK_LAUNCH();
    PLD_CUDA(launch_pdl(boundary_kernel, dim3((unsigned)B), dim3(1024), 0, st, Q, n_valid, round, status));
    PLD_CHECK_LAUNCH();
  }
  return PLD_OK;
}

}  // namespace pld
