// Hand-written segmented LSD radix sort (one segment per image) and radix top-R selection for the
// candidate scores of the score-based sampling strategies (sampling.py:169, 208, 239:
// `result[np.argsort(scores)[::-1]][:batch_size]`).
//
// Layout: every image owns a fixed-stride slice of the key / value buffers; its live length is
// either a host constant or read from a device array (survivors of the selection).  Keys are the
// order-preserving u64 image of the float64 scores, values the candidate index inside the image.
// The sort is ascending and stable; the consumer reads the tail of each segment backwards, which
// yields "score descending, ties: larger candidate index first" (= reversed stable argsort).
#include <cstdlib>
#include "pld_common.cuh"
#include "pld_score.cuh"

namespace pld {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 elements per CTA; warp w owns elements [256 w, 256 (w + 1))
constexpr int RS_BINS = 256;
static_assert(RS_THREADS == RS_BINS, "one thread per digit in the scatter and scan kernels");

struct SegInfo {
  const int* len_dev;                 // per-image live length (nullable)
  int len_fixed;                      // used when len_dev == nullptr
  size_t stride;                      // elements between consecutive images
  const unsigned long long* varying;  // per-image mask of key bits that differ between live keys (nullable)
  int first_pass;                     // passes below this one are known to be constant for every image
  __device__ __forceinline__ int len(int b) const { return len_dev ? len_dev[b] : len_fixed; }
  // a pass over a byte in which all live keys agree cannot reorder anything: it is skipped, and the
  // ping-pong parity of an image is the number of passes it really executed
  __device__ __forceinline__ bool skip(int b, int pass) const {
    return varying != nullptr && ((varying[b] >> (8 * pass)) & 0xFFull) == 0ull;
  }
  __device__ __forceinline__ int parity(int b, int pass) const {  // executed passes before `pass`
    if (varying == nullptr) return (pass - first_pass) & 1;
    const unsigned long long v = varying[b];
    int c = 0;
    for (int p = first_pass; p < pass; ++p) c += ((v >> (8 * p)) & 0xFFull) ? 1 : 0;
    return c & 1;
  }
};

// Digit counting of one warp's 32 elements: lanes holding the same digit elect a leader that adds the group size to
// the warp's counter, so a byte with two or three distinct values (an exponent byte) costs no more than a uniform
// one.  Returns the number of equal digits in lower lanes (rank inside the group).
__device__ __forceinline__ int warp_digit_rank(bool on, int d, int lane, unsigned& group) {
  group = __match_any_sync(0xffffffffu, on ? d : (RS_BINS + lane));   // idle lanes match nobody
  return __popc(group & ((1u << lane) - 1u));
}

// per (image, tile) digit histogram -> hist[(b * nblk + tile) * RS_BINS + digit]
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t* __restrict__ keys_a,
                                                             const uint64_t* __restrict__ keys_b, SegInfo seg, int pass,
                                                             int nblk, int* __restrict__ hist) {
  pdl_sync();
  __shared__ int s_h[RS_BINS];
  const int b = blockIdx.y, tile = blockIdx.x;
  if (seg.skip(b, pass)) return;
  const int n = seg.len(b);
  const int base = tile * RS_TILE;
  if (base >= n) return;
  if (threadIdx.x < RS_BINS) s_h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t* k = (seg.parity(b, pass) ? keys_b : keys_a) + (size_t)b * seg.stride;
  const int shift = pass * 8;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int idx = base + i * RS_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&s_h[(int)((k[idx] >> shift) & 0xFF)], 1);
  }
  __syncthreads();
  if (threadIdx.x < RS_BINS) hist[((size_t)b * nblk + tile) * RS_BINS + threadIdx.x] = s_h[threadIdx.x];
}

// one CTA per image, thread d owns digit d: running prefix over the live tiles (coalesced across
// digits, eight tiles in flight), then an exclusive scan of the 256 digit totals -> dig_off[b][d]
__global__ void __launch_bounds__(RS_BINS) rs_scan_kernel(int* __restrict__ hist, int* __restrict__ dig_off, SegInfo seg,
                                                          int pass, int nblk) {
  pdl_sync();
  __shared__ int s_tot[RS_BINS];
  const int b = blockIdx.x;
  if (seg.skip(b, pass)) return;
  const int n = seg.len(b);
  const int live = (n + RS_TILE - 1) / RS_TILE;
  int* h = hist + (size_t)b * nblk * RS_BINS + threadIdx.x;
  int run = 0;
  for (int t0 = 0; t0 < live; t0 += 8) {
    int c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = (t0 + j < live) ? h[(size_t)(t0 + j) * RS_BINS] : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (t0 + j < live) h[(size_t)(t0 + j) * RS_BINS] = run;
      run += c[j];
    }
  }
  s_tot[threadIdx.x] = run;
  __syncthreads();
  // exclusive scan of 256 totals (Hillis-Steele in shared memory)
  for (int o = 1; o < RS_BINS; o <<= 1) {
    const int v = (threadIdx.x >= o) ? s_tot[threadIdx.x - o] : 0;
    __syncthreads();
    s_tot[threadIdx.x] += v;
    __syncthreads();
  }
  dig_off[b * RS_BINS + threadIdx.x] = s_tot[threadIdx.x] - run;
}

// Stable scatter.  Every warp owns 256 consecutive elements of the tile (8 rounds of 32, kept in registers) and
// ranks them against its own running digit counters: __match_any_sync gives the rank among equal digits of a round,
// the counter the number of equal digits in the warp's earlier rounds.  The tile is then put in digit order in shared
// memory and written out by consecutive threads, so equal digits leave as contiguous runs: scattered 8-byte stores
// are limited by the SM's request rate (0.66 lane/clk), not by bytes, and one request per run
// instead of one per element is what this buys.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(uint64_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a,
                                                                uint64_t* __restrict__ keys_b, uint32_t* __restrict__ vals_b,
                                                                SegInfo seg, int pass, int nblk,
                                                                const int* __restrict__ hist,
                                                                const int* __restrict__ dig_off) {
  pdl_sync();
  __shared__ int s_cnt[RS_WARPS][RS_BINS];   // running count while ranking, then the warp's offset inside its digit run
  __shared__ int s_start[RS_BINS];           // first position of the digit inside the digit-ordered tile
  __shared__ int s_gbase[RS_BINS];           // global position of the first element of (this tile, digit)
  __shared__ int s_wsum[RS_WARPS];
  __shared__ uint64_t s_key[RS_TILE];
  __shared__ uint32_t s_val[RS_TILE];
  const int b = blockIdx.y, tile = blockIdx.x;
  if (seg.skip(b, pass)) return;
  const int n = seg.len(b);
  const int base = tile * RS_TILE;
  if (base >= n) return;
  const bool flip = seg.parity(b, pass) != 0;
  const size_t off = (size_t)b * seg.stride;
  const uint64_t* keys_in = (flip ? keys_b : keys_a) + off;
  const uint32_t* vals_in = (flip ? vals_b : vals_a) + off;
  uint64_t* keys_out = (flip ? keys_a : keys_b) + off;
  uint32_t* vals_out = (flip ? vals_a : vals_b) + off;
  const int shift = pass * 8;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  const int wbase = base + wid * (RS_ITEMS * 32);
  uint64_t k[RS_ITEMS];
  uint32_t v[RS_ITEMS];
  int rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int idx = wbase + r * 32 + lane;
    k[r] = (idx < n) ? keys_in[idx] : 0ull;
    v[r] = (idx < n) ? vals_in[idx] : 0u;
  }
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const bool on = (wbase + r * 32 + lane) < n;
    const int d = (int)((k[r] >> shift) & 0xFF);
    unsigned group;
    const int rw = warp_digit_rank(on, d, lane, group);
    const int prev = on ? s_cnt[wid][d] : 0;
    __syncwarp();
    if (on && rw == 0) s_cnt[wid][d] = prev + __popc(group);
    __syncwarp();
    rank[r] = prev + rw;
  }
  __syncthreads();
  // digit d = threadIdx.x (RS_THREADS == RS_BINS): offsets of the warps inside the digit run, tile count of the digit
  int cnt = 0;
  {
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const int c = s_cnt[w][threadIdx.x];
      s_cnt[w][threadIdx.x] = cnt;
      cnt += c;
    }
    s_gbase[threadIdx.x] = dig_off[b * RS_BINS + threadIdx.x] + hist[((size_t)b * nblk + tile) * RS_BINS + threadIdx.x];
  }
  // exclusive scan of the 256 digit counts (warp shuffles + warp totals)
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += x;
  }
  if (lane == 31) s_wsum[wid] = incl;
  __syncthreads();
  {
    int before = 0;
    for (int w = 0; w < wid; ++w) before += s_wsum[w];
    s_start[threadIdx.x] = before + incl - cnt;
  }
  __syncthreads();
  // the tile in digit order (stable: warp order, then round order, then lane order inside a digit)
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    if (wbase + r * 32 + lane < n) {
      const int d = (int)((k[r] >> shift) & 0xFF);
      const int lp = s_start[d] + s_cnt[wid][d] + rank[r];
      s_key[lp] = k[r];
      s_val[lp] = v[r];
    }
  }
  __syncthreads();
  const int tile_n = (n - base < RS_TILE) ? (n - base) : RS_TILE;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int lp = i * RS_THREADS + threadIdx.x;
    if (lp < tile_n) {
      const uint64_t key = s_key[lp];
      const int d = (int)((key >> shift) & 0xFF);
      const int pos = s_gbase[d] + (lp - s_start[d]);
      keys_out[pos] = key;
      vals_out[pos] = s_val[lp];
    }
  }
}

// Sorts (keys, vals) ascending, stable, per image, with eight 8-bit passes that ping-pong between the
// (a) and (b) buffer pairs.  Without `varying` the result ends in (a).  With `varying` (per-image mask of
// non-constant key bits) passes over constant bytes are skipped per image and the result of image b sits in
// (a) if it executed an even number of passes, else in (b): see seg_sorted_in_b().
// hist: int[B * nblk * 256 + B * 256] scratch.
int seg_radix_sort(pld_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                   const int* len_dev, int len_max, size_t stride, int B, int* hist,
                   const unsigned long long* varying, int first_pass, cudaStream_t st) {
  if (len_max <= 0) return PLD_OK;
  const int nblk = (len_max + RS_TILE - 1) / RS_TILE;
  SegInfo seg{len_dev, len_max, stride, varying, first_pass};
  int* dig_off = hist + (size_t)B * nblk * RS_BINS;
  dim3 grid((unsigned)nblk, (unsigned)B);
  for (int pass = first_pass; pass < 8; ++pass) {
    PLD_CUDA(launch_pdl(rs_hist_kernel, dim3(grid), dim3(RS_THREADS), 0, st, keys, keys_tmp, seg, pass, nblk, hist));
    PLD_CHECK_LAUNCH();
    PLD_CUDA(launch_pdl(rs_scan_kernel, dim3(B), dim3(RS_BINS), 0, st, hist, dig_off, seg, pass, nblk));
    PLD_CHECK_LAUNCH();
    PLD_CUDA(launch_pdl(rs_scatter_kernel, dim3(grid), dim3(RS_THREADS), 0, st, keys, vals, keys_tmp, vals_tmp, seg, pass, nblk, hist, dig_off));
    PLD_CHECK_LAUNCH();
  }
  return PLD_OK;
}

// ------------------------------------------------------------------------------------------
// Small images of candidates (n <= SS_MAX per image, the sizes the reference itself runs: R = 100 ... 1000 lists per
// image, sampling.py:157,190,218): one CTA per image loads every (key, candidate) pair into shared memory, orders
// them with a bitonic network by (key descending, candidate index descending) -- exactly "reversed stable argsort"
// -- and writes the best R.  One launch instead of selection + eight radix passes.
// ------------------------------------------------------------------------------------------
constexpr int SS_THREADS = 1024;
constexpr int SS_MAX = 8192;

template <bool FROM_SCORES>
__global__ void __launch_bounds__(SS_THREADS) sel_small_kernel(const void* __restrict__ src, int n, size_t stride, int R,
                                                              int N2, int ascending_ids, uint32_t* __restrict__ order,
                                                              int32_t* __restrict__ order_out) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char ss_raw[];
  uint64_t* s_key = reinterpret_cast<uint64_t*>(ss_raw);
  uint32_t* s_id = reinterpret_cast<uint32_t*>(ss_raw + sizeof(uint64_t) * (size_t)N2);
  __shared__ int s_cnt[256];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < N2; i += SS_THREADS) {
    uint64_t k = 0ull;
    if (i < n) {
      if (FROM_SCORES) k = score_key(reinterpret_cast<const double*>(src)[(size_t)b * stride + i]);
      else k = reinterpret_cast<const uint64_t*>(src)[(size_t)b * stride + i];
    }
    s_key[i] = k;
    s_id[i] = (i < n) ? (uint32_t)i : 0u;   // pads: (0, 0) -- not larger than any real pair
  }
  __syncthreads();
  for (int k = 2; k <= N2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (N2 >> 1); t += SS_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const uint64_t ka = s_key[i], kb = s_key[p];
        const uint32_t ia = s_id[i], ib = s_id[p];
        const bool a_lt_b = (ka < kb) || (ka == kb && ia < ib);
        if (a_lt_b == ((i & k) == 0)) {   // descending where bit k of i is clear
          s_key[i] = kb; s_key[p] = ka;
          s_id[i] = ib; s_id[p] = ia;
        }
      }
      __syncthreads();
    }
  }
  if (!ascending_ids) {
    for (int t = threadIdx.x; t < R; t += SS_THREADS) {
      const uint32_t c = s_id[t];
      if (order != nullptr) order[(size_t)b * R + t] = c;
      if (order_out != nullptr) order_out[(size_t)b * R + t] = (int32_t)c;
    }
    return;
  }
  // kept candidates in ascending candidate order (the contract of the unordered selection): bitmap + scan
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_key);   // the keys are no longer needed
  const int nwords = N2 >> 5;   // N2 >= 64; N2 / 32 words always fit in the N2 * 8 bytes of the key array
  __syncthreads();
  if ((int)threadIdx.x < nwords) s_bits[threadIdx.x] = 0u;
  __syncthreads();
  for (int t = threadIdx.x; t < R; t += SS_THREADS) {
    const uint32_t c = s_id[t];
    atomicOr(&s_bits[c >> 5], 1u << (c & 31u));
  }
  __syncthreads();
  if (threadIdx.x < 256) s_cnt[threadIdx.x] = ((int)threadIdx.x < nwords) ? __popc(s_bits[threadIdx.x]) : 0;
  __syncthreads();
  if ((int)threadIdx.x < nwords) {
    int pos = 0;
    for (int w = 0; w < (int)threadIdx.x; ++w) pos += s_cnt[w];
    uint32_t bits = s_bits[threadIdx.x];
    while (bits) {
      const uint32_t c = (threadIdx.x << 5) + (uint32_t)(__ffs((int)bits) - 1);
      bits &= bits - 1u;
      if (order != nullptr) order[(size_t)b * R + pos] = c;
      if (order_out != nullptr) order_out[(size_t)b * R + pos] = (int32_t)c;
      ++pos;
    }
  }
}

bool select_small_fits(int n) { return n >= 1 && n <= SS_MAX; }

// raises the dynamic shared-memory limit of the kernels above on the current device (called by pld_ctx_create)
int select_small_init() {
  const int bytes = SS_MAX * (int)(sizeof(uint64_t) + sizeof(uint32_t));
  PLD_CUDA(cudaFuncSetAttribute(sel_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PLD_CUDA(cudaFuncSetAttribute(sel_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return PLD_OK;
}

// best R of the n candidates of every image; exactly one of keys / scores is given
int select_small(const uint64_t* keys, const double* scores, int n, size_t stride, int B, int R, bool ascending_ids,
                 uint32_t* order, int32_t* order_out, cudaStream_t st) {
  int N2 = 64;
  while (N2 < n) N2 <<= 1;
  const size_t smem = (size_t)N2 * (sizeof(uint64_t) + sizeof(uint32_t));
  if (scores != nullptr)
    PLD_CUDA(launch_pdl(sel_small_kernel<true>, dim3(B), dim3(SS_THREADS), smem, st, scores, n, stride, R, N2, ascending_ids ? 1 : 0, order, order_out));
  else
    PLD_CUDA(launch_pdl(sel_small_kernel<false>, dim3(B), dim3(SS_THREADS), smem, st, keys, n, stride, R, N2, ascending_ids ? 1 : 0, order, order_out));
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

size_t seg_radix_sort_hist_bytes(int len_max, int B) {
  const int nblk = (len_max + RS_TILE - 1) / RS_TILE;
  return sizeof(int) * (size_t)B * RS_BINS * ((size_t)(nblk > 0 ? nblk : 1) + 1);
}

}  // namespace pld
