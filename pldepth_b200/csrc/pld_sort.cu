// Hand-written segmented LSD radix sort (one segment per image) and radix top-R selection for the
// candidate scores of the score-based sampling strategies (sampling.py:169, 208, 239:
// `result[np.argsort(scores)[::-1]][:batch_size]`).
//
// Layout: every image owns a fixed-stride slice of the key / value buffers; its live length is
// either a host constant or read from a device array (survivors of the selection).  Keys are the
// order-preserving u64 image of the float64 scores, values the candidate index inside the image.
// The sort is ascending and stable; the consumer reads the tail of each segment backwards, which
// yields "score descending, ties: larger candidate index first" (= reversed stable argsort).
#include "pld_common.cuh"

namespace pld {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // elements per CTA
constexpr int RS_BINS = 256;

struct SegInfo {
  const int* len_dev;  // per-image live length (nullable)
  int len_fixed;       // used when len_dev == nullptr
  size_t stride;       // elements between consecutive images
  __device__ __forceinline__ int len(int b) const { return len_dev ? len_dev[b] : len_fixed; }
};

// per (image, tile) digit histogram -> hist[(b * RS_BINS + digit) * nblk + tile]
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t* __restrict__ keys, SegInfo seg, int shift,
                                                             int nblk, int* __restrict__ hist) {
  __shared__ int s_h[RS_BINS];
  const int b = blockIdx.y, tile = blockIdx.x;
  const int n = seg.len(b);
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const int base = tile * RS_TILE;
  if (base < n) {
    const uint64_t* k = keys + (size_t)b * seg.stride;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
      const int idx = base + i * RS_THREADS + threadIdx.x;
      if (idx < n) atomicAdd(&s_h[(int)((k[idx] >> shift) & 0xFF)], 1);
    }
  }
  __syncthreads();
  hist[((size_t)b * RS_BINS + threadIdx.x) * nblk + tile] = s_h[threadIdx.x];
}

// one CTA per image: exclusive scan of its RS_BINS * nblk counts in (digit, tile) order
__global__ void __launch_bounds__(1024) rs_scan_kernel(int* __restrict__ hist, int nblk) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  int* h = hist + (size_t)blockIdx.x * RS_BINS * nblk;
  const int total = RS_BINS * nblk;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < total; base += 1024) {
    const int i = base + threadIdx.x;
    const int c = (i < total) ? h[i] : 0;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += s_warp[w];
    const int carry = s_carry;
    if (i < total) h[i] = carry + woff + inc - c;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + woff + inc;
    __syncthreads();
  }
}

// stable scatter: elements of a tile are ranked round by round (256 consecutive elements per round);
// inside a round __match_any_sync gives the rank among equal digits of a warp, per-warp digit counts
// give the rank across warps, running counters carry over rounds.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                const uint32_t* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out, SegInfo seg, int shift,
                                                                int nblk, const int* __restrict__ hist) {
  __shared__ int s_base[RS_BINS];              // global offset of (digit, this tile) + elements already placed
  __shared__ int s_wcnt[RS_THREADS / 32][RS_BINS];
  const int b = blockIdx.y, tile = blockIdx.x;
  const int n = seg.len(b);
  const int base = tile * RS_TILE;
  if (base >= n) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  s_base[threadIdx.x] = hist[((size_t)b * RS_BINS + threadIdx.x) * nblk + tile];
#pragma unroll
  for (int w = 0; w < RS_THREADS / 32; ++w) s_wcnt[w][threadIdx.x] = 0;
  __syncthreads();
  const size_t off = (size_t)b * seg.stride;
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int idx = base + r * RS_THREADS + threadIdx.x;
    const bool on = idx < n;
    uint64_t k = 0;
    uint32_t v = 0;
    int d = 0;
    if (on) {
      k = keys_in[off + idx];
      v = vals_in[off + idx];
      d = (int)((k >> shift) & 0xFF);
    }
    // lanes past the end use digit 256 + lane so that they match nobody
    const unsigned m = __match_any_sync(0xffffffffu, on ? d : (256 + lane));
    const int rank_w = __popc(m & ((1u << lane) - 1u));
    if (on && rank_w == 0) s_wcnt[wid][d] = __popc(m);
    __syncthreads();
    if (on) {
      int pre = 0;
      for (int w = 0; w < wid; ++w) pre += s_wcnt[w][d];
      const int pos = s_base[d] + pre + rank_w;
      keys_out[off + pos] = k;
      vals_out[off + pos] = v;
    }
    __syncthreads();
    {
      int t = 0;
#pragma unroll
      for (int w = 0; w < RS_THREADS / 32; ++w) {
        t += s_wcnt[w][threadIdx.x];
        s_wcnt[w][threadIdx.x] = 0;
      }
      s_base[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// Sorts (keys, vals) ascending, stable, per image.  Eight 8-bit passes ping-pong between the two
// buffer pairs; the result ends in (keys, vals).  hist: int[B * 256 * nblk] scratch.
int seg_radix_sort(pld_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                   const int* len_dev, int len_max, size_t stride, int B, int* hist, cudaStream_t st) {
  if (len_max <= 0) return PLD_OK;
  const int nblk = (len_max + RS_TILE - 1) / RS_TILE;
  SegInfo seg{len_dev, len_max, stride};
  dim3 grid((unsigned)nblk, (unsigned)B);
  uint64_t *ki = keys, *ko = keys_tmp;
  uint32_t *vi = vals, *vo = vals_tmp;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = pass * 8;
    rs_hist_kernel<<<grid, RS_THREADS, 0, st>>>(ki, seg, shift, nblk, hist);
    PLD_CHECK_LAUNCH();
    rs_scan_kernel<<<B, 1024, 0, st>>>(hist, nblk);
    PLD_CHECK_LAUNCH();
    rs_scatter_kernel<<<grid, RS_THREADS, 0, st>>>(ki, vi, ko, vo, seg, shift, nblk, hist);
    PLD_CHECK_LAUNCH();
    uint64_t* tk = ki; ki = ko; ko = tk;
    uint32_t* tv = vi; vi = vo; vo = tv;
  }
  return PLD_OK;
}

size_t seg_radix_sort_hist_bytes(int len_max, int B) {
  const int nblk = (len_max + RS_TILE - 1) / RS_TILE;
  return sizeof(int) * (size_t)B * RS_BINS * (size_t)(nblk > 0 ? nblk : 1);
}

}  // namespace pld
