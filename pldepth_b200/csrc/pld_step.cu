// One-call training step: valid-pixel analysis -> per-image lookup tables -> fused
// Philox sampling + ordering + ranking emission + gather + ListMLE forward/backward.
//
//   pass 1  prep_count_kernel   counts valid mask pixels per 4096-pixel chunk; zeroes grad
//   pass 2  prep_build_kernel   per image, one 8-byte table entry per (valid) pixel:
//             full mask at image resolution:  table[j] = (gt[j], pred[j])            (n_valid = -HW)
//             otherwise:                      table[j] = (bits(p_j), gt[p_j])        (n_valid = M)
//           so every draw costs ONE divergent 8-byte gather instead of two or three 4-byte ones
//   pass 3  lists_small_kernel<K, SRC_PHILOX_TAB, true>   (pld_lists_small.cu)
//
// The path is bound by divergent L1TEX sector accesses, not by HBM bytes, which
// is why spending 16-20 streamed bytes per pixel to save one random sector per draw pays off
// once an image receives more draws than about two per pixel.
#include <stdlib.h>

#include "pld_lists.cuh"

namespace pld {
int launch_lists_small(const ListParams& P, int src, bool loss, int num_sms, cudaStream_t st);
int launch_lists_large(const ListParams& P, int src, bool loss, int num_sms, cudaStream_t st);
int launch_acc_finalize(pld_ctx* ctx, float* grad, size_t n, float scale, int accumulate, cudaStream_t st);
int launch_offset_advance(pld_ctx* ctx, cudaStream_t st);
int launch_lists_small_score(const ListParams& P, int num_sms, cudaStream_t st);
int launch_lists_tab_score(const ListParams& P, int num_sms, cudaStream_t st);
bool score_reg_fits(const ListParams& P);                                     // pld_score_reg.cu
int launch_score_reg(const ListParams& P, int num_sms, cudaStream_t st);
int seg_radix_sort(pld_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                   const int* len_dev, int len_max, size_t stride, int B, int* hist,
                   const unsigned long long* varying, int first_pass, cudaStream_t st);
size_t seg_radix_sort_hist_bytes(int len_max, int B);
bool select_small_fits(int n);
int select_small(const uint64_t* keys, const double* scores, int n, size_t stride, int B, int R, bool ascending_ids,
                 uint32_t* order, int32_t* order_out, cudaStream_t st);
bool pilot_select_fits(int n);
size_t pilot_select_bytes(int B, int n, int R, int num_sms, size_t* offs);
int pilot_select(const ListParams& P, int R, int low_bits_zero, void* scratch, int32_t* order_out, int num_sms,
                 uint32_t** order_dev, cudaStream_t st);
int pilot_select_keys(const uint64_t* keys, int B, int n, const int32_t* n_valid, int* status, int R, int low_bits_zero,
                      void* scratch, int32_t* order_out, int num_sms, uint32_t** order_dev, cudaStream_t st);

constexpr int PC_THREADS = 256;
constexpr int PC_ITEMS = 16;
constexpr int PC_CHUNK = PC_THREADS * PC_ITEMS;

__device__ __forceinline__ uint32_t flags16(const float* __restrict__ m, int base, int Nm) {
  uint32_t f = 0;
  if (base + PC_ITEMS <= Nm && ((reinterpret_cast<uintptr_t>(m + base) & 15) == 0)) {
    const float4* v4 = reinterpret_cast<const float4*>(m + base);
#pragma unroll
    for (int q = 0; q < PC_ITEMS / 4; ++q) {
      const float4 v = __ldg(v4 + q);
      f |= (v.x > 0.f ? 1u : 0u) << (q * 4 + 0);
      f |= (v.y > 0.f ? 1u : 0u) << (q * 4 + 1);
      f |= (v.z > 0.f ? 1u : 0u) << (q * 4 + 2);
      f |= (v.w > 0.f ? 1u : 0u) << (q * 4 + 3);
    }
  } else {
#pragma unroll
    for (int i = 0; i < PC_ITEMS; ++i)
      if (base + i < Nm && __ldg(m + base + i) > 0.f) f |= 1u << i;
  }
  return f;
}

// same for a uint8 mask (nonzero = valid): 16 pixels = one 16-byte load
__device__ __forceinline__ uint32_t flags16(const uint8_t* __restrict__ m, int base, int Nm) {
  uint32_t f = 0;
  if (base + PC_ITEMS <= Nm && ((reinterpret_cast<uintptr_t>(m + base) & 15) == 0)) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(m + base));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) f |= (((w[q] >> (8 * j)) & 0xFFu) ? 1u : 0u) << (q * 4 + j);
  } else {
#pragma unroll
    for (int i = 0; i < PC_ITEMS; ++i)
      if (base + i < Nm && __ldg(m + base + i) != 0) f |= 1u << i;
  }
  return f;
}
__device__ __forceinline__ bool mask_on(float v) { return v > 0.f; }      // np.where(mask > 0), sampling.py:135
__device__ __forceinline__ bool mask_on(uint8_t v) { return v != 0; }

__device__ __forceinline__ int block_sum_int(int v, int* s_warp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int i = 0; i < PC_THREADS / 32; ++i) t += s_warp[i];
  return t;
}

// np.amin(gt), np.amax(gt) of an image (sampling.py:219-220; NaNs ignored like fminf / fmaxf) on order-preserving
// encodings, both accumulated with atomicMax from zero: slot 0 = ~ordered(min), slot 1 = ordered(max)
struct MinMaxAcc {
  unsigned int omin = 0u, omax = 0u;
  __device__ __forceinline__ void add(float v) {
    if (v == v) {
      const unsigned int o = float_to_ordered(v);
      omax = o > omax ? o : omax;
      omin = ~o > omin ? ~o : omin;
    }
  }
  // CTA-wide (every thread of the CTA calls it): one pair of global atomics per CTA
  __device__ __forceinline__ void flush(unsigned int* acc) {
    __shared__ unsigned int s_mm[2];
    if (threadIdx.x < 2) s_mm[threadIdx.x] = 0u;
    __syncthreads();
    omin = __reduce_max_sync(0xffffffffu, omin);
    omax = __reduce_max_sync(0xffffffffu, omax);
    if ((threadIdx.x & 31) == 0) {
      if (omin) atomicMax(&s_mm[0], omin);
      if (omax) atomicMax(&s_mm[1], omax);
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_mm[threadIdx.x]) atomicMax(acc + threadIdx.x, s_mm[threadIdx.x]);
  }
};

template <typename MT>
__global__ void __launch_bounds__(PC_THREADS) prep_count_kernel(const MT* __restrict__ mask, int Nm, int nchunks,
                                                               int* __restrict__ counts, uint16_t* __restrict__ bits,
                                                               float4* __restrict__ grad4,
                                                               size_t grad_n4, float* __restrict__ grad_tail,
                                                               int grad_tail_n, unsigned int* __restrict__ mm_acc) {
  __shared__ int s_warp[PC_THREADS / 32];
  pdl_sync();
  const int b = blockIdx.y, chunk = blockIdx.x;
  // min / max accumulators of gt (collected by prep_build_kernel, information strategy) start every call from zero
  if (mm_acc != nullptr && chunk == 0 && threadIdx.x < 2) mm_acc[2 * b + threadIdx.x] = 0u;
  const MT* m = mask + (size_t)b * Nm;
  const int base = chunk * PC_CHUNK + threadIdx.x * PC_ITEMS;
  // the validity flags of the thread's 16 pixels are kept as a bit mask (pixel j of the image = bit j & 15 of word
  // j >> 4): the table build and the gradient expansion read 1/8 byte per pixel instead of the mask again
  const uint32_t fl = (base < Nm) ? flags16(m, base, Nm) : 0u;
  bits[((size_t)b * nchunks + chunk) * PC_THREADS + threadIdx.x] = (uint16_t)fl;
  const int c = __popc(fl);
  const int tot = block_sum_int(c, s_warp);
  if (threadIdx.x == 0) counts[b * nchunks + chunk] = tot;
  if (grad4 != nullptr) {  // zero the dense gradient map (grid-stride, 16-byte stores)
    const size_t nb = (size_t)gridDim.x * gridDim.y, bid = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = bid * PC_THREADS + threadIdx.x; i < grad_n4; i += nb * PC_THREADS) grad4[i] = z;
    if (bid == 0 && (int)threadIdx.x < grad_tail_n) grad_tail[threadIdx.x] = 0.f;
  }
}

template <typename MT>
__global__ void __launch_bounds__(PC_THREADS) prep_build_kernel(
    const MT* __restrict__ mask, const float* __restrict__ gt, const float* __restrict__ pred, int Nm, int Wm,
    int W, int HW, double xs, double ys, int identity_scale, int nchunks, const int* __restrict__ counts,
    float2* __restrict__ table, size_t table_stride, int32_t* __restrict__ n_valid, int32_t* __restrict__ vj_flat,
    float* __restrict__ grad_valid, unsigned int* __restrict__ mm_acc, const uint32_t* __restrict__ bits32,
    float* __restrict__ grad_zero) {
  __shared__ int s_warp[PC_THREADS / 32];
  pdl_sync();
  const int b = blockIdx.y, chunk = blockIdx.x;
  int pre = 0, all = 0;
  for (int i = threadIdx.x; i < nchunks; i += PC_THREADS) {
    const int c = counts[b * nchunks + i];
    all += c;
    if (i < chunk) pre += c;
  }
  const int total = block_sum_int(all, s_warp);
  const int prefix = block_sum_int(pre, s_warp);
  float2* tab = table + (size_t)b * table_stride;
  const float* g = gt + (size_t)b * HW;
  MinMaxAcc mm;
  if (identity_scale && total == Nm) {
    const float* s = pred + (size_t)b * HW;
    const bool vec = ((Nm & 3) == 0) && ((table_stride & 1) == 0) &&
                     ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(table) & 15) == 0;
    if (vec) {
      // 4 pixels per thread and iteration: two 16-byte loads, two 16-byte stores, all coalesced
      const float4* g4 = reinterpret_cast<const float4*>(g);
      const float4* s4 = reinterpret_cast<const float4*>(s);
      float4* t4 = reinterpret_cast<float4*>(tab);
#pragma unroll
      for (int i = 0; i < PC_ITEMS / 4; ++i) {
        const int q = chunk * (PC_CHUNK / 4) + i * PC_THREADS + threadIdx.x;   // index of a 4-pixel group
        if (q * 4 < Nm) {
          const float4 a = __ldg(g4 + q), c = __ldg(s4 + q);
          t4[2 * q] = make_float4(a.x, c.x, a.y, c.y);
          t4[2 * q + 1] = make_float4(a.z, c.z, a.w, c.w);
          if (mm_acc != nullptr) { mm.add(a.x); mm.add(a.y); mm.add(a.z); mm.add(a.w); }
        }
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < PC_ITEMS; ++i) {
        const int j = chunk * PC_CHUNK + i * PC_THREADS + threadIdx.x;
        if (j < Nm) {
          const float gv = __ldg(g + j);
          tab[j] = make_float2(gv, __ldg(s + j));
          if (mm_acc != nullptr) mm.add(gv);
        }
      }
    }
    if (grad_zero != nullptr) {
      // bit-mask mode: only full-mask images accumulate straight into the dense gradient map, so only they need it
      // cleared (holed images get every pixel written by bits_expand_kernel)
      float* gz = grad_zero + (size_t)b * HW;
      if (((reinterpret_cast<uintptr_t>(gz) & 15) == 0) && ((Nm & 3) == 0)) {
        float4* z4 = reinterpret_cast<float4*>(gz);
#pragma unroll
        for (int i = 0; i < PC_ITEMS / 4; ++i) {
          const int q = chunk * (PC_CHUNK / 4) + i * PC_THREADS + threadIdx.x;
          if (q * 4 < Nm) z4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
#pragma unroll 4
        for (int i = 0; i < PC_ITEMS; ++i) {
          const int j = chunk * PC_CHUNK + i * PC_THREADS + threadIdx.x;
          if (j < Nm) gz[j] = 0.f;
        }
      }
    }
    if (mm_acc != nullptr) mm.flush(mm_acc + 2 * b);
    if (chunk == 0 && threadIdx.x == 0) n_valid[b] = -Nm;
    return;
  }
  if (mm_acc != nullptr) {
    // holed / scaled image: the table pass below reads gt at valid pixels only; np.amin / np.amax run over ALL pixels
    for (int i = chunk * PC_THREADS + threadIdx.x; i < HW; i += gridDim.x * PC_THREADS) mm.add(__ldg(g + i));
    mm.flush(mm_acc + 2 * b);
  }
  if (grad_valid != nullptr) {   // holed image in valid-index mode: clear its accumulators (valid count <= Nm)
    float* gv = grad_valid + (size_t)b * table_stride;
#pragma unroll 4
    for (int i = 0; i < PC_ITEMS; ++i) {
      const int j = chunk * PC_CHUNK + i * PC_THREADS + threadIdx.x;
      if (j < Nm) gv[j] = 0.f;
    }
  }
  // Compaction with lane-consecutive pixels: warp w of the CTA owns pixels
  // [chunk*CHUNK + w*512, +512) as 16 rows of 32; ballots give each valid pixel its rank, so mask
  // reads, gt reads and table writes are all coalesced.
  // The 32 pixels of row i are one 32-bit word of the bit mask written by prep_count_kernel (pixels past the end: 0).
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wbase = chunk * PC_CHUNK + wid * (PC_ITEMS * 32);
  uint32_t bal[PC_ITEMS];
  int wcount = 0;
  {
    const uint32_t* bw = bits32 + ((size_t)b * nchunks + chunk) * (PC_THREADS / 2) + wid * PC_ITEMS;
    const uint32_t mine = lane < PC_ITEMS ? bw[lane] : 0u;
#pragma unroll
    for (int i = 0; i < PC_ITEMS; ++i) {
      bal[i] = __shfl_sync(0xffffffffu, mine, i);
      wcount += __popc(bal[i]);
    }
  }
  __syncthreads();
  if (lane == 0) s_warp[wid] = wcount;
  __syncthreads();
  int rank = prefix;
  for (int i = 0; i < wid; ++i) rank += s_warp[i];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < PC_ITEMS; ++i) {
    if ((bal[i] >> lane) & 1u) {
      const int idx = wbase + i * 32 + lane;
      int p;
      if (identity_scale) {
        p = idx;
      } else {
        const int rm = idx / Wm, cm = idx - rm * Wm;
        p = (int)((double)rm * xs) * W + (int)((double)cm * ys);  // sampling.py:115-119
      }
      const int pos = rank + __popc(bal[i] & lt);
      if (grad_valid != nullptr) {   // valid-index mode: (gt, pred) by valid index (+ the pixel of every valid index,
                                     // unless the gradient is expanded through the bit mask)
        tab[pos] = make_float2(__ldg(g + p), __ldg(pred + (size_t)b * HW + p));
        if (vj_flat != nullptr) vj_flat[(size_t)b * table_stride + pos] = p;
      } else {
        tab[pos] = make_float2(__int_as_float(p), __ldg(g + p));
      }
    }
    rank += __popc(bal[i]);
  }
  if (chunk == 0 && threadIdx.x == 0) n_valid[b] = total;
}


// ------------------------------------------------------------------------------------------
// radix top-R selection on the ordered scores: three 12-bit MSB-first histogram passes locate, per
// image, the 36-bit key prefix T such that  #{prefix > T} < R <= #{prefix >= T};  every candidate
// with prefix >= T survives (R' >= R of them, the surplus being the ties of the boundary bucket),
// is compacted in candidate order and fully sorted afterwards.
// ------------------------------------------------------------------------------------------
constexpr int SEL_BINS = 4096;

__global__ void __launch_bounds__(256) sel_init_kernel(uint64_t* prefix, int* remaining, int R, unsigned int* hist, int B,
                                                       unsigned long long* bits_or, unsigned long long* bits_and,
                                                       uint64_t* iprefix) {
  pdl_sync();
  for (int i = blockIdx.x * 256 + threadIdx.x; i < B * SEL_BINS; i += gridDim.x * 256) hist[i] = 0u;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < B; i += 256) {
      prefix[i] = 0ull; remaining[i] = R; bits_or[i] = 0ull; bits_and[i] = ~0ull; iprefix[i] = 0ull;
    }
}

// One CTA per image: largest bin t with  sum_{bin >= t} hist >= remaining; extends the key prefix (or, idx_pass, the
// candidate-index prefix of the exact unordered selection below) by `nbits` bits and clears the histogram for the
// next pass.  (Folding this into the histogram kernel through a last-CTA ticket was measured 6-7 us SLOWER per
// pass than this separate 32-CTA launch inside a CUDA graph: the fence after the histogram atomics is the cost.)
__global__ void __launch_bounds__(256) sel_find_kernel(unsigned int* __restrict__ hist, uint64_t* __restrict__ prefix,
                                                       uint64_t* __restrict__ iprefix, int* __restrict__ remaining,
                                                       int idx_pass, int nbits, const int* __restrict__ n_surv, int R,
                                                       int last) {
  pdl_sync();
  __shared__ unsigned int s_sum[8];
  const int b = blockIdx.x;
  if (n_surv != nullptr && n_surv[b] == R) {
    // exact unordered selection of an image whose survivors are exactly the R kept lists (no ties on the boundary):
    // its histograms were skipped; the final cut (key, index) = (0, 0) keeps every survivor
    if (last && threadIdx.x == 0) { prefix[b] = 0ull; iprefix[b] = 0ull; }
    return;
  }
  unsigned int* h = hist + (size_t)b * SEL_BINS;
  // thread t owns the 16 bins [4096 - 16(t+1), 4096 - 16t): thread 0 holds the top bins
  const int hi = SEL_BINS - 16 * threadIdx.x;
  unsigned int loc[16], tot = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { loc[i] = h[hi - 1 - i]; tot += loc[i]; }
  // inclusive scan of the per-thread totals (thread 0 = top bins): the owner is the thread whose range contains
  // the need-th largest candidate
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_sum[wid] = incl;
  __syncthreads();
  for (int w = 0; w < wid; ++w) incl += s_sum[w];
  const unsigned int need = (unsigned int)remaining[b];
  const unsigned int excl = incl - tot;
  if ((excl < need && need <= incl) || (threadIdx.x == 255 && incl < need)) {
    unsigned int above = excl;          // candidates in bins above this thread's range
    int i = 0;
    while (i < 15 && above + loc[i] < need) { above += loc[i]; ++i; }
    const uint64_t bin = (uint64_t)(hi - 1 - i);
    if (idx_pass) iprefix[b] = (iprefix[b] << nbits) | bin;
    else prefix[b] = (prefix[b] << nbits) | bin;
    remaining[b] = (int)(need - above);   // still to take from inside this bin
  }
  // clear the histogram for the next pass
#pragma unroll
  for (int i = 0; i < 16; ++i) h[hi - 1 - i] = 0u;
}

__global__ void __launch_bounds__(256) sel_hist_kernel(const uint64_t* __restrict__ keys, int n, int pass,
                                                       const uint64_t* __restrict__ prefix, unsigned int* __restrict__ hist) {
  pdl_sync();
  __shared__ unsigned int s_h[SEL_BINS];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < SEL_BINS; i += 256) s_h[i] = 0u;
  __syncthreads();
  const uint64_t* k = keys + (size_t)b * n;
  const int shift = 64 - 12 * (pass + 1);
  const uint64_t pre = prefix[b];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const uint64_t key = k[i];
    const bool match = (pass == 0) || ((key >> (64 - 12 * pass)) == pre);
    if (match) atomicAdd(&s_h[(int)((key >> shift) & 0xFFF)], 1u);
  }
  __syncthreads();
  unsigned int* h = hist + (size_t)b * SEL_BINS;
  for (int i = threadIdx.x; i < SEL_BINS; i += 256)
    if (s_h[i]) atomicAdd(h + i, s_h[i]);
}

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) sel_count_kernel(const uint64_t* __restrict__ keys, int n, int ntiles,
                                                              const uint64_t* __restrict__ prefix, int* __restrict__ counts) {
  pdl_sync();
  __shared__ int s_warp[SC_THREADS / 32];
  const int b = blockIdx.y, tile = blockIdx.x;
  const uint64_t* k = keys + (size_t)b * n;
  const uint64_t T = prefix[b];
  int c = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    const int idx = tile * SC_TILE + i * SC_THREADS + threadIdx.x;
    if (idx < n && (k[idx] >> 28) >= T) ++c;
  }
  const int tot = block_sum_int(c, s_warp);
  if (threadIdx.x == 0) counts[b * ntiles + tile] = tot;
}

__global__ void __launch_bounds__(SC_THREADS) sel_compact_kernel(const uint64_t* __restrict__ keys, int n, int ntiles,
                                                                const uint64_t* __restrict__ prefix,
                                                                const int* __restrict__ counts, uint64_t* __restrict__ keys_s,
                                                                uint32_t* __restrict__ vals_s, int* __restrict__ n_surv,
                                                                unsigned long long* __restrict__ bits_or,
                                                                unsigned long long* __restrict__ bits_and) {
  pdl_sync();
  __shared__ int s_warp[SC_THREADS / 32];
  const int b = blockIdx.y, tile = blockIdx.x;
  int pre = 0, all = 0;
  for (int i = threadIdx.x; i < ntiles; i += SC_THREADS) {
    const int c = counts[b * ntiles + i];
    all += c;
    if (i < tile) pre += c;
  }
  const int total = block_sum_int(all, s_warp);
  const int prefix_cnt = block_sum_int(pre, s_warp);
  if (tile == 0 && threadIdx.x == 0) n_surv[b] = total;
  const uint64_t* k = keys + (size_t)b * n;
  const uint64_t T = prefix[b];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wbase = tile * SC_TILE + wid * (SC_ITEMS * 32);   // each warp owns 512 consecutive candidates
  uint32_t bal[SC_ITEMS];
  uint64_t kv[SC_ITEMS];
  int wcount = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    const int idx = wbase + i * 32 + lane;
    kv[i] = (idx < n) ? k[idx] : 0ull;
    bal[i] = __ballot_sync(0xffffffffu, idx < n && (kv[i] >> 28) >= T);
    wcount += __popc(bal[i]);
  }
  __syncthreads();
  if (lane == 0) s_warp[wid] = wcount;
  __syncthreads();
  int rank = prefix_cnt;
  for (int i = 0; i < wid; ++i) rank += s_warp[i];
  const uint32_t lt = (1u << lane) - 1u;
  const size_t off = (size_t)b * n;
  unsigned long long vor = 0ull, vand = ~0ull;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    if ((bal[i] >> lane) & 1u) {
      const int pos = rank + __popc(bal[i] & lt);
      keys_s[off + pos] = kv[i];
      vals_s[off + pos] = (uint32_t)(wbase + i * 32 + lane);
      vor |= kv[i];
      vand &= kv[i];
    }
    rank += __popc(bal[i]);
  }
  // which key bits differ between survivors: OR / AND over all of them (one atomic pair per warp)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vor |= __shfl_xor_sync(0xffffffffu, vor, o);
    vand &= __shfl_xor_sync(0xffffffffu, vand, o);
  }
  if (lane == 0 && wcount > 0) {
    atomicOr(bits_or + b, vor);
    atomicAnd(bits_and + b, vand);
  }
}

// prefix_shift: the unordered selection continues below the 36-bit prefix; when the low 28 key bits are known to be
// zero (compact float32 score keys) its three key passes are skipped and the prefix is completed here
__global__ void sel_varying_kernel(const unsigned long long* bits_or, const unsigned long long* bits_and,
                                   unsigned long long* varying, int B, uint64_t* prefix, int prefix_shift) {
  pdl_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    varying[b] = bits_or[b] ^ bits_and[b];
    if (prefix_shift) prefix[b] <<= prefix_shift;
  }
}

// ------------------------------------------------------------------------------------------
// Unordered top-R (rankings not materialised: the loss does not depend on the order of the lists).  Instead of
// sorting the survivors, the radix selection is refined on them down to the exact (key, candidate index) cut:
// three more passes over the remaining 28 key bits, two over the 23 index bits (ties: larger index first),
// each restricted to the boundary bucket of the previous pass.  Kept candidates are then compacted in
// candidate order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sel2_hist_kernel(const uint64_t* __restrict__ keys_s, const uint32_t* __restrict__ vals_s,
                                                        const int* __restrict__ n_surv, size_t stride, int idx_pass,
                                                        int shift, int nbits, const uint64_t* __restrict__ prefix,
                                                        const uint64_t* __restrict__ iprefix, unsigned int* __restrict__ hist,
                                                        int R) {
  pdl_sync();
  __shared__ unsigned int s_h[SEL_BINS];
  const int b = blockIdx.y;
  const int n = n_surv[b];
  if (n == R) return;   // no boundary ties in this image: every survivor is kept (sel_find_kernel opens the cut)
  for (int i = threadIdx.x; i < SEL_BINS; i += 256) s_h[i] = 0u;
  __syncthreads();
  const uint64_t* k = keys_s + (size_t)b * stride;
  const uint32_t* v = vals_s + (size_t)b * stride;
  const uint64_t pre = prefix[b], ipre = iprefix[b];
  const uint32_t msk = (1u << nbits) - 1u;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const uint64_t key = k[i];
    if (!idx_pass) {
      if ((key >> (shift + nbits)) == pre) atomicAdd(&s_h[(uint32_t)(key >> shift) & msk], 1u);
    } else if (key == pre) {
      const uint32_t id = v[i];
      if ((uint64_t)(id >> (shift + nbits)) == ipre) atomicAdd(&s_h[(id >> shift) & msk], 1u);
    }
  }
  __syncthreads();
  unsigned int* h = hist + (size_t)b * SEL_BINS;
  for (int i = threadIdx.x; i < SEL_BINS; i += 256)
    if (s_h[i]) atomicAdd(h + i, s_h[i]);
}

__device__ __forceinline__ bool sel2_keep(uint64_t key, uint32_t id, uint64_t tkey, uint32_t tidx) {
  return key > tkey || (key == tkey && id >= tidx);
}

__global__ void __launch_bounds__(SC_THREADS) sel2_count_kernel(const uint64_t* __restrict__ keys_s,
                                                               const uint32_t* __restrict__ vals_s,
                                                               const int* __restrict__ n_surv, size_t stride, int ntiles,
                                                               const uint64_t* __restrict__ prefix,
                                                               const uint64_t* __restrict__ iprefix, int* __restrict__ counts) {
  pdl_sync();
  __shared__ int s_warp[SC_THREADS / 32];
  const int b = blockIdx.y, tile = blockIdx.x;
  const int n = n_surv[b];
  const uint64_t* k = keys_s + (size_t)b * stride;
  const uint32_t* v = vals_s + (size_t)b * stride;
  const uint64_t tkey = prefix[b];
  const uint32_t tidx = (uint32_t)iprefix[b];
  int c = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    const int idx = tile * SC_TILE + i * SC_THREADS + threadIdx.x;
    if (idx < n && sel2_keep(k[idx], v[idx], tkey, tidx)) ++c;
  }
  const int tot = block_sum_int(c, s_warp);
  if (threadIdx.x == 0) counts[b * ntiles + tile] = tot;
}

__global__ void __launch_bounds__(SC_THREADS) sel2_compact_kernel(const uint64_t* __restrict__ keys_s,
                                                                 const uint32_t* __restrict__ vals_s,
                                                                 const int* __restrict__ n_surv, size_t stride, int ntiles,
                                                                 const uint64_t* __restrict__ prefix,
                                                                 const uint64_t* __restrict__ iprefix,
                                                                 const int* __restrict__ counts, int R,
                                                                 uint32_t* __restrict__ order, int32_t* __restrict__ order_out) {
  pdl_sync();
  __shared__ int s_warp[SC_THREADS / 32];
  const int b = blockIdx.y, tile = blockIdx.x;
  int pre = 0;
  for (int i = threadIdx.x; i < tile; i += SC_THREADS) pre += counts[b * ntiles + i];
  const int prefix_cnt = block_sum_int(pre, s_warp);
  const int n = n_surv[b];
  const uint64_t* k = keys_s + (size_t)b * stride;
  const uint32_t* v = vals_s + (size_t)b * stride;
  const uint64_t tkey = prefix[b];
  const uint32_t tidx = (uint32_t)iprefix[b];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wbase = tile * SC_TILE + wid * (SC_ITEMS * 32);
  uint32_t bal[SC_ITEMS], id[SC_ITEMS];
  int wcount = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    const int idx = wbase + i * 32 + lane;
    id[i] = (idx < n) ? v[idx] : 0u;
    bal[i] = __ballot_sync(0xffffffffu, idx < n && sel2_keep(k[idx], id[i], tkey, tidx));
    wcount += __popc(bal[i]);
  }
  __syncthreads();
  if (lane == 0) s_warp[wid] = wcount;
  __syncthreads();
  int rank = prefix_cnt;
  for (int i = 0; i < wid; ++i) rank += s_warp[i];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    if ((bal[i] >> lane) & 1u) {
      const int pos = rank + __popc(bal[i] & lt);
      if (pos < R) {   // exactly R are kept by construction; the guard only protects the buffer
        order[(size_t)b * R + pos] = id[i];
        if (order_out != nullptr) order_out[(size_t)b * R + pos] = (int32_t)id[i];
      }
    }
    rank += __popc(bal[i]);
  }
}

// kept candidates in final order (score descending): tail of the sorted survivors read backwards, taken from
// whichever ping-pong buffer the image's last executed pass wrote
__global__ void __launch_bounds__(256) sel_order_kernel(const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b,
                                                        const unsigned long long* __restrict__ varying,
                                                        const int* __restrict__ n_surv, int n, int R,
                                                        uint32_t* __restrict__ order, int32_t* __restrict__ order_out) {
  pdl_sync();
  const int b = blockIdx.y;
  const int last = n_surv[b] - 1;
  const unsigned long long v = varying[b];
  int passes = 0;
  for (int p = 0; p < 8; ++p) passes += ((v >> (8 * p)) & 0xFFull) ? 1 : 0;
  const uint32_t* src = ((passes & 1) ? vals_b : vals_a) + (size_t)b * n;
  for (int j = blockIdx.x * 256 + threadIdx.x; j < R; j += gridDim.x * 256) {
    const uint32_t c = src[last - j];
    order[(size_t)b * R + j] = c;
    if (order_out != nullptr) order_out[(size_t)b * R + j] = (int32_t)c;
  }
}

// valid-index mode: scatter the per-valid-index gradient of every holed image to its pixels (the dense map was
// zeroed by prep_count_kernel; full-mask images accumulated straight into it)
__global__ void __launch_bounds__(256) vj_expand_kernel(const float* __restrict__ grad_valid,
                                                        const int32_t* __restrict__ vj_flat,
                                                        const int32_t* __restrict__ n_valid, size_t table_stride, int HW,
                                                        float* __restrict__ grad) {
  pdl_sync();
  const int b = blockIdx.y;
  const int M = n_valid[b];
  if (M <= 0) return;   // identity table (negative count) or empty mask
  const float* gv = grad_valid + (size_t)b * table_stride;
  const int32_t* vf = vj_flat + (size_t)b * table_stride;
  float* gr = grad + (size_t)b * HW;
  for (int j = blockIdx.x * 256 + threadIdx.x; j < M; j += gridDim.x * 256) gr[vf[j]] = gv[j];
}

// valid-index mode, mask at image resolution: every pixel of a holed image gets its gradient -- the accumulator of its
// valid index, or zero -- straight from the bit mask (rank of a pixel = valid pixels before it: chunk counts + popcounts),
// so neither a pixel list nor a cleared dense map is needed.  Same pixel ownership as prep_build_kernel.
__global__ void __launch_bounds__(PC_THREADS) bits_expand_kernel(const float* __restrict__ grad_valid,
                                                                 const uint32_t* __restrict__ bits32,
                                                                 const int* __restrict__ counts,
                                                                 const int32_t* __restrict__ n_valid, int nchunks,
                                                                 size_t table_stride, int HW, float* __restrict__ grad) {
  __shared__ int s_warp[PC_THREADS / 32];
  pdl_sync();
  const int b = blockIdx.y, chunk = blockIdx.x;
  if (n_valid[b] < 0) return;   // full mask: the list kernel accumulated straight into the dense map
  int pre = 0;
  for (int i = threadIdx.x; i < chunk; i += PC_THREADS) pre += counts[b * nchunks + i];
  const int prefix = block_sum_int(pre, s_warp);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wbase = chunk * PC_CHUNK + wid * (PC_ITEMS * 32);
  const uint32_t* bw = bits32 + ((size_t)b * nchunks + chunk) * (PC_THREADS / 2) + wid * PC_ITEMS;
  const uint32_t mine = lane < PC_ITEMS ? bw[lane] : 0u;
  uint32_t bal[PC_ITEMS];
  int wcount = 0;
#pragma unroll
  for (int i = 0; i < PC_ITEMS; ++i) {
    bal[i] = __shfl_sync(0xffffffffu, mine, i);
    wcount += __popc(bal[i]);
  }
  __syncthreads();
  if (lane == 0) s_warp[wid] = wcount;
  __syncthreads();
  int rank = prefix;
  for (int i = 0; i < wid; ++i) rank += s_warp[i];
  const uint32_t lt = (1u << lane) - 1u;
  const float* gv = grad_valid + (size_t)b * table_stride;
  float* gr = grad + (size_t)b * HW;
#pragma unroll
  for (int i = 0; i < PC_ITEMS; ++i) {
    const int idx = wbase + i * 32 + lane;
    if (idx < HW) gr[idx] = ((bal[i] >> lane) & 1u) ? gv[rank + __popc(bal[i] & lt)] : 0.f;
    rank += __popc(bal[i]);
  }
}

}  // namespace pld

using namespace pld;

// counts [B, nchunks] followed by the bit mask [B, nchunks, 256] x uint16 (prep_count_kernel -> prep_build_kernel)
static size_t prep_scratch_bytes(int B, int nchunks) {
  const size_t c = (sizeof(int) * (size_t)B * nchunks + 255) & ~(size_t)255;
  return c + sizeof(uint16_t) * (size_t)B * nchunks * PC_THREADS;
}
static uint16_t* prep_bits(int* counts, int B, int nchunks) {
  const size_t c = (sizeof(int) * (size_t)B * nchunks + 255) & ~(size_t)255;
  return reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(counts) + c);
}

// Passes 1-2 of both fused steps: valid pixels counted per chunk while `grad` is zeroed, then the per-image lookup
// tables (layouts: see prep_build_kernel).
template <typename MT>
static int launch_prep(const MT* mask, const float* gt, const float* pred, int B, int Hm, int Wm, int H, int W,
                       int* counts, float2* table, size_t tstride, int32_t* nv, int32_t* vj_flat, float* grad_valid,
                       float* grad, cudaStream_t st, unsigned int* mm_acc = nullptr, bool bit_mode = false) {
  const int HW = H * W, Nm = Hm * Wm;
  const int nchunks = (Nm + PC_CHUNK - 1) / PC_CHUNK;
  dim3 grid((unsigned)nchunks, (unsigned)B);
  const size_t gtotal = (size_t)B * HW;
  float4* g4 = nullptr;
  size_t n4 = 0;
  int tail = 0;
  // bit_mode (valid-index mode with the mask at image resolution): `grad` is cleared by prep_build_kernel for
  // full-mask images and completely written by bits_expand_kernel for holed ones, not here
  if (grad != nullptr && !bit_mode) {
    PLD_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad must be 16-byte aligned");
    g4 = reinterpret_cast<float4*>(grad);
    n4 = gtotal / 4;
    tail = (int)(gtotal - n4 * 4);
  }
  uint16_t* bits = prep_bits(counts, B, nchunks);
  PLD_CUDA(launch_pdl(prep_count_kernel<MT>, grid, dim3(PC_THREADS), 0, st, mask, Nm, nchunks, counts, bits, g4, n4,
                      g4 ? grad + n4 * 4 : nullptr, tail, mm_acc));
  PLD_CHECK_LAUNCH();
  const double xs = (double)H / (double)Hm, ys = (double)W / (double)Wm;
  const int identity_scale = (H == Hm && W == Wm) ? 1 : 0;
  PLD_CUDA(launch_pdl(prep_build_kernel<MT>, grid, dim3(PC_THREADS), 0, st, mask, gt, pred, Nm, Wm, W, HW, xs, ys,
                      identity_scale, nchunks, (const int*)counts, table, tstride, nv, vj_flat, grad_valid, mm_acc,
                      reinterpret_cast<const uint32_t*>(bits), bit_mode ? grad : (float*)nullptr));
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

template <typename MT>
static int fused_step_impl(pld_ctx* ctx, const MT* mask, const float* gt, const float* pred, int B, int Hm,
                           int Wm, int H, int W, int K, int n, uint64_t seed, uint64_t offset, int image_base,
                           float scale, int32_t* n_valid, float* rankings, float* loss, double* loss_sum,
                           float* per_list, float* grad, void* stream) {
  PLD_REQUIRE(ctx && mask && gt && pred && loss, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(B > 0 && B <= 65535 && Hm > 0 && Wm > 0 && H > 0 && W > 0, "bad shape");
  PLD_REQUIRE((long long)H * W <= PLD_MAX_PIXELS && (long long)Hm * Wm <= PLD_MAX_PIXELS, "map too large");
  PLD_REQUIRE(K >= 1 && K <= PLD_MAX_RANKING_SIZE, "ranking_size must be in [1, 512]");
  PLD_REQUIRE(n >= 0 && (long long)B * n < (1ll << 31), "bad list count");
  PLD_REQUIRE((offset >> 48) == 0, "offset must fit in 48 bits");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W, Nm = Hm * Wm;
  const int nchunks = (Nm + PC_CHUNK - 1) / PC_CHUNK;
  const size_t tstride = (size_t)(HW > Nm ? HW : Nm);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t off_counts = 0, off_nv = al(prep_scratch_bytes(B, nchunks));
  const size_t off_tab = off_nv + al(sizeof(int32_t) * (size_t)B);
  // valid-index mode: holed masks, gradient wanted, rankings not materialised (nothing needs the pixel index)
  // -- only while the mask -> image map int(r * H / Hm) is injective (mask not finer than the image): a finer mask sends
  // several valid indices to one pixel, whose contributions must ACCUMULATE, which the dense-map path does
  const bool vj_mode = (rankings == nullptr) && (grad != nullptr) && !ctx->deterministic && Hm <= H && Wm <= W;
  // ... and with the mask at image resolution the pixel of a valid index follows from the bit mask: no pixel list, no
  // cleared dense map for holed images (bits_expand_kernel writes every pixel)
  const bool bit_mode = vj_mode && Hm == H && Wm == W && n > 0;
  const size_t off_vjf = off_tab + al(sizeof(float2) * (size_t)B * tstride);
  const size_t off_gv = off_vjf + ((vj_mode && !bit_mode) ? al(sizeof(int32_t) * (size_t)B * tstride) : 0);
  const size_t scratch_total = off_gv + (vj_mode ? al(sizeof(float) * (size_t)B * tstride) : 0);
  int rc = ctx->ensure_scratch(scratch_total);
  if (rc) return rc;
  int* counts = (int*)((char*)ctx->d_scratch + off_counts);
  int32_t* nv = n_valid ? n_valid : (int32_t*)((char*)ctx->d_scratch + off_nv);
  float2* table = (float2*)((char*)ctx->d_scratch + off_tab);
  int32_t* vj_flat = (vj_mode && !bit_mode) ? (int32_t*)((char*)ctx->d_scratch + off_vjf) : nullptr;
  float* grad_valid = vj_mode ? (float*)((char*)ctx->d_scratch + off_gv) : nullptr;
  const int per_image_cap = lists_per_image_cap(ctx->num_sms, B);
  rc = ctx->ensure_partials(per_image_cap * B + B);
  if (rc) return rc;

  const size_t gtotal = (size_t)B * HW;
  rc = launch_prep(mask, gt, pred, B, Hm, Wm, H, W, counts, table, tstride, nv, vj_flat, grad_valid, grad, st, nullptr,
                   bit_mode);
  if (rc) return rc;
  if (n == 0) {
    PLD_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if (loss_sum) PLD_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), st));
    return PLD_OK;
  }
  ListParams P = {};
  P.gt = gt; P.pred = pred; P.n_valid = nv; P.valid_flat = nullptr; P.table = table; P.table_stride = tstride;
  P.rank_out = rankings; P.per_list = per_list; P.grad = grad; P.loss = loss; P.loss_sum = loss_sum;
  P.partials = ctx->d_partials; P.ticket = ctx->d_ticket; P.status = ctx->d_status;
  P.B = B; P.HW = HW; P.valid_stride = 0; P.n = n; P.K = K; P.scale = scale;
  P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32);
  philox_round_keys(P.seed_lo, P.seed_hi, P.rk0, P.rk1);
  P.off_lo = (uint32_t)offset; P.off_hi16 = (uint32_t)((offset >> 32) & 0xFFFFu) << 16;
  P.image_base = image_base;
  P.grad_valid = grad_valid;
  if (ctx->use_device_offset) P.offset_dev = ctx->d_offset;
  if (grad != nullptr && ctx->deterministic) {
    rc = ctx->ensure_acc(gtotal);
    if (rc) return rc;
    P.acc = ctx->d_acc;
    PLD_CUDA(cudaMemsetAsync(P.acc, 0, sizeof(long long) * gtotal, st));
  }
  ctx->time_begin(st);
  rc = (K <= 16) ? launch_lists_small(P, SRC_PHILOX_TAB, true, ctx->num_sms, st)
                 : launch_lists_large(P, SRC_PHILOX_TAB, true, ctx->num_sms, st);
  ctx->time_end(st);
  if (rc == PLD_OK && bit_mode) {
    PLD_CUDA(launch_pdl(bits_expand_kernel, dim3((unsigned)nchunks, (unsigned)B), dim3(PC_THREADS), 0, st,
                        (const float*)grad_valid, reinterpret_cast<const uint32_t*>(prep_bits(counts, B, nchunks)),
                        (const int*)counts, (const int32_t*)nv, nchunks, tstride, HW, grad));
    PLD_CHECK_LAUNCH();
  } else if (rc == PLD_OK && vj_mode) {
    int gx = (int)((tstride + 255) / 256);
    if (gx > per_image_cap) gx = per_image_cap;
    PLD_CUDA(launch_pdl(vj_expand_kernel, dim3((unsigned)gx, (unsigned)B), dim3(256), 0, st, (const float*)grad_valid,
                        (const int32_t*)vj_flat, (const int32_t*)nv, tstride, HW, grad));
    PLD_CHECK_LAUNCH();
  }
  if (rc == PLD_OK && P.acc != nullptr) rc = launch_acc_finalize(ctx, grad, gtotal, scale, 0, st);
  if (rc == PLD_OK && ctx->use_device_offset) rc = launch_offset_advance(ctx, st);
  return rc;
}

extern "C" int pld_fused_step(pld_ctx* ctx, const float* mask, const float* gt, const float* pred, int B, int Hm,
                              int Wm, int H, int W, int K, int n, uint64_t seed, uint64_t offset, int image_base,
                              float scale, int32_t* n_valid, float* rankings, float* loss, double* loss_sum,
                              float* per_list, float* grad, void* stream) {
  return fused_step_impl<float>(ctx, mask, gt, pred, B, Hm, Wm, H, W, K, n, seed, offset, image_base, scale, n_valid,
                                rankings, loss, loss_sum, per_list, grad, stream);
}

extern "C" int pld_fused_step_m8(pld_ctx* ctx, const uint8_t* mask, const float* gt, const float* pred, int B, int Hm,
                                 int Wm, int H, int W, int K, int n, uint64_t seed, uint64_t offset, int image_base,
                                 float scale, int32_t* n_valid, float* rankings, float* loss, double* loss_sum,
                                 float* per_list, float* grad, void* stream) {
  return fused_step_impl<uint8_t>(ctx, mask, gt, pred, B, Hm, Wm, H, W, K, n, seed, offset, image_base, scale, n_valid,
                                  rankings, loss, loss_sum, per_list, grad, stream);
}

extern "C" int pld_fused_step_scored(pld_ctx* ctx, const float* mask, const float* gt, const float* pred, int B, int Hm,
                                     int Wm, int H, int W, int K, int n, int R, int strategy, double threshold,
                                     double equality_penalty, int promotion, uint64_t seed, uint64_t offset,
                                     int image_base, float scale, int32_t* n_valid, int32_t* order_out, float* rankings,
                                     float* loss, double* loss_sum, float* per_list, float* grad, void* stream) {
  PLD_REQUIRE(ctx && mask && gt, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(pred != nullptr || (loss == nullptr && grad == nullptr && per_list == nullptr), "pred is required for the loss");
  PLD_REQUIRE(rankings != nullptr || loss != nullptr, "no output requested");
  PLD_REQUIRE(B > 0 && B <= 65535 && Hm > 0 && Wm > 0 && H > 0 && W > 0, "bad shape");
  PLD_REQUIRE((long long)H * W <= PLD_MAX_PIXELS && (long long)Hm * Wm <= PLD_MAX_PIXELS, "map too large");
  PLD_REQUIRE(K >= 1 && K <= PLD_MAX_RANKING_SIZE, "ranking_size must be in [1, 512]");
  PLD_REQUIRE(n >= 1 && R >= 1 && R <= n && (long long)B * n < (1ll << 31), "need 1 <= R <= n candidates");
  PLD_REQUIRE(strategy >= PLD_STRATEGY_MASKED && strategy <= PLD_STRATEGY_INFORMATION, "bad strategy");
  PLD_REQUIRE(promotion == PLD_PROMOTION_NEP50 || promotion == PLD_PROMOTION_LEGACY, "bad promotion");
  PLD_REQUIRE((offset >> 48) == 0, "offset must fit in 48 bits");
  PLD_REQUIRE(n <= (1 << 23), "at most 2^23 candidate lists per image");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W, Nm = Hm * Wm;
  const int nchunks = (Nm + PC_CHUNK - 1) / PC_CHUNK;
  const int ntiles = (n + SC_TILE - 1) / SC_TILE;
  const size_t tstride = (size_t)(HW > Nm ? HW : Nm);
  const size_t total = (size_t)B * n;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += al(bytes); return o; };
  const size_t o_counts = take(prep_scratch_bytes(B, nchunks));
  const size_t o_nv = take(sizeof(int32_t) * B);
  const size_t o_mm = take(sizeof(float) * 2 * B);
  // rankings not materialised, many candidates: sampled-window selection (pld_pilot.cu) -- no key array, no histogram
  // passes; the deterministic mode keeps the order-preserving path below (the window path appends with atomics)
  const bool pilot = (rankings == nullptr) && !ctx->deterministic && K <= 16 && pilot_select_fits(n) &&
                     getenv("PLD_NO_PILOT") == nullptr;
  if (pilot) {
    size_t poffs[16];
    const size_t o_pilot = take(pilot_select_bytes(B, n, R, ctx->num_sms, poffs));
    const size_t o_ptab = take(sizeof(float2) * (size_t)B * tstride);
    int prc = ctx->ensure_scratch(off);
    if (prc) return prc;
    char* psb = (char*)ctx->d_scratch;
    int* pcounts = (int*)(psb + o_counts);
    int32_t* pnv = n_valid ? n_valid : (int32_t*)(psb + o_nv);
    float2* ptable = (float2*)(psb + o_ptab);
    const int pcap = lists_per_image_cap(ctx->num_sms, B);
    prc = ctx->ensure_partials(pcap * B + B);
    if (prc) return prc;
    const size_t pgtotal = (size_t)B * HW;
    unsigned int* pmm = nullptr;
    if (strategy == PLD_STRATEGY_INFORMATION) {
      prc = ctx->ensure_mm(B);
      if (prc) return prc;
      pmm = ctx->d_mm_acc;
    }
    prc = launch_prep(mask, gt, pred ? pred : gt, B, Hm, Wm, H, W, pcounts, ptable, tstride, pnv, nullptr, nullptr, grad, st,
                      pmm);
    if (prc) return prc;
    ListParams P = {};
    P.gt = gt; P.pred = pred ? pred : gt; P.n_valid = pnv; P.table = ptable; P.table_stride = tstride;
    P.partials = ctx->d_partials; P.ticket = ctx->d_ticket; P.status = ctx->d_status;
    P.B = B; P.HW = HW; P.n = n; P.K = K; P.scale = scale;
    P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32);
  philox_round_keys(P.seed_lo, P.seed_hi, P.rk0, P.rk1);
    P.off_lo = (uint32_t)offset; P.off_hi16 = (uint32_t)((offset >> 32) & 0xFFFFu) << 16;
    P.image_base = image_base;
    if (ctx->use_device_offset) P.offset_dev = ctx->d_offset;
    P.score_cfg = make_score_cfg(nullptr, strategy, threshold, equality_penalty, promotion);
    P.score_cfg.gt_minmax_enc = pmm;
    const int plow = (promotion == PLD_PROMOTION_NEP50 && strategy != PLD_STRATEGY_INFORMATION) ? 1 : 0;
    uint32_t* porder = nullptr;
    prc = pilot_select(P, R, plow, psb + o_pilot, order_out, ctx->num_sms, &porder, st);
    if (prc) return prc;
    P.n = R;
    P.list_map = porder; P.map_stride = (size_t)R;
    P.per_list = per_list; P.grad = grad; P.loss = loss; P.loss_sum = loss_sum;
    ctx->time_begin(st);
    prc = launch_lists_small(P, SRC_PHILOX_TAB, loss != nullptr, ctx->num_sms, st);
    ctx->time_end(st);
    if (prc == PLD_OK && ctx->use_device_offset) prc = launch_offset_advance(ctx, st);
    (void)pgtotal;
    return prc;
  }
  const size_t o_prefix = take(sizeof(uint64_t) * B);
  const size_t o_rem = take(sizeof(int) * B);
  const size_t o_nsurv = take(sizeof(int) * B);
  const size_t o_shist = take(sizeof(unsigned int) * (size_t)B * SEL_BINS);
  const size_t o_tcnt = take(sizeof(int) * (size_t)B * ntiles);
  const size_t o_keys = take(sizeof(uint64_t) * total);
  const size_t o_k0 = take(sizeof(uint64_t) * total);
  const size_t o_k1 = take(sizeof(uint64_t) * total);
  const size_t o_v0 = take(sizeof(uint32_t) * total);
  const size_t o_v1 = take(sizeof(uint32_t) * total);
  const size_t o_rhist = take(seg_radix_sort_hist_bytes(n, B));
  const size_t o_bits = take(sizeof(unsigned long long) * 4 * (size_t)B);
  const size_t o_order = take(sizeof(uint32_t) * (size_t)B * R);
  const size_t o_tab = take(sizeof(float2) * (size_t)B * tstride);
  // ranking_size > 16, rankings not materialised: the sampled-window selection runs over the stored key array
  const bool pilot_keys = (rankings == nullptr) && !ctx->deterministic && K > 16 && pilot_select_fits(n) &&
                          getenv("PLD_NO_PILOT") == nullptr;
  size_t poffs2[16];
  const size_t o_pilot2 = pilot_keys ? take(pilot_select_bytes(B, n, R, ctx->num_sms, poffs2)) : 0;
  int rc = ctx->ensure_scratch(off);
  if (rc) return rc;
  char* sb = (char*)ctx->d_scratch;
  int* counts = (int*)(sb + o_counts);
  int32_t* nv = n_valid ? n_valid : (int32_t*)(sb + o_nv);
  float* minmax = (float*)(sb + o_mm);
  uint64_t* prefix = (uint64_t*)(sb + o_prefix);
  int* remaining = (int*)(sb + o_rem);
  int* n_surv = (int*)(sb + o_nsurv);
  unsigned int* shist = (unsigned int*)(sb + o_shist);
  int* tcnt = (int*)(sb + o_tcnt);
  uint64_t* keys = (uint64_t*)(sb + o_keys);
  uint64_t* k0 = (uint64_t*)(sb + o_k0);
  uint64_t* k1 = (uint64_t*)(sb + o_k1);
  uint32_t* v0 = (uint32_t*)(sb + o_v0);
  uint32_t* v1 = (uint32_t*)(sb + o_v1);
  int* rhist = (int*)(sb + o_rhist);
  unsigned long long* bits_or = (unsigned long long*)(sb + o_bits);
  unsigned long long* bits_and = bits_or + B;
  unsigned long long* varying = bits_or + 2 * B;
  uint64_t* iprefix = (uint64_t*)(bits_or + 3 * B);
  uint32_t* order = (uint32_t*)(sb + o_order);
  float2* table = (float2*)(sb + o_tab);
  const int per_image_cap = lists_per_image_cap(ctx->num_sms, B);
  rc = ctx->ensure_partials(per_image_cap * B + B);
  if (rc) return rc;

  // 1. valid pixels + zeroed grad + lookup tables (as pld_fused_step)
  const size_t gtotal = (size_t)B * HW;
  unsigned int* mm_acc = nullptr;
  if (strategy == PLD_STRATEGY_INFORMATION) {
    rc = ctx->ensure_mm(B);
    if (rc) return rc;
    mm_acc = ctx->d_mm_acc;
  }
  rc = launch_prep(mask, gt, pred ? pred : gt, B, Hm, Wm, H, W, counts, table, tstride, nv, nullptr, nullptr, grad, st, mm_acc);
  if (rc) return rc;

  // 2. scoring pass over the n candidates of every image: ordered scores only
  ListParams P = {};
  P.gt = gt; P.pred = pred ? pred : gt; P.n_valid = nv; P.table = table; P.table_stride = tstride;
  P.partials = ctx->d_partials; P.ticket = ctx->d_ticket; P.status = ctx->d_status;
  P.B = B; P.HW = HW; P.n = n; P.K = K; P.scale = scale;
  P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32);
  philox_round_keys(P.seed_lo, P.seed_hi, P.rk0, P.rk1);
  P.off_lo = (uint32_t)offset; P.off_hi16 = (uint32_t)((offset >> 32) & 0xFFFFu) << 16;
  P.image_base = image_base;
  if (ctx->use_device_offset) P.offset_dev = ctx->d_offset;
  P.score_keys = keys;
  P.score_cfg = make_score_cfg(nullptr, strategy, threshold, equality_penalty, promotion);
  P.score_cfg.gt_minmax_enc = mm_acc;
  const bool radix_select = !select_small_fits(n) && !pilot_keys;
  const bool fused_hist = K <= 16;     // the thread-per-list scoring kernel takes the first histogram of the selection
  if (radix_select) {
    PLD_CUDA(launch_pdl(sel_init_kernel, dim3((B * SEL_BINS + 255) / 256), dim3(256), 0, st, prefix, remaining, R, shist, B, bits_or, bits_and, iprefix));
    PLD_CHECK_LAUNCH();
    if (fused_hist) P.sel_hist = shist;
  }
  rc = (K <= 16) ? launch_lists_small_score(P, ctx->num_sms, st)
       : score_reg_fits(P) ? launch_score_reg(P, ctx->num_sms, st) : launch_lists_tab_score(P, ctx->num_sms, st);
  if (rc) return rc;
  P.sel_hist = nullptr;

  if (pilot_keys) {
    // 3". unordered top-R set through the sampled window (pld_pilot.cu), sample = the first 8192 stored keys
    const bool low0 = (promotion == PLD_PROMOTION_NEP50 && strategy != PLD_STRATEGY_INFORMATION);
    uint32_t* porder = nullptr;
    rc = pilot_select_keys(keys, B, n, nv, ctx->d_status, R, low0 ? 1 : 0, sb + o_pilot2, order_out, ctx->num_sms, &porder, st);
    if (rc) return rc;
    order = porder;
  } else if (!radix_select) {
    // 3'. few candidates per image (the sizes the reference runs): one shared-memory sort per image
    rc = select_small(keys, nullptr, n, (size_t)n, B, R, rankings == nullptr, order, order_out, st);
    if (rc) return rc;
  } else {
  // 3. radix top-R selection -> survivors in candidate order
  int gsel = (n + 255) / 256;
  if (gsel > per_image_cap) gsel = per_image_cap;
  for (int pass = 0; pass < 3; ++pass) {
    if (pass > 0 || !fused_hist) {   // pass 0 was histogrammed by the thread-per-list scoring kernel
      PLD_CUDA(launch_pdl(sel_hist_kernel, dim3(dim3((unsigned)gsel, (unsigned)B)), dim3(256), 0, st, keys, n, pass, prefix, shist));
      PLD_CHECK_LAUNCH();
    }
    PLD_CUDA(launch_pdl(sel_find_kernel, dim3(B), dim3(256), 0, st, shist, prefix, nullptr, remaining, 0, 12, nullptr, 0, 0));
    PLD_CHECK_LAUNCH();
  }
  dim3 tgrid((unsigned)ntiles, (unsigned)B);
  PLD_CUDA(launch_pdl(sel_count_kernel, dim3(tgrid), dim3(SC_THREADS), 0, st, keys, n, ntiles, prefix, tcnt));
  PLD_CHECK_LAUNCH();
  PLD_CUDA(launch_pdl(sel_compact_kernel, dim3(tgrid), dim3(SC_THREADS), 0, st, keys, n, ntiles, prefix, tcnt, k0, v0, n_surv, bits_or, bits_and));
  PLD_CHECK_LAUNCH();
  // compact float32 score keys (pld_score.cuh: score_key_f32) carry nothing below bit 32
  const bool low_bits_zero = (promotion == PLD_PROMOTION_NEP50 && strategy != PLD_STRATEGY_INFORMATION);
  const int skip_key_passes = (rankings == nullptr && low_bits_zero) ? 1 : 0;
  PLD_CUDA(launch_pdl(sel_varying_kernel, dim3((B + 255) / 256), dim3(256), 0, st, bits_or, bits_and, varying, B, prefix, skip_key_passes ? 28 : 0));
  PLD_CHECK_LAUNCH();

  if (rankings == nullptr) {
    // 4a. nobody sees the order of the kept lists: refine the selection to the exact cut instead of sorting
    static const int kPass[5][3] = {{0, 16, 12}, {0, 4, 12}, {0, 0, 4}, {1, 11, 12}, {1, 0, 11}};  // idx?, shift, bits
    for (int p = skip_key_passes ? 3 : 0; p < 5; ++p) {
      PLD_CUDA(launch_pdl(sel2_hist_kernel, dim3(dim3((unsigned)gsel, (unsigned)B)), dim3(256), 0, st, k0, v0, n_surv, (size_t)n, kPass[p][0], kPass[p][1], kPass[p][2], prefix, iprefix, shist, R));
      PLD_CHECK_LAUNCH();
      PLD_CUDA(launch_pdl(sel_find_kernel, dim3(B), dim3(256), 0, st, shist, prefix, iprefix, remaining, kPass[p][0], kPass[p][2], n_surv, R, p == 4 ? 1 : 0));
      PLD_CHECK_LAUNCH();
    }
    PLD_CUDA(launch_pdl(sel2_count_kernel, dim3(tgrid), dim3(SC_THREADS), 0, st, k0, v0, n_surv, (size_t)n, ntiles, prefix, iprefix, tcnt));
    PLD_CHECK_LAUNCH();
    PLD_CUDA(launch_pdl(sel2_compact_kernel, dim3(tgrid), dim3(SC_THREADS), 0, st, k0, v0, n_surv, (size_t)n, ntiles, prefix, iprefix, tcnt, R, order, order_out));
    PLD_CHECK_LAUNCH();
  } else {
    // 4b. full order of the survivors (ascending, stable LSD radix sort; the four low key bytes are skipped when
    // they are known to be zero); the best R are the tail read backwards
    rc = seg_radix_sort(ctx, k0, v0, k1, v1, n_surv, n, (size_t)n, B, rhist, varying, low_bits_zero ? 4 : 0, st);
    if (rc) return rc;
    int go = (R + 255) / 256;
    if (go > per_image_cap) go = per_image_cap;
    PLD_CUDA(launch_pdl(sel_order_kernel, dim3(dim3((unsigned)go, (unsigned)B)), dim3(256), 0, st, v0, v1, varying, n_surv, n, R, order, order_out));
    PLD_CHECK_LAUNCH();
  }
  }

  // 5. redraw the kept lists from their Philox ids: emit rankings, loss and gradient
  P.score_keys = nullptr;
  P.n = R;
  P.list_map = order; P.map_stride = (size_t)R;
  P.rank_out = rankings; P.per_list = per_list; P.grad = grad; P.loss = loss; P.loss_sum = loss_sum;
  const bool do_loss = loss != nullptr;
  if (do_loss && grad != nullptr && ctx->deterministic) {
    rc = ctx->ensure_acc(gtotal);
    if (rc) return rc;
    P.acc = ctx->d_acc;
    PLD_CUDA(cudaMemsetAsync(P.acc, 0, sizeof(long long) * gtotal, st));
  }
  ctx->time_begin(st);
  rc = (K <= 16) ? launch_lists_small(P, SRC_PHILOX_TAB, do_loss, ctx->num_sms, st)
                 : launch_lists_large(P, SRC_PHILOX_TAB, do_loss, ctx->num_sms, st);
  ctx->time_end(st);
  if (rc == PLD_OK && P.acc != nullptr) rc = launch_acc_finalize(ctx, grad, gtotal, scale, 0, st);
  if (rc == PLD_OK && ctx->use_device_offset) rc = launch_offset_advance(ctx, st);
  return rc;
}
