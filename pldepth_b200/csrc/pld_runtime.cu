// Context, error plumbing and the auxiliary kernels of the sampler: valid-pixel table,
// MT19937 stream + NumPy masked rejection, gt min/max, candidate scores, top-R selection.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "pld_common.cuh"
#include "pld_score.cuh"

namespace pld {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return t_err; }

bool pdl_enabled() {
  static const bool on = !(getenv("PLD_NO_PDL") != nullptr && getenv("PLD_NO_PDL")[0] == '1');
  return on;
}

int lists_grid_mult() {
  static int m = 0;
  if (m == 0) {
    const char* e = getenv("PLD_GRID_MULT");
    m = e ? atoi(e) : 8;
    if (m < 1 || m > 1024) m = 8;
  }
  return m;
}

}  // namespace pld

int pld_ctx::ensure_scratch(size_t bytes) {
  if (bytes <= scratch_cap) return PLD_OK;
  if (d_scratch) {
    cudaDeviceSynchronize();  // growth is rare; make sure nothing still uses the old buffer
    cudaFree(d_scratch);
    d_scratch = nullptr;
    scratch_cap = 0;
  }
  size_t want = bytes + (bytes >> 2) + (1u << 20);
  if (cudaMalloc(&d_scratch, want) != cudaSuccess) {
    cudaGetLastError();
    pld::set_error("scratch allocation of %zu bytes failed", want);
    return PLD_ENOMEM;
  }
  scratch_cap = want;
  return PLD_OK;
}

int pld_ctx::ensure_acc(size_t elems) {
  if (elems <= acc_cap) return PLD_OK;
  if (d_acc) {
    cudaDeviceSynchronize();
    cudaFree(d_acc);
    d_acc = nullptr;
    acc_cap = 0;
  }
  if (cudaMalloc(&d_acc, sizeof(long long) * elems) != cudaSuccess) {
    cudaGetLastError();
    pld::set_error("accumulator allocation failed");
    return PLD_ENOMEM;
  }
  acc_cap = elems;
  return PLD_OK;
}

int pld_ctx::ensure_mm(int B) {
  if (B <= mm_cap) return PLD_OK;
  if (d_mm_acc) {
    cudaDeviceSynchronize();
    cudaFree(d_mm_acc);
    d_mm_acc = nullptr;
    mm_cap = 0;
  }
  if (cudaMalloc(&d_mm_acc, sizeof(unsigned int) * 2 * (size_t)B) != cudaSuccess ||
      cudaMemset(d_mm_acc, 0, sizeof(unsigned int) * 2 * (size_t)B) != cudaSuccess) {
    cudaGetLastError();
    pld::set_error("min/max accumulator allocation failed");
    return PLD_ENOMEM;
  }
  mm_cap = B;
  return PLD_OK;
}

int pld_ctx::ensure_partials(int n) {
  if (n <= partials_cap) return PLD_OK;
  if (d_partials) {
    cudaDeviceSynchronize();
    cudaFree(d_partials);
    d_partials = nullptr;
    partials_cap = 0;
  }
  int want = n < 4096 ? 4096 : n * 2;
  if (cudaMalloc(&d_partials, sizeof(double) * (size_t)want) != cudaSuccess) {
    cudaGetLastError();
    pld::set_error("partials allocation failed");
    return PLD_ENOMEM;
  }
  partials_cap = want;
  return PLD_OK;
}

namespace pld {

// ------------------------------------------------------------------------------------------
// valid-pixel table (np.where(mask > 0) + scaling), two passes over chunks of 4096 pixels
// ------------------------------------------------------------------------------------------
constexpr int MC_THREADS = 256;
constexpr int MC_ITEMS = 16;
constexpr int MC_CHUNK = MC_THREADS * MC_ITEMS;

__device__ __forceinline__ uint32_t mask_flags16(const float* __restrict__ m, int base, int Nm) {
  uint32_t f = 0;
  if (base + MC_ITEMS <= Nm && ((reinterpret_cast<uintptr_t>(m + base) & 15) == 0)) {
    const float4* v4 = reinterpret_cast<const float4*>(m + base);
#pragma unroll
    for (int q = 0; q < MC_ITEMS / 4; ++q) {
      const float4 v = __ldg(v4 + q);
      f |= (v.x > 0.f ? 1u : 0u) << (q * 4 + 0);
      f |= (v.y > 0.f ? 1u : 0u) << (q * 4 + 1);
      f |= (v.z > 0.f ? 1u : 0u) << (q * 4 + 2);
      f |= (v.w > 0.f ? 1u : 0u) << (q * 4 + 3);
    }
  } else {
#pragma unroll
    for (int i = 0; i < MC_ITEMS; ++i)
      if (base + i < Nm && __ldg(m + base + i) > 0.f) f |= 1u << i;
  }
  return f;
}

__global__ void __launch_bounds__(MC_THREADS) mask_count_kernel(const float* __restrict__ mask, int Nm,
                                                               int nchunks, int* __restrict__ counts) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* m = mask + (size_t)b * Nm;
  const int base = chunk * MC_CHUNK + threadIdx.x * MC_ITEMS;
  int c = (base < Nm) ? __popc(mask_flags16(m, base, Nm)) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int s[MC_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < MC_THREADS / 32; ++i) t += s[i];
    counts[b * nchunks + chunk] = t;
  }
}

__global__ void __launch_bounds__(MC_THREADS) mask_scatter_kernel(
    const float* __restrict__ mask, int Nm, int Wm, int W, double xs, double ys, int identity,
    int nchunks, const int* __restrict__ counts, int32_t* __restrict__ valid_flat,
    int32_t* __restrict__ n_valid) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* m = mask + (size_t)b * Nm;
  __shared__ int s_warp[MC_THREADS / 32];
  __shared__ int s_prefix;
  // prefix over preceding chunks of this image
  int pre = 0;
  for (int i = threadIdx.x; i < chunk; i += MC_THREADS) pre += counts[b * nchunks + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = pre;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < MC_THREADS / 32; ++i) t += s_warp[i];
    s_prefix = t;
  }
  __syncthreads();
  if (identity) {
    // full mask at image resolution: the table would be valid_flat[j] == j; it is not written
    // and the count is stored negated so consumers skip the lookup (include/pldepth_b200.h)
    int tot = 0;
    for (int i = threadIdx.x; i < nchunks; i += MC_THREADS) tot += counts[b * nchunks + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    __shared__ int s_tot[MC_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s_tot[threadIdx.x >> 5] = tot;
    __syncthreads();
    tot = 0;
    for (int i = 0; i < MC_THREADS / 32; ++i) tot += s_tot[i];
    if (tot == Nm) {
      if (chunk == 0 && threadIdx.x == 0) n_valid[b] = -Nm;
      return;
    }
  }
  // lane-consecutive pixels + ballots: coalesced mask reads and table writes (see prep_build_kernel)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wbase = chunk * MC_CHUNK + wid * (MC_ITEMS * 32);
  uint32_t bal[MC_ITEMS];
  int wcount = 0;
#pragma unroll
  for (int i = 0; i < MC_ITEMS; ++i) {
    const int idx = wbase + i * 32 + lane;
    const bool v = (idx < Nm) && (__ldg(m + idx) > 0.f);
    bal[i] = __ballot_sync(0xffffffffu, v);
    wcount += __popc(bal[i]);
  }
  __syncthreads();
  if (lane == 0) s_warp[wid] = wcount;
  __syncthreads();
  int rank = s_prefix;
  for (int i = 0; i < wid; ++i) rank += s_warp[i];
  const uint32_t lt = (1u << lane) - 1u;
  int32_t* out = valid_flat + (size_t)b * Nm;
#pragma unroll
  for (int i = 0; i < MC_ITEMS; ++i) {
    if ((bal[i] >> lane) & 1u) {
      const int idx = wbase + i * 32 + lane;
      int p;
      if (identity) {
        p = idx;
      } else {
        const int rm = idx / Wm, cm = idx - rm * Wm;
        p = (int)((double)rm * xs) * W + (int)((double)cm * ys);
      }
      out[rank + __popc(bal[i] & lt)] = p;
    }
    rank += __popc(bal[i]);
  }
  if (chunk == nchunks - 1 && threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < MC_THREADS / 32; ++i) t += s_warp[i];
    n_valid[b] = s_prefix + t;
  }
}

// ------------------------------------------------------------------------------------------
// MT19937
// ------------------------------------------------------------------------------------------
__global__ void mt_init_kernel(uint32_t seed, uint32_t* state, int32_t* pos) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t x = seed;
    state[0] = x;
    for (uint32_t i = 1; i < 624; ++i) {
      x = 1812433253u * (x ^ (x >> 30)) + i;
      state[i] = x;
    }
    *pos = 624;
  }
}

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}

__global__ void __launch_bounds__(256) mt_generate_kernel(uint32_t* state, int32_t* pos,
                                                          uint32_t* __restrict__ out, long long n) {
  __shared__ uint32_t mt[624];
  const int t = threadIdx.x;
  for (int i = t; i < 624; i += 256) mt[i] = state[i];
  int p = *pos;
  __syncthreads();
  long long produced = 0;
  while (produced < n) {
    if (p >= 624) {
      uint32_t v = 0;
      if (t < 227) v = mt_twist(mt[t], mt[t + 1], mt[t + 397]);
      __syncthreads();
      if (t < 227) mt[t] = v;
      __syncthreads();
      if (t < 227) v = mt_twist(mt[227 + t], mt[228 + t], mt[t]);
      __syncthreads();
      if (t < 227) mt[227 + t] = v;
      __syncthreads();
      if (t < 169) v = mt_twist(mt[454 + t], mt[455 + t], mt[227 + t]);
      __syncthreads();
      if (t < 169) mt[454 + t] = v;
      __syncthreads();
      if (t == 0) mt[623] = mt_twist(mt[623], mt[0], mt[396]);
      __syncthreads();
      p = 0;
    }
    const long long rem = n - produced;
    const int take = (int)((rem < (long long)(624 - p)) ? rem : (long long)(624 - p));
    for (int i = t; i < take; i += 256) out[produced + i] = mt_temper(mt[p + i]);
    produced += take;
    p += take;
  }
  __syncthreads();
  for (int i = t; i < 624; i += 256) state[i] = mt[i];
  if (t == 0) *pos = p;
}

// NumPy masked rejection over a window of the raw stream, image by image.
constexpr int MTC_THREADS = 256;
constexpr int MTC_ITEMS = 8;
constexpr int MTC_CHUNK = MTC_THREADS * MTC_ITEMS;

__device__ __forceinline__ uint32_t np_mask(uint32_t M) {  // smallest 2^k - 1 >= M - 1
  const uint32_t r = M - 1u;
  return r == 0 ? 0u : (0xFFFFFFFFu >> __clz(r));
}

__global__ void __launch_bounds__(MTC_THREADS) mt_count_kernel(const uint32_t* __restrict__ raw,
                                                               long long n_raw,
                                                               const long long* __restrict__ cons,
                                                               const int32_t* __restrict__ n_valid,
                                                               int b, int* __restrict__ counts) {
  const uint32_t M = (uint32_t)abs(n_valid[b]);
  int c = 0;
  if ((int)M > 1) {
    const uint32_t msk = np_mask(M);
    const long long start = cons[b];
    const long long base = start + (long long)blockIdx.x * MTC_CHUNK;
#pragma unroll
    for (int i = 0; i < MTC_ITEMS; ++i) {
      const long long idx = base + i * MTC_THREADS + threadIdx.x;
      if (idx < n_raw) c += ((__ldg(raw + idx) & msk) <= M - 1u) ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int s[MTC_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < MTC_THREADS / 32; ++i) t += s[i];
    counts[blockIdx.x] = t;
  }
}

// single block: exclusive scan of counts[nblk] in place; flags exhaustion
__global__ void __launch_bounds__(1024) mt_scan_kernel(int* counts, int nblk, int need,
                                                       const int32_t* __restrict__ n_valid, int b,
                                                       long long* cons, long long n_raw, int* status) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int c = (i < nblk) ? counts[i] : 0;
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
    if (i < nblk) counts[i] = carry + woff + inc - c;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + woff + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int M = abs(n_valid[b]);
    if (M > 1 && s_carry < need) atomicOr(status, PLD_ST_MT_EXHAUSTED);
    if (M <= 0) atomicOr(status, PLD_ST_EMPTY_MASK);
    cons[b + 1] = n_raw;  // "exhausted" marker; the scatter pass overwrites it on success
  }
}

__global__ void __launch_bounds__(MTC_THREADS) mt_scatter_kernel(
    const uint32_t* __restrict__ raw, long long n_raw, long long* __restrict__ cons,
    const int32_t* __restrict__ n_valid, int b, const int* __restrict__ offsets, int need,
    int32_t* __restrict__ sel) {
  const uint32_t M = (uint32_t)abs(n_valid[b]);
  const long long start = cons[b];
  if ((int)M <= 1) {  // randint(1) consumes no word and returns 0
    for (int i = blockIdx.x * MTC_THREADS + threadIdx.x; i < need; i += gridDim.x * MTC_THREADS) sel[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) cons[b + 1] = start;
    return;
  }
  const uint32_t msk = np_mask(M);
  const long long base = start + (long long)blockIdx.x * MTC_CHUNK;
  // blocked arrangement so ranks follow stream order: thread t owns words [t*ITEMS, t*ITEMS+ITEMS)
  uint32_t v[MTC_ITEMS];
  uint32_t f = 0;
#pragma unroll
  for (int i = 0; i < MTC_ITEMS; ++i) {
    const long long idx = base + (long long)threadIdx.x * MTC_ITEMS + i;
    v[i] = 0;
    if (idx < n_raw) {
      v[i] = __ldg(raw + idx) & msk;
      if (v[i] <= M - 1u) f |= 1u << i;
    }
  }
  const int c = __popc(f);
  int inc = c;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __shared__ int s_warp[MTC_THREADS / 32];
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < wid; ++w) woff += s_warp[w];
  int rank = offsets[blockIdx.x] + woff + inc - c;
#pragma unroll
  for (int i = 0; i < MTC_ITEMS; ++i) {
    if ((f >> i) & 1u) {
      if (rank < need) sel[rank] = (int32_t)v[i];
      if (rank == need - 1) cons[b + 1] = base + (long long)threadIdx.x * MTC_ITEMS + i + 1;
      ++rank;
    }
  }
}

__global__ void copy_i64_kernel(const long long* src, long long* dst) { *dst = *src; }
__global__ void offset_advance_kernel(unsigned long long* counter) { *counter += 1ull; }
int launch_offset_advance(pld_ctx* ctx, cudaStream_t st) {
  offset_advance_kernel<<<1, 1, 0, st>>>(ctx->d_offset);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

// deterministic mode: fixed-point accumulators -> float gradient (grad = [grad +] acc * 2^-32 * scale)
__global__ void __launch_bounds__(256) acc_finalize_kernel(const long long* __restrict__ acc, float* __restrict__ grad,
                                                           size_t n, float scale, int accumulate) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float v = (float)((double)acc[i] * (1.0 / 4294967296.0) * (double)scale);
    grad[i] = accumulate ? grad[i] + v : v;
  }
}
int launch_acc_finalize(pld_ctx* ctx, float* grad, size_t n, float scale, int accumulate, cudaStream_t st) {
  int gx = (int)((n + 255) / 256);
  if (gx > ctx->num_sms * 16) gx = ctx->num_sms * 16;
  acc_finalize_kernel<<<gx, 256, 0, st>>>(ctx->d_acc, grad, n, scale, accumulate);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

// ------------------------------------------------------------------------------------------
// gt min / max per image
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) gt_minmax_kernel(const float* __restrict__ gt, int HW,
                                                         float* __restrict__ out) {
  const float* g = gt + (size_t)blockIdx.x * HW;
  float mn = 3.402823466e38f, mx = -3.402823466e38f;
  for (int i = threadIdx.x; i < HW; i += 1024) {
    const float v = __ldg(g + i);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[32], smx[32];
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 32; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    out[blockIdx.x * 2 + 0] = mn;
    out[blockIdx.x * 2 + 1] = mx;
  }
}

// ------------------------------------------------------------------------------------------
// candidate scores, staged API (pld_score.cuh holds the arithmetic)
// ------------------------------------------------------------------------------------------
struct ScoreParams {
  const float* rankings;  // [B, n, K, 2]
  double* scores;         // [B, n]
  int B, n, K;
  ScoreCfg cfg;
};

template <typename T>
__global__ void __launch_bounds__(256) score_kernel(const ScoreParams P) {
  const int b = blockIdx.y;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < P.n; l += gridDim.x * blockDim.x) {
    const size_t list_id = (size_t)b * P.n + l;
    const float2* r = reinterpret_cast<const float2*>(P.rankings) + list_id * P.K;
    auto g = [r](int k) { return __ldg(&r[k].y); };
    P.scores[list_id] = score_list<T>(g, P.K, P.cfg, b);
  }
}

// ------------------------------------------------------------------------------------------
// top-R selection (staged API): per-image ascending stable radix sort of the ordered scores
// (pld_sort.cu); the kept lists are the last R of each image's run, read backwards.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_keys_kernel(const double* __restrict__ scores, int n,
                                                          uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const size_t off = (size_t)blockIdx.y * n;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    keys[off + i] = score_key(scores[off + i]);
    vals[off + i] = (uint32_t)i;
  }
}
__global__ void __launch_bounds__(256) select_gather_kernel(const uint32_t* __restrict__ vals,
                                                            const float* __restrict__ rankings, int n,
                                                            int K, int R, float* __restrict__ out,
                                                            int32_t* __restrict__ order_out, int direct) {
  // one warp per kept list: copies K float2.  vals: sorted candidate ids of the image, best LAST (stride n), or
  // (direct) the kept ids in output order (stride R)
  const int b = blockIdx.y;
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int j = blockIdx.x * warps + (threadIdx.x >> 5); j < R; j += gridDim.x * warps) {
    const uint32_t src = direct ? vals[(size_t)b * R + j] : vals[(size_t)b * n + (size_t)(n - 1 - j)];
    const float2* s = reinterpret_cast<const float2*>(rankings) + ((size_t)b * n + src) * K;
    float2* d = reinterpret_cast<float2*>(out) + ((size_t)b * R + j) * K;
    for (int k = lane; k < K; k += 32) d[k] = __ldg(s + k);
    if (order_out != nullptr && lane == 0) order_out[(size_t)b * R + j] = (int32_t)src;
  }
}

int seg_radix_sort(pld_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                   const int* len_dev, int len_max, size_t stride, int B, int* hist,
                   const unsigned long long* varying, int first_pass, cudaStream_t st);
size_t seg_radix_sort_hist_bytes(int len_max, int B);
bool select_small_fits(int n);
int select_small_init();
int pilot_select_init();
int select_small(const uint64_t* keys, const double* scores, int n, size_t stride, int B, int R, bool ascending_ids,
                 uint32_t* order, int32_t* order_out, cudaStream_t st);

}  // namespace pld

using namespace pld;

// ==========================================================================================
// C ABI (context + auxiliary stages)
// ==========================================================================================
extern "C" {

int pld_version(void) { return 100; }
const char* pld_last_error(void) { return pld::last_error(); }
uint64_t pld_launch_count(void) { return pld::g_launches.load(); }

static int ctx_init(pld_ctx* c, int device) {
  c->device = device;
  int num_sms = 0;
  PLD_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  c->num_sms = num_sms;
  PLD_CUDA(cudaMalloc(&c->d_status, sizeof(int)));
  PLD_CUDA(cudaMalloc(&c->d_ticket, sizeof(unsigned int)));
  PLD_CUDA(cudaMemset(c->d_status, 0, sizeof(int)));
  PLD_CUDA(cudaMemset(c->d_ticket, 0, sizeof(unsigned int)));
  int rc = c->ensure_partials(4096);
  if (rc) return rc;
  rc = select_small_init();
  if (rc) return rc;
  rc = pilot_select_init();
  if (rc) return rc;
  PLD_CUDA(cudaDeviceSynchronize());
  return PLD_OK;
}

int pld_ctx_create(int device, pld_ctx** out) {
  PLD_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  int ndev = 0, prev = 0;
  PLD_CUDA(cudaGetDeviceCount(&ndev));
  PLD_REQUIRE(device >= 0 && device < ndev, "device out of range");
  PLD_CUDA(cudaGetDevice(&prev));
  PLD_CUDA(cudaSetDevice(device));
  pld_ctx* c = new (std::nothrow) pld_ctx();
  if (!c) { set_error("out of host memory"); cudaSetDevice(prev); return PLD_ENOMEM; }
  memset(c, 0, sizeof(*c));
  const int rc = ctx_init(c, device);
  if (rc != PLD_OK) {  // release whatever was allocated; the error message of the failing call is kept
    cudaFree(c->d_status);
    cudaFree(c->d_ticket);
    cudaFree(c->d_partials);
    delete c;
  } else {
    *out = c;
  }
  cudaSetDevice(prev);  // creating a context does not change the caller's current device
  return rc;
}

int pld_ctx_destroy(pld_ctx* ctx) {
  if (!ctx) return PLD_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  cudaFree(ctx->d_status);
  cudaFree(ctx->d_ticket);
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_scratch);
  cudaFree(ctx->d_acc);
  cudaFree(ctx->d_mm_acc);
  cudaFree(ctx->d_offset);
  pld_ctx_kernel_timing(ctx, 0);
  delete ctx;
  cudaSetDevice(prev);
  return PLD_OK;
}

int pld_ctx_device_offset(pld_ctx* ctx, int enable, uint64_t start) {
  PLD_REQUIRE(ctx != nullptr, "null context");
  PLD_REQUIRE((start >> 48) == 0, "offset must fit in 48 bits");
  if (enable) {
    if (!ctx->d_offset) PLD_CUDA(cudaMalloc(&ctx->d_offset, sizeof(unsigned long long)));
    unsigned long long v = start;
    PLD_CUDA(cudaMemcpy(ctx->d_offset, &v, sizeof(v), cudaMemcpyHostToDevice));
  }
  ctx->use_device_offset = enable ? 1 : 0;
  return PLD_OK;
}

int pld_ctx_set_deterministic(pld_ctx* ctx, int on) {
  PLD_REQUIRE(ctx != nullptr, "null context");
  ctx->deterministic = on ? 1 : 0;
  return PLD_OK;
}

int pld_ctx_kernel_timing(pld_ctx* ctx, int slots) {
  PLD_REQUIRE(ctx && slots >= 0 && slots <= 4096, "bad argument");
  for (int i = 0; i < ctx->ev_cap; ++i) { cudaEventDestroy(ctx->ev_start[i]); cudaEventDestroy(ctx->ev_stop[i]); }
  delete[] ctx->ev_start; delete[] ctx->ev_stop;
  ctx->ev_start = ctx->ev_stop = nullptr;
  ctx->ev_cap = ctx->ev_count = 0;
  if (slots > 0) {
    ctx->ev_start = new cudaEvent_t[slots];
    ctx->ev_stop = new cudaEvent_t[slots];
    for (int i = 0; i < slots; ++i) { PLD_CUDA(cudaEventCreate(&ctx->ev_start[i])); PLD_CUDA(cudaEventCreate(&ctx->ev_stop[i])); }
    ctx->ev_cap = slots;
  }
  return PLD_OK;
}

int pld_ctx_kernel_times(pld_ctx* ctx, float* ms_host, int capacity, int* count_host) {
  PLD_REQUIRE(ctx && ms_host && count_host && capacity >= 0, "bad argument");
  int n = ctx->ev_count < capacity ? ctx->ev_count : capacity;
  for (int i = 0; i < n; ++i) {
    PLD_CUDA(cudaEventSynchronize(ctx->ev_stop[i]));
    PLD_CUDA(cudaEventElapsedTime(&ms_host[i], ctx->ev_start[i], ctx->ev_stop[i]));
  }
  *count_host = n;
  ctx->ev_count = 0;
  return PLD_OK;
}

int pld_ctx_status(pld_ctx* ctx, void* stream, int* status_host) {
  PLD_REQUIRE(ctx && status_host, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int h = 0;
  PLD_CUDA(cudaMemcpyAsync(&h, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
  PLD_CUDA(cudaStreamSynchronize(st));
  if (h != 0) PLD_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), st));
  *status_host = h;
  return PLD_OK;
}

int pld_mask_compact(pld_ctx* ctx, const float* mask, int B, int Hm, int Wm, int H, int W,
                     int32_t* valid_flat, int32_t* n_valid, void* stream) {
  PLD_REQUIRE(ctx && mask && valid_flat && n_valid, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(B > 0 && B <= 65535 && Hm > 0 && Wm > 0 && H > 0 && W > 0, "bad shape");
  PLD_REQUIRE((long long)H * W <= PLD_MAX_PIXELS, "H*W exceeds PLD_MAX_PIXELS");
  PLD_REQUIRE((long long)Hm * Wm <= PLD_MAX_PIXELS, "Hm*Wm exceeds PLD_MAX_PIXELS");
  cudaStream_t st = (cudaStream_t)stream;
  const int Nm = Hm * Wm;
  const int nchunks = (Nm + MC_CHUNK - 1) / MC_CHUNK;
  int rc = ctx->ensure_scratch(sizeof(int) * (size_t)B * nchunks);
  if (rc) return rc;
  int* counts = (int*)ctx->d_scratch;
  const double xs = (double)H / (double)Hm, ys = (double)W / (double)Wm;  // sampling.py:126-127
  const int identity = (H == Hm && W == Wm) ? 1 : 0;
  dim3 grid((unsigned)nchunks, (unsigned)B);
  mask_count_kernel<<<grid, MC_THREADS, 0, st>>>(mask, Nm, nchunks, counts);
  PLD_CHECK_LAUNCH();
  mask_scatter_kernel<<<grid, MC_THREADS, 0, st>>>(mask, Nm, Wm, W, xs, ys, identity, nchunks, counts,
                                                  valid_flat, n_valid);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_mt19937_init(pld_ctx* ctx, uint32_t seed, uint32_t* state, int32_t* pos, void* stream) {
  PLD_REQUIRE(ctx && state && pos, "null argument");
  PLD_CHECK_DEVICE(ctx);
  mt_init_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(seed, state, pos);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_mt19937_generate(pld_ctx* ctx, uint32_t* state, int32_t* pos, uint32_t* out, int64_t n,
                         void* stream) {
  PLD_REQUIRE(ctx && state && pos && (out || n == 0), "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(n >= 0, "negative count");
  if (n == 0) return PLD_OK;
  mt_generate_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(state, pos, out, (long long)n);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_gt_minmax(pld_ctx* ctx, const float* gt, int B, int HW, float* gt_minmax, void* stream) {
  PLD_REQUIRE(ctx && gt && gt_minmax, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(B > 0 && HW > 0, "bad shape");
  gt_minmax_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(gt, HW, gt_minmax);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_score_lists(pld_ctx* ctx, const float* rankings, const float* gt_minmax, int B, int n, int K,
                    int strategy, double threshold, double equality_penalty, int promotion,
                    double* scores, void* stream) {
  PLD_REQUIRE(ctx && rankings && scores, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(B > 0 && B <= 65535 && n > 0 && K >= 1 && K <= PLD_MAX_RANKING_SIZE, "bad shape");
  PLD_REQUIRE(strategy >= PLD_STRATEGY_MASKED && strategy <= PLD_STRATEGY_INFORMATION, "bad strategy");
  PLD_REQUIRE(strategy != PLD_STRATEGY_INFORMATION || gt_minmax != nullptr, "gt_minmax required");
  PLD_REQUIRE(promotion == PLD_PROMOTION_NEP50 || promotion == PLD_PROMOTION_LEGACY, "bad promotion");
  ScoreParams P;
  P.rankings = rankings; P.scores = scores; P.B = B; P.n = n; P.K = K;
  P.cfg = make_score_cfg(gt_minmax, strategy, threshold, equality_penalty, promotion);
  int gx = (n + 255) / 256;
  const int cap = (ctx->num_sms * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)B);
  if (promotion == PLD_PROMOTION_NEP50) score_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  else score_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int pld_select_top(pld_ctx* ctx, const double* scores, const float* rankings, int B, int n, int K, int R,
                   float* rankings_out, int32_t* order_out, void* stream) {
  PLD_REQUIRE(ctx && scores && rankings && rankings_out, "null argument");
  PLD_CHECK_DEVICE(ctx);
  PLD_REQUIRE(B > 0 && B <= 65535 && n > 0 && K >= 1 && R >= 0 && R <= n, "bad shape");
  PLD_REQUIRE((long long)B * n < (1ll << 31), "too many candidates");
  if (R == 0) return PLD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * n;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t need = al(total * 8) * 2 + al(total * 4) * 2 + al(seg_radix_sort_hist_bytes(n, B));
  int rc = ctx->ensure_scratch(need);
  if (rc) return rc;
  char* base = (char*)ctx->d_scratch;
  uint64_t* k0 = (uint64_t*)base; base += al(total * 8);
  uint64_t* k1 = (uint64_t*)base; base += al(total * 8);
  uint32_t* v0 = (uint32_t*)base; base += al(total * 4);
  uint32_t* v1 = (uint32_t*)base; base += al(total * 4);
  int* hist = (int*)base;
  const int cap = (ctx->num_sms * 8 + B - 1) / B;
  int gxr = (R + 7) / 8;
  if (gxr > cap) gxr = cap;
  if (select_small_fits(n)) {
    // the sizes the reference runs (R = 100 ... 1000 per image): one shared-memory sort per image
    uint32_t* order = v0;
    rc = select_small(nullptr, scores, n, (size_t)n, B, R, false, order, nullptr, st);
    if (rc) return rc;
    select_gather_kernel<<<dim3((unsigned)gxr, (unsigned)B), 256, 0, st>>>(order, rankings, n, K, R, rankings_out, order_out, 1);
    PLD_CHECK_LAUNCH();
    return PLD_OK;
  }
  int gx = (n + 255) / 256;
  if (gx > cap) gx = cap;
  select_keys_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, st>>>(scores, n, k0, v0);
  PLD_CHECK_LAUNCH();
  rc = seg_radix_sort(ctx, k0, v0, k1, v1, nullptr, n, (size_t)n, B, hist, nullptr, 0, st);
  if (rc) return rc;
  select_gather_kernel<<<dim3((unsigned)gxr, (unsigned)B), 256, 0, st>>>(v0, rankings, n, K, R, rankings_out, order_out, 0);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

}  // extern "C"

// MT-stream compaction driver, used by pld_sample_lists_mt (pld_api.cu)
namespace pld {
int mt_compact_images(pld_ctx* ctx, const int32_t* n_valid, int B, int need_per_image, const uint32_t* raw,
                      int64_t n_raw, int64_t* consumed_io, int32_t* sel_out, cudaStream_t st) {
  // window of the stream scanned per image: accept probability is > 1/2, so 2.5x + slack
  long long window = (long long)need_per_image * 5 / 2 + 8192;
  if (window > n_raw) window = n_raw;
  if (window < 1) window = 1;
  const int nblk = (int)((window + MTC_CHUNK - 1) / MTC_CHUNK);
  const size_t cons_bytes = ((sizeof(long long) * (size_t)(B + 1)) + 255) & ~(size_t)255;
  int rc = ctx->ensure_scratch(cons_bytes + sizeof(int) * (size_t)nblk);
  if (rc) return rc;
  long long* cons = (long long*)ctx->d_scratch;
  int* counts = (int*)((char*)ctx->d_scratch + cons_bytes);
  copy_i64_kernel<<<1, 1, 0, st>>>((const long long*)consumed_io, cons);
  PLD_CHECK_LAUNCH();
  for (int b = 0; b < B; ++b) {
    mt_count_kernel<<<nblk, MTC_THREADS, 0, st>>>(raw, (long long)n_raw, cons, n_valid, b, counts);
    PLD_CHECK_LAUNCH();
    mt_scan_kernel<<<1, 1024, 0, st>>>(counts, nblk, need_per_image, n_valid, b, cons, (long long)n_raw,
                                       ctx->d_status);
    PLD_CHECK_LAUNCH();
    mt_scatter_kernel<<<nblk, MTC_THREADS, 0, st>>>(raw, (long long)n_raw, cons, n_valid, b, counts,
                                                   need_per_image, sel_out + (size_t)b * need_per_image);
    PLD_CHECK_LAUNCH();
  }
  copy_i64_kernel<<<1, 1, 0, st>>>(cons + B, (long long*)consumed_io);
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}
}  // namespace pld
