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
#include "pld_lists.cuh"

namespace pld {
int launch_lists_small(const ListParams& P, int src, bool loss, int num_sms, cudaStream_t st);
int launch_acc_finalize(pld_ctx* ctx, float* grad, size_t n, float scale, int accumulate, cudaStream_t st);

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

__global__ void __launch_bounds__(PC_THREADS) prep_count_kernel(const float* __restrict__ mask, int Nm, int nchunks,
                                                               int* __restrict__ counts, float4* __restrict__ grad4,
                                                               size_t grad_n4, float* __restrict__ grad_tail,
                                                               int grad_tail_n) {
  __shared__ int s_warp[PC_THREADS / 32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* m = mask + (size_t)b * Nm;
  const int base = chunk * PC_CHUNK + threadIdx.x * PC_ITEMS;
  const int c = (base < Nm) ? __popc(flags16(m, base, Nm)) : 0;
  const int tot = block_sum_int(c, s_warp);
  if (threadIdx.x == 0) counts[b * nchunks + chunk] = tot;
  if (grad4 != nullptr) {  // zero the dense gradient map (grid-stride, 16-byte stores)
    const size_t nb = (size_t)gridDim.x * gridDim.y, bid = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = bid * PC_THREADS + threadIdx.x; i < grad_n4; i += nb * PC_THREADS) grad4[i] = z;
    if (bid == 0 && (int)threadIdx.x < grad_tail_n) grad_tail[threadIdx.x] = 0.f;
  }
}

__global__ void __launch_bounds__(PC_THREADS) prep_build_kernel(
    const float* __restrict__ mask, const float* __restrict__ gt, const float* __restrict__ pred, int Nm, int Wm,
    int W, int HW, double xs, double ys, int identity_scale, int nchunks, const int* __restrict__ counts,
    float2* __restrict__ table, size_t table_stride, int32_t* __restrict__ n_valid) {
  __shared__ int s_warp[PC_THREADS / 32];
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
  if (identity_scale && total == Nm) {
    const float* s = pred + (size_t)b * HW;
    // coalesced: thread t handles elements chunk*CHUNK + i*THREADS + t
#pragma unroll 4
    for (int i = 0; i < PC_ITEMS; ++i) {
      const int j = chunk * PC_CHUNK + i * PC_THREADS + threadIdx.x;
      if (j < Nm) tab[j] = make_float2(__ldg(g + j), __ldg(s + j));
    }
    if (chunk == 0 && threadIdx.x == 0) n_valid[b] = -Nm;
    return;
  }
  // Compaction with lane-consecutive pixels: warp w of the CTA owns pixels
  // [chunk*CHUNK + w*512, +512) as 16 rows of 32; ballots give each valid pixel its rank, so mask
  // reads, gt reads and table writes are all coalesced.
  const float* m = mask + (size_t)b * Nm;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wbase = chunk * PC_CHUNK + wid * (PC_ITEMS * 32);
  uint32_t bal[PC_ITEMS];
  int wcount = 0;
#pragma unroll
  for (int i = 0; i < PC_ITEMS; ++i) {
    const int idx = wbase + i * 32 + lane;
    const bool v = (idx < Nm) && (__ldg(m + idx) > 0.f);
    bal[i] = __ballot_sync(0xffffffffu, v);
    wcount += __popc(bal[i]);
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
      tab[rank + __popc(bal[i] & lt)] = make_float2(__int_as_float(p), __ldg(g + p));
    }
    rank += __popc(bal[i]);
  }
  if (chunk == 0 && threadIdx.x == 0) n_valid[b] = total;
}

}  // namespace pld

using namespace pld;

extern "C" int pld_fused_step(pld_ctx* ctx, const float* mask, const float* gt, const float* pred, int B, int Hm,
                              int Wm, int H, int W, int K, int n, uint64_t seed, uint64_t offset, int image_base,
                              float scale, int32_t* n_valid, float* rankings, float* loss, double* loss_sum,
                              float* per_list, float* grad, void* stream) {
  PLD_REQUIRE(ctx && mask && gt && pred && loss, "null argument");
  PLD_REQUIRE(B > 0 && B <= 65535 && Hm > 0 && Wm > 0 && H > 0 && W > 0, "bad shape");
  PLD_REQUIRE((long long)H * W <= PLD_MAX_PIXELS && (long long)Hm * Wm <= PLD_MAX_PIXELS, "map too large");
  PLD_REQUIRE(K >= 1 && K <= 16, "pld_fused_step supports ranking_size 1..16 (use the staged calls above that)");
  PLD_REQUIRE(n >= 0 && (long long)B * n < (1ll << 31), "bad list count");
  PLD_REQUIRE((offset >> 48) == 0, "offset must fit in 48 bits");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W, Nm = Hm * Wm;
  const int nchunks = (Nm + PC_CHUNK - 1) / PC_CHUNK;
  const size_t tstride = (size_t)(HW > Nm ? HW : Nm);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t off_counts = 0, off_nv = al(sizeof(int) * (size_t)B * nchunks);
  const size_t off_tab = off_nv + al(sizeof(int32_t) * (size_t)B);
  int rc = ctx->ensure_scratch(off_tab + sizeof(float2) * (size_t)B * tstride);
  if (rc) return rc;
  int* counts = (int*)((char*)ctx->d_scratch + off_counts);
  int32_t* nv = n_valid ? n_valid : (int32_t*)((char*)ctx->d_scratch + off_nv);
  float2* table = (float2*)((char*)ctx->d_scratch + off_tab);
  const int per_image_cap = (ctx->num_sms * 8 + B - 1) / B;
  rc = ctx->ensure_partials(per_image_cap * B + B);
  if (rc) return rc;

  dim3 grid((unsigned)nchunks, (unsigned)B);
  const size_t gtotal = (size_t)B * HW;
  float4* g4 = nullptr;
  size_t n4 = 0;
  int tail = 0;
  if (grad != nullptr) {
    PLD_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad must be 16-byte aligned");
    g4 = reinterpret_cast<float4*>(grad);
    n4 = gtotal / 4;
    tail = (int)(gtotal - n4 * 4);
  }
  prep_count_kernel<<<grid, PC_THREADS, 0, st>>>(mask, Nm, nchunks, counts, g4, n4, grad ? grad + n4 * 4 : nullptr, tail);
  PLD_CHECK_LAUNCH();
  const double xs = (double)H / (double)Hm, ys = (double)W / (double)Wm;
  const int identity_scale = (H == Hm && W == Wm) ? 1 : 0;
  prep_build_kernel<<<grid, PC_THREADS, 0, st>>>(mask, gt, pred, Nm, Wm, W, HW, xs, ys, identity_scale, nchunks, counts,
                                                table, tstride, nv);
  PLD_CHECK_LAUNCH();
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
  P.off_lo = (uint32_t)offset; P.off_hi16 = (uint32_t)((offset >> 32) & 0xFFFFu) << 16;
  P.image_base = image_base;
  if (grad != nullptr && ctx->deterministic) {
    rc = ctx->ensure_acc(gtotal);
    if (rc) return rc;
    P.acc = ctx->d_acc;
    PLD_CUDA(cudaMemsetAsync(P.acc, 0, sizeof(long long) * gtotal, st));
  }
  ctx->time_begin(st);
  rc = launch_lists_small(P, SRC_PHILOX_TAB, true, ctx->num_sms, st);
  ctx->time_end(st);
  if (rc == PLD_OK && P.acc != nullptr) rc = launch_acc_finalize(ctx, grad, gtotal, scale, 0, st);
  return rc;
}
