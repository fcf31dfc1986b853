// Thread-per-list kernels for ranking_size K <= 16: draw / order / emit / ListMLE fwd+bwd.
// One thread owns one list; all K entries live in registers.
#include "pld_lists.cuh"

namespace pld {

#ifndef PLD_SMALL_MINBLOCKS
#define PLD_SMALL_MINBLOCKS 2
#endif
template <int K, int SRC, bool LOSS, bool SCORE = false>
__global__ void __launch_bounds__(256, PLD_SMALL_MINBLOCKS) lists_small_kernel(const ListParams P) {
  // per-warp staging so the emitted rankings leave as fully coalesced 256-byte rows
#ifdef PLD_SMALL_RK_ALL
  constexpr bool SMALL_RK = true;
#else
  constexpr bool SMALL_RK = SCORE;             // see draw_philox
#endif
  constexpr int STRIDE = (K & 1) ? K : K + 1;  // float2 units; odd => conflict-free 8-byte writes
  // Emitted rankings leave through the TMA engine: the warp's 32 rows lie contiguously in its staging buffer
  // (32 * K * 8 bytes; the lanes' 8-byte writes are conflict-free for odd K, two-way for K = 2 mod 4) and ONE
  // cp.async.bulk shared -> global per warp and iteration writes them, instead of K shared-memory loads + K 8-byte
  // global stores per lane through the L1TEX path that the gathers and reductions of this kernel saturate.  Two
  // buffers per warp while they fit (K <= 9).  K = 0 mod 4 keeps the padded rows + plain stores (4- to 16-way conflicts).
  constexpr bool TMA_ROWS = ((K & 3) != 0) && SRC != SRC_FED_RANK && !SCORE;
  constexpr int STRIDE_E = TMA_ROWS ? K : STRIDE;       // row stride of the emit path
  constexpr int NBUF = (TMA_ROWS && K <= 9) ? 2 : 1;
  __shared__ __align__(128) float2 s_stage[8 * 32 * STRIDE * NBUF];
  int tma_buf = 0;
  bool tma_pending = false;
  // scoring pass: the first histogram of the radix top-R selection (top 12 key bits) is taken on the fly
  __shared__ unsigned int s_hist[SCORE ? 4096 : 1];
  __shared__ double s_lad[SCORE ? 16 : 1];     // information strategy: the image's ladder of expected depths
  pdl_sync();
  const int b = blockIdx.y;
  if (SCORE) {
    if (P.score_cfg.strategy == PLD_STRATEGY_INFORMATION && threadIdx.x < K) {
      if (P.score_cfg.promotion == PLD_PROMOTION_NEP50)
        fill_ladder<float>(P.score_cfg, b, K, (int)threadIdx.x, reinterpret_cast<float*>(s_lad));
      else fill_ladder<double>(P.score_cfg, b, K, (int)threadIdx.x, s_lad);
    }
    if (P.sel_hist != nullptr)
      for (int i = threadIdx.x; i < 4096; i += 256) s_hist[i] = 0u;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t map_off = (size_t)b * (size_t)P.HW;
  const float* __restrict__ gt = P.gt + map_off;
  const float* __restrict__ pred = P.pred + map_off;
  float local = 0.f;
  int bad = 0;
  uint32_t off_lo, off_hi16;
  launch_offset(P, off_lo, off_hi16);

  uint32_t M = 1, thresh = 0;
  bool identity = false;
  const int32_t* __restrict__ vflat = nullptr;
  if (SRC != SRC_FED_RANK) {
    const int mraw = P.n_valid[b];
    const int m = mraw < 0 ? -mraw : mraw;
    if (m == 0) { bad |= PLD_ST_EMPTY_MASK; M = 0; }
    else { M = (uint32_t)m; thresh = (0u - M) % M; }
    identity = mraw < 0;  // full mask at image resolution: valid_flat[j] == j, row not materialised
    vflat = P.valid_flat + (size_t)b * (size_t)P.valid_stride;
  }

  // table layout / gradient destination of this image (see ListParams::grad_valid)
  const bool vj = (SRC == SRC_PHILOX_TAB) && !identity && P.grad_valid != nullptr;
  const bool gs_layout = identity || vj;
  float* grad_dst = P.grad == nullptr ? nullptr
                    : (vj ? P.grad_valid + (size_t)b * P.table_stride : P.grad + map_off);

  // Philox list id of the l-th list of this image: l itself, or the l-th best candidate of a
  // previous scoring pass
  const uint32_t* __restrict__ lmap = nullptr;
  if (SRC == SRC_PHILOX_TAB && P.list_map != nullptr) lmap = P.list_map + (size_t)b * P.map_stride;
  auto philox_list = [&](int l) -> int { return lmap ? (int)__ldg(lmap + l) : l; };

  // software pipeline of the table path: the draws and table gathers of the NEXT list are issued
  // before the current list is ordered / scored, so their latency hides behind that work
  int pre_sel[K];
  float2 pre_t[K];
  const float2* __restrict__ tab = nullptr;
  if (SRC == SRC_PHILOX_TAB && M != 0) {
    tab = P.table + (size_t)b * P.table_stride;
    const int l0 = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x * 256 < P.n) {
      draw_philox<K, SMALL_RK>(P, off_lo, off_hi16, b, philox_list(l0 < P.n ? l0 : P.n - 1), M, thresh, pre_sel);
#pragma unroll
      for (int k = 0; k < K; ++k) pre_t[k] = __ldg(tab + pre_sel[k]);
    }
  }

  if (M != 0) {
    for (int base = blockIdx.x * 256; base < P.n; base += gridDim.x * 256) {
      const int l = base + threadIdx.x;
      const bool active = l < P.n;
      const size_t list_id = (size_t)b * (size_t)P.n + (size_t)(active ? l : P.n - 1);
      int p[K];
      float lab[K];
      uint32_t inval = 0;  // bit k: sorted entry k has an invalid (negative) label
      float s_tab[K];      // predictions delivered by the (gt, pred) table, sorted order
      bool have_s = false;

      if (SRC == SRC_FED_RANK) {
        // the warp's 32 rows are contiguous in memory: read them as full 256-byte rows into the staging
        // buffer, then every thread picks up its own list
        float2* st = s_stage + wid * (32 * STRIDE);
        {
          const int warp_first = base + wid * 32;
          int cnt = P.n - warp_first;
          cnt = cnt > 32 ? 32 : cnt;
          if (cnt > 0) {
            const float2* __restrict__ rin = reinterpret_cast<const float2*>(P.rank_in) +
                                             ((size_t)b * (size_t)P.n + (size_t)warp_first) * K;
            const int total = cnt * K;
#pragma unroll
            for (int i = 0; i < K; ++i) {
              const int e = i * 32 + lane;
              if (e < total) {
                const int li = e / K, kk = e - li * K;
                st[li * STRIDE + kk] = __ldg(rin + e);
              }
            }
          }
          __syncwarp();
        }
        bool sorted = true, valid_all = true;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          // lanes past the end of the image compute on a dummy list (index 0, label 0); nothing of theirs is stored
          const float2 v = active ? st[lane * STRIDE + k] : make_float2(0.f, 0.f);
          int q = (int)v.x;  // tf.cast(point_coords, int32): truncation (depth_utils.py:50)
          if (q < 0 || q >= P.HW) { bad |= PLD_ST_BAD_INDEX; q = 0; }
          p[k] = q;
          lab[k] = v.y;
          valid_all = valid_all && (v.y >= 0.f);
        }
        __syncwarp();
#pragma unroll
        for (int k = 1; k < K; ++k) sorted = sorted && (lab[k - 1] >= lab[k]);
        if (!(sorted && valid_all)) {
          // TF-Ranking ordering: valid labels descending, invalid ones last with key
          // min(labels') - 1e-6; ties keep the earlier position first (stable).
          float mn = 3.402823466e38f;
#pragma unroll
          for (int k = 0; k < K; ++k) mn = fminf(mn, lab[k] >= 0.f ? lab[k] : 0.f);
          const float inv_key = mn - 1e-6f;
          uint64_t key[K];
          uint32_t pay[K];
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const bool v = lab[k] >= 0.f;
            const float kv = v ? lab[k] : inv_key;
            key[k] = ((uint64_t)float_to_ordered(kv) << 32) | (uint32_t)(0xFFFF - k);
            pay[k] = (uint32_t)p[k] | (v ? 0u : 0x80000000u);
          }
          sort_desc_regs<K, true>(key, pay);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            p[k] = (int)(pay[k] & 0x7FFFFFFFu);
            if (pay[k] >> 31) inval |= (1u << k);
          }
        }
      } else {
        uint64_t key[K];
        uint32_t nopay[K];
        int sel[K];
        float2 t[K];
        if (SRC == SRC_PHILOX_TAB) {
#pragma unroll
          for (int k = 0; k < K; ++k) { sel[k] = pre_sel[k]; t[k] = pre_t[k]; }
          const int nbase = base + gridDim.x * 256;
          if (nbase < P.n) {
            const int ln = nbase + threadIdx.x;
            draw_philox<K, SMALL_RK>(P, off_lo, off_hi16, b, philox_list(ln < P.n ? ln : P.n - 1), M, thresh, pre_sel);
#pragma unroll
            for (int k = 0; k < K; ++k) pre_t[k] = __ldg(tab + pre_sel[k]);
          }
        } else if (SRC == SRC_PHILOX) {
          draw_philox<K, SMALL_RK>(P, off_lo, off_hi16, b, active ? l : P.n - 1, M, thresh, sel);
        } else {
          const int32_t* __restrict__ sin = P.sel_in + list_id * K;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            int s = __ldg(sin + k);
            if (s < 0 || (uint32_t)s >= M) { bad |= PLD_ST_BAD_INDEX; s = 0; }
            sel[k] = s;
          }
        }
        if (P.sel_out != nullptr && (SRC == SRC_PHILOX || SRC == SRC_PHILOX_TAB) && active) {
          int32_t* so = P.sel_out + list_id * K;
#pragma unroll
          for (int k = 0; k < K; ++k) so[k] = sel[k];
        }
        if (SCORE && SRC == SRC_PHILOX_TAB) {
          // scoring pass: only the ordered depths matter (no pixel index, no prediction)
          float gs[K];
#pragma unroll
          for (int k = 0; k < K; ++k) gs[k] = gs_layout ? t[k].x : t[k].y;
          sort_desc_floats<K>(gs);
#pragma unroll
          for (int k = 0; k < K; ++k) key[k] = (uint64_t)float_to_ordered(gs[k]) << 32;
        } else if (SRC == SRC_PHILOX_TAB) {
          // per-image lookup table built by prep_build_kernel: one 8-byte gather per draw
          if (gs_layout) {
            // entry j = (gt, pred) of the j-th valid pixel (full mask: of pixel j); the prediction rides through
            // the sort as payload
            uint32_t spay[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
              key[k] = ((uint64_t)float_to_ordered(t[k].x) << 32) | ((uint32_t)k << 23) | (uint32_t)sel[k];
              spay[k] = __float_as_uint(t[k].y);
            }
            sort_desc_regs<K, true>(key, spay);
#pragma unroll
            for (int k = 0; k < K; ++k) s_tab[k] = __uint_as_float(spay[k]);
            have_s = true;
          } else {
            // entry j = (bits of flat index p_j, gt[p_j]) for the j-th valid pixel
#pragma unroll
            for (int k = 0; k < K; ++k)
              key[k] = ((uint64_t)float_to_ordered(t[k].y) << 32) | ((uint32_t)k << 23) | (uint32_t)__float_as_int(t[k].x);
            sort_desc_regs<K, false>(key, nopay);
          }
        } else {
          int q[K];
          if (identity) {
#pragma unroll
            for (int k = 0; k < K; ++k) q[k] = sel[k];
          } else {
#pragma unroll
            for (int k = 0; k < K; ++k) q[k] = __ldg(vflat + sel[k]);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const float g = __ldg(gt + q[k]);
            key[k] = ((uint64_t)float_to_ordered(g) << 32) | ((uint32_t)k << 23) | (uint32_t)q[k];
          }
          sort_desc_regs<K, false>(key, nopay);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          p[k] = (int)((uint32_t)key[k] & 0x7FFFFFu);
          lab[k] = ordered_to_float((uint32_t)(key[k] >> 32));
        }
        if (SCORE) {
          // scoring pass of the score-based strategies: only the ordered score leaves the kernel; the
          // kept lists are redrawn later from their Philox list id (cheaper than storing 8 B/point)
          const double sc = (P.score_cfg.promotion == PLD_PROMOTION_NEP50)
                                ? score_regs<float, K>(lab, P.score_cfg, b, reinterpret_cast<const float*>(s_lad))
                                : score_regs<double, K>(lab, P.score_cfg, b, s_lad);
          const bool f32_exact = P.score_cfg.promotion == PLD_PROMOTION_NEP50 &&
                                 P.score_cfg.strategy != PLD_STRATEGY_INFORMATION;
          const uint64_t skey = f32_exact ? score_key_f32((float)sc) : score_key(sc);
          if (active) P.score_keys[list_id] = skey;
          if (P.sel_hist != nullptr) {
            // lanes with the same bin elect one to add the group size (scores cluster in a few bins)
            const uint32_t bin = (uint32_t)(skey >> 52);
            const unsigned grp = __match_any_sync(0xffffffffu, active ? bin : (4096u + (uint32_t)lane));
            if (active && lane == __ffs((int)grp) - 1) atomicAdd(&s_hist[bin], (unsigned int)__popc(grp));
          }
          continue;
        }
        if (P.rank_out != nullptr) {
          const int warp_first = base + wid * 32;              // first list of this warp
          int cnt = P.n - warp_first;                          // active lists in this warp
          cnt = cnt > 32 ? 32 : cnt;
          float2* ro = reinterpret_cast<float2*>(P.rank_out) + ((size_t)b * (size_t)P.n + (size_t)warp_first) * K;
          const bool bulk = TMA_ROWS && cnt == 32 && (reinterpret_cast<uintptr_t>(ro) & 15) == 0;   // warp-uniform
          float2* st = s_stage + (wid * NBUF + (bulk ? tma_buf : 0)) * (32 * STRIDE_E);
          if (TMA_ROWS && tma_pending) {
            // the bulk copy that last read this buffer must be done with it (the other buffer's may still run)
            if (lane == 0) {
              if (bulk && NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
          }
#pragma unroll
          for (int k = 0; k < K; ++k) st[lane * STRIDE_E + k] = make_float2((float)p[k], lab[k]);
          if (bulk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
            __syncwarp();
            if (lane == 0) {
              const uint32_t src = (uint32_t)__cvta_generic_to_shared(st);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t"
                           "cp.async.bulk.commit_group;"
                           :: "l"(ro), "r"(src), "r"((uint32_t)(32 * K * sizeof(float2))) : "memory");
            }
            tma_pending = true;
            tma_buf ^= (NBUF - 1);
          } else {
            __syncwarp();
            if (cnt > 0) {
              const int total = cnt * K;
#pragma unroll
              for (int i = 0; i < K; ++i) {
                const int e = i * 32 + lane;
                if (e < total) {
                  const int li = e / K, kk = e - li * K;
                  ro[e] = st[li * STRIDE_E + kk];
                }
              }
            }
            __syncwarp();
          }
        }
      }

      if (LOSS) {
        float s[K], g[K];
        if (have_s) {
#pragma unroll
          for (int k = 0; k < K; ++k) s[k] = s_tab[k];
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) s[k] = __ldg(pred + p[k]);
        }
        if (inval) {
#pragma unroll
          for (int k = 0; k < K; ++k)
            if ((inval >> k) & 1u) s[k] = PLD_LOG_EPS;
        }
        const float nll = listmle_regs<K>(s, g);
        if (active) {
          local += nll;
          if (P.per_list != nullptr) P.per_list[list_id] = nll;
          if (P.grad != nullptr) {
#pragma unroll
            for (int k = 0; k < K; ++k)
              if (!((inval >> k) & 1u)) grad_add(P, grad_dst, map_off, p[k], g[k]);
          }
        }
      }
    }
  }
  if (TMA_ROWS && tma_pending && lane == 0)   // shared memory must outlive the bulk copies that read it
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (bad) atomicOr(P.status, bad);
  if (SCORE && P.sel_hist != nullptr) {
    __syncthreads();
    unsigned int* h = P.sel_hist + (size_t)b * 4096;
    for (int i = threadIdx.x; i < 4096; i += 256)
      if (s_hist[i]) atomicAdd(h + i, s_hist[i]);
  }
  if (LOSS) block_loss_epilogue(local, P.partials, P.ticket, P.scale, P.loss, P.loss_sum);
}

template <int SRC, bool LOSS>
static int launch_small_k(const ListParams& P, dim3 grid, cudaStream_t st) {
  switch (P.K) {
#define PLD_CASE(KK)                                               \
  case KK:                                                         \
    { cudaError_t e_ = launch_pdl(lists_small_kernel<KK, SRC, LOSS>, grid, dim3(256), 0, st, P); \
      if (e_ != cudaSuccess) { set_error("launch of lists_small_kernel failed: %s", cudaGetErrorString(e_)); return PLD_ECUDA; } } \
    break;
    PLD_CASE(1) PLD_CASE(2) PLD_CASE(3) PLD_CASE(4) PLD_CASE(5) PLD_CASE(6) PLD_CASE(7) PLD_CASE(8)
    PLD_CASE(9) PLD_CASE(10) PLD_CASE(11) PLD_CASE(12) PLD_CASE(13) PLD_CASE(14) PLD_CASE(15)
    PLD_CASE(16)
#undef PLD_CASE
    default:
      set_error("lists_small: K=%d out of range", P.K);
      return PLD_EINVAL;
  }
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int launch_lists_small_score(const ListParams& P, int num_sms, cudaStream_t st) {
  const int per_image_cap = lists_per_image_cap(num_sms, P.B);
  int gx = (P.n + 255) / 256;
  if (gx > per_image_cap) gx = per_image_cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)P.B);
  switch (P.K) {
#define PLD_CASE(KK)                                                              \
  case KK:                                                                        \
    { cudaError_t e_ = launch_pdl(lists_small_kernel<KK, SRC_PHILOX_TAB, false, true>, grid, dim3(256), 0, st, P); \
      if (e_ != cudaSuccess) { set_error("launch of lists_small_kernel failed: %s", cudaGetErrorString(e_)); return PLD_ECUDA; } } \
    break;
    PLD_CASE(1) PLD_CASE(2) PLD_CASE(3) PLD_CASE(4) PLD_CASE(5) PLD_CASE(6) PLD_CASE(7) PLD_CASE(8)
    PLD_CASE(9) PLD_CASE(10) PLD_CASE(11) PLD_CASE(12) PLD_CASE(13) PLD_CASE(14) PLD_CASE(15)
    PLD_CASE(16)
#undef PLD_CASE
    default:
      set_error("lists_small_score: K=%d out of range", P.K);
      return PLD_EINVAL;
  }
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int launch_lists_small(const ListParams& P, int src, bool loss, int num_sms, cudaStream_t st) {
  const int per_image_cap = lists_per_image_cap(num_sms, P.B);
  int gx = (P.n + 255) / 256;
  if (gx > per_image_cap) gx = per_image_cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)P.B);
  if (src == SRC_PHILOX) return loss ? launch_small_k<SRC_PHILOX, true>(P, grid, st) : launch_small_k<SRC_PHILOX, false>(P, grid, st);
  if (src == SRC_FED_SEL) return loss ? launch_small_k<SRC_FED_SEL, true>(P, grid, st) : launch_small_k<SRC_FED_SEL, false>(P, grid, st);
  if (src == SRC_FED_RANK) return launch_small_k<SRC_FED_RANK, true>(P, grid, st);
  if (src == SRC_PHILOX_TAB) return loss ? launch_small_k<SRC_PHILOX_TAB, true>(P, grid, st)
                                         : launch_small_k<SRC_PHILOX_TAB, false>(P, grid, st);
  set_error("lists_small: bad source %d", src);
  return PLD_EINVAL;
}

}  // namespace pld
