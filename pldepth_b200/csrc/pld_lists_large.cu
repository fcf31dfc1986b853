// Group-per-list kernels for ranking_size 17..512: LPL lanes x IPL register slots per list.
// Ordering = bitonic network over LPL*IPL slots (shuffles across lanes, register swaps inside
// a lane); ListMLE = lane-local scans + shuffle scans for the reverse cumsum / prefix sums.
#include <stdlib.h>

#include "pld_group.cuh"

namespace pld {

#ifndef PLD_LARGE_MINBLOCKS
#define PLD_LARGE_MINBLOCKS 4
#endif
template <int LPL, int IPL, int SRC, bool LOSS>
__global__ void __launch_bounds__(256, PLD_LARGE_MINBLOCKS) lists_large_kernel(const ListParams P) {
  __shared__ uint32_t s_park[(SRC == SRC_PHILOX_TAB) ? IPL * 256 : 1];
  constexpr int GPW = 32 / LPL;    // groups per warp
  constexpr int GPB = 256 / LPL;   // groups per block
  const int K = P.K;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPL - 1);
  // bit i set <=> slot gl*IPL + i holds a real entry (slot index < K)
  const uint32_t emask = (K - gl * IPL >= IPL) ? ((1u << IPL) - 1u) : ((K - gl * IPL > 0) ? ((1u << (K - gl * IPL)) - 1u) : 0u);
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
    identity = mraw < 0;
    vflat = P.valid_flat + (size_t)b * (size_t)P.valid_stride;
  }

  const bool vj = (SRC == SRC_PHILOX_TAB) && !identity && P.grad_valid != nullptr;
  const bool gs_layout = identity || vj;
  float* grad_dst = P.grad == nullptr ? nullptr
                    : (vj ? P.grad_valid + (size_t)b * P.table_stride : P.grad + map_off);

  if (M != 0) {
    for (int l0 = blockIdx.x * GPB + (threadIdx.x >> 5) * GPW; l0 < P.n; l0 += gridDim.x * GPB) {
      const int lraw = l0 + lane / LPL;
      const bool active = lraw < P.n;
      const int l = active ? lraw : (P.n - 1);
      const size_t list_id = (size_t)b * (size_t)P.n + (size_t)l;
      int p[IPL];
      float lab[IPL];
      uint32_t inval = 0;
      uint32_t s_tab[IPL];   // predictions delivered by the (gt, pred) table (bits), sorted order
      bool have_s = false;

      if (SRC == SRC_FED_RANK) {
        const float2* __restrict__ rin = reinterpret_cast<const float2*>(P.rank_in) + list_id * K;
        bool ok = true;  // locally sorted and valid
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          const int e = gl * IPL + i;
          p[i] = 0;
          lab[i] = 0.f;
          if (e < K) {
            const float2 v = __ldg(rin + e);
            int q = (int)v.x;
            if (q < 0 || q >= P.HW) { bad |= PLD_ST_BAD_INDEX; q = 0; }
            p[i] = q;
            lab[i] = v.y;
            ok = ok && (v.y >= 0.f);
          }
        }
#pragma unroll
        for (int i = 1; i < IPL; ++i)
          if (gl * IPL + i < K) ok = ok && (lab[i - 1] >= lab[i]);
        {
          const float nxt = __shfl_down_sync(0xffffffffu, lab[0], 1, LPL);
          if (gl + 1 < LPL && (gl + 1) * IPL < K) ok = ok && (lab[IPL - 1] >= nxt);
        }
        const bool need_sort = __any_sync(0xffffffffu, !ok);
        if (need_sort) {
          float mn = 3.402823466e38f;
#pragma unroll
          for (int i = 0; i < IPL; ++i)
            if (gl * IPL + i < K) mn = fminf(mn, lab[i] >= 0.f ? lab[i] : 0.f);
          mn = group_min<LPL>(mn);
          const float inv_key = mn - 1e-6f;
          uint32_t khi[IPL], klo[IPL], pay[IPL];
#pragma unroll
          for (int i = 0; i < IPL; ++i) {
            const int e = gl * IPL + i;
            const bool on = e < K;
            const bool v = lab[i] >= 0.f;
            khi[i] = on ? float_to_ordered(v ? lab[i] : inv_key) : 0u;
            klo[i] = on ? (uint32_t)(0xFFFF - e) : 0u;
            pay[i] = on ? ((uint32_t)p[i] | (v ? 0u : 0x80000000u)) : 0u;
          }
          bitonic_desc<LPL, IPL, true>(khi, klo, pay, gl);
#pragma unroll
          for (int i = 0; i < IPL; ++i) {
            p[i] = (int)(pay[i] & 0x7FFFFFFFu);
            if (pay[i] >> 31) inval |= (1u << i);
          }
        }
      } else {
        uint32_t khi[IPL], klo[IPL], nopay[IPL];
        const DrawStream ds{(uint32_t)l, (uint32_t)(P.image_base + b), off_lo, off_hi16,
                            P.seed_lo, P.seed_hi};
        int sel[IPL];
#pragma unroll
        for (int i = 0; i < IPL; ++i) sel[i] = 0;
        if (SRC == SRC_PHILOX || SRC == SRC_PHILOX_TAB) {
          bool rej = false;
#pragma unroll
          for (int q = 0; q < IPL / 4; ++q) {
            const int e0 = gl * IPL + q * 4;
            if (e0 < K) {
              const Philox4 r = ds.block((uint32_t)(e0 >> 2));
              sel[q * 4 + 0] = (int)lemire_try(r.x, M, thresh, rej);
              sel[q * 4 + 1] = (int)lemire_try(r.y, M, thresh, rej);
              sel[q * 4 + 2] = (int)lemire_try(r.z, M, thresh, rej);
              sel[q * 4 + 3] = (int)lemire_try(r.w, M, thresh, rej);
            }
          }
          if (rej) {  // rare (P < K * M / 2^32): redo this lane's draws with the redraw stream
#pragma unroll
            for (int q = 0; q < IPL / 4; ++q) {
              const int e0 = gl * IPL + q * 4;
              if (e0 < K) {
                const Philox4 r = ds.block((uint32_t)(e0 >> 2));
                const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  sel[q * 4 + j] = (int)lemire_bounded(w[j], M, thresh, ds, (uint32_t)(e0 + j));
              }
            }
          }
          if (P.sel_out != nullptr && active) {
            int32_t* so = P.sel_out + list_id * K;
#pragma unroll
            for (int i = 0; i < IPL; ++i)
              if ((emask >> i) & 1u) so[gl * IPL + i] = sel[i];
          }
        } else {
          const int32_t* __restrict__ sin = P.sel_in + list_id * K;
#pragma unroll
          for (int i = 0; i < IPL; ++i) {
            const int e = gl * IPL + i;
            int s = __ldg(sin + (e < K ? e : 0));
            if (s < 0 || (uint32_t)s >= M) { bad |= PLD_ST_BAD_INDEX; s = 0; }
            sel[i] = s;
          }
        }
        // pad slots read entry 0 (always valid) and get the smallest key, so they sort last
        if (SRC == SRC_PHILOX_TAB) {
          // per-image 8-byte lookup table (pld_step.cu): one gather per draw
          const float2* __restrict__ tab = P.table + (size_t)b * P.table_stride;
          float2 t[IPL];
#pragma unroll
          for (int i = 0; i < IPL; ++i) t[i] = __ldg(tab + sel[i]);
          if (gs_layout) {  // entry j = (gt, pred) of the j-th valid pixel; the prediction is parked per draw slot
#pragma unroll
            for (int i = 0; i < IPL; ++i) {
              const bool on = (emask >> i) & 1u;
              khi[i] = on ? float_to_ordered(t[i].x) : 0u;
              klo[i] = on ? (((uint32_t)(gl * IPL + i) << 23) | (uint32_t)sel[i]) : 0u;
              // the prediction is parked in shared memory under its draw slot and fetched back by the slot id
              // that survives in klo: 2 shared-memory ops per entry instead of a payload word in all 21+ stages
              s_park[i * 256 + threadIdx.x] = __float_as_uint(t[i].y);
            }
            bitonic_desc<LPL, IPL, false>(khi, klo, nopay, gl);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < IPL; ++i) {
              const uint32_t e = klo[i] >> 23;                       // original draw slot of this sorted entry
              const int owner = (int)(threadIdx.x & ~(LPL - 1)) + (int)(e / IPL);
              s_tab[i] = s_park[(e % IPL) * 256 + owner];
            }
            __syncwarp();
            have_s = true;
          } else {         // entry j = (bits of flat index p_j, gt[p_j])
#pragma unroll
            for (int i = 0; i < IPL; ++i) {
              const bool on = (emask >> i) & 1u;
              khi[i] = on ? float_to_ordered(t[i].y) : 0u;
              klo[i] = on ? (((uint32_t)(gl * IPL + i) << 23) | (uint32_t)__float_as_int(t[i].x)) : 0u;
            }
            bitonic_desc<LPL, IPL, false>(khi, klo, nopay, gl);
          }
        } else {
          int qv[IPL];
#pragma unroll
          for (int i = 0; i < IPL; ++i) qv[i] = identity ? sel[i] : __ldg(vflat + sel[i]);
#pragma unroll
          for (int i = 0; i < IPL; ++i) {
            const float g = __ldg(gt + qv[i]);
            const bool on = (emask >> i) & 1u;
            khi[i] = on ? float_to_ordered(g) : 0u;
            klo[i] = on ? (((uint32_t)(gl * IPL + i) << 23) | (uint32_t)qv[i]) : 0u;
          }
          bitonic_desc<LPL, IPL, false>(khi, klo, nopay, gl);
        }
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          p[i] = (int)(klo[i] & 0x7FFFFFu);
          lab[i] = ordered_to_float(khi[i]);
        }
        if (P.rank_out != nullptr && active) {
          float2* ro = reinterpret_cast<float2*>(P.rank_out) + list_id * K + gl * IPL;
          if ((K & 1) == 0) {  // 16-byte stores: list rows are 16-byte aligned when K is even
#pragma unroll
            for (int i = 0; i < IPL; i += 2)
              if ((emask >> i) & 1u)
                *reinterpret_cast<float4*>(ro + i) = make_float4((float)p[i], lab[i], (float)p[i + 1], lab[i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < IPL; ++i)
              if ((emask >> i) & 1u) ro[i] = make_float2((float)p[i], lab[i]);
          }
        }
      }

      if (LOSS) {
        float s[IPL], g[IPL];
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          const float sv = have_s ? __uint_as_float(s_tab[i]) : __ldg(pred + p[i]);  // pads: p == 0 is valid
          s[i] = ((inval >> i) & 1u) ? PLD_LOG_EPS : sv;
        }
        const int nreal = min(max(K - gl * IPL, 0), IPL);
        const float nll = group_listmle<LPL, IPL>(s, nreal, gl, g);
        if (active) {
          if (gl == 0) {
            local += nll;
            if (P.per_list != nullptr) P.per_list[list_id] = nll;
          }
          if (P.grad != nullptr) {
#pragma unroll
            for (int i = 0; i < IPL; ++i)
              if (((emask & ~inval) >> i) & 1u) grad_add(P, grad_dst, map_off, p[i], g[i]);
          }
        }
      }
    }
  }
  if (bad) atomicOr(P.status, bad);
  if (LOSS) block_loss_epilogue(local, P.partials, P.ticket, P.scale, P.loss, P.loss_sum);
}

template <int LPL, int IPL>
static int launch_large_cfg(const ListParams& P, int src, bool loss, dim3 grid, cudaStream_t st) {
  if (src == SRC_PHILOX) {
    if (loss) lists_large_kernel<LPL, IPL, SRC_PHILOX, true><<<grid, 256, 0, st>>>(P);
    else lists_large_kernel<LPL, IPL, SRC_PHILOX, false><<<grid, 256, 0, st>>>(P);
  } else if (src == SRC_FED_SEL) {
    if (loss) lists_large_kernel<LPL, IPL, SRC_FED_SEL, true><<<grid, 256, 0, st>>>(P);
    else lists_large_kernel<LPL, IPL, SRC_FED_SEL, false><<<grid, 256, 0, st>>>(P);
  } else if (src == SRC_PHILOX_TAB) {
    lists_large_kernel<LPL, IPL, SRC_PHILOX_TAB, true><<<grid, 256, 0, st>>>(P);
  } else if (src == SRC_FED_RANK) {
    lists_large_kernel<LPL, IPL, SRC_FED_RANK, true><<<grid, 256, 0, st>>>(P);
  } else {
    set_error("lists_large: bad source %d", src);
    return PLD_EINVAL;
  }
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

int launch_lists_tab(const ListParams& P, bool loss, int num_sms, cudaStream_t st);   // pld_lists_tab.cu

int launch_lists_large(const ListParams& P, int src, bool loss, int num_sms, cudaStream_t st) {
  const int K = P.K;
  // the one-call steps run the software-pipelined kernel (PLD_LARGE_LEGACY=1 keeps the round-1 kernel for A/B runs)
  static const bool legacy = getenv("PLD_LARGE_LEGACY") != nullptr && getenv("PLD_LARGE_LEGACY")[0] == '1';
  if (src == SRC_PHILOX_TAB && !legacy) return launch_lists_tab(P, loss, num_sms, st);
  int lpl;
  if (K <= 32) lpl = 4;
  else if (K <= 64) lpl = 8;
  else if (K <= 128) lpl = 16;
  else lpl = 32;
  const int gpb = 256 / lpl;
  const int per_image_cap = lists_per_image_cap(num_sms, P.B);
  int gx = (P.n + gpb - 1) / gpb;
  if (gx > per_image_cap) gx = per_image_cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)P.B);
  if (K <= 32) return launch_large_cfg<4, 8>(P, src, loss, grid, st);
  if (K <= 64) return launch_large_cfg<8, 8>(P, src, loss, grid, st);
  if (K <= 128) return launch_large_cfg<16, 8>(P, src, loss, grid, st);
  if (K <= 256) return launch_large_cfg<32, 8>(P, src, loss, grid, st);
  if (K <= 512) return launch_large_cfg<32, 16>(P, src, loss, grid, st);
  set_error("lists_large: K=%d out of range", K);
  return PLD_EINVAL;
}

}  // namespace pld
