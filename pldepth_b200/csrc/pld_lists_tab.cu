// Software-pipelined group-per-list kernel of the one-call steps for ranking_size 17..512 (BASELINE config 3: K = 50):
// Philox draws -> one 8-byte lookup-table gather per draw -> per-list order by gt depth -> emit rankings ->
// ListMLE forward / backward -> gradient reductions.   Same results as lists_large_kernel<.., SRC_PHILOX_TAB, ..>
// (pld_lists_large.cu), which it replaces on that path; organised around what round 1's profile showed
// (profiles/r01_c3_final_summary.txt: ALU pipe 68 %, issue 59 %, L1TEX 74 % -- nothing saturated, nothing overlapped):
//
//  * the draws and table gathers of list group i+1 are issued BEFORE group i is ordered, so the L1TEX queue is never
//    empty while a warp runs its ordering network; they land in registers and are parked in shared memory after
//    group i is finished (cp.async straight into shared memory was measured and rejected: a scattered LDGSTS spends one
//    shared-memory wavefront per LANE on its write-back, 48 M wavefronts per config-3 launch, profiles/r02_c3_v1);
//  * the ordering network runs on ONE 32-bit key per entry -- depth prefix | draw slot -- so a compare-exchange is a
//    min + a max (2 ALU-pipe instructions) instead of a 64-bit compare + four selects; payloads (depth, prediction,
//    pixel) never travel through the network: they stay parked under their draw slot and are fetched once, by the
//    slot id that survives in the key's low bits;
//  * the truncated depth prefix is verified after the fetch (adjacent full depths must be non-increasing); a
//    group that fails -- two depths of one list equal in their top 32 - log2(slots) bits but not equal -- is re-ordered
//    by the exact 64-bit network.  Exact ties keep the rule "later draw first" in both networks.
//
// Reference semantics: sample_single_masked_ranking (pldepth/data/sampling.py:110-122), prepare_fully_fledged_loss_input
// (pldepth/data/depth_utils.py:39-61), TF-Ranking ListMLE behind nll_loss.py:43-62.
#include <type_traits>

#include "pld_group.cuh"

namespace pld {

template <int N> struct ILog2 { static constexpr int v = 1 + ILog2<N / 2>::v; };
template <> struct ILog2<1> { static constexpr int v = 0; };

// 8-byte read-only load under a predicate, without a branch; (0, 0) when off
__device__ __forceinline__ float2 ldg_f2_if(const float2* ptr, bool on) {
  float2 v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
               "@p ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
               : "=f"(v.x), "=f"(v.y) : "l"(ptr), "r"((uint32_t)on));
  return v;
}

// min or max by a per-lane predicate in two issue slots (max; @!keep_max min) -- the select form costs three
__device__ __forceinline__ uint32_t minmax_pred(uint32_t a, uint32_t o, bool keep_max) {
  uint32_t r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmax.u32 %0, %1, %2;\n\t@!p min.u32 %0, %1, %2;\n\t}"
      : "=r"(r) : "r"(a), "r"(o), "r"((uint32_t)keep_max));
  return r;
}

// Bitonic network, descending, on single 32-bit keys (slot e = lane*IPL + i).  Blocks that the classic network sorts
// ASCENDING hold their keys complemented instead (x -> ~x reverses the order), so every compare-exchange inside a
// lane is the same static max / min pair; a lane complements its keys only where its block direction changes between
// merge levels (IPL LOP3 per level >= IPL).  Across lanes the lower lane keeps the maximum, the upper the minimum.
template <int LPL, int IPL>
__device__ __forceinline__ void bitonic_desc32(uint32_t (&key)[IPL], int gl) {
  constexpr int N = LPL * IPL;
  uint32_t flipped = 0u;   // all-ones while this lane's keys are stored complemented
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
    if (k >= IPL) {
      // direction of this lane's block at level k depends on the lane only
      const uint32_t want = (k < N && ((gl * IPL) & k) != 0) ? 0xFFFFFFFFu : 0u;
      const uint32_t d = flipped ^ want;
#pragma unroll
      for (int i = 0; i < IPL; ++i) key[i] ^= d;
      flipped = want;
    }
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= IPL) {
        const int lm = j / IPL;
        const bool lower = (gl & lm) == 0;
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          const uint32_t o = __shfl_xor_sync(0xffffffffu, key[i], lm);
          key[i] = minmax_pred(key[i], o, lower);
        }
      } else {
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          if ((i & j) == 0) {
            const bool up = (k >= IPL) || ((i & k) == 0);   // static
            const uint32_t a = key[i], c = key[i | j];
            key[i] = up ? max(a, c) : min(a, c);
            key[i | j] = up ? min(a, c) : max(a, c);
          }
        }
      }
    }
  }
}

// Exact fallback: full 64-bit keys (ordered depth, draw slot) rebuilt from the parked entries; returns the draw slot of
// every sorted position.
template <int LPL, int IPL>
__device__ __noinline__ void exact_order(const float2* own, bool gs_layout, uint32_t emask, int gl, uint32_t (&eslot)[IPL]) {
  uint32_t khi[IPL], klo[IPL], nopay[IPL];
#pragma unroll
  for (int i = 0; i < IPL; ++i) {
    const bool on = (emask >> i) & 1u;
    const float2 t = on ? own[i] : make_float2(0.f, 0.f);
    khi[i] = on ? float_to_ordered(gs_layout ? t.x : t.y) : 0u;
    klo[i] = on ? ((uint32_t)(gl * IPL + i) << 23) : 0u;
  }
  bitonic_desc<LPL, IPL, false>(khi, klo, nopay, gl);
#pragma unroll
  for (int i = 0; i < IPL; ++i) eslot[i] = klo[i] >> 23;
}

#ifndef PLD_TAB_MINBLOCKS
#define PLD_TAB_MINBLOCKS 3
#endif

// SCORE: scoring pass of the score-based strategies -- the ordered depths of every candidate list are turned into its
// 8-byte score key (NumPy-exact arithmetic, pld_score.cuh) and nothing else leaves the kernel.
template <int LPL, int IPL, int THREADS, bool LOSS, bool SCORE = false>
__global__ void __launch_bounds__(THREADS, (THREADS == 256) ? PLD_TAB_MINBLOCKS : 4) lists_tab_kernel(const ListParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_sync();
  // Parking buffers.  A lane owns ROW = IPL + 1 consecutive entries (the odd stride keeps its own 8-byte accesses at the
  // two-wavefront minimum); draw slot e of a list group sits at  group_base + e + (e >> log2 IPL).  s_ent: one stage (a
  // warp refills its own rows after it has fetched everything from them); s_aux (selection index): two stages, because
  // it is written when the draws are made, one group ahead.
  constexpr int ROW = IPL + 1;
  constexpr int STAGE = ROW * THREADS;
  float2* const s_ent = reinterpret_cast<float2*>(smem_raw);
  uint32_t* const s_aux = reinterpret_cast<uint32_t*>(smem_raw + (size_t)STAGE * sizeof(float2));
  constexpr int GPW = 32 / LPL;         // list groups per warp
  constexpr int GPB = THREADS / LPL;    // list groups per CTA
  constexpr int SLOT_BITS = ILog2<LPL * IPL>::v;
  constexpr uint32_t SLOT_MASK = (1u << SLOT_BITS) - 1u;
  constexpr int LOG_IPL = ILog2<IPL>::v;
  const int K = P.K;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int gl = lane & (LPL - 1);
  const int own = tid * ROW;                       // this lane's row
  const int grp = (tid & ~(LPL - 1)) * ROW;        // first row of its list group
  // real entries of this lane: slots / sorted positions gl*IPL + i with i < nreal
  const int nreal = min(max(K - gl * IPL, 0), IPL);
  const uint32_t emask = (1u << nreal) - 1u;
  const size_t map_off = (size_t)b * (size_t)P.HW;
  const float* __restrict__ pred = P.pred + map_off;
  float local = 0.f;
  int bad = 0;
  uint32_t off_lo, off_hi16;
  launch_offset(P, off_lo, off_hi16);

  uint32_t M = 0, thresh = 0;
  bool identity = false;
  {
    const int mraw = P.n_valid[b];
    const int m = mraw < 0 ? -mraw : mraw;
    if (m == 0) bad |= PLD_ST_EMPTY_MASK;
    else { M = (uint32_t)m; thresh = (0u - M) % M; }
    identity = mraw < 0;
  }
  const bool vj = !identity && P.grad_valid != nullptr;
  const bool gs_layout = identity || vj;      // table entry = (gt, pred); else (bits of the pixel index, gt)
  float* grad_dst = P.grad == nullptr ? nullptr : (vj ? P.grad_valid + (size_t)b * P.table_stride : P.grad + map_off);
  const float2* __restrict__ tab = P.table + (size_t)b * P.table_stride;
  const uint32_t* __restrict__ lmap = P.list_map != nullptr ? P.list_map + (size_t)b * P.map_stride : nullptr;
  const uint32_t image = (uint32_t)(P.image_base + b);

  // ---- stage A: draws of one list group; table gathers issued into registers, selections parked in aux[stage] ---------
  auto issue = [&](int l0, int stage, float2 (&nt)[IPL]) {
    const int lraw = l0 + lane / LPL;
    const int l = lraw < P.n ? lraw : P.n - 1;
    const uint32_t lid = lmap != nullptr ? __ldg(lmap + l) : (uint32_t)l;
    uint32_t sel[IPL];
#pragma unroll
    for (int i = 0; i < IPL; ++i) sel[i] = 0u;
    bool rej = false;
#pragma unroll
    for (int q = 0; q < IPL / 4; ++q) {
      const int e0 = gl * IPL + q * 4;
      if (e0 < K) {
        const Philox4 r = philox4x32_10_rk(lid, image, (uint32_t)(e0 >> 2) | off_hi16, off_lo, P.rk0, P.rk1);
        sel[q * 4 + 0] = lemire_try(r.x, M, thresh, rej);
        sel[q * 4 + 1] = lemire_try(r.y, M, thresh, rej);
        sel[q * 4 + 2] = lemire_try(r.z, M, thresh, rej);
        sel[q * 4 + 3] = lemire_try(r.w, M, thresh, rej);
      }
    }
    if (rej) {  // rare (P < K * M / 2^32): redo this lane's draws with the redraw stream
      const DrawStream ds{lid, image, off_lo, off_hi16, P.seed_lo, P.seed_hi};
#pragma unroll
      for (int q = 0; q < IPL / 4; ++q) {
        const int e0 = gl * IPL + q * 4;
        if (e0 < K) {
          const Philox4 r = ds.block((uint32_t)(e0 >> 2));
          const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) sel[q * 4 + j] = lemire_bounded(w[j], M, thresh, ds, (uint32_t)(e0 + j));
        }
      }
    }
    // pad slots gather nothing (predicated load: an extra sector per load instruction otherwise) and park (0, 0)
    uint32_t* aux = s_aux + stage * STAGE + own;
#pragma unroll
    for (int i = 0; i < IPL; ++i) {
      nt[i] = ldg_f2_if(tab + sel[i], i < nreal);
      aux[i] = sel[i];
    }
    if (P.sel_out != nullptr && lraw < P.n) {
      int32_t* so = P.sel_out + ((size_t)b * (size_t)P.n + (size_t)l) * K;
#pragma unroll
      for (int i = 0; i < IPL; ++i)
        if (i < nreal) so[gl * IPL + i] = (int32_t)sel[i];
    }
  };

  // ---- stage B: order, emit, loss, gradient of the group parked in s_ent / aux[stage] -------------------------------
  auto process = [&](int l0, int stage, auto gs_tag) {
    constexpr bool GS = decltype(gs_tag)::value;
    const int lraw = l0 + lane / LPL;
    const bool active = lraw < P.n;
    const int l = active ? lraw : (P.n - 1);
    const size_t list_id = (size_t)b * (size_t)P.n + (size_t)l;
    uint32_t* aux = s_aux + stage * STAGE;

    // own draw slots -> 32-bit keys: ordered depth with its low SLOT_BITS replaced by the draw slot; pads = bare slot
    uint32_t key[IPL];
    float spre[GS ? 1 : IPL];
#pragma unroll
    for (int i = 0; i < IPL; ++i) {
      const float2 t = s_ent[own + i];
      const uint32_t slot = (uint32_t)(gl * IPL + i);
      key[i] = (i < nreal) ? ((float_to_ordered(GS ? t.x : t.y) & ~SLOT_MASK) | slot) : slot;
      // pixel layout (holed mask, rankings emitted): the prediction needs its own gather, issued here so that its
      // latency hides behind the ordering network (pads hold entry 0 = a valid pixel)
      if (!GS && LOSS) spre[i] = __ldg(pred + __float_as_int(t.x));
    }
    bitonic_desc32<LPL, IPL>(key, gl);
    if (!GS && LOSS) {
      // park the predictions under their draw slots (the selection index is not needed in this layout)
#pragma unroll
      for (int i = 0; i < IPL; ++i) aux[own + i] = __float_as_uint(spre[i]);
    }
    __syncwarp();

    // fetch the parked payloads by surviving slot id
    float2 e2[IPL];
    uint32_t ax[IPL];
    uint32_t eslot[IPL];
#pragma unroll
    for (int i = 0; i < IPL; ++i) eslot[i] = key[i] & SLOT_MASK;
    auto unpark = [&]() {
#pragma unroll
      for (int i = 0; i < IPL; ++i) {
        // pad positions re-read this lane's own row (no new bank, nothing used from it)
        const int src = (i < nreal) ? grp + (int)(eslot[i] + (eslot[i] >> LOG_IPL)) : own + i;
        e2[i] = s_ent[src];
        ax[i] = aux[src];
      }
    };
    unpark();
    // verify the truncated order on the full depths
    {
      bool viol = false;
      uint32_t og[IPL];
#pragma unroll
      for (int i = 0; i < IPL; ++i) og[i] = float_to_ordered(GS ? e2[i].x : e2[i].y);
#pragma unroll
      for (int i = 0; i < IPL; ++i) {
        if (i < nreal && eslot[i] >= (uint32_t)K) viol = true;                  // a pad sorted in front of a real entry
        if (i + 1 < IPL && i + 1 < nreal && og[i] < og[i + 1]) viol = true;
      }
      const uint32_t nxt = __shfl_down_sync(0xffffffffu, og[0], 1, LPL);
      if (gl + 1 < LPL && (gl + 1) * IPL < K && og[IPL - 1] < nxt) viol = true;
      if (__any_sync(0xffffffffu, viol)) {
        uint32_t ex_slot[IPL];     // separate array: its address escapes into the call, eslot must stay in registers
        exact_order<LPL, IPL>(s_ent + own, GS, emask, gl, ex_slot);
#pragma unroll
        for (int i = 0; i < IPL; ++i) eslot[i] = ex_slot[i];
        unpark();
      }
    }

    int p[IPL];
    float lab[IPL], sv[IPL];
#pragma unroll
    for (int i = 0; i < IPL; ++i) {
      if (GS) { lab[i] = e2[i].x; sv[i] = e2[i].y; p[i] = (int)ax[i]; }
      else { lab[i] = e2[i].y; sv[i] = __uint_as_float(ax[i]); p[i] = __float_as_int(e2[i].x); }
    }
    if constexpr (SCORE) {
      // Per-position parts in parallel (lane gl owns sorted positions gl*IPL + i): chi term or |difference| into the
      // lane's own parking row, "equal relation with the next position" into its aux row; then lane 0 of the group
      // combines them serially in NumPy's order of operations (score_combine) -- the same code the staged
      // pld_score_lists runs, so both pipelines produce identical keys.
      const ScoreCfg& C = P.score_cfg;
      const bool info = C.strategy == PLD_STRATEGY_INFORMATION;
      const bool want_eq = info || C.strategy == PLD_STRATEGY_THRESHOLDED;
      const float nxt0 = __shfl_down_sync(0xffffffffu, lab[0], 1, LPL);
      __syncwarp();     // every lane has fetched its payloads: the rows are free
      if constexpr (LPL * IPL <= 128 && IPL == 8) {
        if (C.promotion == PLD_PROMOTION_NEP50) {
          // float32 scores of lists of up to 128 entries, combined by the whole group instead of one lane:
          //  * per-position terms go to a compact row of floats (sorted position q at s_term[q]), the "equal" relations
          //    into a bit per position;
          //  * information: NumPy's pairwise summation of n <= 128 terms IS eight strided accumulators
          //    r[j] = a[j] + a[8 + j] + ... combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail --
          //    one accumulator per lane (two for four-lane groups) and an xor-butterfly reproduce it bit for bit
          //    (IEEE addition is commutative);
          //  * masked / thresholded: the sum is a strictly sequential chain (penalties interleaved), run by one lane
          //    from vector loads of the compact row.
          float* const s_term = reinterpret_cast<float*>(s_ent + grp);
          float start = 0.f, stop = 0.f, delta = 0.f, step = 0.f;
          if (info) ladder_setup<float>(C, b, K, start, stop, delta, step);
          float term[IPL];
          uint32_t fl = 0u;
#pragma unroll
          for (int i = 0; i < IPL; ++i) {
            const int ppos = gl * IPL + i;
            const bool on = i < nreal;
            const float gn = (i + 1 < IPL) ? lab[(i + 1 < IPL) ? i + 1 : i] : nxt0;
            const bool has_next = ppos + 1 < K;
            if (on && want_eq && has_next && relation_equal<float>(lab[i], gn, C)) fl |= 1u << i;
            if (info) term[i] = on ? chi_term<float>(lab[i], ppos, K, start, stop, delta, step) : 0.f;
            else term[i] = (on && has_next) ? fabsf(__fsub_rn(lab[i], gn)) : 0.f;
          }
          {
            float4* dst = reinterpret_cast<float4*>(s_term + gl * IPL);
            dst[0] = make_float4(term[0], term[1], term[2], term[3]);
            dst[1] = make_float4(term[4], term[5], term[6], term[7]);
          }
          __syncwarp();
          double sc;
          if (info) {
            const int n8 = K & ~7;            // K >= 17: NumPy's unrolled branch
            auto strided = [&](int j) {
              float r = s_term[j];
#pragma unroll 4
              for (int q = 8 + j; q < n8; q += 8) r = __fadd_rn(r, s_term[q]);
              return r;
            };
            float res;
            if constexpr (LPL >= 8) {      // lanes 0..7 of the group hold the eight accumulators
              float t = strided(gl & 7);
              t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 1, LPL));
              t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 2, LPL));
              res = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 4, LPL));
            } else {
              float ta = strided(gl), tb = strided(gl + 4);
              ta = __fadd_rn(ta, __shfl_xor_sync(0xffffffffu, ta, 1, LPL));
              tb = __fadd_rn(tb, __shfl_xor_sync(0xffffffffu, tb, 1, LPL));
              ta = __fadd_rn(ta, __shfl_xor_sync(0xffffffffu, ta, 2, LPL));
              tb = __fadd_rn(tb, __shfl_xor_sync(0xffffffffu, tb, 2, LPL));
              res = __fadd_rn(ta, tb);
            }
            for (int q = n8; q < K; ++q) res = __fadd_rn(res, s_term[q]);
            int cnt = __popc(fl);
#pragma unroll
            for (int o = LPL / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o, LPL);
            sc = (double)(-res);
            for (int c = 0; c < cnt; ++c) sc = __dadd_rn(sc, C.penalty);
            if (gl == 0 && active) P.score_keys[list_id] = score_key(sc);
          } else {
            // the group's relation bits, position q at bit q & 31 of word q >> 5
            constexpr int NW = (LPL * IPL + 31) / 32;
            uint32_t mw[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) mw[w] = ((gl * IPL) >> 5) == w ? fl << ((gl * IPL) & 31) : 0u;
#pragma unroll
            for (int o = LPL / 2; o > 0; o >>= 1) {
#pragma unroll
              for (int w = 0; w < NW; ++w) mw[w] |= __shfl_xor_sync(0xffffffffu, mw[w], o, LPL);
            }
            if (gl == 0) {
              const bool thr = C.strategy == PLD_STRATEGY_THRESHOLDED;
              const float pen = (float)C.penalty;
              const float4* src = reinterpret_cast<const float4*>(s_term);
              float acc = 0.f;
#pragma unroll
              for (int c4 = 0; c4 < (LPL * IPL) / 4; ++c4) {
                if (c4 * 4 + 1 < K) {     // uniform: some position of this vector has a successor
                  const float4 v = src[c4];
                  const float d[4] = {v.x, v.y, v.z, v.w};
                  const uint32_t m4 = mw[(c4 * 4) >> 5] >> ((c4 * 4) & 31);
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    if (c4 * 4 + u + 1 < K) {
                      if (thr && ((m4 >> u) & 1u)) acc = __fadd_rn(acc, pen);
                      acc = __fadd_rn(acc, d[u]);
                    }
                  }
                }
              }
              if (active) P.score_keys[list_id] = score_key_f32(acc);
            }
          }
          __syncwarp();   // the combination is done before the rows are refilled
          return;
        }
      }
      auto parts = [&](auto tag) {
        using T = decltype(tag);
        T start = (T)0, stop = (T)0, delta = (T)0, step = (T)0;
        if (info) ladder_setup<T>(C, b, K, start, stop, delta, step);
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          if (i < nreal) {
            const int ppos = gl * IPL + i;
            const float gn = (i + 1 < IPL) ? lab[(i + 1 < IPL) ? i + 1 : i] : nxt0;
            const bool has_next = ppos + 1 < K;
            const bool e = (want_eq && has_next) ? relation_equal<T>(lab[i], gn, C) : false;
            aux[own + i] = e ? 1u : 0u;
            if (info) {
              const T c = chi_term<T>(lab[i], ppos, K, start, stop, delta, step);
              if (sizeof(T) == 4) s_ent[own + i] = make_float2((float)c, 0.f);
              else *reinterpret_cast<double*>(&s_ent[own + i]) = (double)c;
            } else {
              s_ent[own + i] = make_float2(has_next ? fabsf(__fsub_rn(lab[i], gn)) : 0.f, 0.f);
            }
          }
        }
      };
      if (C.promotion == PLD_PROMOTION_NEP50) parts(0.f);
      else parts(0.0);
      __syncwarp();
      if (gl == 0) {
        auto posf = [&](int q) { return grp + q + (q >> LOG_IPL); };
        auto eqf = [&](int j) { return aux[posf(j)] != 0u; };
        auto difff = [&](int j) { return s_ent[posf(j)].x; };
        double sc;
        if (C.promotion == PLD_PROMOTION_NEP50) {
          auto chif = [&](int k) { return s_ent[posf(k)].x; };
          sc = score_combine<float>(chif, difff, eqf, K, C);
        } else {
          auto chif = [&](int k) { return *reinterpret_cast<const double*>(&s_ent[posf(k)]); };
          sc = score_combine<double>(chif, difff, eqf, K, C);
        }
        const bool f32_exact = C.promotion == PLD_PROMOTION_NEP50 && !info;
        if (active) P.score_keys[list_id] = f32_exact ? score_key_f32((float)sc) : score_key(sc);
      }
      __syncwarp();   // the combination is done before the rows are refilled
      return;
    }
    if (P.rank_out != nullptr && active) {
      float2* ro = reinterpret_cast<float2*>(P.rank_out) + list_id * K + gl * IPL;
      if ((K & 1) == 0) {  // 16-byte stores: list rows are 16-byte aligned when K is even
#pragma unroll
        for (int i = 0; i < IPL; i += 2)
          if (i < nreal)
            *reinterpret_cast<float4*>(ro + i) = make_float4((float)p[i], lab[i], (float)p[i + 1], lab[i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < IPL; ++i)
          if (i < nreal) ro[i] = make_float2((float)p[i], lab[i]);
      }
    }

    if (LOSS) {
      float g[IPL];
      const float nll = group_listmle<LPL, IPL>(sv, nreal, gl, g);
      if (active && gl == 0) {
        local += nll;
        if (P.per_list != nullptr) P.per_list[list_id] = nll;
      }
      if (P.grad != nullptr) {
        const int lim = active ? nreal : 0;
        if (P.acc != nullptr) {   // deterministic mode: 64-bit fixed-point atomics
#pragma unroll
          for (int i = 0; i < IPL; ++i)
            if (i < lim) grad_add(P, grad_dst, map_off, p[i], g[i]);
        } else {
          const float sc = P.scale;
#pragma unroll
          for (int i = 0; i < IPL; ++i)
            if (i < lim) red_add_f32(grad_dst + p[i], g[i] * sc);
        }
      }
    }
    __syncwarp();   // every lane is done with the parked group before it is overwritten
  };

  if (M != 0) {
    const int stride = gridDim.x * GPB;
    int l0 = blockIdx.x * GPB + (tid >> 5) * GPW;
    int stage = 0;
    float2 nt[IPL];
    if (l0 < P.n) {
      issue(l0, 0, nt);
#pragma unroll
      for (int i = 0; i < IPL; ++i) s_ent[own + i] = nt[i];
      __syncwarp();
    }
    for (; l0 < P.n; l0 += stride) {
      const bool more = l0 + stride < P.n;     // uniform per warp
      if (more) issue(l0 + stride, stage ^ 1, nt);
      if (gs_layout) process(l0, stage, std::true_type{});
      else process(l0, stage, std::false_type{});
      if (more) {
#pragma unroll
        for (int i = 0; i < IPL; ++i) s_ent[own + i] = nt[i];
        __syncwarp();
      }
      stage ^= 1;
    }
  }
  if (bad) atomicOr(P.status, bad);
  if (LOSS) block_loss_epilogue(local, P.partials, P.ticket, P.scale, P.loss, P.loss_sum);
}

template <int LPL, int IPL, int THREADS>
static int launch_tab_cfg(const ListParams& P, bool loss, int num_sms, cudaStream_t st, bool score = false) {
  constexpr size_t SMEM = (size_t)(IPL + 1) * THREADS * (sizeof(float2) + 2 * sizeof(uint32_t));
  static bool configured[3] = {false, false, false};
  const int which = score ? 2 : (loss ? 1 : 0);
  if (!configured[which]) {
    cudaError_t e = score ? cudaFuncSetAttribute(lists_tab_kernel<LPL, IPL, THREADS, false, true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM)
                    : loss ? cudaFuncSetAttribute(lists_tab_kernel<LPL, IPL, THREADS, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM)
                           : cudaFuncSetAttribute(lists_tab_kernel<LPL, IPL, THREADS, false>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(lists_tab_kernel) failed: %s", cudaGetErrorString(e));
      return PLD_ECUDA;
    }
    configured[which] = true;
  }
  constexpr int GPB = THREADS / LPL;
  const int per_image_cap = lists_per_image_cap(num_sms, P.B);
  int gx = (P.n + GPB - 1) / GPB;
  if (gx > per_image_cap) gx = per_image_cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)P.B);
  ListParams Q = P;
  philox_round_keys(P.seed_lo, P.seed_hi, Q.rk0, Q.rk1);
  if (score) PLD_CUDA(launch_pdl(lists_tab_kernel<LPL, IPL, THREADS, false, true>, grid, dim3(THREADS), SMEM, st, Q));
  else if (loss) PLD_CUDA(launch_pdl(lists_tab_kernel<LPL, IPL, THREADS, true>, grid, dim3(THREADS), SMEM, st, Q));
  else PLD_CUDA(launch_pdl(lists_tab_kernel<LPL, IPL, THREADS, false>, grid, dim3(THREADS), SMEM, st, Q));
  PLD_CHECK_LAUNCH();
  return PLD_OK;
}

// SRC_PHILOX_TAB for 17 <= K <= 512
static int launch_tab_any(const ListParams& P, bool loss, bool score, int num_sms, cudaStream_t st) {
  const int K = P.K;
  if (K <= 32) return launch_tab_cfg<4, 8, 256>(P, loss, num_sms, st, score);
  if (K <= 64) return launch_tab_cfg<8, 8, 256>(P, loss, num_sms, st, score);
  if (K <= 128) return launch_tab_cfg<16, 8, 256>(P, loss, num_sms, st, score);
  if (K <= 256) return launch_tab_cfg<32, 8, 256>(P, loss, num_sms, st, score);
  if (K <= 512) return launch_tab_cfg<32, 16, 128>(P, loss, num_sms, st, score);
  set_error("lists_tab: K=%d out of range", K);
  return PLD_EINVAL;
}
int launch_lists_tab(const ListParams& P, bool loss, int num_sms, cudaStream_t st) {
  return launch_tab_any(P, loss, false, num_sms, st);
}
// scoring pass (P.score_keys, P.score_cfg) of the score-based strategies for 17 <= K <= 512
int launch_lists_tab_score(const ListParams& P, int num_sms, cudaStream_t st) {
  return launch_tab_any(P, false, true, num_sms, st);
}

}  // namespace pld
