// Building blocks of the group-per-list kernels (ranking_size 17..512): LPL lanes x IPL register slots per list.
// Ordering networks over the LPL*IPL slots (slot e = lane*IPL + i) and group-wide scans by warp shuffles.
#pragma once
#include "pld_lists.cuh"

namespace pld {

// Bitonic network over LPL*IPL slots (slot e = lane*IPL + i), descending, keys as (hi, lo)
// 32-bit halves.  All keys are distinct (pads: all zero, interchangeable), so "take the partner"
// is a single predicate and the whole network is branch-free SEL code.
template <int LPL, int IPL, bool PAYLOAD>
__device__ __forceinline__ void bitonic_desc(uint32_t (&khi)[IPL], uint32_t (&klo)[IPL], uint32_t (&pay)[IPL], int gl) {
  constexpr int N = LPL * IPL;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= IPL) {
        const int lm = j / IPL;
        const bool lower = (gl & lm) == 0;
        const bool up = (gl & (k / IPL)) == 0;  // k/IPL == LPL on the last merge -> always up
        const bool keep_max = (up == lower);
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          const uint32_t ohi = __shfl_xor_sync(0xffffffffu, khi[i], lm);
          const uint32_t olo = __shfl_xor_sync(0xffffffffu, klo[i], lm);
          const bool gt = (((uint64_t)ohi << 32) | olo) > (((uint64_t)khi[i] << 32) | klo[i]);
          const bool take = (gt == keep_max);
          if (PAYLOAD) {
            const uint32_t op = __shfl_xor_sync(0xffffffffu, pay[i], lm);
            pay[i] = take ? op : pay[i];
          }
          khi[i] = take ? ohi : khi[i];
          klo[i] = take ? olo : klo[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < IPL; ++i) {
          if ((i & j) == 0) {
            const bool up = (((gl * IPL + i) & k) == 0);
            const uint32_t ah = khi[i], al = klo[i], bh = khi[i | j], bl = klo[i | j];
            const bool lt = (((uint64_t)ah << 32) | al) < (((uint64_t)bh << 32) | bl);
            const bool sw = (lt == up);
            khi[i] = sw ? bh : ah;
            klo[i] = sw ? bl : al;
            khi[i | j] = sw ? ah : bh;
            klo[i | j] = sw ? al : bl;
            if (PAYLOAD) {
              const uint32_t pa = pay[i], pb = pay[i | j];
              pay[i] = sw ? pb : pa;
              pay[i | j] = sw ? pa : pb;
            }
          }
        }
      }
    }
  }
}

template <int LPL>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = LPL / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPL>
__device__ __forceinline__ float group_min(float v) {
#pragma unroll
  for (int o = LPL / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPL>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum of `v` over lanes of the group with a larger / smaller group-lane index
template <int LPL>
__device__ __forceinline__ float group_excl_suffix(float v, int gl) {
  float x = __shfl_down_sync(0xffffffffu, v, 1, LPL);
  if (gl + 1 >= LPL) x = 0.f;
#pragma unroll
  for (int d = 1; d < LPL; d <<= 1) {
    const float t = __shfl_down_sync(0xffffffffu, x, d, LPL);
    if (gl + d < LPL) x += t;
  }
  return x;
}
template <int LPL>
__device__ __forceinline__ float group_excl_prefix(float v, int gl) {
  float x = __shfl_up_sync(0xffffffffu, v, 1, LPL);
  if (gl == 0) x = 0.f;
#pragma unroll
  for (int d = 1; d < LPL; d <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, x, d, LPL);
    if (gl >= d) x += t;
  }
  return x;
}

// MUFU approximations without the denormal pre/post-scaling of __expf / __logf / __fdividef (each of those costs 6-9
// instructions per element; results differ only for denormal arguments, which flush to zero)
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ListMLE of one list spread over LPL lanes x IPL slots, scores `sv` in sorted (label-descending) order; sorted position
// gl*IPL + i is real for i < nreal.  Returns the list's NLL on every lane of the group and d nll / d score in `g`.
//   nll = sum_k log(S_k) - (s_k - m),  S_k = sum_{j>=k} exp(s_j - m);   g_k = exp(s_k - m) * sum_{i<=k} 1/S_i - 1
// evaluated in base 2 (one FFMA + MUFU.EX2 per exponential): t_k = (s_k - m) log2 e, e_k = 2^t_k,
// nll = ln 2 * sum_k (log2 S_k - t_k).  Both group-per-list kernels call this, so their results are bit-identical.
template <int LPL, int IPL>
__device__ __forceinline__ float group_listmle(float (&sv)[IPL], int nreal, int gl, float (&g)[IPL]) {
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  float t2[IPL], S[IPL];
  float m = -3.402823466e38f;
#pragma unroll
  for (int i = 0; i < IPL; ++i) {
    sv[i] = (i < nreal) ? sv[i] : -3.402823466e38f;
    m = fmaxf(m, sv[i]);
  }
  m = group_max<LPL>(m);
  const float ml = m * LOG2E;
  float run = 0.f;
#pragma unroll
  for (int i = IPL - 1; i >= 0; --i) {
    t2[i] = fmaf(sv[i], LOG2E, -ml);   // pads: -inf
    g[i] = fast_ex2(t2[i]);            // pads: 0
    run += g[i];
    S[i] = run;
  }
  const float carry = group_excl_suffix<LPL>(run, gl);
  float nll2 = 0.f, c = 0.f;
  float cl[IPL];
#pragma unroll
  for (int i = 0; i < IPL; ++i) {
    S[i] += carry;
    nll2 += (i < nreal) ? (fast_lg2(S[i]) - t2[i]) : 0.f;
    c += (i < nreal) ? fast_rcp(S[i]) : 0.f;
    cl[i] = c;
  }
  const float cpre = group_excl_prefix<LPL>(c, gl);
#pragma unroll
  for (int i = 0; i < IPL; ++i) g[i] = fmaf(g[i], cl[i] + cpre, -1.0f);
  return group_sum<LPL>(nll2) * LN2;
}

}  // namespace pld
