// Candidate scores of the score-based strategies, reproducing NumPy's arithmetic operation for
// operation (no FMA contraction): sampling.py:161-167 (masked), 194-205 (thresholded), 219-237
// (information), get_depth_relation depth_utils.py:5-21.  Shared by the staged score kernel
// (runtime K, lists read from memory) and the fused list kernel (K <= 16, list in registers).
#pragma once
#include "pld_common.cuh"

namespace pld {

template <typename T> struct Arith;
template <> struct Arith<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Arith<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

struct ScoreCfg {
  const float* gt_minmax;          // [B, 2] (information): min, max of gt -- or, when null,
  const unsigned int* gt_minmax_enc;  // [B, 2] (~ordered(min), ordered(max)) as collected by prep_build_kernel
  int strategy;                    // PLD_STRATEGY_*
  int promotion;                   // PLD_PROMOTION_*
  double thr_hi, thr_lo, penalty;  // legacy (float64) thresholds
  float thr_hi_f, thr_lo_f;        // nep50 (float32) thresholds
  // float32 relation test without a division (relation_equal): for a >= b > 0 the ratio r = RN(a / b) is >= 1 > thr_lo,
  // and  a < b * c_in  =>  r < thr_hi,   a > b * c_out  =>  r >= thr_hi   (c = thr_hi (1 -+ 2^-20): far outside the
  // half-ulp in which rounding decides); anything in between takes the exact division.  0 = off.
  float c_in, c_out;
  int fast_rel;
};

inline ScoreCfg make_score_cfg(const float* gt_minmax, int strategy, double threshold, double penalty, int promotion) {
  ScoreCfg c;
  c.gt_minmax = gt_minmax;
  c.gt_minmax_enc = nullptr;
  c.strategy = strategy;
  c.promotion = promotion;
  c.thr_hi = 1.0 + threshold;          // depth_utils.py:16
  c.thr_lo = 1.0 / (1.0 + threshold);  // depth_utils.py:18
  c.thr_hi_f = (float)c.thr_hi;
  c.thr_lo_f = (float)c.thr_lo;
  c.penalty = penalty;
  c.c_in = (float)(c.thr_hi * (1.0 - 9.5367431640625e-07));
  c.c_out = (float)(c.thr_hi * (1.0 + 9.5367431640625e-07));
  c.fast_rel = (c.thr_hi_f > 1.0f && c.thr_lo_f < 1.0f && c.thr_hi_f < 1e30f && c.c_in < c.thr_hi_f && c.c_out > c.thr_hi_f) ? 1 : 0;
  return c;
}

// the exact float32 relation of a ratio a / c that the margin test of relation_equal could not decide (one in ~2^19);
// out of line: kernels that unroll the test dozens of times must not carry a division per copy
static __device__ __noinline__ bool relation_equal_exact_f32(float a, float c, float thr_hi, float thr_lo) {
  const float r = __fdiv_rn(a, c);
  return !(r >= thr_hi) && !(r <= thr_lo);
}

template <typename T>
__device__ __forceinline__ bool relation_equal(float g1, float g2, const ScoreCfg& P) {
  if (sizeof(T) == 4) {
    const float a = __fadd_rn(g1, 1e-10f), c = __fadd_rn(g2, 1e-10f);
    // decided without dividing in all but ~2^-19 of the cases (ScoreCfg::c_in); NaNs fail every test and divide
    const bool in = a < __fmul_rn(c, P.c_in), out = a > __fmul_rn(c, P.c_out);
    if (P.fast_rel != 0 && a >= c && c > 1e-30f && a < 1e30f && (in || out)) return in;
    return relation_equal_exact_f32(a, c, P.thr_hi_f, P.thr_lo_f);
  } else {
    const double r = __ddiv_rn(__dadd_rn((double)g1, 1e-10), __dadd_rn((double)g2, 1e-10));
    return !(r >= P.thr_hi) && !(r <= P.thr_lo);
  }
}

// element k (0-based) of linspace(start, stop, K+1)[1:]  (numpy/_core/function_base.py)
template <typename T>
__device__ __forceinline__ T ladder(int k, int K, T start, T stop, T delta, T step) {
  if (k == K - 1) return stop;
  const T i = (T)(k + 1);
  if (step == (T)0) return Arith<T>::add(Arith<T>::mul(Arith<T>::div(i, (T)K), delta), start);
  return Arith<T>::add(Arith<T>::mul(i, step), start);
}

template <typename T>
__device__ __forceinline__ void ladder_setup(const ScoreCfg& C, int b, int K, T& start, T& stop, T& delta, T& step) {
  float mn, mx;
  if (C.gt_minmax != nullptr) { mn = C.gt_minmax[b * 2]; mx = C.gt_minmax[b * 2 + 1]; }
  else { mn = ordered_to_float(~C.gt_minmax_enc[b * 2]); mx = ordered_to_float(C.gt_minmax_enc[b * 2 + 1]); }
  if (sizeof(T) == 4) start = (T)__fadd_rn(mn, 0.001f);      // sampling.py:223
  else start = (T)__dadd_rn((double)mn, 0.001);
  stop = (T)mx;
  delta = Arith<T>::sub(stop, start);
  step = Arith<T>::div(delta, (T)K);
}

// the K expected depths of one image (information strategy), computed once instead of per list
template <typename T>
__device__ __forceinline__ void fill_ladder(const ScoreCfg& C, int b, int K, int k, T* out) {
  T start, stop, delta, step;
  ladder_setup<T>(C, b, K, start, stop, delta, step);
  out[k] = ladder<T>(k, K, start, stop, delta, step);
}

// NumPy pairwise summation (umath loops_utils.h pairwise_sum): n < 8 sequential; n <= 128 eight
// running accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail; larger n
// split at n/2 rounded down to a multiple of 8.
template <typename T, typename F>
__device__ T np_pairwise_sum(const F& f, int lo, int n) {
  using A = Arith<T>;
  if (n < 8) {
    T res = (T)0;
    for (int i = 0; i < n; ++i) res = A::add(res, f(lo + i));
    return res;
  }
  if (n <= 128) {
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = A::add(r[j], f(lo + i + j));
    }
    T res = A::add(A::add(A::add(r[0], r[1]), A::add(r[2], r[3])),
                   A::add(A::add(r[4], r[5]), A::add(r[6], r[7])));
    for (; i < n; ++i) res = A::add(res, f(lo + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  const T a = np_pairwise_sum<T, F>(f, lo, n2);
  const T b = np_pairwise_sum<T, F>(f, lo + n2, n - n2);
  return A::add(a, b);
}

// Combination step of a list score from its per-position parts (runtime K), in NumPy's order of operations:
//   chi(k)  = (g_k - e_k)^2 / e_k            (information, type T)        -- see chi_term
//   diff(j) = |g_j - g_{j+1}|                (masked / thresholded, float32)
//   eq(j)   = get_depth_relation(g_j, g_{j+1}) == 0
template <typename T, typename FC, typename FD, typename FE>
__device__ __forceinline__ double score_combine(const FC& chi, const FD& diff, const FE& eq, int K, const ScoreCfg& C) {
  double score;
  if (C.strategy == PLD_STRATEGY_INFORMATION) {
    const T sum = np_pairwise_sum<T>(chi, 0, K);
    score = (double)(-sum);
    for (int j = 0; j + 1 < K; ++j)
      if (eq(j)) score = __dadd_rn(score, C.penalty);
  } else {
    T acc = (T)0;
    const T pen = (T)C.penalty;
    for (int j = 0; j + 1 < K; ++j) {
      if (C.strategy == PLD_STRATEGY_THRESHOLDED && eq(j)) acc = Arith<T>::add(acc, pen);
      acc = Arith<T>::add(acc, (T)diff(j));
    }
    score = (double)acc;
  }
  return score;
}

template <typename T>
__device__ __forceinline__ T chi_term(float g, int k, int K, T start, T stop, T delta, T step) {
  const T e = ladder<T>(k, K, start, stop, delta, step);
  const T d = Arith<T>::sub((T)g, e);
  return Arith<T>::div(Arith<T>::mul(d, d), e);
}

// Score of a list whose depths are read through `g(k)` (runtime K).
template <typename T, typename G>
__device__ __forceinline__ double score_list(const G& g, int K, const ScoreCfg& C, int b) {
  T start = (T)0, stop = (T)0, delta = (T)0, step = (T)0;
  if (C.strategy == PLD_STRATEGY_INFORMATION) ladder_setup<T>(C, b, K, start, stop, delta, step);
  auto chi = [&](int k) { return chi_term<T>(g(k), k, K, start, stop, delta, step); };
  auto diff = [&](int j) { return fabsf(__fsub_rn(g(j), g(j + 1))); };
  auto eq = [&](int j) { return relation_equal<T>(g(j), g(j + 1), C); };
  return score_combine<T>(chi, diff, eq, K, C);
}

// Same arithmetic for a list held in registers (compile-time K <= 16, fully unrolled).
// `lad`: the image's ladder (fill_ladder) when the caller precomputed it, else null.
template <typename T, int K>
__device__ __forceinline__ double score_regs(const float (&g)[K], const ScoreCfg& C, int b, const T* lad = nullptr) {
  static_assert(K <= 16, "register scoring supports K <= 16");
  using A = Arith<T>;
  double score;
  if (C.strategy == PLD_STRATEGY_INFORMATION) {
    T start = (T)0, stop = (T)0, delta = (T)0, step = (T)0;
    if (lad == nullptr) ladder_setup<T>(C, b, K, start, stop, delta, step);
    T chi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const T e = lad != nullptr ? lad[k] : ladder<T>(k, K, start, stop, delta, step);
      const T d = A::sub((T)g[k], e);
      chi[k] = A::div(A::mul(d, d), e);
    }
    T sum;
    if constexpr (K < 8) {
      sum = (T)0;
#pragma unroll
      for (int k = 0; k < K; ++k) sum = A::add(sum, chi[k]);
    } else {
      T r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = chi[j];
      if constexpr (K == 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = A::add(r[j], chi[8 + j]);
      }
      sum = A::add(A::add(A::add(r[0], r[1]), A::add(r[2], r[3])), A::add(A::add(r[4], r[5]), A::add(r[6], r[7])));
#pragma unroll
      for (int k = K - (K % 8); k < K; ++k) sum = A::add(sum, chi[k]);
    }
    score = (double)(-sum);
#pragma unroll
    for (int j = 0; j + 1 < K; ++j)
      if (relation_equal<T>(g[j], g[j + 1], C)) score = __dadd_rn(score, C.penalty);
  } else {
    T acc = (T)0;
    const T pen = (T)C.penalty;
#pragma unroll
    for (int j = 0; j + 1 < K; ++j) {
      const float diff = fabsf(__fsub_rn(g[j], g[j + 1]));
      if (C.strategy == PLD_STRATEGY_THRESHOLDED && relation_equal<T>(g[j], g[j + 1], C)) acc = A::add(acc, pen);
      acc = A::add(acc, (T)diff);
    }
    score = (double)acc;
  }
  return score;
}

// order-preserving u64 image of a score (== what np.argsort compares; -0.0 == +0.0)
__device__ __forceinline__ uint64_t score_key(double s) {
  if (s == 0.0) s = 0.0;
  return double_to_ordered(s);
}
// Scores of the masked / thresholded strategies under NEP-50 promotion are float32 values stored in a float64
// array: their order is carried by 32 bits.  Placing those in the HIGH word leaves the low four key bytes
// constant, so the radix sort skips four of its eight passes and the 36-bit selection prefix is exact.
__device__ __forceinline__ uint64_t score_key_f32(float s) {
  if (s == 0.f) s = 0.f;
  uint32_t u = __float_as_uint(s);
  u ^= ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
  return (uint64_t)u << 32;
}

}  // namespace pld
