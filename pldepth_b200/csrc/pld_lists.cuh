// Per-list device building blocks: draws, ordering networks, ListMLE forward/backward.
#pragma once
#include "pld_common.cuh"
#include "pld_score.cuh"

namespace pld {

enum { SRC_PHILOX = 0, SRC_FED_SEL = 1, SRC_FED_RANK = 2, SRC_PHILOX_TAB = 3 };

struct ListParams {
  const float* gt;             // [B, HW]
  const int32_t* valid_flat;   // [B, valid_stride]
  const int32_t* n_valid;      // [B]
  const float* pred;           // [B, HW]
  const int32_t* sel_in;       // [B, n, K]   (SRC_FED_SEL)
  const float* rank_in;        // [B, n, K, 2] (SRC_FED_RANK)
  const float2* table;         // [B, table_stride] (SRC_PHILOX_TAB), see prep_build_kernel
  size_t table_stride;
  float* rank_out;             // [B, n, K, 2] nullable
  uint64_t* score_keys;        // [B, n] ordered scores (score mode of the small kernel)
  unsigned int* sel_hist;      // [B, 4096] nullable: histogram of the top 12 key bits (first pass of the radix top-R)
  ScoreCfg score_cfg;
  // optional indirection: list l of image b redraws candidate list_map[b*map_stride + l]
  const uint32_t* list_map;
  size_t map_stride;
  int32_t* sel_out;            // [B, n, K] nullable
  float* per_list;             // [B*n] nullable
  float* grad;                 // [B, HW] nullable
  long long* acc;              // [B, HW] fixed-point (2^-32) accumulators: deterministic mode, else null
  // valid-index mode (holed masks, rankings not emitted): tables hold (gt, pred) by valid index for EVERY image and
  // the gradient of a holed image accumulates by valid index here [B, table_stride]; expanded to pixels afterwards
  float* grad_valid;
  float* loss;                 // [1]
  double* loss_sum;            // [1] nullable
  double* partials;
  unsigned int* ticket;
  int* status;
  int B, HW, valid_stride, n, K;
  uint32_t seed_lo, seed_hi, off_lo, off_hi16;
  uint32_t rk0[10], rk1[10];             // Philox round keys seed + r * Weyl (filled by the launchers that use them)
  const unsigned long long* offset_dev;  // when set, the Philox offset is read from device memory (graph replay)
  int image_base;
  float scale;
};

#define PLD_LOG_EPS (-23.025850929940457f) /* float32(log(1e-10)), TF-Ranking _EPSILON */

// Philox offset of this launch: by value, or from the context's device counter (CUDA-graph friendly)
__device__ __forceinline__ void launch_offset(const ListParams& P, uint32_t& off_lo, uint32_t& off_hi16) {
  off_lo = P.off_lo;
  off_hi16 = P.off_hi16;
  if (P.offset_dev != nullptr) {
    const unsigned long long o = *P.offset_dev;
    off_lo = (uint32_t)o;
    off_hi16 = (uint32_t)((o >> 32) & 0xFFFFull) << 16;
  }
}

// K Philox draws of list l of image (image_base + b), mapped to [0, M)
// RK: Philox round keys taken from the kernel parameters (P.rk0 / P.rk1, constant-bank operands) instead of being
// advanced in registers: 40 instructions fewer per list, which pays in the ALU-heavy scoring kernels; the
// gather-bound loss kernel measured 6 % SLOWER with it (80 instead of 93 registers -> a third resident CTA per SM),
// so it keeps the register form.
template <int K, bool RK = false>
__device__ __forceinline__ void draw_philox(const ListParams& P, uint32_t off_lo, uint32_t off_hi16, int b, int l,
                                            uint32_t M, uint32_t thresh, int (&sel)[K]) {
  const DrawStream ds{(uint32_t)l, (uint32_t)(P.image_base + b), off_lo, off_hi16, P.seed_lo, P.seed_hi};
  bool rej = false;
#pragma unroll
  for (int q = 0; q < (K + 3) / 4; ++q) {
    const Philox4 r = RK ? philox4x32_10_rk(ds.list, ds.image, (uint32_t)q | off_hi16, off_lo, P.rk0, P.rk1)
                         : ds.block((uint32_t)q);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = q * 4 + j;
      if (k < K) sel[k] = (int)lemire_try(w[j], M, thresh, rej);
    }
  }
  if (rej) {  // rare (P < K * M / 2^32): redo with the redraw stream
#pragma unroll
    for (int q = 0; q < (K + 3) / 4; ++q) {
      const Philox4 r = ds.block((uint32_t)q);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = q * 4 + j;
        if (k < K) sel[k] = (int)lemire_bounded(w[j], M, thresh, ds, (uint32_t)k);
      }
    }
  }
}

// fire-and-forget float add into the dense gradient map (RED, no return value)
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
#ifdef PLD_USE_ATOMG
  atomicAdd(addr, v);
#else
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
#endif
}

// gradient contribution of one point: float RED (fast) or 64-bit fixed-point atomic (integer addition is
// associative, so the dense gradient is bit-reproducible); `g` is the unscaled d nll / d score
__device__ __forceinline__ void grad_add(const ListParams& P, float* dst_base, size_t acc_base, int idx, float g) {
  if (P.acc != nullptr) {
    atomicAdd(reinterpret_cast<unsigned long long*>(P.acc) + acc_base + (size_t)idx,
              (unsigned long long)__double2ll_rn((double)g * 4294967296.0));
  } else {
    red_add_f32(dst_base + idx, g * P.scale);
  }
}

// compare-exchange, descending (a >= b afterwards)
__device__ __forceinline__ void ce_desc(uint64_t& a, uint64_t& b) {
  const bool sw = a < b;
  const uint64_t hi = sw ? b : a, lo = sw ? a : b;
  a = hi; b = lo;
}
__device__ __forceinline__ void ce_desc(uint64_t& a, uint64_t& b, uint32_t& pa, uint32_t& pb) {
  const bool sw = a < b;
  const uint64_t hi = sw ? b : a, lo = sw ? a : b;
  const uint32_t ph = sw ? pb : pa, pl = sw ? pa : pb;
  a = hi; b = lo; pa = ph; pb = pl;
}

// Register sorting network for K keys, descending.  K == 5 uses the optimal 9-comparator
// network; other K use the insertion network (K(K-1)/2 comparators, fully unrolled).
template <int K, bool PAYLOAD>
__device__ __forceinline__ void sort_desc_regs(uint64_t (&key)[K], uint32_t (&pay)[K]) {
#define PLD_CE(i, j)                                        \
  do {                                                      \
    if (PAYLOAD) ce_desc(key[i], key[j], pay[i], pay[j]);   \
    else ce_desc(key[i], key[j]);                           \
  } while (0)
  if constexpr (K == 5) {
    PLD_CE(0, 1); PLD_CE(3, 4); PLD_CE(2, 4); PLD_CE(2, 3); PLD_CE(1, 4);
    PLD_CE(0, 3); PLD_CE(0, 2); PLD_CE(1, 3); PLD_CE(1, 2);
  } else {
#pragma unroll
    for (int i = 1; i < K; ++i) {
#pragma unroll
      for (int j = i; j > 0; --j) PLD_CE(j - 1, j);
    }
  }
#undef PLD_CE
}

// The same networks on bare depths (descending), two FMNMX per comparator: enough for the scoring pass, whose result
// depends on the ordered depths only (tied depths are interchangeable there).
template <int K>
__device__ __forceinline__ void sort_desc_floats(float (&g)[K]) {
#define PLD_CEF(i, j)                                  \
  do {                                                 \
    const float hi_ = fmaxf(g[i], g[j]), lo_ = fminf(g[i], g[j]); \
    g[i] = hi_; g[j] = lo_;                            \
  } while (0)
  if constexpr (K == 5) {
    PLD_CEF(0, 1); PLD_CEF(3, 4); PLD_CEF(2, 4); PLD_CEF(2, 3); PLD_CEF(1, 4);
    PLD_CEF(0, 3); PLD_CEF(0, 2); PLD_CEF(1, 3); PLD_CEF(1, 2);
  } else {
#pragma unroll
    for (int i = 1; i < K; ++i) {
#pragma unroll
      for (int j = i; j > 0; --j) PLD_CEF(j - 1, j);
    }
  }
#undef PLD_CEF
}

// ListMLE on K scores already in sorted (label-descending) order, all in registers.
// nll = sum_k log(S_k) - (s_k - m),  S_k = sum_{j>=k} exp(s_j - m)   (reverse cumsum)
// g_k = exp(s_k - m) * sum_{i<=k} 1/S_i - 1
template <int K>
__device__ __forceinline__ float listmle_regs(const float (&s)[K], float (&g)[K]) {
  float m = s[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m = fmaxf(m, s[k]);
  float e[K], S[K];
#pragma unroll
  for (int k = 0; k < K; ++k) e[k] = expf(s[k] - m);
  float run = 0.f;
#pragma unroll
  for (int k = K - 1; k >= 0; --k) { run += e[k]; S[k] = run; }
  float nll = 0.f, c = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // the last term is log(e_k) - (s_k - m) == 0 analytically; keep TF's arithmetic anyway
    nll += logf(S[k]) - (s[k] - m);
    c += 1.0f / S[k];
    g[k] = e[k] * c - 1.0f;
  }
  return nll;
}

}  // namespace pld
