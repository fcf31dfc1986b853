// Shared device/host helpers for the pldepth_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/pldepth_b200.h"

namespace pld {

// ------------------------------------------------------------------------------------------
// host side: errors, launch accounting, context
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define PLD_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      pld::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,    \
                     __LINE__);                                                            \
      return PLD_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

#define PLD_REQUIRE(cond, msg)                           \
  do {                                                   \
    if (!(cond)) {                                       \
      pld::set_error("invalid argument: %s", msg);       \
      return PLD_EINVAL;                                 \
    }                                                    \
  } while (0)

// every entry point runs on the context's device: refuse a call made while another device is current
#define PLD_CHECK_DEVICE(ctx)                                                                             \
  do {                                                                                                    \
    int dev__ = -1;                                                                                       \
    if (cudaGetDevice(&dev__) != cudaSuccess || dev__ != (ctx)->device) {                                 \
      pld::set_error("context was created on device %d but device %d is current", (ctx)->device, dev__);  \
      return PLD_EINVAL;                                                                                  \
    }                                                                                                     \
  } while (0)

#define PLD_CHECK_LAUNCH()                                                                  \
  do {                                                                                     \
    pld::count_launch();                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      pld::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                            \
      return PLD_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched through launch_pdl may be scheduled while the kernel before
// it in the stream is still draining -- its CTAs become resident once every CTA of that kernel has started, and block
// in pdl_sync() until it has completed and its memory is visible -- so the launch latency and the tail of one kernel
// overlap the ramp-up of the next.  Every kernel launched this way calls pdl_sync() FIRST, on every path (a kernel that
// returned without waiting would let its own dependents run ahead of the grandparent kernel).  PLD_NO_PDL=1 launches
// them plainly (A/B runs).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// CTAs per image of the list kernels: enough CTAs per SM that the hardware scheduler evens out the tail
// (measured at config 2: 8 CTAs/SM 153 us, 4 -> 176 us, 2 -> 202 us; PLD_GRID_MULT overrides for experiments).
int lists_grid_mult();
inline int lists_per_image_cap(int num_sms, int B) { return (num_sms * lists_grid_mult() + B - 1) / B; }

}  // namespace pld

struct pld_ctx {
  int device;
  int num_sms;
  int* d_status;          // device status word (PLD_ST_* bits)
  unsigned int* d_ticket;  // "last block done" ticket for the loss reduction
  double* d_partials;     // per-block loss partials
  int partials_cap;
  void* d_scratch;  // growable scratch (mask chunk counts, MT compaction, radix sort)
  size_t scratch_cap;
  // optional timing of the dominant (list) kernel: ring of event pairs recorded on the launch stream
  cudaEvent_t* ev_start;
  cudaEvent_t* ev_stop;
  int ev_cap, ev_count;
  int deterministic;     // 1: gradients accumulate in 64-bit fixed point (bit-reproducible)
  long long* d_acc;
  size_t acc_cap;
  int ensure_acc(size_t elems);
  unsigned long long* d_offset;  // device-resident Philox offset counter (pld_ctx_device_offset)
  int use_device_offset;
  unsigned int* d_mm_acc;        // [mm_cap, 2] per-image (~ordered min, ordered max) of gt, zero between calls
  int mm_cap;
  int ensure_mm(int B);
  int ensure_scratch(size_t bytes);
  int ensure_partials(int n);
  inline void time_begin(cudaStream_t st) { if (ev_cap > 0 && ev_count < ev_cap) cudaEventRecord(ev_start[ev_count], st); }
  inline void time_end(cudaStream_t st) { if (ev_cap > 0 && ev_count < ev_cap) { cudaEventRecord(ev_stop[ev_count], st); ++ev_count; } }
};

namespace pld {

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
// see launch_pdl: wait for the kernel(s) before this one, then let the kernel after this one be scheduled as soon as
// every CTA of this grid has started
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t u = o ^ ((~(uint32_t)((int32_t)o >> 31)) | 0x80000000u);
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t double_to_ordered(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  return u ^ ((uint64_t)((int64_t)u >> 63) | 0x8000000000000000ull);
}

// Philox4x32-10 (Salmon et al., SC'11).  ctr/key as in Random123.
struct Philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// Same function with the ten round keys precomputed (kernel-parameter constants: they enter the XORs as constant-bank
// operands instead of costing two additions per round)
__device__ __forceinline__ Philox4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    const uint32_t (&rk0)[10], const uint32_t (&rk1)[10]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ rk0[r];
    const uint32_t n2 = hi0 ^ c3 ^ rk1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return Philox4{c0, c1, c2, c3};
}
inline void philox_round_keys(uint32_t k0, uint32_t k1, uint32_t (&rk0)[10], uint32_t (&rk1)[10]) {
  for (int r = 0; r < 10; ++r) { rk0[r] = k0 + (uint32_t)r * 0x9E3779B9u; rk1[r] = k1 + (uint32_t)r * 0xBB67AE85u; }
}

// Draw stream of one list (see DESIGN.md "Philox stream"):
//   first attempt of draw k = word (k & 3) of block (k >> 2);
//   a-th redraw of draw k (Lemire rejection, probability < M / 2^32 each) = word 0 of block
//   0x8000 | ((a - 1) << 9) | k.
struct DrawStream {
  uint32_t list, image, off_lo, off_hi16, k0, k1;
  __device__ __forceinline__ Philox4 block(uint32_t blk) const {
    return philox4x32_10(list, image, blk | off_hi16, off_lo, k0, k1);
  }
};
__device__ __forceinline__ uint32_t lemire_bounded(uint32_t word, uint32_t M, uint32_t thresh,
                                                   const DrawStream& ds, uint32_t k) {
  uint64_t m = (uint64_t)word * (uint64_t)M;
  uint32_t a = 0;
  while ((uint32_t)m < thresh) {  // rare: P < M / 2^32
    ++a;
    Philox4 r = ds.block(0x8000u | (((a - 1u) & 63u) << 9) | k);
    m = (uint64_t)r.x * (uint64_t)M;
  }
  return (uint32_t)(m >> 32);
}

// branch-free first attempt: returns the mapped value and ORs "needs the redraw path" into rej
__device__ __forceinline__ uint32_t lemire_try(uint32_t word, uint32_t M, uint32_t thresh, bool& rej) {
  const uint64_t m = (uint64_t)word * (uint64_t)M;
  rej = rej || ((uint32_t)m < thresh);
  return (uint32_t)(m >> 32);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of per-thread doubles -> partials[blockIdx linear]; the last block to finish
// (ticket) adds all partials in index order and writes the loss.  Deterministic.
__device__ __forceinline__ void block_loss_epilogue(float local, double* partials, unsigned int* ticket,
                                                    float scale, float* loss, double* loss_sum) {
  __shared__ double s_warp[32];
  __shared__ bool s_last;
  double v = (double)warp_sum(local);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) s_warp[wid] = v;
  __syncthreads();
  const unsigned int nblocks = gridDim.x * gridDim.y;
  const unsigned int bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += s_warp[i];
    partials[bid] = t;
    __threadfence();
    unsigned int prev = atomicAdd(ticket, 1u);
    s_last = (prev == nblocks - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // fixed-order tree: thread t sums partials t, t+T, ... then a fixed shuffle/smem tree
    double acc = 0.0;
    for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += __ldcg(partials + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();
    if (lane == 0) s_warp[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < nw; ++i) t += s_warp[i];
      if (loss_sum) *loss_sum = t;
      if (loss) *loss = (float)(t * (double)scale);
      *ticket = 0u;
    }
  }
}

}  // namespace pld
