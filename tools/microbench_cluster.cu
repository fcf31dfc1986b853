// Round-2 micro-benchmarks for the one open road to the SMEM-resident design named in north_star: can a thread-block
// cluster whose distributed shared memory (DSMEM) holds one image's lookup table and gradient map serve the random
// per-point accesses of the list kernel faster than the L2 path (1.0 divergent 8-byte gather and 0.66 float reduction
// per clock and SM, profiles/r01_microbench.jsonl)?
//
// Round 1 measured DSMEM at 16 resident warps per SM only (latency-bound: ~215 cycles per remote access).  Here every
// test runs at 32 and 64 resident warps per SM, clusters of 2 / 4 / 8 / 16, and rates are normalised per PARTICIPATING
// SM (clusters that are co-resident in one wave x cluster size), as VERDICT r01 item 8 asks.
//
// Tests (one JSON line each):
//   dsmem_gather_{4,8}B   random loads from a uniformly random CTA of the cluster (ld.shared::cluster)
//   dsmem_red_f32         random float reductions into a random CTA (red.shared::cluster.add.f32)
//   cluster_step          memory skeleton of one list: K gathers (8 B, DSMEM) + K reductions (DSMEM), the cluster holding
//                         table + gradient of one image: 224 x 224 in a 4-CTA cluster (the reference's training size,
//                         pldepth/PLDepth.py:112) and 448 x 448 in a 16-CTA cluster (BASELINE configs 1-4)
//   hybrid_*              gathers from the L2-resident global table + reductions into DSMEM (and the reverse): are the
//                         two paths additive?
//   global_step           the same skeleton entirely through L2 (reference point, same launch shape)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t xs32(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 ld_cluster_v2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void red_cluster_f32(uint32_t a, float v) {
  asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

constexpr int ITERS = 32;
constexpr int UNR = 8;
constexpr int THREADS = 1024;

// mode 0: 4-byte loads, 1: 8-byte loads, 2: float reductions.  `elems` 4-byte words of shared memory per CTA.
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k_dsmem(uint32_t elems, float* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  for (uint32_t i = threadIdx.x; i < elems; i += THREADS) sm[i] = (MODE == 2) ? 0.f : (float)i;
  cluster_sync_all();
  const uint32_t cl = cluster_nctarank();
  const uint32_t base0 = mapa(smem_u32(sm), 0);
  const uint32_t stride = cl > 1 ? mapa(smem_u32(sm), 1) - base0 : 0u;
  const uint32_t n = (MODE == 1) ? elems / 2 : elems;
  uint32_t s = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 31u;
  float acc = 0.f;
  for (int it = 0; it < ITERS; ++it) {
    uint32_t a[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const uint32_t r = xs32(s);
      const uint32_t rank = __umulhi(r, cl);
      const uint32_t idx = __umulhi(r * 2246822519u, n);
      a[u] = base0 + rank * stride + idx * (MODE == 1 ? 8u : 4u);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (MODE == 0) acc += ld_cluster_f32(a[u]);
      else if (MODE == 1) { const float2 v = ld_cluster_v2(a[u]); acc += v.x + v.y; }
      else red_cluster_f32(a[u], 1.0f);
    }
  }
  cluster_sync_all();
  if (acc == 123.456f || sm[threadIdx.x % elems] == 123.456f) out[0] = acc;
}

// Memory skeleton of the list kernel with K points per list.  TAB / GRD: 0 = global (L2), 1 = DSMEM.
//   table: per CTA `tab_elems` float2 in shared memory (cluster holds cl * tab_elems entries) or a global region
//   grad : per CTA `grd_elems` floats                                                      or a global region
template <int K, int TAB, int GRD>
__global__ void __launch_bounds__(THREADS, 1) k_cluster_step(uint32_t tab_elems, uint32_t grd_elems, const float2* __restrict__ gtab,
                                                            float* __restrict__ ggrd, uint32_t region, int lists_per_thread,
                                                            float* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* stab = reinterpret_cast<float2*>(smem_raw);
  float* sgrd = reinterpret_cast<float*>(smem_raw + (TAB ? (size_t)tab_elems * 8 : 0));
  if (TAB) for (uint32_t i = threadIdx.x; i < tab_elems; i += THREADS) stab[i] = make_float2((float)i, 1.0f);
  if (GRD) for (uint32_t i = threadIdx.x; i < grd_elems; i += THREADS) sgrd[i] = 0.f;
  cluster_sync_all();
  const uint32_t cl = cluster_nctarank();
  const uint32_t cluster_id = blockIdx.x / cl;
  const uint32_t tb0 = mapa(smem_u32(stab), 0), gb0 = mapa(smem_u32(sgrd), 0);
  const uint32_t tstride = cl > 1 ? mapa(smem_u32(stab), 1) - tb0 : 0u;
  // the cluster's "image": cl * per-CTA elements when in DSMEM, else `region` global elements
  const uint32_t n_tab = TAB ? cl * tab_elems : region;
  const uint32_t n_grd = GRD ? cl * grd_elems : region;
  const uint32_t npix = n_tab < n_grd ? n_tab : n_grd;
  const float2* gt_ = gtab + (size_t)(cluster_id % 32u) * region;
  float* gg_ = ggrd + (size_t)(cluster_id % 32u) * region;
  uint32_t x = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  // a pixel = (owner CTA, offset inside its slice); per = pixels per CTA slice
  const uint32_t per = (TAB ? tab_elems : (GRD ? grd_elems : npix));
  const uint32_t owners = (TAB || GRD) ? cl : 1u;
  for (int l = 0; l < lists_per_thread; ++l) {
    uint32_t rk[K], of[K];
    float2 t[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      x = x * 1664525u + 1013904223u;
      rk[k] = __umulhi(x, owners);
      of[k] = __umulhi(x * 2246822519u, per);
      if (TAB) t[k] = ld_cluster_v2(tb0 + rk[k] * tstride + of[k] * 8u);
      else t[k] = __ldg(gt_ + (rk[k] * per + of[k]));
    }
    float sacc = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) sacc += t[k].x * t[k].y;
    acc += sacc;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (GRD) red_cluster_f32(gb0 + rk[k] * tstride + of[k] * 4u, sacc * 1e-9f);
      else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gg_ + (rk[k] * per + of[k])), "f"(sacc * 1e-9f) : "memory");
    }
  }
  cluster_sync_all();
  if (acc == 123.456f) out[0] = acc;
}

static int g_sms = 148;
static double g_clock_ghz = 1.9;

struct Launch {
  void* fn;
  int cl;
  int smem;
  int clusters;   // co-resident clusters launched (one wave)
};

static int max_clusters(void* fn, int cl, int smem) {
  if (cl > 8) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl * 64); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static void launch(void* fn, int cl, int nblocks, int smem, void** args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelExC(&cfg, fn, args));
}

template <typename F>
static float best_ms(F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 2; ++i) f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

static void report(const char* name, int cl, int smem_kb, int ctas_per_sm, int clusters, double ops, float ms, const char* unit) {
  const int sms = clusters * cl / ctas_per_sm;
  printf("{\"test\": \"%s\", \"cluster\": %d, \"smem_kb_per_cta\": %d, \"warps_per_sm\": %d, \"clusters_resident\": %d, "
         "\"participating_sms\": %d, \"ms\": %.4f, \"%s_per_clk_per_participating_sm\": %.4f}\n",
         name, cl, smem_kb, ctas_per_sm * THREADS / 32, clusters, sms, ms, unit,
         ops / (ms * 1e-3) / (g_clock_ghz * 1e9) / sms);
  fflush(stdout);
}

template <int MODE>
static void run_dsmem(const char* name, int cl, int smem, float* out) {
  void* fn = (void*)k_dsmem<MODE>;
  int clusters = max_clusters(fn, cl, smem);
  int ctas_per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, fn, THREADS, smem));
  if (clusters <= 0) { printf("{\"test\": \"%s\", \"cluster\": %d, \"skipped\": \"no co-resident cluster\"}\n", name, cl); return; }
  uint32_t elems = (uint32_t)smem / 4;
  void* args[] = {&elems, &out};
  const int nblocks = clusters * cl;
  const float ms = best_ms([&] { launch(fn, cl, nblocks, smem, args); });
  // subtract nothing: fill + cluster syncs are a few microseconds against >= 100 us of accesses
  report(name, cl, smem / 1024, ctas_per_sm, clusters, (double)nblocks * THREADS * ITERS * UNR, ms, "lane_ops");
}

template <int K, int TAB, int GRD>
static void run_step(const char* name, int cl, uint32_t tab_elems, uint32_t grd_elems, const float2* gtab, float* ggrd,
                     uint32_t region, float* out) {
  void* fn = (void*)k_cluster_step<K, TAB, GRD>;
  const int smem = (int)((TAB ? tab_elems * 8 : 0) + (GRD ? grd_elems * 4 : 0)) + 16;
  int clusters = max_clusters(fn, cl, smem);
  if (clusters <= 0) { printf("{\"test\": \"%s\", \"cluster\": %d, \"skipped\": \"no co-resident cluster\"}\n", name, cl); return; }
  int ctas_per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, fn, THREADS, smem));
  int lists_per_thread = 96;
  void* args[] = {&tab_elems, &grd_elems, (void*)&gtab, (void*)&ggrd, &region, &lists_per_thread, &out};
  const int nblocks = clusters * cl;
  const float ms = best_ms([&] { launch(fn, cl, nblocks, smem, args); });
  report(name, cl, smem / 1024, ctas_per_sm, clusters, (double)nblocks * THREADS * lists_per_thread, ms, "lists");
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  g_clock_ghz = prop.clockRate * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", prop.name, g_sms, g_clock_ghz);
  float* out; CK(cudaMalloc(&out, 16));
  const uint32_t REGION = 448 * 448;
  float2* gtab; CK(cudaMalloc(&gtab, sizeof(float2) * (size_t)REGION * 32));
  float* ggrd; CK(cudaMalloc(&ggrd, sizeof(float) * (size_t)REGION * 32));
  CK(cudaMemset(gtab, 0, sizeof(float2) * (size_t)REGION * 32));
  CK(cudaMemset(ggrd, 0, sizeof(float) * (size_t)REGION * 32));

  const int cls[5] = {1, 2, 4, 8, 16};
  for (int smem : {96 * 1024, 200 * 1024}) {
    for (int cl : cls) {
      run_dsmem<0>("dsmem_gather_4B", cl, smem, out);
      run_dsmem<1>("dsmem_gather_8B", cl, smem, out);
      run_dsmem<2>("dsmem_red_f32", cl, smem, out);
    }
  }
  // one image resident in the cluster: (table 8 B + grad 4 B) per pixel
  const uint32_t P224 = 224 * 224, P448 = 448 * 448;
  run_step<5, 1, 1>("cluster_step_K5_224x224_table+grad_in_dsmem", 4, P224 / 4, P224 / 4, gtab, ggrd, P224, out);
  run_step<5, 1, 1>("cluster_step_K5_224x224_table+grad_in_dsmem", 8, P224 / 8, P224 / 8, gtab, ggrd, P224, out);
  run_step<5, 1, 1>("cluster_step_K5_448x448_table+grad_in_dsmem", 16, P448 / 16, P448 / 16, gtab, ggrd, P448, out);
  // hybrids at 448 x 448: only one of the two maps in DSMEM
  run_step<5, 0, 1>("hybrid_K5_448x448_gather_L2_red_dsmem", 4, 0, P448 / 4, gtab, ggrd, P448, out);
  run_step<5, 0, 1>("hybrid_K5_448x448_gather_L2_red_dsmem", 8, 0, P448 / 8, gtab, ggrd, P448, out);
  run_step<5, 0, 1>("hybrid_K5_448x448_gather_L2_red_dsmem", 16, 0, P448 / 16, gtab, ggrd, P448, out);
  run_step<5, 1, 0>("hybrid_K5_448x448_gather_dsmem_red_L2", 8, P448 / 8, 0, gtab, ggrd, P448, out);
  run_step<5, 1, 0>("hybrid_K5_448x448_gather_dsmem_red_L2", 16, P448 / 16, 0, gtab, ggrd, P448, out);
  // single-CTA "cluster": a 224 x 224 gradient map alone fits one SM (196 KB): gathers from L2, reductions local
  run_step<5, 0, 1>("hybrid_K5_224x224_gather_L2_red_local_smem", 1, 0, P224, gtab, ggrd, P224, out);
  // reference points: everything through L2, same launch shape (1024 threads, one CTA per SM)
  run_step<5, 0, 0>("global_step_K5_448x448_L2_only", 1, 0, 0, gtab, ggrd, P448, out);
  run_step<5, 0, 0>("global_step_K5_224x224_L2_only", 1, 0, 0, gtab, ggrd, P224, out);
  return 0;
}
