// Phase stamps for the in-pipeline timelines (tools/debug/coop_timeline.py, tools/microbench/finish_phases.cu).
// NOT part of the product library: only variant builds made by tools/build_variant.sh include this
// header (-DPC_STAMPS_HEADER='"../../tools/debug/stamps.cuh"'); the default build compiles the
// PC_*_CLOCK / PC_PROP_* macros to nothing and exports no pc_debug_* symbol.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pc_dbg_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

#ifdef PC_STAMPS_UKF
__device__ long long g_cstamp[12][256];
__device__ unsigned long long g_gt[8];   // globaltimer: [0] min regular entry, [1] max regular exit, [2] min update entry, [3] max update exit
#define PC_COOP_CLOCK(i)                                                                     \
  do {                                                                                       \
    if ((threadIdx.x & 31) == 0 && blockIdx.x < 256) g_cstamp[i][blockIdx.x] = clock64();    \
    if ((threadIdx.x & 31) == 0 && (i) == 0) atomicMin(&g_gt[2], pc_dbg_gtimer());           \
    if ((threadIdx.x & 31) == 0 && (i) == 9) atomicMax(&g_gt[3], pc_dbg_gtimer());           \
  } while (0)
__device__ long long g_pstamp[10][2048];   // regular path: per-warp phase stamps (clock64)
#define PC_PHASE_CLOCK(i)                                                                    \
  do {                                                                                       \
    if ((threadIdx.x & 31) == 0 && (i) == 0) atomicMin(&g_gt[0], pc_dbg_gtimer());           \
    if ((threadIdx.x & 31) == 0 && (i) == 9) atomicMax(&g_gt[1], pc_dbg_gtimer());           \
    if ((threadIdx.x & 31) == 0) {                                                           \
      const int wid_ = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;                         \
      if (wid_ < 2048) g_pstamp[i][wid_] = clock64();                                        \
    }                                                                                        \
  } while (0)
extern "C" int pc_debug_phase_stamps(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_pstamp, sizeof(g_pstamp)); }
extern "C" int pc_debug_stamps(long long* cst, unsigned long long* gt, int reset) {
  if (cst) cudaMemcpyFromSymbol(cst, g_cstamp, sizeof(g_cstamp));
  if (gt) cudaMemcpyFromSymbol(gt, g_gt, sizeof(g_gt));
  if (reset) {
    unsigned long long z[8] = {~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0};
    cudaMemcpyToSymbol(g_gt, z, sizeof(z));
  }
  return 0;
}
#endif  // PC_STAMPS_UKF

#ifdef PC_STAMPS_PROP
__device__ unsigned long long g_pt[4];
__device__ unsigned long long g_cta[2048][4];   // per CTA: entry, loads landed (thread 0), exit, max substeps
#define PC_PROP_STAMP(i, op) do { if ((threadIdx.x & 31) == 0) op(&g_pt[i], pc_dbg_gtimer()); } while (0)
#define PC_PROP_CTA_ENTRY()                                                                  \
  const int cta_id = blockIdx.y * gridDim.x + blockIdx.x;                                    \
  if (threadIdx.x == 0 && cta_id < 2048) { g_cta[cta_id][0] = pc_dbg_gtimer(); g_cta[cta_id][3] = 0; }
#define PC_PROP_CTA_LOADED(h, r0, v0, ns)                                                    \
  if (cta_id < 2048 && (h) == 0) {                                                           \
    if ((r0) + (v0) != 1e300 && threadIdx.x == 0) g_cta[cta_id][1] = pc_dbg_gtimer();        \
    atomicMax(&g_cta[cta_id][3], (unsigned long long)(ns));                                  \
  }
#define PC_PROP_CTA_EXIT() if (cta_id < 2048) atomicMax(&g_cta[cta_id][2], pc_dbg_gtimer());
extern "C" int pc_debug_prop_ctas(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_cta, sizeof(g_cta));
}
extern "C" int pc_debug_prop_stamps(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_pt, sizeof(g_pt));
  if (reset) {
    unsigned long long z[4] = {~0ull, 0, ~0ull, 0};
    cudaMemcpyToSymbol(g_pt, z, sizeof(z));
  }
  return 0;
}
#endif  // PC_STAMPS_PROP
