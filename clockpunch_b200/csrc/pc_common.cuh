// Shared definitions for libpunch_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/punch_b200.h"

namespace pc {

// Reference constants (punchclock/common/constants.py:21-27).  The reference
// deliberately uses the low-precision values below; they are the spec.
constexpr double kMu = 398600.0;
constexpr double kRE = 6378.1363;
constexpr double kOmegaEarth = 7.27e-5;
// J2 is NOT in the reference (dynamics.py:16-40 is two-body only); opt-in.
constexpr double kJ2 = 1.08262668e-3;

// Auto-substep rule: fixed-step DOP853 substeps = ceil(|dt| * Omega / kTheta),
// Omega = sqrt(mu / rp^3) (1 + e)^2 with rp, e from the osculating orbit.
// kTheta is chosen so that the truncation error stays at the FP64 round-off
// floor of the state (measured: DESIGN.md "integrator").
constexpr double kTheta = 0.12;
// Orbit rates above kOmegaMax belong to osculating orbits whose perigee lies deep below the
// surface (a surface-grazing e -> 1 orbit has Omega = 5e-3 rad/s; 2e-2 rad/s is a perigee radius
// of ~2500 km): diverged filter estimates.  Their rate is clamped and the target flagged
// (PC_FLAG_SUBSTEP_CLAMP) instead of integrating a meaningless trajectory to tolerance.
constexpr float kOmegaMax = 2.0e-2f;
constexpr int kMaxAutoSubsteps = 1 << 16;

struct UkfWeights {
  double gamma;    // sqrt(n + lambda)                      ukf_v2.py:117
  double wm0;      // lambda / (n + lambda)                 ukf_v2.py:120
  double wi;       // 1 / (2 (n + lambda))                  ukf_v2.py:121
  double wc0;      // wm0 + 1 - alpha^2 + beta              ukf_v2.py:128
  double eps_w;    // (wm0 + 12 wi) - 1 evaluated exactly from the float64 weights
};

}  // namespace pc

namespace pc {
// Programmatic dependent launch (griddepcontrol): a kernel launched with the
// programmatic-stream-serialization attribute may start while its stream predecessor is still
// running; it must call pdl_wait() before touching anything the predecessor produces.  A kernel
// calls pdl_trigger() to let its successor's CTAs be scheduled as soon as all of its own are
// resident.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Bounded wait of thread r for peer r's flag to reach the number of the push in flight (see PeerPush,
// ukf_args.cuh): table[world + rank] is this rank's flag array, flag[r] the number of the last push
// peer r has published here; *epoch the number of pushes this rank has completed, so the push in
// flight is *epoch + 1 (nobody changes *epoch while a kernel that waits is running).  Ranks step in
// lockstep, so "peer r has published push n" means its summaries of that push are all here.
__device__ __forceinline__ void peer_wait_thread(const uint64_t* __restrict__ table, int world, int rank,
                                                 const unsigned long long* __restrict__ epoch,
                                                 unsigned int* __restrict__ err, int r) {
  const uint64_t want = *reinterpret_cast<const volatile unsigned long long*>(epoch) + 1;
  const volatile uint64_t* flag = reinterpret_cast<const volatile uint64_t*>(__ldg(table + world + rank)) + r;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    if (*flag >= want) return;
    __nanosleep(100);
  }
  atomicOr(err, 1u << r);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, bool programmatic, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = programmatic ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace pc

#define PC_CUDA_TRY(ctx, expr)                                        \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) return pc_fail_cuda((ctx), _e, #expr);     \
  } while (0)

int pc_fail(pc_ctx* ctx, int code, const char* msg);
int pc_fail_cuda(pc_ctx* ctx, cudaError_t e, const char* what);
