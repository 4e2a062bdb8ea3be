// Per-step scheduler glue on the packed maps (first slice of SURVEY.md row N1):
//   * tasking mask from a MultiDiscrete action vector, masked by the ESTIMATED visibility of
//     the previous step  (actionSpace2Array utilities.py:147-172 + SSAScheduler._taskAgents
//     env.py:487-501: target i is tasked iff some sensor j chose i and vis_map_est[i, j] == 1)
//   * float32 observation buffers exactly as SSAScheduler emits them
//     (getEstStates env.py:508-530, est_cov env.py:419-429, obs_staleness env.py:446-452)
#include "pc_common.cuh"

namespace pc {

__global__ void __launch_bounds__(256)
task_from_actions_kernel(const int64_t* __restrict__ actions, int64_t M,
                         const uint32_t* __restrict__ vis_est, int64_t N, int64_t wpr,
                         uint8_t* __restrict__ tasked) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= M) return;
  const int64_t a = actions[j];
  if (a < 0 || a >= N) return;  // a == N is "inaction"
  if ((vis_est[a * wpr + (j >> 5)] >> (j & 31)) & 1u) tasked[a] = 1;
}

__global__ void __launch_bounds__(256)
obs_pack_kernel(const double* __restrict__ sens_x, int64_t M, int64_t ld_s,
                const double* __restrict__ est_x, const double* __restrict__ est_p, int64_t N,
                int64_t ld, const float* __restrict__ last_time_tasked, double t_now_arg,
                const pc_clock* __restrict__ clock, float* __restrict__ out_eci,
                float* __restrict__ out_cov, float* __restrict__ out_stale) {
  const double t_now = clock ? clock->t_now : t_now_arg;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t A = M + N;
  const int64_t n_eci = 6 * A, n_cov = 36 * N;
  if (idx < n_eci) {
    const int64_t c = idx / A, k = idx - c * A;
    out_eci[idx] = (float)(k < M ? sens_x[c * ld_s + k] : est_x[c * ld + (k - M)]);
  } else if (idx < n_eci + n_cov) {
    const int64_t k = idx - n_eci;
    out_cov[k] = (float)est_p[k];
  } else if (idx < n_eci + n_cov + N) {
    const int64_t k = idx - n_eci - n_cov;
    // float(info["time_now"] - agent.last_time_tasked), last_time_tasked is np.float32
    out_stale[k] = (float)(t_now - (double)last_time_tasked[k]);
  }
}

// Stand-in scheduler policy for benchmarks (SURVEY.md section 8d "Actions"): every sensor probes up to
// `probes` uniformly random targets and takes the first one its column of the estimated map
// shows visible, else inaction (action = N).  Counter-based hash, no state.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
policy_random_visible_kernel(const uint32_t* __restrict__ vis_est, int64_t N, int64_t M,
                             int64_t wpr, uint64_t seed, uint64_t step_arg,
                             const pc_clock* __restrict__ clock, int probes,
                             int64_t* __restrict__ actions) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= M) return;
  const uint64_t step = clock ? clock->step : step_arg;
  int64_t pick = N;
  uint64_t h = splitmix64(seed ^ splitmix64(step * 0x100000001B3ull + (uint64_t)j));
  for (int k = 0; k < probes; ++k) {
    h = splitmix64(h);
    const int64_t i = (int64_t)(h % (uint64_t)N);
    if ((vis_est[i * wpr + (j >> 5)] >> (j & 31)) & 1u) {
      pick = i;
      break;
    }
  }
  actions[j] = pick;
}

// FP64 FMA throughput probe: the roofline denominator for the propagate / UKF kernels
// (MEASURED_PEAKS.json holds only HBM and bf16 tensor numbers).
__global__ void __launch_bounds__(256)
dfma_probe_kernel(double* __restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

cudaError_t launch_policy_random_visible(const uint32_t* vis_est, int64_t N, int64_t M, int64_t wpr,
                                         uint64_t seed, uint64_t step, const pc_clock* clock,
                                         int probes, int64_t* actions, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  policy_random_visible_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(
      vis_est, N, M, wpr, seed, step, clock, probes, actions);
  return cudaGetLastError();
}

__global__ void advance_clock_kernel(pc_clock* clock, double dt) {
  clock->t_now += dt;
  clock->step += 1;
}

cudaError_t launch_advance_clock(pc_clock* clock, double dt, cudaStream_t s) {
  advance_clock_kernel<<<1, 1, 0, s>>>(clock, dt);
  return cudaGetLastError();
}

// returns FLOP (2 per DFMA) issued by one launch
double launch_dfma_probe(double* scratch, int blocks, int iters, cudaStream_t s, cudaError_t* err) {
  dfma_probe_kernel<<<blocks, 256, 0, s>>>(scratch, iters, 0.999999999, 1e-9);
  *err = cudaGetLastError();
  return 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
}

cudaError_t launch_task_from_actions(const int64_t* actions, int64_t M, const uint32_t* vis_est,
                                     int64_t N, int64_t wpr, uint8_t* tasked, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(tasked, 0, (size_t)N, s);
  if (e != cudaSuccess || M <= 0) return e;
  task_from_actions_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(actions, M, vis_est, N, wpr,
                                                                      tasked);
  return cudaGetLastError();
}

cudaError_t launch_obs_pack(const double* sens_x, int64_t M, int64_t ld_s, const double* est_x,
                            const double* est_p, int64_t N, int64_t ld,
                            const float* last_time_tasked, double t_now, const pc_clock* clock,
                            float* out_eci, float* out_cov, float* out_stale, cudaStream_t s) {
  const int64_t total = 6 * (M + N) + 36 * N + N;
  if (total <= 0) return cudaSuccess;
  obs_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
      sens_x, M, ld_s, est_x, est_p, N, ld, last_time_tasked, t_now, clock, out_eci, out_cov,
      out_stale);
  return cudaGetLastError();
}

}  // namespace pc
