// Per-step scheduler glue on the packed maps (first slice of SURVEY.md row N1):
//   * tasking mask from a MultiDiscrete action vector, masked by the ESTIMATED visibility of
//     the previous step  (actionSpace2Array utilities.py:147-172 + SSAScheduler._taskAgents
//     env.py:487-501: target i is tasked iff some sensor j chose i and vis_map_est[i, j] == 1)
//   * float32 observation buffers exactly as SSAScheduler emits them
//     (getEstStates env.py:508-530, est_cov env.py:419-429, obs_staleness env.py:446-452)
#include "propagate.cuh"
#include "ukf_args.cuh"
#include "vis.cuh"

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
  pdl_wait();  // launched programmatically behind the per-target stage
  const double t_now = clock ? clock->t_now : t_now_arg;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t A = M + N;
  const int64_t n_eci = 6 * A, n_cov = out_cov ? 36 * N : 0;   // out_cov == NULL: written elsewhere
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

// The same observation for a batch of E environments, one block per environment:
// eci_state (E, 6, Me + Ne) = per environment [its sensors' truth | its targets' estimates];
// est_cov (E, Ne, 6, 6) and obs_staleness (E, Ne) coincide with the flat layouts.
__global__ void __launch_bounds__(256)
obs_pack_envs_kernel(const double* __restrict__ sens_x, int64_t ld_s, const double* __restrict__ est_x,
                     const double* __restrict__ est_p, int64_t N, int64_t ld,
                     const float* __restrict__ last_time_tasked, double t_now_arg,
                     const pc_clock* __restrict__ clock, int64_t E, int64_t Me, int64_t Ne,
                     float* __restrict__ out_eci, float* __restrict__ out_cov,
                     float* __restrict__ out_stale) {
  const double t_now = clock ? clock->t_now : t_now_arg;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t Ae = Me + Ne;
  const int64_t n_eci = 6 * E * Ae, n_cov = 36 * N;
  if (idx < n_eci) {
    const int64_t e = idx / (6 * Ae), rem = idx - e * 6 * Ae;
    const int64_t c = rem / Ae, k = rem - c * Ae;
    out_eci[idx] = (float)(k < Me ? sens_x[c * ld_s + e * Me + k] : est_x[c * ld + e * Ne + (k - Me)]);
  } else if (idx < n_eci + n_cov) {
    const int64_t k = idx - n_eci;
    out_cov[k] = (float)est_p[k];
  } else if (idx < n_eci + n_cov + N) {
    const int64_t k = idx - n_eci - n_cov;
    out_stale[k] = (float)(t_now - (double)last_time_tasked[k]);
  }
}

cudaError_t launch_obs_pack_envs(const double* sens_x, int64_t ld_s, const double* est_x,
                                 const double* est_p, int64_t N, int64_t ld, const float* last_time_tasked,
                                 double t_now, const pc_clock* clock, int64_t E, int64_t Me, int64_t Ne,
                                 float* out_eci, float* out_cov, float* out_stale, cudaStream_t s) {
  const int64_t total = 6 * E * (Me + Ne) + 36 * N + N;
  if (total <= 0) return cudaSuccess;
  obs_pack_envs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
      sens_x, ld_s, est_x, est_p, N, ld, last_time_tasked, t_now, clock, E, Me, Ne, out_eci, out_cov, out_stale);
  return cudaGetLastError();
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

__global__ void __launch_bounds__(128)
policy_random_visible_kernel(const uint32_t* __restrict__ vis_est, int64_t N, int64_t M,
                             int64_t wpr, uint64_t seed, uint64_t step_arg,
                             const pc_clock* __restrict__ clock, int probes,
                             int64_t* __restrict__ actions) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);  // one warp per sensor
  if (j >= M) return;
  const uint64_t step = clock ? clock->step : step_arg;
  const uint64_t base = splitmix64(seed ^ splitmix64(step * 0x100000001B3ull + (uint64_t)j));
  int64_t pick = N;
  for (int k0 = 0; k0 < probes && pick == N; k0 += 32) {
    const int k = k0 + lane;
    const int64_t i = (int64_t)(splitmix64(base + (uint64_t)k) % (uint64_t)N);
    const bool hit = k < probes && ((vis_est[i * wpr + (j >> 5)] >> (j & 31)) & 1u);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) pick = __shfl_sync(0xffffffffu, i, __ffs(m) - 1);
  }
  if (lane == 0) actions[j] = pick;
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
  policy_random_visible_kernel<<<(unsigned)((M + 3) / 4), 128, 0, s>>>(
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

// ---- pre-step kernel: everything the fused UKF kernel needs from the M sensors ---------------
//   blocks [0, nb_sens): one thread per sensor -- Agent.propagate (agents.py:59-71; terrestrial
//       ones are a rotation), result written back and, with its horizon term, to sens_q (4, M)
//   blocks [nb_sens, ...): one WARP per sensor -- tasking: the sensor's action (given, or the
//       stand-in policy: the first of `probes` random targets its column of the PREVIOUS
//       estimated map shows visible, probed 32 at a time) marks tasked[target]
//   block 0 thread 0 also latches the step clock (t_cur / step_cur) for the kernels that follow.
//   Multi-GPU: blocks [nb_sens + nb_task, ...) copy this rank's summary buffer -- the per-target
//   summaries the PREVIOUS step left there -- into every peer's gathered buffer and publish the push
//   number.  The target propagation runs beside this kernel and only the per-target stage waits for it,
//   so the push of step k's summaries overlaps the propagation of step k + 1 instead of ending step k;
//   the wait for the peers' pushes is the last thing the last CTA of the per-target kernel does.  (A peer that is a step ahead writes the OTHER parity slot; it cannot be two
//   steps ahead, because its own wait needs this rank's push of the step in between.)
struct PreStepArgs {
  double* sens_x;
  const uint8_t* sens_kind;
  int64_t M, ld_s;
  double dt;
  int substeps;
  double cs, sn;            // rotation of Earth-fixed sensors over dt
  double4* sens_q;          // may be NULL
  double body_radius, vis_tol;
  pc_clock* clock;          // may be NULL
  // tasking (actions == NULL: skip)
  int64_t* actions;
  const uint32_t* vis_est;
  uint8_t* tasked;
  int64_t N, wpr;
  int probes;
  uint64_t seed, step_arg;
  int nb_sens;
  int64_t env_targets, env_sensors;  // batch of environments (single catalog: N, M)
  int64_t* task_of;         // may be NULL; [j] = global index of the target sensor j validly tasks, else -1
  // multi-GPU (pp.world == 0: off): push the summaries of the PREVIOUS step into every peer (blocks
  // [nb_sens + nb_task, + nb_push)); the wait for the peers' pushes ends the per-target kernel's last CTA
  PeerPush pp;
  int nb_task, nb_push;
};

template <bool J2>
__global__ void __launch_bounds__(128)
pre_step_kernel(const __grid_constant__ PreStepArgs a) {
  pdl_trigger();  // the target propagation launched behind this kernel does not depend on it
  if ((int)blockIdx.x < a.nb_sens) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.clock) {
      a.clock->t_cur = a.clock->t_now;
      a.clock->step_cur = a.clock->step;
    }
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j < a.M) {
      double r[3], v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        r[c] = a.sens_x[(int64_t)c * a.ld_s + j];
        v[c] = a.sens_x[(int64_t)(c + 3) * a.ld_s + j];
      }
      if (a.sens_kind[j] == PC_KIND_TERRESTRIAL) {
        rotate_z(r, v, a.cs, a.sn);
      } else {
        const int nsub = a.substeps > 0 ? a.substeps : auto_substeps(r, v, a.dt, nullptr);
        propagate_state<J2>(r, v, a.dt, nsub);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        a.sens_x[(int64_t)c * a.ld_s + j] = r[c];
        a.sens_x[(int64_t)(c + 3) * a.ld_s + j] = v[c];
      }
      if (a.sens_q)
        a.sens_q[j] = make_double4(r[0], r[1], r[2], horizon_q(r[0], r[1], r[2], a.body_radius, a.vis_tol));
    }
    return;
  }
  if ((int)blockIdx.x >= a.nb_sens + a.nb_task) {
    peer_push_block(a.pp, (int)blockIdx.x - a.nb_sens - a.nb_task, a.nb_push);
    return;
  }
  // tasking: sensor j of environment e = j / env_sensors addresses that environment's targets
  // with LOCAL indices 0 .. env_targets (env_targets = inaction), exactly the MultiDiscrete action
  // space of one SSAScheduler (env.py:172-179)
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)((int)blockIdx.x - a.nb_sens) * 4 + (threadIdx.x >> 5);
  if (j >= a.M || a.actions == nullptr) return;
  const int64_t Ne = a.env_targets, Me = a.env_sensors;
  const int64_t e = j / Me, jl = j - e * Me, t0 = e * Ne;
  int64_t pick = Ne;
  if (a.probes > 0) {
    const uint64_t step = a.clock ? a.clock->step : a.step_arg;
    const uint64_t base = splitmix64(a.seed ^ splitmix64(step * 0x100000001B3ull + (uint64_t)j));
    for (int k0 = 0; k0 < a.probes && pick == Ne; k0 += 32) {
      const int k = k0 + lane;
      const int64_t i = (int64_t)(splitmix64(base + (uint64_t)k) % (uint64_t)Ne);
      const bool hit = k < a.probes && ((a.vis_est[(t0 + i) * a.wpr + (jl >> 5)] >> (jl & 31)) & 1u);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) pick = __shfl_sync(0xffffffffu, i, __ffs(m) - 1);
    }
    if (lane == 0) {
      a.actions[j] = pick;
      if (pick < Ne) a.tasked[t0 + pick] = 1;  // visible by construction
      if (a.task_of) a.task_of[j] = pick < Ne ? t0 + pick : -1;
    }
  } else if (lane == 0) {
    pick = a.actions[j];
    const bool valid =
        pick >= 0 && pick < Ne && ((a.vis_est[(t0 + pick) * a.wpr + (jl >> 5)] >> (jl & 31)) & 1u);
    if (valid) a.tasked[t0 + pick] = 1;
    if (a.task_of) a.task_of[j] = valid ? t0 + pick : -1;
  }
}

cudaError_t launch_pre_step(double* sens_x, const uint8_t* sens_kind, int64_t M, int64_t ld_s, double t0,
                            double tf, int substeps, int force_model, double* sens_q,
                            double body_radius, double vis_tol, pc_clock* clock, int64_t* actions,
                            const uint32_t* vis_est, uint8_t* tasked, int64_t N, int64_t wpr,
                            int probes, uint64_t seed, uint64_t step, int64_t env_targets,
                            int64_t env_sensors, int64_t* task_of, const PeerPush& pp, cudaStream_t s) {
  if (M <= 0 && clock == nullptr && pp.world <= 0) return cudaSuccess;
  PreStepArgs a;
  a.task_of = task_of;
  a.pp = pp;
  a.env_targets = env_targets > 0 ? env_targets : N;
  a.env_sensors = env_sensors > 0 ? env_sensors : M;
  a.sens_x = sens_x;
  a.sens_kind = sens_kind;
  a.M = M;
  a.ld_s = ld_s;
  a.dt = tf - t0;
  a.substeps = substeps;
  a.cs = cos(kOmegaEarth * a.dt);
  a.sn = sin(kOmegaEarth * a.dt);
  a.sens_q = reinterpret_cast<double4*>(sens_q);
  a.body_radius = body_radius;
  a.vis_tol = vis_tol;
  a.clock = clock;
  a.actions = actions;
  a.vis_est = vis_est;
  a.tasked = tasked;
  a.N = N;
  a.wpr = wpr;
  a.probes = probes;
  a.seed = seed;
  a.step_arg = step;
  a.nb_sens = (int)((M + 127) / 128);
  if (a.nb_sens < 1) a.nb_sens = 1;
  a.nb_task = actions ? (int)((M + 3) / 4) : 0;
  a.nb_push = pp.world > 0 ? peer_push_blocks(pp.bytes) : 0;
  const int blocks = a.nb_sens + a.nb_task + a.nb_push;
  if (force_model == PC_FORCE_TWOBODY_J2)
    pre_step_kernel<true><<<blocks, 128, 0, s>>>(a);
  else
    pre_step_kernel<false><<<blocks, 128, 0, s>>>(a);
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
  const int64_t total = 6 * (M + N) + (out_cov ? 36 * N : 0) + N;
  if (total <= 0) return cudaSuccess;
  return launch_pdl(obs_pack_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, true, sens_x, M, ld_s,
                    est_x, est_p, N, ld, last_time_tasked, t_now, clock, out_eci, out_cov, out_stale);
}

}  // namespace pc
