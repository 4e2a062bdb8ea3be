/* libpunch_b200 -- C-ABI of the B200-native punchclock hot path.
 *
 * The reference (punchclock v0.8.5) is pure Python and has no FFI layer; its
 * boundary for this path is the Python module API.  Each entry
 * point below names the reference function(s) whose arithmetic it replaces; the
 * Python mirror in clockpunch_b200/ binds them with ctypes.
 *
 * Conventions
 *   - All arithmetic is FP64.  State arrays are structure-of-arrays `(6, ld)`
 *     row-major (component c of agent i at `x[c*ld + i]`), i.e. exactly a C-order
 *     numpy `(6, N)` array when ld == N.  Covariances are `(N, 6, 6)` row-major
 *     (`p[i*36 + a*6 + b]`), exactly the reference's `est_cov` layout.
 *   - Pointers are DEVICE pointers unless the function name ends in `_host`.
 *     The library never frees or retains caller arrays past the call.
 *   - Every call is asynchronous on `stream` (a cudaStream_t; NULL = legacy
 *     default stream); `_host` variants synchronise before returning.
 *   - Return value: 0 = OK, < 0 = pc_status.  `pc_last_error(ctx)` gives text.
 *   - Per-target numerical trouble (non-PD covariance, NaN) is DATA, reported
 *     in `out_flags`, never an error code.
 */
#ifndef PUNCH_B200_H
#define PUNCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pc_ctx pc_ctx;
typedef void* pc_stream; /* cudaStream_t */

enum pc_status {
  PC_OK = 0,
  PC_ERR_BAD_ARG = -1,
  PC_ERR_CUDA = -2,
  PC_ERR_EQUAL_TIMES = -3, /* t0 == tf: reference raises ValueError (propagator.py:40-41) */
  PC_ERR_NO_DEVICE = -4
};

enum pc_agent_kind { PC_KIND_SATELLITE = 0, PC_KIND_TERRESTRIAL = 1 };
enum pc_force_model { PC_FORCE_TWOBODY = 0, PC_FORCE_TWOBODY_J2 = 1 };

enum pc_flag_bits {
  PC_FLAG_NONPD_EST = 1,     /* cholesky(est_p) failed   (reference: nearestPD fallback, ukf_v2.py:170-176) */
  PC_FLAG_NONPD_PRED = 2,    /* cholesky(pred_p) failed in the measurement update */
  PC_FLAG_NONPD_INNOV = 4,   /* innovation covariance not PD */
  PC_FLAG_NONFINITE = 8,     /* NaN/Inf in an output of this target */
  PC_FLAG_SUBSTEP_CLAMP = 16 /* auto substep count hit the clamp */
};

/* UKF constants, evaluated by the caller in the reference's own float64
 * operation order (ukf_v2.py:113-128) so that the weights are bit-identical. */
typedef struct pc_ukf_params {
  double gamma;   /* sqrt(n + lambda) */
  double wm0;     /* mean weight of the centre point */
  double wi;      /* weight of the other 2n points */
  double wc0;     /* covariance weight of the centre point */
  double eps_w;   /* (wm0 + 2n*wi) - 1, exact value for the float64 weights above */
  double Q[36];   /* process noise   (ukf_v2.py:291) */
  double R[36];   /* measurement noise (ukf_v2.py:328) */
  double Lr[36];  /* lower Cholesky factor of R; only used to draw z on device */
} pc_ukf_params;

/* Device-resident step clock: lets a captured CUDA graph of pc_step_fused be replayed without
 * touching kernel arguments.  When pc_step_args.clock != NULL the step runs from clock->t_now to
 * clock->t_now + (tf - t0), draws measurements with clock->step, and advances the clock
 * (t_now += tf - t0, step += 1) for the next replay. */
typedef struct pc_clock {
  double t_now;       /* persistent: time / index of the NEXT step to run */
  uint64_t step;
  double t_cur;       /* copy taken by the running step's first kernel; what its kernels read */
  uint64_t step_cur;
} pc_clock;

/* ---- lifetime -------------------------------------------------------- */
int pc_version(void);
int pc_ctx_create(int device_ordinal, pc_ctx** out);
int pc_ctx_destroy(pc_ctx* ctx);
const char* pc_last_error(pc_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py: gpu_launches) */
int64_t pc_launch_count(pc_ctx* ctx);

/* Kernel-mapping overrides.  The library picks the thread mapping of two kernels from the catalog size;
 * tests (cross-kernel parity at sizes the oracle can follow) and tuning runs can pin the choice per
 * context.  Every mapping evaluates the same equations; value 0 restores the size rule.
 *   PC_OPT_FINISH_KERNEL     per-target stage (transform + update + next sigma points + vis rows):
 *       1 four lanes per target, measurement updates by dedicated warps when pc_step_args.task_of is given
 *       2 four lanes per target, update in line
 *       3 one thread per target, 64-thread CTAs      4 one thread per target, 128-thread CTAs
 *   PC_OPT_PROPAGATE_KERNEL  sigma-point propagation: 1 one state per thread, 2 two states per thread
 *       out of a sorted 256-target window */
enum pc_option { PC_OPT_FINISH_KERNEL = 1, PC_OPT_PROPAGATE_KERNEL = 2 };
int pc_set_option(pc_ctx* ctx, int option, int value);
int pc_get_option(pc_ctx* ctx, int option, int* value);

/* ---- K1: propagation -------------------------------------------------
 * pc_propagate_twobody   replaces SatDynamicsModel.propagate / simplePropagate(satelliteDynamics)
 *                        (dynamics_classes.py:51-76, propagator.py:12-81, dynamics.py:16-40,71-86)
 * pc_rotate_terrestrial  replaces StaticTerrestrial.propagate (dynamics_classes.py:79-121,
 *                        dynamics.py:43-68, transforms.py:55-129)
 * pc_propagate_agents    both, selected per agent by kind[] (Agent.propagate, agents.py:59-71;
 *                        SSAScheduler._propagateAgents, env.py:503-506)
 * substeps: >0 fixed number of DOP853 steps per call; 0 = choose per state
 * from its osculating orbit.  x_out may alias x_in. */
int pc_propagate_twobody(pc_ctx* ctx, const double* x_in, double* x_out, int64_t n, int64_t ld,
                         double t0, double tf, int substeps, int force_model, pc_stream stream);
int pc_rotate_terrestrial(pc_ctx* ctx, const double* x_in, double* x_out, int64_t n, int64_t ld,
                          double t0, double tf, pc_stream stream);
int pc_propagate_agents(pc_ctx* ctx, const double* x_in, double* x_out, const uint8_t* kind,
                        int64_t n, int64_t ld, double t0, double tf, int substeps,
                        int force_model, pc_stream stream);
int pc_propagate_agents_host(pc_ctx* ctx, const double* h_x_in, double* h_x_out,
                             const uint8_t* h_kind, int64_t n, int64_t ld, double t0, double tf,
                             int substeps, int force_model);

/* ---- frames and initial conditions -------------------------------------
 * pc_frame_change replaces ecef2eci (to_eci != 0) / eci2ecef (to_eci == 0)
 *                 (common/transforms.py:55-129,143-178); jd = seconds since the frames were aligned.
 * pc_coe2eci      replaces coe2eci (transforms.py:240-292); coe is (6, ld) rows
 *                 [sma, ecc, inc, raan, argp, true_anom].  Range checks (ValueError in the
 *                 reference, transforms.py:268-279) are done by the caller.
 * pc_lla2eci      replaces ecef2eci(lla2ecef(x), jd) (transforms.py:34-52; the ground-sensor
 *                 IC path of genInitStates, simulation/sim_utils.py:96-101); lla is (3, ld). */
int pc_frame_change(pc_ctx* ctx, const double* x_in, double* x_out, int64_t n, int64_t ld,
                    double jd, int to_eci, pc_stream stream);
int pc_coe2eci(pc_ctx* ctx, const double* coe, double* x_out, int64_t n, int64_t ld, double mu,
               pc_stream stream);
int pc_lla2eci(pc_ctx* ctx, const double* lla, double* x_out, int64_t n, int64_t ld, double jd,
               pc_stream stream);

/* ---- K3: visibility ----------------------------------------------------
 * pc_vismap_packed replaces calcVisMap(binary=True) (common/visibility.py:56-116) and the two
 * calls per step in updateInfoPostTasking (env.py:455-462).  Output is bit-packed:
 * bit (j % 32) of word out[i*words_per_row + j/32] is 1 iff sensor j sees target i.
 * targ_b / out_b may be NULL; when given, a second map (e.g. estimated states) is
 * produced in the same pass.  Sensors with sens_kind[j] == PC_KIND_TERRESTRIAL are
 * rotated about +z by sens_rot_angle while the sensor tile is staged (fused K2);
 * pass sens_kind = NULL for "use as given".
 * pc_vismap_value replaces calcVisMap(binary=False): out is (N, M) float64. */
int pc_vismap_packed(pc_ctx* ctx, const double* sens, int64_t M, int64_t ld_s,
                     const uint8_t* sens_kind, double sens_rot_angle, const double* targ_a,
                     const double* targ_b, int64_t N, int64_t ld_t, double body_radius, double tol,
                     uint32_t* out_a, uint32_t* out_b, int64_t words_per_row, pc_stream stream);
int pc_vismap_value(pc_ctx* ctx, const double* sens, int64_t M, int64_t ld_s, const double* targ,
                    int64_t N, int64_t ld_t, double body_radius, double tol, double* out,
                    pc_stream stream);
/* Host-buffer variant behind calcVisMap(): binary != 0 writes int64 (N, M) into out_i64,
 * otherwise float64 (N, M) into out_f64. */
int pc_vismap_host(pc_ctx* ctx, const double* h_sens, int64_t M, const double* h_targ, int64_t N,
                   double body_radius, double tol, int binary, int64_t* out_i64, double* out_f64);

/* Batch-of-environments forms (vectorised SSAScheduler envs, BASELINE.json configs[2]): E
 * environments stored environment-major, each with env_sensors sensors and env_targets targets.
 * pc_vismap_packed_envs: the E block-diagonal maps, out words (E env_targets, ceil(env_sensors/32)).
 * pc_obs_pack_envs: eci_state (E, 6, env_sensors + env_targets) | est_cov (E, env_targets, 6, 6) |
 * obs_staleness (E, env_targets), i.e. E reference observations back to back. */
int pc_vismap_packed_envs(pc_ctx* ctx, const double* sens, int64_t ld_s, const double* targ_a,
                          const double* targ_b, int64_t ld_t, int64_t E, int64_t env_sensors,
                          int64_t env_targets, double body_radius, double tol, uint32_t* out_a,
                          uint32_t* out_b, pc_stream stream);
int pc_obs_pack_envs(pc_ctx* ctx, const double* sens_x, int64_t ld_s, const double* est_x,
                     const double* est_p, int64_t ld, const float* last_time_tasked, double t_now,
                     const pc_clock* clock, int64_t E, int64_t env_sensors, int64_t env_targets,
                     float* out_eci, float* out_cov, float* out_stale, pc_stream stream);

/* pc_vismap_der replaces calcVisMapAndDerivative (common/visibility.py:119-180): continuous
 * visibility value and its time derivative for every pair, (N, M) float64 each (states need their
 * velocity rows; the radial rates of orbits.py:39-69 are formed in the kernel). */
int pc_vismap_der(pc_ctx* ctx, const double* sens, int64_t M, int64_t ld_s, const double* targ,
                  int64_t N, int64_t ld_t, double body_radius, double tol, double* out_v,
                  double* out_dv, pc_stream stream);
int pc_vismap_der_host(pc_ctx* ctx, const double* h_sens, int64_t M, const double* h_targ, int64_t N,
                       double body_radius, double tol, double* out_v, double* out_dv);

/* pc_vis_history replaces AccessWindowCalculator.calcVisHist (schedule_tree/access_windows.py:148-195,
 * 382-402): the packed visibility map at each of T epochs.  Agents (kind per agent, NULL = all
 * satellites) are advanced IN PLACE from t_now through h_times[0..T-1] (a HOST array; an epoch equal
 * to the current time is evaluated without propagating) -- the reference's own chained
 * Agent.propagate calls -- and out_words receives (T, N, ceil(M/32)) uint32.  T x (N + M) solve_ivp
 * calls and T N M Python pair tests in the reference; 3 T launches here. */
int pc_vis_history(pc_ctx* ctx, double* sens, const uint8_t* sens_kind, int64_t M, int64_t ld_s,
                   double* targ, const uint8_t* targ_kind, int64_t N, int64_t ld_t, double t_now,
                   const double* h_times, int64_t T, int substeps, int force_model,
                   double body_radius, double tol, uint32_t* out_words, pc_stream stream);
int pc_vis_history_host(pc_ctx* ctx, const double* h_sens, const uint8_t* h_sens_kind, int64_t M,
                        const double* h_targ, const uint8_t* h_targ_kind, int64_t N, double t_now,
                        const double* h_times, int64_t T, int substeps, int force_model,
                        double body_radius, double tol, uint32_t* h_out_words);

/* ---- K4: UKF predict + update -------------------------------------------
 * Replaces UnscentedKalmanFilter.predictAndUpdate for the 6-D identity-measurement
 * filter built by ezUKF (ukf_v2.py:185-375, ez_ukf.py:17-95), one filter per target.
 * If truth_x != NULL the target's truth state is propagated in the same launch
 * (Agent.propagate, agents.py:59-71).  tasked[i] != 0 selects the measurement update;
 * the measurement is z[:, i] when z != NULL, else truth' + Lr * N(0, I) drawn with
 * Philox(rng_seed, rng_step, i) (Target.getMeasurement, agents.py:173-183; distributional
 * parity only -- the reference draws from numpy's legacy global RNG). */
int pc_ukf_predict_update(pc_ctx* ctx, double* est_x, double* est_p, double* truth_x,
                          const double* z, const uint8_t* tasked, int64_t n, int64_t ld, double t0,
                          double tf, int substeps, int force_model, const pc_ukf_params* params,
                          uint64_t rng_seed, uint64_t rng_step, double* out_pred_x,
                          double* out_pred_p, double* out_nis, uint32_t* out_flags,
                          double* out_z, pc_stream stream);
/* Stage 1 of the UKF on its own: upper Cholesky of est_p and the 13 sigma points
 * x + gamma [0, U, -U] (generateSigmaPoints, ukf_v2.py:158-183) into blocks 0..12 of a sigma
 * buffer (layout: pc_step_args.sigma_ws).  est_x may be block 0 of sigma_ws itself.  sig_flags is
 * the (2, N) uint32 companion: [i] receives PC_FLAG_NONPD_EST where the factorisation failed,
 * [N + i] the packed rates of the substep rule (see pc_step_args.sig_flags). */
int pc_sigma_points(pc_ctx* ctx, const double* est_x, const double* est_p, int64_t n, int64_t ld,
                    double gamma, double* sigma_ws, uint32_t* sig_flags, pc_stream stream);
int pc_ukf_predict_update_host(pc_ctx* ctx, double* h_est_x, double* h_est_p, const double* h_z,
                               const uint8_t* h_tasked, int64_t n, double t0, double tf,
                               int substeps, int force_model, const pc_ukf_params* params,
                               double* h_pred_x, double* h_pred_p, double* h_nis,
                               uint32_t* h_flags);

/* ---- scheduler glue on the packed maps -----------------------------------
 * pc_task_from_actions replaces actionSpace2Array (common/utilities.py:147-172) + the masking
 *   loop of SSAScheduler._taskAgents (env.py:487-501): tasked[i] = 1 iff some sensor j has
 *   actions[j] == i and bit (i, j) of the ESTIMATED vis map is set (actions[j] == N: inaction).
 * pc_obs_pack writes the float32 observation arrays of SSAScheduler._getObs: eci_state
 *   (6, M+N) = [sensor truth | target estimates] (env.py:508-530), est_cov (N, 6, 6)
 *   (env.py:419-429), obs_staleness (N) = t_now - last_time_tasked (env.py:446-452). */
int pc_task_from_actions(pc_ctx* ctx, const int64_t* actions, int64_t M, const uint32_t* vis_est,
                         int64_t N, int64_t words_per_row, uint8_t* tasked, pc_stream stream);
int pc_obs_pack(pc_ctx* ctx, const double* sens_x, int64_t M, int64_t ld_s, const double* est_x,
                const double* est_p, int64_t N, int64_t ld, const float* last_time_tasked,
                double t_now, const pc_clock* clock /* device or NULL: use t_now */, float* out_eci,
                float* out_cov, float* out_stale, pc_stream stream);

/* Benchmark stand-in policy: each sensor takes the first of `probes` random
 * targets that its column of the estimated map shows visible, else inaction (N). */
int pc_policy_random_visible(pc_ctx* ctx, const uint32_t* vis_est, int64_t N, int64_t M,
                             int64_t words_per_row, uint64_t seed, uint64_t step,
                             const pc_clock* clock /* device, or NULL: use `step` */, int probes,
                             int64_t* actions, pc_stream stream);
/* The same policy on the HOST, for a caller that schedules from the observation it received (bench.py's
 * end-to-end loop): vis_est and actions are host arrays (the packed estimated map as copied out by
 * pc_step_host, the M actions handed to the next pc_step_host).  Same probe sequence as the device form
 * for equal (seed, step).  No context, no device work; returns PC_ERR_BAD_ARG on bad sizes. */
int pc_policy_random_visible_host(const uint32_t* vis_est, int64_t N, int64_t M, int64_t words_per_row,
                                  uint64_t seed, uint64_t step, int probes, int64_t* actions);
/* The same for a batch of E environments stored environment-major (pc_step_args.env_targets / env_sensors):
 * actions[e * env_sensors + j] is LOCAL to environment e (env_targets = inaction); same picks as the device
 * form inside the step for equal (seed, step). */
int pc_policy_random_visible_envs_host(const uint32_t* vis_est, int64_t E, int64_t env_targets, int64_t env_sensors,
                                       int64_t words_per_row, uint64_t seed, uint64_t step, int probes,
                                       int64_t* actions);
/* Host-side expansion of a packed map into the (N, M) int64 array the reference's observation / info carry
 * (`vis_map_est`, `vis_map_truth`: env.py:455-462, 508-530): out[i * M + j] = bit j % 32 of words[i * words_per_row
 * + j / 32].  No device needed; large maps are split over the cores the process may use. */
int pc_unpack_vis_host(const uint32_t* words, int64_t n, int64_t words_per_row, int64_t m, int64_t* out);

/* ---- measurement helpers (bench.py) ------------------------------------
 * pc_set_timing(1) makes pc_step_fused record three CUDA events on the launch stream: before the
 * sigma-point propagation kernel, between it and the per-target kernel, and after the per-target
 * kernel (recorded as external event nodes when the stream is being captured, so they are re-recorded
 * by every replay of the graph); the two kernels then run one after the other instead of chained by
 * programmatic launch.  pc_get_timing waits and returns the device time of the most recent
 * propagation launch in milliseconds, pc_get_kernel_timing that and the per-target kernel's.
 * pc_fp64_peak runs a
 * register-resident DFMA probe and returns the best TFLOP/s of `repeats` launches: the FP64
 * roofline denominator. */
int pc_set_timing(pc_ctx* ctx, int enabled);
int pc_get_timing(pc_ctx* ctx, double* ukf_ms_last);
int pc_get_kernel_timing(pc_ctx* ctx, double* propagate_ms_last, double* finish_ms_last);
int pc_fp64_peak(pc_ctx* ctx, int repeats, double* tflops);

/* ---- fused step ---------------------------------------------------------
 * One SSAScheduler physics step (env.py:533-597 minus host bookkeeping): sensors
 * propagated (kind per sensor), targets' truth + UKF, both packed vis maps, and the
 * per-target scheduler summaries.  All pointers device; see pc_step_args. */
typedef struct pc_step_args {
  /* sensors */
  double* sens_x;            /* (6, ld_s) in/out */
  const uint8_t* sens_kind;  /* (M) */
  int64_t M, ld_s;
  /* targets */
  double* truth_x;           /* (6, ld) in/out */
  double* est_x;             /* (6, ld) in/out */
  double* est_p;             /* (N, 36) in/out */
  const double* z;           /* (6, ld) or NULL -> drawn on device */
  const uint8_t* tasked;     /* (N) or NULL */
  int64_t N, ld;
  double t0, tf;
  int substeps, force_model;
  uint64_t rng_seed, rng_step;
  /* outputs */
  double* pred_p;            /* (N, 36) or NULL */
  double* nis;               /* (N) or NULL */
  uint32_t* flags;           /* (N) or NULL */
  uint32_t* vis_truth;       /* (N, words_per_row) */
  uint32_t* vis_est;         /* (N, words_per_row) */
  int64_t words_per_row;
  double body_radius, vis_tol;
  /* scheduler summaries (Target.updateNonPhysical agents.py:144-171, env.py:437-452,599-609) */
  int64_t* num_tasked;       /* (N) in/out or NULL */
  float* last_time_tasked;   /* (N) in/out or NULL */
  float* tr_pos;             /* (N) trace est_p[:3,:3] or NULL */
  float* tr_vel;             /* (N) trace est_p[3:,3:] or NULL */
  pc_clock* clock;           /* device pointer or NULL (then t0 / tf / rng_step above are used) */
  /* tasking inside the step (SSAScheduler.step: actionSpace2Array + _taskAgents masking,
   * env.py:564-571).  With actions != NULL the step itself builds the tasking mask in `tasked`
   * (which must then be a writable, zero-initialised N-byte workspace; the step leaves it zeroed):
   * policy_probes == 0 reads actions[j] in [0, N] (N = inaction); policy_probes > 0 first fills
   * actions[] with the benchmark stand-in policy (pc_policy_random_visible).  Either way a target
   * is tasked only where the PREVIOUS step's estimated vis map shows the pair visible. */
  int64_t* actions;          /* (M) or NULL */
  int policy_probes;
  int sigma_valid;           /* see sigma_ws */
  uint64_t policy_seed;
  /* Sigma-point buffer carried from step to step: 14 blocks of (6, ld) -- block 0 = centre point =
   * est_x, blocks 1..6 / 7..12 = est_x +- gamma U[:, j], block 13 = truth state.  When given,
   * est_x MUST be sigma_ws and truth_x MUST be sigma_ws + 13*6*ld (the step then copies nothing,
   * and finishes by laying out the sigma points of the new estimate for the next step).
   * sigma_valid != 0 promises that blocks 1..12 already hold the sigma points of the current
   * (est_x, est_p), i.e. nothing touched them since the previous pc_step_fused on these buffers;
   * with 0 they are regenerated first.  sig_flags is a (2, N) uint32 companion workspace: row 0 the
   * flags raised while factoring / propagating, row 1 the per-target rates of the substep rule
   * (two bfloat16: high half the perigee angular rate with its margin, low half the bound for steps
   * of at most 128 s -- all 13 sigma points of a target take the same number of DOP853 substeps).
   * With sigma_ws == NULL a context-owned buffer is used and est_x / truth_x are ordinary arrays. */
  double* sigma_ws;
  uint32_t* sig_flags;
  /* (4, M) workspace, 32-byte aligned: sensor positions + horizon terms at the end of the step,
   * written by the sensor kernel and read by the visibility phase of the target kernel.  NULL: a
   * context-owned buffer is used (allocated on demand, so not usable under stream capture). */
  double* sens_q;
  /* A batch of E independent environments in one catalog (vectorised SSAScheduler envs): targets
   * and sensors are stored environment-major, N = E env_targets, M = E env_sensors; target i is
   * tested only against the sensors of environment i / env_targets, words_per_row =
   * ceil(env_sensors / 32), and actions[j] is LOCAL to the environment of sensor j (0 ..
   * env_targets, env_targets = inaction).  Both 0: one environment (N targets, M sensors). */
  int64_t env_targets, env_sensors;
  /* Multi-GPU exchange fused into the step (one process per GPU, targets partitioned, SURVEY.md
   * section 8e).  peer_table is a DEVICE array of 2 * peer_world pointers: first the base address of every
   * rank's gathered buffer, then of every rank's flag array (peer_world uint64), all peer-mapped over
   * NVLink (pc_comm_* below sets them up with CUDA IPC; any other peer-mapped memory works).  A gathered
   * buffer is [2][peer_world][summary_bytes]: parity of the push, source rank, a copy of that rank's
   * summary buffer.  summary_base / summary_bytes describe the local buffer (16-byte aligned, a multiple
   * of 4 bytes) that contains num_tasked, vis_est, tr_pos, tr_vel and last_time_tasked -- everything
   * another shard's scheduler needs (metrics.py:161, greedy_cov_v2.py:104-120, env.py:599-609).
   * The summaries a step leaves there are needed by the scheduler of the NEXT step and by nothing else of
   * its physics, so the exchange is not the end of step k but a side task of step k + 1: its sensor
   * kernel copies the summary buffer into this rank's slot of EVERY peer's gathered buffer (16-byte
   * stores over NVLink) and publishes the push number in every peer's flag array -- beside the target
   * propagation, which needs none of it -- and the last CTA of the per-target kernel, after its own
   * targets, waits (bounded) until all ranks' flags carry that number.  No collective call; a rank stalls
   * only when a peer is a full step behind.  When pc_step_fused number k completes, the gathered buffers hold every rank's
   * summaries of step k - 1 (parity slot = push count & 1); a caller that needs the summaries of step k
   * before step k + 1 runs (a scheduler on the host) enqueues pc_peer_exchange, the same push + wait as
   * two stand-alone kernels.  pc_peer_status reports (and clears) a bit per peer whose flag never came,
   * and the number of pushes completed so far.  All ranks must step in lockstep (same sequence of steps
   * and exchanges). */
  const uint64_t* peer_table;
  int32_t peer_world, peer_rank;
  const void* summary_base;
  int64_t summary_bytes;
  /* (M) int64 workspace or NULL.  With actions != NULL the sensor kernel records here which target
   * each sensor validly tasks (-1: none), and -- for catalogs small enough that the step is latency
   * bound (N <= 32768, M <= 256) -- the measurement updates of the few tasked targets are done by
   * dedicated warps of the per-target kernel (one warp per tasked target, work split over its lanes)
   * beside the untasked targets instead of lengthening the warps those targets sit in.  Results are
   * the same filter equations either way; without the workspace the in-line update is used. */
  int64_t* task_of;
  /* Host mirror of the step's outputs (used by pc_step_host's zero-copy form; leave NULL otherwise).
   * mirror_host is DEVICE-ACCESSIBLE host memory (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch
   * tensor) laid out like the device block [mirror_dev, mirror_dev + mirror_bytes): every summary
   * array of this struct that lies inside the device block (num_tasked, vis_est, tr_pos, tr_vel,
   * last_time_tasked) is ALSO stored, value by value as the per-target kernel produces it, at the
   * same offset of the host block; obs_cov_host (N, 36) receives the float32 est_cov the same way.
   * Only values that change are written (num_tasked / last_time_tasked of targets that were not
   * measured are not): the caller initialises the host block with one copy of the device block.
   * Supported by the small-catalog kernel with update warps (task_of given, N <= 32768, M <= 256);
   * ignored otherwise. */
  const void* mirror_dev;
  void* mirror_host;
  int64_t mirror_bytes;
  float* obs_cov_host;
} pc_step_args;
int pc_step_fused(pc_ctx* ctx, const pc_step_args* args, const pc_ukf_params* params,
                  pc_stream stream);
int pc_peer_status(pc_ctx* ctx, uint32_t* out_error_mask, uint64_t* out_pushes);
/* Push `summary` (summary_bytes of it; 16-byte aligned, a multiple of 4) into this rank's slot of every
 * peer's gathered buffer and wait (bounded) for every peer's push of the same number, as two kernels on
 * `stream`; what follows on the stream may read the gathered buffer, parity slot = push count & 1.
 * table / world / rank as in pc_step_args.  Every rank must call it (lockstep). */
int pc_peer_exchange(pc_ctx* ctx, const uint64_t* peer_table, int32_t peer_world, int32_t peer_rank,
                     const void* summary, int64_t summary_bytes, pc_stream stream);

/* ---- peer exchange set-up without any framework (pc_comm_init / pc_allgather_step) --
 * One process per GPU on one node.  pc_comm_init allocates this rank's gathered buffer
 * [2][world][summary_bytes] and flag array on the context's device and returns a 64-byte handle
 * (CUDA IPC) for it.  The caller gathers the handles of all ranks with whatever transport it has
 * (MPI, files, torch.distributed, ...) -- the exchange is also the barrier that orders every rank's
 * allocation before anybody's first push -- and hands the world x 64 bytes to pc_comm_attach,
 * which maps every peer's buffer over NVLink and builds the device table.  pc_comm_table then
 * yields what pc_step_args.peer_table / peer_world / peer_rank expect, and the local gathered
 * buffer.  pc_allgather_step is pc_peer_exchange on the attached buffers (summary_bytes of `summary`):
 * the stand-alone form of what the step does in its sensor kernel.  Tear-down: pc_comm_detach on every rank (unmaps the peers' buffers; no rank may still
 * be pushing), synchronise the ranks, then pc_comm_destroy (frees this rank's buffer; also detaches if
 * that has not been done).  pc_ctx_destroy and a second pc_comm_init destroy an existing exchange. */
#define PC_COMM_HANDLE_BYTES 64
int pc_comm_init(pc_ctx* ctx, int32_t world, int32_t rank, int64_t summary_bytes,
                 uint8_t handle_out[PC_COMM_HANDLE_BYTES]);
int pc_comm_attach(pc_ctx* ctx, const uint8_t* all_handles /* world x PC_COMM_HANDLE_BYTES */);
int pc_comm_table(pc_ctx* ctx, const uint64_t** out_table, void** out_gathered, int32_t* out_world,
                  int32_t* out_rank, int64_t* out_summary_bytes);
int pc_allgather_step(pc_ctx* ctx, const void* summary, pc_stream stream);
int pc_comm_detach(pc_ctx* ctx);
int pc_comm_destroy(pc_ctx* ctx);

/* ---- the step as the reference's caller sees it: host actions in, host observation out -------
 * Replaces one SSAScheduler.step(action) (env.py:533-597) for the whole catalog: copies the M
 * MultiDiscrete actions (target index per sensor, N = inaction; utilities.py:147-172) to the device,
 * runs pc_step_fused, packs the float32 observation (pc_obs_pack: eci_state (6, M+N) | est_cov
 * (N, 6, 6) | obs_staleness (N), env.py:261-270) and copies it, the packed estimated vis map and
 * num_tasked back to the host, then synchronises.  When the argument block is replayable (device
 * clock, caller-owned sigma buffer with sigma_valid, measurements drawn on device) the step + pack
 * is captured into a CUDA graph on first use and replayed while the same buffers keep being passed.
 * Host buffers may be pageable; pinned ones make the copies asynchronous to the kernels' tail. */
typedef struct pc_step_io {
  const int64_t* h_actions;  /* (M) host or NULL (actions already in args->actions) */
  float* d_obs;              /* device workspace, 6(M+N) + 37N floats, or NULL (no observation) */
  float* h_obs;              /* host destination of the same size, or NULL */
  uint32_t* h_vis_est;       /* host (N, words_per_row) or NULL */
  int64_t* h_num_tasked;     /* host (N) or NULL */
  /* One-copy form: when the caller has laid its outputs out in ONE device block (e.g. d_obs, the
   * packed vis map and the summaries as adjacent sub-buffers), d_block / h_block / block_bytes make
   * the whole device-to-host transfer a single copy (each small copy costs ~12 us of fixed latency);
   * h_obs / h_vis_est / h_num_tasked are then ignored. */
  const void* d_block;
  void* h_block;
  int64_t block_bytes;
  /* Zero-copy form of the one-copy form: with zero_copy != 0, a device-accessible h_block (see
   * pc_step_args.mirror_host) whose first 6(M+N) + 37N floats mirror d_obs, a replayable argument block and
   * a catalog the update-warp kernel handles (single environment, task_of, N <= 32768, M <= 256), the
   * kernels store the observation and the summaries straight into h_block while they run -- the
   * transfer overlaps the per-target kernel instead of following it -- and no copy is issued (one
   * initialising copy when the graph is (re)built).  With zero_copy != 0 and a device-accessible h_actions
   * the sensor kernel likewise reads the actions from the host buffer itself (args->actions is then not
   * written), so that the step does not start behind a copy.  The device observation d_obs is then NOT refreshed.
   * Falls back to the one-copy form when any condition fails. */
  int32_t zero_copy;
} pc_step_io;
int pc_step_host(pc_ctx* ctx, const pc_step_args* args, const pc_ukf_params* params,
                 const pc_step_io* io, pc_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* PUNCH_B200_H */
