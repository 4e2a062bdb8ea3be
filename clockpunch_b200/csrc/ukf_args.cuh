// Argument block of the fused UKF step kernel (kernels_ukf.cu), shared with api.cu.
#pragma once
#include "pc_common.cuh"

namespace pc {

// Multi-GPU exchange fused into the per-target stage: every per-target summary the other shards'
// schedulers need (num_tasked, packed estimated vis words, covariance traces, last_time_tasked) is
// stored, as it is produced, into the corresponding slot of EVERY peer's gathered buffer over
// NVLink (peer-mapped symmetric memory), instead of into a local buffer that a collective then
// ships.  Layout of a gathered buffer: [step parity][rank][slot_bytes], a slot being a copy of the
// shard's summary buffer.  The last CTA of the grid publishes the step number in every peer's flag
// array.  Nobody waits at the end of the step: the sensor kernel of the NEXT step waits (bounded) for
// every peer's flag of this push, beside the target propagation (the gathered summaries are only
// consumed by the next step's scheduler); peer_wait_kernel is the stand-alone form of that wait.
struct PeerPush {
  int world, rank;             // world == 0: disabled
  const uint64_t* table;       // device array: [0, world) gathered bases, [world, 2 world) flag bases
  int64_t slot_bytes;
  int64_t off_num_tasked, off_vis_est, off_tr_pos, off_tr_vel, off_ltt;   // field offsets inside a slot
  unsigned int* ticket;        // device counter for "last CTA done"
  unsigned long long* seq;     // pushes completed so far (never reset: engine resets must not rewind the flags)
};

struct UkfArgs {
  double* sb;              // sigma buffer: 14 blocks of (6, ld)
  double* est_p;           // (N, 36) out
  const double* z;
  uint8_t* tasked;         // read; cleared after reading when clear_tasked != 0
  int clear_tasked;
  int64_t n, ld;
  double dt, t_now;
  uint64_t rng_seed, rng_step;
  pc_clock* clock;         // NULL, or: kernels read t_cur / step_cur (latched by the pre-step kernel)
  int advance_clock;       // thread 0 advances clock->t_now / step (nobody reads those in this launch)
  double* out_pred_x;
  double* out_pred_p;
  double* out_nis;
  uint32_t* out_flags;
  double* out_z;
  int64_t* num_tasked;
  float* last_time_tasked;
  float* tr_pos;
  float* tr_vel;
  uint32_t* sig_flags;     // in: flags raised while generating / propagating; out: next generation
  int gen_next;            // write the next step's sigma points (est_x = block 0) instead of est_x_out
  double* est_x_out;       // (6, ld) when !gen_next
  double* truth_out;       // (6, ld) copy of block 13, or NULL
  // fused visibility phase (vis_truth == NULL: skipped): sensor positions + horizon terms at the
  // end of the step, as written by the pre-step kernel
  const double4* sens_q;
  int64_t M, wpr;
  // batch of environments: target i belongs to environment i / env_targets and is tested against
  // that environment's env_sensors sensors; a single catalog has env_targets = 0, env_sensors = M
  int64_t env_targets, env_sensors;
  uint32_t* vis_truth;
  uint32_t* vis_est;
  double body_radius, vis_tol;
  // cooperative measurement updates (quad kernel only; task_of == NULL: in-line update): task_of[j]
  // = global index of the target sensor j validly tasks this step, or -1 (written by the pre-step
  // kernel); sensors of one environment are task_env_sensors consecutive entries
  const int64_t* task_of;
  int64_t task_sensors, task_env_sensors;
  // host mirrors (mapped pinned memory, NULL = off; update-warp kernel only): values are stored here
  // as well as in the device arrays, so that the PCIe transfer runs under the kernel
  float* h_cov;              // (N, 36) float32 est_cov
  int64_t* h_num_tasked;
  uint32_t* h_vis_est;
  float* h_tr_pos;
  float* h_tr_vel;
  float* h_ltt;
  PeerPush pp;
  int variant;             // pc_option PC_OPT_FINISH_KERNEL (0 = choose by catalog size)
  pc_ukf_params p;
};

cudaError_t launch_sigma_gen(const double* est_x, const double* est_p, const double* truth_x, double* sb,
                             int64_t n, int64_t ld, double gamma, uint32_t* sig_flags,
                             cudaStream_t stream);
cudaError_t launch_propagate_sigma(double* sb, int64_t n, int64_t ld, int have_truth, double t0,
                                   double tf, int substeps, int force_model, uint32_t* sig_flags,
                                   bool beside_predecessor, int mapping, cudaStream_t stream);
cudaError_t launch_ukf_finish(const UkfArgs& a, cudaStream_t stream);
bool finish_fills_host_mirrors(int64_t n, int64_t sensors, int variant);
cudaError_t launch_peer_push(const void* summary, int64_t bytes, const PeerPush& pp, bool programmatic,
                             cudaStream_t stream);
cudaError_t launch_peer_wait(const uint64_t* table, int world, int rank, const unsigned long long* seq,
                             unsigned int* err, bool programmatic, cudaStream_t stream);

}  // namespace pc
