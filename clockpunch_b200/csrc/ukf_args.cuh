// Argument block of the fused UKF step kernel (kernels_ukf.cu), shared with api.cu.
#pragma once
#include "pc_common.cuh"

namespace pc {

struct UkfArgs {
  double* est_x;
  double* est_p;
  double* truth_x;
  const double* z;
  uint8_t* tasked;          // read; cleared after reading when clear_tasked != 0
  int64_t n, ld;
  double dt, t_now;
  int substeps;
  uint64_t rng_seed, rng_step;
  double* out_pred_x;
  double* out_pred_p;
  double* out_nis;
  uint32_t* out_flags;
  double* out_z;
  int64_t* num_tasked;
  float* last_time_tasked;
  float* tr_pos;
  float* tr_vel;
  pc_clock* clock;         // NULL, or: kernels read t_cur / step_cur (copied by the pre-step kernel)
  int advance_clock;       // block 0 advances clock->t_now / step (nobody reads those in this launch)
  int clear_tasked;        // reset tasked[i] after reading it (graph replay: the mask is rebuilt per step)
  // fused visibility (both maps for the next step); sens_q == NULL disables it
  const double4* sens_q;   // (M) sensor ECI position + horizon term, from the pre-step kernel
  int M, wpr;
  uint32_t* vis_truth;
  uint32_t* vis_est;
  double body_radius, vis_tol;
  pc_ukf_params p;
};

cudaError_t launch_ukf(const UkfArgs& a, int force_model, cudaStream_t stream);

}  // namespace pc
