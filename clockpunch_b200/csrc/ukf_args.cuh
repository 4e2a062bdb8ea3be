// Argument block of the fused UKF step kernel (kernels_ukf.cu), shared with api.cu.
#pragma once
#include "pc_common.cuh"

namespace pc {

// Multi-GPU exchange of the per-target summaries (num_tasked, packed estimated vis words, covariance
// traces, last_time_tasked -- one contiguous summary buffer per shard): every rank copies its summary
// buffer into its slot of EVERY peer's gathered buffer over NVLink (peer-mapped memory, plain 16-byte
// stores) and publishes the push number in every peer's flag array; a rank then waits (bounded) until
// all peers' flags carry that number.  Layout of a gathered buffer: [push parity][rank][slot_bytes].
// The summaries of step k are needed by the scheduler of step k + 1 and by nothing else of step
// k + 1's physics, so the push of step k's summaries sits in the SENSOR kernel of step k + 1, which runs
// beside the target propagation, and the wait for the peers' pushes is the last thing the LAST CTA of
// step k + 1's per-target kernel does (after its own targets): a rank stalls only when a peer is a full
// step behind, and the exchange is off the critical path of the step (profiles/README.md).
struct PeerPush {
  int world, rank;             // world == 0: disabled
  const uint64_t* table;       // device array: [0, world) gathered bases, [world, 2 world) flag bases
  const void* src;             // this rank's summary buffer (16-byte aligned)
  int64_t bytes;               // its size = slot size (multiple of 4)
  unsigned int* ticket;        // device counter for "last push CTA done"
  unsigned long long* epoch;   // pushes completed so far (never rewound: engine resets must not rewind the flags);
                               // the push in flight is number *epoch + 1, counted by whoever waited for it (the per-
                               // target kernel's last CTA, or peer_wait_kernel) once every rank's flag carries it
  unsigned int* err;           // bit r: peer r's flag never arrived
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
  PeerPush peer;           // multi-GPU (world == 0: off): the kernel's last CTA ends with the wait for the peers' pushes
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
cudaError_t launch_peer_push(const PeerPush& pp, cudaStream_t stream);
cudaError_t launch_peer_wait(const uint64_t* table, int world, int rank, unsigned long long* epoch,
                             unsigned int* err, int advance, cudaStream_t stream);

// End of the per-target kernel, called by every thread of its LAST CTA after the CTA's own work: wait
// (bounded) for every rank's flag of the push in flight, then count the push.  The sensor kernel that
// pushed completed before this kernel started, so nobody else reads *epoch any more.
static __device__ __noinline__ void peer_wait_and_count(const PeerPush& pp) {
  if ((int)threadIdx.x < pp.world) peer_wait_thread(pp.table, pp.world, pp.rank, pp.epoch, pp.err, (int)threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) *pp.epoch += 1;
}

// CTAs of a push: enough 16-byte stores in flight to fill NVLink for megabyte summaries, one CTA for small ones
inline int peer_push_blocks(int64_t bytes) {
  const int64_t want = (bytes / 16 + 128 * 8 - 1) / (128 * 8);   // eight 16-byte elements per thread and pass
  return (int)(want < 1 ? 1 : (want > 64 ? 64 : want));
}

// One CTA (128 threads) of a push: copy its share of the summary buffer into this rank's slot of every
// peer's gathered buffer, make the stores visible system wide, and let the last CTA publish the push
// number in every peer's flag array.
__device__ __forceinline__ void peer_push_block(const PeerPush& pp, int block, int nblocks) {
  const uint64_t n = *reinterpret_cast<const volatile unsigned long long*>(pp.epoch) + 1;
  const int64_t off = ((int64_t)(n & 1) * pp.world + pp.rank) * pp.bytes;
  const int64_t n16 = pp.bytes / 16;
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(pp.src);
  // eight loads in flight per thread before the first store: the buffer comes cold from HBM (~1 us per
  // dependent round trip), and a load-store-load-store loop was the whole cost of the push (6 of 8 us)
  const int64_t stride = (int64_t)nblocks * 128;
  for (int64_t i0 = (int64_t)block * 128 + threadIdx.x; i0 < n16; i0 += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i0 + u * stride < n16) v[u] = src[i0 + u * stride];
    for (int r = 0; r < pp.world; ++r) {
      uint4* dst = reinterpret_cast<uint4*>(__ldg(pp.table + r) + off);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * stride < n16) dst[i0 + u * stride] = v[u];
    }
  }
  if (block == 0) {   // tail words (slot sizes are multiples of 4 bytes, not always of 16)
    const int64_t w0 = n16 * 4, nw = pp.bytes / 4;
    const uint32_t* __restrict__ s32 = reinterpret_cast<const uint32_t*>(pp.src);
    for (int64_t i = w0 + threadIdx.x; i < nw; i += 128) {
      const uint32_t v = s32[i];
      for (int r = 0; r < pp.world; ++r) reinterpret_cast<uint32_t*>(__ldg(pp.table + r) + off)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(pp.ticket, 1u);
    if (done == (unsigned)nblocks - 1) {
      *pp.ticket = 0;                   // ready for the next push (everybody else has already arrived)
      __threadfence_system();
      for (int r = 0; r < pp.world; ++r)
        *reinterpret_cast<volatile uint64_t*>(__ldg(pp.table + pp.world + r) + 8 * (uint64_t)pp.rank) = n;
    }
  }
}

}  // namespace pc
