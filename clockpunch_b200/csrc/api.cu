// C-ABI of libpunch_b200 (see include/punch_b200.h for the contract).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <sched.h>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "pc_common.cuh"
#include "ukf_args.cuh"

namespace pc {
cudaError_t launch_propagate_agents(const double*, double*, const uint8_t*, int, int64_t, int64_t,
                                    double, double, int, int, cudaStream_t);
cudaError_t launch_vismap_packed(const double*, int64_t, int64_t, const uint8_t*, double,
                                 const double*, const double*, int64_t, int64_t, double, double,
                                 uint32_t*, uint32_t*, int64_t, cudaStream_t);
cudaError_t launch_vismap_value(const double*, int64_t, int64_t, const double*, int64_t, int64_t,
                                double, double, double*, cudaStream_t);
cudaError_t launch_vismap_envs(const double*, int64_t, const double*, const double*, int64_t, int64_t, int64_t,
                               int64_t, double, double, uint32_t*, uint32_t*, cudaStream_t);
cudaError_t launch_obs_pack_envs(const double*, int64_t, const double*, const double*, int64_t, int64_t,
                                 const float*, double, const pc_clock*, int64_t, int64_t, int64_t, float*,
                                 float*, float*, cudaStream_t);
cudaError_t launch_vismap_der(const double*, int64_t, int64_t, const double*, int64_t, int64_t, double,
                              double, double*, double*, cudaStream_t);
cudaError_t launch_pre_step(double* sens_x, const uint8_t* sens_kind, int64_t M, int64_t ld_s, double t0,
                            double tf, int substeps, int force_model, double* sens_q,
                            double body_radius, double vis_tol, pc_clock* clock, int64_t* actions,
                            const uint32_t* vis_est, uint8_t* tasked, int64_t N, int64_t wpr,
                            int probes, uint64_t seed, uint64_t step, int64_t env_targets,
                            int64_t env_sensors, int64_t* task_of, const PeerPush& pp, cudaStream_t s);

cudaError_t launch_frame_change(const double*, double*, int64_t, int64_t, double, double, cudaStream_t);
cudaError_t launch_coe2eci(const double*, double*, int64_t, int64_t, double, cudaStream_t);
cudaError_t launch_lla2eci(const double*, double*, int64_t, int64_t, double, cudaStream_t);
cudaError_t launch_task_from_actions(const int64_t*, int64_t, const uint32_t*, int64_t, int64_t,
                                     uint8_t*, cudaStream_t);
cudaError_t launch_obs_pack(const double*, int64_t, int64_t, const double*, const double*, int64_t,
                            int64_t, const float*, double, const pc_clock*, float*, float*, float*,
                            cudaStream_t);
cudaError_t launch_policy_random_visible(const uint32_t*, int64_t, int64_t, int64_t, uint64_t,
                                         uint64_t, const pc_clock*, int, int64_t*, cudaStream_t);
double launch_dfma_probe(double*, int, int, cudaStream_t, cudaError_t*);
cudaError_t launch_advance_clock(pc_clock*, double, cudaStream_t);
cudaError_t launch_unpack_bits(const uint32_t* words, int64_t N, int64_t M, int64_t wpr,
                               int64_t* out, cudaStream_t stream);
}  // namespace pc

struct pc_ctx {
  int device = 0;
  std::string err;
  int64_t launches = 0;
  // device workspace for the *_host entry points (grown on demand, freed with the ctx)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  cudaStream_t host_stream = nullptr;
  // sigma-point workspace for UKF calls whose caller does not bring one (not graph-capture safe)
  double* sb = nullptr;
  uint32_t* sbf = nullptr;
  size_t sb_states = 0;
  // optional per-kernel timing of pc_step_fused (bench.py roofline leg)
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;  // before / between / after the propagation and the per-target kernel
  // sensor positions + horizon terms for the fused visibility phase when the caller brings no
  // pc_step_args.sens_q (not graph-capture safe)
  double* sq = nullptr;
  size_t sq_sensors = 0;
  // peer push: "last CTA done" ticket and the wait kernel's error mask (device), see PeerPush
  unsigned int* peer_dev = nullptr;   // [0] ticket, [1] error mask, [2..3] 64-bit push sequence number

  // pc_step_host: captured steps (pc_step_fused + pc_obs_pack) and the arguments each was captured
  // with, replayed while a caller keeps passing the same buffers.  Several engines share one context
  // (deep-copied environments do: the reference's wrappers and SimRunner copy the env every step), so
  // the cache holds a few graphs, least recently used one evicted.
  struct HsEntry {
    cudaGraphExec_t exec = nullptr;
    pc_step_args args;
    pc_ukf_params params;
    pc_step_io io;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    uint64_t last_use = 0;
  };
  static constexpr int kHsEntries = 8;
  HsEntry hs[kHsEntries];
  uint64_t hs_tick = 0;
  const void* hs_zc_ptr = nullptr;   // last host block checked for device accessibility (zero-copy form)
  bool hs_zc_ok = false;
  const void* hs_act_ptr = nullptr;  // the same for the host actions buffer
  bool hs_act_ok = false;
  // kernel-mapping overrides (pc_set_option)
  int opt_finish = 0, opt_prop = 0;
  // pc_comm_*: this rank's gathered buffer + flags (one cudaMalloc, shared by CUDA IPC), the peers'
  // mappings and the device table pc_step_args.peer_table points at
  struct Comm {
    int world = 0, rank = 0;
    int64_t slot_bytes = 0;
    void* base = nullptr;              // [flags: 64 x uint64][gathered: 2 x world x slot_bytes]
    std::vector<void*> opened;         // peers' bases as mapped here (own rank: base)
    uint64_t* table = nullptr;         // device: [world] gathered bases, [world] flag bases
    bool attached = false;
  } comm;
};

int pc_fail(pc_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}
int pc_fail_cuda(pc_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx) ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
  return PC_ERR_CUDA;
}

namespace {
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

int ensure_ws(pc_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return PC_OK;
  if (ctx->ws) cudaFree(ctx->ws);
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  size_t want = bytes + (bytes >> 2) + 4096;
  PC_CUDA_TRY(ctx, cudaMalloc(&ctx->ws, want));
  ctx->ws_bytes = want;
  return PC_OK;
}
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int ensure_sq(pc_ctx* ctx, int64_t M) {
  if ((size_t)M <= ctx->sq_sensors) return PC_OK;
  if (ctx->sq) cudaFree(ctx->sq);
  ctx->sq = nullptr;
  ctx->sq_sensors = 0;
  const size_t want = (size_t)M + (size_t)(M >> 2) + 64;
  PC_CUDA_TRY(ctx, cudaMalloc(&ctx->sq, 4 * want * sizeof(double)));
  ctx->sq_sensors = want;
  return PC_OK;
}

int ensure_sb(pc_ctx* ctx, int64_t ld) {
  if ((size_t)ld <= ctx->sb_states) return PC_OK;
  if (ctx->sb) cudaFree(ctx->sb);
  if (ctx->sbf) cudaFree(ctx->sbf);
  ctx->sb = nullptr;
  ctx->sbf = nullptr;
  ctx->sb_states = 0;
  const size_t want = (size_t)ld + (size_t)(ld >> 2) + 64;
  PC_CUDA_TRY(ctx, cudaMalloc(&ctx->sb, 14 * 6 * want * sizeof(double)));
  PC_CUDA_TRY(ctx, cudaMalloc(&ctx->sbf, 2 * want * sizeof(uint32_t)));
  ctx->sb_states = want;
  return PC_OK;
}
}  // namespace

#define PC_REQUIRE(ctx, cond, msg) \
  do {                             \
    if (!(cond)) return pc_fail((ctx), PC_ERR_BAD_ARG, msg); \
  } while (0)

extern "C" {

int pc_version(void) { return 100; }

int pc_ctx_create(int device_ordinal, pc_ctx** out) {
  if (!out) return PC_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return PC_ERR_NO_DEVICE;
  if (device_ordinal < 0 || device_ordinal >= count) return PC_ERR_BAD_ARG;
  pc_ctx* ctx = new (std::nothrow) pc_ctx();
  if (!ctx) return PC_ERR_BAD_ARG;
  ctx->device = device_ordinal;
  DeviceGuard g(device_ordinal);
  if (cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return PC_ERR_CUDA;
  }
  if (cudaMalloc(&ctx->peer_dev, 4 * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(ctx->peer_dev, 0, 4 * sizeof(unsigned int)) != cudaSuccess) {
    delete ctx;
    return PC_ERR_CUDA;
  }
  *out = ctx;
  return PC_OK;
}

int pc_ctx_destroy(pc_ctx* ctx) {
  if (!ctx) return PC_OK;
  {
    DeviceGuard g(ctx->device);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->sb) cudaFree(ctx->sb);
    if (ctx->sbf) cudaFree(ctx->sbf);
    if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
    if (ctx->sq) cudaFree(ctx->sq);
    if (ctx->peer_dev) cudaFree(ctx->peer_dev);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    for (auto& e : ctx->hs)
      if (e.exec) cudaGraphExecDestroy(e.exec);
    pc_comm_destroy(ctx);
  }
  delete ctx;
  return PC_OK;
}

const char* pc_last_error(pc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t pc_launch_count(pc_ctx* ctx) { return ctx ? ctx->launches : 0; }

static int* option_slot(pc_ctx* ctx, int option, int* max_value) {
  switch (option) {
    case PC_OPT_FINISH_KERNEL: *max_value = 4; return &ctx->opt_finish;
    case PC_OPT_PROPAGATE_KERNEL: *max_value = 2; return &ctx->opt_prop;
    default: return nullptr;
  }
}

int pc_set_option(pc_ctx* ctx, int option, int value) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  int mx = 0;
  int* slot = option_slot(ctx, option, &mx);
  PC_REQUIRE(ctx, slot != nullptr, "unknown option");
  PC_REQUIRE(ctx, value >= 0 && value <= mx, "option value out of range");
  *slot = value;
  return PC_OK;
}

int pc_get_option(pc_ctx* ctx, int option, int* value) {
  PC_REQUIRE(ctx, ctx != nullptr && value != nullptr, "null pointer");
  int mx = 0;
  int* slot = option_slot(ctx, option, &mx);
  PC_REQUIRE(ctx, slot != nullptr, "unknown option");
  *value = *slot;
  return PC_OK;
}

/* ---- K1 ---------------------------------------------------------------- */
static int propagate_common(pc_ctx* ctx, const double* x_in, double* x_out, const uint8_t* kind,
                            int default_kind, int64_t n, int64_t ld, double t0, double tf,
                            int substeps, int force_model, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, n >= 0 && ld >= n, "need 0 <= n <= ld");
  PC_REQUIRE(ctx, n == 0 || (x_in && x_out), "null state pointer");
  PC_REQUIRE(ctx, substeps >= 0, "substeps must be >= 0 (0 = auto)");
  PC_REQUIRE(ctx, force_model == PC_FORCE_TWOBODY || force_model == PC_FORCE_TWOBODY_J2,
             "unknown force model");
  if (t0 == tf) return pc_fail(ctx, PC_ERR_EQUAL_TIMES, "t0 and tf are equal");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_propagate_agents(x_in, x_out, kind, default_kind, n, ld, t0, tf,
                                               substeps, force_model, (cudaStream_t)stream));
  if (n) ctx->launches += 1;
  return PC_OK;
}

int pc_propagate_twobody(pc_ctx* ctx, const double* x_in, double* x_out, int64_t n, int64_t ld,
                         double t0, double tf, int substeps, int force_model, pc_stream stream) {
  return propagate_common(ctx, x_in, x_out, nullptr, PC_KIND_SATELLITE, n, ld, t0, tf, substeps,
                          force_model, stream);
}

int pc_rotate_terrestrial(pc_ctx* ctx, const double* x_in, double* x_out, int64_t n, int64_t ld,
                          double t0, double tf, pc_stream stream) {
  // The reference's StaticTerrestrial does not reject t0 == tf (identity rotation).
  if (ctx && t0 == tf && n > 0 && x_in && x_out) {
    DeviceGuard g(ctx->device);
    if (x_in != x_out)
      PC_CUDA_TRY(ctx, cudaMemcpy2DAsync(x_out, ld * sizeof(double), x_in, ld * sizeof(double),
                                         n * sizeof(double), 6, cudaMemcpyDeviceToDevice,
                                         (cudaStream_t)stream));
    return PC_OK;
  }
  return propagate_common(ctx, x_in, x_out, nullptr, PC_KIND_TERRESTRIAL, n, ld, t0, tf, 0,
                          PC_FORCE_TWOBODY, stream);
}

int pc_propagate_agents(pc_ctx* ctx, const double* x_in, double* x_out, const uint8_t* kind,
                        int64_t n, int64_t ld, double t0, double tf, int substeps, int force_model,
                        pc_stream stream) {
  return propagate_common(ctx, x_in, x_out, kind, PC_KIND_SATELLITE, n, ld, t0, tf, substeps,
                          force_model, stream);
}

int pc_propagate_agents_host(pc_ctx* ctx, const double* h_in, double* h_out, const uint8_t* h_kind,
                             int64_t n, int64_t ld, double t0, double tf, int substeps,
                             int force_model) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, n >= 0 && ld >= n, "need 0 <= n <= ld");
  PC_REQUIRE(ctx, n == 0 || (h_in && h_out), "null state pointer");
  if (t0 == tf) return pc_fail(ctx, PC_ERR_EQUAL_TIMES, "t0 and tf are equal");
  if (n == 0) return PC_OK;
  DeviceGuard g(ctx->device);
  const size_t xb = align256(6 * (size_t)ld * sizeof(double));
  const size_t kb = align256((size_t)n);
  int rc = ensure_ws(ctx, xb + kb);
  if (rc) return rc;
  double* d_x = (double*)ctx->ws;
  uint8_t* d_k = h_kind ? (uint8_t*)ctx->ws + xb : nullptr;
  cudaStream_t s = ctx->host_stream;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_x, h_in, 6 * (size_t)ld * sizeof(double), cudaMemcpyHostToDevice, s));
  if (h_kind) PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_k, h_kind, (size_t)n, cudaMemcpyHostToDevice, s));
  rc = pc_propagate_agents(ctx, d_x, d_x, d_k, n, ld, t0, tf, substeps, force_model, s);
  if (rc) return rc;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_out, d_x, 6 * (size_t)ld * sizeof(double), cudaMemcpyDeviceToHost, s));
  PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return PC_OK;
}

/* ---- frames / initial conditions -------------------------------------- */
int pc_frame_change(pc_ctx* ctx, const double* x_in, double* x_out, int64_t n, int64_t ld,
                    double jd, int to_eci, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, n >= 0 && ld >= n, "need 0 <= n <= ld");
  PC_REQUIRE(ctx, n == 0 || (x_in && x_out), "null state pointer");
  DeviceGuard g(ctx->device);
  const double w = to_eci ? pc::kOmegaEarth : -pc::kOmegaEarth;
  PC_CUDA_TRY(ctx, pc::launch_frame_change(x_in, x_out, n, ld, w * jd, w, (cudaStream_t)stream));
  if (n) ctx->launches += 1;
  return PC_OK;
}

int pc_coe2eci(pc_ctx* ctx, const double* coe, double* x_out, int64_t n, int64_t ld, double mu,
               pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, n >= 0 && ld >= n, "need 0 <= n <= ld");
  PC_REQUIRE(ctx, n == 0 || (coe && x_out), "null pointer");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_coe2eci(coe, x_out, n, ld, mu, (cudaStream_t)stream));
  if (n) ctx->launches += 1;
  return PC_OK;
}

int pc_lla2eci(pc_ctx* ctx, const double* lla, double* x_out, int64_t n, int64_t ld, double jd,
               pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, n >= 0 && ld >= n, "need 0 <= n <= ld");
  PC_REQUIRE(ctx, n == 0 || (lla && x_out), "null pointer");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_lla2eci(lla, x_out, n, ld, jd, (cudaStream_t)stream));
  if (n) ctx->launches += 1;
  return PC_OK;
}

/* ---- K3 ---------------------------------------------------------------- */
int pc_vismap_packed(pc_ctx* ctx, const double* sens, int64_t M, int64_t ld_s,
                     const uint8_t* sens_kind, double sens_rot_angle, const double* targ_a,
                     const double* targ_b, int64_t N, int64_t ld_t, double body_radius, double tol,
                     uint32_t* out_a, uint32_t* out_b, int64_t words_per_row, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0 && ld_s >= M && ld_t >= N, "bad sizes");
  PC_REQUIRE(ctx, words_per_row * 32 >= M, "words_per_row too small for M sensors");
  if (M == 0 || N == 0) return PC_OK;
  PC_REQUIRE(ctx, sens && targ_a && out_a, "null pointer");
  PC_REQUIRE(ctx, (targ_b == nullptr) == (out_b == nullptr), "targ_b and out_b go together");
  PC_REQUIRE(ctx, words_per_row == (M + 31) / 32, "words_per_row must be ceil(M/32)");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_vismap_packed(sens, M, ld_s, sens_kind, sens_rot_angle, targ_a, targ_b,
                                            N, ld_t, body_radius, tol, out_a, out_b, words_per_row,
                                            (cudaStream_t)stream));
  ctx->launches += 1;
  return PC_OK;
}

int pc_vismap_value(pc_ctx* ctx, const double* sens, int64_t M, int64_t ld_s, const double* targ,
                    int64_t N, int64_t ld_t, double body_radius, double tol, double* out,
                    pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0 && ld_s >= M && ld_t >= N, "bad sizes");
  if (M == 0 || N == 0) return PC_OK;
  PC_REQUIRE(ctx, sens && targ && out, "null pointer");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_vismap_value(sens, M, ld_s, targ, N, ld_t, body_radius, tol, out,
                                           (cudaStream_t)stream));
  ctx->launches += 1;
  return PC_OK;
}

int pc_vismap_host(pc_ctx* ctx, const double* h_sens, int64_t M, const double* h_targ, int64_t N,
                   double body_radius, double tol, int binary, int64_t* out_i64, double* out_f64) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0, "bad sizes");
  if (M == 0 || N == 0) return PC_OK;
  PC_REQUIRE(ctx, h_sens && h_targ, "null pointer");
  PC_REQUIRE(ctx, binary ? out_i64 != nullptr : out_f64 != nullptr, "null output");
  DeviceGuard g(ctx->device);
  const int64_t wpr = (M + 31) / 32;
  const size_t sb = align256(6 * (size_t)M * 8), tb = align256(6 * (size_t)N * 8);
  const size_t wb = align256((size_t)N * wpr * 4), ob = align256((size_t)N * M * 8);
  int rc = ensure_ws(ctx, sb + tb + wb + ob);
  if (rc) return rc;
  char* base = (char*)ctx->ws;
  double* d_s = (double*)base;
  double* d_t = (double*)(base + sb);
  uint32_t* d_w = (uint32_t*)(base + sb + tb);
  void* d_o = base + sb + tb + wb;
  cudaStream_t s = ctx->host_stream;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, h_sens, 6 * (size_t)M * 8, cudaMemcpyHostToDevice, s));
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_t, h_targ, 6 * (size_t)N * 8, cudaMemcpyHostToDevice, s));
  if (binary) {
    rc = pc_vismap_packed(ctx, d_s, M, M, nullptr, 0.0, d_t, nullptr, N, N, body_radius, tol, d_w,
                          nullptr, wpr, s);
    if (rc) return rc;
    PC_CUDA_TRY(ctx, pc::launch_unpack_bits(d_w, N, M, wpr, (int64_t*)d_o, s));
    ctx->launches += 1;
    PC_CUDA_TRY(ctx, cudaMemcpyAsync(out_i64, d_o, (size_t)N * M * 8, cudaMemcpyDeviceToHost, s));
  } else {
    rc = pc_vismap_value(ctx, d_s, M, M, d_t, N, N, body_radius, tol, (double*)d_o, s);
    if (rc) return rc;
    PC_CUDA_TRY(ctx, cudaMemcpyAsync(out_f64, d_o, (size_t)N * M * 8, cudaMemcpyDeviceToHost, s));
  }
  PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return PC_OK;
}

int pc_vismap_packed_envs(pc_ctx* ctx, const double* sens, int64_t ld_s, const double* targ_a,
                          const double* targ_b, int64_t ld_t, int64_t E, int64_t env_sensors,
                          int64_t env_targets, double body_radius, double tol, uint32_t* out_a,
                          uint32_t* out_b, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, E >= 0 && env_sensors >= 0 && env_targets >= 0, "bad sizes");
  PC_REQUIRE(ctx, ld_s >= E * env_sensors && ld_t >= E * env_targets, "leading dimensions too small");
  if (E == 0 || env_sensors == 0 || env_targets == 0) return PC_OK;
  PC_REQUIRE(ctx, sens && targ_a && out_a, "null pointer");
  PC_REQUIRE(ctx, (targ_b == nullptr) == (out_b == nullptr), "targ_b and out_b go together");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_vismap_envs(sens, ld_s, targ_a, targ_b, E * env_targets, ld_t, env_targets,
                                          env_sensors, body_radius, tol, out_a, out_b, (cudaStream_t)stream));
  ctx->launches += 1;
  return PC_OK;
}

int pc_obs_pack_envs(pc_ctx* ctx, const double* sens_x, int64_t ld_s, const double* est_x,
                     const double* est_p, int64_t ld, const float* last_time_tasked, double t_now,
                     const pc_clock* clock, int64_t E, int64_t env_sensors, int64_t env_targets,
                     float* out_eci, float* out_cov, float* out_stale, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, E >= 0 && env_sensors >= 0 && env_targets >= 0, "bad sizes");
  PC_REQUIRE(ctx, ld_s >= E * env_sensors && ld >= E * env_targets, "leading dimensions too small");
  PC_REQUIRE(ctx, out_eci && out_cov && out_stale, "null output");
  if (E == 0) return PC_OK;
  PC_REQUIRE(ctx, (env_sensors == 0 || sens_x) && (env_targets == 0 || (est_x && est_p && last_time_tasked)),
             "null input");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_obs_pack_envs(sens_x, ld_s, est_x, est_p, E * env_targets, ld, last_time_tasked,
                                            t_now, clock, E, env_sensors, env_targets, out_eci, out_cov,
                                            out_stale, (cudaStream_t)stream));
  ctx->launches += 1;
  return PC_OK;
}

int pc_vismap_der(pc_ctx* ctx, const double* sens, int64_t M, int64_t ld_s, const double* targ,
                  int64_t N, int64_t ld_t, double body_radius, double tol, double* out_v,
                  double* out_dv, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0 && ld_s >= M && ld_t >= N, "bad sizes");
  if (M == 0 || N == 0) return PC_OK;
  PC_REQUIRE(ctx, sens && targ && out_v && out_dv, "null pointer");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_vismap_der(sens, M, ld_s, targ, N, ld_t, body_radius, tol, out_v, out_dv,
                                         (cudaStream_t)stream));
  ctx->launches += 1;
  return PC_OK;
}

int pc_vismap_der_host(pc_ctx* ctx, const double* h_sens, int64_t M, const double* h_targ, int64_t N,
                       double body_radius, double tol, double* out_v, double* out_dv) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0, "bad sizes");
  if (M == 0 || N == 0) return PC_OK;
  PC_REQUIRE(ctx, h_sens && h_targ && out_v && out_dv, "null pointer");
  DeviceGuard g(ctx->device);
  const size_t sb = align256(6 * (size_t)M * 8), tb = align256(6 * (size_t)N * 8);
  const size_t ob = align256((size_t)N * M * 8);
  int rc = ensure_ws(ctx, sb + tb + 2 * ob);
  if (rc) return rc;
  char* base = (char*)ctx->ws;
  double* d_s = (double*)base;
  double* d_t = (double*)(base + sb);
  double* d_v = (double*)(base + sb + tb);
  double* d_d = (double*)(base + sb + tb + ob);
  cudaStream_t s = ctx->host_stream;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, h_sens, 6 * (size_t)M * 8, cudaMemcpyHostToDevice, s));
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_t, h_targ, 6 * (size_t)N * 8, cudaMemcpyHostToDevice, s));
  rc = pc_vismap_der(ctx, d_s, M, M, d_t, N, N, body_radius, tol, d_v, d_d, s);
  if (rc) return rc;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(out_v, d_v, (size_t)N * M * 8, cudaMemcpyDeviceToHost, s));
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(out_dv, d_d, (size_t)N * M * 8, cudaMemcpyDeviceToHost, s));
  PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return PC_OK;
}

/* ---- visibility history over a time axis (access windows) --------------------------------- */
int pc_vis_history(pc_ctx* ctx, double* sens, const uint8_t* sens_kind, int64_t M, int64_t ld_s,
                   double* targ, const uint8_t* targ_kind, int64_t N, int64_t ld_t, double t_now,
                   const double* h_times, int64_t T, int substeps, int force_model,
                   double body_radius, double tol, uint32_t* out_words, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0 && ld_s >= M && ld_t >= N && T >= 0, "bad sizes");
  if (T == 0 || M == 0 || N == 0) return PC_OK;
  PC_REQUIRE(ctx, sens && targ && h_times && out_words, "null pointer");
  DeviceGuard g(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t wpr = (M + 31) / 32;
  double t_cur = t_now;
  for (int64_t i = 0; i < T; ++i) {
    if (h_times[i] != t_cur) {
      PC_CUDA_TRY(ctx, pc::launch_propagate_agents(sens, sens, sens_kind, PC_KIND_SATELLITE, M, ld_s, t_cur,
                                                   h_times[i], substeps, force_model, s));
      PC_CUDA_TRY(ctx, pc::launch_propagate_agents(targ, targ, targ_kind, PC_KIND_SATELLITE, N, ld_t, t_cur,
                                                   h_times[i], substeps, force_model, s));
      ctx->launches += 2;
      t_cur = h_times[i];
    }
    PC_CUDA_TRY(ctx, pc::launch_vismap_packed(sens, M, ld_s, nullptr, 0.0, targ, nullptr, N, ld_t, body_radius,
                                              tol, out_words + i * N * wpr, nullptr, wpr, s));
    ctx->launches += 1;
  }
  return PC_OK;
}

int pc_vis_history_host(pc_ctx* ctx, const double* h_sens, const uint8_t* h_sens_kind, int64_t M,
                        const double* h_targ, const uint8_t* h_targ_kind, int64_t N, double t_now,
                        const double* h_times, int64_t T, int substeps, int force_model,
                        double body_radius, double tol, uint32_t* h_out_words) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0 && T >= 0, "bad sizes");
  if (T == 0 || M == 0 || N == 0) return PC_OK;
  PC_REQUIRE(ctx, h_sens && h_targ && h_times && h_out_words, "null pointer");
  DeviceGuard g(ctx->device);
  const int64_t wpr = (M + 31) / 32;
  const size_t sb = align256(6 * (size_t)M * 8), tb = align256(6 * (size_t)N * 8);
  const size_t skb = align256((size_t)M), tkb = align256((size_t)N);
  const size_t ob = align256((size_t)T * N * wpr * 4);
  int rc = ensure_ws(ctx, sb + tb + skb + tkb + ob);
  if (rc) return rc;
  char* b = (char*)ctx->ws;
  double* d_s = (double*)b;            b += sb;
  double* d_t = (double*)b;            b += tb;
  uint8_t* d_sk = (uint8_t*)b;         b += skb;
  uint8_t* d_tk = (uint8_t*)b;         b += tkb;
  uint32_t* d_o = (uint32_t*)b;
  cudaStream_t s = ctx->host_stream;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, h_sens, 6 * (size_t)M * 8, cudaMemcpyHostToDevice, s));
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_t, h_targ, 6 * (size_t)N * 8, cudaMemcpyHostToDevice, s));
  if (h_sens_kind) PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_sk, h_sens_kind, (size_t)M, cudaMemcpyHostToDevice, s));
  if (h_targ_kind) PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_tk, h_targ_kind, (size_t)N, cudaMemcpyHostToDevice, s));
  rc = pc_vis_history(ctx, d_s, h_sens_kind ? d_sk : nullptr, M, M, d_t, h_targ_kind ? d_tk : nullptr, N, N,
                      t_now, h_times, T, substeps, force_model, body_radius, tol, d_o, s);
  if (rc) return rc;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_out_words, d_o, (size_t)T * N * wpr * 4, cudaMemcpyDeviceToHost, s));
  PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return PC_OK;
}

/* ---- K4 ---------------------------------------------------------------- */
static int ukf_check(pc_ctx* ctx, const double* est_x, const double* est_p, const double* truth_x,
                     const double* z, const uint8_t* tasked, int64_t n, int64_t ld, double t0,
                     double tf, int substeps, int force_model, const pc_ukf_params* params) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, n >= 0 && ld >= n, "need 0 <= n <= ld");
  PC_REQUIRE(ctx, params != nullptr, "null params");
  PC_REQUIRE(ctx, n == 0 || (est_x && est_p), "null filter state pointer");
  PC_REQUIRE(ctx, substeps >= 0, "substeps must be >= 0 (0 = auto)");
  PC_REQUIRE(ctx, force_model == PC_FORCE_TWOBODY || force_model == PC_FORCE_TWOBODY_J2,
             "unknown force model");
  PC_REQUIRE(ctx, !(tasked && !z && !truth_x),
             "tasked targets need z or truth_x (to draw the measurement)");
  if (t0 == tf) return pc_fail(ctx, PC_ERR_EQUAL_TIMES, "t0 and tf are equal");
  return PC_OK;
}

int pc_ukf_predict_update(pc_ctx* ctx, double* est_x, double* est_p, double* truth_x,
                          const double* z, const uint8_t* tasked, int64_t n, int64_t ld, double t0,
                          double tf, int substeps, int force_model, const pc_ukf_params* params,
                          uint64_t rng_seed, uint64_t rng_step, double* out_pred_x,
                          double* out_pred_p, double* out_nis, uint32_t* out_flags, double* out_z,
                          pc_stream stream) {
  int rc = ukf_check(ctx, est_x, est_p, truth_x, z, tasked, n, ld, t0, tf, substeps, force_model, params);
  if (rc) return rc;
  if (n == 0) return PC_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  rc = ensure_sb(ctx, ld);
  if (rc) return rc;
  // stage 1: Cholesky + sigma points into the context's sigma buffer (ld == row stride there too)
  PC_CUDA_TRY(ctx, pc::launch_sigma_gen(est_x, est_p, truth_x, ctx->sb, n, ld, params->gamma, ctx->sbf, st));
  // stage 2: all 13 (+1) n states
  PC_CUDA_TRY(ctx, pc::launch_propagate_sigma(ctx->sb, n, ld, truth_x != nullptr, t0, tf, substeps,
                                              force_model, ctx->sbf, false, ctx->opt_prop, st));
  // stage 3: unscented transform + update
  pc::UkfArgs u{};
  u.sb = ctx->sb;
  u.est_p = est_p;
  u.z = z;
  u.tasked = const_cast<uint8_t*>(tasked);
  u.n = n;
  u.ld = ld;
  u.dt = tf - t0;
  u.t_now = tf;
  u.rng_seed = rng_seed;
  u.rng_step = rng_step;
  u.out_pred_x = out_pred_x;
  u.out_pred_p = out_pred_p;
  u.out_nis = out_nis;
  u.out_flags = out_flags;
  u.out_z = out_z;
  u.sig_flags = ctx->sbf;
  u.est_x_out = est_x;
  u.truth_out = truth_x;
  u.variant = ctx->opt_finish;
  u.p = *params;
  PC_CUDA_TRY(ctx, pc::launch_ukf_finish(u, st));
  ctx->launches += 2;
  ctx->launches += 1;
  return PC_OK;
}

int pc_sigma_points(pc_ctx* ctx, const double* est_x, const double* est_p, int64_t n, int64_t ld,
                    double gamma, double* sigma_ws, uint32_t* sig_flags, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, n >= 0 && ld >= n, "need 0 <= n <= ld");
  if (n == 0) return PC_OK;
  PC_REQUIRE(ctx, est_x && est_p && sigma_ws && sig_flags, "null pointer");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_sigma_gen(est_x, est_p, nullptr, sigma_ws, n, ld, gamma, sig_flags,
                                        (cudaStream_t)stream));
  ctx->launches += 1;
  return PC_OK;
}

int pc_ukf_predict_update_host(pc_ctx* ctx, double* h_est_x, double* h_est_p, const double* h_z,
                               const uint8_t* h_tasked, int64_t n, double t0, double tf,
                               int substeps, int force_model, const pc_ukf_params* params,
                               double* h_pred_x, double* h_pred_p, double* h_nis,
                               uint32_t* h_flags) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, !(h_tasked && !h_z), "tasked targets need z on the host path");
  int rc = ukf_check(ctx, h_est_x, h_est_p, nullptr, h_z, h_tasked, n, n, t0, tf, substeps, force_model, params);
  if (rc) return rc;
  if (n == 0) return PC_OK;
  DeviceGuard g(ctx->device);
  const size_t xb = align256(6 * (size_t)n * 8), pb = align256(36 * (size_t)n * 8);
  const size_t nb = align256((size_t)n * 8), fb = align256((size_t)n * 4), kb = align256((size_t)n);
  rc = ensure_ws(ctx, 3 * xb + 2 * pb + nb + fb + kb);
  if (rc) return rc;
  char* b = (char*)ctx->ws;
  double* d_x = (double*)b;           b += xb;
  double* d_z = (double*)b;           b += xb;
  double* d_px = (double*)b;          b += xb;
  double* d_p = (double*)b;           b += pb;
  double* d_pp = (double*)b;          b += pb;
  double* d_nis = (double*)b;         b += nb;
  uint32_t* d_f = (uint32_t*)b;       b += fb;
  uint8_t* d_k = (uint8_t*)b;
  cudaStream_t s = ctx->host_stream;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_x, h_est_x, 6 * (size_t)n * 8, cudaMemcpyHostToDevice, s));
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_p, h_est_p, 36 * (size_t)n * 8, cudaMemcpyHostToDevice, s));
  if (h_z) PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_z, h_z, 6 * (size_t)n * 8, cudaMemcpyHostToDevice, s));
  if (h_tasked) PC_CUDA_TRY(ctx, cudaMemcpyAsync(d_k, h_tasked, (size_t)n, cudaMemcpyHostToDevice, s));
  rc = pc_ukf_predict_update(ctx, d_x, d_p, nullptr, h_z ? d_z : nullptr, h_tasked ? d_k : nullptr, n,
                             n, t0, tf, substeps, force_model, params, 0, 0, d_px, d_pp, d_nis, d_f,
                             nullptr, s);
  if (rc) return rc;
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_est_x, d_x, 6 * (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_est_p, d_p, 36 * (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  if (h_pred_x) PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_pred_x, d_px, 6 * (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  if (h_pred_p) PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_pred_p, d_pp, 36 * (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  if (h_nis) PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_nis, d_nis, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  if (h_flags) PC_CUDA_TRY(ctx, cudaMemcpyAsync(h_flags, d_f, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return PC_OK;
}

/* ---- scheduler glue --------------------------------------------------- */
int pc_task_from_actions(pc_ctx* ctx, const int64_t* actions, int64_t M, const uint32_t* vis_est,
                         int64_t N, int64_t words_per_row, uint8_t* tasked, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0, "bad sizes");
  if (N == 0) return PC_OK;
  PC_REQUIRE(ctx, tasked && (M == 0 || (actions && vis_est)), "null pointer");
  PC_REQUIRE(ctx, words_per_row == (M + 31) / 32, "words_per_row must be ceil(M/32)");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_task_from_actions(actions, M, vis_est, N, words_per_row, tasked,
                                                (cudaStream_t)stream));
  if (M) ctx->launches += 1;
  return PC_OK;
}

int pc_obs_pack(pc_ctx* ctx, const double* sens_x, int64_t M, int64_t ld_s, const double* est_x,
                const double* est_p, int64_t N, int64_t ld, const float* last_time_tasked,
                double t_now, const pc_clock* clock, float* out_eci, float* out_cov,
                float* out_stale, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, M >= 0 && N >= 0 && ld_s >= M && ld >= N, "bad sizes");
  PC_REQUIRE(ctx, out_eci && out_stale, "null output");   // out_cov may be NULL (est_cov written elsewhere)
  PC_REQUIRE(ctx, (M == 0 || sens_x) && (N == 0 || (est_x && est_p && last_time_tasked)), "null input");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_obs_pack(sens_x, M, ld_s, est_x, est_p, N, ld, last_time_tasked, t_now,
                                       clock, out_eci, out_cov, out_stale, (cudaStream_t)stream));
  ctx->launches += 1;
  return PC_OK;
}

int pc_policy_random_visible(pc_ctx* ctx, const uint32_t* vis_est, int64_t N, int64_t M,
                             int64_t words_per_row, uint64_t seed, uint64_t step,
                             const pc_clock* clock, int probes, int64_t* actions, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, N > 0 && M >= 0 && probes >= 0, "bad sizes");
  PC_REQUIRE(ctx, M == 0 || (vis_est && actions), "null pointer");
  PC_REQUIRE(ctx, words_per_row == (M + 31) / 32, "words_per_row must be ceil(M/32)");
  DeviceGuard g(ctx->device);
  PC_CUDA_TRY(ctx, pc::launch_policy_random_visible(vis_est, N, M, words_per_row, seed, step, clock,
                                                    probes, actions, (cudaStream_t)stream));
  if (M) ctx->launches += 1;
  return PC_OK;
}

// Work items [0, total) over the cores this process may use (at most 8 threads, at least min_per_thread items each);
// no exception crosses the C ABI: chunks whose thread could not be started are taken by the caller.
extern "C++" {
template <typename F>
static void host_parallel_for(int64_t total, int64_t min_per_thread, F range) {
  int64_t threads = 1;
  if (total >= 2 * min_per_thread) {
    cpu_set_t set;
    CPU_ZERO(&set);
    const int allowed = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : 1;
    threads = std::min<int64_t>(std::min<int64_t>(allowed, 8), total / min_per_thread);
  }
  if (threads <= 1) {
    range(0, total);
    return;
  }
  std::vector<std::thread> pool;
  const int64_t per = (total + threads - 1) / threads;
  int64_t started = 1;   // chunk 0 is this thread's
  try {
    for (; started < threads; ++started) pool.emplace_back(range, started * per, std::min(total, (started + 1) * per));
  } catch (...) {
  }
  range(0, std::min(total, per));
  if (started < threads) range(started * per, total);
  for (auto& th : pool) th.join();
}
}  // extern "C++"

static inline uint64_t splitmix64_host(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

int pc_policy_random_visible_host(const uint32_t* vis_est, int64_t N, int64_t M, int64_t words_per_row,
                                  uint64_t seed, uint64_t step, int probes, int64_t* actions) {
  if (N <= 0 || M < 0 || probes < 0 || words_per_row != (M + 31) / 32) return PC_ERR_BAD_ARG;
  if (M > 0 && (!vis_est || !actions)) return PC_ERR_BAD_ARG;
  for (int64_t j = 0; j < M; ++j) {   // policy_random_visible_kernel, one sensor at a time
    const uint64_t base = splitmix64_host(seed ^ splitmix64_host(step * 0x100000001B3ull + (uint64_t)j));
    int64_t pick = N;
    for (int k = 0; k < probes; ++k) {
      const int64_t i = (int64_t)(splitmix64_host(base + (uint64_t)k) % (uint64_t)N);
      if ((vis_est[i * words_per_row + (j >> 5)] >> (j & 31)) & 1u) {
        pick = i;
        break;
      }
    }
    actions[j] = pick;
  }
  return PC_OK;
}

int pc_policy_random_visible_envs_host(const uint32_t* vis_est, int64_t E, int64_t env_targets, int64_t env_sensors,
                                       int64_t words_per_row, uint64_t seed, uint64_t step, int probes,
                                       int64_t* actions) {
  if (E <= 0 || env_targets <= 0 || env_sensors < 0 || probes < 0 || words_per_row != (env_sensors + 31) / 32)
    return PC_ERR_BAD_ARG;
  if (env_sensors > 0 && (!vis_est || !actions)) return PC_ERR_BAD_ARG;
  const int64_t total = E * env_sensors;
  auto range = [=](int64_t j0, int64_t j1) {   // the tasking warps of pre_step_kernel, one sensor at a time
    for (int64_t j = j0; j < j1; ++j) {
      const int64_t e = j / env_sensors, jl = j - e * env_sensors, t0 = e * env_targets;
      const uint64_t base = splitmix64_host(seed ^ splitmix64_host(step * 0x100000001B3ull + (uint64_t)j));
      int64_t pick = env_targets;
      for (int k = 0; k < probes; ++k) {
        const int64_t i = (int64_t)(splitmix64_host(base + (uint64_t)k) % (uint64_t)env_targets);
        if ((vis_est[(t0 + i) * words_per_row + (jl >> 5)] >> (jl & 31)) & 1u) {
          pick = i;
          break;
        }
      }
      actions[j] = pick;
    }
  };
  // sensors are independent: a batch of environments (10 240 sensors at BASELINE.json configs[2]: 0.9 ms on one
  // core, more than the device step and the observation copy together) is split over the cores this process may use
  host_parallel_for(total, 2048, range);
  return PC_OK;
}

int pc_unpack_vis_host(const uint32_t* words, int64_t n, int64_t words_per_row, int64_t m, int64_t* out) {
  if (n < 0 || m < 0 || words_per_row != (m + 31) / 32) return PC_ERR_BAD_ARG;
  if (n == 0 || m == 0) return PC_OK;
  if (!words || !out) return PC_ERR_BAD_ARG;
  host_parallel_for(n, std::max<int64_t>(1, 262144 / m), [=](int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; ++i) {
      const uint32_t* w = words + i * words_per_row;
      int64_t* o = out + i * m;
      for (int64_t j = 0; j < m; ++j) o[j] = (w[j >> 5] >> (j & 31)) & 1u;
    }
  });
  return PC_OK;
}

/* ---- measurement helpers ------------------------------------------------ */
int pc_set_timing(pc_ctx* ctx, int enabled) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  DeviceGuard g(ctx->device);
  if (enabled && !ctx->ev0) {
    PC_CUDA_TRY(ctx, cudaEventCreate(&ctx->ev0));
    PC_CUDA_TRY(ctx, cudaEventCreate(&ctx->ev1));
    PC_CUDA_TRY(ctx, cudaEventCreate(&ctx->ev2));
  }
  ctx->timing = enabled != 0;
  return PC_OK;
}

int pc_get_timing(pc_ctx* ctx, double* ukf_ms_last) {
  PC_REQUIRE(ctx, ctx != nullptr && ukf_ms_last, "null pointer");
  PC_REQUIRE(ctx, ctx->ev0 != nullptr, "timing was never enabled");
  DeviceGuard g(ctx->device);
  float ms = 0.f;
  PC_CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
  PC_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ukf_ms_last = ms;
  return PC_OK;
}

int pc_get_kernel_timing(pc_ctx* ctx, double* propagate_ms_last, double* finish_ms_last) {
  PC_REQUIRE(ctx, ctx != nullptr && propagate_ms_last && finish_ms_last, "null pointer");
  PC_REQUIRE(ctx, ctx->ev0 != nullptr, "timing was never enabled");
  DeviceGuard g(ctx->device);
  float a = 0.f, b = 0.f;
  PC_CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev2));
  PC_CUDA_TRY(ctx, cudaEventElapsedTime(&a, ctx->ev0, ctx->ev1));
  PC_CUDA_TRY(ctx, cudaEventElapsedTime(&b, ctx->ev1, ctx->ev2));
  *propagate_ms_last = a;
  *finish_ms_last = b;
  return PC_OK;
}

int pc_fp64_peak(pc_ctx* ctx, int repeats, double* tflops) {
  PC_REQUIRE(ctx, ctx != nullptr && tflops, "null pointer");
  DeviceGuard g(ctx->device);
  cudaDeviceProp prop;
  PC_CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  int rc = ensure_ws(ctx, 4096);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  PC_CUDA_TRY(ctx, cudaEventCreate(&e0));
  PC_CUDA_TRY(ctx, cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8;
  cudaError_t err;
  pc::launch_dfma_probe((double*)ctx->ws, blocks, 256, ctx->host_stream, &err);  // warm-up
  PC_CUDA_TRY(ctx, err);
  double best = 0.0;
  for (int r = 0; r < (repeats > 0 ? repeats : 5); ++r) {
    PC_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->host_stream));
    const double flop = pc::launch_dfma_probe((double*)ctx->ws, blocks, 4096, ctx->host_stream, &err);
    PC_CUDA_TRY(ctx, err);
    PC_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->host_stream));
    PC_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    PC_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    const double tf = flop / (ms * 1e-3) * 1e-12;
    if (tf > best) best = tf;
    ctx->launches += 1;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return PC_OK;
}

/* ---- fused step -------------------------------------------------------- */
int pc_step_fused(pc_ctx* ctx, const pc_step_args* a, const pc_ukf_params* params,
                  pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, a != nullptr && params != nullptr, "null args");
  PC_REQUIRE(ctx, a->M >= 0 && a->ld_s >= a->M, "bad sensor sizes");
  PC_REQUIRE(ctx, a->truth_x != nullptr || a->N == 0, "step needs truth_x");
  int rc = ukf_check(ctx, a->est_x, a->est_p, a->truth_x, a->z, a->tasked, a->N, a->ld, a->t0, a->tf,
                     a->substeps, a->force_model, params);
  if (rc) return rc;
  PC_REQUIRE(ctx, a->M == 0 || (a->sens_x && a->sens_kind), "null sensor pointer");
  const bool want_vis = a->vis_truth || a->vis_est;
  PC_REQUIRE(ctx, !want_vis || (a->vis_truth && a->vis_est), "vis_truth and vis_est go together");
  const int64_t Ne = a->env_targets > 0 ? a->env_targets : a->N;
  const int64_t Me = a->env_sensors > 0 ? a->env_sensors : a->M;
  PC_REQUIRE(ctx, (a->env_targets > 0) == (a->env_sensors > 0), "env_targets and env_sensors go together");
  PC_REQUIRE(ctx, a->env_targets <= 0 || (a->N % Ne == 0 && a->M % Me == 0 && a->N / Ne == a->M / Me),
             "a batch of environments needs N = E env_targets and M = E env_sensors");
  PC_REQUIRE(ctx, !want_vis || a->words_per_row == (Me + 31) / 32,
             "words_per_row must be ceil(M/32) (ceil(env_sensors/32) for a batch of environments)");
  PC_REQUIRE(ctx, !a->actions || (a->tasked && (a->vis_est || a->M == 0)),
             "actions need a tasked workspace and the estimated vis map");
  PC_REQUIRE(ctx, a->policy_probes >= 0 && a->policy_probes <= 4096, "bad policy_probes");
  if (a->peer_table) {
    PC_REQUIRE(ctx, a->peer_world >= 1 && a->peer_world <= 64 && a->peer_rank >= 0 && a->peer_rank < a->peer_world,
               "bad peer_world / peer_rank");
    PC_REQUIRE(ctx, a->summary_base && a->summary_bytes > 0 && a->summary_bytes % 4 == 0 &&
                        reinterpret_cast<uintptr_t>(a->summary_base) % 16 == 0,
               "peer push needs a 16-byte aligned summary buffer whose size is a multiple of 4");
    PC_REQUIRE(ctx, !(ctx->comm.attached && a->peer_table == ctx->comm.table) ||
                        (a->summary_bytes == ctx->comm.slot_bytes && a->peer_world == ctx->comm.world &&
                         a->peer_rank == ctx->comm.rank),
               "summary_bytes / peer_world / peer_rank differ from what pc_comm_init was given");
  }
  const bool own_sb = a->sigma_ws != nullptr;
  if (own_sb && a->N > 0) {
    PC_REQUIRE(ctx, a->sig_flags != nullptr, "sigma_ws needs sig_flags");
    PC_REQUIRE(ctx, a->est_x == a->sigma_ws && a->truth_x == a->sigma_ws + 13 * 6 * a->ld,
               "with sigma_ws, est_x must be its block 0 and truth_x its block 13");
  }
  DeviceGuard g(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (!own_sb && a->N > 0) {
    rc = ensure_sb(ctx, a->ld);
    if (rc) return rc;
  }
  double* sb = own_sb ? a->sigma_ws : ctx->sb;
  uint32_t* sbf = own_sb ? a->sig_flags : ctx->sbf;
  const bool fuse_vis = want_vis && a->M > 0 && a->N > 0;
  double* sq = a->sens_q;
  if (fuse_vis && !sq) {
    rc = ensure_sq(ctx, a->M);
    if (rc) return rc;
    sq = ctx->sq;
  }
  // 1. sensors (Agent.propagate per sensor; terrestrial ones are a rotation), their horizon terms,
  //    tasking mask from the actions and the PREVIOUS estimated vis map, step clock latch.
  //    Independent of the target propagation, which is launched programmatically behind it and
  //    therefore runs beside it (no fork / join, no second stream).
  //    Multi-GPU: the same kernel pushes the summaries the PREVIOUS step left in the summary buffer into
  //    every peer (see PeerPush), beside the propagation; the wait for the peers' pushes of the same number
  //    is the last thing the last CTA of the per-target kernel does, so that a rank only stalls when a
  //    peer is a full step behind.
  pc::PeerPush pp{};
  if (a->peer_table && a->N > 0) {
    pp.world = a->peer_world;
    pp.rank = a->peer_rank;
    pp.table = a->peer_table;
    pp.src = a->summary_base;
    pp.bytes = a->summary_bytes;
    pp.ticket = ctx->peer_dev;
    pp.epoch = reinterpret_cast<unsigned long long*>(ctx->peer_dev + 2);
    pp.err = ctx->peer_dev + 1;
  }
  const bool pre = a->M > 0 || a->clock || pp.world > 0;
  if (pre) {
    PC_CUDA_TRY(ctx, pc::launch_pre_step(a->sens_x, a->sens_kind, a->M, a->ld_s, a->t0, a->tf, a->substeps,
                                         a->force_model, fuse_vis ? sq : nullptr, a->body_radius, a->vis_tol,
                                         a->clock, a->actions, a->vis_est, const_cast<uint8_t*>(a->tasked),
                                         a->N, a->words_per_row, a->policy_probes, a->policy_seed,
                                         a->rng_step, Ne, Me, a->actions ? a->task_of : nullptr, pp, s));
    ctx->launches += 1;
  }
  // 2. targets: sigma points (only when not carried over from the previous step), all 14 N
  //    states propagated, then transform + update + summaries + next step's sigma points + both
  //    visibility rows of every target (env.py:455-462)
  if (a->N > 0) {
    bool beside = pre && !ctx->timing;  // the roofline pass wants the propagation alone between its events
    if (!own_sb || !a->sigma_valid) {
      PC_CUDA_TRY(ctx, pc::launch_sigma_gen(a->est_x, a->est_p, a->truth_x, sb, a->N, a->ld, params->gamma,
                                            sbf, s));
      ctx->launches += 1;
      beside = false;  // the propagation reads what this kernel writes: ordinary dependency
    }
    unsigned evflags = cudaEventRecordDefault;
    if (ctx->timing) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      PC_CUDA_TRY(ctx, cudaStreamIsCapturing(s, &cap));
      if (cap == cudaStreamCaptureStatusActive) evflags = cudaEventRecordExternal;
      PC_CUDA_TRY(ctx, cudaEventRecordWithFlags(ctx->ev0, s, evflags));
    }
    PC_CUDA_TRY(ctx, pc::launch_propagate_sigma(sb, a->N, a->ld, 1, a->t0, a->tf, a->substeps,
                                                a->force_model, sbf, beside, ctx->opt_prop, s));
    if (ctx->timing) PC_CUDA_TRY(ctx, cudaEventRecordWithFlags(ctx->ev1, s, evflags));
    pc::UkfArgs u{};
    u.sb = sb;
    u.est_p = a->est_p;
    u.z = a->z;
    u.tasked = const_cast<uint8_t*>(a->tasked);
    u.clear_tasked = a->actions != nullptr;
    u.n = a->N;
    u.ld = a->ld;
    u.dt = a->tf - a->t0;
    u.t_now = a->tf;
    u.rng_seed = a->rng_seed;
    u.rng_step = a->rng_step;
    u.clock = a->clock;
    u.advance_clock = a->clock != nullptr;
    u.out_pred_p = a->pred_p;
    u.out_nis = a->nis;
    u.out_flags = a->flags;
    u.num_tasked = a->num_tasked;
    u.last_time_tasked = a->last_time_tasked;
    u.tr_pos = a->tr_pos;
    u.tr_vel = a->tr_vel;
    u.sig_flags = sbf;
    u.gen_next = own_sb ? 1 : 0;
    u.est_x_out = own_sb ? nullptr : a->est_x;
    u.truth_out = own_sb ? nullptr : a->truth_x;
    if (a->actions && a->task_of && a->M > 0) {
      u.task_of = a->task_of;
      u.task_sensors = a->M;
      u.task_env_sensors = Me;
      if (a->mirror_host && a->mirror_dev && a->mirror_bytes > 0 && pc::finish_fills_host_mirrors(a->N, a->M, ctx->opt_finish)) {
        // host mirrors of the summary arrays that live inside the caller's device block
        const char* d0 = static_cast<const char*>(a->mirror_dev);
        char* h0 = static_cast<char*>(a->mirror_host);
        auto mirror = [&](const void* p) -> char* {
          const char* q = static_cast<const char*>(p);
          return (p && q >= d0 && q < d0 + a->mirror_bytes) ? h0 + (q - d0) : nullptr;
        };
        u.h_num_tasked = reinterpret_cast<int64_t*>(mirror(a->num_tasked));
        u.h_vis_est = reinterpret_cast<uint32_t*>(mirror(a->vis_est));
        u.h_tr_pos = reinterpret_cast<float*>(mirror(a->tr_pos));
        u.h_tr_vel = reinterpret_cast<float*>(mirror(a->tr_vel));
        u.h_ltt = reinterpret_cast<float*>(mirror(a->last_time_tasked));
      }
      if (pc::finish_fills_host_mirrors(a->N, a->M, ctx->opt_finish)) u.h_cov = a->obs_cov_host;
    }
    if (fuse_vis) {
      u.sens_q = reinterpret_cast<const double4*>(sq);
      u.M = a->M;
      u.wpr = a->words_per_row;
      u.env_targets = a->env_targets > 0 ? Ne : 0;   // 0: one environment (no index divisions)
      u.env_sensors = Me;
      u.vis_truth = a->vis_truth;
      u.vis_est = a->vis_est;
      u.body_radius = a->body_radius;
      u.vis_tol = a->vis_tol;
    }
    u.variant = ctx->opt_finish;
    u.peer = pp;     // its last CTA waits for the peers' pushes and counts the push
    u.p = *params;
    PC_CUDA_TRY(ctx, pc::launch_ukf_finish(u, s));
    if (ctx->timing) PC_CUDA_TRY(ctx, cudaEventRecordWithFlags(ctx->ev2, s, evflags));
    ctx->launches += 2;
    // (no exchange here: the next step's sensor kernel pushes these summaries)
  } else if (a->clock) {
    PC_CUDA_TRY(ctx, pc::launch_advance_clock(a->clock, a->tf - a->t0, s));
    ctx->launches += 1;
  }
  return PC_OK;
}

int pc_peer_status(pc_ctx* ctx, uint32_t* out_error_mask, uint64_t* out_pushes) {
  PC_REQUIRE(ctx, ctx != nullptr && out_error_mask && out_pushes, "null pointer");
  DeviceGuard g(ctx->device);
  unsigned int v[4] = {0, 0, 0, 0};
  PC_CUDA_TRY(ctx, cudaMemcpy(v, ctx->peer_dev, sizeof(v), cudaMemcpyDeviceToHost));
  PC_CUDA_TRY(ctx, cudaMemset(ctx->peer_dev + 1, 0, sizeof(unsigned int)));
  *out_error_mask = v[1];
  *out_pushes = (uint64_t)v[2] | ((uint64_t)v[3] << 32);
  return PC_OK;
}

int pc_peer_exchange(pc_ctx* ctx, const uint64_t* peer_table, int32_t peer_world, int32_t peer_rank,
                     const void* summary, int64_t summary_bytes, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr && peer_table != nullptr && summary != nullptr, "null pointer");
  PC_REQUIRE(ctx, peer_world >= 1 && peer_world <= 64 && peer_rank >= 0 && peer_rank < peer_world,
             "bad peer_world / peer_rank");
  PC_REQUIRE(ctx, summary_bytes > 0 && summary_bytes % 4 == 0 && reinterpret_cast<uintptr_t>(summary) % 16 == 0,
             "the summary buffer must be 16-byte aligned and a multiple of 4 bytes long");
  DeviceGuard g(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  pc::PeerPush pp{};
  pp.world = peer_world;
  pp.rank = peer_rank;
  pp.table = peer_table;
  pp.src = summary;
  pp.bytes = summary_bytes;
  pp.ticket = ctx->peer_dev;
  pp.epoch = reinterpret_cast<unsigned long long*>(ctx->peer_dev + 2);
  pp.err = ctx->peer_dev + 1;
  PC_CUDA_TRY(ctx, pc::launch_peer_push(pp, s));
  PC_CUDA_TRY(ctx, pc::launch_peer_wait(peer_table, peer_world, peer_rank, pp.epoch, pp.err, 1, s));
  ctx->launches += 2;
  return PC_OK;
}

/* ---- peer exchange set-up over CUDA IPC (no framework needed) ----------------------------------- */
static constexpr size_t kCommFlagBytes = 64 * sizeof(uint64_t);

int pc_comm_init(pc_ctx* ctx, int32_t world, int32_t rank, int64_t summary_bytes,
                 uint8_t handle_out[PC_COMM_HANDLE_BYTES]) {
  PC_REQUIRE(ctx, ctx != nullptr && handle_out != nullptr, "null pointer");
  PC_REQUIRE(ctx, world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad world / rank");
  PC_REQUIRE(ctx, summary_bytes > 0, "summary_bytes must be positive");
  static_assert(sizeof(cudaIpcMemHandle_t) == PC_COMM_HANDLE_BYTES, "handle size");
  DeviceGuard g(ctx->device);
  int rc = pc_comm_destroy(ctx);
  if (rc) return rc;
  auto& c = ctx->comm;
  const size_t bytes = kCommFlagBytes + 2 * (size_t)world * (size_t)summary_bytes;
  PC_CUDA_TRY(ctx, cudaMalloc(&c.base, bytes));
  PC_CUDA_TRY(ctx, cudaMemset(c.base, 0, bytes));
  // the push counter restarts with the (zeroed) flags
  PC_CUDA_TRY(ctx, cudaMemset(ctx->peer_dev, 0, 4 * sizeof(unsigned int)));
  PC_CUDA_TRY(ctx, cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  PC_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, c.base));
  std::memcpy(handle_out, &h, sizeof(h));
  c.world = world;
  c.rank = rank;
  c.slot_bytes = summary_bytes;
  return PC_OK;
}

int pc_comm_attach(pc_ctx* ctx, const uint8_t* all_handles) {
  PC_REQUIRE(ctx, ctx != nullptr && all_handles != nullptr, "null pointer");
  auto& c = ctx->comm;
  PC_REQUIRE(ctx, c.base != nullptr && !c.attached, "pc_comm_init first (and attach once)");
  DeviceGuard g(ctx->device);
  c.opened.assign((size_t)c.world, nullptr);
  for (int r = 0; r < c.world; ++r) {
    if (r == c.rank) {
      c.opened[r] = c.base;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all_handles + (size_t)r * PC_COMM_HANDLE_BYTES, sizeof(h));
    PC_CUDA_TRY(ctx, cudaIpcOpenMemHandle(&c.opened[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  std::vector<uint64_t> tab(2 * (size_t)c.world);
  for (int r = 0; r < c.world; ++r) {
    tab[r] = reinterpret_cast<uint64_t>(static_cast<char*>(c.opened[r]) + kCommFlagBytes);
    tab[c.world + r] = reinterpret_cast<uint64_t>(c.opened[r]);
  }
  PC_CUDA_TRY(ctx, cudaMalloc(&c.table, tab.size() * sizeof(uint64_t)));
  PC_CUDA_TRY(ctx, cudaMemcpy(c.table, tab.data(), tab.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  c.attached = true;
  return PC_OK;
}

int pc_comm_table(pc_ctx* ctx, const uint64_t** out_table, void** out_gathered, int32_t* out_world,
                  int32_t* out_rank, int64_t* out_summary_bytes) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  auto& c = ctx->comm;
  PC_REQUIRE(ctx, c.attached, "no attached peer exchange on this context");
  if (out_table) *out_table = c.table;
  if (out_gathered) *out_gathered = static_cast<char*>(c.base) + kCommFlagBytes;
  if (out_world) *out_world = c.world;
  if (out_rank) *out_rank = c.rank;
  if (out_summary_bytes) *out_summary_bytes = c.slot_bytes;
  return PC_OK;
}

int pc_allgather_step(pc_ctx* ctx, const void* summary, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr && summary != nullptr, "null pointer");
  auto& c = ctx->comm;
  PC_REQUIRE(ctx, c.attached, "no attached peer exchange on this context");
  return pc_peer_exchange(ctx, c.table, c.world, c.rank, summary, c.slot_bytes, stream);
}

int pc_comm_detach(pc_ctx* ctx) {
  if (!ctx) return PC_OK;
  auto& c = ctx->comm;
  if (!c.base) return PC_OK;
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < (int)c.opened.size(); ++r)
    if (r != c.rank && c.opened[r]) cudaIpcCloseMemHandle(c.opened[r]);
  c.opened.clear();
  if (c.table) cudaFree(c.table);
  c.table = nullptr;
  c.attached = false;
  return PC_OK;
}

int pc_comm_destroy(pc_ctx* ctx) {
  if (!ctx) return PC_OK;
  auto& c = ctx->comm;
  if (!c.base) return PC_OK;
  pc_comm_detach(ctx);
  DeviceGuard g(ctx->device);
  cudaFree(c.base);
  c = pc_ctx::Comm{};
  return PC_OK;
}

/* ---- the step with host buffers (reference-facing call: actions in, observation out) ---- */
static int step_and_pack(pc_ctx* ctx, const pc_step_args* a, const pc_ukf_params* params,
                         float* d_obs, cudaStream_t s) {
  int rc = pc_step_fused(ctx, a, params, s);
  if (rc) return rc;
  if (d_obs && a->obs_cov_host && a->env_targets <= 0) {
    // zero-copy form: est_cov was written by the per-target kernel; eci_state and obs_staleness go
    // straight to the host block as well (obs_cov_host sits 6 (M + N) floats into it)
    const int64_t n0 = 6 * (a->M + a->N), n1 = 36 * a->N;
    float* h_obs = a->obs_cov_host - n0;
    return pc_obs_pack(ctx, a->sens_x, a->M, a->ld_s, a->est_x, a->est_p, a->N, a->ld, a->last_time_tasked,
                       a->tf, a->clock, h_obs, nullptr, h_obs + n0 + n1, s);
  }
  if (d_obs) {
    const int64_t n0 = 6 * (a->M + a->N), n1 = 36 * a->N;
    if (a->env_targets > 0)
      rc = pc_obs_pack_envs(ctx, a->sens_x, a->ld_s, a->est_x, a->est_p, a->ld, a->last_time_tasked, a->tf,
                            a->clock, a->N / a->env_targets, a->env_sensors, a->env_targets, d_obs, d_obs + n0,
                            d_obs + n0 + n1, s);
    else
      rc = pc_obs_pack(ctx, a->sens_x, a->M, a->ld_s, a->est_x, a->est_p, a->N, a->ld,
                       a->last_time_tasked, a->tf, a->clock, d_obs, d_obs + n0, d_obs + n0 + n1, s);
  }
  return rc;
}

int pc_step_host(pc_ctx* ctx, const pc_step_args* a, const pc_ukf_params* params,
                 const pc_step_io* io, pc_stream stream) {
  PC_REQUIRE(ctx, ctx != nullptr, "null context");
  PC_REQUIRE(ctx, a != nullptr && params != nullptr && io != nullptr, "null args");
  PC_REQUIRE(ctx, !io->h_actions || a->actions, "h_actions needs a device actions buffer in the step args");
  PC_REQUIRE(ctx, !io->h_obs || (io->d_obs && a->last_time_tasked), "h_obs needs d_obs and last_time_tasked");
  PC_REQUIRE(ctx, !io->h_vis_est || a->vis_est, "h_vis_est needs vis_est in the step args");
  PC_REQUIRE(ctx, !io->h_num_tasked || a->num_tasked, "h_num_tasked needs num_tasked in the step args");
  DeviceGuard g(ctx->device);
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->host_stream;
  // Replayable only when nothing in the argument block changes from step to step: device clock,
  // caller-owned sigma buffer already holding the current sigma points, measurements drawn on device.
  const bool replayable = a->clock && a->sigma_ws && a->sigma_valid && !a->z && !ctx->timing;
  const bool have_block = io->d_block && io->h_block && io->block_bytes > 0;
  const bool copy_actions = io->h_actions && a->M > 0;
  // zero-copy form: the kernels store the outputs into the (device-accessible) host block themselves
  const int64_t obs_floats = 6 * (a->M + a->N) + 37 * a->N;
  bool zero_copy = io->zero_copy && replayable && have_block && io->d_obs && a->env_targets <= 0 &&
                   a->actions && a->task_of && pc::finish_fills_host_mirrors(a->N, a->M, ctx->opt_finish) &&
                   io->d_obs == io->d_block && io->block_bytes >= obs_floats * (int64_t)sizeof(float);
  if (zero_copy) {
    // the kernels will dereference h_block: it must be pinned / registered host memory (checked once per buffer)
    if (ctx->hs_zc_ptr != io->h_block) {
      cudaPointerAttributes at{};
      const cudaError_t e = cudaPointerGetAttributes(&at, io->h_block);
      if (e != cudaSuccess) cudaGetLastError();
      ctx->hs_zc_ok = e == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer == io->h_block;
      ctx->hs_zc_ptr = io->h_block;
    }
    zero_copy = ctx->hs_zc_ok;
  }
  // ... and the sensor kernel reads the M actions straight from the (device-accessible) host buffer: 8 bytes
  // per sensor over PCIe inside the tasking warps, off the critical path, instead of a copy node that the
  // whole step -- propagation included -- would wait for
  bool direct_actions = io->zero_copy && replayable && copy_actions && a->policy_probes == 0;
  if (direct_actions) {
    if (ctx->hs_act_ptr != io->h_actions) {
      cudaPointerAttributes at{};
      const cudaError_t e = cudaPointerGetAttributes(&at, io->h_actions);
      if (e != cudaSuccess) cudaGetLastError();
      ctx->hs_act_ok = e == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer == io->h_actions;
      ctx->hs_act_ptr = io->h_actions;
    }
    direct_actions = ctx->hs_act_ok;
  }
  int rc = PC_OK;
  pc_step_args key;
  if (!replayable) {
    if (copy_actions)
      PC_CUDA_TRY(ctx, cudaMemcpyAsync(a->actions, io->h_actions, (size_t)a->M * sizeof(int64_t),
                                       cudaMemcpyHostToDevice, s));
    rc = step_and_pack(ctx, a, params, io->d_obs, s);
    if (rc) return rc;
  } else {
    // the two host copies (actions in, output block out) are nodes of the same graph: one launch and
    // one synchronise per step
    // with a device clock the kernels take the step's start time and RNG counter from it; only
    // tf - t0 matters, so the key is normalised and a caller that advances t0 / rng_step still hits
    key = *a;
    key.tf = a->tf - a->t0;
    key.t0 = 0.0;
    key.rng_step = 0;
    if (zero_copy) {
      key.mirror_dev = io->d_block;
      key.mirror_host = io->h_block;
      key.mirror_bytes = io->block_bytes;
      key.obs_cov_host = static_cast<float*>(io->h_block) + 6 * (a->M + a->N);
    } else {
      key.mirror_dev = nullptr;
      key.mirror_host = nullptr;
      key.mirror_bytes = 0;
      key.obs_cov_host = nullptr;
    }
    if (direct_actions) key.actions = const_cast<int64_t*>(io->h_actions);
    a = &key;
    pc_ctx::HsEntry* ent = nullptr;
    for (auto& e : ctx->hs)
      if (e.exec && e.stream == s && std::memcmp(&e.io, io, sizeof(*io)) == 0 &&
          std::memcmp(&e.args, a, sizeof(*a)) == 0 && std::memcmp(&e.params, params, sizeof(*params)) == 0) {
        ent = &e;
        break;
      }
    if (!ent) {
      // a free slot, else the slot of the same engine (same sigma buffer: its old graph is stale), else
      // the least recently used one
      for (auto& e : ctx->hs)
        if (!e.exec) { ent = &e; break; }
      if (!ent)
        for (auto& e : ctx->hs)
          if (e.args.sigma_ws == a->sigma_ws) { ent = &e; break; }
      if (!ent) {
        ent = &ctx->hs[0];
        for (auto& e : ctx->hs)
          if (e.last_use < ent->last_use) ent = &e;
      }
      if (ent->exec) {
        cudaGraphExecDestroy(ent->exec);
        ent->exec = nullptr;
      }
      const int64_t before = ctx->launches;
      cudaGraph_t graph = nullptr;
      PC_CUDA_TRY(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      cudaError_t ec = cudaSuccess;
      if (copy_actions && !direct_actions)
        ec = cudaMemcpyAsync(a->actions, io->h_actions, (size_t)a->M * sizeof(int64_t), cudaMemcpyHostToDevice, s);
      rc = ec == cudaSuccess ? step_and_pack(ctx, a, params, io->d_obs, s) : PC_ERR_CUDA;
      if (rc == PC_OK && have_block && !zero_copy)
        ec = cudaMemcpyAsync(io->h_block, io->d_block, (size_t)io->block_bytes, cudaMemcpyDeviceToHost, s);
      cudaError_t e = cudaStreamEndCapture(s, &graph);
      if (rc == PC_OK && ec != cudaSuccess) e = ec;
      ent->launches = ctx->launches - before;
      ctx->launches = before;
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      PC_CUDA_TRY(ctx, e);
      e = cudaGraphInstantiate(&ent->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) ent->exec = nullptr;
      PC_CUDA_TRY(ctx, e);
      std::memcpy(&ent->args, a, sizeof(*a));
      std::memcpy(&ent->params, params, sizeof(*params));
      std::memcpy(&ent->io, io, sizeof(*io));
      ent->stream = s;
      // the kernels only store what changes: start from a full copy of the device block
      if (zero_copy)
        PC_CUDA_TRY(ctx, cudaMemcpyAsync(io->h_block, io->d_block, (size_t)io->block_bytes, cudaMemcpyDeviceToHost, s));
    }
    ent->last_use = ++ctx->hs_tick;
    PC_CUDA_TRY(ctx, cudaGraphLaunch(ent->exec, s));
    ctx->launches += ent->launches;
    if (have_block) {
      PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
      return PC_OK;
    }
  }
  if (have_block) {
    PC_CUDA_TRY(ctx, cudaMemcpyAsync(io->h_block, io->d_block, (size_t)io->block_bytes, cudaMemcpyDeviceToHost, s));
    PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return PC_OK;
  }
  if (io->h_obs)
    PC_CUDA_TRY(ctx, cudaMemcpyAsync(io->h_obs, io->d_obs,
                                     (size_t)(6 * (a->M + a->N) + 37 * a->N) * sizeof(float),
                                     cudaMemcpyDeviceToHost, s));
  if (io->h_vis_est)
    PC_CUDA_TRY(ctx, cudaMemcpyAsync(io->h_vis_est, a->vis_est,
                                     (size_t)a->N * a->words_per_row * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, s));
  if (io->h_num_tasked)
    PC_CUDA_TRY(ctx, cudaMemcpyAsync(io->h_num_tasked, a->num_tasked, (size_t)a->N * sizeof(int64_t),
                                     cudaMemcpyDeviceToHost, s));
  PC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return PC_OK;
}

}  // extern "C"
