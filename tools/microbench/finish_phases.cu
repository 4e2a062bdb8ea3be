// Per-phase cycle stamps of the quad finish kernel on a synthetic 10k-target catalog.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I clockpunch_b200/csrc tools/microbench/finish_phases.cu -o tools/microbench/finish_phases.bin
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
__device__ long long g_stamp[16][2048];
#define PC_PHASE_CLOCK(i)                                                                     \
  do {                                                                                        \
    if ((threadIdx.x & 31) == 0) {                                                             \
      const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;                            \
      if (wid < 2048) g_stamp[i][wid] = clock64();                                             \
    }                                                                                         \
  } while (0)
__device__ long long g_cstamp[12][256];
#define PC_COOP_CLOCK(i)                                                     \
  do {                                                                       \
    if ((threadIdx.x & 31) == 0 && blockIdx.x < 256) g_cstamp[i][blockIdx.x] = clock64(); \
  } while (0)
#include "kernels_ukf.cu"
int pc_fail(pc_ctx*, int c, const char*) { return c; }
int pc_fail_cuda(pc_ctx*, cudaError_t, const char*) { return -2; }
namespace pc {
cudaError_t launch_propagate_sigma(double*, int64_t, int64_t, int, double, double, int, int, uint32_t*, bool, cudaStream_t) { return cudaSuccess; }
}
using namespace pc;

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 10000, ld = (n + 15) / 16 * 16, M = 100, wpr = 4;
  const int tasked_every = argc > 2 ? atoi(argv[2]) : 0;
  const int do_flush = argc > 3 ? atoi(argv[3]) : 1;
  const int coop = argc > 4 ? atoi(argv[4]) : 0;   // 1: update warps (task_of lists the tasked targets, at most M)
  std::vector<double> ex(6 * ld), P(36 * n, 0.0), tx(6 * ld), sq(4 * 128);
  for (int64_t t = 0; t < n; ++t) {
    const double a = 6778.0 + 36000.0 * ((t * 2654435761u) % 1000) / 1000.0, vc = sqrt(398600.0 / a), th = 0.001 * t;
    double x[6] = {a * cos(th), a * sin(th), 10.0, -vc * sin(th), vc * cos(th) * 0.9, vc * 0.43};
    for (int c = 0; c < 6; ++c) { ex[c * ld + t] = x[c] * (1 + 1e-4); tx[c * ld + t] = x[c]; }
    for (int c = 0; c < 6; ++c) P[t * 36 + c * 7] = c < 3 ? 10.0 : 0.1;
  }
  for (int j = 0; j < M; ++j) { sq[4 * j] = 6378.2 * cos(0.3 * j); sq[4 * j + 1] = 6378.2 * sin(0.3 * j); sq[4 * j + 2] = 0; sq[4 * j + 3] = 50.0; }
  double *d_ex, *d_P, *d_tx, *d_sb, *d_pp, *d_sq; uint32_t *d_sf, *d_vt, *d_ve; uint8_t* d_task; float *d_f1, *d_f2, *d_f3; int64_t* d_nt;
  cudaMalloc(&d_ex, ex.size() * 8); cudaMalloc(&d_P, P.size() * 8); cudaMalloc(&d_tx, tx.size() * 8);
  cudaMalloc(&d_sb, 14 * 6 * ld * 8); cudaMalloc(&d_pp, P.size() * 8); cudaMalloc(&d_sq, sq.size() * 8);
  cudaMalloc(&d_sf, 2 * n * 4); cudaMalloc(&d_vt, n * wpr * 4); cudaMalloc(&d_ve, n * wpr * 4); cudaMalloc(&d_task, n);
  cudaMalloc(&d_f1, n * 4); cudaMalloc(&d_f2, n * 4); cudaMalloc(&d_f3, n * 4); cudaMalloc(&d_nt, n * 8);
  cudaMemcpy(d_ex, ex.data(), ex.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_P, P.data(), P.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_tx, tx.data(), tx.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_sq, sq.data(), sq.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(d_nt, 0, n * 8); cudaMemset(d_f3, 0, n * 4);
  std::vector<uint8_t> tk(n, 0);
  if (tasked_every) for (int64_t t = 0; t < n; t += tasked_every) tk[t] = 1;
  std::vector<int64_t> tof(M, -1);
  if (coop && tasked_every) {
    std::fill(tk.begin(), tk.end(), 0);
    for (int j = 0; j < M && (int64_t)j * tasked_every < n; ++j) { tof[j] = (int64_t)j * tasked_every; tk[tof[j]] = 1; }
  }
  int64_t* d_tof; cudaMalloc(&d_tof, M * 8); cudaMemcpy(d_tof, tof.data(), M * 8, cudaMemcpyHostToDevice);
  pc_ukf_params prm{};
  prm.gamma = 1.7320508075688772e-3; prm.wm0 = -1999999.0000000002; prm.wi = 166666.66666666669; prm.wc0 = prm.wm0 + 3 - 1e-6; prm.eps_w = 0.0;
  for (int c = 0; c < 6; ++c) { prm.Q[c * 7] = c < 3 ? 0.1 : 1e-3; prm.R[c * 7] = c < 3 ? 1 : 1e-2; prm.Lr[c * 7] = c < 3 ? 1 : 0.1; }
  launch_sigma_gen(d_ex, d_P, d_tx, d_sb, n, ld, prm.gamma, d_sf, 0);
  UkfArgs u{};
  u.sb = d_sb; u.est_p = d_P; u.tasked = d_task; u.n = n; u.ld = ld; u.dt = 100; u.t_now = 100; u.out_pred_p = d_pp;
  u.num_tasked = d_nt; u.last_time_tasked = d_f3; u.tr_pos = d_f1; u.tr_vel = d_f2; u.sig_flags = d_sf; u.gen_next = 1;
  u.sens_q = (const double4*)d_sq; u.M = M; u.wpr = wpr; u.env_targets = 0; u.env_sensors = M; u.vis_truth = d_vt; u.vis_est = d_ve; u.body_radius = 6378.1363; u.vis_tol = 1e-13; u.p = prm;
  if (coop) { u.task_of = d_tof; u.task_sensors = M; u.task_env_sensors = M; u.clear_tasked = 1; }
  char* flush; cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemcpy(d_task, tk.data(), n, cudaMemcpyHostToDevice);
    if (coop) { UkfArgs v = u; v.task_of = nullptr; launch_ukf_finish(v, 0); cudaMemcpy(d_task, tk.data(), n, cudaMemcpyHostToDevice); }
    if (do_flush) cudaMemset(flush, rep, 256 << 20);
    cudaEventRecord(e0);
    launch_ukf_finish(u, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d: %.2f us (%s)\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  static long long h[16][2048];
  cudaMemcpyFromSymbol(h, g_stamp, sizeof(h));
  const int nw = (int)std::min<int64_t>(2048, 4 * n / 32);
  const char* names[] = {"entry+tile issue", "X0/truth loads issued", "pair loads + F accumulate", "quad reduce", "m, pred_p", "tasked flag + update", "outputs (P stores)", "chol + sigma writes", "tile wait", "vis words"};
  double tot = 0;
  for (int p = 1; p < 10; ++p) {
    double s = 0, mx = 0;
    for (int w = 0; w < nw; ++w) { const double d = (double)(h[p][w] - h[p - 1][w]); s += d; mx = std::max(mx, d); }
    printf("  phase %d %-28s mean %8.0f cycles  max %8.0f\n", p, names[p], s / nw, mx);
    tot += s / nw;
  }
  if (coop) {
    static long long hc[12][256];
    cudaMemcpyFromSymbol(hc, g_cstamp, sizeof(hc));
    const char* cn[] = {"", "loads issued", "loads + rng done", "transform, pred_p", "chol pred_p", "S, chol S (end part 1)", "gain, xe, est_p", "stores, summaries", "chol est_p, sigma writes", "vis row (end)", "helper end"};
    for (int p = 1; p <= 10; ++p) {
      double s = 0; int c = 0;
      for (int j = 0; j < M; ++j) if (tof[j] >= 0) { s += (double)(hc[p][j] - hc[p == 10 ? 0 : p - 1][j]); ++c; }
      printf("  update warp %2d %-28s mean %8.0f cycles\n", p, cn[p], c ? s / c : 0.0);
    }
  }
  long long first = h[0][0], last = h[9][0];
  for (int w = 0; w < nw; ++w) { first = std::min(first, h[0][w]); last = std::max(last, h[9][w]); }
  printf("  mean per-warp total %.0f cycles; first entry -> last exit (SM clocks not synchronised across SMs) %lld\n", tot, last - first);
  return 0;
}
