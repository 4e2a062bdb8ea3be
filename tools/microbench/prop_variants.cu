// Launch-shape / code-shape variants of the sigma-buffer propagation kernel, timed on a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I clockpunch_b200/csrc tools/microbench/prop_variants.cu -o tools/microbench/prop_variants.bin
#include <cstdio>
#include <vector>
#include <cmath>
#include "propagate.cuh"

int pc_fail(pc_ctx*, int c, const char*) { return c; }
int pc_fail_cuda(pc_ctx*, cudaError_t, const char*) { return -2; }

using namespace pc;

template <int T, int MINB, int FAST>
__global__ void __launch_bounds__(T, MINB)
prop_variant(double* __restrict__ sb, int64_t n, int64_t ld, double dt, int substeps,
             const uint32_t* __restrict__ aux) {
  const int64_t t = (int64_t)blockIdx.x * T + threadIdx.x;
  if (t >= n) return;
  const int i = blockIdx.y;
  double* base = sb + (int64_t)i * 6 * ld + t;
  double r[3], v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r[c] = base[(int64_t)c * ld];
    v[c] = base[(int64_t)(c + 3) * ld];
  }
  int nsub = substeps;
  if (nsub <= 0) nsub = substeps_from_rate(__uint_as_float(aux[t]), dt, nullptr);
  if (FAST && nsub == 1) {
    double rr[1][3] = {{r[0], r[1], r[2]}}, vv[1][3] = {{v[0], v[1], v[2]}};
    rkn_step<1, false>(rr, vv, dt, -kMu * dt * dt, 1.0 / dt);
#pragma unroll
    for (int c = 0; c < 3; ++c) { r[c] = rr[0][c]; v[c] = vv[0][c]; }
  } else {
    propagate_state<false>(r, v, dt, nsub);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    base[(int64_t)c * ld] = r[c];
    base[(int64_t)(c + 3) * ld] = v[c];
  }
}

template <int T, int MINB, int FAST>
float run(double* d, const std::vector<double>& h, int64_t n, const uint32_t* aux, int reps, const char* name) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dim3 grid((unsigned)((n + T - 1) / T), 14);
  float best = 1e30f, sum = 0;
  for (int r = 0; r < reps + 2; ++r) {
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    prop_variant<T, MINB, FAST><<<grid, T>>>(d, n, n, 100.0, 0, aux);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2) { best = fminf(best, ms); sum += ms; }
  }
  printf("  %-22s n=%8lld: best %8.2f us  mean %8.2f us  -> %.2f TF (1150 flop/state)\n", name, (long long)n,
         best * 1e3, sum / reps * 1e3, 14.0 * n * 1150 / (best * 1e-3) * 1e-12);
  return best;
}

int main() {
  for (int64_t n : {10000LL, 125000LL, 1000000LL}) {
    std::vector<double> h(14 * 6 * n);
    std::vector<uint32_t> aux(n);
    for (int64_t t = 0; t < n; ++t) {
      const double a = 6778.0 + 36000.0 * ((t * 2654435761u) % 1000) / 1000.0;
      const double vc = sqrt(398600.0 / a), th = 0.001 * t;
      for (int b = 0; b < 14; ++b) {
        double* x = h.data() + (int64_t)b * 6 * n;
        x[0 * n + t] = a * cos(th) + 0.01 * b; x[1 * n + t] = a * sin(th); x[2 * n + t] = 0.1 * b;
        x[3 * n + t] = -vc * sin(th); x[4 * n + t] = vc * cos(th) * 0.9; x[5 * n + t] = vc * 0.43;
      }
      const float om = (float)sqrt(398600.0 / (a * a * a));
      aux[t] = *reinterpret_cast<const uint32_t*>(&om);
    }
    double* d; uint32_t* da;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&da, n * 4);
    cudaMemcpy(da, aux.data(), n * 4, cudaMemcpyHostToDevice);
    const int reps = n > 500000 ? 5 : 20;
    printf("n = %lld targets (x14 states)\n", (long long)n);
    run<128, 5, 0>(d, h, n, da, reps, "T128 minb5 (current)");
    run<128, 4, 0>(d, h, n, da, reps, "T128 minb4");
    run<128, 3, 0>(d, h, n, da, reps, "T128 minb3");
    run<96, 5, 0>(d, h, n, da, reps, "T96 minb5");
    run<64, 8, 0>(d, h, n, da, reps, "T64 minb8");
    run<64, 10, 0>(d, h, n, da, reps, "T64 minb10");
    run<128, 5, 1>(d, h, n, da, reps, "T128 minb5 fast1");
    run<128, 4, 1>(d, h, n, da, reps, "T128 minb4 fast1");
    run<96, 5, 1>(d, h, n, da, reps, "T96 minb5 fast1");
    run<64, 8, 1>(d, h, n, da, reps, "T64 minb8 fast1");
    run<256, 2, 1>(d, h, n, da, reps, "T256 minb2 fast1");
    cudaFree(d); cudaFree(da);
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
