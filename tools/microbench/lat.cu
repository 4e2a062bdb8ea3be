// FP64 latency / throughput microbenchmarks for B200 (run under gpurun).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dep_dfma(double* out, long long* cyc, int iters, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = fma(x, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
template <int ILP>
__global__ void ilp_dfma(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dep_rsqrt(double* out, long long* cyc, int iters) {
  double x = 1.0e7 + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) x = rsqrt(x) + 1.0e7;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const int it = 1000;
  dep_dfma<<<1, 32>>>(out, cyc, it, 0.999, 1e-3); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.2f cycles\n", (double)h / (it * 16));
  dep_rsqrt<<<1, 32>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent rsqrt+DADD latency: %.2f cycles\n", (double)h / (it * 4));
#define RUN(ILP, W) ilp_dfma<ILP><<<1, 32 * W>>>(out, cyc, it, 0.999, 1e-3); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("ILP=%d warps/SM=%d: %.2f cycles per warp-DFMA per SMSP\n", ILP, W, (double)h / (it * 8.0 * ILP * ((W + 3) / 4)));
  RUN(1, 4) RUN(2, 4) RUN(3, 4) RUN(4, 4) RUN(8, 4) RUN(1, 8) RUN(1, 16) RUN(3, 16) RUN(1, 32) RUN(3, 32) RUN(8, 16)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
