// Can a kernel stream its float32 observation straight into mapped pinned host memory instead of
// a device buffer + cudaMemcpyAsync?  Times both for a few sizes and store widths.
#include <cstdio>
#include <cuda_runtime.h>
template <typename V>
__global__ void writer(V* __restrict__ dst, size_t n, float seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    V v;
    float* f = reinterpret_cast<float*>(&v);
    for (int k = 0; k < (int)(sizeof(V) / 4); ++k) f[k] = seed + i;
    dst[i] = v;
  }
}
int main() {
  for (size_t bytes : {(size_t)2 << 20, (size_t)24 << 20, (size_t)196 << 20}) {
    void *h, *d;
    cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    cudaMalloc(&d, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    auto time = [&](auto fn, const char* name) {
      for (int i = 0; i < 3; ++i) fn();
      cudaEventRecord(e0);
      for (int i = 0; i < 10; ++i) fn();
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      printf("  %-44s %8.1f us  %6.1f GB/s\n", name, ms * 100, bytes / (ms * 1e-4) * 1e-9);
    };
    printf("%zu MB\n", bytes >> 20);
    time([&] { writer<float4><<<1250, 64>>>((float4*)d, bytes / 16, 1.f); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost); },
         "kernel -> device, then cudaMemcpyAsync D2H");
    time([&] { writer<float4><<<1250, 64>>>((float4*)h, bytes / 16, 1.f); }, "kernel -> mapped host, float4 stores, 1250x64");
    time([&] { writer<float4><<<592, 256>>>((float4*)h, bytes / 16, 1.f); }, "kernel -> mapped host, float4 stores, 592x256");
    time([&] { writer<float><<<1250, 64>>>((float*)h, bytes / 4, 1.f); }, "kernel -> mapped host, float stores");
    time([&] { writer<float4><<<1250, 64>>>((float4*)d, bytes / 16, 1.f); }, "kernel -> device only");
    cudaFreeHost(h); cudaFree(d);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
