// Microbenchmark: write-only and mixed read/write HBM bandwidth (the assembly kernels are write-dominated:
// ~400 MB of nzval out for ~90 MB of inputs in).  Feeds DESIGN.md's roofline discussion.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(double2* __restrict__ out, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_double2(v, v);
}
__global__ void k_write64(double* __restrict__ out, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
// 1 read : 4.5 writes (like the assembly step)
__global__ void k_mixed(const double2* __restrict__ in, double2* __restrict__ out, size_t n_in, int ratio) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_in; i += (size_t)gridDim.x * blockDim.x) {
    double2 v = in[i];
    for (int r = 0; r < ratio; ++r) out[i + (size_t)r * n_in] = v;
  }
}
template <class F> float timeit(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}
int main() {
  const size_t bytes = (size_t)435 << 20;       // ~ nzval of config 2
  double2* out; cudaMalloc(&out, bytes * 5);
  double2* in; cudaMalloc(&in, bytes);
  cudaMemset(in, 1, bytes);
  const size_t n = bytes / 16;
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    float ms = timeit([&] { k_write<<<g, 512>>>(out, n, 1.0); }, 20);
    printf("write-only STG.128 435 MB grid=%d: %.4f ms  %.0f GB/s\n", g, ms, bytes / ms / 1e6);
  }
  { float ms = timeit([&] { k_write64<<<148 * 16, 512>>>((double*)out, n * 2, 1.0); }, 20);
    printf("write-only STG.64 435 MB: %.4f ms  %.0f GB/s\n", ms, bytes / ms / 1e6); }
  { float ms = timeit([&] { cudaMemsetAsync(out, 0, bytes); }, 20);
    printf("cudaMemset 435 MB: %.4f ms  %.0f GB/s\n", ms, bytes / ms / 1e6); }
  { float ms = timeit([&] { k_write<<<148 * 16, 512>>>(out, n * 4, 1.0); }, 10);
    printf("write-only STG.128 1740 MB: %.4f ms  %.0f GB/s\n", ms, 4 * bytes / ms / 1e6); }
  { const size_t n_in = ((size_t)90 << 20) / 16; float ms = timeit([&] { k_mixed<<<148 * 16, 512>>>(in, out, n_in, 4); }, 20);
    printf("mixed 90 MB read + 360 MB write: %.4f ms  %.0f GB/s total\n", ms, 5.0 * n_in * 16 / ms / 1e6); }
  { float ms = timeit([&] { cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice); }, 20);
    printf("D2D copy 435 MB: %.4f ms  %.0f GB/s (read+write)\n", ms, 2.0 * bytes / ms / 1e6); }
  return 0;
}
