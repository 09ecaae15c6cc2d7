// Microbenchmark: sustained FP64 FMA rate and plain HBM copy bandwidth of this GPU (denominators for DESIGN.md).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}
__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
template <int ILP>
void run(int blocks, int threads, int iters) {
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k_dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)blocks * threads * iters * ILP;
  printf("DFMA ILP=%d blocks=%d threads=%d: %.3f ms  %.2f TFMA/s = %.2f TFLOP/s\n", ILP, blocks, threads, ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d clock=%d kHz smemOptin=%zu regs/SM=%d\n", p.name, p.multiProcessorCount, p.clockRate, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor);
  run<8>(148 * 8, 256, 4096);
  run<4>(148 * 8, 256, 8192);
  run<8>(148 * 2, 160, 8192);
  run<1>(148 * 8, 256, 8192);
  size_t n = (size_t)1 << 28;  // 4 GiB each way of double2? no: 2^28 double2 = 4 GiB
  double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16);
  cudaMemset(a, 1, n * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_copy<<<148 * 16, 512>>>(a, b, n);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k_copy<<<148 * 16, 512>>>(a, b, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("copy: %.1f GB/s (read+write)\n", 5.0 * 2 * n * 16 / ms / 1e6);
  return 0;
}
