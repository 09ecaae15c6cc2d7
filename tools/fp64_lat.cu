// Microbenchmark: FP64 pipe latency / throughput on this GPU as a function of resident warps per SM and
// independent chains per thread (how much parallelism phase A of the sweep kernel needs).  Output feeds DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) x[i] = fma(x[i], a, b);
      else if (OP == 1) x[i] = x[i] + b;
      else x[i] = x[i] * a;
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP, int OP>
void run(int warps_per_sm, int iters) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 148);
  k_chain<ILP, OP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_chain<ILP, OP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double per_smsp_instr = (double)iters * ILP * warps_per_sm / 4.0;
  printf("op=%d warps/SM=%2d ILP=%d: %8.0f cyc, %.2f cyc per warp-instr per SMSP, chain step %.1f cyc, %.2f TFLOP/s-equiv (%.3f ms)\n", OP, warps_per_sm, ILP, c,
         c / per_smsp_instr, c / iters, (OP == 0 ? 2.0 : 1.0) * 148.0 * warps_per_sm * 32 * iters * ILP / ms / 1e9, ms);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  const int it = 20000;
  run<1, 0>(4, it); run<2, 0>(4, it); run<4, 0>(4, it); run<8, 0>(4, it); run<16, 0>(4, it);
  run<1, 0>(8, it); run<2, 0>(8, it); run<4, 0>(8, it); run<8, 0>(8, it); run<16, 0>(8, it);
  run<1, 0>(12, it); run<2, 0>(12, it); run<4, 0>(12, it); run<8, 0>(12, it);
  run<1, 0>(16, it); run<2, 0>(16, it); run<4, 0>(16, it); run<8, 0>(16, it);
  run<1, 0>(32, it); run<2, 0>(32, it); run<4, 0>(32, it);
  run<1, 1>(4, it); run<4, 1>(4, it); run<1, 2>(4, it); run<4, 2>(4, it);
  // long run for sustained clocks
  run<8, 0>(16, 2000000);
  return 0;
}
