// Microbenchmark: FP64 tensor-core (DMMA) issue rate on this GPU for the mma.sync f64 shapes, with
// register-resident operands (no memory traffic).  Denominator for the high-order (Q3) roofline in DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int ILP>
__global__ void k_dmma(double* out, int iters) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x % 7) + i * 1e-4;
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x % 5) - i * 1e-4;
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if constexpr (SHAPE == 0) { double (&cc)[2] = *reinterpret_cast<double (*)[2]>(&c[i][0]); dmma884(cc, a[0], b[0]); }
      if constexpr (SHAPE == 1) { double (&aa)[2] = *reinterpret_cast<double (*)[2]>(&a[0]); dmma1684(c[i], aa, b[0]); }
      if constexpr (SHAPE == 2) { double (&aa)[4] = *reinterpret_cast<double (*)[4]>(&a[0]); double (&bb)[2] = *reinterpret_cast<double (*)[2]>(&b[0]); dmma1688(c[i], aa, bb); }
      if constexpr (SHAPE == 3) dmma16816(c[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

template <int SHAPE, int ILP>
void run(int blocks, int threads, int iters) {
  static const char* names[] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  static const double fma_per[] = {256, 512, 1024, 2048};
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dmma<SHAPE, ILP><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e0);
  k_dmma<SHAPE, ILP><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)blocks * (threads / 32) * iters * ILP * fma_per[SHAPE];
  printf("DMMA %-9s ILP=%d blocks=%d threads=%d: %.3f ms  %.2f TFLOP/s  (%.2f cycles/instr/SMSP at 1.965 GHz, %d warps/SMSP)\n",
         names[SHAPE], ILP, blocks, threads, ms, 2 * fma / ms / 1e9,
         ms * 1e-3 * 1.965e9 / ((double)blocks / 148 * (threads / 32) / 4 * iters * ILP), blocks / 148 * (threads / 32) / 4);
  cudaFree(out);
}

int main2();
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
  run<0, 8>(148, 256, 4096);
  run<0, 8>(148 * 2, 256, 4096);
  run<0, 1>(148, 128, 8192);
  run<0, 2>(148, 128, 8192);
  run<0, 4>(148, 128, 8192);
  run<1, 8>(148, 256, 4096);
  run<1, 8>(148 * 2, 256, 4096);
  run<1, 1>(148, 128, 8192);
  run<2, 8>(148, 256, 2048);
  run<2, 1>(148, 128, 4096);
  run<3, 8>(148, 256, 1024);
  run<3, 1>(148, 128, 2048);
  main2();
  return 0;
}

// ---- operand-pattern variants (what a real register-blocked kernel issues) --------------------------------
// PAT 0: consecutive DMMAs share B, distinct A and C (column-block ownership)   PAT 1: share A, distinct B and C
template <int PAT, int NT, int NK>
__global__ void k_dmma_pat(double* out, int iters) {
  double a[NK * NT], b[NK * NT];
  for (int i = 0; i < NK * NT; ++i) { a[i] = 1e-3 * (threadIdx.x % 7) + i * 1e-4; b[i] = 1e-3 * (threadIdx.x % 5) - i * 1e-4; }
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if constexpr (PAT == 0) dmma884(*reinterpret_cast<double (*)[2]>(&c[t][0]), a[k * NT + t], b[k]);
        else dmma884(*reinterpret_cast<double (*)[2]>(&c[t][0]), a[k], b[k * NT + t]);
      }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
template <int PAT, int NT, int NK>
void run_pat(int threads, int iters) {
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dmma_pat<PAT, NT, NK><<<148, threads>>>(out, iters);
  cudaEventRecord(e0);
  k_dmma_pat<PAT, NT, NK><<<148, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)(threads / 32) / 4 * iters * NT * NK;   // DMMAs per SM sub-partition
  printf("DMMA pattern %d (%s) NT=%d NK=%d threads=%d: %.3f ms  %.2f cycles/instr/SMSP  %.2f TFLOP/s\n", PAT,
         PAT == 0 ? "shared B, distinct A,C" : "shared A, distinct B,C", NT, NK, threads, ms, ms * 1e-3 * 1.965e9 / n,
         148.0 * 4 * n * 512 / ms / 1e9);
  cudaFree(out);
}
int main2() {
  run_pat<0, 5, 8>(256, 2048);
  run_pat<1, 5, 8>(256, 2048);
  run_pat<0, 4, 8>(256, 2048);
  run_pat<0, 8, 4>(256, 2048);
  run_pat<1, 8, 4>(256, 2048);
  run_pat<0, 5, 8>(128, 2048);
  return 0;
}
