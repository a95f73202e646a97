// Probe: do the FP64 tensor instruction (mma.sync.m8n8k4.f64, SASS DMMA) and the plain FP64 FMA (DFMA) share one pipe on sm_100?
// Runs DMMA only, DFMA only, and an interleaved mix from the same warps, and two-kernel concurrency on separate streams.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_fp64_mix tools/probe_fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>   // 0: DMMA only, 1: DFMA only, 2: both interleaved
__global__ void k(int iters, double* out) {
  double c[8][2], f[8];
  for (int i = 0; i < 8; i++) { c[i][0] = c[i][1] = 0.0; f[i] = threadIdx.x * 1e-3 + i; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE != 1) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      if (MODE != 0) { asm volatile("fma.rn.f64 %0, %1, %2, %0;\n" : "+d"(f[i]) : "d"(a), "d"(b)); asm volatile("fma.rn.f64 %0, %1, %2, %0;\n" : "+d"(f[i]) : "d"(b), "d"(a)); }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> double run(int iters, double* out, cudaStream_t st = 0) {
  k<MODE><<<148 * 4, 256, 0, st>>>(16, out);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st); k<MODE><<<148 * 4, 256, 0, st>>>(iters, out); cudaEventRecord(e1, st); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 4 * 256 * 2);
  const int iters = 1 << 14;
  const double warps = 148.0 * 4 * 8;
  const double fl_mma = 2.0 * 8 * 8 * 4 * 8 * iters * warps;            // per DMMA: 8x8x4 MACs
  const double fl_fma = 2.0 * 32 * 2 * 8 * iters * warps;                // 2 DFMA per slot, 32 lanes
  double t0 = run<0>(iters, out), t1 = run<1>(iters, out), t2 = run<2>(iters, out);
  printf("DMMA only : %.2f ms  %.1f TFLOP/s\n", t0, fl_mma / t0 / 1e9);
  printf("DFMA only : %.2f ms  %.1f TFLOP/s\n", t1, fl_fma / t1 / 1e9);
  printf("interleave: %.2f ms  %.1f TFLOP/s total (DMMA %.1f + DFMA %.1f); sum of the separate times %.2f ms\n", t2, (fl_mma + fl_fma) / t2 / 1e9,
         fl_mma / t2 / 1e9, fl_fma / t2 / 1e9, t0 + t1);
  // two kernels on two streams
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  cudaDeviceSynchronize();
  cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
  k<0><<<148 * 2, 256, 0, s1>>>(iters, out); k<1><<<148 * 2, 256, 0, s2>>>(iters, out + 148 * 4 * 256);
  cudaEventRecord(e1, s1); cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("two streams (half the CTAs each): %.2f ms for %.1f TFLOP -> %.1f TFLOP/s\n", ms, (fl_mma + fl_fma) / 2 / 1e12, (fl_mma + fl_fma) / 2 / ms / 1e9);
  return 0;
}
