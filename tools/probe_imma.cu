// Probe: throughput of the legacy warp-level int8 MMA (mma.sync.m16n8k32.s8, SASS IMMA) on sm_100a -- is a cp.async + mma.sync
// int8 GEMM (the DMMA kernel's structure) a viable hand-written home for the int8-sliced contractions, or is tcgen05 needed?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_imma tools/probe_imma.cu && tools/probe_imma
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int iters, int* out) {
  int acc[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+r"(acc[i][0]), "+r"(acc[i][1]), "+r"(acc[i][2]), "+r"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int* out; cudaMalloc(&out, sizeof(int) * 148 * 8 * 256);
  const int iters = 1 << 15;
  for (int ctas = 1; ctas <= 8; ctas *= 2) {
    k<<<148 * ctas, 256>>>(16, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148 * ctas, 256>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 2.0 * 16 * 8 * 32 * 8.0 * iters * (148.0 * ctas * 8);
    printf("IMMA m16n8k32 s8: %d CTA/SM x 8 warps: %.1f TOP/s\n", ctas, ops / ms / 1e9);
  }
  return 0;
}
