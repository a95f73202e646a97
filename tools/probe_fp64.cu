// Probe: FP64 pipe throughput on sm_100a -- DFMA vs DMMA (mma.sync f64 shapes).
// Not part of the product; used once to choose the GEMM inner instruction.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void k_dfma(double* out, int iters) {
  double a[16]; double x = threadIdx.x * 1e-9 + 1.0, y = 0.999999;
  #pragma unroll
  for (int i = 0; i < 16; i++) a[i] = i * 0.5;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0; for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_mma884(double* out, int iters) {
  double c[NACC][2]; double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  #pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++) mma884(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, int KIND>
__global__ void k_mma16(double* out, int iters) {
  double c[NACC][4]; double a[8], b[4];
  #pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  #pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-6 * i;
  #pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (KIND == 4) mma1684(c[i], a, b[0]);
      if (KIND == 8) mma1688(c[i], a, b);
      if (KIND == 16) mma16816(c[i], a, b);
    }
  }
  double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\",\n", p.name, nsm, p.major, p.minor);
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  const int iters = 20000;
  for (int wpsm : {4, 8, 16, 32}) {
    int threads = 256, blocks = nsm * wpsm * 32 / threads;
    float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf(" \"dfma_w%d_tflops\": %.2f,\n", wpsm, fl / ms / 1e9);
    ms = timeit([&] { k_mma884<8><<<blocks, threads>>>(out, iters); });
    fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * threads / 32;
    printf(" \"mma884_w%d_tflops\": %.2f,\n", wpsm, fl / ms / 1e9);
    ms = timeit([&] { k_mma16<8, 4><<<blocks, threads>>>(out, iters); });
    fl = 2.0 * 16 * 8 * 4 * 8 * iters * (double)blocks * threads / 32;
    printf(" \"mma1684_w%d_tflops\": %.2f,\n", wpsm, fl / ms / 1e9);
    ms = timeit([&] { k_mma16<8, 8><<<blocks, threads>>>(out, iters / 2); });
    fl = 2.0 * 16 * 8 * 8 * 8 * (iters / 2) * (double)blocks * threads / 32;
    printf(" \"mma1688_w%d_tflops\": %.2f,\n", wpsm, fl / ms / 1e9);
    ms = timeit([&] { k_mma16<8, 16><<<blocks, threads>>>(out, iters / 4); });
    fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * (double)blocks * threads / 32;
    printf(" \"mma16816_w%d_tflops\": %.2f,\n", wpsm, fl / ms / 1e9);
  }
  printf(" \"done\": 1}\n");
  return 0;
}
