// Development harness for the hand-written tcgen05 int8 GEMM (C int32 = A^T B, both operands K-contiguous int8) that the
// int8-sliced gemm_nonlop (csrc/ozaki.cu) needs instead of cuBLASLt.  Checks bit-exactness against a naive kernel on small and
// ragged shapes, then times the Si-512 shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/igemm_lab tools/igemm_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../abinit_b200/csrc/igemm_tc.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_ref(int M, int N, int K, const int8_t* A, long long lda, const int8_t* B, long long ldb, int32_t* C, long long ldc) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (m >= M) return;
  int s = 0;
  for (int k = 0; k < K; k++) s += (int)A[(long long)m * lda + k] * (int)B[(long long)n * ldb + k];
  C[(long long)n * ldc + m] = s;
}
__global__ void k_fill(int8_t* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u + seed; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
    p[i] = (int8_t)((int)(x % 129) - 64);
  }
}

int main(int argc, char** argv) {
  const bool big = argc > 1 && argv[1][0] == 'b';
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  struct Shape { int M, N, K; } small[] = {{128, 256, 128}, {128, 256, 512}, {256, 512, 1024}, {100, 36, 256}, {388, 900, 1280}, {9216, 128, 2048}, {256, 256, 16384}, {300, 100, 32768}};
  Shape bigs[] = {{9216, 896, 288128}, {9216, 512, 288128}, {9216, 256, 288128}, {9216, 128, 288128}, {288116, 896, 9216}, {288116, 512, 9216}, {288116, 128, 9216}};
  for (const Shape& sh : (big ? std::vector<Shape>(bigs, bigs + 7) : std::vector<Shape>(small, small + 8))) {
    const long long lda = sh.K, ldb = sh.K, ldc = sh.M;
    int8_t *A, *B; int32_t *C, *R;
    CK(cudaMalloc(&A, (size_t)sh.M * lda)); CK(cudaMalloc(&B, (size_t)sh.N * ldb));
    CK(cudaMalloc(&C, sizeof(int32_t) * (size_t)sh.N * ldc));
    k_fill<<<1024, 256>>>(A, (size_t)sh.M * lda, 1u); k_fill<<<1024, 256>>>(B, (size_t)sh.N * ldb, 77u);
    CK(cudaMemset(C, 0xff, sizeof(int32_t) * (size_t)sh.N * ldc));
    CK(cudaDeviceSynchronize());
    abi::igemm_tc(sh.M, sh.N, sh.K, A, lda, B, ldb, C, ldc, 0, variant);
    CK(cudaDeviceSynchronize());
    if (!big) {
      CK(cudaMalloc(&R, sizeof(int32_t) * (size_t)sh.N * ldc));
      k_ref<<<dim3((sh.M + 127) / 128, sh.N), 128>>>(sh.M, sh.N, sh.K, A, lda, B, ldb, R, ldc);
      CK(cudaDeviceSynchronize());
      std::vector<int32_t> hc((size_t)sh.N * ldc), hr((size_t)sh.N * ldc);
      CK(cudaMemcpy(hc.data(), C, sizeof(int32_t) * hc.size(), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hr.data(), R, sizeof(int32_t) * hr.size(), cudaMemcpyDeviceToHost));
      size_t bad = 0; long long first = -1;
      for (size_t i = 0; i < hc.size(); i++) if (hc[i] != hr[i]) { if (first < 0) first = (long long)i; bad++; }
      printf("M=%d N=%d K=%d: %s (%zu mismatches, first at %lld: got %d want %d)\n", sh.M, sh.N, sh.K, bad ? "FAIL" : "ok", bad, first,
             first >= 0 ? hc[first] : 0, first >= 0 ? hr[first] : 0);
      cudaFree(R);
    } else {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      for (int r = 0; r < 5; r++) abi::igemm_tc(sh.M, sh.N, sh.K, A, lda, B, ldb, C, ldc, 0, variant);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      printf("M=%d N=%d K=%d: %.3f ms  %.1f TOP/s\n", sh.M, sh.N, sh.K, ms, 2.0 * sh.M * sh.N * sh.K / ms / 1e9);
    }
    fflush(stdout);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  return 0;
}
