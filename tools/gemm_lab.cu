// Developer lab for the FP64 DMMA GEMMs of gemm_nonlop (not part of the product): times kernel variants on the two
// bench shapes and checks sampled outputs against a host recomputation.
//   TN: C(M,N) = A(K,M)^T B(K,N)   M = nprojs, N = ndat, K = 2 npw      (opernla, split-K)
//   NN: C(M,N) = A(M,K)   B(K,N)   M = 2 npw,  N = ndat, K = nprojs     (opernlb)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/gemm_lab tools/gemm_lab.cu
// run:   tools/gemm_lab [npw=144057] [nprojs=9216] [ndat=128]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

#ifndef FULLMANT
#define FULLMANT 0
#endif
__host__ __device__ inline double hashval(uint64_t i, uint64_t seed) {
  uint64_t x = i * 0x9E3779B97F4A7C15ULL + seed;
  x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 29; x *= 0x94D049BB133111EBULL; x ^= x >> 32;
  return FULLMANT ? (double)(int64_t)(x >> 11) / 4503599627370496.0 - 1.0 : (double)(int64_t)(x & 0xFFFFF) / 524288.0 - 1.0;      // [-1, 1)
}
__global__ void k_fill(double* p, size_t n, uint64_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = hashval(i, seed);
}

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int BK = 16;
// K-major tile [rows][16]: 16-byte chunk c of row r lives at chunk c ^ swz(r)  (conflict-free LDS.64 fragment reads)
__device__ __forceinline__ int swz(int r) { return ((r & 3) << 1) | ((r >> 2) & 1); }

struct GemmParams {
  int M, N, K;
  const double* A; long long lda;     // TN: A[k + m*lda] ; NN: A[m + k*lda]
  const double* B; long long ldb;     // B[k + n*ldb]
  double* C; long long ldc;           // NN: C[m + n*ldc] ; TN: partials [z][n][m]
  int nsplit, kchunk, tiles_m, tiles_n;
};

// 8 warps, warp tile WM x WN, CTA tile (WM*WARPS_M) x (WN*WARPS_N), fragment double buffering in registers
template <bool TN, int WM, int WN, int WARPS_M, int WARPS_N, int STAGES, int MINB = 1>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MINB) k_gemm(GemmParams p) {
  constexpr int BM = WM * WARPS_M, BN = WN * WARPS_N, NT = WARPS_M * WARPS_N * 32;
  constexpr int FM = WM / 8, FN = WN / 8;
  constexpr int PITCH = BM + 4;                       // M-major A tile pitch: 32 bytes mod 128 -> 4 k rows hit 4 bank quarters
  constexpr int A_STAGE = TN ? BM * BK : BK * PITCH;
  constexpr int B_STAGE = BN * BK;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp / WARPS_N) * WM, wn = (warp % WARPS_N) * WN;
  int bid = blockIdx.x;
  const int tn = bid % p.tiles_n; bid /= p.tiles_n;
  const int tm = bid % p.tiles_m; const int z = bid / p.tiles_m;
  const int m0 = tm * BM, n0 = tn * BN;
  const int k_begin = z * p.kchunk, k_end = min(p.K, k_begin + p.kchunk);
  const int nkt = (k_end - k_begin + BK - 1) / BK;

  auto load_stage = [&](int s, int kt) {
    const int k0 = k_begin + kt * BK;
    double* as = As + s * A_STAGE;
    double* bs = Bs + s * B_STAGE;
    if (TN) {
#pragma unroll
      for (int c = tid; c < BM * 8; c += NT) {
        const int row = c >> 3, ch = c & 7;
        const int gm = m0 + row, gk = k0 + ch * 2;
        const bool ok = gm < p.M && gk < k_end;
        cp_async16(as + row * BK + ((ch ^ swz(row)) << 1), ok ? p.A + (long long)gm * p.lda + gk : p.A, ok);
      }
    } else {
#pragma unroll
      for (int c = tid; c < BK * (BM / 2); c += NT) {
        const int krow = c / (BM / 2), ch = c % (BM / 2);
        const int gk = k0 + krow, gm = m0 + ch * 2;
        const bool ok = gk < k_end && gm < p.M;
        cp_async16(as + krow * PITCH + ch * 2, ok ? p.A + (long long)gk * p.lda + gm : p.A, ok);
      }
    }
#pragma unroll
    for (int c = tid; c < BN * 8; c += NT) {
      const int row = c >> 3, ch = c & 7;
      const int gn = n0 + row, gk = k0 + ch * 2;
      const bool ok = gn < p.N && gk < k_end;
      cp_async16(bs + row * BK + ((ch ^ swz(row)) << 1), ok ? p.B + (long long)gn * p.ldb + gk : p.B, ok);
    }
  };

  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; i++)
#pragma unroll
    for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // per-lane fragment offsets inside a stage (doubles), for kk = 0; kk advances by a compile-time stride
  int a_off[FM], b_off[FN];
#pragma unroll
  for (int i = 0; i < FM; i++) {
    const int row = wm + 8 * i + g;
    a_off[i] = TN ? row * BK : row;          // TN: + swizzled k ; NN: + k*PITCH
  }
#pragma unroll
  for (int j = 0; j < FN; j++) b_off[j] = (wn + 8 * j + g) * BK;
  // swizzle terms: rows = base + g with base multiple of 8 -> swz(row) = swz(g)
  const int sw = swz(g);
  auto ld_frags = [&](const double* as, const double* bs, int kk, double (&a)[FM], double (&b)[FN]) {
    const int k = kk * 4 + t;
    const int ksw = ((((k >> 1) ^ sw) << 1) | (k & 1));
#pragma unroll
    for (int i = 0; i < FM; i++) a[i] = TN ? as[a_off[i] + ksw] : as[a_off[i] + k * PITCH];
#pragma unroll
    for (int j = 0; j < FN; j++) b[j] = bs[b_off[j] + ksw];
  };

  for (int s = 0; s < STAGES - 1; s++) { if (s < nkt) load_stage(s, s); cp_async_commit(); }
  cp_async_wait<STAGES - 2>();
  __syncthreads();
  double a[2][FM], b[2][FN];
  ld_frags(As, Bs, 0, a[0], b[0]);
  for (int kt = 0; kt < nkt; kt++) {
    const double* as = As + (kt % STAGES) * A_STAGE;
    const double* bs = Bs + (kt % STAGES) * B_STAGE;
    { const int nx = kt + STAGES - 1; if (nx < nkt) load_stage(nx % STAGES, nx); cp_async_commit(); }
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      if (kk < BK / 4 - 1) {
        ld_frags(as, bs, kk + 1, a[(kk + 1) & 1], b[(kk + 1) & 1]);
      } else {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = (kt + 1) % STAGES;
        ld_frags(As + nk * A_STAGE, Bs + nk * B_STAGE, 0, a[(kk + 1) & 1], b[(kk + 1) & 1]);
      }
#pragma unroll
      for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], a[kk & 1][i], b[kk & 1][j]);
    }
  }
  cp_async_wait<0>();
  if (TN) {
    double* out = p.C + (size_t)z * p.N * p.M;
#pragma unroll
    for (int i = 0; i < FM; i++) {
      const int m = m0 + wm + 8 * i + g;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < FN; j++) {
        const int n = n0 + wn + 8 * j + 2 * t;
        if (n < p.N) out[(size_t)n * p.M + m] = acc[i][j][0];
        if (n + 1 < p.N) out[(size_t)(n + 1) * p.M + m] = acc[i][j][1];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < FM; i++) {
      const int m = m0 + wm + 8 * i + g;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < FN; j++) {
        const int n = n0 + wn + 8 * j + 2 * t;
        if (n < p.N) p.C[(size_t)n * p.ldc + m] = acc[i][j][0];
        if (n + 1 < p.N) p.C[(size_t)(n + 1) * p.ldc + m] = acc[i][j][1];
      }
    }
  }
}

template <bool TN, int WM, int WN, int WARPS_M, int WARPS_N, int STAGES, int MINB = 1>
float run(const char* name, GemmParams p, int nsplit, int reps) {
  constexpr int BM = WM * WARPS_M, BN = WN * WARPS_N;
  constexpr int A_STAGE = TN ? BM * BK : BK * (BM + 4);
  const size_t smem = sizeof(double) * STAGES * (A_STAGE + BN * BK);
  auto kern = k_gemm<TN, WM, WN, WARPS_M, WARPS_N, STAGES, MINB>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.tiles_m = (p.M + BM - 1) / BM; p.tiles_n = (p.N + BN - 1) / BN;
  p.nsplit = nsplit; p.kchunk = (((p.K + nsplit - 1) / nsplit) + BK - 1) / BK * BK;
  const int grid = p.tiles_m * p.tiles_n * nsplit;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<grid, WARPS_M * WARPS_N * 32, smem>>>(p); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) kern<<<grid, WARPS_M * WARPS_N * 32, smem>>>(p);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  const double tf = 2.0 * p.M * p.N * (double)p.K / (ms * 1e-3) / 1e12;
  printf("%-44s grid %5d smem %6zu  %8.3f ms  %6.2f TFLOP/s\n", name, grid, smem, ms, tf);
  return ms;
}

int main(int argc, char** argv) {
  const int npw = argc > 1 ? atoi(argv[1]) : 144057, nprojs = argc > 2 ? atoi(argv[2]) : 9216, ndat = argc > 3 ? atoi(argv[3]) : 128;
  const long long K2 = 2LL * npw;
  double *P, *X, *G, *Y, *part;
  CK(cudaMalloc(&P, sizeof(double) * K2 * nprojs));
  CK(cudaMalloc(&X, sizeof(double) * K2 * ndat));
  CK(cudaMalloc(&Y, sizeof(double) * K2 * ndat));
  CK(cudaMalloc(&G, sizeof(double) * (size_t)nprojs * ndat));
  CK(cudaMalloc(&part, sizeof(double) * (size_t)nprojs * ndat * 8));
  k_fill<<<148 * 8, 256>>>(P, (size_t)K2 * nprojs, 1);
  k_fill<<<148 * 8, 256>>>(X, (size_t)K2 * ndat, 2);
  k_fill<<<148 * 8, 256>>>(G, (size_t)nprojs * ndat, 3);
  CK(cudaDeviceSynchronize());
  const int reps = 3;
  // ---- NN: Y(2npw, ndat) = P(2npw, nprojs) G(nprojs, ndat)
  GemmParams nn{}; nn.M = (int)K2; nn.N = ndat; nn.K = nprojs; nn.A = P; nn.lda = K2; nn.B = G; nn.ldb = nprojs; nn.C = Y; nn.ldc = K2;
  run<false, 64, 32, 2, 4, 4>("NN 8 warps 64x32 (128x128) 4 stages", nn, 1, reps);
  {
    std::vector<double> y(4); const int ms[4] = {0, 77, 4099, (int)K2 - 1}, ns[4] = {0, 5, 64, ndat - 1};
    double maxerr = 0;
    for (int q = 0; q < 4; q++) {
      CK(cudaMemcpy(&y[q], Y + (size_t)ns[q] * K2 + ms[q], 8, cudaMemcpyDeviceToHost));
      double ref = 0; for (int k = 0; k < nprojs; k++) ref += hashval((uint64_t)k * K2 + ms[q], 1) * hashval((uint64_t)ns[q] * nprojs + k, 3);
      maxerr = fmax(maxerr, fabs(ref - y[q]) / (fabs(ref) + 1e-30));
    }
    printf("   NN sampled rel err %.2e\n", maxerr);
  }
  run<false, 32, 32, 2, 4, 4, 2>("NN 8 warps 32x32 (64x128) 4 stages 2 CTA/SM", nn, 1, reps);
  run<false, 32, 32, 4, 2, 4, 2>("NN 8 warps 32x32 (128x64) 4 stages 2 CTA/SM", nn, 1, reps);
  run<false, 32, 32, 2, 4, 3, 2>("NN 8 warps 32x32 (64x128) 3 stages 2 CTA/SM", nn, 1, reps);
  run<false, 32, 32, 4, 4, 4>("NN 16 warps 32x32 (128x128) 4 stages", nn, 1, reps);
  run<false, 16, 64, 4, 2, 4, 2>("NN 8 warps 16x64 (64x128) 4 stages 2 CTA/SM", nn, 1, reps);
  // ---- TN: partials(nprojs, ndat) = P^T X
  GemmParams tn{}; tn.M = nprojs; tn.N = ndat; tn.K = (int)K2; tn.A = P; tn.lda = K2; tn.B = X; tn.ldb = K2; tn.C = part;
  run<true, 64, 32, 2, 4, 4>("TN 8 warps 64x32 (128x128) 4 stages split 2", tn, 2, reps);
  {
    double maxerr = 0; const int ms[3] = {0, 1234, nprojs - 1}, ns[3] = {0, 77, ndat - 1};
    for (int q = 0; q < 3; q++) {
      double v0, v1;
      CK(cudaMemcpy(&v0, part + (size_t)ns[q] * nprojs + ms[q], 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&v1, part + (size_t)nprojs * ndat + (size_t)ns[q] * nprojs + ms[q], 8, cudaMemcpyDeviceToHost));
      double ref = 0; for (long long k = 0; k < K2; k++) ref += hashval((uint64_t)ms[q] * K2 + k, 1) * hashval((uint64_t)ns[q] * K2 + k, 2);
      maxerr = fmax(maxerr, fabs(ref - (v0 + v1)) / (fabs(ref) + 1e-30));
    }
    printf("   TN sampled rel err %.2e\n", maxerr);
  }
  run<true, 32, 32, 4, 4, 4>("TN 16 warps 32x32 (128x128) 4 stages split 2", tn, 2, reps);
  run<true, 32, 32, 4, 4, 4>("TN 16 warps 32x32 (128x128) 4 stages split 4", tn, 4, reps);
  run<true, 32, 32, 2, 4, 4, 2>("TN 8 warps 32x32 (64x128) 4 st split 2, 2 CTA/SM", tn, 2, reps);
  run<true, 32, 32, 2, 4, 4, 2>("TN 8 warps 32x32 (64x128) 4 st split 4, 2 CTA/SM", tn, 4, reps);
  run<true, 32, 32, 4, 4, 3>("TN 16 warps 32x32 (128x128) 3 stages split 2", tn, 2, reps);
  printf("done\n");
  return 0;
}
