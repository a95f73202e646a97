// xgBlock linear algebra for the eigensolver side of getghc: Gram matrices and rotations on the DMMA GEMMs of nonlop.cu,
// fused column-wise streaming kernels, dense sub-space eigenproblem (cuSOLVER, bound at run time), Rayleigh-Ritz.
//
// Reference semantics (not code): src/45_xgTools/m_xg.F90:1674-1976 (xgBlock_gemm, SPACE_CR G=0 correction :1802-1882),
// :3301-3413 (colwiseCymax), :4341-4846 (colwiseNorm2 / colwiseDotProduct), :5851-5898 (zero_im_g0),
// :2239-2861 (heevd / hegvd); src/45_xgTools/m_xg_ortho_RR.F90:251-571 (xg_RayleighRitz).
//
// Design: every streaming primitive the ChebFi2 loop strings together (scale, saxpy, scale back, saxpy ...) is ONE pass
// here (xg_cheb_next reads AX, X, Xprev once and writes Xnext once); the column-wise reductions use one CTA per column
// with a fixed-order tree, so results do not depend on the launch geometry.
#include "xg.cuh"
#include "nonlop.cuh"
#include "fourwf.cuh"   // g_kernel_launches
#include "context.cuh"
#include <algorithm>
#include <vector>
#ifndef ABI_EMU
#include <dlfcn.h>
#include <cusolverDn.h>
#endif

namespace abi {

#ifndef ABI_EMU
namespace {
struct XgWorkspace {
  double* d = nullptr; size_t cap = 0;
  double* get(size_t n) {
    if (n > cap) { if (d) cudaFree(d); CUDA_CHECK(cudaMalloc(&d, sizeof(double) * n)); cap = n; }
    return d;
  }
  void release() { if (d) cudaFree(d); d = nullptr; cap = 0; }
};
XgWorkspace g_xgws[4];   // 0: rotation slab, 1: subA, 2: subB, 3: cusolver work

constexpr int kRedThreads = 512;

// fixed-order block reduction of up to 2 values
template <int NV> ABI_DEV void block_reduce(double (&v)[NV], double* red) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; i++) red[i * kRedThreads + tid] = v[i];
  __syncthreads();
  for (int w = kRedThreads / 2; w > 0; w >>= 1) {
    if (tid < w) {
#pragma unroll
      for (int i = 0; i < NV; i++) red[i * kRedThreads + tid] += red[i * kRedThreads + tid + w];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = red[i * kRedThreads];
}

// One CTA per column.  mode 0: <A|B> ; mode 1: |A|^2 (B unused)
// SPACE_CR: 2 sum(a b) over the 2*rows reals, minus a(1) b(1) (dot) or a(1)^2 + a(2)^2 (norm2) when me_g0 = 1
// SPACE_C : (sum ar br + ai bi, sum ar bi - ai br) ; SPACE_R: plain sum over `rows` reals
__global__ void __launch_bounds__(kRedThreads) k_colwise_dot(int space, int mode, int rows, const double* __restrict__ A, long long lda,
                                                             const double* __restrict__ B, long long ldb, double* __restrict__ out,
                                                             int me_g0) {
  __shared__ double red[2 * kRedThreads];
  const int col = blockIdx.x;
  const int cplx = space != SPACE_R;
  const double* a = A + (cplx ? 2 : 1) * lda * col;
  const double* b = mode == 0 ? B + (cplx ? 2 : 1) * ldb * col : a;
  double v[2] = {0.0, 0.0};
  if (cplx) {
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    for (int i = threadIdx.x; i < rows; i += kRedThreads) {
      const double2 x = a2[i], y = b2[i];
      v[0] += x.x * y.x + x.y * y.y;
      v[1] += x.x * y.y - x.y * y.x;
    }
  } else {
    for (int i = threadIdx.x; i < rows; i += kRedThreads) v[0] += a[i] * b[i];
  }
  block_reduce<2>(v, red);
  if (threadIdx.x == 0) {
    if (space == SPACE_CR) {
      double r = 2.0 * v[0];
      if (me_g0 == 1 && rows > 0) r -= (mode == 0) ? a[0] * b[0] : a[0] * a[0] + a[1] * a[1];
      out[col] = r;
    } else if (space == SPACE_C && mode == 0) {
      out[2 * col] = v[0]; out[2 * col + 1] = v[1];
    } else {
      out[col] = v[0];
    }
  }
}

// G=0 correction of the SPACE_CR Gram matrix (m_xg.F90:1862-1882): the K=2*rows product counted the G=0 row twice
//   W += -2 alpha (a0r b0r + a0i b0i) + alpha a0r b0r        (alpha = 1 here)
__global__ void k_gram_g0(int na, int nb, const double* __restrict__ A, long long lda, const double* __restrict__ B, long long ldb,
                          double* __restrict__ W, long long ldw) {
  const long long total = (long long)na * nb;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % na); const long long j = idx / na;
    const double a0 = A[2 * lda * i], a1 = A[2 * lda * i + 1], b0 = B[2 * ldb * j], b1 = B[2 * ldb * j + 1];
    W[j * ldw + i] += -2.0 * (a0 * b0 + a1 * b1) + a0 * b0;
  }
}

__global__ void k_zero_im_g0(int ncols, double* __restrict__ X, long long ldx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ncols) X[2 * ldx * j + 1] = 0.0;
}

// elementwise kernels on the real view: blockIdx.y = column, nreal doubles per column (even when complex)
__global__ void k_cymax(int nreal, double* __restrict__ A, long long lda, const double* __restrict__ da, const double* __restrict__ B,
                        long long ldb, const double* __restrict__ W, long long ldw) {
  const int col = blockIdx.y;
  const double d = da[col];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreal; i += gridDim.x * blockDim.x)
    A[lda * col + i] = -d * B[ldb * col + i] + W[ldw * col + i];
}

__global__ void k_scale_cols(int nreal, double* __restrict__ X, long long ldx, const double* __restrict__ s) {
  const int col = blockIdx.y;
  const double f = s[col];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreal; i += gridDim.x * blockDim.x) X[ldx * col + i] *= f;
}

__global__ void k_cheb_next(int nreal, double* __restrict__ Xn, long long ldn, const double* __restrict__ AX, long long lda,
                            const double* __restrict__ X, long long ldx, const double* __restrict__ Xp, long long ldp, double center,
                            double scale) {
  const int col = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreal; i += gridDim.x * blockDim.x) {
    double v = (AX[lda * col + i] - center * X[ldx * col + i]) * scale;
    if (Xp) v -= Xp[ldp * col + i];
    Xn[ldn * col + i] = v;
  }
}

__global__ void k_add(int nreal, double* __restrict__ X, long long ldx, const double* __restrict__ P, long long ldp) {
  const int col = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreal; i += gridDim.x * blockDim.x) X[ldx * col + i] += P[ldp * col + i];
}

__global__ void k_sub(int nreal, double* __restrict__ X, long long ldx, const double* __restrict__ P, long long ldp) {
  const int col = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreal; i += gridDim.x * blockDim.x) X[ldx * col + i] -= P[ldp * col + i];
}

__global__ void k_apply_diag(int nreal, int shift, double* __restrict__ X, long long ldx, const double* __restrict__ d) {
  const int col = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreal; i += gridDim.x * blockDim.x) X[ldx * col + i] *= d[i >> shift];
}

// keep the upper triangle (incl. diagonal) of an n x n column-major matrix, zero the rest (cplx doubles per element)
__global__ void k_zero_lower(int n, int cplx, double* __restrict__ A, long long lda) {
  const long long total = (long long)n * n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n); const long long j = idx / n;
    if (i > j) for (int c = 0; c < cplx; c++) A[cplx * (j * lda + i) + c] = 0.0;
  }
}

dim3 ew_grid(int nreal, int ncols) { return dim3(std::max(1, std::min(64, ceil_div(nreal, 256 * 4))), ncols); }
int real_rows(int space, int rows) { return space == SPACE_R ? rows : 2 * rows; }
long long real_ld(int space, long long ld) { return space == SPACE_R ? ld : 2 * ld; }

// ---------------------------------------------------------------------------------------------------------
// cuSOLVER, bound at run time (the library is not a link dependency: only the Rayleigh-Ritz needs it)
// ---------------------------------------------------------------------------------------------------------
struct Cusolver {
  void* lib = nullptr; cusolverDnHandle_t h = nullptr;
  decltype(&cusolverDnCreate) create; decltype(&cusolverDnSetStream) set_stream;
  decltype(&cusolverDnDsygvd_bufferSize) dsygvd_bs; decltype(&cusolverDnDsygvd) dsygvd;
  decltype(&cusolverDnZhegvd_bufferSize) zhegvd_bs; decltype(&cusolverDnZhegvd) zhegvd;
  decltype(&cusolverDnDsyevd_bufferSize) dsyevd_bs; decltype(&cusolverDnDsyevd) dsyevd;
  decltype(&cusolverDnZheevd_bufferSize) zheevd_bs; decltype(&cusolverDnZheevd) zheevd;
  decltype(&cusolverDnDpotrf_bufferSize) dpotrf_bs; decltype(&cusolverDnDpotrf) dpotrf;
  decltype(&cusolverDnZpotrf_bufferSize) zpotrf_bs; decltype(&cusolverDnZpotrf) zpotrf;
  decltype(&cusolverDnXtrtri_bufferSize) xtrtri_bs; decltype(&cusolverDnXtrtri) xtrtri;
};
Cusolver& cusolver() {
  static Cusolver cs;
  if (cs.h) return cs;
  const char* cands[] = {getenv("ABI_B200_CUSOLVER"), "libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11",
                         "/usr/local/cuda/targets/x86_64-linux/lib/libcusolver.so.11", "libcusolver.so"};
  for (const char* c : cands) { if (c && !cs.lib) cs.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL); }
  ABI_CHECK(cs.lib != nullptr, "xg_hegvd: cannot load libcusolver.so.11 (set ABI_B200_CUSOLVER to its path); there is no CPU fallback");
#define ABI_SYM(field, name) do { cs.field = reinterpret_cast<decltype(cs.field)>(dlsym(cs.lib, #name)); \
    ABI_CHECK(cs.field != nullptr, "xg_hegvd: symbol " #name " missing in libcusolver"); } while (0)
  ABI_SYM(create, cusolverDnCreate); ABI_SYM(set_stream, cusolverDnSetStream);
  ABI_SYM(dsygvd_bs, cusolverDnDsygvd_bufferSize); ABI_SYM(dsygvd, cusolverDnDsygvd);
  ABI_SYM(zhegvd_bs, cusolverDnZhegvd_bufferSize); ABI_SYM(zhegvd, cusolverDnZhegvd);
  ABI_SYM(dsyevd_bs, cusolverDnDsyevd_bufferSize); ABI_SYM(dsyevd, cusolverDnDsyevd);
  ABI_SYM(zheevd_bs, cusolverDnZheevd_bufferSize); ABI_SYM(zheevd, cusolverDnZheevd);
  ABI_SYM(dpotrf_bs, cusolverDnDpotrf_bufferSize); ABI_SYM(dpotrf, cusolverDnDpotrf);
  ABI_SYM(zpotrf_bs, cusolverDnZpotrf_bufferSize); ABI_SYM(zpotrf, cusolverDnZpotrf);
  ABI_SYM(xtrtri_bs, cusolverDnXtrtri_bufferSize); ABI_SYM(xtrtri, cusolverDnXtrtri);
#undef ABI_SYM
  ABI_CHECK(cs.create(&cs.h) == CUSOLVER_STATUS_SUCCESS, "xg_hegvd: cusolverDnCreate failed");
  return cs;
}
#define CUSOLVER_CHECK(call) ABI_CHECK((call) == CUSOLVER_STATUS_SUCCESS, "cuSOLVER failure in " #call)
}  // namespace

void xg_release_workspace() { for (auto& w : g_xgws) w.release(); }

void xg_gram(int space, int rows, int ncols_a, int ncols_b, const double* A, long long lda, const double* B, long long ldb,
             double* W, long long ldw, int me_g0, cudaStream_t st) {
  ABI_CHECK(space == SPACE_R || space == SPACE_C || space == SPACE_CR, "xgBlock_gemm: bad space");
  if (ncols_a == 0 || ncols_b == 0) return;
  if (space == SPACE_C) {
    zgemm_cn(ncols_a, ncols_b, rows, A, lda, B, ldb, W, ldw, 1.0, st);
  } else if (space == SPACE_R) {
    ABI_CHECK(rows % 2 == 0 && lda % 2 == 0 && ldb % 2 == 0, "xgBlock_gemm(SPACE_R): rows and leading dimensions must be even");
    dgemm_tn(ncols_a, ncols_b, rows, A, lda, B, ldb, W, ldw, 1.0, st);
  } else {
    ABI_CHECK(me_g0 == 0 || me_g0 == 1, "xgBlock me_g0 is not initialized");                  // m_xg.F90:4591-4596
    dgemm_tn(ncols_a, ncols_b, 2 * rows, A, 2 * lda, B, 2 * ldb, W, ldw, 2.0, st);
    if (me_g0 == 1 && rows > 0) {
      const int blocks = std::min(kNumSM * 4, (int)ceil_div<long long>((long long)ncols_a * ncols_b, 256));
      k_gram_g0<<<blocks, 256, 0, st>>>(ncols_a, ncols_b, A, lda, B, ldb, W, ldw);
      CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
  }
}

// upper = true: C is upper triangular (the inverse Cholesky factor of xg_Borthonormalize): output columns [j0, j1) only need
// the first j1 columns of A, which halves the flops of the trsm-equivalent product; blocks run from the last to the first so
// that the in-place case never reads a column block it has already overwritten.
// subtract = true: OUT -= A.C instead of OUT = A.C (xgBlock_gemm 'n','n' with alpha = -1, beta = 1).
static void gemm_nn_impl(int space, int rows, int k, int ncols_out, const double* A, long long lda, const double* C, long long ldc,
                         double* OUT, long long ldo, bool upper, cudaStream_t st, bool subtract = false) {
  if (rows == 0 || ncols_out == 0) return;
  const int M = real_rows(space, rows);
  const long long ldar = real_ld(space, lda), ldor = real_ld(space, ldo);
  ABI_CHECK(space == SPACE_C || ldc % 2 == 0, "xg_rotate: the sub-space matrix needs an even leading dimension");
  ABI_CHECK(space != SPACE_R || (rows % 2 == 0 && lda % 2 == 0), "xg_rotate(SPACE_R): rows and ld must be even");
  // row slabs are independent in A.C: product into a slab buffer, then store (the reference uses a full-size temporary
  // block, m_xg_ortho_RR.F90:524-531); slab = whole waves of 64-row CTA tiles, <= 256 MB
  const long long budget = (256LL << 20) / (8LL * ncols_out);
  long long slab = std::max<long long>(2 * kNumSM * 64, budget / (2 * kNumSM * 64) * (2 * kNumSM * 64));
  slab = std::min<long long>(slab, (M + 1) & ~1LL);
  double* tmp = g_xgws[0].get((size_t)slab * ncols_out);
  const int sc = sub_cplex(space);
  const int cb = upper ? 256 : ncols_out;                          // column block of the triangular variant
  for (long long m0 = 0; m0 < M; m0 += slab) {
    const int mlen = (int)std::min<long long>(slab, M - m0);
    for (int j0 = ((ncols_out - 1) / cb) * cb; j0 >= 0; j0 -= cb) {
      const int jb = std::min(cb, ncols_out - j0);
      const int kk = upper ? std::min(k, j0 + jb) : k;
      const double* Cj = C + (size_t)sc * ldc * j0;
      if (space == SPACE_C) zgemm_nn(mlen / 2, jb, kk, A + m0, lda, Cj, ldc, tmp, slab / 2, st);
      else dgemm_nn(mlen, jb, kk, A + m0, ldar, Cj, ldc, tmp, slab, st);
      if (subtract) {
        k_sub<<<ew_grid(mlen, jb), 256, 0, st>>>(mlen, OUT + m0 + (size_t)ldor * j0, ldor, tmp, slab);
        CUDA_CHECK(cudaGetLastError());
        g_kernel_launches++;
      } else {
        CUDA_CHECK(cudaMemcpy2DAsync(OUT + m0 + (size_t)ldor * j0, sizeof(double) * ldor, tmp, sizeof(double) * slab, sizeof(double) * mlen, jb,
                                     cudaMemcpyDeviceToDevice, st));
      }
    }
  }
}

void xg_gemm_nn(int space, int rows, int k, int ncols_out, const double* A, long long lda, const double* C, long long ldc, double* OUT,
                long long ldo, cudaStream_t st) {
  gemm_nn_impl(space, rows, k, ncols_out, A, lda, C, ldc, OUT, ldo, false, st);
}

void xg_gemm_nn_upper(int space, int rows, int k, int ncols_out, const double* A, long long lda, const double* C, long long ldc, double* OUT,
                      long long ldo, cudaStream_t st) {
  gemm_nn_impl(space, rows, k, ncols_out, A, lda, C, ldc, OUT, ldo, true, st);
}

// W(0:j1, j0:j1) for every column block: the upper block-triangle of A^H B, all that potrf / heevd / hegvd ('u') read
static void gram_upper(int space, int rows, int n, const double* A, long long lda, const double* B, long long ldb, double* W, long long ldw,
                       int me_g0, cudaStream_t st) {
  const int sc = sub_cplex(space), cb = 512;
  const long long colB = real_ld(space, ldb);
  for (int j0 = 0; j0 < n; j0 += cb) {
    const int jb = std::min(cb, n - j0);
    xg_gram(space, rows, j0 + jb, jb, A, lda, B + colB * j0, ldb, W + (size_t)sc * ldw * j0, ldw, me_g0, st);
  }
}

void xg_rotate(int space, int rows, int k, int ncols_out, double* X, long long ldx, const double* C, long long ldc, cudaStream_t st) {
  xg_gemm_nn(space, rows, k, ncols_out, X, ldx, C, ldc, X, ldx, st);
}

void xg_ortho_wrt_blocks(int space, int rows, int nprev, int n, double* V, long long ldv, const double* X0, long long ldx0,
                         const double* BX0, long long ldbx0, int me_g0, cudaStream_t st) {
  if (nprev == 0 || n == 0 || rows == 0) return;
  const int sc = sub_cplex(space);
  const long long ldw = (nprev + 1) & ~1LL;                              // even: K-padding of the product below
  double* buf = g_xgws[1].get((size_t)sc * ldw * n);
  CUDA_CHECK(cudaMemsetAsync(buf, 0, sizeof(double) * sc * ldw * n, st));
  xg_gram(space, rows, nprev, n, BX0, ldbx0, V, ldv, buf, ldw, me_g0, st);           // buffer = BX0^H var   (m_lobpcg2.F90:829)
  gemm_nn_impl(space, rows, nprev, n, X0, ldx0, buf, ldw, V, ldv, false, st, true);  // var = var - X0 buffer (:833)
}

void xg_add(int space, int rows, int ncols, double* X, long long ldx, const double* P, long long ldp, cudaStream_t st) {
  if (ncols == 0 || rows == 0) return;
  const int nreal = real_rows(space, rows);
  k_add<<<ew_grid(nreal, ncols), 256, 0, st>>>(nreal, X, real_ld(space, ldx), P, real_ld(space, ldp));
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void xg_apply_diag(int space, int rows, int ncols, double* X, long long ldx, const double* d, cudaStream_t st) {
  if (ncols == 0 || rows == 0) return;
  const int nreal = real_rows(space, rows);
  k_apply_diag<<<ew_grid(nreal, ncols), 256, 0, st>>>(nreal, space == SPACE_R ? 0 : 1, X, real_ld(space, ldx), d);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void xg_zero_im_g0(int space, int ncols, double* X, long long ldx, int me_g0, cudaStream_t st) {
  if (space != SPACE_CR || ncols == 0) return;
  ABI_CHECK(me_g0 >= 0, "xgBlock me_g0 is not initialized");
  if (me_g0 != 1) return;
  k_zero_im_g0<<<ceil_div(ncols, 256), 256, 0, st>>>(ncols, X, ldx);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void xg_colwise_dot(int space, int rows, int ncols, const double* A, long long lda, const double* B, long long ldb, double* dots,
                    int me_g0, cudaStream_t st) {
  if (ncols == 0) return;
  ABI_CHECK(space != SPACE_CR || me_g0 >= 0, "xgBlockA me_g0 is not initialized");
  k_colwise_dot<<<ncols, kRedThreads, 0, st>>>(space, 0, rows, A, lda, B, ldb, dots, me_g0);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void xg_colwise_norm2(int space, int rows, int ncols, const double* A, long long lda, double* norms, int me_g0, cudaStream_t st) {
  if (ncols == 0) return;
  ABI_CHECK(space != SPACE_CR || me_g0 >= 0, "xgBlock me_g0 is not initialized");
  k_colwise_dot<<<ncols, kRedThreads, 0, st>>>(space, 1, rows, A, lda, A, lda, norms, me_g0);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void xg_colwise_cymax(int space, int rows, int ncols, double* A, long long lda, const double* da, const double* B, long long ldb,
                      const double* W, long long ldw, cudaStream_t st) {
  if (ncols == 0 || rows == 0) return;
  const int nreal = real_rows(space, rows);
  k_cymax<<<ew_grid(nreal, ncols), 256, 0, st>>>(nreal, A, real_ld(space, lda), da, B, real_ld(space, ldb), W, real_ld(space, ldw));
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void xg_scale_cols(int space, int rows, int ncols, double* X, long long ldx, const double* s, cudaStream_t st) {
  if (ncols == 0 || rows == 0) return;
  const int nreal = real_rows(space, rows);
  k_scale_cols<<<ew_grid(nreal, ncols), 256, 0, st>>>(nreal, X, real_ld(space, ldx), s);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void xg_cheb_next(int space, int rows, int ncols, double* Xnext, long long ldn, const double* AX, long long lda, const double* X,
                  long long ldx, const double* Xprev, long long ldp, double center, double scale, cudaStream_t st) {
  if (ncols == 0 || rows == 0) return;
  const int nreal = real_rows(space, rows);
  ProfScope ps("cheb_next");
  k_cheb_next<<<ew_grid(nreal, ncols), 256, 0, st>>>(nreal, Xnext, real_ld(space, ldn), AX, real_ld(space, lda), X, real_ld(space, ldx),
                                                     Xprev, real_ld(space, ldp), center, scale);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

int xg_hegvd(int space, int n, double* A, long long lda, double* B, long long ldb, double* w, cudaStream_t st) {
  if (n == 0) return 0;
  Cusolver& cs = cusolver();
  CUSOLVER_CHECK(cs.set_stream(cs.h, st));
  ProfScope ps("hegvd");
  int lwork = 0;
  const cusolverEigType_t it = CUSOLVER_EIG_TYPE_1; const cusolverEigMode_t jz = CUSOLVER_EIG_MODE_VECTOR;
  const cublasFillMode_t up = CUBLAS_FILL_MODE_UPPER;
  int* d_info = reinterpret_cast<int*>(g_xgws[3].get(4));
  if (space == SPACE_C) {
    auto* a = reinterpret_cast<cuDoubleComplex*>(A); auto* b = reinterpret_cast<cuDoubleComplex*>(B);
    if (B) CUSOLVER_CHECK(cs.zhegvd_bs(cs.h, it, jz, up, n, a, (int)lda, b, (int)ldb, w, &lwork));
    else CUSOLVER_CHECK(cs.zheevd_bs(cs.h, jz, up, n, a, (int)lda, w, &lwork));
    double* work = g_xgws[3].get(4 + 2 * (size_t)lwork) + 4;
    d_info = reinterpret_cast<int*>(g_xgws[3].d);
    if (B) CUSOLVER_CHECK(cs.zhegvd(cs.h, it, jz, up, n, a, (int)lda, b, (int)ldb, w, reinterpret_cast<cuDoubleComplex*>(work), lwork, d_info));
    else CUSOLVER_CHECK(cs.zheevd(cs.h, jz, up, n, a, (int)lda, w, reinterpret_cast<cuDoubleComplex*>(work), lwork, d_info));
  } else {
    if (B) CUSOLVER_CHECK(cs.dsygvd_bs(cs.h, it, jz, up, n, A, (int)lda, B, (int)ldb, w, &lwork));
    else CUSOLVER_CHECK(cs.dsyevd_bs(cs.h, jz, up, n, A, (int)lda, w, &lwork));
    double* work = g_xgws[3].get(4 + (size_t)lwork) + 4;
    d_info = reinterpret_cast<int*>(g_xgws[3].d);
    if (B) CUSOLVER_CHECK(cs.dsygvd(cs.h, it, jz, up, n, A, (int)lda, B, (int)ldb, w, work, lwork, d_info));
    else CUSOLVER_CHECK(cs.dsyevd(cs.h, jz, up, n, A, (int)lda, w, work, lwork, d_info));
  }
  g_kernel_launches++;
  int info = 0;
  CUDA_CHECK(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  return info;
}

// U^-1 of the Cholesky factor of the m x m sub-space matrix A (upper): potrf 'u' + trtri, strictly-lower part zeroed so
// that the result can be used as a plain rotation matrix.  Returns potrf's info.
static int chol_inverse_upper(int sub_space, int m, double* A, long long lda, cudaStream_t st) {
  Cusolver& cs = cusolver();
  CUSOLVER_CHECK(cs.set_stream(cs.h, st));
  const int sc = sub_space == SPACE_C ? 2 : 1;
  int lwork = 0;
  const cublasFillMode_t up = CUBLAS_FILL_MODE_UPPER;
  if (sc == 2) CUSOLVER_CHECK(cs.zpotrf_bs(cs.h, up, m, reinterpret_cast<cuDoubleComplex*>(A), (int)lda, &lwork));
  else CUSOLVER_CHECK(cs.dpotrf_bs(cs.h, up, m, A, (int)lda, &lwork));
  size_t wdev = 0, whost = 0;
  const cudaDataType dt = sc == 2 ? CUDA_C_64F : CUDA_R_64F;
  CUSOLVER_CHECK(cs.xtrtri_bs(cs.h, up, CUBLAS_DIAG_NON_UNIT, m, dt, A, lda, &wdev, &whost));
  const size_t ndbl = 4 + std::max<size_t>((size_t)sc * lwork, (wdev + 7) / 8);
  double* base = g_xgws[3].get(ndbl);
  int* d_info = reinterpret_cast<int*>(base);
  double* work = base + 4;
  std::vector<char> hbuf(whost + 8);
  if (sc == 2) CUSOLVER_CHECK(cs.zpotrf(cs.h, up, m, reinterpret_cast<cuDoubleComplex*>(A), (int)lda, reinterpret_cast<cuDoubleComplex*>(work), lwork, d_info));
  else CUSOLVER_CHECK(cs.dpotrf(cs.h, up, m, A, (int)lda, work, lwork, d_info));
  int info = 0;
  CUDA_CHECK(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  g_kernel_launches++;
  if (info != 0) return info;
  CUSOLVER_CHECK(cs.xtrtri(cs.h, up, CUBLAS_DIAG_NON_UNIT, m, dt, A, lda, work, wdev, hbuf.data(), whost, d_info));
  CUDA_CHECK(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  k_zero_lower<<<std::min(kNumSM * 4, (int)ceil_div<long long>((long long)m * m, 256)), 256, 0, st>>>(m, sc, A, lda);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(st));
  g_kernel_launches += 2;
  return info;
}

int xg_chol_inverse(int sub_space, int m, double* A, long long lda, cudaStream_t st) { return chol_inverse_upper(sub_space, m, A, lda, st); }

int xg_b_orthonormalize(int space, int rows, int m, double* X, long long ldx, double* BX, long long ldbx, double* AX, long long ldax,
                        int me_g0, cudaStream_t st) {
  if (m == 0) return 0;
  const int sc = sub_cplex(space);
  const int sub_space = (space == SPACE_C) ? SPACE_C : SPACE_R;
  const long long ldw = (m + 1) & ~1LL;
  double* buf = g_xgws[1].get((size_t)sc * ldw * m);
  CUDA_CHECK(cudaMemsetAsync(buf, 0, sizeof(double) * sc * ldw * m, st));
  xg_zero_im_g0(space, m, X, ldx, me_g0, st);                          // m_xg_ortho_RR.F90:115-119
  if (BX != X) xg_zero_im_g0(space, m, BX, ldbx, me_g0, st);            // BX == X: norm-conserving caller (B = 1)
  if (AX) xg_zero_im_g0(space, m, AX, ldax, me_g0, st);
  gram_upper(space, rows, m, X, ldx, BX, ldbx, buf, ldw, me_g0, st);    // :122 (potrf 'u' reads the upper triangle only)
  const int info = chol_inverse_upper(sub_space, m, buf, ldw, st);      // potrf :125 (+ the inverse the trsm calls apply)
  if (info != 0) return info;                                           // "Cholesky decomposition did not work"
  gemm_nn_impl(space, rows, m, m, X, ldx, buf, ldw, X, ldx, true, st);  // trsm 'r','u','n' :135-142 (U^-1 is upper triangular)
  if (BX != X) gemm_nn_impl(space, rows, m, m, BX, ldbx, buf, ldw, BX, ldbx, true, st);
  if (AX) gemm_nn_impl(space, rows, m, m, AX, ldax, buf, ldw, AX, ldax, true, st);
  return 0;
}

int xg_rayleigh_ritz_xwp(int space, int rows, int n, int nvar, double* XWP, double* AXWP, double* BXWP, long long ld, double* eig,
                         int me_g0, cudaStream_t st) {
  ABI_CHECK(nvar == 2 || nvar == 3, "xg_RayleighRitz: nvar must be 2 (XW) or 3 (XWP)");
  if (n == 0) return 0;
  const int sc = sub_cplex(space);
  const int sub_space = (space == SPACE_C) ? SPACE_C : SPACE_R;
  const int sub = nvar * n;
  const long long ldw = (sub + 1) & ~1LL;
  const size_t blk = (size_t)2 * ld * n;                                 // doubles per n-column block
  double* subA = g_xgws[1].get((size_t)sc * ldw * sub);
  double* subB = g_xgws[2].get((size_t)sc * ldw * sub);
  CUDA_CHECK(cudaMemsetAsync(subA, 0, sizeof(double) * sc * ldw * sub, st));
  CUDA_CHECK(cudaMemsetAsync(subB, 0, sizeof(double) * sc * ldw * sub, st));
  xg_zero_im_g0(space, sub, XWP, ld, me_g0, st);                          // :376-398
  xg_zero_im_g0(space, sub, AXWP, ld, me_g0, st);
  if (BXWP != XWP) xg_zero_im_g0(space, sub, BXWP, ld, me_g0, st);       // BXWP == XWP: norm-conserving caller (B = 1)
  // upper-triangular block columns of the sub-space matrices (:384-412)
  for (int v = 0; v < nvar; v++) {
    const int rws = (v + 1) * n;
    xg_gram(space, rows, rws, n, XWP, ld, AXWP + v * blk, ld, subA + (size_t)sc * ldw * v * n, ldw, me_g0, st);
    xg_gram(space, rows, rws, n, XWP, ld, BXWP + v * blk, ld, subB + (size_t)sc * ldw * v * n, ldw, me_g0, st);
  }
  const int info = xg_hegvd(sub_space, sub, subA, ldw, subB, ldw, eig, st);   // EIGENVD, :465
  if (info != 0) return info;
  // X <- X Cwp(0:n) ; P <- WP Cwp(n:sub) ; X += P  (:524-560), same for the A and B blocks.  For VAR_XW the reference's
  // product runs over [W P] with P zeroed beforehand: the W rows alone give the same result.
  // rows n:sub of the first n eigenvectors, copied to a 16-byte aligned, K-padded matrix (subB is free after hegvd)
  const long long ldc1 = (sub - n + 1) & ~1LL;
  double* c1 = subB;
  CUDA_CHECK(cudaMemsetAsync(c1, 0, sizeof(double) * sc * ldc1 * n, st));
  CUDA_CHECK(cudaMemcpy2DAsync(c1, sizeof(double) * sc * ldc1, subA + (size_t)sc * n, sizeof(double) * sc * ldw,
                               sizeof(double) * sc * (sub - n), n, cudaMemcpyDeviceToDevice, st));
  double* blocks[3] = {XWP, AXWP, BXWP};
  for (int ib = 0; ib < 3; ib++) {
    if (ib == 2 && BXWP == XWP) continue;                                 // B blocks alias the X blocks: already rotated
    double* B0 = blocks[ib];
    xg_rotate(space, rows, n, n, B0, ld, subA, ldw, st);
    xg_gemm_nn(space, rows, sub - n, n, B0 + blk, ld, c1, ldc1, B0 + 2 * blk, ld, st);
    xg_add(space, rows, n, B0, ld, B0 + 2 * blk, ld, st);
  }
  return 0;
}

int xg_rayleigh_ritz(int space, int rows, int n, double* X, long long ldx, double* AX, long long ldax, double* BX, long long ldbx,
                     double* eig, bool solve_ax_bx, int me_g0, cudaStream_t st) {
  if (n == 0) return 0;
  const bool bx_is_x = (BX == nullptr || BX == X);
  const int sc = sub_cplex(space);
  const int sub_space = (space == SPACE_C) ? SPACE_C : SPACE_R;
  const long long ldw = (n + 1) & ~1LL;                                  // even: K-padding of the rotation GEMM
  double* subA = g_xgws[1].get((size_t)sc * ldw * n);
  CUDA_CHECK(cudaMemsetAsync(subA, 0, sizeof(double) * sc * ldw * n, st));
  double* subB = nullptr;
  // m_xg_ortho_RR.F90:376-380
  xg_zero_im_g0(space, n, X, ldx, me_g0, st);
  xg_zero_im_g0(space, n, AX, ldax, me_g0, st);
  if (!bx_is_x) xg_zero_im_g0(space, n, BX, ldbx, me_g0, st);
  gram_upper(space, rows, n, X, ldx, AX, ldax, subA, ldw, me_g0, st);     // :384 (heevd / hegvd 'u' read the upper triangle only)
  if (solve_ax_bx) {
    subB = g_xgws[2].get((size_t)sc * ldw * n);
    CUDA_CHECK(cudaMemsetAsync(subB, 0, sizeof(double) * sc * ldw * n, st));
    gram_upper(space, rows, n, X, ldx, bx_is_x ? X : BX, bx_is_x ? ldx : ldbx, subB, ldw, me_g0, st);   // :388
  }
  const int info = xg_hegvd(sub_space, n, subA, ldw, subB, ldw, eig, st);  // heevd :441 / hegvd :465
  if (info != 0) return info;
  // X, AX, BX <- . Cwp (:524-531)
  xg_rotate(space, rows, n, n, X, ldx, subA, ldw, st);
  xg_rotate(space, rows, n, n, AX, ldax, subA, ldw, st);
  if (!bx_is_x) xg_rotate(space, rows, n, n, BX, ldbx, subA, ldw, st);
  return 0;
}

#else   // ABI_EMU
void xg_release_workspace() {}
#endif

}  // namespace abi
