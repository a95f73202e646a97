// EXPERIMENTAL (opt-in, ABI_B200_OZAKI=1): the two projector contractions of gemm_nonlop by error-free int8 slicing
// (Ozaki scheme I) instead of FP64 DMMA.  Default OFF: the product path is the FP64 DMMA kernels of nonlop.cu.
//
//   A (K x M, the K-contiguous operand) and B (K x N) are split, column by column, against a per-column power-of-two scale
//   into S = 7 signed slices of 7 bits:  x = 2^e * sum_s q_s 2^(-7(s+1)),  q_s integer in [-64, 64].
//   Every slice product A_s^T B_t is EXACT in int32 (K * 64 * 64 < 2^31 for K <= 524 288) and runs on the int8 tensor pipe;
//   the FP64 result is  2^(ea+eb) * sum_{s+t<S} 2^(-7(s+t+2)) (A_s^T B_t), summed smallest terms first.
//   28 slice products reproduce the FP64 GEMM to ~3e-13 relative (tools/ozaki_study.py).
//
// The int8 GEMMs are the hand-written tcgen05 kernel of igemm_tc.cuh (TMA + mbarrier ring + kind::i8 MMA into TMEM),
// measured at 97 % (K = 2 npw) / 78 % (K = nprojs) of cuBLASLt's int8 GEMM on the Si-512 shapes (tools/igemm_lab.cu,
// tools/probe_int8.py).  Nothing here is used unless the knob is set; bench.py reports it under a separate key.
#include "nonlop.cuh"
#include "fourwf.cuh"
#include "context.cuh"
#include <algorithm>
#ifndef ABI_EMU
#include "igemm_tc.cuh"
#endif

namespace abi {

#ifndef ABI_EMU
namespace {
constexpr int kS = 7, kBits = 7;

// C(M x N, int32, ldc) = A(K x M, int8, lda)^T B(K x N, int8, ldb)
void igemm_tn(int M, int N, int K, const int8_t* A, long long lda, const int8_t* B, long long ldb, int32_t* C, long long ldc, cudaStream_t st) {
  ProfScope ps("ozaki_igemm");
  igemm_tc(M, N, K, A, lda, B, ldb, C, ldc, st);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

// ---- slicing of a K-contiguous operand: X (K x C, FP64, ld) -> q[s][c][Kp] int8 and e[c] ----
// pairs = 1 (complex opernla): effective column 2n is column n of X, effective column 2n+1 is (-i psi_n) in the real view,
// i.e. value(i) = x[i+1] for even i, -x[i-1] for odd i -- formed while slicing, never materialised in FP64
__global__ void __launch_bounds__(256) k_slice_cols(const double* __restrict__ X, long long ld, int K, long long Kp, int ncols,
                                                     int8_t* __restrict__ q, double* __restrict__ e, int pairs) {
  __shared__ double red[256];
  const int c = blockIdx.x;
  const bool rot = pairs && (c & 1);
  const double* x = X + ld * (pairs ? (c >> 1) : c);
  double m = 0.0;
  for (int i = threadIdx.x; i < K; i += 256) m = fmax(m, fabs(x[i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) { if (threadIdx.x < w) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + w]); __syncthreads(); }
  const double amax = red[0];
  int ex = 0;
  if (amax > 0.0) { frexp(amax, &ex); ex += 1; }                 // |x| / 2^ex < 1/2
  if (threadIdx.x == 0) e[c] = (double)ex;
  const double inv = ldexp(1.0, -ex);
  for (long long i = threadIdx.x; i < Kp; i += 256) {
    double r = 0.0;
    if (i < K) r = (rot ? ((i & 1) ? -x[i - 1] : x[i + 1]) : x[i]) * inv;
#pragma unroll
    for (int s = 0; s < kS; s++) {
      r *= 128.0;
      const double v = rint(r);
      r -= v;
      q[((size_t)s * ncols + c) * Kp + i] = (int8_t)(int)v;
    }
  }
}

// ---- slicing of the M-contiguous P for opernlb: P (M x K, ld) -> q[s][m][Kp] (K-contiguous per row m) and e[m] ----
// cplx = 1: row m of the effective operand holds P[m,:] and (iP)[m,:] = -+P[m^1,:], so its scale covers both rows
__global__ void k_row_exponent(const double* __restrict__ P, long long ld, int M, int K, double* __restrict__ e, int cplx) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double mx = 0.0;
  for (int k = 0; k < K; k++) {
    mx = fmax(mx, fabs(P[(long long)k * ld + m]));
    if (cplx) mx = fmax(mx, fabs(P[(long long)k * ld + (m ^ 1)]));
  }
  int ex = 0;
  if (mx > 0.0) { frexp(mx, &ex); ex += 1; }
  e[m] = (double)ex;
}
// cplx = 1: effective K index k = 2p + c: c = 0 -> P[m,p], c = 1 -> (iP)[m,p] = (m even ? -P[m+1,p] : P[m-1,p])
__global__ void __launch_bounds__(256) k_slice_rows_t(const double* __restrict__ P, long long ld, int M, int K, long long Kp, long long Mp,
                                                       const double* __restrict__ e, int8_t* __restrict__ q, int cplx) {
  __shared__ int8_t tile[kS][32][33];
  const int m0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
  for (int kk = ty; kk < 32; kk += 8) {
    const int m = m0 + tx, k = k0 + kk;
    double r = 0.0;
    if (m < M && k < K) {
      double v;
      if (!cplx) v = P[(long long)k * ld + m];
      else if (!(k & 1)) v = P[(long long)(k >> 1) * ld + m];
      else v = (m & 1) ? P[(long long)(k >> 1) * ld + m - 1] : -P[(long long)(k >> 1) * ld + m + 1];
      r = v * ldexp(1.0, -(int)e[m]);
    }
#pragma unroll
    for (int s = 0; s < kS; s++) { r *= 128.0; const double v = rint(r); r -= v; tile[s][kk][tx] = (int8_t)(int)v; }
  }
  __syncthreads();
  for (int mm = ty; mm < 32; mm += 8) {
    const int m = m0 + mm, k = k0 + tx;
    if (m < Mp && k < Kp) {
#pragma unroll
      for (int s = 0; s < kS; s++) q[((size_t)s * Mp + m) * Kp + k] = tile[s][tx][mm];
    }
  }
}

// ---- FP64 recombination: out(m, n) = 2^(ea[m]+eb[n]) sum_g 2^(-7(g+2)) sum_{s+t=g} C_s(m, t N + n), smallest g first ----
// mode 0: store to out (ldo) ; mode 1: getghc fusion  ghc = kin[m/2] < filter ? ghc + v : 0, optional copy of v to vout
struct CombineParams {
  const int32_t* C[kS]; long long ldc; int M, N, Nstride;
  const double* ea; const double* eb;
  double* out; long long ldo; int mode;
  double* vout; const double* kin; double kin_filter; const double* add;
};
__global__ void k_combine(CombineParams p) {
  const long long total = (long long)p.M * p.N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % p.M); const int n = (int)(idx / p.M);
    double acc = 0.0;
#pragma unroll
    for (int g = kS - 1; g >= 0; g--) {
      double part = 0.0;
#pragma unroll
      for (int s = 0; s <= g; s++) part += (double)p.C[s][(long long)((g - s) * p.Nstride + n) * p.ldc + m];
      acc += ldexp(part, -kBits * (g + 2));
    }
    double v = ldexp(acc, (int)(p.ea[m] + p.eb[n]));
    const long long o = (long long)n * p.ldo + m;
    if (p.mode == 0) { if (p.add) v += p.add[o]; p.out[o] = v; }
    else {
      if (p.vout) p.vout[o] = v;
      v += p.out[o];
      if (!(p.kin[m >> 1] < p.kin_filter)) v = 0.0;
      p.out[o] = v;
    }
  }
}

struct Buf8 { void* d = nullptr; size_t cap = 0;
  void* get(size_t bytes) { if (bytes > cap) { if (d) cudaFree(d); CUDA_CHECK(cudaMalloc(&d, bytes)); cap = bytes; } return d; }
  void release() { if (d) cudaFree(d); d = nullptr; cap = 0; } };
Buf8 g_oz[4];   // 0: B slices, 1: B exponents, 2: int32 products, 3: spare
}  // namespace

static int g_ozaki_on = -1;
bool ozaki_enabled() {
  if (g_ozaki_on < 0) { const char* e = getenv("ABI_B200_OZAKI"); g_ozaki_on = (e && atoi(e) != 0) ? 1 : 0; }
  return g_ozaki_on == 1;
}
void ozaki_set_enabled(int flag) { g_ozaki_on = flag ? 1 : 0; }

void OzakiP::release() {
  for (void** p : {(void**)&a_k, (void**)&a_m, (void**)&ea_k, (void**)&ea_m}) { if (*p) cudaFree(*p); *p = nullptr; }
  npw = nprojs = 0; failed = false;
}

// slices of P for both contractions (once per k-point): 2 x 7 int8 copies of P
void ozaki_prepare(const Projectors& P, OzakiP& oz, cudaStream_t st) {
  oz.release();
  const int K1 = 2 * P.npw, M1 = P.nprojs;
  const int cplx = P.istwf_k == 1 ? 1 : 0;
  const int K2 = (cplx ? 2 : 1) * M1;                              // opernlb: K = nprojs (real) or 2 nprojs ((P, iP) pairs)
  oz.npw = P.npw; oz.nprojs = P.nprojs; oz.cplx = cplx;
  oz.kp1 = ((long long)K1 + 127) / 128 * 128; oz.kp2 = ((long long)K2 + 127) / 128 * 128; oz.mp2 = ((long long)K1 + 3) / 4 * 4;
  // the int8 copies are an optimisation: when the device cannot hold them this k-point stays on the FP64 DMMA kernels
  if (cudaMalloc(&oz.a_k, (size_t)kS * M1 * oz.kp1) != cudaSuccess || cudaMalloc(&oz.a_m, (size_t)kS * oz.mp2 * oz.kp2) != cudaSuccess) {
    cudaGetLastError();
    oz.release();
    oz.failed = true;
    fprintf(stderr, "\n--- !WARNING\nmessage: |\n    abinit_b200: not enough device memory for the int8-sliced projectors (%.1f GB); "
                    "this k-point uses the FP64 path\n...\n", 1e-9 * ((double)kS * M1 * oz.kp1 + (double)kS * oz.mp2 * oz.kp2));
    return;
  }
  CUDA_CHECK(cudaMalloc(&oz.ea_k, sizeof(double) * M1));
  CUDA_CHECK(cudaMalloc(&oz.ea_m, sizeof(double) * oz.mp2));
  CUDA_CHECK(cudaMemsetAsync(oz.ea_m, 0, sizeof(double) * oz.mp2, st));
  k_slice_cols<<<M1, 256, 0, st>>>(P.d_p, (long long)K1, K1, oz.kp1, M1, oz.a_k, oz.ea_k, 0);
  k_row_exponent<<<ceil_div(K1, 256), 256, 0, st>>>(P.d_p, (long long)K1, K1, M1, oz.ea_m, cplx);
  k_slice_rows_t<<<dim3(ceil_div<long long>(oz.mp2, 32), ceil_div<long long>(oz.kp2, 32)), 256, 0, st>>>(P.d_p, (long long)K1, K1, K2, oz.kp2, oz.mp2,
                                                                                                  oz.ea_m, oz.a_m, cplx);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches += 3;
}

// part[n][m] (FP64, the nsplit = 1 partial buffer of opernla) = P^T psi through the slice products
void ozaki_project(const OzakiP& oz, const double* vectin, int ndat_bands, double* part, cudaStream_t st) {
  const int M = oz.nprojs, K = 2 * oz.npw;
  const int ndat = (oz.cplx ? 2 : 1) * ndat_bands;                  // complex: (psi, -i psi) column pairs -> (Re, Im) of P^H psi
  const int nd4 = (ndat + 3) & ~3;                                // cuBLASLt int8: every extent a multiple of 4
  int8_t* bq = (int8_t*)g_oz[0].get((size_t)kS * nd4 * oz.kp1);
  double* eb = (double*)g_oz[1].get(sizeof(double) * nd4);
  if (nd4 != ndat) CUDA_CHECK(cudaMemsetAsync(bq, 0, (size_t)kS * nd4 * oz.kp1, st));
  k_slice_cols<<<ndat, 256, 0, st>>>(vectin, (long long)K, K, oz.kp1, nd4, bq, eb, oz.cplx);
  size_t tot = 0; for (int s = 0; s < kS; s++) tot += (size_t)M * nd4 * (kS - s);
  int32_t* c = (int32_t*)g_oz[2].get(sizeof(int32_t) * tot);
  CombineParams p{};
  size_t off = 0;
  for (int s = 0; s < kS; s++) {
    const int ns = nd4 * (kS - s);
    igemm_tn(M, ns, (int)oz.kp1, oz.a_k + (size_t)s * M * oz.kp1, oz.kp1, bq, oz.kp1, c + off, M, st);
    p.C[s] = c + off; off += (size_t)M * ns;
  }
  p.ldc = M; p.M = M; p.N = ndat; p.Nstride = nd4; p.ea = oz.ea_k; p.eb = eb; p.out = part; p.ldo = M; p.mode = 0;
  ProfScope ps("ozaki_combine");
  k_combine<<<std::min(kNumSM * 8, (int)ceil_div<long long>((long long)M * ndat, 256)), 256, 0, st>>>(p);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches += 2;
}

// vect = P z through the slice products; mode as CombineParams (fusion with the getghc epilogue)
void ozaki_expand(const OzakiP& oz, const double* z, long long ldz, int ndat, double* out, int fuse, double* vout, const double* kin,
                  double kin_filter, const double* add, cudaStream_t st, int nslabs, void (*after_slab)(void*, int, int), void* user) {
  const int M = 2 * oz.npw, K = (oz.cplx ? 2 : 1) * oz.nprojs;     // complex: z is (re, im) interleaved = the K index of (P, iP)
  const int nd4 = (ndat + 3) & ~3;
  int8_t* bq = (int8_t*)g_oz[0].get((size_t)kS * nd4 * oz.kp2);
  double* eb = (double*)g_oz[1].get(sizeof(double) * nd4);
  if (nd4 != ndat) CUDA_CHECK(cudaMemsetAsync(bq, 0, (size_t)kS * nd4 * oz.kp2, st));
  k_slice_cols<<<ndat, 256, 0, st>>>(z, ldz, K, oz.kp2, nd4, bq, eb, 0);
  g_kernel_launches++;
  // row slabs (whole 128-row tiles): bounds the int32 workspace and lets the caller ship finished rows of ghc to the host
  // while the next slab is computed (the after_slab hook of NonlopFusion)
  nslabs = std::max(1, std::min(nslabs, (M + 127) / 128));
  long long slab = ((((long long)M + nslabs - 1) / nslabs) + 127) / 128 * 128;
  size_t tot = 0; for (int s = 0; s < kS; s++) tot += (size_t)slab * nd4 * (kS - s);
  int32_t* c = (int32_t*)g_oz[2].get(sizeof(int32_t) * tot);
  for (long long m0 = 0; m0 < M; m0 += slab) {
    const int mlen = (int)std::min<long long>(slab, M - m0);
    const int mrows = (int)std::min<long long>(slab, oz.mp2 - m0);          // rows present in the int8 copy (multiple of 4)
    CombineParams p{};
    size_t off = 0;
    for (int s = 0; s < kS; s++) {
      const int ns = nd4 * (kS - s);
      igemm_tn(mrows, ns, (int)oz.kp2, oz.a_m + ((size_t)s * oz.mp2 + m0) * oz.kp2, oz.kp2, bq, oz.kp2, c + off, slab, st);
      p.C[s] = c + off; off += (size_t)slab * ns;
    }
    p.ldc = slab; p.M = mlen; p.N = ndat; p.Nstride = nd4; p.ea = oz.ea_m + m0; p.eb = eb; p.out = out + m0; p.ldo = M; p.mode = fuse ? 1 : 0;
    p.vout = vout ? vout + m0 : nullptr; p.kin = kin ? kin + m0 / 2 : nullptr; p.kin_filter = kin_filter; p.add = add ? add + m0 : nullptr;
    {
      ProfScope ps("ozaki_combine");
      k_combine<<<std::min(kNumSM * 16, (int)ceil_div<long long>((long long)mlen * ndat, 256)), 256, 0, st>>>(p);
      CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
    if (after_slab) after_slab(user, (int)(m0 / 2), (int)((m0 + mlen) / 2));
  }
}

void ozaki_release_workspace() { for (auto& b : g_oz) b.release(); }

#else
bool ozaki_enabled() { return false; }
void ozaki_set_enabled(int) {}
void OzakiP::release() {}
void ozaki_release_workspace() {}
#endif

}  // namespace abi
