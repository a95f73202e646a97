// gemm_nonlop for sm_100a: projector overlaps P^H psi, D_ij/S_ij application, P.z back-accumulation.
//
// Reference semantics (not code): src/66_nonlocal/m_gemm_nonlop.F90:191-1242 (driver),
// m_gemm_nonlop_projectors.F90:792-1038 (prep_projectors), m_opernla_gemm.F90:361-712 (P^H psi),
// m_opernlc_ylm_allwf.F90:308-447,1253-1295 (D/S), m_opernlb_gemm.F90:353-837 (P z).
//
// Design: P is kept as ONE real matrix, the real view of P(2,npw,nprojs): (2 npw) x nprojs, column-major.
//  * istwf_k>=2 (real projections): gx = 2 (P_r^T psi_r + P_i^T psi_i) is a single real GEMM with K = 2 npw on
//    the interleaved data -- no de-interleave pass (the reference splits into P_r/P_i and psi_r/psi_i,
//    m_opernla_gemm.F90:569-689); the istwf_k=2 G=0 correction is a rank-1 fix-up in the reduction epilogue.
//  * istwf_k==1 (complex): P^H psi is the same real GEMM on 2 ndat columns, the odd ones being (-i psi), which
//    is formed while loading B fragments from shared memory (swap + sign), never materialised.
//    P z is a real GEMM with K = 2 nprojs whose odd k are (i P), formed while loading A fragments.
//  * The GEMM kernels issue FP64 tensor-core MMAs (mma.sync m8n8k4 f64 = DMMA.8x8x4 in SASS; tcgen05 has no
//    f64 kind) from a 4-stage cp.async pipeline; 16 warps x (32x32) warp tiles per 128x128 CTA tile.
//  * opernla is K-long and output-small: split-K across CTAs with deterministic partial buffers; the reduction
//    kernel fuses the x2 / G=0 fix-up, the projections store and the NC ekb scaling (opernlc).
#include "nonlop.cuh"
#include <type_traits>
#include "fourwf.cuh"   // g_kernel_launches
#include "context.cuh"
#include <algorithm>

namespace abi {

#ifndef ABI_EMU
// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
constexpr int kBK = 16;
// tile configurations (measured on B200 with tools/gemm_lab.cu, Si-512 shapes, cuBLAS DGEMM = 35.45 TFLOP/s):
//   TN (opernla): 16 warps x (32x32) = 128x128 CTA tile, 4 stages, 1 CTA/SM            -> 34.97 TFLOP/s
//   NN (opernlb):  8 warps x (32x32) =  64x128 CTA tile, 3 stages, 2 CTAs/SM (the two CTAs of an SM share the DMMA
//                  pipe, which removes the 15.2-wave tail of the 128x128 tiling)         -> 34.18 TFLOP/s
// Narrow band blocks (effective N <= 64 / <= 32) use 64- and 32-column CTA tiles of the same warp tile.
template <int WARPS_M_, int WARPS_N_, int STAGES_, int MINB_> struct GemmCfg {
  static constexpr int WM = 32, WN = 32, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, STAGES = STAGES_, MINB = MINB_;
};
using TnCfg = GemmCfg<4, 4, 4, 1>;      // 128 x 128
using TnCfg64 = GemmCfg<8, 2, 4, 1>;    // 256 x 64
using TnCfg32 = GemmCfg<16, 1, 3, 1>;   // 512 x 32
using NnCfg = GemmCfg<2, 4, 3, 2>;      //  64 x 128
using NnCfg64 = GemmCfg<4, 2, 3, 2>;    // 128 x 64
using NnCfg32 = GemmCfg<8, 1, 3, 2>;    // 256 x 32

// K-major tile [rows][16]: the 16-byte chunk c of row r is stored at chunk c ^ swz(r) -> conflict-free LDS.64 fragments
ABI_DEV int swz(int r) { return ((r & 3) << 1) | ((r >> 2) & 1); }

ABI_DEV void cp_async16(double* smem_dst, const double* gsrc, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(bytes));
}
ABI_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> ABI_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

ABI_DEV void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// v with its sign bit xor-ed by `mask` (0 or 0x80000000): integer pipe, keeps the FP64 pipe for the DMMAs
ABI_DEV double flip_sign(double v, int mask) { return __hiloint2double(__double2hiint(v) ^ mask, __double2loint(v)); }

// ---------------------------------------------------------------------------------------------------------
// TN: part[z][n][m] = sum_{k in split z} A[k + m lda] * Beff[k][n]
//   CPLX: Beff columns are (psi_j, -i psi_j) pairs read from one smem row per band
// NN: C[m + n ldc] = sum_k Aeff[m][k] * B[k + n ldb]   (+ add[m + n ldc])
//   A is M-contiguous (P real view). CPLX: k = 2p+c, Aeff[m][2p] = A[m][p], Aeff[m][2p+1] = (iP)[m][p]
// Both: cp.async multi-stage pipeline into shared memory, fragments double-buffered in registers (the loads of
// k-step kk+1 are in flight while the 16 DMMAs of k-step kk issue), one block barrier per 16-wide k tile.
// ---------------------------------------------------------------------------------------------------------
struct GemmParams {
  int M, N, K;                 // TN: N = effective columns (2*ndat when CPLX); NN: K = number of A columns (nprojs)
  const double* A; long long lda;
  const double* B; long long ldb;   // K-contiguous columns
  double* C; long long ldc;    // NN: output ; TN: partial buffer [nsplit][N][M]
  const double* add;           // NN only: optional, same layout as C
  double* C2;                  // NN only: optional copy of the bare product (gvnlxc when the sum goes to ghc)
  const double* kin; double kin_filter;   // NN only: optional getghc filter, C = 0 where kin[m/2] >= kin_filter
  int nsplit, kchunk, tiles_m, tiles_n;
};

// RAG: variant for launches whose last column tile is ragged (fewer than BN - 32 columns): the warps of one COLUMN block sit on
// different SM sub-partitions (the FP64 pipe is per sub-partition: with the default mapping the warps that still have columns
// would share one), and the warps whose 32-column block lies beyond N skip the fragment loads and DMMAs -- they only take part
// in the copies and barriers -- so a 76-band tail costs 3/4 of a 128-band block instead of all of it.
template <bool TN, bool CPLX, class Cfg, bool RAG = false>
__global__ void __launch_bounds__(Cfg::WARPS_M * Cfg::WARPS_N * 32, Cfg::MINB) k_dgemm(GemmParams p) {
  constexpr int WM = Cfg::WM, WN = Cfg::WN, WARPS_N = Cfg::WARPS_N, STAGES = Cfg::STAGES;
  constexpr int BM = WM * Cfg::WARPS_M, BN = WN * WARPS_N, NT = Cfg::WARPS_M * WARPS_N * 32;
  constexpr int FM = WM / 8, FN = WN / 8;
  constexpr int PITCH = BM + 4;                            // M-major A tile: 32 bytes mod 128 -> 4 k rows = 4 bank quarters
  constexpr int AROWS = (!TN && CPLX) ? kBK / 2 : kBK;     // NN: A columns (p) per stage
  constexpr int BROWS = (TN && CPLX) ? BN / 2 : BN;        // TN complex: one smem row per band = two effective columns
  constexpr int A_STAGE = TN ? BM * kBK : AROWS * PITCH;
  constexpr int B_STAGE = BROWS * kBK;
  extern __shared__ __align__(16) double smem_d[];
  double* As = smem_d;
  double* Bs = smem_d + STAGES * A_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = RAG ? (warp % Cfg::WARPS_M) * WM : (warp / WARPS_N) * WM;
  const int wn = RAG ? (warp / Cfg::WARPS_M) * WN : (warp % WARPS_N) * WN;
  int bid = blockIdx.x;
  const int tn = bid % p.tiles_n; bid /= p.tiles_n;
  const int tm = bid % p.tiles_m; const int z = bid / p.tiles_m;
  const int m0 = tm * BM, n0 = tn * BN;
  // k range in units of the smem k index (TN: rows of A^T; NN real: A columns; NN complex: 2 per A column)
  const int k_total = (!TN && CPLX) ? 2 * p.K : p.K;
  const int k_begin = TN ? z * p.kchunk : 0, k_end = TN ? min(k_total, k_begin + p.kchunk) : k_total;
  const int nkt = (k_end - k_begin + kBK - 1) / kBK;
  const int brow0 = (TN && CPLX) ? n0 / 2 : n0;
  const int nrows_b = (TN && CPLX) ? p.N / 2 : p.N;

  // ---- global -> shared copy plan: every thread owns fixed (row, 16-byte chunk) slots of a stage; per k tile it only
  // advances its global pointers and re-evaluates the k-range predicate (no 64-bit index arithmetic in the main loop)
  constexpr int KROWS_PER_PASS = NT / 8;                    // K-major tiles: rows covered by one pass of the CTA
  static_assert(KROWS_PER_PASS % 8 == 0, "swizzle term must be identical for every pass");
  constexpr int A_IT = TN ? (BM * 8) / NT : (AROWS * (BM / 2)) / NT;
  constexpr int B_IT = (BROWS * 8 + NT - 1) / NT;           // narrow B tiles: only the first BROWS*8 threads copy
  static_assert(A_IT >= 1 && (TN ? (BM * 8) % NT == 0 : (AROWS * (BM / 2)) % NT == 0), "tile/threads mismatch");
  static_assert(B_IT == 1 || (BROWS * 8) % NT == 0, "tile/threads mismatch");
  constexpr int MROWS_PER_PASS = NT / (BM / 2);             // M-major A tile: k rows covered by one pass
  const int kch = (tid & 7) * 2;                            // k offset of this thread's chunk in K-major tiles
  const double* a_ptr[A_IT]; bool a_ok[A_IT]; int a_dst;
  const double* b_ptr[B_IT]; bool b_ok[B_IT]; int b_dst;
  if (TN) {
    const int row = tid >> 3;
    a_dst = row * kBK + (((tid & 7) ^ swz(row)) << 1);
#pragma unroll
    for (int i = 0; i < A_IT; i++) {
      const int gm = m0 + row + i * KROWS_PER_PASS;
      a_ok[i] = gm < p.M;
      a_ptr[i] = p.A + (long long)(a_ok[i] ? gm : 0) * p.lda + k_begin + kch;
    }
  } else {
    const int krow = tid / (BM / 2), ch = tid % (BM / 2);
    a_dst = krow * PITCH + ch * 2;
    const int gm = m0 + ch * 2;
#pragma unroll
    for (int i = 0; i < A_IT; i++) {
      a_ok[i] = gm < p.M;
      a_ptr[i] = p.A + (long long)(krow + i * MROWS_PER_PASS) * p.lda + (a_ok[i] ? gm : 0);
    }
  }
  {
    const int row = tid >> 3;
    b_dst = row * kBK + (((tid & 7) ^ swz(row)) << 1);
#pragma unroll
    for (int i = 0; i < B_IT; i++) {
      const int gn = brow0 + row + i * KROWS_PER_PASS;
      b_ok[i] = gn < nrows_b;
      b_ptr[i] = p.B + (long long)(b_ok[i] ? gn : 0) * p.ldb + k_begin + kch;
    }
  }
  const long long a_step = TN ? kBK : (long long)AROWS * p.lda;
  auto load_stage = [&](int s, int kt) {
    const int k0 = k_begin + kt * kBK;
    double* as = As + s * A_STAGE + a_dst;
    double* bs = Bs + s * B_STAGE + b_dst;
    const bool kok = k0 + kch < k_end;                      // K-major chunks: same k offset for all of this thread's slots
    if (TN) {
#pragma unroll
      for (int i = 0; i < A_IT; i++) { cp_async16(as + i * KROWS_PER_PASS * kBK, a_ptr[i], a_ok[i] && kok); a_ptr[i] += a_step; }
    } else {
      const int p0 = (CPLX ? k0 / 2 : k0) + tid / (BM / 2);
#pragma unroll
      for (int i = 0; i < A_IT; i++) {
        cp_async16(as + i * MROWS_PER_PASS * PITCH, a_ptr[i], a_ok[i] && (p0 + i * MROWS_PER_PASS < p.K));
        a_ptr[i] += a_step;
      }
    }
#pragma unroll
    for (int i = 0; i < B_IT; i++) {
      // narrow B tiles (BROWS*8 < threads): the surplus threads own no slot (a zero-fill copy would land outside the tile)
      if ((BROWS * 8) % NT == 0 || (tid >> 3) < BROWS) cp_async16(bs + i * KROWS_PER_PASS * kBK, b_ptr[i], b_ok[i] && kok);
      b_ptr[i] += kBK;
    }
  };

  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; i++)
#pragma unroll
    for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // Per-lane fragment addresses (doubles inside a stage).  K-major tiles: the k-step only changes the 16-byte chunk
  // index by kk*2, which commutes with the swizzle xor -> address(kk) = address(0) ^ (kk*4); the swizzle term depends on
  // (row & 7) = g only, so one address per k-step serves every fragment row block through an immediate offset.
  int a_k[kBK / 4], b_k[kBK / 4][2];
  int a_flip = 0, b_flip = 0;
  {
    const int row = wm + g;
    int a0;
    if (TN) a0 = row * kBK + ((((t >> 1) ^ swz(row)) << 1) | (t & 1));
    else if (!CPLX) a0 = t * PITCH + row;
    else a0 = (t >> 1) * PITCH + ((t & 1) ? (row ^ 1) : row);             // odd k: (iP)[m] = -+P[m^1]
#pragma unroll
    for (int kk = 0; kk < kBK / 4; kk++) a_k[kk] = TN ? (a0 ^ (kk * 4)) : (a0 + kk * (CPLX ? 2 : 4) * PITCH);
  }
  if (!TN && CPLX) a_flip = ((t & 1) && !(g & 1)) ? (int)0x80000000 : 0;  // (iP) even rows (real parts) = -Im P
#pragma unroll
  for (int jp = 0; jp < 2; jp++) {
    const int col = wn + 8 * jp + g;
    int b0;
    if (TN && CPLX) {
      const int row = col >> 1, kq = (col & 1) ? (t ^ 1) : t;            // odd effective column: -i psi
      b0 = row * kBK + ((((kq >> 1) ^ swz(row)) << 1) | (kq & 1));
    } else {
      b0 = col * kBK + ((((t >> 1) ^ swz(col)) << 1) | (t & 1));
    }
#pragma unroll
    for (int kk = 0; kk < kBK / 4; kk++) b_k[kk][jp] = b0 ^ (kk * 4);
  }
  if (TN && CPLX) b_flip = ((g & 1) && (t & 1)) ? (int)0x80000000 : 0;    // Re(-i psi) = Im psi, Im(-i psi) = -Re psi
  auto ld_frags = [&](const double* as, const double* bs, int kk, double (&a)[FM], double (&b)[FN]) {
#pragma unroll
    for (int i = 0; i < FM; i++) {
      const double v = as[a_k[kk] + i * 8 * (TN ? kBK : 1)];
      a[i] = (!TN && CPLX) ? flip_sign(v, a_flip) : v;
    }
#pragma unroll
    for (int j = 0; j < FN; j++) {
      // complex TN: effective columns 8j+g live in smem row 4j+(g>>1): j and j+2 share the swizzle term
      const double v = (TN && CPLX) ? bs[b_k[kk][j & 1] + (j >> 1) * 8 * kBK] : bs[b_k[kk][0] + j * 8 * kBK];
      b[j] = (TN && CPLX) ? flip_sign(v, b_flip) : v;
    }
  };

  for (int s = 0; s < STAGES - 1; s++) { if (s < nkt) load_stage(s, s); cp_async_commit(); }
  cp_async_wait<STAGES - 2>();
  __syncthreads();
  double a[2][FM], b[2][FN];
  const bool active = !RAG || n0 + wn < p.N;               // warp-uniform
  // RAG: 8-column fragments of this warp that hold columns (warp-uniform): the DMMAs of the others are skipped
  const int nfr = RAG ? min(FN, max(0, (p.N - n0 - wn + 7) >> 3)) : FN;
  if (active) ld_frags(As, Bs, 0, a[0], b[0]);
  for (int kt = 0; kt < nkt; kt++) {
    const double* as = As + (kt % STAGES) * A_STAGE;
    const double* bs = Bs + (kt % STAGES) * B_STAGE;
    { const int nx = kt + STAGES - 1; if (nx < nkt) load_stage(nx % STAGES, nx); cp_async_commit(); }
#pragma unroll
    for (int kk = 0; kk < kBK / 4; kk++) {
      if (kk < kBK / 4 - 1) {
        if (active) ld_frags(as, bs, kk + 1, a[(kk + 1) & 1], b[(kk + 1) & 1]);
      } else {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = (kt + 1) % STAGES;
        if (active) ld_frags(As + nk * A_STAGE, Bs + nk * B_STAGE, 0, a[(kk + 1) & 1], b[(kk + 1) & 1]);
      }
      if (active) {
        // (predicated-off DMMAs still occupy the FP64 pipe -- measured: no gain -- so the fragment counts are separate branches)
        auto mma = [&](auto nf) {
#pragma unroll
          for (int j = 0; j < decltype(nf)::value; j++)
#pragma unroll
            for (int i = 0; i < FM; i++) dmma884(acc[i][j][0], acc[i][j][1], a[kk & 1][i], b[kk & 1][j]);
        };
        if (!RAG || nfr == FN) mma(std::integral_constant<int, FN>());
        else if (nfr * 2 > FN) mma(std::integral_constant<int, (3 * FN) / 4>());
        else if (nfr * 4 > FN) mma(std::integral_constant<int, FN / 2>());
        else mma(std::integral_constant<int, FN / 4>());
      }
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
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (n + h < p.N) {
            const size_t o = (size_t)(n + h) * p.ldc + m;
            double v = acc[i][j][h];
            if (p.C2) p.C2[o] = v;
            if (p.add) v += p.add[o];
            if (p.kin && !(p.kin[m >> 1] < p.kin_filter)) v = 0.0;     // m_getghc.F90:1272-1277
            p.C[o] = v;
          }
        }
      }
    }
  }
}

template <bool TN, bool CPLX, class Cfg> constexpr size_t gemm_smem() {
  constexpr int BM = Cfg::WM * Cfg::WARPS_M, BN = Cfg::WN * Cfg::WARPS_N;
  constexpr int a_stage = TN ? BM * kBK : ((CPLX ? kBK / 2 : kBK) * (BM + 4));
  constexpr int b_stage = ((TN && CPLX) ? BN / 2 : BN) * kBK;
  return sizeof(double) * Cfg::STAGES * (a_stage + b_stage);
}

// ---------------------------------------------------------------------------------------------------------
// reduction of the split-K partials + opernla post-processing + NC opernlc
// ---------------------------------------------------------------------------------------------------------
struct ReduceParams {
  int M, ndat, cplex, nsplit, neff;      // M = nprojs, neff = cplex*ndat
  const double* part;                    // [nsplit][neff][M]
  double scale;                          // 2 for istwf_k>=2 (m_opernla_gemm.F90:681-689), 1 otherwise
  int g0fix;                             // istwf_k==2 && me_g0
  const double* A; long long lda;        // P real view (rows 0,1 = G=0)
  const double* B; long long ldb;        // psi real view
  double* gx; long long ldg;             // internal gx [ndat][ldg] (cplex interleaved)
  double* proj_out;                      // caller's projections(cplex,nprojs,ndat) or null
  // NC opernlc fused: gxfac = ekb(iln,itypat) * gx (m_opernlc_ylm_allwf.F90:318-331)
  double* gxfac; const double* ekb; int dimenl1; const int* proj_typ; const int* proj_iln;
};

__global__ void k_reduce_proj(ReduceParams r) {
  const long long total = (long long)r.M * r.ndat;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % r.M), n = (int)(idx / r.M);
    double v[2] = {0.0, 0.0};
    for (int c = 0; c < r.cplex; c++) {
      const int ne = r.cplex * n + c;
      double s = 0.0;
      for (int z = 0; z < r.nsplit; z++) s += r.part[((size_t)z * r.neff + ne) * r.M + m];
      v[c] = s * r.scale;
    }
    if (r.g0fix) {
      const double a0 = r.A[(long long)m * r.lda], a1 = r.A[(long long)m * r.lda + 1];
      const double b0 = r.B[(long long)n * r.ldb], b1 = r.B[(long long)n * r.ldb + 1];
      v[0] -= a0 * b0 + 2.0 * a1 * b1;
    }
    double e = 1.0;
    if (r.gxfac && r.ekb) e = r.ekb[r.proj_iln[m] + r.dimenl1 * r.proj_typ[m]];
    for (int c = 0; c < r.cplex; c++) {
      const size_t o = (size_t)n * r.ldg + (size_t)r.cplex * m + c;
      r.gx[o] = v[c];
      if (r.proj_out) r.proj_out[((size_t)n * r.M + m) * r.cplex + c] = v[c];
      if (r.gxfac && r.ekb) r.gxfac[o] = e * v[c];
    }
  }
}

__global__ void k_load_proj(const double* __restrict__ proj, double* __restrict__ gx, long long ldg, int M, int ndat, int cplex) {
  const long long total = (long long)M * ndat * cplex;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / ((long long)M * cplex), r = idx % ((long long)M * cplex);
    gx[n * ldg + r] = proj[idx];
  }
}

__global__ void k_nc_scale(const double* __restrict__ gx, double* __restrict__ gxfac, long long ldg, int M, int ndat,
                           int cplex, const double* __restrict__ ekb, int dimenl1, const int* __restrict__ proj_typ,
                           const int* __restrict__ proj_iln) {
  const long long total = (long long)M * ndat;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M); const long long n = idx / M;
    const double e = ekb[proj_iln[m] + dimenl1 * proj_typ[m]];
    for (int c = 0; c < cplex; c++) gxfac[n * ldg + (long long)cplex * m + c] = e * gx[n * ldg + (long long)cplex * m + c];
  }
}

// PAW: per (sorted atom, band): gxfac = D_ij gx (packed symmetric, real), optional -lambda S_ij, gxs = S_ij gx
// (m_opernlc_ylm_allwf.F90:336-447, 1253-1295)
__global__ void k_paw_opernlc(const double* __restrict__ gx, double* __restrict__ gxfac, double* __restrict__ gxs,
                              long long ldg, int cplex, const int* __restrict__ atom_first, const int* __restrict__ atom_typ,
                              const int* __restrict__ atom_enl, const double* __restrict__ enl, const double* __restrict__ sij,
                              int dimenl1, int paw_opt, const double* __restrict__ lambda) {
  const int a = blockIdx.x, n = blockIdx.y;
  const int first = atom_first[a], nlmn = atom_first[a + 1] - first;
  const double* D = enl + (size_t)dimenl1 * atom_enl[a];
  const double* S = sij ? sij + (size_t)dimenl1 * atom_typ[a] : nullptr;
  const double lam = (paw_opt == 2) ? lambda[n] : 0.0;
  const double* x = gx + (size_t)n * ldg + (size_t)cplex * first;
  for (int w = threadIdx.x; w < nlmn * cplex; w += blockDim.x) {
    const int j = w / cplex, c = w - j * cplex;
    double sd = 0.0, ss = 0.0;
    for (int i = 0; i < nlmn; i++) {
      const int hi = max(i, j), lo = min(i, j);
      const int pk = hi * (hi + 1) / 2 + lo;
      const double xv = x[i * cplex + c];
      if (paw_opt == 1 || paw_opt == 2 || paw_opt == 4) sd += (D[pk] - (paw_opt == 2 ? lam * S[pk] : 0.0)) * xv;
      if (paw_opt == 3 || paw_opt == 4) ss += S[pk] * xv;
    }
    const size_t o = (size_t)n * ldg + (size_t)cplex * (first + j) + c;
    if (gxfac && (paw_opt == 1 || paw_opt == 2 || paw_opt == 4)) gxfac[o] = sd;
    if (gxs && (paw_opt == 3 || paw_opt == 4)) gxs[o] = ss;
  }
}

// PAW with complex Hermitian D_ij (cplex_enl = 2) and / or spinor wavefunctions (nspinortot = 2), complex gx (istwf_k = 1):
//   gxfac_s = M^{ss} gx_s + M^{ss'} gx_s'      (m_opernlc_ylm_allwf.F90:453-655 spin-diagonal, :660-737 off-diagonal blocks)
// with the packed upper triangles E_b(i <= j) of the blocks b = [up-up, dn-dn, up-dn, dn-up]:
//   M^{ss}[j,i]  = conj(E_s[i,j]) (i < j),  Re E_s[j,j] (i = j: the stored imaginary part is not used),  E_s[j,i] (i > j)
//   M^{ud}[j,i]  = E_ud[j,i] (i > j),  conj(E_du[i,j]) (i <= j);      M^{du}[j,i] = E_du[j,i] (i > j),  conj(E_ud[i,j]) (i <= j)
// paw_opt = 2 subtracts lambda S_ij from the spin-diagonal blocks; gxs = S_ij gx per spinor component (S real, :1253-1295).
// One block per (sorted atom, band); work items (j, spinor).
__global__ void k_paw_opernlc_cplx(const double* __restrict__ gx, double* __restrict__ gxfac, double* __restrict__ gxs, long long ldg,
                                   const int* __restrict__ atom_first, const int* __restrict__ atom_typ,
                                   const int* __restrict__ atom_enl, const double* __restrict__ enl, const double* __restrict__ sij,
                                   int dimenl1, int sij_dim1, int cplex_enl, int nspinor, long long blk_stride, int paw_opt,
                                   const double* __restrict__ lambda) {
  const int a = blockIdx.x, b = blockIdx.y;
  const int first = atom_first[a], nlmn = atom_first[a + 1] - first;
  const double* S = sij ? sij + (size_t)sij_dim1 * atom_typ[a] : nullptr;
  const double lam = (paw_opt == 2) ? lambda[b] : 0.0;
  const bool want_d = paw_opt == 1 || paw_opt == 2 || paw_opt == 4, want_s = paw_opt == 3 || paw_opt == 4;
  auto elem = [&](int blk, int pk) {      // packed element pk of spin block blk as a complex number
    const double* D = enl + blk_stride * blk + (size_t)dimenl1 * atom_enl[a];
    return cplex_enl == 2 ? make_double2(D[2 * pk], D[2 * pk + 1]) : make_double2(D[pk], 0.0);
  };
  for (int w = threadIdx.x; w < nlmn * nspinor; w += blockDim.x) {
    const int s = w / nlmn, j = w - s * nlmn;
    const double2* xs = reinterpret_cast<const double2*>(gx + (size_t)(b * nspinor + s) * ldg) + first;
    const double2* xo = reinterpret_cast<const double2*>(gx + (size_t)(b * nspinor + (1 - s)) * ldg) + first;   // other spinor
    double2 sd = make_double2(0.0, 0.0), ss = make_double2(0.0, 0.0);
    for (int i = 0; i < nlmn; i++) {
      const int hi = max(i, j), lo = min(i, j);
      const int pk = hi * (hi + 1) / 2 + lo;
      const double2 x = xs[i];
      if (want_d) {
        double2 m = elem(s, pk);
        if (i < j) m.y = -m.y; else if (i == j) m.y = 0.0;
        if (paw_opt == 2) m.x -= lam * S[pk];
        sd.x += m.x * x.x - m.y * x.y; sd.y += m.x * x.y + m.y * x.x;
        if (nspinor == 2) {
          // s = 0 (up): upper triangle from the up-dn block, lower + diagonal from conj(dn-up); s = 1 (dn): the other way round
          double2 mo = (i > j) ? elem(2 + s, pk) : elem(2 + (1 - s), pk);
          if (i <= j) mo.y = -mo.y;
          const double2 y = xo[i];
          sd.x += mo.x * y.x - mo.y * y.y; sd.y += mo.x * y.y + mo.y * y.x;
        }
      }
      if (want_s) { ss.x += S[pk] * x.x; ss.y += S[pk] * x.y; }
    }
    const size_t o = (size_t)(b * nspinor + s) * ldg + 2 * (size_t)(first + j);
    if (gxfac && want_d) { gxfac[o] = sd.x; gxfac[o + 1] = sd.y; }
    if (gxs && want_s) { gxs[o] = ss.x; gxs[o + 1] = ss.y; }
  }
}

// prep_projectors: P = 4 pi / sqrt(ucvol) * ffnl(:,1,ilmn,itypat) * (-i)^l * conj(ph3d(:,ia))
__global__ void k_prep_projectors(double2* __restrict__ P, int npw, int nprojs, const double* __restrict__ ffnl, int dimffnl,
                                  int lmnmax, const double2* __restrict__ ph3d, const int* __restrict__ proj_typ,
                                  const int* __restrict__ proj_lmn, const int* __restrict__ proj_atom,
                                  const int* __restrict__ proj_l, double wt) {
  const int ip = blockIdx.y;
  const int typ = proj_typ[ip], lmn = proj_lmn[ip], ia = proj_atom[ip], il = proj_l[ip] & 3;
  const double* f = ffnl + (size_t)npw * ((size_t)dimffnl * (lmn + (size_t)lmnmax * typ));
  const double2* ph = ph3d + (size_t)npw * ia;
  for (int ig = blockIdx.x * blockDim.x + threadIdx.x; ig < npw; ig += gridDim.x * blockDim.x) {
    const double a = wt * f[ig];
    double re, im;                       // a * (-i)^l
    switch (il) { case 0: re = a; im = 0.0; break; case 1: re = 0.0; im = -a; break;
                  case 2: re = -a; im = 0.0; break; default: re = 0.0; im = a; break; }
    const double2 e = ph[ig];            // times conj(ph3d)
    P[(size_t)ip * npw + ig] = make_double2(re * e.x + im * e.y, im * e.x - re * e.y);
  }
}
// prep_projectors with the structure-factor phases built in place (ph1d3d, src/56_recipspace/m_kg.F90:644-700:
// ph3d(G, ia) = exp(2 pi i (k+G).xred_ia)): one sincospi per (plane wave, atom), reused by the atom's nlmn projectors;
// the npw x natom phase array (1.2 GB at Si-512) never exists and is never uploaded.
__global__ void k_prep_projectors_xred(double2* __restrict__ P, int npw, const double* __restrict__ ffnl, int dimffnl, int lmnmax,
                                       const int* __restrict__ kg, const double* __restrict__ xred, double k1, double k2, double k3,
                                       const int* __restrict__ atom_first, const int* __restrict__ atom_typ,
                                       const int* __restrict__ proj_l, double wt) {
  const int ia = blockIdx.y;
  const int first = atom_first[ia], nlmn = atom_first[ia + 1] - first, typ = atom_typ[ia];
  const double x1 = xred[3 * ia], x2 = xred[3 * ia + 1], x3 = xred[3 * ia + 2];
  for (int ig = blockIdx.x * blockDim.x + threadIdx.x; ig < npw; ig += gridDim.x * blockDim.x) {
    const double arg = (k1 + kg[3 * ig]) * x1 + (k2 + kg[3 * ig + 1]) * x2 + (k3 + kg[3 * ig + 2]) * x3;
    double es, ec;
    sincospi(2.0 * (arg - rint(arg)), &es, &ec);
    for (int i = 0; i < nlmn; i++) {
      const double a = wt * ffnl[(size_t)npw * ((size_t)dimffnl * (i + (size_t)lmnmax * typ)) + ig];
      double re, im;
      switch (proj_l[first + i] & 3) { case 0: re = a; im = 0.0; break; case 1: re = 0.0; im = -a; break;
                                       case 2: re = -a; im = 0.0; break; default: re = 0.0; im = a; break; }
      P[(size_t)(first + i) * npw + ig] = make_double2(re * ec + im * es, im * ec - re * es);
    }
  }
}
// initylmg, optder = 0, one k-point (src/56_recipspace/m_initylmg.F90:94-396; ass_leg_pol of shared/libpaw/src/m_paw_sphharm.F90):
// ylm(ig, l^2+l+1+-m) real spherical harmonics of k+G with the reference's sign conventions
__device__ double ass_leg_pol_dev(int l, int m, double x) {
  if (fabs(x) > 1.0) x = 1.0;
  double polmm = 1.0;
  if (m > 0) {
    const double sqrx = sqrt(fabs((1.0 - x) * (1.0 + x)));
    for (int i = 1; i <= m; i++) polmm *= (1.0 - 2.0 * i) * sqrx;
  }
  if (l == m) return polmm;
  double tmp1 = x * (2.0 * m + 1.0) * polmm;
  if (l == m + 1) return tmp1;
  double pll = 0.0;
  for (int ll = m + 2; ll <= l; ll++) { pll = (x * (2.0 * ll - 1.0) * tmp1 - (ll + m - 1.0) * polmm) / (double)(ll - m); polmm = tmp1; tmp1 = pll; }
  return pll;
}
__global__ void k_initylmg(double* __restrict__ ylm, int npw, int mpsang, const int* __restrict__ kg, double k1, double k2, double k3,
                           const double* __restrict__ gprimd) {
  const double tol = 1e-10, four_pi = 4.0 * 3.14159265358979323846;
  for (int ig = blockIdx.x * blockDim.x + threadIdx.x; ig < npw; ig += gridDim.x * blockDim.x) {
    const double a = k1 + kg[3 * ig], b = k2 + kg[3 * ig + 1], c = k3 + kg[3 * ig + 2];
    const double xx = a * gprimd[0] + b * gprimd[3] + c * gprimd[6];
    const double yy = a * gprimd[1] + b * gprimd[4] + c * gprimd[7];
    const double zz = a * gprimd[2] + b * gprimd[5] + c * gprimd[8];
    const double rr = sqrt(xx * xx + yy * yy + zz * zz);
    ylm[ig] = 1.0 / sqrt(four_pi);
    for (int i = 1; i < mpsang * mpsang; i++) ylm[(size_t)npw * i + ig] = 0.0;
    if (!(rr > tol)) continue;
    double cphi = 1.0, sphi = 0.0;
    const double ctheta = zz / rr, stheta = sqrt(fabs((1.0 - ctheta) * (1.0 + ctheta)));
    if (stheta > tol) { cphi = xx / (rr * stheta); sphi = yy / (rr * stheta); }
    for (int ll = 1; ll < mpsang; ll++) {
      const int l0 = ll * ll + ll;
      double fact = 1.0 / (double)(ll * (ll + 1));
      const double ylmcst = sqrt((double)(2 * ll + 1) / four_pi);
      ylm[(size_t)npw * l0 + ig] = ylmcst * ass_leg_pol_dev(ll, 0, ctheta);
      double onem = 1.0, er = 1.0, ei = 0.0;
      for (int mm = 1; mm <= ll; mm++) {
        onem = -onem;
        const double t = er * cphi - ei * sphi; ei = er * sphi + ei * cphi; er = t;      // (cphi + i sphi)^mm
        const double work1 = ylmcst * sqrt(fact) * onem * ass_leg_pol_dev(ll, mm, ctheta) * sqrt(2.0);
        ylm[(size_t)npw * (l0 + mm) + ig] = work1 * er;
        ylm[(size_t)npw * (l0 - mm) + ig] = work1 * ei;
        if (mm != ll) fact /= (double)((ll + mm + 1) * (ll - mm));
      }
    }
  }
}

// mkffnl, ider = 0, useylm = 1 (src/66_nonlocal/m_mkffnl.F90:238-560): ffnl(ig,1,ilmn,itypat) = ylm(ig, l^2+l+1+m) * f_ln(|k+G|),
// f_ln = splfit of ffspl(:,:,iln,itypat) on the uniform qgrid (shared/common/src/28_numeric_noabirule/m_splines.F90 splfit, ider 0),
// |k+G| = |gprimd (k+G)| (no 2 pi); channels with indlmn(6)/=1 (and pspso=0) or |ekb| <= tol10 (NC) stay zero.
__global__ void k_mkffnl(double* __restrict__ ffnl, int npw, int lmnmax, int ntypat, const int* __restrict__ indlmn,
                         const int* __restrict__ kg, double k1, double k2, double k3, const double* __restrict__ gprimd,
                         const double* __restrict__ ffspl, int mqgrid, int lnmax, double q0, double dq, const double* __restrict__ ylm,
                         const unsigned char* __restrict__ active) {
  const int ilmn = blockIdx.y % lmnmax, ityp = blockIdx.y / lmnmax;
  const int* il = indlmn + 6 * (ilmn + (size_t)lmnmax * ityp);
  double* out = ffnl + (size_t)npw * (ilmn + (size_t)lmnmax * ityp);
  const bool on = active[ilmn + lmnmax * ityp] != 0;
  const int l = il[0], m = il[1], iln = il[4];
  const double* f = ffspl + (size_t)mqgrid * 2 * ((iln - 1) + (size_t)lnmax * ityp);
  const double* y = ylm + (size_t)npw * (l * l + l + m);
  const double qmax = q0 + dq * (mqgrid - 1), de2 = dq * dq / 6.0;
  for (int ig = blockIdx.x * blockDim.x + threadIdx.x; ig < npw; ig += gridDim.x * blockDim.x) {
    if (!on) { out[ig] = 0.0; continue; }
    const double a = k1 + kg[3 * ig], b = k2 + kg[3 * ig + 1], c = k3 + kg[3 * ig + 2];
    const double c1 = a * gprimd[0] + b * gprimd[3] + c * gprimd[6];     // gprimd(1,1:3) in Fortran order gprimd(i,j) = gprimd[i + 3 j]
    const double c2 = a * gprimd[1] + b * gprimd[4] + c * gprimd[7];
    const double c3 = a * gprimd[2] + b * gprimd[5] + c * gprimd[8];
    const double q = sqrt(c1 * c1 + c2 * c2 + c3 * c3);
    double v;
    if (q >= qmax) v = f[mqgrid - 1];
    else if (q <= q0) v = f[0];
    else {
      const int j = (int)((q - q0) / dq);                                 // jspl - 1
      const double d = q - (q0 + dq * j), bb = d / dq, aa = 1.0 - bb;
      const double cc = aa * (aa * aa - 1.0) * de2, dd = bb * (bb * bb - 1.0) * de2;
      v = aa * f[j] + bb * f[j + 1] + cc * f[mqgrid + j] + dd * f[mqgrid + j + 1];
    }
    out[ig] = y[ig] * v;
  }
}
#endif  // !ABI_EMU

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
void Projectors::alloc(int npw_, int nprojs_, int istwf_k_) {
  size_t need = (size_t)2 * npw_ * nprojs_ + 64;
  if (need > cap) {
    if (d_p) CUDA_CHECK(cudaFree(d_p));
    CUDA_CHECK(cudaMalloc(&d_p, sizeof(double) * need));
    cap = need;
  }
  npw = npw_; nprojs = nprojs_; istwf_k = istwf_k_;
  stamp++;                                    // the caller is about to (re)write P: any int8-sliced copy is stale
}
void Projectors::release() { if (d_p) cudaFree(d_p); d_p = nullptr; cap = 0; npw = nprojs = 0; oz.release(); oz_stamp = 0; }

template <typename T> static T* upload(const std::vector<T>& v) {
  T* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(1, v.size())));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return d;
}

void NonlopAtoms::build(int natom_, int ntypat_, int lmnmax_, const int* indlmn_, const int* nattyp_, const int* atindx1_) {
  release();
  natom = natom_; ntypat = ntypat_; lmnmax = lmnmax_;
  indlmn.assign(indlmn_, indlmn_ + (size_t)6 * lmnmax * ntypat);
  nattyp.assign(nattyp_, nattyp_ + ntypat);
  atindx1.assign(atindx1_, atindx1_ + natom);
  nlmn.resize(ntypat);
  std::vector<int> ptyp, plmn, patom, pl, piln, afirst, atyp, aenl;
  int ia = 0;
  for (int t = 0; t < ntypat; t++) {
    int c = 0;
    for (int i = 0; i < lmnmax; i++) if (indlmn[2 + 6 * (i + (size_t)lmnmax * t)] > 0) c++;   // count(indlmn(3,:,itypat)>0)
    nlmn[t] = c;
    for (int a = 0; a < nattyp[t]; a++, ia++) {
      ABI_CHECK(ia < natom, "sum(nattyp) exceeds natom");
      afirst.push_back((int)ptyp.size()); atyp.push_back(t); aenl.push_back(atindx1[ia] - 1);
      for (int i = 0; i < c; i++) {
        ptyp.push_back(t); plmn.push_back(i); patom.push_back(ia);
        pl.push_back(indlmn[0 + 6 * (i + (size_t)lmnmax * t)]);
        piln.push_back(indlmn[4 + 6 * (i + (size_t)lmnmax * t)] - 1);
      }
    }
  }
  ABI_CHECK(ia == natom, "sum(nattyp) differs from natom");
  afirst.push_back((int)ptyp.size());
  nprojs = (int)ptyp.size();                                  // m_gemm_nonlop.F90:400-403
  d_proj_typ = upload(ptyp); d_proj_lmn = upload(plmn); d_proj_atom = upload(patom); d_proj_l = upload(pl);
  d_proj_iln = upload(piln); d_atom_first = upload(afirst); d_atom_typ = upload(atyp); d_atom_enl = upload(aenl);
}
void NonlopAtoms::release() {
  int** ptrs[] = {&d_proj_typ, &d_proj_lmn, &d_proj_atom, &d_proj_l, &d_proj_iln, &d_atom_first, &d_atom_typ, &d_atom_enl};
  for (auto pp : ptrs) { if (*pp) cudaFree(*pp); *pp = nullptr; }
}

void NonlopEnl::load(const double* enl, int d1, int d2, const double* sij, int ntypat, cudaStream_t st, int nb, int sd1) {
  // gemm_nonlop passes enl on every call (m_nonlop.F90:800-808): keep the device buffers while the sizes fit
  if (sd1 <= 0) sd1 = d1;
  const size_t ne = std::max<size_t>(1, (size_t)d1 * d2 * nb), ns = std::max<size_t>(1, (size_t)sd1 * ntypat);
  dimenl1 = d1; dimenl2 = d2; nblk = nb; sij_dim1 = sd1;
  if (ne > enl_cap || d_enl == nullptr) {
    if (d_enl) CUDA_CHECK(cudaFree(d_enl));
    CUDA_CHECK(cudaMalloc(&d_enl, sizeof(double) * ne)); enl_cap = ne;
  }
  CUDA_CHECK(cudaMemcpyAsync(d_enl, enl, sizeof(double) * (size_t)d1 * d2 * nb, cudaMemcpyDefault, st));
  if (sij) {
    if (ns > sij_cap || d_sij == nullptr) {
      if (d_sij) CUDA_CHECK(cudaFree(d_sij));
      CUDA_CHECK(cudaMalloc(&d_sij, sizeof(double) * ns)); sij_cap = ns;
    }
    CUDA_CHECK(cudaMemcpyAsync(d_sij, sij, sizeof(double) * (size_t)sd1 * ntypat, cudaMemcpyDefault, st));
  } else if (d_sij) {
    CUDA_CHECK(cudaFree(d_sij)); d_sij = nullptr; sij_cap = 0;
  }
  CUDA_CHECK(cudaStreamSynchronize(st));     // the caller owns enl / sij and may change them after the call returns
}
void NonlopEnl::release() {
  if (d_enl) cudaFree(d_enl);
  if (d_sij) cudaFree(d_sij);
  d_enl = d_sij = nullptr; enl_cap = sij_cap = 0;
}

struct NlWorkspace {
  double* p = nullptr; size_t cap = 0;
  double* get(size_t n) {
    if (n > cap) { if (p) CUDA_CHECK(cudaFree(p)); CUDA_CHECK(cudaMalloc(&p, sizeof(double) * (n + n / 8 + 64))); cap = n + n / 8 + 64; }
    return p;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
static NlWorkspace g_nlws_all[kMaxLanes][4];   // per lane: 0: partials, 1: gx, 2: gxfac, 3: gxs
#define g_nlws g_nlws_all[ctx().lane]
void nonlop_release_workspace() { for (auto& l : g_nlws_all) for (auto& w : l) w.release(); ozaki_release_workspace(); }

#ifndef ABI_EMU
void prep_projectors_device(Projectors& P, const NonlopAtoms& at, const double* d_ffnl, int dimffnl, const double* d_ph3d,
                            int matblk, double ucvol, cudaStream_t st) {
  ABI_CHECK(matblk >= at.natom, "ph3d must hold one phase column per atom (matblk >= natom)");
  const double wt = 4.0 * 3.14159265358979323846 / sqrt(ucvol);
  if (at.nprojs == 0 || P.npw == 0) return;
  dim3 grid(std::min(64, ceil_div(P.npw, 256)), at.nprojs);
  k_prep_projectors<<<grid, 256, 0, st>>>(reinterpret_cast<double2*>(P.d_p), P.npw, at.nprojs, d_ffnl, dimffnl, at.lmnmax,
                                          reinterpret_cast<const double2*>(d_ph3d), at.d_proj_typ, at.d_proj_lmn,
                                          at.d_proj_atom, at.d_proj_l, wt);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void prep_projectors_xred_device(Projectors& P, const NonlopAtoms& at, const double* d_ffnl, int dimffnl, const int* d_kg,
                                 const double* d_xred, const double* kpt, double ucvol, cudaStream_t st) {
  const double wt = 4.0 * 3.14159265358979323846 / sqrt(ucvol);
  if (at.nprojs == 0 || P.npw == 0) return;
  dim3 grid(std::min(64, ceil_div(P.npw, 256)), at.natom);
  k_prep_projectors_xred<<<grid, 256, 0, st>>>(reinterpret_cast<double2*>(P.d_p), P.npw, d_ffnl, dimffnl, at.lmnmax, d_kg, d_xred,
                                               kpt[0], kpt[1], kpt[2], at.d_atom_first, at.d_atom_typ, at.d_proj_l, wt);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void initylmg_device(double* d_ylm, int npw, int mpsang, const int* d_kg, const double* kpt, const double* d_gprimd, cudaStream_t st) {
#ifndef ABI_EMU
  if (npw == 0) return;
  k_initylmg<<<std::min(kNumSM * 4, ceil_div(npw, 256)), 256, 0, st>>>(d_ylm, npw, mpsang, d_kg, kpt[0], kpt[1], kpt[2], d_gprimd);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
#endif
}

void mkffnl_device(double* d_ffnl, int npw, int lmnmax, int ntypat, const int* d_indlmn, const int* d_kg, const double* kpt,
                   const double* d_gprimd, const double* d_ffspl, int mqgrid, int lnmax, double q0, double dq, const double* d_ylm,
                   const unsigned char* d_active, cudaStream_t st) {
#ifndef ABI_EMU
  if (npw == 0) return;
  k_mkffnl<<<dim3(std::min(64, ceil_div(npw, 256)), lmnmax * ntypat), 256, 0, st>>>(d_ffnl, npw, lmnmax, ntypat, d_indlmn, d_kg, kpt[0], kpt[1],
                                                                                   kpt[2], d_gprimd, d_ffspl, mqgrid, lnmax, q0, dq, d_ylm,
                                                                                   d_active);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
#endif
}

// developer knob (abi_b200_fourwf_set_tuning("nonlop_rag", mask)): ragged variants of bit 0: TN 128-wide, bit 1: TN 64-wide, bit 2: NN 64-wide, bit 3: NN 65..104 columns as 64 + rest, bit 4 / 5: TN / NN 32-wide (TN: no gain, off by default)
int g_nonlop_rag = 47;
void nonlop_set_rag(int mask) { g_nonlop_rag = mask; }

template <bool TN, bool CPLX, class Cfg, bool RAG = false>
static void launch_gemm(const GemmParams& p, int nblocks, cudaStream_t st) {
  auto kern = k_dgemm<TN, CPLX, Cfg, RAG>;
  constexpr size_t smem = gemm_smem<TN, CPLX, Cfg>();
  static bool attr_done = false;
  if (!attr_done) { CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; }
  kern<<<nblocks, Cfg::WARPS_M * Cfg::WARPS_N * 32, smem, st>>>(p);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

// Small-system TN product (Si-2 / Fe-2 per k-point: M = nprojs <= 64, N <= 64 effective columns, K = 2 npw ~ 10^3): one CTA per
// K chunk holds its slices of A and B in shared memory and every thread accumulates a few of the M x N outputs with plain
// FP64 FMAs -- the 128-wide DMMA tiles above would be 90 % padding here and cost the latency of their cp.async pipeline.
// Same partial-buffer layout [z][n][m] and the same reduction kernel as the tiled path.
constexpr int kSmallK = 32;
template <bool CPLX>
__global__ void __launch_bounds__(256) k_small_tn(GemmParams p) {
  __shared__ double As[64][kSmallK + 1];
  __shared__ double Bs[64][kSmallK + 2];
  const int z = blockIdx.x, k0 = z * p.kchunk, kc = min(p.kchunk, p.K - k0);
  const int nrows_b = CPLX ? p.N / 2 : p.N;
  for (int w = threadIdx.x; w < p.M * kSmallK; w += 256) {
    const int m = w / kSmallK, k = w - m * kSmallK;
    As[m][k] = k < kc ? p.A[(size_t)m * p.lda + k0 + k] : 0.0;
  }
  for (int w = threadIdx.x; w < nrows_b * kSmallK; w += 256) {
    const int n = w / kSmallK, k = w - n * kSmallK;
    Bs[n][k] = k < kc ? p.B[(size_t)n * p.ldb + k0 + k] : 0.0;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < p.M * p.N; o += 256) {
    const int m = o % p.M, n = o / p.M;
    double acc = 0.0;
    if (!CPLX || !(n & 1)) {
      const double* b = Bs[CPLX ? n >> 1 : n];
#pragma unroll 8
      for (int k = 0; k < kSmallK; k++) acc = fma(As[m][k], b[k], acc);
    } else {
      // column (-i psi): real view (Im psi, -Re psi) -- k chunks are even, so pairs never straddle a chunk
      const double* b = Bs[n >> 1];
#pragma unroll 8
      for (int k = 0; k < kSmallK; k += 2) { acc = fma(As[m][k], b[k + 1], acc); acc = fma(-As[m][k + 1], b[k], acc); }
    }
    p.C[((size_t)z * p.N + n) * p.M + m] = acc;
  }
}

// split-K TN GEMM into partial buffers; returns nsplit
static int launch_tn(bool cplx, int M, int Neff, int K, const double* A, long long lda, const double* B, long long ldb,
                     double*& part, cudaStream_t st, const char* prof_name = "dgemm_tn_opernla") {
  // (narrow blocks, Neff <= 32, on the ragged 64- or 128-wide tiles instead of the 512 x 32 one: measured 5.0 / 6.2 ms against 5.0-5.2
  //  at 10 columns -- the narrow TN product is bound by its 128-byte-per-row access pattern (4.2 TB/s), not by DMMAs or tile shape)
  const int BN = Neff <= 32 ? 32 : (Neff <= 64 ? 64 : 128), BM = 128 * 128 / BN;
  GemmParams p{};
  p.M = M; p.N = Neff; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.add = nullptr; p.ldc = 0;
  if (M <= 64 && Neff <= 64 && K <= 128 * kSmallK && K % 2 == 0) {
    p.kchunk = kSmallK; p.nsplit = ceil_div(K, kSmallK);
    part = g_nlws[0].get((size_t)p.nsplit * Neff * M);
    p.C = part;
    ProfScope ps(prof_name);
    if (cplx) k_small_tn<true><<<p.nsplit, 256, 0, st>>>(p); else k_small_tn<false><<<p.nsplit, 256, 0, st>>>(p);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
    return p.nsplit;
  }
  p.tiles_m = ceil_div(M, BM); p.tiles_n = ceil_div(Neff, BN);
  const int tiles = p.tiles_m * p.tiles_n;
  // pick the split count that minimises the makespan (waves of 148 CTAs per unit of work)
  // few output tiles (small systems: Si-2, Fe-2 per k-point): the GPU is filled along K instead -- short chunks, or one or
  // two CTAs walk the whole K loop at the latency of one k tile per step (111 us for K = 1480 before, launch-bound after)
  const int min_chunk = (tiles >= kNumSM ? 64 : 8) * kBK;
  const int max_split = std::max(1, std::min(64, ceil_div(K, min_chunk)));
  int nsplit = 1; double best = 1e30;
  for (int s = 1; s <= max_split; s++) {
    const double cost = (double)ceil_div(tiles * s, kNumSM) / s + 0.002 * s;
    if (cost < best - 1e-12) { best = cost; nsplit = s; }
  }
  int kchunk = ceil_div(ceil_div(K, nsplit), kBK) * kBK;
  nsplit = ceil_div(K, kchunk);
  p.nsplit = nsplit; p.kchunk = kchunk;
  part = g_nlws[0].get((size_t)nsplit * Neff * M);
  p.C = part;
  ProfScope ps(prof_name);
  const int nb = tiles * nsplit;
  // ragged single column tile: the variant that spreads column blocks over the SM sub-partitions and skips the DMMAs of empty
  // 8-column fragments (cost follows the columns present, rounded up to 8)
  // (a 32-column tile issues 4.75 ms of DMMAs per pass over the Si-512 P against 3.3 ms of HBM time: with 8 or 16 columns present
  //  the skipped fragments turn it from pipe-bound into HBM-bound)
  const bool ragt = p.tiles_n == 1 && Neff <= BN - 8 && M >= 4 * BM && (g_nonlop_rag & (BN == 128 ? 1 : (BN == 64 ? 2 : 16)));
  if (ragt && BN == 128) { if (cplx) launch_gemm<true, true, TnCfg, true>(p, nb, st); else launch_gemm<true, false, TnCfg, true>(p, nb, st); }
  else if (ragt && BN == 64) { if (cplx) launch_gemm<true, true, TnCfg64, true>(p, nb, st); else launch_gemm<true, false, TnCfg64, true>(p, nb, st); }
  else if (ragt) { if (cplx) launch_gemm<true, true, TnCfg32, true>(p, nb, st); else launch_gemm<true, false, TnCfg32, true>(p, nb, st); }
  else if (BN == 128) { if (cplx) launch_gemm<true, true, TnCfg>(p, nb, st); else launch_gemm<true, false, TnCfg>(p, nb, st); }
  else if (BN == 64) { if (cplx) launch_gemm<true, true, TnCfg64>(p, nb, st); else launch_gemm<true, false, TnCfg64>(p, nb, st); }
  else { if (cplx) launch_gemm<true, true, TnCfg32>(p, nb, st); else launch_gemm<true, false, TnCfg32>(p, nb, st); }
  return nsplit;
}

static void launch_nn(bool cplx, int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C,
                      long long ldc, const double* add, cudaStream_t st, double* C2 = nullptr, const double* kin = nullptr,
                      double kin_filter = 0.0, const char* prof_name = "dgemm_nn_opernlb") {
  const int BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128), BM = 64 * 128 / BN;
  // 65..104 columns on a large M: a full 64-wide pass plus a ragged / narrow pass over the rest costs less than one 128-wide
  // pass whose empty warp columns cannot be spread over the sub-partitions (two passes over A: 9.9 + 5.4 ms against 19.0 at 76 columns)
  // (a 128 x 128 one-CTA-per-SM ragged NN tile was measured too: 15.2 ms at 76 columns, 18.6 at 100 -- slower than the split below)
  if (BN == 128 && N > 64 && N <= 104 && M >= 4 * BM && (g_nonlop_rag & 8)) {
    launch_nn(cplx, M, 64, K, A, lda, B, ldb, C, ldc, add, st, C2, kin, kin_filter, prof_name);
    const long long n0 = 64;
    launch_nn(cplx, M, N - 64, K, A, lda, B + n0 * ldb, ldb, C + n0 * ldc, ldc, add ? add + n0 * ldc : nullptr, st, C2 ? C2 + n0 * ldc : nullptr, kin,
              kin_filter, prof_name);
    return;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.add = add;
  p.C2 = C2; p.kin = kin; p.kin_filter = kin_filter;
  p.tiles_m = ceil_div(M, BM); p.tiles_n = ceil_div(N, BN);
  p.nsplit = 1; p.kchunk = 0;
  ProfScope ps(prof_name);
  const int nb = p.tiles_m * p.tiles_n;
  // (no RAG variant here: the 64 x 128 tile has two warps per column block, which cannot cover the four sub-partitions of an SM --
  //  measured on B200, 76 columns: 20.2 ms against 18.9 ms for the plain kernel; the TN kernel gains 19 %)
  //  the narrower tiles have four / eight warps per column block: there the ragged variant's fragment skipping pays)
  const bool ragn = p.tiles_n == 1 && N <= BN - 8 && BN <= 64 && M >= 4 * BM && (g_nonlop_rag & (BN == 64 ? 4 : 32));
  if (BN == 128) { if (cplx) launch_gemm<false, true, NnCfg>(p, nb, st); else launch_gemm<false, false, NnCfg>(p, nb, st); }
  else if (ragn && BN == 64) { if (cplx) launch_gemm<false, true, NnCfg64, true>(p, nb, st); else launch_gemm<false, false, NnCfg64, true>(p, nb, st); }
  else if (ragn) { if (cplx) launch_gemm<false, true, NnCfg32, true>(p, nb, st); else launch_gemm<false, false, NnCfg32, true>(p, nb, st); }
  else if (BN == 64) { if (cplx) launch_gemm<false, true, NnCfg64>(p, nb, st); else launch_gemm<false, false, NnCfg64>(p, nb, st); }
  else { if (cplx) launch_gemm<false, true, NnCfg32>(p, nb, st); else launch_gemm<false, false, NnCfg32>(p, nb, st); }
}

__global__ void k_reduce_plain(const double* __restrict__ part, double* __restrict__ C, long long ldc, int M, int N, int nsplit, double alpha) {
  const long long total = (long long)M * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M); const long long n = idx / M;
    double s = 0.0;
    for (int z = 0; z < nsplit; z++) s += part[((size_t)z * N + n) * M + m];
    C[n * ldc + m] = alpha * s;
  }
}

void dgemm_tn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
              double alpha, cudaStream_t st) {
  double* part = nullptr;
  const int nsplit = launch_tn(false, M, N, K, A, lda, B, ldb, part, st, "dgemm_tn_gram");
  k_reduce_plain<<<std::min(kNumSM * 8, (int)ceil_div<long long>((long long)M * N, 256)), 256, 0, st>>>(part, C, ldc, M, N, nsplit, alpha);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

// complex partials [z][2n+c][m] -> C(m,n) = alpha * sum_z (part[2n] + i part[2n+1])
__global__ void k_reduce_cplx(const double* __restrict__ part, double2* __restrict__ C, long long ldc, int M, int N, int nsplit, double alpha) {
  const long long total = (long long)M * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M); const long long n = idx / M;
    double sr = 0.0, si = 0.0;
    for (int z = 0; z < nsplit; z++) {
      sr += part[((size_t)z * 2 * N + 2 * n) * M + m];
      si += part[((size_t)z * 2 * N + 2 * n + 1) * M + m];
    }
    C[n * ldc + m] = make_double2(alpha * sr, alpha * si);
  }
}

void zgemm_cn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
              double alpha, cudaStream_t st) {
  double* part = nullptr;
  const int nsplit = launch_tn(true, M, 2 * N, 2 * K, A, 2 * lda, B, 2 * ldb, part, st, "dgemm_tn_gram");
  k_reduce_cplx<<<std::min(kNumSM * 8, (int)ceil_div<long long>((long long)M * N, 256)), 256, 0, st>>>(part, (double2*)C, ldc, M, N, nsplit, alpha);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void dgemm_nn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
              cudaStream_t st) {
  launch_nn(false, M, N, K, A, lda, B, ldb, C, ldc, nullptr, st, nullptr, nullptr, 0.0, "dgemm_nn_rotate");
}

void zgemm_nn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
              cudaStream_t st) {
  launch_nn(true, 2 * M, N, K, A, 2 * lda, B, 2 * ldb, C, 2 * ldc, nullptr, st, nullptr, nullptr, 0.0, "dgemm_nn_rotate");
}

// opernla / opernlb as stand-alone steps (used by apply_invovl): gx is the padded internal layout [ndat][ldg]
long long nonlop_ldg(const Projectors& P) { return ((long long)(P.istwf_k == 1 ? 2 : 1) * P.nprojs + 1) & ~1LL; }

void nonlop_project(const Projectors& P, int me_g0, const double* vectin, int ndat, double* gx, cudaStream_t st) {
  const bool cplx = P.istwf_k == 1; const int cplex = cplx ? 2 : 1;
  const long long ldv = 2LL * P.npw, ldg = nonlop_ldg(P);
  if (ldg != (long long)cplex * P.nprojs) CUDA_CHECK(cudaMemsetAsync(gx, 0, sizeof(double) * ldg * ndat, st));
  double* part = nullptr;
  const int nsplit = launch_tn(cplx, P.nprojs, cplex * ndat, 2 * P.npw, P.d_p, ldv, vectin, ldv, part, st);
  ReduceParams r{};
  r.M = P.nprojs; r.ndat = ndat; r.cplex = cplex; r.nsplit = nsplit; r.neff = cplex * ndat; r.part = part;
  r.scale = cplx ? 1.0 : 2.0; r.g0fix = (P.istwf_k == 2 && me_g0 == 1) ? 1 : 0;
  r.A = P.d_p; r.lda = ldv; r.B = vectin; r.ldb = ldv; r.gx = gx; r.ldg = ldg;
  const int blocks = std::min(kNumSM * 8, (int)ceil_div<long long>((long long)P.nprojs * ndat, 256));
  k_reduce_proj<<<blocks, 256, 0, st>>>(r);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void nonlop_expand(const Projectors& P, const double* z, int ndat, double* vectout, const double* add, cudaStream_t st) {
  launch_nn(P.istwf_k == 1, 2 * P.npw, ndat, P.nprojs, P.d_p, 2LL * P.npw, z, nonlop_ldg(P), vectout, 2LL * P.npw, add, st);
}

// opernld, choice 1 (m_opernld_ylm_allwf.F90:160-203): enlout(idat) = sum_{ilmn, cplex} gxfac * gx
__global__ void k_opernld(const double* __restrict__ gx, const double* __restrict__ gxfac, long long ldg, int n, double* __restrict__ enlout) {
  __shared__ double red[256];
  const int idat = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += gxfac[(size_t)idat * ldg + i] * gx[(size_t)idat * ldg + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) { if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) enlout[idat] = red[0];
}

void gemm_nonlop_device(const Projectors& P, const NonlopAtoms& at, const NonlopEnl& enl, int choice, int cpopt, int paw_opt,
                        int me_g0, const double* d_lambda, int ndat, const double* vectin, double* vectout, double* svectout,
                        double* projections, cudaStream_t st, const NonlopFusion* fuse, int signs, double* enlout) {
  ABI_CHECK(choice == 0 || choice == 1 || choice == 7, "gemm_nonlop: only choice 0, 1, 7 are on the getghc path");
  ABI_CHECK(signs == 2 || (signs == 1 && choice == 1), "gemm_nonlop: signs=1 is implemented for choice=1 only (energy contribution)");
  ABI_CHECK(signs == 2 || enlout != nullptr, "gemm_nonlop: signs=1 needs enlout");
  ABI_CHECK(paw_opt >= 0 && paw_opt <= 4, "gemm_nonlop: bad paw_opt");
  ABI_CHECK(P.nprojs == at.nprojs, "gemm_nonlop: projectors were prepared for a different atom table");
  const int npw = P.npw, nprojs = P.nprojs;
  if (signs == 1 && (nprojs == 0 || ndat == 0)) {             // m_gemm_nonlop.F90:420-423
    if (ndat > 0) CUDA_CHECK(cudaMemsetAsync(enlout, 0, sizeof(double) * ndat, st));
    return;
  }
  if (nprojs == 0 || ndat == 0) {
    if (vectout) CUDA_CHECK(cudaMemsetAsync(vectout, 0, sizeof(double) * 2 * (size_t)npw * ndat, st));
    if (svectout && vectin) CUDA_CHECK(cudaMemcpyAsync(svectout, vectin, sizeof(double) * 2 * (size_t)npw * ndat, cudaMemcpyDeviceToDevice, st));
    return;
  }
  const bool cplx = P.istwf_k == 1;
  const int cplex = cplx ? 2 : 1;
  const long long ldv = 2LL * npw;
  const long long ldg = ((long long)cplex * nprojs + 1) & ~1LL;      // even: 16-byte aligned columns for cp.async
  double* gx = g_nlws[1].get((size_t)ldg * ndat);
  double* gxfac = g_nlws[2].get((size_t)ldg * ndat);
  double* gxs = (paw_opt == 3 || paw_opt == 4) ? g_nlws[3].get((size_t)ldg * ndat) : nullptr;
  if (ldg != (long long)cplex * nprojs) {   // keep the pad element finite (it multiplies zero-filled A rows)
    CUDA_CHECK(cudaMemsetAsync(gx, 0, sizeof(double) * ldg * ndat, st));
    CUDA_CHECK(cudaMemsetAsync(gxfac, 0, sizeof(double) * ldg * ndat, st));
    if (gxs) CUDA_CHECK(cudaMemsetAsync(gxs, 0, sizeof(double) * ldg * ndat, st));
  }
  const bool nc_fused = (choice == 1 && paw_opt == 0);
  const int blocks = std::min(kNumSM * 8, (int)ceil_div<long long>((long long)nprojs * ndat, 256));
  if (cpopt >= 2) {
    // <p|c> already in memory (m_gemm_nonlop.F90:719-734)
    ABI_CHECK(projections != nullptr, "gemm_nonlop: cpopt>=2 needs the projections buffer");
    k_load_proj<<<blocks, 256, 0, st>>>(projections, gx, ldg, nprojs, ndat, cplex);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
    if (nc_fused) {
      k_nc_scale<<<blocks, 256, 0, st>>>(gx, gxfac, ldg, nprojs, ndat, cplex, enl.d_enl, enl.dimenl1, at.d_proj_typ, at.d_proj_iln);
      CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
  } else {
    ABI_CHECK(vectin != nullptr, "gemm_nonlop: vectin is required");
    double* part = nullptr;
    int nsplit;
    bool oz = ozaki_enabled();                     // opt-in int8-sliced contractions (ozaki.cu), off by default
    if (oz && P.oz_stamp != P.stamp) { ozaki_prepare(P, P.oz, st); P.oz_stamp = P.stamp; }
    if (oz && P.oz.failed) oz = false;
    if (oz) {
      part = g_nlws[0].get((size_t)cplex * ndat * nprojs);
      ozaki_project(P.oz, vectin, ndat, part, st);
      nsplit = 1;
    } else {
      nsplit = launch_tn(cplx, nprojs, cplex * ndat, 2 * npw, P.d_p, ldv, vectin, ldv, part, st);
    }
    ReduceParams r;
    r.M = nprojs; r.ndat = ndat; r.cplex = cplex; r.nsplit = nsplit; r.neff = cplex * ndat; r.part = part;
    r.scale = cplx ? 1.0 : 2.0;
    r.g0fix = (P.istwf_k == 2 && me_g0 == 1) ? 1 : 0;
    r.A = P.d_p; r.lda = ldv; r.B = vectin; r.ldb = ldv; r.gx = gx; r.ldg = ldg;
    r.proj_out = (cpopt >= 0 || choice == 0) ? projections : nullptr;
    r.gxfac = nc_fused ? gxfac : nullptr; r.ekb = nc_fused ? enl.d_enl : nullptr; r.dimenl1 = enl.dimenl1;
    r.proj_typ = at.d_proj_typ; r.proj_iln = at.d_proj_iln;
    ProfScope ps("reduce_opernlc");
    k_reduce_proj<<<blocks, 256, 0, st>>>(r);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  if (choice == 0) return;
  const double* zfac = gxfac;
  const double* zs = gxs;
  if (choice == 7) {
    zs = gx;                                   // s_projections = projections (m_gemm_nonlop.F90:856-861)
  } else if (paw_opt != 0) {
    ABI_CHECK(enl.d_enl != nullptr, "gemm_nonlop: D_ij not loaded");
    // k_paw_opernlc indexes real packed-symmetric D_ij / S_ij (m_opernlc_ylm_allwf.F90:395-447): cplex_dij = 2 or a
    // q-dependent layout (dimekbq = 2) would be read as something else
    const int lmn2 = at.lmnmax * (at.lmnmax + 1) / 2;
    ABI_CHECK(enl.dimenl1 == enl.cplex_enl * lmn2, "gemm_nonlop: PAW D_ij must be packed symmetric / Hermitian, dimenl1 = cplex_dij * lmnmax*(lmnmax+1)/2");
    ABI_CHECK(!(paw_opt >= 2) || enl.d_sij != nullptr, "gemm_nonlop: S_ij not loaded");
    if (enl.cplex_enl == 1 && enl.nspinor == 1) {
      k_paw_opernlc<<<dim3(at.natom, ndat), 64, 0, st>>>(gx, gxfac, gxs, ldg, cplex, at.d_atom_first, at.d_atom_typ, at.d_atom_enl,
                                                         enl.d_enl, enl.d_sij, enl.dimenl1, paw_opt, d_lambda);
    } else {
      // complex Hermitian D_ij and / or spinor-mixing blocks (m_opernlc_ylm_allwf.F90:453-737): complex projections only
      ABI_CHECK(cplex == 2, "gemm_nonlop: complex D_ij / spinor blocks need istwf_k = 1 (cplex_fac = cplex = 2)");
      ABI_CHECK(enl.nspinor == 1 || (enl.nblk == 4 && enl.cplex_enl == 2),
                "gemm_nonlop: nspinor = 2 with PAW needs enl(cplex_dij*lmn2, natom, 4) with cplex_dij = 2 (m_opernlc_ylm_allwf.F90:665)");
      ABI_CHECK(ndat % enl.nspinor == 0, "gemm_nonlop: the column count is not a multiple of nspinor");
      k_paw_opernlc_cplx<<<dim3(at.natom, ndat / enl.nspinor), 64, 0, st>>>(
          gx, gxfac, gxs, ldg, at.d_atom_first, at.d_atom_typ, at.d_atom_enl, enl.d_enl, enl.d_sij, enl.dimenl1,
          enl.sij_dim1 > 0 ? enl.sij_dim1 : lmn2, enl.cplex_enl, enl.nspinor, (long long)enl.dimenl1 * enl.dimenl2, paw_opt, d_lambda);
    }
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  if (signs == 1) {                                          // opernld (m_gemm_nonlop.F90:881-890)
    const double* fac = (paw_opt == 3) ? zs : zfac;
    k_opernld<<<ndat, 256, 0, st>>>(gx, fac, ldg, cplex * nprojs, enlout);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
    return;
  }
  // opernlb
  if (choice == 7 || paw_opt == 3 || paw_opt == 4) {
    ABI_CHECK(svectout != nullptr && (vectin != nullptr || choice == 7), "gemm_nonlop: svectout/vectin required for the overlap");
    // + vectin except for choice 7 (m_opernlb_gemm.F90:654-665)
    if (ozaki_enabled() && P.oz_stamp == P.stamp && !P.oz.failed)
      ozaki_expand(P.oz, zs, ldg, ndat, svectout, 0, nullptr, nullptr, 0.0, choice == 7 ? nullptr : vectin, st);
    else if (fuse != nullptr && fuse->kinpw != nullptr) {     // getghc: gsc = 0 where the kinetic filter strikes, in the epilogue
      launch_nn(cplx, 2 * npw, ndat, nprojs, P.d_p, ldv, zs, ldg, svectout, ldv, choice == 7 ? nullptr : vectin, st, nullptr, fuse->kinpw,
                fuse->kin_filter);
      fuse->gsc_filtered = true;
    } else
    launch_nn(cplx, 2 * npw, ndat, nprojs, P.d_p, ldv, zs, ldg, svectout, ldv, choice == 7 ? nullptr : vectin, st);
  }
  if (choice == 1 && (paw_opt == 0 || paw_opt == 1 || paw_opt == 2 || paw_opt == 4)) {
    const bool oz_b = ozaki_enabled() && P.oz_stamp == P.stamp && !P.oz.failed;
    if (fuse == nullptr) {
      ABI_CHECK(vectout != nullptr, "gemm_nonlop: vectout required");
      if (oz_b) ozaki_expand(P.oz, zfac, ldg, ndat, vectout, 0, nullptr, nullptr, 0.0, nullptr, st);
      else launch_nn(cplx, 2 * npw, ndat, nprojs, P.d_p, ldv, zfac, ldg, vectout, ldv, nullptr, st);
    } else if (oz_b) {
      ABI_CHECK(fuse->ghc != nullptr && fuse->kinpw != nullptr, "gemm_nonlop: fusion needs ghc and kinpw");
      ozaki_expand(P.oz, zfac, ldg, ndat, fuse->ghc, 1, vectout, fuse->kinpw, fuse->kin_filter, nullptr, st,
                   std::max(4, fuse->nslabs), fuse->after_slab, fuse->user);
    } else {
      // getghc fusion: ghc += P.gxfac with the kinetic filter in the GEMM epilogue, in row slabs (each complete for
      // every band, so the caller can ship it to the host while the next slab is computed)
      ABI_CHECK(fuse->ghc != nullptr && fuse->kinpw != nullptr, "gemm_nonlop: fusion needs ghc and kinpw");
      const int M = 2 * npw;
      // slabs are whole waves of CTA tiles (2 CTAs/SM x 148 SMs x BM rows) so that cutting the GEMM adds no tail
      const int BN = ndat <= 32 ? 32 : (ndat <= 64 ? 64 : 128), BM = 64 * 128 / BN;
      const long long wave_rows = 2LL * kNumSM * BM;
      const int waves = (int)ceil_div<long long>(M, wave_rows);
      const int nslabs = std::max(1, std::min(fuse->nslabs, waves));
      const int slab = (int)std::min<long long>(M, (long long)ceil_div(waves, nslabs) * wave_rows);
      for (int mb = 0; mb < M; mb += slab) {
        const int mlen = std::min(slab, M - mb);
        launch_nn(cplx, mlen, ndat, nprojs, P.d_p + mb, ldv, zfac, ldg, fuse->ghc + mb, ldv, fuse->ghc + mb, st,
                  vectout ? vectout + mb : nullptr, fuse->kinpw + mb / 2, fuse->kin_filter);
        if (fuse->after_slab) fuse->after_slab(fuse->user, mb / 2, (mb + mlen) / 2);
      }
    }
  }
}
#else
void prep_projectors_device(Projectors&, const NonlopAtoms&, const double*, int, const double*, int, double, cudaStream_t) {}
void prep_projectors_xred_device(Projectors&, const NonlopAtoms&, const double*, int, const int*, const double*, const double*, double, cudaStream_t) {}
void dgemm_nn(int, int, int, const double*, long long, const double*, long long, double*, long long, cudaStream_t) {}
void zgemm_nn(int, int, int, const double*, long long, const double*, long long, double*, long long, cudaStream_t) {}
void zgemm_cn(int, int, int, const double*, long long, const double*, long long, double*, long long, double, cudaStream_t) {}
void gemm_nonlop_device(const Projectors&, const NonlopAtoms&, const NonlopEnl&, int, int, int, int, const double*, int,
                        const double*, double*, double*, double*, cudaStream_t, const NonlopFusion*, int, double*) {}
void dgemm_tn(int, int, int, const double*, long long, const double*, long long, double*, long long, double, cudaStream_t) {}
#endif

}  // namespace abi
