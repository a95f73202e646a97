// C-ABI entry points of the eigensolver side: nonlop dispatcher, xgBlock algebra, Rayleigh-Ritz, ChebFi2.
// Reference semantics (not code): src/66_nonlocal/m_nonlop.F90:336-976, src/45_xgTools/m_xg.F90,
// src/45_xgTools/m_xg_ortho_RR.F90:251-571, src/48_diago/m_chebfi2.F90:466-1210, src/79_seqpar_mpi/m_chebfiwf.F90:110-385.
#include "../../include/abinit_b200.h"
#include "context.cuh"
#include "ham.cuh"
#include "xg.cuh"
#include "comm.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

using namespace abi;

namespace {

// ---- Chebyshev helpers (m_chebfi2.F90:1031-1064, 1084-1106) ----
int cheb_oracle1(double xx, double aa, double bb, double tol, int nmax) {
  const double xred = (xx - (aa + bb) / 2) / (bb - aa) * 2;
  double yy = xred, yim1 = 1.0;
  int nn = nmax;
  if (1.0 / (yy * yy) < tol) {
    nn = 1;
  } else {
    for (int ii = 2; ii <= nmax - 1; ii++) {
      const double temp = yy;
      yy = 2 * xred * yy - yim1;
      yim1 = temp;
      if (1.0 / (yy * yy) < tol) { nn = ii; break; }
    }
  }
  return nn;
}
double cheb_poly1(double xx, int nn, double aa, double bb) {
  const double xred = (xx - (aa + bb) / 2) / (bb - aa) * 2;
  double yy = xred, yim1 = 1.0;
  for (int ii = 2; ii <= nn; ii++) { const double temp = yy; yy = 2 * xred * yy - yim1; yim1 = temp; }
  return yy;
}

struct Buf {
  double* d = nullptr; size_t cap = 0;
  double* get(size_t n) {
    if (n > cap) { if (d) cudaFree(d); CUDA_CHECK(cudaMalloc(&d, sizeof(double) * n)); cap = n; }
    return d;
  }
  void release() { if (d) cudaFree(d); d = nullptr; cap = 0; }
};
Buf g_cheb[6];      // AX, BX, X_next, X_prev, X (work copy), LOBPCG AllBX0
Buf g_small[2];     // device scalars per band
Buf g_par[4];       // band-parallel drivers: transposer pack buffer, BX (row layout), sub-space matrices, scalars

struct AsyncGuard {   // inner calls must not synchronise per block
  bool old; AsyncGuard() : old(ctx().async) { ctx().async = true; }
  ~AsyncGuard() { ctx().async = old; }
};

// getAX_BX bound to getghc (getghc_gsc1, m_chebfiwf.F90:341-385): AX = H X, BX = S X (PAW) in band blocks of `bandpp`,
// followed by xgBlock_zero_im_g0 on AX and BX (m_chebfi2.F90:580-581).  BX == nullptr for norm-conserving (BX = X).
void get_ax_bx(abi_b200_ham_t* h, int space, int me_g0, int npw, int ncols, int bandpp, double* X, double* AX, double* BX) {
  NvtxRange nvtx("GET_AX_BX");                            // NVTX_CHEBFI2_GET_AX_BX / NVTX_LOBPCG2_GET_AX_BX
  int cpopt = -1, prtvol = 0, tim = 0, type_calc = 0, sij_opt = BX ? 1 : 0;
  const size_t col = 2 * (size_t)npw;
  for (int b0 = 0; b0 < ncols; b0 += bandpp) {
    int nd = std::min(bandpp, ncols - b0);
    abi_b200_getghc_(&cpopt, X + col * b0, nullptr, AX + col * b0, BX ? BX + col * b0 : nullptr, &h, nullptr, nullptr, &nd, &prtvol,
                     &sij_opt, &tim, &type_calc);
  }
  xg_zero_im_g0(space, ncols, AX, npw, me_g0, ctx().stream);
  if (BX) xg_zero_im_g0(space, ncols, BX, npw, me_g0, ctx().stream);
}

// chebfi_rayleighRitzQuotients (m_chebfi2.F90:761-810): eig = <X|AX> / <X|BX>, host result
void rr_quotients(int space, int me_g0, int npw, int ncols, const double* X, const double* AX, const double* BX,
                  std::vector<double>& div, double& maxeig, double& mineig) {
  cudaStream_t st = ctx().stream;
  const int w = (space == SPACE_C) ? 2 : 1;
  double* d1 = g_small[0].get((size_t)2 * w * ncols);
  double* d2 = d1 + (size_t)w * ncols;
  xg_colwise_dot(space, npw, ncols, X, npw, AX, npw, d1, me_g0, st);
  xg_colwise_dot(space, npw, ncols, X, npw, BX ? BX : X, npw, d2, me_g0, st);
  std::vector<double> hbuf((size_t)2 * w * ncols);
  CUDA_CHECK(cudaMemcpyAsync(hbuf.data(), d1, sizeof(double) * hbuf.size(), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  div.resize(ncols);
  maxeig = -1.7976931348623157e308; mineig = 1.7976931348623157e308;
  for (int j = 0; j < ncols; j++) {
    double q;
    if (w == 1) {
      q = hbuf[j] / hbuf[ncols + j];
    } else {   // complex division, real part kept (xgBlock_colwiseDivision SPACE_C, max/min on dble())
      const double ar = hbuf[2 * j], ai = hbuf[2 * j + 1], br = hbuf[2 * ncols + 2 * j], bi = hbuf[2 * ncols + 2 * j + 1];
      q = (ar * br + ai * bi) / (br * br + bi * bi);
    }
    div[j] = q; maxeig = std::max(maxeig, q); mineig = std::min(mineig, q);
  }
}

struct ChebOpts {
  double tolerance, ecut; int ndeg_filter, nbdbuf, oracle; double oracle_factor, oracle_min_occ; int bandpp;
};

// chebfi_set_ndeg_from_residu (m_chebfi2.F90:1131-1210), single band group (shift = 0); work = one block of scratch
// band-parallel: `occ` points at the occupations of THIS rank's bands, `shift` = global index of its first band, `ntot` = all bands
// (shift = xmpi_comm_rank(comm_cols) * bandpp, :1161); the caller takes the MAX over the ranks (:1203)
int ndeg_from_residu(const ChebOpts& o, int space, int me_g0, int npw, int ncols, double lm, double lp, const double* occ,
                     const std::vector<double>& div, int ndeg_max, const double* AX, const double* BX, double* work, int shift = 0,
                     int ntot = -1) {
  if (ntot < 0) ntot = ncols;
  cudaStream_t st = ctx().stream;
  double* d_eig = g_small[1].get((size_t)2 * ncols);
  double* d_res = d_eig + ncols;
  CUDA_CHECK(cudaMemcpyAsync(d_eig, div.data(), sizeof(double) * ncols, cudaMemcpyHostToDevice, st));
  xg_colwise_cymax(space, npw, ncols, work, npw, d_eig, BX, npw, AX, npw, st);     // H|psi> - eig S|psi>
  xg_colwise_norm2(space, npw, ncols, work, npw, d_res, me_g0, st);
  std::vector<double> res(ncols);
  CUDA_CHECK(cudaMemcpyAsync(res.data(), d_res, sizeof(double) * ncols, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  int nbdbuf = 0;
  if (o.nbdbuf > 0) nbdbuf = o.nbdbuf;
  int ndeg = 0;
  for (int i = 0; i < ncols; i++) {
    double r = res[i];
    const double occ_i = occ ? occ[i] : 1.0;
    if (o.nbdbuf == -101) r *= occ_i;                                              // xgBlock_apply_diag(residu, occ)
    const bool test1 = r < o.tolerance;
    const bool test2 = (i + 1 + shift) > ntot - nbdbuf;
    const bool test3 = o.nbdbuf == -101 && occ_i < o.oracle_min_occ;
    int nd = 0;
    if (!(test1 || test2 || test3)) {
      const int n_tol = cheb_oracle1(div[i], lm, lp, o.tolerance / r, 1000);
      if (o.oracle == 1) nd = std::min(std::min(ndeg_max, n_tol), o.ndeg_filter);
      else if (o.oracle == 2) nd = std::min(std::min(ndeg_max, n_tol), cheb_oracle1(div[i], lm, lp, o.oracle_factor, 15));
      else ABI_ERROR("Wrong value for chebfi%oracle");
    }
    ndeg = std::max(ndeg, nd);
  }
  return ndeg;
}

// The filter loop + amplification factors (m_chebfi2.F90:634-677, 944-995) on `ncols` bands held by this process.
// X/AX/BX/Xnext/Xprev are device blocks (2, npw, ncols); on return *X_io points to the buffer holding the filtered X.
void cheb_core(abi_b200_ham_t* h, int space, int me_g0, int npw, int ncols, int bandpp, double** X_io, double* AX, double* BX,
               double** Xnext_io, double** Xprev_io, double lm, double lp, int ndeg, const std::vector<double>& div) {
  cudaStream_t st = ctx().stream;
  double *X = *X_io, *Xn = *Xnext_io, *Xp = *Xprev_io;
  const double center = (lp + lm) * 0.5, radius = (lp - lm) * 0.5;
  const double one_over_r = 1 / radius, two_over_r = 2 / radius;
  for (int ideg = 0; ideg < ndeg; ideg++) {
    // X_next = (AX - center X) * (1/r | 2/r) [- X_prev]   (chebfi_computeNextOrderChebfiPolynom, one pass)
    const double* src = AX;
    if (BX) {                                  // PAW: X_next = getBm1X(AX) = S^-1 AX (apply_invovl, m_chebfiwf.F90:390-440)
      apply_invovl_device(h, AX, Xn, nullptr, ncols * h->nspinor, st);   // spinor components are separate columns of npw rows
      src = Xn;
    }
    xg_cheb_next(space, npw, ncols, Xn, npw, src, npw, X, npw, ideg == 0 ? nullptr : Xp, npw, center,
                 ideg == 0 ? one_over_r : two_over_r, st);
    double* t = Xp; Xp = X; X = Xn; Xn = t;                                         // chebfi_swapInnerBuffers
    get_ax_bx(h, space, me_g0, npw, ncols, bandpp, X, AX, BX);
  }
  // chebfi_ampfactor
  std::vector<double> s(ncols);
  for (int j = 0; j < ncols; j++) {
    double amp = cheb_poly1(div[j], ndeg, lm, lp);
    if (std::fabs(amp) < 1e-3) amp = 1e-3;
    s[j] = 1 / amp;
  }
  double* d_s = g_small[1].get((size_t)2 * ncols);
  CUDA_CHECK(cudaMemcpyAsync(d_s, s.data(), sizeof(double) * ncols, cudaMemcpyHostToDevice, st));
  xg_scale_cols(space, npw, ncols, X, npw, d_s, st);
  xg_scale_cols(space, npw, ncols, AX, npw, d_s, st);
  if (BX) xg_scale_cols(space, npw, ncols, BX, npw, d_s, st);
  CUDA_CHECK(cudaStreamSynchronize(st));     // s[] leaves scope
  *X_io = X; *Xnext_io = Xn; *Xprev_io = Xp;
}

int space_of(const abi_b200_ham_t* h) { return h->istwf_k > 1 ? SPACE_CR : SPACE_C; }
int me_g0_of(const abi_b200_ham_t* h) { return h->istwf_k > 1 ? ((h->istwf_k == 2 && h->me_g0 == 1) ? 1 : 0) : -1; }   // m_chebfiwf.F90:198-206

}  // namespace

namespace abi {
void chebfi_release_workspace() {
  for (auto& b : g_cheb) b.release();
  for (auto& b : g_small) b.release();
  for (auto& b : g_par) b.release();
  invovl_release_workspace();
}
}

extern "C" {

void abi_b200_nonlop_(int* choice, int* cpopt, double* cprjin, double* enlout, abi_b200_ham_t** hamk, int* idir, double* lambda,
                      int* ndat, int* nnlout, int* paw_opt, int* signs, double* svectout, int* tim_nonlop, double* vectin,
                      double* vectout) {
  (void)idir; (void)tim_nonlop;
  ensure_init();
  NvtxRange nvtx("NONLOP");                               // NVTX_NONLOP
  Context& c = ctx();
  abi_b200_ham* h = *hamk;
  const int nd = *ndat;
  c.nonlop_counter += nd;                                                          // m_nonlop.F90:389-392
  ABI_CHECK(*signs == 1 || *signs == 2, "nonlop: signs must be 1 or 2");
  ABI_CHECK(h->P.d_p != nullptr || h->atoms.nprojs == 0, "nonlop: projectors not loaded (load_k)");
  ABI_CHECK(*signs == 2 || *nnlout >= 1, "nonlop: nnlout must be >= 1 for signs=1, choice=1");
#ifndef ABI_EMU
  const int cplex = (h->istwf_k == 1) ? 2 : 1;
  const size_t nv = sizeof(double) * 2 * (size_t)h->npw * nd;
  DevArg a_in(0, vectin, nv, true);
  DevArg a_out(1, (*signs == 2) ? vectout : nullptr, nv, false);
  DevArg a_sout(2, (*signs == 2) ? svectout : nullptr, nv, false);
  DevArg a_prj(3, (*cpopt >= 0) ? cprjin : nullptr, sizeof(double) * (size_t)cplex * h->atoms.nprojs * nd, *cpopt >= 2);
  DevArg a_lam(4, lambda, sizeof(double) * nd, true);
  DevArg a_enl(5, (*signs == 1) ? enlout : nullptr, sizeof(double) * nd, false);
  gemm_nonlop_device(h->P, h->atoms, h->enl, *choice, *cpopt, *paw_opt, h->me_g0, a_lam.as<double>(), nd, a_in.as<double>(),
                     a_out.as<double>(), a_sout.as<double>(), a_prj.as<double>(), c.stream, nullptr, *signs, a_enl.as<double>());
  if (*signs == 2) {
    if (*choice == 1 && *paw_opt != 3) a_out.copy_back();
    if (*choice == 7 || (*choice == 1 && (*paw_opt == 3 || *paw_opt == 4))) a_sout.copy_back();
  } else {
    a_enl.copy_back();
  }
  if ((*cpopt >= 0 && *cpopt < 2) || *choice == 0) a_prj.copy_back();
  if (!c.async || a_in.staged || a_out.staged || a_sout.staged || a_prj.staged || a_enl.staged) CUDA_CHECK(cudaStreamSynchronize(c.stream));
#endif
}

#ifndef ABI_EMU
void abi_b200_xg_gram_(int* space, int* rows, int* ncols_a, int* ncols_b, double* a, int* lda, double* b, int* ldb, double* cmat,
                       int* ldc, int* me_g0) {
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(is_device_ptr(a) && is_device_ptr(b) && is_device_ptr(cmat), "xg_gram: device pointers required");
  xg_gram(*space, *rows, *ncols_a, *ncols_b, a, *lda, b, *ldb, cmat, *ldc, *me_g0, c.stream);
  if (!c.async) CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_xg_rotate_(int* space, int* rows, int* k, int* ncols_out, double* x, int* ldx, double* cmat, int* ldc) {
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(is_device_ptr(x) && is_device_ptr(cmat), "xg_rotate: device pointers required");
  xg_rotate(*space, *rows, *k, *ncols_out, x, *ldx, cmat, *ldc, c.stream);
  if (!c.async) CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_xg_gemm_nn_(int* space, int* rows, int* k, int* ncols_out, double* a, int* lda, double* cmat, int* ldc, double* out, int* ldo,
                          int* upper) {
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(is_device_ptr(a) && is_device_ptr(cmat) && is_device_ptr(out), "xg_gemm_nn: device pointers required");
  if (*upper) xg_gemm_nn_upper(*space, *rows, *k, *ncols_out, a, *lda, cmat, *ldc, out, *ldo, c.stream);
  else xg_gemm_nn(*space, *rows, *k, *ncols_out, a, *lda, cmat, *ldc, out, *ldo, c.stream);
  if (!c.async) CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_xg_chol_inverse_(int* space, int* m, double* a, int* lda, int* info) {
  ensure_init();
  ABI_CHECK(is_device_ptr(a), "xg_chol_inverse: device pointer required");
  *info = xg_chol_inverse(*space == SPACE_C ? SPACE_C : SPACE_R, *m, a, *lda, ctx().stream);
}

void abi_b200_xg_hegvd_(int* space, int* n, double* a, int* lda, double* b, int* ldb, double* w, int* info) {
  ensure_init();
  ABI_CHECK(is_device_ptr(a) && is_device_ptr(w) && (b == nullptr || is_device_ptr(b)), "xg_hegvd: device pointers required");
  *info = xg_hegvd(*space == SPACE_C ? SPACE_C : SPACE_R, *n, a, *lda, b, *ldb, w, ctx().stream);
}

void abi_b200_xg_colwise_(int* op, int* space, int* rows, int* ncols, double* a, int* lda, double* b, int* ldb, double* w, int* ldw,
                          double* da, double* out, int* me_g0) {
  ensure_init();
  Context& c = ctx();
  switch (*op) {
    case 0: xg_colwise_dot(*space, *rows, *ncols, a, *lda, b, *ldb, out, *me_g0, c.stream); break;
    case 1: xg_colwise_norm2(*space, *rows, *ncols, a, *lda, out, *me_g0, c.stream); break;
    case 2: xg_colwise_cymax(*space, *rows, *ncols, a, *lda, da, b, *ldb, w, *ldw, c.stream); break;
    case 3: xg_scale_cols(*space, *rows, *ncols, a, *lda, da, c.stream); break;
    case 4: xg_zero_im_g0(*space, *ncols, a, *lda, *me_g0, c.stream); break;
    case 5: xg_add(*space, *rows, *ncols, a, *lda, b, *ldb, c.stream); break;
    case 6: xg_apply_diag(*space, *rows, *ncols, a, *lda, da, c.stream); break;
    default: ABI_ERROR("xg_colwise: op must be 0 (dot), 1 (norm2), 2 (cymax), 3 (scale), 4 (zero_im_g0), 5 (add), 6 (apply_diag)");
  }
  if (!c.async) CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_xg_rayleigh_ritz_(int* space, int* rows, int* blockdim, double* x, int* ldx, double* ax, int* ldax, double* bx,
                                int* ldbx, double* eigenvalues, int* info, int* solve_ax_bx, int* me_g0) {
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(is_device_ptr(x) && is_device_ptr(ax) && (bx == nullptr || is_device_ptr(bx)), "xg_RayleighRitz: device blocks required");
  DevArg a_eig(8, eigenvalues, sizeof(double) * (*blockdim), false);
  *info = xg_rayleigh_ritz(*space, *rows, *blockdim, x, *ldx, ax, *ldax, bx, bx ? *ldbx : *ldx, a_eig.as<double>(), *solve_ax_bx != 0,
                           *me_g0, c.stream);
  a_eig.copy_back();
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_make_invovl_(abi_b200_ham_t** ham) {
  ensure_init();
  make_invovl(*ham, ctx().stream);
}

void abi_b200_apply_invovl_(abi_b200_ham_t** ham, double* cwavef, double* sm1cwavef, double* cwaveprj, int* npw, int* ndat,
                            int* nspinor, int* block_sliced) {
  (void)block_sliced;
  ensure_init();
  NvtxRange nvtx("INVOVL");                               // NVTX_INVOVL, m_invovl.F90:790
  Context& c = ctx();
  abi_b200_ham* h = *ham;
  ABI_CHECK(*nspinor == h->nspinor, "apply_invovl: nspinor differs from the Hamiltonian's (abi_b200_ham_set_nspinor)");
  ABI_CHECK(*npw == h->npw, "apply_invovl: npw differs from the k-point loaded in ham");
  ABI_CHECK(h->usepaw == 1, "apply_invovl: PAW only");
  // S has no spin structure: every spinor component is a column of npw coefficients (ndat*nspinor columns, m_invovl.F90:851-958)
  const int ncol = *ndat * *nspinor;
  const size_t nv = sizeof(double) * 2 * (size_t)h->npw * ncol;
  const int cplex = h->istwf_k == 1 ? 2 : 1;
  DevArg a_c(0, cwavef, nv, true);
  DevArg a_s(1, sm1cwavef, nv, false);
  DevArg a_p(3, cwaveprj, sizeof(double) * (size_t)cplex * h->atoms.nprojs * ncol, false);
  apply_invovl_device(h, a_c.as<double>(), a_s.as<double>(), a_p.as<double>(), ncol, c.stream);
  a_s.copy_back(); a_p.copy_back();
  if (!c.async || a_c.staged || a_s.staged || a_p.staged) CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// ---- ChebFi2, split in the phases between which a band-parallel run communicates (m_chebfi2.F90:596-611, 687-705) ----
void abi_b200_chebfi_rq_(abi_b200_ham_t** gs_hamk, int* ncols, int* bandpp, double* x, double* ax, double* bx, double* div,
                         double* maxeig, double* mineig) {
  ensure_init();
  AsyncGuard g;
  abi_b200_ham* h = *gs_hamk;
  ABI_CHECK(is_device_ptr(x) && is_device_ptr(ax), "chebfi_rq: device blocks required");
  const int space = space_of(h), me_g0 = me_g0_of(h);
  const int rows = h->npw * h->nspinor;                      // blocks hold npw*nspinor rows per band
  ABI_CHECK(!(h->usepaw && bx == nullptr), "chebfi_rq: a PAW Hamiltonian needs the BX block (S X)");
  get_ax_bx(h, space, me_g0, rows, *ncols, *bandpp, x, ax, h->usepaw ? bx : nullptr);
  std::vector<double> d;
  rr_quotients(space, me_g0, rows, *ncols, x, ax, h->usepaw ? bx : nullptr, d, *maxeig, *mineig);
  std::copy(d.begin(), d.end(), div);
}

void abi_b200_chebfi_core_(abi_b200_ham_t** gs_hamk, int* ncols, int* bandpp, double** x, double* ax, double* bx, double** x_next,
                           double** x_prev, double* lambda_minus, double* lambda_plus, int* ndeg_filter, double* div) {
  ensure_init();
  AsyncGuard g;
  abi_b200_ham* h = *gs_hamk;
  std::vector<double> d(div, div + *ncols);
  ABI_CHECK(!(h->usepaw && bx == nullptr), "chebfi_core: a PAW Hamiltonian needs the BX block (S X)");
  cheb_core(h, space_of(h), me_g0_of(h), h->npw * h->nspinor, *ncols, *bandpp, x, ax, h->usepaw ? bx : nullptr, x_next, x_prev, *lambda_minus,
            *lambda_plus, *ndeg_filter, d);
}

// <psi|Vnl|psi> per band through nonlop(choice=1, signs=1, paw_opt=0) (m_chebfiwf.F90:289-316, m_lobpcgwf.F90:229-236); every
// (band, spinor) pair is a column of npw coefficients, the two spinor contributions of a band are summed
static void enl_per_band(abi_b200_ham_t* h, int nband, int nsp, int bandpp, const double* X, double* enl_out, cudaStream_t st) {
  const size_t colb = 2 * (size_t)h->npw * nsp;             // doubles per band
  double* d_enl = g_small[1].get((size_t)2 * nband * nsp);
  for (int b0 = 0; b0 < nband; b0 += bandpp) {
    const int nd = std::min(bandpp, nband - b0);
    gemm_nonlop_device(h->P, h->atoms, h->enl, 1, -1, 0, h->me_g0, nullptr, nd * nsp, X + colb * b0, nullptr, nullptr, nullptr, st,
                       nullptr, 1, d_enl + (size_t)b0 * nsp);
  }
  if (nsp == 1) {
    CUDA_CHECK(cudaMemcpyAsync(enl_out, d_enl, sizeof(double) * nband, cudaMemcpyDeviceToHost, st));
    return;
  }
  std::vector<double> e((size_t)nband * nsp);
  CUDA_CHECK(cudaMemcpyAsync(e.data(), d_enl, sizeof(double) * e.size(), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  for (int b = 0; b < nband; b++) { double a = 0.0; for (int is = 0; is < nsp; is++) a += e[(size_t)b * nsp + is]; enl_out[b] = a; }
}

// lobpcgwf2 (src/79_seqpar_mpi/m_lobpcgwf.F90:100-250) -> lobpcg_run (src/48_diago/m_lobpcg2.F90:340-765), nblock_lobpcg
// blocks of nband / nblock_lobpcg bands (lobpcg_orthoXwrtBlocks against the finished blocks, final Borthonormalize +
// Rayleigh-Ritz over all bands), paral_kgb = 0
__global__ void k_build_pcon(int npw, int nspinor, const double* __restrict__ kinpw, double* __restrict__ pcon, double filter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npw * nspinor) return;
  const double k = kinpw[i % npw];                                                // xgBlock_apply_diag(W, pcond, nspinor)
  if (k > filter) { pcon[i] = 0.0; return; }                                      // m_lobpcgwf.F90:326-331
  const double num = 27 + k * (18 + k * (12 + 8 * k));
  pcon[i] = num / (num + 16 * k * k * k * k);
}

void abi_b200_lobpcgwf2_(double* cg, double* eig, double* occ, double* enl_out, abi_b200_ham_t** gs_hamk, int* nband, int* npw,
                         int* nspinor, int* prtvol, double* resid, double* tolwfr_diago, int* nline, int* nblock_lobpcg, int* nbdbuf,
                         int* bandpp) {
  (void)prtvol;
  ensure_init();
  NvtxRange nvtx("LOBPCG2");                              // NVTX_LOBPCG2, m_lobpcgwf.F90
  Context& c = ctx();
  cudaStream_t st = c.stream;
  abi_b200_ham* h = *gs_hamk;
  const int nb_all = *nband, nsp = *nspinor, nblock = *nblock_lobpcg;
  ABI_CHECK(nsp == h->nspinor, "lobpcgwf2: nspinor differs from the Hamiltonian's (abi_b200_ham_set_nspinor)");
  ABI_CHECK(*npw == h->npw && h->plan != nullptr, "lobpcgwf2: npw differs from the k-point loaded in gs_hamk");
  const int np = *npw * nsp;                                  // rows of the blocks: npw*nspinor (m_lobpcgwf.F90:192)
  ABI_CHECK(nblock >= 1 && nb_all % nblock == 0, "lobpcgwf2: nband must be a multiple of nblock_lobpcg");   // m_lobpcgwf.F90:133
  ABI_CHECK(*nbdbuf >= 0 || (*nbdbuf == -101 && occ != nullptr), "Bad value of nbdbuf");
  const bool paw = h->usepaw == 1;
  const int space = space_of(h), me_g0 = me_g0_of(h);
  AsyncGuard g;
  const int n = nb_all / nblock;                              // blockdim (m_lobpcgwf.F90:133)
  const size_t col = 2 * (size_t)np;                          // doubles per band
  const size_t blk = col * n;
  DevArg a_cg(10, cg, sizeof(double) * col * nb_all, true);
  double* AllX0 = a_cg.as<double>();                          // X0 of all blocks (m_lobpcg2.F90:768-782): updated block by block
  // AX / BX of the finished blocks (lobpcg_transferAX_BX :877-900); norm-conserving: B X0 = X0
  double* AllAX0 = nblock > 1 ? g_cheb[3].get(col * nb_all) : nullptr;
  double* AllBX0 = nblock > 1 ? (paw ? g_cheb[5].get(col * nb_all) : AllX0) : nullptr;
  // [X | W | P], [AX | AW | AP], [BX | BW | BP] (m_lobpcg2.F90:224-268); norm-conserving: BX = X, so B blocks alias the X ones
  double* XWP = g_cheb[4].get(3 * blk);
  double* AXWP = g_cheb[0].get(3 * blk);
  double* BXWP = paw ? g_cheb[1].get(3 * blk) : nullptr;
  double* d_pcon = g_cheb[2].get((size_t)np + 8);
  k_build_pcon<<<ceil_div(np, 256), 256, 0, st>>>(*npw, nsp, h->d_kinpw, d_pcon, 1.7976931348623157e308 * 1.0e-11);
  CUDA_CHECK(cudaGetLastError());
  double* d_eig = g_small[0].get((size_t)4 * n + nb_all);    // 3n eigenvalues + n residuals of a block, nband final eigenvalues
  double* d_res = d_eig + 3 * n;
  double* d_eig_all = d_eig + 4 * n;
  double *X = XWP, *W = XWP + blk, *AX = AXWP, *AW = AXWP + blk;
  // NC: B X = X.  The blocks the reference keeps separately (BX copy of X) are the SAME vectors, rotated identically.
  double* BX = paw ? BXWP : XWP; double* BW = paw ? BXWP + blk : W;
  double* Bblk = paw ? BXWP : XWP;
  auto getax = [&](double* src, double* a_dst, double* b_dst) { get_ax_bx(h, space, me_g0, np, n, *bandpp, src, a_dst, paw ? b_dst : nullptr); };
  const int nband_eff = (*nbdbuf > 0) ? nb_all - *nbdbuf : nb_all;                  // m_lobpcg2.F90:394-398
  std::vector<double> r(n);
  for (int iblock = 0; iblock < nblock; iblock++) {                                 // "big loop over blocks" :456-695
    const int prev = iblock * n;                                                    // bands of the previous blocks
    const double* occ_b = occ ? occ + prev : nullptr;
    CUDA_CHECK(cudaMemcpyAsync(X, AllX0 + col * prev, sizeof(double) * blk, cudaMemcpyDeviceToDevice, st));   // lobpcg_getX0
    if (iblock > 0) xg_ortho_wrt_blocks(space, np, prev, n, X, np, AllX0, np, AllBX0, np, me_g0, st);         // :464-469
    getax(X, AX, BX);
    int info = xg_b_orthonormalize(space, np, n, X, np, paw ? BX : X, np, AX, np, me_g0, st);   // :497
    (void)info;                                              // BX == X (NC) is rotated once: the xg routines skip aliased blocks
    info = xg_rayleigh_ritz(space, np, n, X, np, AX, np, paw ? BX : nullptr, np, d_eig, false, me_g0, st);   // VAR_X, heevd :500
    ABI_CHECK(info == 0, "lobpcg: the sub-space eigenproblem (X) failed");
    bool compute_residu = true;
    double min_res = 0.0, max_res = 0.0;
    auto residuals = [&]() {
      xg_colwise_cymax(space, np, n, W, np, d_eig, BX, np, AX, np, st);             // lobpcg_getResidu :842-854
      xg_colwise_norm2(space, np, n, W, np, d_res, me_g0, st);
      xg_apply_diag(space, np, n, W, np, d_pcon, st);                               // preconditioner :513
      CUDA_CHECK(cudaMemcpyAsync(r.data(), d_res, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (*nbdbuf >= 0) {                                                           // bands of this block below nband_eff :526-538
        const int cnt = std::min(n, std::max(0, nband_eff - prev));
        min_res = max_res = 0.0;
        for (int i = 0; i < cnt; i++) { if (i == 0) min_res = max_res = r[0]; min_res = std::min(min_res, r[i]); max_res = std::max(max_res, r[i]); }
      } else {
        min_res = r[0]; max_res = r[0] * occ_b[0];
        for (int i = 0; i < n; i++) { min_res = std::min(min_res, r[i]); max_res = std::max(max_res, r[i] * occ_b[i]); }
      }
    };
    for (int iline = 1; iline <= *nline; iline++) {
      residuals();
      if (max_res < *tolwfr_diago) { compute_residu = false; break; }
      if (iblock > 0) xg_ortho_wrt_blocks(space, np, prev, n, W, np, AllX0, np, AllBX0, np, me_g0, st);       // :553-555
      getax(W, AW, BW);
      bool use_xw = (iline == 1 || min_res < 1e-27);
      if (!use_xw) {
        const int ierr = xg_b_orthonormalize(space, np, 3 * n, XWP, np, Bblk, np, AXWP, np, me_g0, st);   // :569
        if (ierr != 0) use_xw = true;                                               // "did not work, try on XW" :583
      }
      if (use_xw) {
        xg_b_orthonormalize(space, np, 2 * n, XWP, np, Bblk, np, AXWP, np, me_g0, st);                   // :556
        CUDA_CHECK(cudaMemsetAsync(XWP + 2 * blk, 0, sizeof(double) * blk, st));    // P = AP = BP = 0 :557-559
        CUDA_CHECK(cudaMemsetAsync(AXWP + 2 * blk, 0, sizeof(double) * blk, st));
        if (paw) CUDA_CHECK(cudaMemsetAsync(BXWP + 2 * blk, 0, sizeof(double) * blk, st));
      }
      info = xg_rayleigh_ritz_xwp(space, np, n, use_xw ? 2 : 3, XWP, AXWP, Bblk, np, d_eig, me_g0, st);
      if (info != 0) { fprintf(stderr, "\n--- !WARNING\nmessage: |\n    RayleighRitz (XW/XWP) did not work, but continue anyway.\n...\n"); break; }
    }
    if (compute_residu) residuals();
    CUDA_CHECK(cudaMemcpyAsync(eig + prev, d_eig, sizeof(double) * n, cudaMemcpyDeviceToHost, st));           // :681-684
    std::copy(r.begin(), r.end(), resid + prev);
    CUDA_CHECK(cudaMemcpyAsync(AllX0 + col * prev, X, sizeof(double) * blk, cudaMemcpyDeviceToDevice, st));   // lobpcg_setX0
    if (nblock > 1) {                                                                                         // lobpcg_transferAX_BX
      CUDA_CHECK(cudaMemcpyAsync(AllAX0 + col * prev, AX, sizeof(double) * blk, cudaMemcpyDeviceToDevice, st));
      if (paw) CUDA_CHECK(cudaMemcpyAsync(AllBX0 + col * prev, BX, sizeof(double) * blk, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_CHECK(cudaStreamSynchronize(st));                   // eig / r are read on the host before the next block reuses d_eig
  }
  if (nblock > 1) {                                          // all bands once more (m_lobpcg2.F90:744-751)
    xg_b_orthonormalize(space, np, nb_all, AllX0, np, paw ? AllBX0 : AllX0, np, AllAX0, np, me_g0, st);
    const int info = xg_rayleigh_ritz(space, np, nb_all, AllX0, np, AllAX0, np, paw ? AllBX0 : nullptr, np, d_eig_all, false, me_g0, st);
    ABI_CHECK(info == 0, "lobpcg: the final sub-space eigenproblem over all blocks failed");
    CUDA_CHECK(cudaMemcpyAsync(eig, d_eig_all, sizeof(double) * nb_all, cudaMemcpyDeviceToHost, st));
  }
  a_cg.copy_back();
  if (!paw && enl_out) enl_per_band(h, nb_all, nsp, *bandpp, AllX0, enl_out, st);
  CUDA_CHECK(cudaStreamSynchronize(st));
}

int abi_b200_cheb_oracle1_(double* xx, double* aa, double* bb, double* tol, int* nmax) { return cheb_oracle1(*xx, *aa, *bb, *tol, *nmax); }
double abi_b200_cheb_poly1_(double* xx, int* nn, double* aa, double* bb) { return cheb_poly1(*xx, *nn, *aa, *bb); }

void abi_b200_chebfiwf2_(double* cg, double* eig, double* occ, double* enl_out, abi_b200_ham_t** gs_hamk, int* nband, int* npw,
                         int* nspinor, int* prtvol, double* resid, double* tolwfr_diago, double* ecut, int* nline, int* nbdbuf,
                         int* chebfi_oracle, double* oracle_factor, double* oracle_min_occ, int* bandpp) {
  (void)prtvol;
  ensure_init();
  NvtxRange nvtx("CHEBFI2");                              // NVTX_CHEBFI2, m_chebfiwf.F90
  Context& c = ctx();
  cudaStream_t st = c.stream;
  abi_b200_ham* h = *gs_hamk;
  const int nb = *nband, nsp = *nspinor;
  ABI_CHECK(nsp == h->nspinor, "chebfiwf2: nspinor differs from the Hamiltonian's (abi_b200_ham_set_nspinor)");
  ABI_CHECK(*npw == h->npw, "chebfiwf2: npw differs from the k-point loaded in gs_hamk");
  const int np = *npw * nsp;                                  // rows of the blocks: npw*nspinor (m_chebfiwf.F90:227)
  ABI_CHECK(h->plan != nullptr, "chebfiwf2: load_k has not been called");
  ABI_CHECK(*bandpp >= 1, "chebfiwf2: bandpp must be >= 1");
  ABI_CHECK(!is_device_ptr(eig) && !is_device_ptr(resid) && (occ == nullptr || !is_device_ptr(occ)), "chebfiwf2: eig, resid, occ are host arrays");
  const bool paw = h->usepaw == 1;
  const int space = space_of(h), me_g0 = me_g0_of(h);
  ChebOpts o{*tolwfr_diago, *ecut, *nline, *nbdbuf, *chebfi_oracle, *oracle_factor, *oracle_min_occ, *bandpp};
  AsyncGuard g;
  const size_t blk = 2 * (size_t)np * nb;
  // X0 = cg (m_chebfiwf.F90:252); chebfi%X is a work copy (m_chebfi2.F90:549), AX, BX, X_next, X_prev are the solver's blocks
  DevArg a_cg(10, cg, sizeof(double) * blk, true);
  double* X = g_cheb[4].get(blk);
  CUDA_CHECK(cudaMemcpyAsync(X, a_cg.as<double>(), sizeof(double) * blk, cudaMemcpyDeviceToDevice, st));
  double* AX = g_cheb[0].get(blk);
  double* BX = paw ? g_cheb[1].get(blk) : nullptr;
  double* Xn = g_cheb[2].get(blk);
  double* Xp = g_cheb[3].get(blk);
  // occupancies are used by the oracle only (m_chebfiwf.F90:256-264: halved when nbdbuf=-101, nspinor=1, nsppol=1 -- the
  // caller passes the array it wants compared with oracle_min_occ)
  get_ax_bx(h, space, me_g0, np, nb, o.bandpp, X, AX, BX);                          // m_chebfi2.F90:578-581
  std::vector<double> div; double maxeig, mineig;
  { NvtxRange nq("RAYLRITZ_Q"); rr_quotients(space, me_g0, np, nb, X, AX, BX, div, maxeig, mineig); }   // :613
  const double lambda_minus = maxeig, lambda_plus = o.ecut;                         // :547, :619
  const int ndeg_max = cheb_oracle1(mineig, lambda_minus, lambda_plus, 1e-16, 40);  // :625
  int ndeg = std::min(ndeg_max, o.ndeg_filter);
  if (o.oracle > 0) ndeg = ndeg_from_residu(o, space, me_g0, np, nb, lambda_minus, lambda_plus, occ, div, ndeg_max, AX, BX ? BX : X, Xn);
  { NvtxRange nc("CHEBFI2_CORE"); cheb_core(h, space, me_g0, np, nb, o.bandpp, &X, AX, BX, &Xn, &Xp, lambda_minus, lambda_plus, ndeg, div); }
  // Rayleigh-Ritz (:705) and residuals (:709-716)
  NvtxRange nrr("RAYLRITZ");
  double* d_eig = g_small[0].get((size_t)2 * nb);
  double* d_res = d_eig + nb;
  const int info = xg_rayleigh_ritz(space, np, nb, X, np, AX, np, BX, np, d_eig, true, me_g0, st);
  ABI_CHECK(info == 0, "chebfi: the sub-space eigenproblem failed (hegvd info /= 0)");
  xg_colwise_cymax(space, np, nb, AX, np, d_eig, BX ? BX : X, np, AX, np, st);
  xg_colwise_norm2(space, np, nb, AX, np, d_res, me_g0, st);
  CUDA_CHECK(cudaMemcpyAsync(eig, d_eig, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(resid, d_res, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(a_cg.as<double>(), X, sizeof(double) * blk, cudaMemcpyDeviceToDevice, st));   // xgBlock_copy(X, X0) :720
  a_cg.copy_back();
  if (!paw && enl_out) enl_per_band(h, nb, nsp, o.bandpp, X, enl_out, st);   // m_chebfiwf.F90:289-316
  CUDA_CHECK(cudaStreamSynchronize(st));
}
// ---- band-parallel ChebFi2 inside the library (paral_kgb = 1, npband = ranks of the library communicator) ----
void abi_b200_comm_get_unique_id_(char* id128) { ensure_init(); comm_get_unique_id(id128); }
void abi_b200_comm_init_rank_(const char* id128, int* nranks, int* rank) { comm_init(id128, *nranks, *rank); }
void abi_b200_comm_adopt_(void* nccl_comm, int* nranks, int* rank) { ensure_init(); comm_adopt(nccl_comm, *nranks, *rank); }
void abi_b200_comm_destroy_(void) { comm_destroy(); }

void abi_b200_xg_transpose_(int* to_rows, double* cols, double* lin, int* rows, int* nband) {
  ensure_init();
  cudaStream_t st = ctx().stream;
  ABI_CHECK(is_device_ptr(cols) && is_device_ptr(lin), "xg_transpose: device blocks required");
  long long f, l;
  block_range(*nband, comm_state().nranks, comm_state().rank, &f, &l);
  double* pack = g_par[0].get(2 * (size_t)(*rows) * std::max<long long>(l - f, 1));
  if (*to_rows) transpose_cols_to_rows(cols, lin, pack, *rows, *nband, st);
  else transpose_rows_to_cols(lin, cols, pack, *rows, *nband, st);
  if (!ctx().async) CUDA_CHECK(cudaStreamSynchronize(st));
}

void abi_b200_chebfiwf2_paral_(double* cg, double* eig, double* occ, double* enl_out, double* resid, abi_b200_ham_t** gs_hamk, int* nband,
                               int* ncols_mine, int* npw, int* nspinor, double* tolwfr_diago, double* ecut, int* nline, int* nbdbuf,
                               int* chebfi_oracle, double* oracle_factor, double* oracle_min_occ, int* bandpp) {
  ensure_init();
  NvtxRange nvtx("CHEBFI2");
  Context& c = ctx();
  cudaStream_t st = c.stream;
  abi_b200_ham* h = *gs_hamk;
  const CommState& cm = comm_state();
  const int R = cm.nranks, nb = *nband, nsp = *nspinor;
  ABI_CHECK(nsp == h->nspinor, "chebfiwf2_paral: nspinor differs from the Hamiltonian's (abi_b200_ham_set_nspinor)");
  ABI_CHECK(*npw == h->npw && h->plan != nullptr, "chebfiwf2_paral: npw differs from the k-point loaded in gs_hamk");
  ABI_CHECK(!is_device_ptr(eig) && !is_device_ptr(resid), "chebfiwf2_paral: eig, resid are host arrays");
  const int np = *npw * nsp;
  long long f, l, lo, hi;
  block_range(nb, R, cm.rank, &f, &l);
  block_range(np, R, cm.rank, &lo, &hi);
  const int ncols = (int)(l - f), nrows = (int)(hi - lo);
  ABI_CHECK(*ncols_mine == ncols, "chebfiwf2_paral: cg does not hold this rank's band block (contiguous blocks, larger ones first)");
  const bool paw = h->usepaw == 1;
  const int space = space_of(h), me_g0 = me_g0_of(h);
  AsyncGuard g;
  const size_t blk_c = 2 * (size_t)np * std::max(ncols, 1), blk_r = 2 * (size_t)std::max(nrows, 1) * nb;
  const size_t blk = std::max(blk_c, blk_r);
  DevArg a_cg(10, cg, sizeof(double) * 2 * (size_t)np * ncols, true);
  double* X = g_cheb[4].get(blk);
  double* AX = g_cheb[0].get(blk);
  double* BX = paw ? g_cheb[1].get(blk) : nullptr;
  double* Xn = g_cheb[2].get(blk);
  double* Xp = g_cheb[3].get(blk);
  CUDA_CHECK(cudaMemcpyAsync(X, a_cg.as<double>(), sizeof(double) * 2 * (size_t)np * ncols, cudaMemcpyDeviceToDevice, st));
  // ---- filter on my band block: no communication except the extrema of the Rayleigh quotients (m_chebfi2.F90:606-611)
  get_ax_bx(h, space, me_g0, np, ncols, *bandpp, X, AX, BX);
  std::vector<double> div; double maxeig, mineig;
  rr_quotients(space, me_g0, np, ncols, X, AX, BX, div, maxeig, mineig);
  if (R > 1) {
    double* d_mm = g_par[3].get(8);
    double mm[2] = {maxeig, -mineig};
    CUDA_CHECK(cudaMemcpyAsync(d_mm, mm, sizeof(mm), cudaMemcpyHostToDevice, st));
    comm_allreduce(d_mm, 2, true, st);
    CUDA_CHECK(cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    maxeig = mm[0]; mineig = -mm[1];
  }
  const double lambda_minus = maxeig, lambda_plus = *ecut;
  const int ndeg_max = cheb_oracle1(mineig, lambda_minus, lambda_plus, 1e-16, 40);
  int ndeg = std::min(ndeg_max, *nline);
  if (*chebfi_oracle > 0) {                                   // chebfi_set_ndeg_from_residu on my bands, MAX over the ranks (:1203)
    ABI_CHECK(*nbdbuf >= 0 || (*nbdbuf == -101 && occ != nullptr), "Bad value of nbdbuf");
    ChebOpts o{*tolwfr_diago, *ecut, *nline, *nbdbuf, *chebfi_oracle, *oracle_factor, *oracle_min_occ, *bandpp};
    ndeg = ndeg_from_residu(o, space, me_g0, np, ncols, lambda_minus, lambda_plus, occ ? occ + f : nullptr, div, ndeg_max, AX, BX ? BX : X, Xn,
                            (int)f, nb);
    if (R > 1) {
      double* d_nd = g_par[3].get(8);
      double nd = ndeg;
      CUDA_CHECK(cudaMemcpyAsync(d_nd, &nd, sizeof(double), cudaMemcpyHostToDevice, st));
      comm_allreduce(d_nd, 1, true, st);
      CUDA_CHECK(cudaMemcpyAsync(&nd, d_nd, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      ndeg = (int)nd;
    }
  }
  { NvtxRange nc("CHEBFI2_CORE"); cheb_core(h, space, me_g0, np, ncols, *bandpp, &X, AX, BX, &Xn, &Xp, lambda_minus, lambda_plus, ndeg, div); }
  // ---- Rayleigh-Ritz in the row-sharded layout (m_chebfi2.F90:687-705): all bands, my rows
  NvtxRange nrr("RAYLRITZ");
  ProfScope* ps_rr = new ProfScope("rr_paral");
  double *Xr = X, *AXr = AX, *BXr = BX;
  double* pack = nullptr;
  if (R > 1) {
    pack = g_par[0].get(blk_c);
    Xr = Xn; AXr = Xp; BXr = paw ? g_par[1].get(blk_r) : nullptr;    // the filter's X_next / X_prev blocks are free now
    transpose_cols_to_rows(X, Xr, pack, np, nb, st);
    transpose_cols_to_rows(AX, AXr, pack, np, nb, st);
    if (paw) transpose_cols_to_rows(BX, BXr, pack, np, nb, st);
  }
  const int me_g0_rows = space == SPACE_CR ? ((h->istwf_k == 2 && cm.rank == 0 && h->me_g0 == 1) ? 1 : 0) : -1;   // row 0 lives on rank 0
  xg_zero_im_g0(space, nb, Xr, nrows, me_g0_rows, st);
  xg_zero_im_g0(space, nb, AXr, nrows, me_g0_rows, st);
  if (paw) xg_zero_im_g0(space, nb, BXr, nrows, me_g0_rows, st);
  const int sc = sub_cplex(space);
  const long long ldw = (nb + 1) & ~1LL;
  const size_t nsub = (size_t)sc * ldw * nb;
  double* subA = g_par[2].get(2 * nsub + nb);
  double* subB = subA + nsub;
  double* d_eig = subB + nsub;
  CUDA_CHECK(cudaMemsetAsync(subA, 0, sizeof(double) * 2 * nsub, st));
  static cudaStream_t cs = nullptr; static cudaEvent_t evA = nullptr, evB = nullptr, evC = nullptr;
  if (!cs) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&evA, cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&evB, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&evC, cudaEventDisableTiming));
  }
  // X^H A X, then its allreduce on the side stream while X^H B X is computed (xgBlock_gemm(..., comm=), m_xg.F90:1969-1974)
  xg_gram(space, nrows, nb, nb, Xr, nrows, AXr, nrows, subA, ldw, me_g0_rows, st);
  if (R > 1) { CUDA_CHECK(cudaEventRecord(evA, st)); CUDA_CHECK(cudaStreamWaitEvent(cs, evA, 0)); comm_allreduce(subA, nsub, false, cs); }
  xg_gram(space, nrows, nb, nb, Xr, nrows, paw ? BXr : Xr, nrows, subB, ldw, me_g0_rows, st);
  if (R > 1) {
    CUDA_CHECK(cudaEventRecord(evB, st)); CUDA_CHECK(cudaStreamWaitEvent(cs, evB, 0)); comm_allreduce(subB, nsub, false, cs);
    CUDA_CHECK(cudaEventRecord(evC, cs)); CUDA_CHECK(cudaStreamWaitEvent(st, evC, 0));
  }
  const int info = xg_hegvd(space == SPACE_C ? SPACE_C : SPACE_R, nb, subA, ldw, subB, ldw, d_eig, st);   // replicated on every rank
  ABI_CHECK(info == 0, "chebfi: the sub-space eigenproblem failed (hegvd info /= 0)");
  xg_rotate(space, nrows, nb, nb, Xr, nrows, subA, ldw, st);
  xg_rotate(space, nrows, nb, nb, AXr, nrows, subA, ldw, st);
  if (paw) xg_rotate(space, nrows, nb, nb, BXr, nrows, subA, ldw, st);
  if (R > 1) {
    transpose_rows_to_cols(Xr, X, pack, np, nb, st);
    transpose_rows_to_cols(AXr, AX, pack, np, nb, st);
    if (paw) transpose_rows_to_cols(BXr, BX, pack, np, nb, st);
  }
  delete ps_rr;
  // ---- residuals of my bands (m_chebfi2.F90:709-716)
  double* d_res = g_small[0].get((size_t)2 * std::max(ncols, 1));
  xg_colwise_cymax(space, np, ncols, AX, np, d_eig + f, BX ? BX : X, np, AX, np, st);
  xg_colwise_norm2(space, np, ncols, AX, np, d_res, me_g0, st);
  CUDA_CHECK(cudaMemcpyAsync(eig, d_eig, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
  if (ncols) CUDA_CHECK(cudaMemcpyAsync(resid, d_res, sizeof(double) * ncols, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(a_cg.as<double>(), X, sizeof(double) * 2 * (size_t)np * ncols, cudaMemcpyDeviceToDevice, st));
  a_cg.copy_back();
  if (!paw && enl_out && ncols) enl_per_band(h, ncols, nsp, *bandpp, X, enl_out, st);   // <psi|Vnl|psi> of my bands (m_chebfiwf.F90:289-316)
  CUDA_CHECK(cudaStreamSynchronize(st));
}

// build_pcon on the rows [lo, lo + nr) of a block with npw*nspinor rows (row-sharded layout)
__global__ void k_build_pcon_rows(int npw, long long lo, int nr, const double* __restrict__ kinpw, double* __restrict__ pcon, double filter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr) return;
  const double k = kinpw[(lo + i) % npw];
  if (k > filter) { pcon[i] = 0.0; return; }
  const double num = 27 + k * (18 + k * (12 + 8 * k));
  pcon[i] = num / (num + 16 * k * k * k * k);
}

// lobpcg_run with paral_kgb = 1, npband = ranks of the library communicator, one block of all bands
// (src/48_diago/m_lobpcg2.F90:340-765): getAX_BX on the rank's own band block (band-sharded layout), everything else --
// B-orthonormalisation, X / XW / XWP Rayleigh-Ritz, residuals, preconditioner -- on the rank's plane-wave rows with the Gram matrices
// summed over the ranks and the small dense problems solved redundantly; per iteration one xgTransposer exchange out (W) and one
// (PAW: two) back (AW, BW).  The scheme of abinit_b200/parallel.py:lobpcg_band_parallel with NCCL on the library stream.
void abi_b200_lobpcgwf2_paral_(double* cg, double* eig, double* resid, abi_b200_ham_t** gs_hamk, int* nband, int* ncols_mine, int* npw,
                               int* nspinor, double* tolwfr_diago, int* nline, int* bandpp) {
  ensure_init();
  NvtxRange nvtx("LOBPCG2");
  Context& c = ctx();
  cudaStream_t st = c.stream;
  abi_b200_ham* h = *gs_hamk;
  const CommState& cm = comm_state();
  const int R = cm.nranks, n = *nband, nsp = *nspinor;
  ABI_CHECK(nsp == h->nspinor, "lobpcgwf2_paral: nspinor differs from the Hamiltonian's (abi_b200_ham_set_nspinor)");
  ABI_CHECK(*npw == h->npw && h->plan != nullptr, "lobpcgwf2_paral: npw differs from the k-point loaded in gs_hamk");
  ABI_CHECK(!is_device_ptr(eig) && !is_device_ptr(resid), "lobpcgwf2_paral: eig, resid are host arrays");
  const int np = *npw * nsp;
  long long f, l, lo, hi;
  block_range(n, R, cm.rank, &f, &l);
  block_range(np, R, cm.rank, &lo, &hi);
  const int ncols = (int)(l - f), nr = (int)(hi - lo);
  ABI_CHECK(*ncols_mine == ncols, "lobpcgwf2_paral: cg does not hold this rank's band block (contiguous blocks, larger ones first)");
  const bool paw = h->usepaw == 1;
  const int space = space_of(h), me_g0_cols = me_g0_of(h);
  const int me_g0 = space == SPACE_CR ? ((h->istwf_k == 2 && cm.rank == 0 && h->me_g0 == 1) ? 1 : 0) : -1;   // row 0 lives on rank 0
  const int sub_space = space == SPACE_C ? SPACE_C : SPACE_R, sc = sub_cplex(space);
  AsyncGuard g;
  const size_t colr = 2 * (size_t)std::max(nr, 1);            // doubles per column in the row layout
  const size_t blk_r = colr * n, blk_c = 2 * (size_t)np * std::max(ncols, 1);
  DevArg a_cg(10, cg, sizeof(double) * 2 * (size_t)np * ncols, true);
  double* XWP = g_cheb[4].get(3 * blk_r);
  double* AXWP = g_cheb[0].get(3 * blk_r);
  double* BXWP = paw ? g_cheb[1].get(3 * blk_r) : XWP;
  CUDA_CHECK(cudaMemsetAsync(XWP, 0, sizeof(double) * 3 * blk_r, st));
  CUDA_CHECK(cudaMemsetAsync(AXWP, 0, sizeof(double) * 3 * blk_r, st));
  if (paw) CUDA_CHECK(cudaMemsetAsync(BXWP, 0, sizeof(double) * 3 * blk_r, st));
  double* blocks[3] = {XWP, AXWP, BXWP};
  const int nblocks = paw ? 3 : 2;
  double* cols_in = g_cheb[2].get(blk_c);                     // band-sharded work blocks of getAX_BX
  double* cols_a = g_cheb[3].get(blk_c);
  double* cols_b = paw ? g_cheb[5].get(blk_c) : nullptr;
  double* pack = g_par[0].get(blk_c);
  const long long ldw3 = (3LL * n + 1) & ~1LL;
  double* sub = g_par[2].get((size_t)2 * sc * ldw3 * 3 * n + (size_t)sc * ldw3 * n);   // A / B sub-space matrices + the c1 rotation block
  double* d_pcon = g_par[1].get((size_t)nr + 8);
  double* d_eig = g_small[0].get((size_t)4 * n);              // 3n eigenvalues + n residuals
  double* d_res = d_eig + 3 * n;
  if (nr) { k_build_pcon_rows<<<ceil_div(nr, 256), 256, 0, st>>>(*npw, lo, nr, h->d_kinpw, d_pcon, 1.7976931348623157e308 * 1.0e-11); CUDA_CHECK(cudaGetLastError()); }
  double *X = XWP, *W = XWP + blk_r, *AX = AXWP, *AW = AXWP + blk_r, *BX = BXWP, *BW = BXWP + blk_r;

  auto to_rows = [&](const double* cols, double* rows) {
    if (R > 1) transpose_cols_to_rows(cols, rows, pack, np, n, st);
    else CUDA_CHECK(cudaMemcpyAsync(rows, cols, sizeof(double) * blk_r, cudaMemcpyDeviceToDevice, st));
  };
  auto to_cols = [&](const double* rows, double* cols) {
    if (R > 1) transpose_rows_to_cols(rows, cols, pack, np, n, st);
    else CUDA_CHECK(cudaMemcpyAsync(cols, rows, sizeof(double) * blk_r, cudaMemcpyDeviceToDevice, st));
  };
  auto apply_h = [&](const double* src_rows, double* dst_rows, double* dst_b_rows) {
    to_cols(src_rows, cols_in);
    get_ax_bx(h, space, me_g0_cols, np, ncols, *bandpp, cols_in, cols_a, cols_b);
    to_rows(cols_a, dst_rows);
    if (paw) to_rows(cols_b, dst_b_rows);
  };
  auto zero_all = [&](int m) { for (int b = 0; b < nblocks; b++) xg_zero_im_g0(space, m, blocks[b], nr, me_g0, st); };
  auto b_orthonormalize = [&](int m) -> int {
    const long long ldw = (m + 1) & ~1LL;
    CUDA_CHECK(cudaMemsetAsync(sub, 0, sizeof(double) * sc * ldw * m, st));
    zero_all(m);
    xg_gram(space, nr, m, m, XWP, nr, BXWP, nr, sub, ldw, me_g0, st);
    comm_allreduce(sub, (size_t)sc * ldw * m, false, st);
    const int info = xg_chol_inverse(sub_space, m, sub, ldw, st);
    if (info != 0) return info;
    for (int b = 0; b < nblocks; b++) xg_gemm_nn_upper(space, nr, m, m, blocks[b], nr, sub, ldw, blocks[b], nr, st);
    return 0;
  };
  auto rayleigh_ritz = [&](int nvar) {
    const int m = nvar * n;
    const long long ldw = (m + 1) & ~1LL;
    const size_t nsub = (size_t)sc * ldw * m;
    double* subA = sub; double* subB = sub + nsub;
    CUDA_CHECK(cudaMemsetAsync(sub, 0, sizeof(double) * 2 * nsub, st));
    zero_all(m);
    for (int v = 0; v < nvar; v++) {                          // upper block columns of [X W P]^H A [X W P] (and B)
      xg_gram(space, nr, (v + 1) * n, n, XWP, nr, AXWP + (size_t)v * blk_r, nr, subA + (size_t)sc * ldw * v * n, ldw, me_g0, st);
      if (nvar > 1) xg_gram(space, nr, (v + 1) * n, n, XWP, nr, BXWP + (size_t)v * blk_r, nr, subB + (size_t)sc * ldw * v * n, ldw, me_g0, st);
    }
    comm_allreduce(sub, (nvar > 1 ? 2 : 1) * nsub, false, st);
    const int info = xg_hegvd(sub_space, m, subA, ldw, nvar > 1 ? subB : nullptr, ldw, d_eig, st);
    ABI_CHECK(info == 0, "lobpcg: the sub-space eigenproblem failed");
    double* c1 = nullptr; long long ldc1 = 0;
    if (nvar > 1) {                                           // rows n..m of the first n eigenvectors -> K-padded (m - n) x n block
      ldc1 = (m - n + 1) & ~1LL;
      c1 = sub + 2 * nsub;
      CUDA_CHECK(cudaMemsetAsync(c1, 0, sizeof(double) * sc * ldc1 * n, st));
      CUDA_CHECK(cudaMemcpy2DAsync(c1, sizeof(double) * sc * ldc1, subA + (size_t)sc * n, sizeof(double) * sc * ldw, sizeof(double) * sc * (m - n), n,
                                   cudaMemcpyDeviceToDevice, st));
    }
    for (int b = 0; b < nblocks; b++) {
      double* blk = blocks[b];
      xg_rotate(space, nr, n, n, blk, nr, subA, ldw, st);                                             // X <- X C(0:n)
      if (nvar > 1) {
        xg_gemm_nn(space, nr, m - n, n, blk + blk_r, nr, c1, ldc1, blk + 2 * blk_r, nr, st);          // P <- [W P] C(n:m)
        xg_add(space, nr, n, blk, nr, blk + 2 * blk_r, nr, st);                                        // X += P
      }
    }
  };
  std::vector<double> r(n);
  double min_res = 0.0, max_res = 0.0;
  auto residuals = [&]() {
    xg_colwise_cymax(space, nr, n, W, nr, d_eig, BX, nr, AX, nr, st);           // W = AX - eig BX (BX = X when norm-conserving)
    xg_colwise_norm2(space, nr, n, W, nr, d_res, me_g0, st);
    comm_allreduce(d_res, n, false, st);
    xg_apply_diag(space, nr, n, W, nr, d_pcon, st);
    CUDA_CHECK(cudaMemcpyAsync(r.data(), d_res, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    min_res = max_res = r[0];
    for (int i = 0; i < n; i++) { min_res = std::min(min_res, r[i]); max_res = std::max(max_res, r[i]); }
  };
  to_rows(a_cg.as<double>(), X);
  apply_h(X, AX, BX);
  b_orthonormalize(n);
  rayleigh_ritz(1);
  bool compute_residu = true;
  for (int iline = 1; iline <= *nline; iline++) {
    residuals();
    if (max_res < *tolwfr_diago) { compute_residu = false; break; }
    apply_h(W, AW, BW);
    bool use_xw = (iline == 1 || min_res < 1e-27);
    if (!use_xw && b_orthonormalize(3 * n) != 0) use_xw = true;
    if (use_xw) {
      b_orthonormalize(2 * n);
      for (int b = 0; b < nblocks; b++) CUDA_CHECK(cudaMemsetAsync(blocks[b] + 2 * blk_r, 0, sizeof(double) * blk_r, st));
    }
    rayleigh_ritz(use_xw ? 2 : 3);
  }
  if (compute_residu) residuals();
  CUDA_CHECK(cudaMemcpyAsync(eig, d_eig, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  std::copy(r.begin(), r.end(), resid);
  to_cols(X, cols_in);
  CUDA_CHECK(cudaMemcpyAsync(a_cg.as<double>(), cols_in, sizeof(double) * 2 * (size_t)np * ncols, cudaMemcpyDeviceToDevice, st));
  a_cg.copy_back();
  CUDA_CHECK(cudaStreamSynchronize(st));
}
#endif   // ABI_EMU



}  // extern "C"
