// C-ABI entry points: gemm_nonlop life cycle + apply, Hamiltonian handle, fused getghc, Gram matrices.
// See include/abinit_b200.h for the reference interfaces replaced.
#include "../../include/abinit_b200.h"
#include <tuple>
#include <map>
#include "context.cuh"
#include "fourwf.cuh"
#include "nonlop.cuh"
#include "ham.cuh"
#include <cmath>
#include <map>
#include <memory>

using namespace abi;

namespace abi {
void nonlop_release_workspace();

struct NonlopSlot { Projectors P; };
static std::map<int, std::unique_ptr<NonlopSlot>> g_slots;
static int g_cur_slot = 1;
static NonlopAtoms g_call_atoms; static uint64_t g_call_atoms_key = 0;
static NonlopEnl g_call_enl;

static uint64_t hash_ints(const int* p, size_t n, uint64_t h = 1469598103934665603ULL) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n * sizeof(int); i++) { h ^= b[i]; h *= 1099511628211ULL; }
  return h;
}
static const NonlopAtoms& call_atoms(int natom, int ntypat, int lmnmax, const int* indlmn, const int* nattyp, const int* atindx1) {
  uint64_t k = hash_ints(indlmn, (size_t)6 * lmnmax * ntypat);
  k = hash_ints(nattyp, ntypat, k);
  if (atindx1) k = hash_ints(atindx1, natom, k);
  int meta[3] = {natom, ntypat, lmnmax};
  k = hash_ints(meta, 3, k);
  // the hash only short-cuts the comparison: a hit must also match the stored tables themselves
  const bool same = k == g_call_atoms_key && g_call_atoms.d_proj_typ != nullptr && g_call_atoms.natom == natom &&
                    g_call_atoms.ntypat == ntypat && g_call_atoms.lmnmax == lmnmax &&
                    g_call_atoms.indlmn.size() == (size_t)6 * lmnmax * ntypat &&
                    memcmp(g_call_atoms.indlmn.data(), indlmn, sizeof(int) * 6 * (size_t)lmnmax * ntypat) == 0 &&
                    g_call_atoms.nattyp.size() == (size_t)ntypat && memcmp(g_call_atoms.nattyp.data(), nattyp, sizeof(int) * ntypat) == 0 &&
                    (atindx1 == nullptr || (g_call_atoms.atindx1.size() == (size_t)natom &&
                                            memcmp(g_call_atoms.atindx1.data(), atindx1, sizeof(int) * natom) == 0));
  if (!same) {
    std::vector<int> ident(natom);
    for (int i = 0; i < natom; i++) ident[i] = i + 1;
    g_call_atoms.build(natom, ntypat, lmnmax, indlmn, nattyp, atindx1 ? atindx1 : ident.data());
    g_call_atoms_key = k;
  }
  return g_call_atoms;
}

void nonlop_release_all() {
  for (auto& kv : g_slots) kv.second->P.release();
  g_slots.clear();
  g_call_atoms.release(); g_call_atoms_key = 0;
  g_call_enl.release();
  nonlop_release_workspace();
}

#ifndef ABI_EMU
// ghc = (kinpw < huge*1e-11) ? ghc + kinpw*cwavef + gvnlxc : 0 ; gsc zeroed likewise (m_getghc.F90:1266-1280)
__global__ void k_assemble(double2* __restrict__ ghc, double2* __restrict__ gsc, const double* __restrict__ kinpw,
                           const double2* __restrict__ cwavef, const double2* __restrict__ gvnlxc, int npw, int ndat,
                           double kin_filter) {
  const long long total = (long long)npw * ndat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ig = (int)(i % npw);
    const double k = kinpw[ig];
    double2 v = make_double2(0.0, 0.0);
    if (k < kin_filter) {
      v = ghc[i];
      const double2 c = cwavef[i];
      v.x = v.x + k * c.x; v.y = v.y + k * c.y;
      if (gvnlxc) { v.x += gvnlxc[i].x; v.y += gvnlxc[i].y; }
    } else if (gsc) {
      gsc[i] = make_double2(0.0, 0.0);
    }
    ghc[i] = v;
  }
}
// strict: type_calc=1 filter "kinpw > huge*1e-11" (m_getghc.F90:1003-1031); otherwise "not (kinpw < huge*1e-11)" (:1272-1277)
// nspinor = 2: cwavef(2, npw, 2, ndat) -> compact up / dn blocks (m_getghc.F90:495-520)
__global__ void k_spinor_split(const double2* __restrict__ cw, double2* __restrict__ up, double2* __restrict__ dn, int npw, int ndat) {
  const long long total = (long long)npw * ndat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / npw; const int g = (int)(i - b * npw);
    up[i] = cw[(2 * b) * npw + g]; dn[i] = cw[(2 * b + 1) * npw + g];
  }
}
// ghc_up = g1 + g4, ghc_dn = g3 + g2 (m_getghc.F90:806-830) [+ kinpw psi, filtered (:1266-1280) when with_kin;
// filter only (type_calc = 1, :1003-1031) otherwise]
__global__ void k_spinor_combine(double2* __restrict__ ghc, const double2* __restrict__ g1, const double2* __restrict__ g2,
                                 const double2* __restrict__ g3, const double2* __restrict__ g4, const double2* __restrict__ cw,
                                 const double* __restrict__ kinpw, int npw, int ndat, double kin_filter, int with_kin) {
  const long long total = (long long)npw * ndat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / npw; const int g = (int)(i - b * npw);
    const double k = kinpw[g];
    const long long ou = (2 * b) * npw + g, od = (2 * b + 1) * npw + g;
    double2 u = make_double2(g1[i].x + g4[i].x, g1[i].y + g4[i].y), d = make_double2(g3[i].x + g2[i].x, g3[i].y + g2[i].y);
    if (with_kin) {
      if (k < kin_filter) { const double2 cu = cw[ou], cd = cw[od]; u.x += k * cu.x; u.y += k * cu.y; d.x += k * cd.x; d.y += k * cd.y; }
      else { u = make_double2(0.0, 0.0); d = u; }
    } else if (k > kin_filter) { u = make_double2(0.0, 0.0); d = u; }
    ghc[ou] = u; ghc[od] = d;
  }
}

__global__ void k_filter_only(double2* __restrict__ ghc, const double* __restrict__ kinpw, int npw, int ndat, double kin_filter,
                              bool strict) {
  const long long total = (long long)npw * ndat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const double k = kinpw[i % npw];
    if (strict ? (k > kin_filter) : !(k < kin_filter)) ghc[i] = make_double2(0.0, 0.0);
  }
}
#endif
}  // namespace abi


extern "C" {

void abi_b200_init_gemm_nonlop_(int* nkpt) { (void)nkpt; ensure_init(); }
void abi_b200_destroy_gemm_nonlop_(void) { nonlop_release_all(); }
void abi_b200_set_gemm_nonlop_ikpt_(int* ikpt) { g_cur_slot = *ikpt; }
long long abi_b200_nonlop_counter(void) { return ctx().nonlop_counter; }

static NonlopSlot& slot(int ikpt) {
  auto& s = g_slots[ikpt];
  if (!s) s = std::make_unique<NonlopSlot>();
  return *s;
}

void abi_b200_prep_projectors_(int* ikpt, int* npw, int* lmnmax, int* ntypat, int* indlmn, int* nattyp, int* istwf_k,
                               double* ucvol, double* ffnl, double* ph3d, int* dimffnl, int* matblk) {
  ensure_init();
  Context& c = ctx();
  int natom = 0;
  for (int t = 0; t < *ntypat; t++) natom += nattyp[t];
  const NonlopAtoms& at = call_atoms(natom, *ntypat, *lmnmax, indlmn, nattyp, nullptr);
  NonlopSlot& s = slot(*ikpt);
  s.P.alloc(*npw, at.nprojs, *istwf_k);
  DevArg a_ffnl(6, ffnl, sizeof(double) * (size_t)(*npw) * (*dimffnl) * (*lmnmax) * (*ntypat), true);
  DevArg a_ph3d(7, ph3d, sizeof(double) * 2 * (size_t)(*npw) * (*matblk), true);
  prep_projectors_device(s.P, at, a_ffnl.as<double>(), *dimffnl, a_ph3d.as<double>(), *matblk, *ucvol, c.stream);
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  g_cur_slot = *ikpt;
}

void abi_b200_set_projectors_(int* ikpt, int* npw, int* nprojs, int* istwf_k, double* projs) {
  ensure_init();
  Context& c = ctx();
  NonlopSlot& s = slot(*ikpt);
  s.P.alloc(*npw, *nprojs, *istwf_k);
  CUDA_CHECK(cudaMemcpyAsync(s.P.d_p, projs, sizeof(double) * 2 * (size_t)(*npw) * (*nprojs), cudaMemcpyDefault, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  g_cur_slot = *ikpt;
}

void abi_b200_initylmg_k_(double* gprimd, int* kg, double* kpt, int* mpsang, int* npw, double* ylm) {
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(*mpsang >= 1 && *mpsang <= 4, "initylmg: mpsang must be in 1..4");
  ABI_CHECK(!is_device_ptr(gprimd) && !is_device_ptr(kpt), "initylmg: gprimd and kpt are host arrays");
  DevArg a_ylm(0, ylm, sizeof(double) * (size_t)(*npw) * (*mpsang) * (*mpsang), false);
  DevArg a_kg(3, kg, sizeof(int) * 3 * (size_t)(*npw), true);
  DevArg a_gp(5, gprimd, sizeof(double) * 9, true);
  initylmg_device(a_ylm.as<double>(), *npw, *mpsang, a_kg.as<int>(), kpt, a_gp.as<double>(), c.stream);
  a_ylm.copy_back();
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_mkffnl_(int* dimekb, int* dimffnl, double* ekb, double* ffnl, double* ffspl, double* gmet, double* gprimd, int* ider, int* idir,
                      int* indlmn, int* kg, double* kpg, double* kpt, int* lmnmax, int* lnmax, int* mpsang, int* mqgrid, int* nkpg, int* npw,
                      int* ntypat, int* pspso, double* qgrid, double* rmet, int* usepaw, int* useylm, double* ylm, double* ylm_gr) {
  (void)gmet; (void)kpg; (void)rmet; (void)ylm_gr; (void)nkpg;
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(*ider == 0 && *idir == 0 && *dimffnl == 1, "mkffnl: only ider=0, idir=0 (dimffnl=1) is on the getghc path");
  ABI_CHECK(*useylm == 1, "mkffnl: useylm=1 is required (gemm_nonlop path)");
  ABI_CHECK(*mpsang <= 4, "Called with mpsang > 4: this subroutine will not accept lmax+1 > 4.");      // m_mkffnl.F90:291-296
  ABI_CHECK(*mqgrid >= 2, "mkffnl: mqgrid must be >= 2");
  ABI_CHECK(!is_device_ptr(indlmn) && !is_device_ptr(qgrid) && !is_device_ptr(kpt) && !is_device_ptr(gprimd), "mkffnl: small tables are host arrays");
  const int np = *npw, lm = *lmnmax, nt = *ntypat;
  // testnl (m_mkffnl.F90:452-456): PAW always; NC only where |ekb(iln,itypat)| > tol10; channel must have indlmn(6)=1 or pspso/=0
  std::vector<unsigned char> active((size_t)lm * nt, 0);
  for (int t = 0; t < nt; t++) for (int i = 0; i < lm; i++) {
    const int* il = indlmn + 6 * (i + (size_t)lm * t);
    if (il[2] <= 0) continue;
    bool on = (il[5] == 1) || (pspso && pspso[t] != 0);
    if (on && *usepaw == 0) on = std::fabs(ekb[(il[4] - 1) + (size_t)(*dimekb) * t]) > 1e-10;
    active[i + (size_t)lm * t] = on ? 1 : 0;
  }
  const double q0 = qgrid[0], dq = (qgrid[*mqgrid - 1] - qgrid[0]) / (double)(*mqgrid - 1);
  ABI_CHECK(dq >= 1e-12, "spacing should be strictly positive");
  DevArg a_ffnl(0, ffnl, sizeof(double) * (size_t)np * lm * nt, false);
  DevArg a_ffspl(1, ffspl, sizeof(double) * (size_t)(*mqgrid) * 2 * (*lnmax) * nt, true);
  DevArg a_ylm(2, ylm, sizeof(double) * (size_t)np * (*mpsang) * (*mpsang), true);
  DevArg a_kg(3, kg, sizeof(int) * 3 * (size_t)np, true);
  DevArg a_ind(4, indlmn, sizeof(int) * 6 * (size_t)lm * nt, true);
  DevArg a_gp(5, gprimd, sizeof(double) * 9, true);
  DevArg a_act(6, active.data(), active.size(), true);
  mkffnl_device(a_ffnl.as<double>(), np, lm, nt, a_ind.as<int>(), a_kg.as<int>(), kpt, a_gp.as<double>(), a_ffspl.as<double>(), *mqgrid, *lnmax,
                q0, dq, a_ylm.as<double>(), a_act.as<unsigned char>(), c.stream);
  a_ffnl.copy_back();
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_gemm_nonlop_(int* atindx1, int* choice, int* cpopt, double* vectproj, int* dimenl1, int* dimenl2, int* dimekbq,
                           double* enl, int* indlmn, int* istwf_k, double* lambda, int* lmnmax, int* natom, int* nattyp,
                           int* ndat, int* nnlout, int* npwin, int* npwout, int* nspinor, int* nspinortot, int* ntypat,
                           int* paw_opt, double* sij, double* svectout, int* useylm, double* vectin, double* vectout,
                           int* signs) {
  (void)nnlout;
  ensure_init();
  Context& c = ctx();
  c.nonlop_counter += *ndat;                               // m_nonlop.F90:389-392
  ABI_CHECK(*signs == 2, "gemm_nonlop: only signs=2 is on the getghc path (signs=1 contractions are out of scope)");
  ABI_CHECK(*useylm == 1, "gemm_nonlop requires useylm=1 (m_invars2.F90:2829)");
  ABI_CHECK((*nspinor == 1 || *nspinor == 2) && *nspinor == *nspinortot,
            "gemm_nonlop: nspinor must be 1 or 2 and equal to nspinortot (no parallelisation over spinors, as in the reference's GPU paths)");
  ABI_CHECK(*dimekbq == 1, "gemm_nonlop: dimekbq=2 (q-dependent D_ij, DFPT) is outside the getghc path");
  ABI_CHECK(*npwin == *npwout, "gemm_nonlop: k/=k' is not implemented");
  auto it = g_slots.find(g_cur_slot);
  ABI_CHECK(it != g_slots.end() && it->second->P.d_p != nullptr, "gemm_nonlop: projectors not prepared for the current ikpt");
  const Projectors& P = it->second->P;
  ABI_CHECK(P.npw == *npwin, "gemm_nonlop: npw differs from the prepared projectors");
  ABI_CHECK(P.istwf_k == *istwf_k, "gemm_nonlop: istwf_k differs from the prepared projectors");
  const NonlopAtoms& at = call_atoms(*natom, *ntypat, *lmnmax, indlmn, nattyp, atindx1);
  if (*choice != 0 && *choice != 7) {
    ABI_CHECK(enl != nullptr, "gemm_nonlop: enl is required");
    // enl(dimenl1, dimenl2, nspinortot**2): NC ekb(lnmax, ntypat, nspinortot**2) uses its first block (identical for both spinor
    // components without spin-orbit); PAW D_ij real or complex (cplex_dij = dimenl1 / lmn2), four blocks with spinors
    const int lmn2 = *lmnmax * (*lmnmax + 1) / 2;
    const int nblk = (*paw_opt == 0) ? 1 : (*nspinortot) * (*nspinortot);
    g_call_enl.load(enl, *dimenl1, *dimenl2, (*paw_opt >= 2) ? sij : nullptr, *ntypat, c.stream, nblk, *paw_opt == 0 ? 0 : lmn2);
    g_call_enl.cplex_enl = (*paw_opt != 0 && *dimenl1 == 2 * lmn2) ? 2 : 1;
  }
  g_call_enl.nspinor = (*paw_opt == 0) ? 1 : *nspinor;       // NC: every (band, spinor) column is independent
  const int cplex = (*istwf_k == 1) ? 2 : 1;
  const int ncol = *ndat * *nspinor;                         // vectin(2, npw*nspinor*ndat)
  const size_t nv = sizeof(double) * 2 * (size_t)P.npw * ncol;
  const size_t np = sizeof(double) * (size_t)cplex * P.nprojs * ncol;
  DevArg a_in(0, vectin, nv, true);
  DevArg a_out(1, vectout, nv, false);
  DevArg a_sout(2, svectout, nv, false);
  DevArg a_proj(3, vectproj, np, *cpopt >= 2);
  DevArg a_lam(4, lambda, sizeof(double) * (*ndat), true);
  ABI_CHECK(!(*nspinor == 2 && *paw_opt == 2), "gemm_nonlop: paw_opt=2 with nspinor=2 needs lambda per band; use the handle-based nonlop");
  gemm_nonlop_device(P, at, g_call_enl, *choice, *cpopt, *paw_opt, c.me_g0, a_lam.as<double>(), ncol, a_in.as<double>(),
                     a_out.as<double>(), a_sout.as<double>(), a_proj.as<double>(), c.stream);
  const bool want_v = *choice == 1 && (*paw_opt == 0 || *paw_opt == 1 || *paw_opt == 2 || *paw_opt == 4);
  const bool want_s = *choice == 7 || (*choice == 1 && (*paw_opt == 3 || *paw_opt == 4));
  if (want_v) a_out.copy_back();
  if (want_s) a_sout.copy_back();
  if ((*cpopt >= 0 && *cpopt < 2) || *choice == 0) a_proj.copy_back();
  if (!c.async || a_in.staged || a_out.staged || a_sout.staged || a_proj.staged) CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// ---------------------------------------------------------------------------------------------------------
// Hamiltonian handle + getghc
// ---------------------------------------------------------------------------------------------------------
abi_b200_ham_t* abi_b200_ham_create(const int* ngfft, int natom, int ntypat, int lmnmax, const int* indlmn, const int* nattyp,
                                    const int* atindx1, int usepaw, double ucvol) {
  ensure_init();
  auto* h = new abi_b200_ham();
  memcpy(h->ngfft, ngfft, sizeof(int) * 18);
  h->natom = natom; h->ntypat = ntypat; h->lmnmax = lmnmax; h->usepaw = usepaw; h->ucvol = ucvol;
  h->atoms.build(natom, ntypat, lmnmax, indlmn, nattyp, atindx1);
  for (int i = 0; i < 3; i++) fft_tables(ngfft[i]);
  return h;
}

void abi_b200_ham_destroy(abi_b200_ham_t* h) {
  if (!h) return;
  if (ctx().initialized) CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
  h->atoms.release(); h->enl.release(); h->P.release(); h->invovl.release();
  for (VlocDev* v : {&h->vloc, &h->vloc22, &h->vloc_ud, &h->vloc_du}) { if (v->d_v) cudaFree(v->d_v); if (v->d_vT) cudaFree(v->d_vT); }
  if (h->d_spin_tmp) cudaFree(h->d_spin_tmp);
  if (h->d_kinpw) cudaFree(h->d_kinpw);
  if (h->d_gvnlxc) cudaFree(h->d_gvnlxc);
  delete h;
}

void abi_b200_ham_load_spin(abi_b200_ham_t* h, const double* vlocal, int cplex_vloc, int n4, int n5, int n6) {
  h->epoch = ++ham_epoch_counter();
  ensure_init();
  ABI_CHECK(n4 == h->ngfft[0] && n5 == h->ngfft[1] && n6 == h->ngfft[2],
            "FFT SIZE ERROR: when gpu mode is on the fft grid must not be augmented (n4,n5,n6 must equal n1,n2,n3)");
  ABI_CHECK(cplex_vloc == 1 || cplex_vloc == 2, "vlocal must be real (cplex=1) or complex (cplex=2)");
  vloc_upload(h->vloc, vlocal, is_device_ptr(vlocal), cplex_vloc, n4, n5, n6, ctx().stream);
  h->nvloc = 1;
  CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
}

void abi_b200_ham_set_nspinor(abi_b200_ham_t* h, int nspinor) {
  h->epoch = ++ham_epoch_counter();
  ABI_CHECK(nspinor == 1 || nspinor == 2, "nspinor must be 1 or 2");
  // PAW spinors: D_ij comes as four complex blocks (abi_b200_ham_load_enl_spinor), m_opernlc_ylm_allwf.F90:660-737
  h->nspinor = nspinor;
  h->enl.nspinor = h->usepaw ? nspinor : 1;
}

#ifndef ABI_EMU
// (re, im) planes -> one interleaved complex potential, im scaled by `sign`
__global__ void k_pack_vud(const double* __restrict__ v3, const double* __restrict__ v4, double* __restrict__ out, long long n, double sign) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out[2 * i] = v3[i]; out[2 * i + 1] = sign * v4[i];
  }
}
#endif

void abi_b200_ham_load_spin_nvloc(abi_b200_ham_t* h, const double* vlocal, int nvloc, int n4, int n5, int n6) {
  h->epoch = ++ham_epoch_counter();
  ensure_init();
  ABI_CHECK(nvloc == 1 || nvloc == 4, "nvloc must be 1 or 4");
  if (nvloc == 1) { abi_b200_ham_load_spin(h, vlocal, 1, n4, n5, n6); h->nvloc = 1; return; }
  ABI_CHECK(h->nspinor == 2, "nvloc=4 requires nspinor=2 (abi_b200_ham_set_nspinor)");
  ABI_CHECK(n4 == h->ngfft[0] && n5 == h->ngfft[1] && n6 == h->ngfft[2],
            "FFT SIZE ERROR: when gpu mode is on the fft grid must not be augmented (n4,n5,n6 must equal n1,n2,n3)");
#ifndef ABI_EMU
  Context& c = ctx();
  const size_t N = (size_t)n4 * n5 * n6;
  const bool dev = is_device_ptr(vlocal);
  vloc_upload(h->vloc, vlocal, dev, 1, n4, n5, n6, c.stream);                  // vlocal(:,:,:,1)
  vloc_upload(h->vloc22, vlocal + N, dev, 1, n4, n5, n6, c.stream);            // vlocal(:,:,:,2)
  DevArg a34(9, vlocal + 2 * N, sizeof(double) * 2 * N, true);                  // vlocal(:,:,:,3:4)
  if (h->spin_tmp_cap < 2 * N) { if (h->d_spin_tmp) cudaFree(h->d_spin_tmp); CUDA_CHECK(cudaMalloc(&h->d_spin_tmp, sizeof(double) * 2 * N)); h->spin_tmp_cap = 2 * N; }
  // m_getghc.F90:771-776: psi_dn -> ghc_up sees (V3, +V4); :737-742: psi_up -> ghc_dn sees (V3, -V4)
  k_pack_vud<<<kNumSM * 4, 256, 0, c.stream>>>(a34.as<double>(), a34.as<double>() + N, h->d_spin_tmp, (long long)N, 1.0);
  vloc_upload(h->vloc_ud, h->d_spin_tmp, true, 2, n4, n5, n6, c.stream);
  k_pack_vud<<<kNumSM * 4, 256, 0, c.stream>>>(a34.as<double>(), a34.as<double>() + N, h->d_spin_tmp, (long long)N, -1.0);
  vloc_upload(h->vloc_du, h->d_spin_tmp, true, 2, n4, n5, n6, c.stream);
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
#endif
  h->nvloc = 4;
}

void abi_b200_ham_load_enl_spinor(abi_b200_ham_t* h, const double* enl, int dimenl1, int dimenl2, int nspinortot2, const double* sij) {
  h->epoch = ++ham_epoch_counter();
  ensure_init();
  ABI_CHECK(nspinortot2 == 1 || nspinortot2 == 4, "load_enl: enl(dimenl1, dimenl2, nspinortot**2) needs nspinortot**2 = 1 or 4");
  const int lmn2 = h->lmnmax * (h->lmnmax + 1) / 2;
  int cplex_enl = 1;
  if (h->usepaw) {
    ABI_CHECK(dimenl1 == lmn2 || dimenl1 == 2 * lmn2, "load_enl: PAW D_ij must be packed, dimenl1 = cplex_dij * lmnmax*(lmnmax+1)/2");
    cplex_enl = dimenl1 / lmn2;
    ABI_CHECK(nspinortot2 == 1 || cplex_enl == 2, "load_enl: spinor D_ij blocks must be complex (m_opernlc_ylm_allwf.F90:665)");
  }
  h->enl.load(enl, dimenl1, dimenl2, sij, h->ntypat, ctx().stream, h->usepaw ? nspinortot2 : 1, h->usepaw ? lmn2 : 0);
  h->enl.cplex_enl = cplex_enl;
  h->enl.nspinor = h->usepaw ? h->nspinor : 1;
  h->invovl.release();
}
void abi_b200_ham_load_enl(abi_b200_ham_t* h, const double* enl, int dimenl1, int dimenl2, const double* sij) {
  abi_b200_ham_load_enl_spinor(h, enl, dimenl1, dimenl2, 1, sij);
}

void abi_b200_ham_load_k(abi_b200_ham_t* h, int istwf_k, int npw, const int* kg_k, const double* kinpw, const double* ffnl,
                         int dimffnl, const double* ph3d, int matblk, int me_g0) {
  h->epoch = ++ham_epoch_counter();
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(!is_device_ptr(kg_k), "kg_k must be a host array");
  h->istwf_k = istwf_k; h->npw = npw; h->me_g0 = me_g0;
  h->invovl.release();
  h->kg.assign(kg_k, kg_k + (size_t)3 * npw);
  h->plan_ref = fourwf_get_plan_shared(h->kg.data(), npw, h->kg.data(), npw, h->ngfft, istwf_k, me_g0);
  h->plan = h->plan_ref.get();
  if (h->d_kinpw) cudaFree(h->d_kinpw);
  CUDA_CHECK(cudaMalloc(&h->d_kinpw, sizeof(double) * std::max(1, npw)));
  CUDA_CHECK(cudaMemcpyAsync(h->d_kinpw, kinpw, sizeof(double) * npw, cudaMemcpyDefault, c.stream));
  if (ffnl && ph3d) {
    h->P.alloc(npw, h->atoms.nprojs, istwf_k);
    DevArg a_ffnl(6, ffnl, sizeof(double) * (size_t)npw * dimffnl * h->lmnmax * h->ntypat, true);
    DevArg a_ph3d(7, ph3d, sizeof(double) * 2 * (size_t)npw * matblk, true);
    prep_projectors_device(h->P, h->atoms, a_ffnl.as<double>(), dimffnl, a_ph3d.as<double>(), matblk, h->ucvol, c.stream);
  } else {
    h->P.npw = npw; h->P.istwf_k = istwf_k;     // projectors to be installed with abi_b200_ham_set_projectors
  }
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_ham_load_k_xred(abi_b200_ham_t* h, int istwf_k, int npw, const int* kg_k, const double* kinpw, const double* ffnl,
                              int dimffnl, const double* kpt, const double* xred, int me_g0) {
  h->epoch = ++ham_epoch_counter();
  ABI_CHECK(ffnl != nullptr && kpt != nullptr && xred != nullptr, "load_k_xred: ffnl, kpt and xred are required");
  abi_b200_ham_load_k(h, istwf_k, npw, kg_k, kinpw, nullptr, 0, nullptr, 0, me_g0);
  Context& c = ctx();
  h->P.alloc(npw, h->atoms.nprojs, istwf_k);
  DevArg a_ffnl(6, ffnl, sizeof(double) * (size_t)npw * dimffnl * h->lmnmax * h->ntypat, true);
  DevArg a_kg(7, kg_k, sizeof(int) * 3 * (size_t)npw, true);
  DevArg a_x(8, xred, sizeof(double) * 3 * (size_t)h->natom, true);
  prep_projectors_xred_device(h->P, h->atoms, a_ffnl.as<double>(), dimffnl, a_kg.as<int>(), a_x.as<double>(), kpt, h->ucvol, c.stream);
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void abi_b200_ham_set_projectors(abi_b200_ham_t* h, const double* projs, int nprojs) {
  h->epoch = ++ham_epoch_counter();
  ensure_init();
  ABI_CHECK(nprojs == h->atoms.nprojs, "set_projectors: nprojs differs from sum(nlmn*nattyp)");
  h->P.alloc(h->npw, nprojs, h->istwf_k);
  h->invovl.release();
  CUDA_CHECK(cudaMemcpyAsync(h->P.d_p, projs, sizeof(double) * 2 * (size_t)h->npw * nprojs, cudaMemcpyDefault, ctx().stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
}

int abi_b200_ham_nprojs(const abi_b200_ham_t* h) { return h->atoms.nprojs; }

namespace {
// host <-> device pipeline of one getghc call with HOST wavefunction arrays: band chunks of cwavef arrive on the copy
// stream while fourwf runs on earlier chunks; finished row slabs of ghc leave while the last GEMM computes the next.
struct GhcPipe {
  Context* c; cudaEvent_t ev[2]; int flip = 0;
  double* h_ghc; double* d_ghc; int npw, nd;
};
cudaEvent_t pipe_event(int i) {
  static cudaEvent_t ev[32]; static bool init = false;
  if (!init) { for (auto& e : ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); init = true; }
  return ev[i & 31];
}
int g_pipe_ev = 0;
void ship_slab(void* user, int ipw_begin, int ipw_end) {
  GhcPipe* gp = static_cast<GhcPipe*>(user);
  cudaEvent_t e = pipe_event(g_pipe_ev++);
  CUDA_CHECK(cudaEventRecord(e, gp->c->stream));
  CUDA_CHECK(cudaStreamWaitEvent(gp->c->copy_stream, e, 0));
  const size_t pitch = sizeof(double) * 2 * (size_t)gp->npw;
  CUDA_CHECK(cudaMemcpy2DAsync(gp->h_ghc + 2 * (size_t)ipw_begin, pitch, gp->d_ghc + 2 * (size_t)ipw_begin, pitch,
                               sizeof(double) * 2 * (size_t)(ipw_end - ipw_begin), gp->nd, cudaMemcpyDeviceToHost,
                               gp->c->copy_stream));
}
}  // namespace

void abi_b200_getghc_(int* cpopt, double* cwavef, double* cwaveprj, double* ghc, double* gsc, abi_b200_ham_t** gs_ham,
                      double* gvnlxc, double* lambda, int* ndat, int* prtvol, int* sij_opt, int* tim_getghc, int* type_calc) {
  (void)prtvol; (void)tim_getghc;
  ensure_init();
  NvtxRange nvtx_getghc("GETGHC");                       // NVTX_GETGHC, m_getghc.F90:264
  Context& c = ctx();
  abi_b200_ham* h = *gs_ham;
  // nspinor = 2 (NC, istwf_k = 1): the block is cwavef(2, npw*nspinor*ndat) -- every (band, spinor) pair is a column of npw
  // coefficients on which the kinetic and non-local terms act separately (m_getghc.F90:1266-1280, opernlc NC branch with
  // ekb(:,:,ispinor) equal for both components); only the local part couples them, and only when nvloc = 4.
  if (h->nspinor == 2) {
    ABI_CHECK(h->istwf_k == 1, "getghc: nspinor=2 requires istwf_k=1");
    ABI_CHECK(h->usepaw == 0 || (h->enl.nblk == 4 && h->enl.cplex_enl == 2),
              "getghc: nspinor=2 with PAW needs the four complex D_ij blocks (abi_b200_ham_load_enl_spinor)");
    ABI_CHECK(!(h->usepaw == 1 && *sij_opt == -1), "getghc: sij_opt=-1 (H - lambda S) with nspinor=2 is not implemented");
  }
  const int tc = *type_calc, nd = *ndat * h->nspinor, npw = h->npw;
  ABI_CHECK(tc >= 0 && tc <= 3, "getghc: type_calc must be 0, 1, 2 or 3");
  ABI_CHECK(h->plan != nullptr, "getghc: load_k has not been called");
  ABI_CHECK(!(*sij_opt != 0 && h->usepaw == 0), "getghc: sij_opt/=0 requires PAW");   // m_getghc.F90:333-336
  const bool local = (tc == 0 || tc == 1 || tc == 3), nonlocal = (tc == 0 || tc == 2);
  if (local) ABI_CHECK(h->vloc.d_v != nullptr, "We need vlocal in gs_ham!");             // m_getghc.F90:404
  const size_t nv = sizeof(double) * 2 * (size_t)npw * nd;
  const int cplex = (h->istwf_k == 1) ? 2 : 1;
  const bool fused_fw = local && h->plan->fused_ok && c.fourwf_impl != 1;
  // pipelined staging needs the fused fourwf (band-chunked) and host input + output arrays
  const bool pipe = fused_fw && nd >= 8 && cwavef && ghc && !is_device_ptr(cwavef) && !is_device_ptr(ghc) && c.pipeline &&
                    h->nvloc == 1;
  DevArg a_c(0, cwavef, nv, !pipe);
  DevArg a_ghc(1, ghc, nv, tc == 2);
  DevArg a_gsc(2, (*sij_opt == 1) ? gsc : nullptr, nv, false);
  DevArg a_gv(3, gvnlxc, nv, false);
  const int cpopt_here = (h->usepaw == 1) ? *cpopt : -1;                                 // m_getghc.F90:1046
  DevArg a_prj(4, (cpopt_here >= 0) ? cwaveprj : nullptr, sizeof(double) * (size_t)cplex * h->atoms.nprojs * nd, cpopt_here >= 2);
  DevArg a_lam(5, lambda, sizeof(double) * (*ndat), true);     // the reference passes lambda(ndat): one value per band, spinors share it
  const double kin_filter = 1.7976931348623157e308 * 1.0e-11;
  int paw_opt = h->usepaw; if (*sij_opt != 0) paw_opt = *sij_opt + 3;                    // m_getghc.F90:1067
  if (nonlocal)
    ABI_CHECK(h->P.d_p != nullptr || h->atoms.nprojs == 0, "getghc: projectors not loaded (load_k with ffnl/ph3d or set_projectors)");
  bool ghc_shipped = false;

#ifndef ABI_EMU
  if (local && h->nvloc == 4) {
    // ---- non-collinear local part (m_getghc.F90:655-830): four applications on the compacted spinor components
    //   ghc_up = F[V11] psi_up + F[V3 + i V4] psi_dn ;  ghc_dn = F[V3 - i V4] psi_up + F[V22] psi_dn
    ABI_CHECK(h->vloc22.d_v && h->vloc_ud.d_v && h->vloc_du.d_v, "We need vlocal(:,:,:,1:4) in gs_ham! (load_spin_nvloc)");
    const int nb = *ndat;
    c.fourwf_counter += 8 * nb;
    const size_t cb = (size_t)npw * nb;                                   // double2 per compact block
    double2* t = reinterpret_cast<double2*>(c.stage[11].get(sizeof(double2) * 6 * cb));
    double2 *up = t, *dn = t + cb, *g1 = t + 2 * cb, *g2 = t + 3 * cb, *g3 = t + 4 * cb, *g4 = t + 5 * cb;
    const int blocks = std::min(kNumSM * 8, (int)ceil_div<long long>((long long)npw * nb, 256));
    k_spinor_split<<<blocks, 256, 0, c.stream>>>(a_c.as<double2>(), up, dn, npw, nb);
    auto apply_local = [&](const VlocDev& v, const double2* in, double2* out) {
      if (fused_fw) { FourwfEpilogue e0; fourwf_fused_opt2(*h->plan, v, in, out, nb, e0, c.stream); }
      else fourwf_generic(*h->plan, 2, v.cplex, v.d_v, in, out, nullptr, nb, nullptr, nullptr, c.stream);
    };
    apply_local(h->vloc, up, g1); apply_local(h->vloc22, dn, g2);
    apply_local(h->vloc_du, up, g3); apply_local(h->vloc_ud, dn, g4);
    k_spinor_combine<<<blocks, 256, 0, c.stream>>>(a_ghc.as<double2>(), g1, g2, g3, g4, a_c.as<double2>(), h->d_kinpw, npw, nb,
                                                   kin_filter, tc == 1 ? 0 : 1);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches += 2;
  } else
#endif
  if (local) {
    // ---- local part first: ghc = V_loc psi + T psi (filtered); the non-local term is added by the last GEMM's epilogue
    NvtxRange nvtx_loc("LOCPOT");                        // NVTX_GETGHC_LOCPOT, m_getghc.F90:400
    NvtxRange nvtx_kin("KINETIC");                       // NVTX_GETGHC_KIN (m_getghc.F90:1162): fused into the last fourwf kernel here
    c.fourwf_counter += 2 * nd;
    FourwfEpilogue epi;
    if (tc == 1) { epi.mode = 2; epi.kinpw = h->d_kinpw; }
    else { epi.mode = 1; epi.kinpw = h->d_kinpw; epi.cwavef = a_c.as<double2>(); epi.gvnlxc = nullptr; epi.gsc = nullptr; }
    if (fused_fw) {
      if (!pipe) {
        fourwf_fused_opt2(*h->plan, h->vloc, a_c.as<double2>(), a_ghc.as<double2>(), nd, epi, c.stream);
      } else {
        // band chunks (even sizes keep the Gamma-point pairs together): H2D on the copy stream, fourwf behind an event
        // (16 chunks: the exposed head of the pipeline is the H2D of the first chunk only)
        const int nchunk = std::max(1, std::min(c.pipe_chunks, nd / 8));
        const int chunk = ceil_div(ceil_div(nd, nchunk), 2) * 2;
        for (int b0 = 0; b0 < nd; b0 += chunk) {
          const int nb = std::min(chunk, nd - b0);
          const size_t off = 2 * (size_t)npw * b0;
          CUDA_CHECK(cudaMemcpyAsync(a_c.as<double>() + off, cwavef + off, sizeof(double) * 2 * (size_t)npw * nb,
                                     cudaMemcpyHostToDevice, c.copy_stream));
          cudaEvent_t e = pipe_event(g_pipe_ev++);
          CUDA_CHECK(cudaEventRecord(e, c.copy_stream));
          CUDA_CHECK(cudaStreamWaitEvent(c.stream, e, 0));
          FourwfEpilogue ec = epi;
          if (ec.cwavef) ec.cwavef += (size_t)npw * b0;
          fourwf_fused_opt2(*h->plan, h->vloc, a_c.as<double2>() + (size_t)npw * b0, a_ghc.as<double2>() + (size_t)npw * b0, nb,
                            ec, c.stream);
        }
      }
    } else {
#ifndef ABI_EMU
      fourwf_generic(*h->plan, 2, h->vloc.cplex, h->vloc.d_v, a_c.as<double2>(), a_ghc.as<double2>(), nullptr, nd, nullptr, nullptr, c.stream);
      const int blocks = std::min(kNumSM * 8, (int)ceil_div<long long>((long long)npw * nd, 256));
      if (tc == 1) k_filter_only<<<blocks, 256, 0, c.stream>>>(a_ghc.as<double2>(), h->d_kinpw, npw, nd, kin_filter, true);
      else k_assemble<<<blocks, 256, 0, c.stream>>>(a_ghc.as<double2>(), nullptr, h->d_kinpw, a_c.as<double2>(), nullptr, npw, nd,
                                                    kin_filter);
      CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
#endif
    }
  }
#ifndef ABI_EMU
  if (nonlocal) {
    NvtxRange nvtx_nl("NLOCPOT");                        // NVTX_GETGHC_NLOCPOT, m_getghc.F90:1042
    c.nonlop_counter += nd;
    bool gsc_filtered = false;
    if (tc == 0) {
      NonlopFusion fuse;
      fuse.ghc = a_ghc.as<double>(); fuse.kinpw = h->d_kinpw; fuse.kin_filter = kin_filter;
      GhcPipe gp{&c, {}, 0, ghc, a_ghc.as<double>(), npw, nd};
      if (pipe) { fuse.nslabs = c.pipe_chunks; fuse.after_slab = ship_slab; fuse.user = &gp; ghc_shipped = true; }
      gemm_nonlop_device(h->P, h->atoms, h->enl, 1, cpopt_here, paw_opt, h->me_g0, a_lam.as<double>(), nd, a_c.as<double>(),
                         a_gv.as<double>(), a_gsc.as<double>(), a_prj.as<double>(), c.stream, &fuse);
      if (h->atoms.nprojs == 0 || nd == 0) ghc_shipped = false;          // nothing was launched: plain copy below
      gsc_filtered = fuse.gsc_filtered;
    } else {
      // type_calc == 2: non-local + kinetic added to the caller's ghc (m_getghc.F90:152)
      double* d_gv = a_gv.as<double>();
      if (d_gv == nullptr) {
        if (h->gvnlxc_cap < nv) { if (h->d_gvnlxc) cudaFree(h->d_gvnlxc); CUDA_CHECK(cudaMalloc(&h->d_gvnlxc, nv)); h->gvnlxc_cap = nv; }
        d_gv = h->d_gvnlxc;
      }
      gemm_nonlop_device(h->P, h->atoms, h->enl, 1, cpopt_here, paw_opt, h->me_g0, a_lam.as<double>(), nd, a_c.as<double>(), d_gv,
                         a_gsc.as<double>(), a_prj.as<double>(), c.stream);
      const int blocks = std::min(kNumSM * 8, (int)ceil_div<long long>((long long)npw * nd, 256));
      k_assemble<<<blocks, 256, 0, c.stream>>>(a_ghc.as<double2>(), nullptr, h->d_kinpw, a_c.as<double2>(), (const double2*)d_gv,
                                               npw, nd, kin_filter);
      CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
    if (*sij_opt == 1 && a_gsc.dev && !gsc_filtered) {                   // gsc = 0 where the kinetic filter strikes
      const int blocks = std::min(kNumSM * 8, (int)ceil_div<long long>((long long)npw * nd, 256));
      k_filter_only<<<blocks, 256, 0, c.stream>>>(a_gsc.as<double2>(), h->d_kinpw, npw, nd, kin_filter, false);
      CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
  } else if (tc == 3 && a_gv.dev) {
    CUDA_CHECK(cudaMemsetAsync(a_gv.dev, 0, nv, c.stream));                              // m_getghc.F90:1148-1154
  }
#endif
  if (!ghc_shipped) a_ghc.copy_back();
  if (*sij_opt == 1) a_gsc.copy_back();
  if (tc != 1) a_gv.copy_back();
  if (cpopt_here >= 0 && cpopt_here < 2) a_prj.copy_back();
  if (!c.async || pipe || a_c.staged || a_ghc.staged || a_gsc.staged || a_gv.staged || a_prj.staged) {
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    if (pipe) CUDA_CHECK(cudaStreamSynchronize(c.copy_stream));
  }
}


// ------------------------------------------------------------------------------------------------------
// Batched getghc for the many-small-k-points regime (BASELINE configs[2]: Fe-2 PAW, hundreds of (k, spin) pairs of ~24 bands and
// ~700 plane waves: 6-9 kernels of a few microseconds per call, i.e. launch-latency bound on one stream).
//   * the calls of a batch are dealt round-robin to kMaxLanes streams, each with its own set of internal workspaces, so the
//     kernels of different k-points overlap on the device;
//   * the kernel sequence of one (handle, arrays, ndat) combination is captured into a CUDA graph the second time it is seen
//     and replayed afterwards (one launch per call instead of 6-9); any load_* / set_* on the handle bumps its epoch and
//     drops the graph.
// Reference: the (k, spin) loop of src/79_seqpar_mpi/m_vtorho.F90:789-1045 around the eigensolver's getghc calls.
// ------------------------------------------------------------------------------------------------------
namespace {
struct GraphEntry { int state = 0; cudaGraphExec_t exec = nullptr; cudaGraph_t graph = nullptr; };
struct GraphKey {
  const void* h; unsigned epoch; const void* c; const void* g; const void* s; int ndat, sij, tc, lane;
  bool operator<(const GraphKey& o) const {
    return std::tie(h, epoch, c, g, s, ndat, sij, tc, lane) < std::tie(o.h, o.epoch, o.c, o.g, o.s, o.ndat, o.sij, o.tc, o.lane);
  }
};
std::map<GraphKey, GraphEntry>& graph_cache() { static std::map<GraphKey, GraphEntry> c; return c; }
cudaEvent_t g_lane_ev[kMaxLanes + 1] = {};
// One graph for a whole sweep: the per-call graphs as child nodes, chained per lane (8 parallel branches), built the first time
// every call of a batch has its own graph and replayed with ONE launch while the batch (handles, epochs, arrays, order) is unchanged.
struct SweepGraph { std::vector<GraphKey> keys; cudaGraphExec_t exec = nullptr; int seen = 0; };
SweepGraph& sweep_graph() { static SweepGraph g; return g; }
bool same_keys(const std::vector<GraphKey>& a, const std::vector<GraphKey>& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++) if (a[i] < b[i] || b[i] < a[i]) return false;
  return true;
}
void sweep_graph_drop() {
  SweepGraph& g = sweep_graph();
#ifndef ABI_EMU
  if (g.exec) cudaGraphExecDestroy(g.exec);
#endif
  g = SweepGraph();
}
}  // namespace

void abi_b200_graphs_clear(void) {
#ifndef ABI_EMU
  for (auto& kv : graph_cache()) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); if (kv.second.graph) cudaGraphDestroy(kv.second.graph); }
#endif
  sweep_graph_drop();
  graph_cache().clear();
}

void abi_b200_getghc_batch_(int* nk, abi_b200_ham_t** hams, double** cwavef, double** ghc, double** gsc, int* ndat, int* sij_opt,
                            int* type_calc, int* use_graphs) {
  ensure_init();
  Context& c = ctx();
  ABI_CHECK(*nk >= 0, "getghc_batch: nk must be >= 0");
#ifdef ABI_EMU
  for (int i = 0; i < *nk; i++) {
    int cpopt = -1, prtvol = 0, tim = 0;
    abi_b200_getghc_(&cpopt, cwavef[i], nullptr, ghc[i], gsc ? gsc[i] : nullptr, &hams[i], nullptr, nullptr, ndat, &prtvol, sij_opt, &tim, type_calc);
  }
  (void)use_graphs;
#else
  const bool was_async = c.async;
  cudaStream_t main_stream = c.stream;
  const int nlanes = std::min(kMaxLanes, std::max(1, *nk));
  // ---- whole-sweep graph: one launch for the batch when nothing changed since it was built
  std::vector<GraphKey> keys;
  if (*use_graphs && *nk > 1) {
    keys.reserve(*nk);
    for (int i = 0; i < *nk; i++)
      keys.push_back(GraphKey{hams[i], hams[i]->epoch, cwavef[i], ghc[i], gsc ? gsc[i] : nullptr, *ndat, *sij_opt, *type_calc, i % nlanes});
    SweepGraph& sg = sweep_graph();
    if (sg.exec && same_keys(sg.keys, keys)) {
      CUDA_CHECK(cudaGraphLaunch(sg.exec, main_stream));
      c.fourwf_counter += 2LL * (*ndat) * (*nk); c.nonlop_counter += (long long)(*ndat) * (*nk); g_kernel_launches += *nk;
      if (!c.async) CUDA_CHECK(cudaStreamSynchronize(main_stream));
      return;
    }
  }
  for (int l = 0; l < nlanes; l++) {
    if (!c.lane_stream[l]) CUDA_CHECK(cudaStreamCreateWithFlags(&c.lane_stream[l], cudaStreamNonBlocking));
    if (!g_lane_ev[l]) CUDA_CHECK(cudaEventCreateWithFlags(&g_lane_ev[l], cudaEventDisableTiming));
  }
  if (!g_lane_ev[kMaxLanes]) CUDA_CHECK(cudaEventCreateWithFlags(&g_lane_ev[kMaxLanes], cudaEventDisableTiming));
  // the lanes start behind whatever the caller queued on the library stream (the blocks may have been produced there)
  CUDA_CHECK(cudaEventRecord(g_lane_ev[kMaxLanes], main_stream));
  for (int l = 0; l < nlanes; l++) CUDA_CHECK(cudaStreamWaitEvent(c.lane_stream[l], g_lane_ev[kMaxLanes], 0));
  c.async = true;
  for (int i = 0; i < *nk; i++) {
    abi_b200_ham* h = hams[i];
    double* gs = gsc ? gsc[i] : nullptr;
    ABI_CHECK(is_device_ptr(cwavef[i]) && is_device_ptr(ghc[i]) && (gs == nullptr || is_device_ptr(gs)),
              "getghc_batch: the wavefunction blocks must be device-resident");
    const int lane = i % nlanes;
    c.lane = lane; c.stream = c.lane_stream[lane];
    int cpopt = -1, prtvol = 0, tim = 0;
    auto eager = [&]() {
      abi_b200_getghc_(&cpopt, cwavef[i], nullptr, ghc[i], gs, &hams[i], nullptr, nullptr, ndat, &prtvol, sij_opt, &tim, type_calc);
    };
    if (!*use_graphs) { eager(); continue; }
    GraphKey key{h, h->epoch, cwavef[i], ghc[i], gs, *ndat, *sij_opt, *type_calc, lane};
    GraphEntry& ge = graph_cache()[key];
    if (ge.state == 0) {                       // first sight: run eagerly (plans, V_loc permutation, workspaces get allocated)
      eager(); ge.state = 1;
    } else if (ge.state == 1) {                // second sight: capture the same call into a graph
      const long long fw0 = c.fourwf_counter, nl0 = c.nonlop_counter, kl0 = g_kernel_launches;
      c.force_scratch_clear = true;
      CUDA_CHECK(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
      eager();
      cudaGraph_t graph = nullptr;
      CUDA_CHECK(cudaStreamEndCapture(c.stream, &graph));
      c.force_scratch_clear = false;
      CUDA_CHECK(cudaGraphInstantiate(&ge.exec, graph, 0));
      ge.graph = graph;                        // kept: child node of the whole-sweep graph
      ge.state = 2;
      // the captured call did not run: undo its bookkeeping, the launch below redoes it
      c.fourwf_counter = fw0; c.nonlop_counter = nl0; g_kernel_launches = kl0;
      ge.state = 2;
      CUDA_CHECK(cudaGraphLaunch(ge.exec, c.stream));
      c.fourwf_counter += 2 * (*ndat); c.nonlop_counter += *ndat; g_kernel_launches += 1;
    } else {
      CUDA_CHECK(cudaGraphLaunch(ge.exec, c.stream));
      c.fourwf_counter += 2 * (*ndat); c.nonlop_counter += *ndat; g_kernel_launches += 1;
    }
  }
  // join: the library stream continues behind every lane
  c.lane = 0; c.stream = main_stream; c.async = was_async;
  for (int l = 0; l < nlanes; l++) {
    CUDA_CHECK(cudaEventRecord(g_lane_ev[l], c.lane_stream[l]));
    CUDA_CHECK(cudaStreamWaitEvent(main_stream, g_lane_ev[l], 0));
  }
  // every call of this batch has its own graph now: compose them (third occurrence of the batch) for the following sweeps
  if (!keys.empty()) {
    bool all = true;
    for (const GraphKey& k : keys) { auto it = graph_cache().find(k); if (it == graph_cache().end() || it->second.state != 2 || !it->second.graph) { all = false; break; } }
    SweepGraph& sg = sweep_graph();
    if (all && !(sg.exec && same_keys(sg.keys, keys))) {
      sweep_graph_drop();
      cudaGraph_t bg = nullptr;
      CUDA_CHECK(cudaGraphCreate(&bg, 0));
      std::vector<cudaGraphNode_t> prev(nlanes, nullptr);
      for (int i = 0; i < *nk; i++) {
        const int lane = i % nlanes;
        cudaGraphNode_t node = nullptr;
        CUDA_CHECK(cudaGraphAddChildGraphNode(&node, bg, prev[lane] ? &prev[lane] : nullptr, prev[lane] ? 1 : 0, graph_cache()[keys[i]].graph));
        prev[lane] = node;
      }
      SweepGraph& ng = sweep_graph();
      CUDA_CHECK(cudaGraphInstantiate(&ng.exec, bg, 0));
      CUDA_CHECK(cudaGraphDestroy(bg));
      ng.keys = keys;
    }
  }
  if (!c.async) CUDA_CHECK(cudaStreamSynchronize(main_stream));
#endif
}

}  // extern "C"
