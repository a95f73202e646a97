// gemm_nonlop on sm_100a (internal header).
// Reference semantics: src/66_nonlocal/m_gemm_nonlop.F90:191-1242, m_gemm_nonlop_projectors.F90:792-1038,
// m_opernla_gemm.F90:361-712, m_opernlc_ylm_allwf.F90:231-1349, m_opernlb_gemm.F90:353-837.
#pragma once
#include "common.cuh"
#include <vector>

namespace abi {

// EXPERIMENTAL int8-sliced copies of P (ozaki.cu, opt-in with ABI_B200_OZAKI=1; never used by default)
struct OzakiP {
  int npw = 0, nprojs = 0, cplx = 0;      // cplx = 1 (istwf_k = 1): (psi, -i psi) column pairs / (P, iP) K pairs
  long long kp1 = 0, kp2 = 0, mp2 = 0;     // padded K of opernla, padded K and M of opernlb
  int8_t* a_k = nullptr;                   // [7][nprojs][kp1]  slices of P columns (K = 2 npw contiguous)
  int8_t* a_m = nullptr;                   // [7][mp2][kp2]     slices of P rows (K = nprojs contiguous)
  double* ea_k = nullptr; double* ea_m = nullptr;   // binary exponents per projector / per row
  bool failed = false;                     // allocation failed: this P stays on the FP64 kernels
  void release();
};
bool ozaki_enabled();
void ozaki_set_enabled(int flag);
void nonlop_set_rag(int mask);   // developer knob: which ragged GEMM variants may be used (nonlop.cu)

// Projectors of one k-point, resident on the device as the real view of P(2, npw, nprojs):
// a column-major (2*npw) x nprojs FP64 matrix (rows = re/im interleaved plane-wave coefficients).
struct Projectors {
  int npw = 0, nprojs = 0, istwf_k = 1;
  double* d_p = nullptr;          // (2*npw) * nprojs doubles (+ padding)
  size_t cap = 0;
  mutable OzakiP oz;              // experimental, built lazily when ABI_B200_OZAKI=1
  mutable unsigned long long oz_stamp = 0, stamp = 1;   // oz is valid when oz_stamp == stamp (stamp bumps on every (re)build of P)
  void alloc(int npw_, int nprojs_, int istwf_k_);
  void release();
};

// Per-type / per-atom description of the non-local operator (flattened gs_hamiltonian_type fields)
struct NonlopAtoms {
  int natom = 0, ntypat = 0, lmnmax = 0, nprojs = 0;
  std::vector<int> indlmn;        // 6 * lmnmax * ntypat (Fortran order)
  std::vector<int> nattyp, atindx1, nlmn;
  // device tables, one entry per projector
  int* d_proj_typ = nullptr;      // type of the projector's atom
  int* d_proj_lmn = nullptr;      // ilmn (0-based) inside the atom
  int* d_proj_atom = nullptr;     // atom index in the type-sorted order (0-based)
  int* d_proj_l = nullptr;        // l quantum number
  int* d_proj_iln = nullptr;      // iln (0-based)
  int* d_atom_first = nullptr;    // natom+1: first projector of each sorted atom
  int* d_atom_typ = nullptr;      // natom
  int* d_atom_enl = nullptr;      // natom: atindx1-1 (column of the D_ij array)
  void build(int natom, int ntypat, int lmnmax, const int* indlmn, const int* nattyp, const int* atindx1);
  void release();
};

struct NonlopEnl {
  int dimenl1 = 0, dimenl2 = 0;
  int nblk = 1;                   // spin blocks of enl(dimenl1, dimenl2, nspinortot**2): 1, or 4 = [up-up, dn-dn, up-dn, dn-up]
  int cplex_enl = 1;              // 2: complex Hermitian D_ij stored as (re, im) pairs of the packed upper triangle
  int sij_dim1 = 0;               // leading dimension of sij (lmn2_size: S_ij is always real)
  int nspinor = 1;                // spinor components per band in the blocks handed to gemm_nonlop (columns = ndat * nspinor)
  double* d_enl = nullptr;        // dimenl1 * dimenl2 * nblk
  double* d_sij = nullptr;        // sij_dim1 * ntypat or null
  size_t enl_cap = 0, sij_cap = 0; // allocated doubles (the buffers are reused while the dimensions do not grow)
  void load(const double* enl, int dimenl1, int dimenl2, const double* sij, int ntypat, cudaStream_t st, int nblk = 1,
            int sij_dim1 = 0);
  void release();
};

// Build P on the device (prep_projectors). ffnl / ph3d are DEVICE pointers in Fortran layout.
void prep_projectors_device(Projectors& P, const NonlopAtoms& at, const double* d_ffnl, int dimffnl,
                            const double* d_ph3d, int matblk, double ucvol, cudaStream_t st);

// Same, with ph3d built in place from the reduced coordinates (ph1d3d): d_kg (3,npw) int, d_xred (3,natom) type-sorted,
// kpt[3] host.
void prep_projectors_xred_device(Projectors& P, const NonlopAtoms& at, const double* d_ffnl, int dimffnl, const int* d_kg,
                                 const double* d_xred, const double* kpt, double ucvol, cudaStream_t st);

// initylmg (optder = 0) for one k-point: d_ylm(npw, mpsang^2) on the device
void initylmg_device(double* d_ylm, int npw, int mpsang, const int* d_kg, const double* kpt, const double* d_gprimd, cudaStream_t st);
// mkffnl (ider = 0, useylm = 1) on the device; all pointers are DEVICE pointers except kpt; d_active(lmnmax*ntypat) = channel mask
void mkffnl_device(double* d_ffnl, int npw, int lmnmax, int ntypat, const int* d_indlmn, const int* d_kg, const double* kpt,
                   const double* d_gprimd, const double* d_ffspl, int mqgrid, int lnmax, double q0, double dq, const double* d_ylm,
                   const unsigned char* d_active, cudaStream_t st);

// getghc fusion of the last GEMM (opernlb): ghc <- (kinpw < filter) ? ghc + P.gxfac : 0 (m_getghc.F90:1266-1280 done in
// the GEMM epilogue; `vectout`, when given, still receives the bare non-local term).  The rows are cut in `nslabs`
// slabs; after each one `after_slab(user, ipw_begin, ipw_end)` runs on the host (used to queue the device->host copy
// of the finished rows behind an event while the next slab is computed).
struct NonlopFusion {
  double* ghc = nullptr;            // (2, npw, ndat): holds V_loc psi + T psi on entry
  const double* kinpw = nullptr;    // npw, device
  double kin_filter = 0.0;
  int nslabs = 1;
  void (*after_slab)(void* user, int ipw_begin, int ipw_end) = nullptr;
  void* user = nullptr;
  mutable bool gsc_filtered = false; // out: the kinetic filter was applied to svectout in the epilogue of its GEMM (one kernel less)
};

// gemm_nonlop, choice in {0,1,7}, signs=2, or choice=1, signs=1 (enlout(ndat) = <psi|Vnl|psi>, opernld).
// All data pointers are DEVICE pointers (may be null when unused):
//   vectin, vectout, svectout : (2, npw, ndat) ; projections : (cplex, nprojs, ndat)
void gemm_nonlop_device(const Projectors& P, const NonlopAtoms& at, const NonlopEnl& enl, int choice, int cpopt,
                        int paw_opt, int me_g0, const double* d_lambda, int ndat, const double* vectin,
                        double* vectout, double* svectout, double* projections, cudaStream_t st,
                        const NonlopFusion* fuse = nullptr, int signs = 2, double* enlout = nullptr);

// opernla / opernlb as stand-alone steps on the padded projection layout gx[ndat][nonlop_ldg(P)] (cplex interleaved):
//   nonlop_project: gx = P^H psi (with the istwf_k>=2 factor 2 / G=0 correction);  nonlop_expand: vectout = P.z (+ add)
long long nonlop_ldg(const Projectors& P);
void nonlop_project(const Projectors& P, int me_g0, const double* vectin, int ndat, double* gx, cudaStream_t st);
void nonlop_expand(const Projectors& P, const double* z, int ndat, double* vectout, const double* add, cudaStream_t st);

void ozaki_prepare(const Projectors& P, OzakiP& oz, cudaStream_t st);
void ozaki_project(const OzakiP& oz, const double* vectin, int ndat, double* part, cudaStream_t st);
void ozaki_expand(const OzakiP& oz, const double* z, long long ldz, int ndat, double* out, int fuse, double* vout, const double* kin,
                  double kin_filter, const double* add, cudaStream_t st, int nslabs = 4, void (*after_slab)(void*, int, int) = nullptr,
                  void* user = nullptr);
void ozaki_release_workspace();

// plain tensor-core GEMMs (also used by the Gram kernels of xg.cu); all device pointers, column-major
//   TN: C(M,N) = alpha * A(K,M)^T B(K,N)      NN: C(M,N) = A(M,K) B(K,N)
void dgemm_tn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C,
              long long ldc, double alpha, cudaStream_t st);
void dgemm_nn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C,
              long long ldc, cudaStream_t st);
// complex flavours on interleaved (re,im) data; M, N, K, lda, ldb, ldc count COMPLEX elements:
//   zgemm_cn: C(M,N) = alpha * A(K,M)^H B(K,N)     zgemm_nn: C(M,N) = A(M,K) B(K,N)
// (same real DMMA kernels: the -i psi / i P operands are formed at fragment-load time, see nonlop.cu)
void zgemm_cn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C,
              long long ldc, double alpha, cudaStream_t st);
void zgemm_nn(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C,
              long long ldc, cudaStream_t st);

}  // namespace abi
