/* abinit_b200 -- B200-native (sm_100a) drop-in for ABINIT's getghc hot path: fourwf + gemm_nonlop + assembly.
 *
 * C-ABI boundary.  Every entry point below is what the reference's Fortran would bind through
 * iso_c_binding for this path; the citation next to each one names the reference interface it replaces
 * (paths relative to the ABINIT source tree).  Conventions follow the reference's own C/CUDA plug points
 * (src/46_manage_cuda/gpu_fourwf.cu:133-156): Fortran array layouts (column-major, re/im interleaved
 * real(dp)(2,*) complex data, kg(3,npw) integer triplets), no torch or C++ types in any signature.
 *
 * Pointer residency: every data pointer may be a HOST pointer or a DEVICE pointer (cudaMalloc'ed on the
 * current device).  The library inspects it (cudaPointerGetAttributes) and only stages host arrays --
 * the same contract as the reference's offload path, which skips transfers for arrays the caller already
 * mapped (src/66_wfs/m_getghc.F90:378-394, src/46_ghc_omp/m_ompgpu_fourwf.F90:247-262).
 *
 * Error convention: none of these functions returns a status.  Invalid input or a CUDA failure prints a
 * YAML-style "--- !ERROR" document to stderr and calls abort(), like ABI_ERROR -> abi_abort
 * (shared/common/src/incs/abi_common.h:307-309) and abi_cabort() (gpu_fourwf.cu:196-207).
 * There is no CPU fallback anywhere in the library.
 *
 * Threading: one host thread per GPU drives the library (src/66_wfs/m_getghc.F90:2442-2449).
 * Calls are synchronous from the caller's point of view unless abi_b200_set_async(1) was called.
 */
#ifndef ABINIT_B200_H
#define ABINIT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------
 * Library lifecycle.  Replaces alloc_hamilt_gpu / dealloc_hamilt_gpu
 * (src/66_nonlocal/m_alloc_hamilt_gpu.F90:96,237) and the device pick of
 * shared/common/src/17_gpu_toolbox/m_initcuda.F90:326-333 (device = mod(rank, ndevices)).
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_init(int rank);                 /* selects device rank % ndevices, creates the stream */
void abi_b200_finalize(void);                 /* frees plans, tables, workspaces */
void abi_b200_set_stream(void* cuda_stream);  /* run on a caller-owned cudaStream_t (NULL -> internal stream) */
void abi_b200_set_async(int flag);            /* 1: do not synchronise before returning (device pointers only) */
void abi_b200_synchronize(void);              /* gpu_device_synchronize, src/79_seqpar_mpi/m_chebfiwf.F90:377-379 */
long long abi_b200_kernel_launches(void);     /* kernels launched by this library so far (bench accounting) */
const char* abi_b200_version(void);
/* Per-kernel-class device timers (CUDA events on the library stream), the measurement twin of the reference's
 * timab slots 840+ (fourwf) / 220+ (nonlop) (shared/common/src/18_timing/m_time.F90:746).  collect() synchronises
 * and returns the number of classes; names are ';'-separated, ms are summed durations, counts are launches. */
void abi_b200_profile_enable(int on);
int abi_b200_profile_collect(char* names, int names_cap, double* ms, long long* counts, int cap);
/* FP64 pipe peak of the current device, measured now (register-resident DFMA and DMMA m8n8k4 loops, TFLOP/s): the roofline
 * denominator bench.py reports is measured in the run it belongs to. */
void abi_b200_probe_fp64_peak(double* dfma_tflops, double* dmma_tflops);

/* ------------------------------------------------------------------------------------------------------
 * fourwf.  Same symbol shape as the legacy CUDA plug point
 *   extern "C" void gpu_fourwf_(...)            src/46_manage_cuda/gpu_fourwf.cu:133-156
 * called from fourwf (src/53_ffts/m_fft.F90:2383-2387): all scalars by reference, weight_r/weight_i arrays
 * of ndat, trailing underscore.  Supports option 0,1,2,3; cplex 1,2 (option 2); istwf_k 1..9; ndat>=1.
 * Requires n4,n5,n6 == n1,n2,n3 (as every GPU path of the reference, gpu_fourwf.cu:196-201).
 * mpi_enreg is opaque and ignored except through me_g0 (set with abi_b200_set_me_g0, default 1).
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_fourwf_(int* cplex, double* denpot, double* fofgin, double* fofgout, double* fofr,
                      int* gboundin, int* gboundout, int* istwf_k, int* kg_kin, int* kg_kout, int* mgfft,
                      void* mpi_enreg, int* ndat, int* ngfft, int* npwin, int* npwout, int* n4, int* n5, int* n6,
                      int* option, int* paral_kgb, int* tim_fourwf, double* weight_r, double* weight_i);
/* alloc_gpu_fourwf_ / free_gpu_fourwf_ (gpu_fourwf.cu:290,350): pre-size / release the work buffers */
void abi_b200_alloc_fourwf_(int* ngfft, int* ndat, int* npwin, int* npwout);
void abi_b200_free_fourwf_(void);
/* The legacy symbol names themselves, so that an ABINIT built with the legacy CUDA plug point can link this
 * library in place of src/46_manage_cuda/gpu_fourwf.cu without touching m_fft.F90 (identical signatures). */
void gpu_fourwf_(int* cplex, double* denpot, double* fofgin, double* fofgout, double* fofr, int* gboundin,
                 int* gboundout, int* istwf_k, int* kg_kin, int* kg_kout, int* mgfft, void* mpi_enreg, int* ndat,
                 int* ngfft, int* npwin, int* npwout, int* n4, int* n5, int* n6, int* option, int* paral_kgb,
                 int* tim_fourwf, double* weight_r, double* weight_i);
void alloc_gpu_fourwf_(int* ngfft, int* ndat, int* npwin, int* npwout);
void free_gpu_fourwf_(void);
void abi_b200_set_me_g0(int me_g0);           /* mpi_enreg%me_g0_fft (src/53_ffts/m_fft.F90:2346) */
/* force the generic (full-box) or the fused implementation of option 2: 0 auto, 1 generic, 2 fused */
void abi_b200_fourwf_set_impl(int impl);
/* tuning knobs of the fused path (no reference counterpart; also read from ABI_B200_FOURWF_* at first use):
 * "plane" 0/1, "plane_cfg" 0 auto / 1 (8 columns x 4 warps) / 2 (4 columns x 8 warps) / 3 (4 x 16), "plane_ctas_per_sm",
 * "pack2" 0/1 (istwf_k=2: two bands per complex transform),
 * "cluster", "lines_x", "smem_kb_mid", "band_chunk", "pipe_chunks" (host-array getghc pipeline depth),
 * "nonlop_ozaki" 0/1 (EXPERIMENTAL, default 0: gemm_nonlop's real contractions through exact int8 slice products,
 * csrc/ozaki.cu; also ABI_B200_OZAKI=1).  Unknown names abort. */
void abi_b200_fourwf_set_tuning(const char* name, int value);
/* fourwf_counter of src/53_ffts/m_fft.F90:2333-2336 */
long long abi_b200_fourwf_counter(void);

/* ------------------------------------------------------------------------------------------------------
 * gemm_nonlop.  Projector lifecycle mirrors init_gemm_nonlop / set_gemm_nonlop_ikpt / prep_projectors /
 * destroy_gemm_nonlop (src/66_nonlocal/m_gemm_nonlop_projectors.F90:199-253,555-575,792-1038).
 * The apply call has the argument list of gemm_nonlop_gpu (src/66_nonlocal/m_gemm_nonlop_gpu.F90:134,
 * called at src/66_nonlocal/m_nonlop.F90:800-808) restricted to what choice in {0,1,7}, signs=2 reads, with
 * cprjin flattened to the `vectproj` buffer projections(cplex, nprojs, nspinor*ndat)
 * (src/66_nonlocal/m_gemm_nonlop.F90:503-508); the Fortran shim does the pawcprj pack/unpack (:719-789).
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_init_gemm_nonlop_(int* nkpt);
void abi_b200_destroy_gemm_nonlop_(void);
/* Build and cache P for k-point slot *ikpt (1-based).  ffnl(npw,dimffnl,lmnmax,ntypat), ph3d(2,npw,matblk)
 * with atoms sorted by type, indlmn(6,lmnmax,ntypat), nattyp(ntypat).  P stays on the device. */
void abi_b200_prep_projectors_(int* ikpt, int* npw, int* lmnmax, int* ntypat, int* indlmn, int* nattyp,
                               int* istwf_k, double* ucvol, double* ffnl, double* ph3d, int* dimffnl,
                               int* matblk);
/* Test/benchmark hook: install an explicit P(2,npw,nprojs) instead of building it from ffnl/ph3d. */
void abi_b200_set_projectors_(int* ikpt, int* npw, int* nprojs, int* istwf_k, double* projs);
void abi_b200_set_gemm_nonlop_ikpt_(int* ikpt);
/* mkffnl (src/66_nonlocal/m_mkffnl.F90:238, same argument list, optional arguments dropped) restricted to what the getghc
 * path needs: ider=0, idir=0, dimffnl=1, useylm=1: ffnl(npw,1,lmnmax,ntypat) = ylm * splfit(ffspl) at |k+G|, computed on the
 * device.  ffnl / ffspl / ylm / kg may be host or device arrays (a device ffnl can be handed to abi_b200_ham_load_k or
 * abi_b200_prep_projectors_ as is); indlmn, qgrid (uniform), kpt, gprimd, ekb, pspso are host arrays. */
/* initylmg (src/56_recipspace/m_initylmg.F90:94) for ONE k-point, optder = 0: ylm(npw, mpsang^2), host or device. */
void abi_b200_initylmg_k_(double* gprimd, int* kg, double* kpt, int* mpsang, int* npw, double* ylm);
void abi_b200_mkffnl_(int* dimekb, int* dimffnl, double* ekb, double* ffnl, double* ffspl, double* gmet, double* gprimd,
                      int* ider, int* idir, int* indlmn, int* kg, double* kpg, double* kpt, int* lmnmax, int* lnmax,
                      int* mpsang, int* mqgrid, int* nkpg, int* npw, int* ntypat, int* pspso, double* qgrid, double* rmet,
                      int* usepaw, int* useylm, double* ylm, double* ylm_gr);
void abi_b200_gemm_nonlop_(int* atindx1, int* choice, int* cpopt, double* vectproj, int* dimenl1, int* dimenl2,
                           int* dimekbq, double* enl, int* indlmn, int* istwf_k, double* lambda, int* lmnmax,
                           int* natom, int* nattyp, int* ndat, int* nnlout, int* npwin, int* npwout,
                           int* nspinor, int* nspinortot, int* ntypat, int* paw_opt, double* sij,
                           double* svectout, int* useylm, double* vectin, double* vectout, int* signs);
long long abi_b200_nonlop_counter(void);      /* nonlop_counter, src/66_nonlocal/m_nonlop.F90:389-392 */

/* ------------------------------------------------------------------------------------------------------
 * getghc (fused fast path).  gs_hamiltonian_type (src/66_nonlocal/m_hamiltonian.F90:99-467) is flattened
 * into an opaque handle filled by the same three steps the reference performs:
 *   init      -> abi_b200_ham_create      (gs_hamk%init,      src/79_seqpar_mpi/m_vtorho.F90:640)
 *   load_spin -> abi_b200_ham_load_spin   (vlocal per spin,   m_vtorho.F90:804)
 *   load_k    -> abi_b200_ham_load_k      (kg,kinpw,ffnl,ph3d m_vtorho.F90:1035-1045; P built here)
 * abi_b200_getghc_ has the argument list of getghc (src/66_wfs/m_getghc.F90:182-202) with gs_ham -> handle,
 * cwaveprj -> flattened projections, mpi_enreg dropped.  k==k' (select_k default); nspinor=2 / nvloc=4: NC only (below).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct abi_b200_ham abi_b200_ham_t;
abi_b200_ham_t* abi_b200_ham_create(const int* ngfft, int natom, int ntypat, int lmnmax, const int* indlmn,
                                    const int* nattyp, const int* atindx1, int usepaw, double ucvol);
void abi_b200_ham_destroy(abi_b200_ham_t* h);
void abi_b200_ham_load_spin(abi_b200_ham_t* h, const double* vlocal, int cplex_vloc, int n4, int n5, int n6);
/* nspinor = 2 (gs_hamk%nspinor, norm-conserving, istwf_k = 1): blocks are cwavef(2, npw*nspinor*ndat); ndat keeps counting bands.
 * load_spin_nvloc: vlocal(n4,n5,n6,nvloc), nvloc = 4 = [V11, V22, Re V12, Im V12] (non-collinear magnetism): the four local
 * applications of src/66_wfs/m_getghc.F90:655-830.  PAW spinors take their D_ij through abi_b200_ham_load_enl_spinor. */
void abi_b200_ham_set_nspinor(abi_b200_ham_t* h, int nspinor);
void abi_b200_ham_load_spin_nvloc(abi_b200_ham_t* h, const double* vlocal, int nvloc, int n4, int n5, int n6);
/* enl: NC ekb(dimenl1=lnmax, ntypat); PAW dij(dimenl1=cplex_dij*lmn2_size, natom): real packed symmetric (cplex_dij = 1) or
 * complex Hermitian, (re, im) pairs of the packed upper triangle (cplex_dij = 2, istwf_k = 1 only;
 * src/66_nonlocal/m_opernlc_ylm_allwf.F90:453-576).  sij(lmn2_size, ntypat) (PAW, always real) or NULL.
 * load_enl_spinor: enl(dimenl1, dimenl2, nspinortot**2) with the four blocks [up-up, dn-dn, up-dn, dn-up] of a spinor
 * Hamiltonian (gs_hamk%ekb, nspinor = 2 with PAW; off-diagonal blocks :660-737); nspinortot2 = 1 is load_enl. */
void abi_b200_ham_load_enl(abi_b200_ham_t* h, const double* enl, int dimenl1, int dimenl2, const double* sij);
void abi_b200_ham_load_enl_spinor(abi_b200_ham_t* h, const double* enl, int dimenl1, int dimenl2, int nspinortot2, const double* sij);
void abi_b200_ham_load_k(abi_b200_ham_t* h, int istwf_k, int npw, const int* kg_k, const double* kinpw,
                         const double* ffnl, int dimffnl, const double* ph3d, int matblk, int me_g0);
/* load_k with the structure-factor phases built on the device: ph3d(G,ia) = exp(2 pi i (k+G).xred_ia) is what load_k computes
 * with ph1d3d when compute_ph3d is set (src/66_nonlocal/m_hamiltonian.F90:1132-1140, src/56_recipspace/m_kg.F90:644-700);
 * here it is fused into prep_projectors, so the npw x natom phase array is neither built on the host nor uploaded.
 * kpt(3); xred(3,natom) with atoms sorted by type (the order of ph3d's columns). */
void abi_b200_ham_load_k_xred(abi_b200_ham_t* h, int istwf_k, int npw, const int* kg_k, const double* kinpw,
                              const double* ffnl, int dimffnl, const double* kpt, const double* xred, int me_g0);
/* benchmark/test hook: explicit projectors instead of ffnl/ph3d (pass ffnl=ph3d=NULL to load_k) */
void abi_b200_ham_set_projectors(abi_b200_ham_t* h, const double* projs, int nprojs);
int abi_b200_ham_nprojs(const abi_b200_ham_t* h);
void abi_b200_getghc_(int* cpopt, double* cwavef, double* cwaveprj, double* ghc, double* gsc,
                      abi_b200_ham_t** gs_ham, double* gvnlxc, double* lambda, int* ndat, int* prtvol,
                      int* sij_opt, int* tim_getghc, int* type_calc);
/* Batched getghc for the many-small-k-points regime (the (k, spin) loop of src/79_seqpar_mpi/m_vtorho.F90:789-1045 around the
 * eigensolver's getghc calls): nk independent applications, call i on handle hams[i] (one handle per (k, spin) pair: its own
 * load_spin / load_k) with DEVICE-resident blocks cwavef[i] -> ghc[i] (, gsc[i]; gsc may be NULL), cpopt = -1, no gvnlxc / lambda.
 * The calls are dealt to concurrent streams with private workspaces and, when *use_graphs != 0, replayed from CUDA graphs
 * captured on their second occurrence (dropped when the handle is reloaded).  Results are identical to nk calls of
 * abi_b200_getghc_.  abi_b200_graphs_clear drops every captured graph (call it before freeing the arrays they refer to). */
void abi_b200_getghc_batch_(int* nk, abi_b200_ham_t** hams, double** cwavef, double** ghc, double** gsc, int* ndat,
                            int* sij_opt, int* type_calc, int* use_graphs);
void abi_b200_graphs_clear(void);

/* ------------------------------------------------------------------------------------------------------
 * nonlop dispatcher on the Hamiltonian handle (src/66_nonlocal/m_nonlop.F90:336-976, gemm_nonlop route :782-811):
 * argument list of nonlop(choice,cpopt,cprjin,enlout,hamk,idir,lambda,mpi_enreg,ndat,nnlout,paw_opt,signs,svectout,
 * tim_nonlop,vectin,vectout) with mpi_enreg dropped and cprjin flattened to projections(cplex,nprojs,ndat).
 * signs=2: choice 0/1/7 as abi_b200_gemm_nonlop_.  signs=1, choice=1: enlout(ndat) = <psi|Vnl|psi>
 * (opernld, src/66_nonlocal/m_opernld_ylm_allwf.F90:160-203; the call of src/79_seqpar_mpi/m_chebfiwf.F90:296-297).
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_nonlop_(int* choice, int* cpopt, double* cprjin, double* enlout, abi_b200_ham_t** hamk, int* idir,
                      double* lambda, int* ndat, int* nnlout, int* paw_opt, int* signs, double* svectout,
                      int* tim_nonlop, double* vectin, double* vectout);

/* ------------------------------------------------------------------------------------------------------
 * xgBlock algebra of the eigensolvers (src/45_xgTools/m_xg.F90).  Blocks are DEVICE arrays with the memory layout
 * of cg(2, npw*nband) (xgBlock_map, m_xg.F90:716-781): `rows` complex coefficients per column, leading dimensions
 * in complex elements.  space: 1 SPACE_R, 2 SPACE_C, 3 SPACE_CR (m_xg.F90:63-65).  Sub-space matrices (Gram matrices,
 * eigenvectors) are real for SPACE_R/SPACE_CR and complex for SPACE_C; their leading dimension counts their elements.
 *
 * xg_gram     : W(ncols_a,ncols_b) = A^H B, xgBlock_gemm('t','n') (m_xg.F90:1674-1976); SPACE_CR: 2 A^T B on the real
 *               view minus the doubled G=0 row when me_g0=1 (:1802-1882).  Row-sharded callers sum W over ranks
 *               (xgBlock_mpi_sum :3636-3663 -> NCCL allreduce in abinit_b200.parallel).
 * xg_rotate   : X(:,1:ncols_out) <- X(:,1:k) C(1:k,1:ncols_out), xgBlock_gemm('n','n') + xgBlock_copy
 *               (m_xg_ortho_RR.F90:524-531); ldc must be even for the real spaces, pad row zero when k is odd.
 * xg_hegvd    : xgBlock_hegvd(1,'v','u') / xgBlock_heevd('v','u') when b is NULL (m_xg.F90:2239-2861); eigenvectors
 *               overwrite a, w(n) on the device.
 * xg_colwise  : op 0 colwiseDotProduct (out(ncols), (re,im) pairs for SPACE_C) :4550-4846; 1 colwiseNorm2 :4341-4542;
 *               2 colwiseCymax a = w - da(col) b :3301-3413; 3 per-column scale a(:,j) *= da(j) (chebfi_ampfactor);
 *               4 zero_im_g0 :5851-5898; 5 xgBlock_add a += b; 6 xgBlock_apply_diag a(i,:) *= da(i) (da per row: the LOBPCG
 *               preconditioner).
 * xg_rayleigh_ritz : xg_RayleighRitz, VAR_X branch (m_xg_ortho_RR.F90:251-571); bx NULL -> overlap block is x
 *               (norm-conserving); eigenvalues may be a host or device array.
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_xg_gram_(int* space, int* rows, int* ncols_a, int* ncols_b, double* a, int* lda, double* b, int* ldb,
                       double* c, int* ldc, int* me_g0);
void abi_b200_xg_rotate_(int* space, int* rows, int* k, int* ncols_out, double* x, int* ldx, double* c, int* ldc);
void abi_b200_xg_hegvd_(int* space, int* n, double* a, int* lda, double* b, int* ldb, double* w, int* info);
/* out(:,1:ncols_out) = a(:,1:k) c(1:k,1:ncols_out) (xgBlock_gemm 'n','n'), out may alias a; upper=1: c upper triangular
 * (the trsm of xg_Borthonormalize, m_xg_ortho_RR.F90:135-142, with c = U^-1).  xg_chol_inverse: potrf 'u' + inverse of the
 * factor on the m x m sub-space matrix a (xgBlock_potrf, m_xg_ortho_RR.F90:125), strictly-lower triangle zeroed. */
void abi_b200_xg_gemm_nn_(int* space, int* rows, int* k, int* ncols_out, double* a, int* lda, double* c, int* ldc,
                          double* out, int* ldo, int* upper);
void abi_b200_xg_chol_inverse_(int* space, int* m, double* a, int* lda, int* info);
void abi_b200_xg_colwise_(int* op, int* space, int* rows, int* ncols, double* a, int* lda, double* b, int* ldb,
                          double* w, int* ldw, double* da, double* out, int* me_g0);
void abi_b200_xg_rayleigh_ritz_(int* space, int* rows, int* blockdim, double* x, int* ldx, double* ax, int* ldax,
                                double* bx, int* ldbx, double* eigenvalues, int* info, int* solve_ax_bx, int* me_g0);

/* ------------------------------------------------------------------------------------------------------
 * ChebFi2 (src/48_diago/m_chebfi2.F90:466-735 chebfi_run) driven as chebfiwf2 drives it
 * (src/79_seqpar_mpi/m_chebfiwf.F90:110-330): getAX_BX is bound to the fused getghc on band blocks of `bandpp`
 * columns (getghc_gsc1 :341-385), lambda_plus = ecut, the dtset scalars are passed flattened
 * (tolwfr_diago, ecut, nline, nbdbuf, chebfi_oracle, oracle_factor, oracle_min_occ).
 * cg(2,npw*nband) host or device, in/out; eig, resid, occ (may be NULL when chebfi_oracle=0), enl_out (NC only, may
 * be NULL): host arrays of nband.  PAW: BX = S X from getghc(sij_opt=1), getBm1X = abi_b200_apply_invovl_.
 * The three phase functions are what a band-parallel caller strings together around its collectives
 * (max/min of the Rayleigh quotients, m_chebfi2.F90:606-611; transposition + Rayleigh-Ritz, :687-705):
 *   chebfi_rq   : AX,BX = getAX_BX(X); div(ncols) = <X|AX>/<X|BX>; max / min      (:575-613)
 *   chebfi_core : the filter loop of degree ndeg_filter + chebfi_ampfactor           (:634-677)
 *                 x, x_next, x_prev are rotated like chebfi_swapInnerBuffers: on return *x holds the filtered block.
 * cheb_oracle1 / cheb_poly1: m_chebfi2.F90:1031-1064, 1084-1106.
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_chebfiwf2_(double* cg, double* eig, double* occ, double* enl_out, abi_b200_ham_t** gs_hamk, int* nband,
                         int* npw, int* nspinor, int* prtvol, double* resid, double* tolwfr_diago, double* ecut,
                         int* nline, int* nbdbuf, int* chebfi_oracle, double* oracle_factor, double* oracle_min_occ,
                         int* bandpp);
void abi_b200_chebfi_rq_(abi_b200_ham_t** gs_hamk, int* ncols, int* bandpp, double* x, double* ax, double* bx,
                         double* div, double* maxeig, double* mineig);
void abi_b200_chebfi_core_(abi_b200_ham_t** gs_hamk, int* ncols, int* bandpp, double** x, double* ax, double* bx,
                           double** x_next, double** x_prev, double* lambda_minus, double* lambda_plus,
                           int* ndeg_filter, double* div);
/* ------------------------------------------------------------------------------------------------------
 * Band-parallel ChebFi2 inside the library: chebfi_run with paral_kgb = 1, npband = the ranks of the library communicator
 * (src/48_diago/m_chebfi2.F90:466-735 with xmpi_max/xmpi_min of the Rayleigh quotients :606-611, xgTransposer
 * src/45_xgTools/m_xgTransposer.F90:640-900 (TRANS_ALL2ALL) :687-689, xgBlock_gemm(..., comm=) -> xgBlock_mpi_sum
 * src/45_xgTools/m_xg.F90:1969-1974, 3636-3663).  One process per GPU; NCCL (dlopen'ed libnccl.so.2) carries the three collectives
 * on the library stream, so a Fortran caller reaches the multi-GPU solver without Python:
 *   comm_get_unique_id : rank 0 obtains the 128-byte ncclUniqueId and broadcasts it with the application's own MPI (xmpi_bcast)
 *   comm_init_rank     : every rank joins (ncclCommInitRank); nranks = 1 needs no id and no NCCL
 *   comm_adopt         : alternative: hand in an ncclComm_t the application already owns (not destroyed by the library)
 *   xg_transpose       : xgTransposer_transpose on device blocks: to_rows = 1: cols(2, rows, my_ncols) -> lin(2, my_nrows, nband),
 *                        to_rows = 0 the inverse.  Bands and rows are distributed in contiguous blocks whose sizes differ by at
 *                        most one, the larger blocks on the lower ranks.
 *   chebfiwf2_paral    : the argument list of abi_b200_chebfiwf2_ with nband = ALL bands and ncols_mine = this rank's:
 *                        cg(2, npw*nspinor*ncols_mine): the rank's band block (host or device, in/out); eig(nband): all
 *                        eigenvalues (host, replicated); occ(nband): all occupations (host, may be NULL when chebfi_oracle = 0);
 *                        enl_out(ncols_mine) (NC only, may be NULL) and resid(ncols_mine): the rank's bands (host).
 *                        chebfi_oracle 1 / 2: chebfi_set_ndeg_from_residu on the rank's bands (shift = its first band,
 *                        m_chebfi2.F90:1161) and the MAX of the degrees over the ranks (:1203).  NC and PAW.
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_comm_get_unique_id_(char* id128);
void abi_b200_comm_init_rank_(const char* id128, int* nranks, int* rank);
void abi_b200_comm_adopt_(void* nccl_comm, int* nranks, int* rank);
void abi_b200_comm_destroy_(void);
void abi_b200_xg_transpose_(int* to_rows, double* cols, double* lin, int* rows, int* nband);
void abi_b200_chebfiwf2_paral_(double* cg, double* eig, double* occ, double* enl_out, double* resid, abi_b200_ham_t** gs_hamk,
                               int* nband, int* ncols_mine, int* npw, int* nspinor, double* tolwfr_diago, double* ecut, int* nline,
                               int* nbdbuf, int* chebfi_oracle, double* oracle_factor, double* oracle_min_occ, int* bandpp);
/* lobpcg_run with paral_kgb = 1 on the same communicator (src/48_diago/m_lobpcg2.F90:340-765, one block of all bands, NC and PAW):
 * getAX_BX on the rank's band block, B-orthonormalisation / Rayleigh-Ritz / residuals / preconditioner on the rank's plane-wave
 * rows, Gram and residual sums over NCCL, one xgTransposer exchange out and one (PAW: two) back per iteration.
 * cg: this rank's band block (in/out); eig(nband), resid(nband): host, replicated. */
void abi_b200_lobpcgwf2_paral_(double* cg, double* eig, double* resid, abi_b200_ham_t** gs_hamk, int* nband, int* ncols_mine,
                               int* npw, int* nspinor, double* tolwfr_diago, int* nline, int* bandpp);
/* LOBPCG (src/48_diago/m_lobpcg2.F90:340-765 lobpcg_run) driven as lobpcgwf2 drives it (src/79_seqpar_mpi/m_lobpcgwf.F90:
 * 100-250): getAX_BX = fused getghc, preconditioner = build_pcon(kinpw) (:316-334), xg_Borthonormalize (Cholesky) and the
 * X / XW / XWP Rayleigh-Ritz of src/45_xgTools/m_xg_ortho_RR.F90:86-150, 251-571.  nblock_lobpcg blocks of
 * blockdim = nband / nblock_lobpcg bands (m_lobpcgwf.F90:133; lobpcg_orthoXwrtBlocks m_lobpcg2.F90:803-840 against the
 * finished blocks, final xg_Borthonormalize + xg_RayleighRitz over all bands :744-751); paral_kgb = 0 in this entry (the
 * band-parallel scheme is abi_b200_lobpcgwf2_paral_ above); dtset scalars flattened (tolwfr_diago, nline,
 * nblock_lobpcg, nbdbuf).  nspinor = 2 (NC): blocks have npw*nspinor rows, nspinor must match abi_b200_ham_set_nspinor.
 * cg in/out (host or device); eig, resid, occ (used when nbdbuf = -101), enl_out (NC, may be NULL): host arrays. */
void abi_b200_lobpcgwf2_(double* cg, double* eig, double* occ, double* enl_out, abi_b200_ham_t** gs_hamk, int* nband,
                         int* npw, int* nspinor, int* prtvol, double* resid, double* tolwfr_diago, int* nline,
                         int* nblock_lobpcg, int* nbdbuf, int* bandpp);
int abi_b200_cheb_oracle1_(double* xx, double* aa, double* bb, double* tol, int* nmax);
double abi_b200_cheb_poly1_(double* xx, int* nn, double* aa, double* bb);

/* ------------------------------------------------------------------------------------------------------
 * PAW inverse overlap (src/66_wfs/m_invovl.F90): S^-1 = 1 - P (s^-1 + P^H P)^-1 P^H, the getBm1X of ChebFi2-PAW
 * (src/79_seqpar_mpi/m_chebfiwf.F90:390-440).
 *   make_invovl  (m_invovl.F90:469-776): builds inv_sij, inv_s_approx and gram_projs = P^H P for the k-point loaded in
 *                the handle (from the resident projectors; ffnl/ph3d were consumed by load_k).  Called lazily by
 *                apply_invovl; load_k / load_enl / set_projectors invalidate it.
 *   apply_invovl (m_invovl.F90:790-1039): sm1cwavef = S^-1 cwavef, argument list of the reference with mpi_enreg dropped
 *                and cwaveprj flattened to (cplex,nprojs,ndat) (may be NULL); block_sliced is accepted and ignored (both
 *                settings give the same numbers, it only selects a BLAS call pattern, :1165-1231).
 * ---------------------------------------------------------------------------------------------------- */
void abi_b200_make_invovl_(abi_b200_ham_t** ham);
void abi_b200_apply_invovl_(abi_b200_ham_t** ham, double* cwavef, double* sm1cwavef, double* cwaveprj, int* npw,
                            int* ndat, int* nspinor, int* block_sliced);

#ifdef __cplusplus
}
#endif
#endif /* ABINIT_B200_H */
