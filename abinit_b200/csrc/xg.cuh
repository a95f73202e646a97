// xgBlock linear algebra on sm_100a for the eigensolver side of getghc (internal header).
// Reference semantics: src/45_xgTools/m_xg.F90 (xgBlock_gemm :1674-1976, colwise* :3301-5066, zero_im_g0 :5851-5898,
// heevd/hegvd :2239-2861), src/45_xgTools/m_xg_ortho_RR.F90:251-571 (xg_RayleighRitz).
//
// A block is the memory of cg(2, npw*nband): column-major, `rows` COMPLEX coefficients per column, leading dimension
// `ld` in complex elements.  space: 1 SPACE_R (plain real, rows real numbers), 2 SPACE_C (complex), 3 SPACE_CR (complex
// storage of istwf_k>=2 data handled as 2*rows reals, with the G=0 conventions of the reference when me_g0 = 1).
#pragma once
#include "common.cuh"

namespace abi {

constexpr int SPACE_R = 1, SPACE_C = 2, SPACE_CR = 3;   // m_xg.F90:63-65

// element size of a sub-space matrix (Gram matrices, eigenvectors): real for SPACE_R/SPACE_CR, complex for SPACE_C
inline int sub_cplex(int space) { return space == SPACE_C ? 2 : 1; }

// W(ncols_a, ncols_b) = alpha * A^H B  (xgBlock_gemm 't','n').  ldw counts sub-space elements.
void xg_gram(int space, int rows, int ncols_a, int ncols_b, const double* A, long long lda, const double* B, long long ldb,
             double* W, long long ldw, int me_g0, cudaStream_t st);
// X(:, 0:ncols_out) <- X(:, 0:k) . C(0:k, 0:ncols_out)   (xgBlock_gemm 'n','n' + xgBlock_copy), in place by row slabs.
// C must be K-padded: ldc even (real spaces) and the pad row zero when k is odd.
void xg_rotate(int space, int rows, int k, int ncols_out, double* X, long long ldx, const double* C, long long ldc,
               cudaStream_t st);
// OUT(:, 0:ncols_out) = A(:, 0:k) . C(0:k, 0:ncols_out) by row slabs; OUT may be any block, including columns of A itself
// (each slab's product is complete before it is stored).  xg_rotate is the case OUT == A.
void xg_gemm_nn(int space, int rows, int k, int ncols_out, const double* A, long long lda, const double* C, long long ldc,
                double* OUT, long long ldo, cudaStream_t st);
// same with an upper-triangular C (inverse Cholesky factor): triangular K ranges, in place allowed
void xg_gemm_nn_upper(int space, int rows, int k, int ncols_out, const double* A, long long lda, const double* C, long long ldc,
                      double* OUT, long long ldo, cudaStream_t st);
// A(m x m, upper triangle of a Hermitian positive matrix) -> U^-1 with A = U^H U (potrf 'u' + trtri, strictly-lower part zeroed);
// sub_space: SPACE_R or SPACE_C.  Returns potrf's info.
int xg_chol_inverse(int sub_space, int m, double* A, long long lda, cudaStream_t st);
// lobpcg_orthoXwrtBlocks (src/48_diago/m_lobpcg2.F90:803-840): V(:, 0:n) -= X0(:, 0:nprev) . (BX0^H V), in place
void xg_ortho_wrt_blocks(int space, int rows, int nprev, int n, double* V, long long ldv, const double* X0, long long ldx0,
                         const double* BX0, long long ldbx0, int me_g0, cudaStream_t st);
// X += P (xgBlock_add)
void xg_add(int space, int rows, int ncols, double* X, long long ldx, const double* P, long long ldp, cudaStream_t st);
// X(i, j) *= d(i), d real per (complex) row (xgBlock_apply_diag with a SPACE_R diagonal: the LOBPCG preconditioner)
void xg_apply_diag(int space, int rows, int ncols, double* X, long long ldx, const double* d, cudaStream_t st);
// xg_Borthonormalize (m_xg_ortho_RR.F90:86-150): X^H BX = U^H U (potrf 'u'), X, BX, AX <- . U^-1.  Returns potrf's info.
int xg_b_orthonormalize(int space, int rows, int m, double* X, long long ldx, double* BX, long long ldbx, double* AX, long long ldax,
                        int me_g0, cudaStream_t st);
// xg_RayleighRitz, VAR_XW (nvar = 2) / VAR_XWP (nvar = 3) branches (m_xg_ortho_RR.F90:300-571) on contiguous blocks
// [X | W | P] of n columns each: X, AX, BX, P, AP, BP updated; eig: DEVICE array of nvar*n eigenvalues.
int xg_rayleigh_ritz_xwp(int space, int rows, int n, int nvar, double* XWP, double* AXWP, double* BXWP, long long ld, double* eig,
                         int me_g0, cudaStream_t st);
// xgBlock_zero_im_g0
void xg_zero_im_g0(int space, int ncols, double* X, long long ldx, int me_g0, cudaStream_t st);
// dots(ncols) = colwise <A|B> with the SPACE_CR conventions (xgBlock_colwiseDotProduct); SPACE_C stores (re, im) pairs
void xg_colwise_dot(int space, int rows, int ncols, const double* A, long long lda, const double* B, long long ldb,
                    double* dots, int me_g0, cudaStream_t st);
// norms(ncols) = colwise |A|^2 (xgBlock_colwiseNorm2; always real)
void xg_colwise_norm2(int space, int rows, int ncols, const double* A, long long lda, double* norms, int me_g0,
                      cudaStream_t st);
// A <- W - da(col) * B (xgBlock_colwiseCymax); A may alias W
void xg_colwise_cymax(int space, int rows, int ncols, double* A, long long lda, const double* da, const double* B,
                      long long ldb, const double* W, long long ldw, cudaStream_t st);
// X(:,j) *= s(j)  (xgBlock_scale per column, chebfi_ampfactor); s is a DEVICE array
void xg_scale_cols(int space, int rows, int ncols, double* X, long long ldx, const double* s, cudaStream_t st);
// one Chebyshev recurrence step (chebfi_computeNextOrderChebfiPolynom, m_chebfi2.F90:837-896):
//   Xnext = scale * (AX - center * X) - (Xprev ? Xprev : 0)        (Bm1AX replaces AX for PAW)
void xg_cheb_next(int space, int rows, int ncols, double* Xnext, long long ldn, const double* AX, long long lda,
                  const double* X, long long ldx, const double* Xprev, long long ldp, double center, double scale,
                  cudaStream_t st);
// Dense (generalised) Hermitian eigenproblem of the sub-space, eigenvectors overwrite A (xgBlock_heevd / xgBlock_hegvd
// 'v','u'); B == nullptr -> standard problem.  w: DEVICE array of n eigenvalues.  Returns LAPACK-style info.
int xg_hegvd(int space, int n, double* A, long long lda, double* B, long long ldb, double* w, cudaStream_t st);
// xg_RayleighRitz, VAR_X branch (m_xg_ortho_RR.F90:251-571): eigenvalues (device, n) and rotated X, AX, BX.
// BX == nullptr or BX == X: the overlap block is X itself (norm-conserving getAX_BX copies X into BX).
int xg_rayleigh_ritz(int space, int rows, int n, double* X, long long ldx, double* AX, long long ldax, double* BX,
                     long long ldbx, double* eig, bool solve_ax_bx, int me_g0, cudaStream_t st);
void xg_release_workspace();

}  // namespace abi
