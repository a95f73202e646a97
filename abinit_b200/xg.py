"""Host-side mirror of the eigensolver-side interfaces around getghc (same names / argument meaning as the reference):

  xgBlock_gemm ('t','n') , rotation, colwise ops   src/45_xgTools/m_xg.F90
  xg_RayleighRitz                                 src/45_xgTools/m_xg_ortho_RR.F90:251
  chebfiwf2 / chebfi_run                          src/79_seqpar_mpi/m_chebfiwf.F90:110, src/48_diago/m_chebfi2.F90:466

Blocks are CUDA torch tensors (or any object with data_ptr()) holding cg(2, npw*nband) memory: shape (ncols, rows)
complex128 or (ncols, rows, 2) float64.  Everything runs in libabinit_b200.so; there is no CPU path here.
"""
from __future__ import annotations
import ctypes as C
import numpy as np
from .api import L, _ptr, _iref, _dref, _F, Hamiltonian

SPACE_R, SPACE_C, SPACE_CR = 1, 2, 3      # m_xg.F90:63-65


def xg_gram(space, rows, ncols_a, ncols_b, a, lda, b, ldb, w, ldw, me_g0=1):
    """W(ncols_a, ncols_b) = A^H B (xgBlock_gemm 't','n'); W column-major, real (SPACE_R/CR) or complex (SPACE_C)."""
    L().abi_b200_xg_gram_(_iref(space), _iref(rows), _iref(ncols_a), _iref(ncols_b), _ptr(a), _iref(lda), _ptr(b), _iref(ldb),
                          _ptr(w), _iref(ldw), _iref(me_g0))


def xg_rotate(space, rows, k, ncols_out, x, ldx, c, ldc):
    """X(:, :ncols_out) <- X(:, :k) C(:k, :ncols_out) in place."""
    L().abi_b200_xg_rotate_(_iref(space), _iref(rows), _iref(k), _iref(ncols_out), _ptr(x), _iref(ldx), _ptr(c), _iref(ldc))


def xg_gemm_nn(space, rows, k, ncols_out, a, lda, c, ldc, out, ldo, upper=False):
    """out(:, :ncols_out) = a(:, :k) c(:k, :ncols_out); out may alias a; upper: c is upper triangular (U^-1 of xg_Borthonormalize)."""
    L().abi_b200_xg_gemm_nn_(_iref(space), _iref(rows), _iref(k), _iref(ncols_out), _ptr(a), _iref(lda), _ptr(c), _iref(ldc), _ptr(out),
                             _iref(ldo), _iref(1 if upper else 0))


def xg_chol_inverse(space, m, a, lda) -> int:
    """a (upper triangle of a Hermitian positive m x m sub-space matrix) -> U^-1 with a = U^H U; returns potrf's info."""
    info = C.c_int(0)
    L().abi_b200_xg_chol_inverse_(_iref(space), _iref(m), _ptr(a), _iref(lda), C.byref(info))
    return info.value


def xg_hegvd(space, n, a, lda, b, ldb, w) -> int:
    """hegvd(1,'v','u') (b=None: heevd('v','u')): eigenvectors overwrite a, eigenvalues in the device array w."""
    info = C.c_int(0)
    L().abi_b200_xg_hegvd_(_iref(space), _iref(n), _ptr(a), _iref(lda), _ptr(b), _iref(ldb), _ptr(w), C.byref(info))
    return info.value


def xg_colwise(op, space, rows, ncols, a, lda, b=None, ldb=0, w=None, ldw=0, da=None, out=None, me_g0=1):
    """op: 'dot' | 'norm2' | 'cymax' | 'scale' | 'zero_im_g0' (include/abinit_b200.h: abi_b200_xg_colwise_)."""
    code = {"dot": 0, "norm2": 1, "cymax": 2, "scale": 3, "zero_im_g0": 4, "add": 5, "apply_diag": 6}[op]
    L().abi_b200_xg_colwise_(_iref(code), _iref(space), _iref(rows), _iref(ncols), _ptr(a), _iref(lda), _ptr(b), _iref(ldb),
                             _ptr(w), _iref(ldw), _ptr(da), _ptr(out), _iref(me_g0))


def xg_RayleighRitz(x, ax, bx, eigenvalues, space, rows, blockdim, me_g0=1, solve_ax_bx=True, ldx=None) -> int:
    """xg_RayleighRitz(X, AX, BX, eigenvalues, info, ..., solve_ax_bx): VAR_X branch.  bx=None -> BX is X (NC)."""
    info = C.c_int(0)
    ld = rows if ldx is None else ldx
    L().abi_b200_xg_rayleigh_ritz_(_iref(space), _iref(rows), _iref(blockdim), _ptr(x), _iref(ld), _ptr(ax), _iref(ld),
                                   _ptr(bx), _iref(ld), _ptr(eigenvalues, _F, "eigenvalues"), C.byref(info),
                                   _iref(1 if solve_ax_bx else 0), _iref(me_g0))
    return info.value


def cheb_oracle1(xx, aa, bb, tol, nmax) -> int:
    return int(L().abi_b200_cheb_oracle1_(_dref(xx), _dref(aa), _dref(bb), _dref(tol), _iref(nmax)))


def cheb_poly1(xx, nn, aa, bb) -> float:
    return float(L().abi_b200_cheb_poly1_(_dref(xx), _iref(nn), _dref(aa), _dref(bb)))


def chebfiwf2(cg, eig, occ, enl_out, gs_hamk: Hamiltonian, nband, npw, nspinor, resid, tolwfr_diago, ecut, nline,
              nbdbuf=0, chebfi_oracle=0, oracle_factor=1e-2, oracle_min_occ=1e-8, bandpp=None, prtvol=0):
    """chebfiwf2(cg, dtset, eig, occ, enl_out, gs_hamk, ..., nband, npw, nspinor, prtvol, resid): one ChebFi2 call on the
    nband wavefunctions cg(2, npw*nband) (host array or CUDA tensor, updated in place); eig / resid / enl_out: host
    float64 arrays of nband (output).  dtset scalars are passed by name."""
    hp = C.c_void_p(gs_hamk.h)
    bp = int(nband if bandpp is None else bandpp)
    L().abi_b200_chebfiwf2_(_ptr(cg, _F, "cg"), _ptr(eig, _F, "eig"), _ptr(occ, _F, "occ"), _ptr(enl_out, _F, "enl_out"),
                            C.byref(hp), _iref(nband), _iref(npw), _iref(nspinor), _iref(prtvol), _ptr(resid, _F, "resid"),
                            _dref(tolwfr_diago), _dref(ecut), _iref(nline), _iref(nbdbuf), _iref(chebfi_oracle),
                            _dref(oracle_factor), _dref(oracle_min_occ), _iref(bp))


def chebfi_rq(gs_hamk: Hamiltonian, ncols, bandpp, x, ax, bx=None):
    """Phase 1 of chebfi_run on the local band group: AX (BX) = getAX_BX(X), Rayleigh quotients.  Returns
    (div[ncols], maxeig, mineig); a band-parallel caller reduces maxeig/mineig over its ranks (m_chebfi2.F90:606-611)."""
    hp = C.c_void_p(gs_hamk.h)
    div = np.zeros(ncols); mx = C.c_double(0.0); mn = C.c_double(0.0)
    L().abi_b200_chebfi_rq_(C.byref(hp), _iref(ncols), _iref(bandpp), _ptr(x), _ptr(ax), _ptr(bx), div.ctypes.data,
                            C.byref(mx), C.byref(mn))
    return div, mx.value, mn.value


def chebfi_core(gs_hamk: Hamiltonian, ncols, bandpp, x, ax, bx, x_next, x_prev, lambda_minus, lambda_plus, ndeg_filter, div):
    """Phase 2: filter loop + amplification.  x, x_next, x_prev are rotated as chebfi_swapInnerBuffers does; returns the
    three objects in their new roles (filtered X first)."""
    hp = C.c_void_p(gs_hamk.h)
    objs = {_ptr(x): x, _ptr(x_next): x_next, _ptr(x_prev): x_prev}
    px, pn, pp = C.c_void_p(_ptr(x)), C.c_void_p(_ptr(x_next)), C.c_void_p(_ptr(x_prev))
    d = np.ascontiguousarray(div, dtype=np.float64)
    L().abi_b200_chebfi_core_(C.byref(hp), _iref(ncols), _iref(bandpp), C.byref(px), _ptr(ax), _ptr(bx), C.byref(pn), C.byref(pp),
                              _dref(lambda_minus), _dref(lambda_plus), _iref(ndeg_filter), d.ctypes.data)
    return objs[px.value], objs[pn.value], objs[pp.value]


def lobpcgwf2(cg, eig, occ, enl_out, gs_hamk: Hamiltonian, nband, npw, nspinor, resid, tolwfr_diago, nline, nblock_lobpcg=1,
              nbdbuf=0, bandpp=None, prtvol=0):
    """lobpcgwf2(cg, dtset, eig, occ, enl_out, gs_hamk, ..., nband, npw, nspinor, prtvol, resid, nbdbuf): one LOBPCG call on the
    nband wavefunctions in nblock_lobpcg blocks of nband / nblock_lobpcg bands (m_lobpcgwf.F90:133); arguments as chebfiwf2."""
    hp = C.c_void_p(gs_hamk.h)
    bp = int(nband // nblock_lobpcg if bandpp is None else bandpp)
    L().abi_b200_lobpcgwf2_(_ptr(cg, _F, "cg"), _ptr(eig, _F, "eig"), _ptr(occ, _F, "occ"), _ptr(enl_out, _F, "enl_out"),
                            C.byref(hp), _iref(nband), _iref(npw), _iref(nspinor), _iref(prtvol), _ptr(resid, _F, "resid"),
                            _dref(tolwfr_diago), _iref(nline), _iref(nblock_lobpcg), _iref(nbdbuf), _iref(bp))


# ---- band-parallel ChebFi2 inside the library (NCCL communicator owned by libabinit_b200.so) ----
def comm_get_unique_id() -> bytes:
    """ncclGetUniqueId (128 bytes): call on rank 0 and broadcast with the application's own transport."""
    buf = C.create_string_buffer(128)
    L().abi_b200_comm_get_unique_id_(buf)
    return buf.raw


def comm_init_rank(uid: bytes, nranks: int, rank: int):
    if len(uid) != 128:
        raise ValueError("the ncclUniqueId is 128 bytes")
    L().abi_b200_comm_init_rank_(C.create_string_buffer(uid, 128), _iref(nranks), _iref(rank))


def comm_destroy():
    L().abi_b200_comm_destroy_()


def xg_transpose(to_rows: bool, cols, lin, rows, nband):
    """xgTransposer_transpose between cols (2, rows, my_ncols) and lin (2, my_nrows, nband), device blocks."""
    L().abi_b200_xg_transpose_(_iref(1 if to_rows else 0), _ptr(cols), _ptr(lin), _iref(rows), _iref(nband))


def chebfiwf2_paral(cg, eig, occ, enl_out, resid, gs_hamk: Hamiltonian, nband, ncols_mine, npw, nspinor, tolwfr_diago, ecut, nline,
                    nbdbuf=0, chebfi_oracle=0, oracle_factor=1e-2, oracle_min_occ=1e-8, bandpp=None):
    """chebfiwf2 with paral_kgb = 1 over the ranks of the library communicator: cg holds this rank's band block (in/out); eig and occ
    (nband, replicated) and enl_out / resid (ncols_mine) are host float64 arrays; the dtset scalars as in chebfiwf2."""
    hp = C.c_void_p(gs_hamk.h)
    bp = int(max(ncols_mine, 1) if bandpp is None else bandpp)
    L().abi_b200_chebfiwf2_paral_(_ptr(cg, _F, "cg"), _ptr(eig, _F, "eig"), _ptr(occ, _F, "occ"), _ptr(enl_out, _F, "enl_out"),
                                  _ptr(resid, _F, "resid"), C.byref(hp), _iref(nband), _iref(ncols_mine), _iref(npw), _iref(nspinor),
                                  _dref(tolwfr_diago), _dref(ecut), _iref(nline), _iref(nbdbuf), _iref(chebfi_oracle), _dref(oracle_factor),
                                  _dref(oracle_min_occ), _iref(bp))


def lobpcgwf2_paral(cg, eig, resid, gs_hamk: Hamiltonian, nband, ncols_mine, npw, nspinor, tolwfr_diago, nline, bandpp=None):
    """lobpcgwf2 (one block of all bands) with paral_kgb = 1 over the ranks of the library communicator: cg holds this rank's band
    block (in/out); eig and resid (nband, replicated) are host float64 arrays."""
    hp = C.c_void_p(gs_hamk.h)
    bp = int(max(ncols_mine, 1) if bandpp is None else bandpp)
    L().abi_b200_lobpcgwf2_paral_(_ptr(cg, _F, "cg"), _ptr(eig, _F, "eig"), _ptr(resid, _F, "resid"), C.byref(hp), _iref(nband),
                                  _iref(ncols_mine), _iref(npw), _iref(nspinor), _dref(tolwfr_diago), _iref(nline), _iref(bp))
