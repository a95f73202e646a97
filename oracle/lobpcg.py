"""LOBPCG (oracle; test infrastructure only): restates src/48_diago/m_lobpcg2.F90:340-765 (lobpcg_run), one or several
blocks (blockdim = neigenpairs / nblock, lobpcg_orthoXwrtBlocks :803-840, final RR on all bands :744-751), paral_kgb = 0, as driven by src/79_seqpar_mpi/m_lobpcgwf.F90:100-250, with
  xg_Borthonormalize            src/45_xgTools/m_xg_ortho_RR.F90:86-150   (X^H B X = U^H U, X <- X U^-1, also BX, AX)
  xg_RayleighRitz VAR_X/XW/XWP  src/45_xgTools/m_xg_ortho_RR.F90:251-571
  build_pcon                    src/79_seqpar_mpi/m_lobpcgwf.F90:316-334
`apply_h(X) -> (AX, BX)` is the getAX_BX callback on band-major blocks (ncols, npw)."""
from __future__ import annotations
import numpy as np
import scipy.linalg as sla
from . import xg
from .gsphere import KIN_FILTER


def build_pcon(kinpw):
    k = np.where(kinpw > KIN_FILTER, 0.0, kinpw)
    num = 27 + k * (18 + k * (12 + 8 * k))
    return np.where(kinpw > KIN_FILTER, 0.0, num / (num + 16 * k ** 4))


def b_orthonormalize(space, x, bx, ax, me_g0):
    """in place; returns info (0 ok)"""
    xg.zero_im_g0(space, x, me_g0); xg.zero_im_g0(space, bx, me_g0); xg.zero_im_g0(space, ax, me_g0)
    buf = xg.gram(space, x, bx, me_g0)
    try:
        u = sla.cholesky(buf, lower=False)                      # potrf 'u': buf = U^H U
    except np.linalg.LinAlgError:
        return 1
    uinv = sla.solve_triangular(u, np.eye(u.shape[0]), lower=False)
    for blk in (x, bx, ax):
        blk[...] = uinv.T @ blk                                  # trsm 'r','u','n': X <- X U^-1  (band-major: rows mix)
    return 0


def rayleigh_ritz_xwp(space, me_g0, n, xwp, axwp, bxwp, nvar):
    """nvar = 2 (VAR_XW) or 3 (VAR_XWP); blocks are views [X | W | P] of n rows each (band-major).  Updates X, AX, BX, P, AP, BP
    in place and returns the subdim eigenvalues."""
    sub = nvar * n
    for blk in (xwp, axwp, bxwp):
        xg.zero_im_g0(space, blk[:sub], me_g0)
    dt = np.complex128 if space == xg.SPACE_C else np.float64
    sa = np.zeros((sub, sub), dtype=dt); sb = np.zeros((sub, sub), dtype=dt)
    sa[:n, :n] = xg.gram(space, xwp[:n], axwp[:n], me_g0); sb[:n, :n] = xg.gram(space, xwp[:n], bxwp[:n], me_g0)
    sa[:2 * n, n:2 * n] = xg.gram(space, xwp[:2 * n], axwp[n:2 * n], me_g0)
    sb[:2 * n, n:2 * n] = xg.gram(space, xwp[:2 * n], bxwp[n:2 * n], me_g0)
    if nvar == 3:
        sa[:, 2 * n:] = xg.gram(space, xwp[:sub], axwp[2 * n:sub], me_g0)
        sb[:, 2 * n:] = xg.gram(space, xwp[:sub], bxwp[2 * n:sub], me_g0)
    w, vec = sla.eigh(sa, sb, lower=False)                       # hegvd(1,'v','u'): only the upper triangles are read
    c0 = vec[:n, :n]; c1 = vec[n:sub, :n]
    for blk in (xwp, axwp, bxwp):
        xn = c0.T @ blk[:n]                                       # X <- X Cwp
        p = c1.T @ blk[n:sub]                                     # P <- WP Cwp (after the cshift of the eigenvector rows)
        blk[2 * n:3 * n] = p
        blk[:n] = xn + p                                          # xgBlock_add(X, P)
    return w


def ortho_x_wrt_blocks(space, var, x0_prev, bx0_prev, me_g0):
    """lobpcg_orthoXwrtBlocks (m_lobpcg2.F90:803-840): var <- var - X0 (BX0^H var) with the previous blocks, in place
    (band-major blocks: rows are bands)."""
    buf = xg.gram(space, bx0_prev, var, me_g0)                  # (nprev, n) = BX0^H var  (xgBlock_gemm 't','n')
    var -= buf.T @ x0_prev                                      # var(:, j) -= sum_i X0(:, i) buf(i, j)


def lobpcg_run(apply_h, x0, pcon, space, me_g0, nline, tolerance=1e-20, nbdbuf=0, occ=None, info=None, nblock=1):
    """lobpcg_run (m_lobpcg2.F90:340-765), paral_kgb = 0, blocks of blockdim = nband / nblock bands.  Returns
    (eigenvalues, residuals, X) for all bands."""
    nband, npw = x0.shape
    assert nband % nblock == 0
    n = nband // nblock
    all_x0 = np.array(x0, dtype=np.complex128)
    all_ax0 = np.zeros_like(all_x0); all_bx0 = np.zeros_like(all_x0)
    eig_all = np.zeros(nband); resid_all = np.zeros(nband)
    lines = []
    for iblock in range(nblock):
        sl = slice(iblock * n, (iblock + 1) * n)
        sub = {}
        e, r, x, ax, bx = _lobpcg_block(apply_h, all_x0[sl], pcon, space, me_g0, nline, tolerance, nbdbuf,
                                        None if occ is None else occ[sl], sub, iblock, nband,
                                        all_x0[:iblock * n], all_bx0[:iblock * n])
        eig_all[sl] = e; resid_all[sl] = r
        all_x0[sl] = x                                           # lobpcg_setX0
        all_ax0[sl] = ax; all_bx0[sl] = bx                       # lobpcg_transferAX_BX
        lines.append(sub["nline_done"])
    if nblock > 1:                                               # m_lobpcg2.F90:744-751
        b_orthonormalize(space, all_x0, all_bx0, all_ax0, me_g0)
        w, xr, _, _, _ = xg.rayleigh_ritz(space, all_x0, all_ax0, all_bx0, me_g0, solve_ax_bx=False)
        all_x0 = xr; eig_all = w.copy()
    if info is not None:
        info.update(nline_done=lines[0] if nblock == 1 else lines)
    return eig_all, resid_all, all_x0


def _lobpcg_block(apply_h, x0, pcon, space, me_g0, nline, tolerance, nbdbuf, occ, info, iblock, nband, x0_prev, bx0_prev):
    """One block of the big loop over blocks (m_lobpcg2.F90:456-695).  Returns (eig, resid, X, AX, BX)."""
    n, npw = x0.shape
    xwp = np.zeros((3 * n, npw), dtype=np.complex128); axwp = np.zeros_like(xwp); bxwp = np.zeros_like(xwp)
    xwp[:n] = x0
    X, W, P = xwp[:n], xwp[n:2 * n], xwp[2 * n:]
    AX, AW = axwp[:n], axwp[n:2 * n]
    BX, BW = bxwp[:n], bxwp[n:2 * n]

    def get_ax_bx(src, a_dst, b_dst):
        a, b = apply_h(src)
        a_dst[...] = a; b_dst[...] = b
        xg.zero_im_g0(space, a_dst, me_g0); xg.zero_im_g0(space, b_dst, me_g0)
    nband_eff = nband - nbdbuf if nbdbuf > 0 else nband         # m_lobpcg2.F90:394-398 (counted over ALL bands)
    iband_min = n * iblock                                       # 0-based first band of this block
    if iblock > 0:
        ortho_x_wrt_blocks(space, X, x0_prev, bx0_prev, me_g0)   # :464-469
    get_ax_bx(X, AX, BX)
    b_orthonormalize(space, X, BX, AX, me_g0)
    w, xr, axr, bxr, _ = xg.rayleigh_ritz(space, X, AX, BX, me_g0, solve_ax_bx=False)      # VAR_X, heevd
    X[...] = xr; AX[...] = axr; BX[...] = bxr
    eig = w.copy()
    compute_residu = True
    resid = np.zeros(n)
    nline_done = 0

    def residuals():
        W[...] = xg.colwise_cymax(eig, BX, AX)                   # lobpcg_getResidu
        r = xg.colwise_norm2(space, W, me_g0)
        W[...] = W * pcon[None, :]                               # xgBlock_apply_diag(W, pcond)
        if nbdbuf >= 0:
            eff = r[:min(n, max(nband_eff - iband_min, 0))]      # bands of this block below nband_eff (:526-538)
            mn, mx = (float(eff.min()), float(eff.max())) if eff.size else (0.0, 0.0)
        elif nbdbuf == -101:
            mn = float(r.min()); mx = float((r * occ).max())
        else:
            raise ValueError("Bad value of nbdbuf")
        return r, mn, mx
    for iline in range(1, nline + 1):
        resid, min_res, max_res = residuals()
        if max_res < tolerance:
            compute_residu = False
            break
        if iblock > 0:
            ortho_x_wrt_blocks(space, W, x0_prev, bx0_prev, me_g0)   # :553-555
        get_ax_bx(W, AW, BW)
        if iline == 1 or min_res < 1e-27:
            b_orthonormalize(space, xwp[:2 * n], bxwp[:2 * n], axwp[:2 * n], me_g0)
            xwp[2 * n:] = 0; axwp[2 * n:] = 0; bxwp[2 * n:] = 0
            w = rayleigh_ritz_xwp(space, me_g0, n, xwp, axwp, bxwp, 2)
        else:
            ierr = b_orthonormalize(space, xwp, bxwp, axwp, me_g0)
            if ierr == 0:
                w = rayleigh_ritz_xwp(space, me_g0, n, xwp, axwp, bxwp, 3)
            else:
                b_orthonormalize(space, xwp[:2 * n], bxwp[:2 * n], axwp[:2 * n], me_g0)
                xwp[2 * n:] = 0; axwp[2 * n:] = 0; bxwp[2 * n:] = 0
                w = rayleigh_ritz_xwp(space, me_g0, n, xwp, axwp, bxwp, 2)
        eig = w[:n].copy()
        nline_done = iline
    if compute_residu:
        resid, _, _ = residuals()
    if info is not None:
        info.update(nline_done=nline_done)
    return eig, resid, X.copy(), AX.copy(), BX.copy()
