"""ChebFi2 (oracle; test infrastructure only): restates src/48_diago/m_chebfi2.F90:466-1210 (chebfi_run, paral_kgb=0)
as it is driven by src/79_seqpar_mpi/m_chebfiwf.F90:110-330 (lambda_plus = ecut, getAX_BX = getghc, NC: BX = copy of X).
`apply_h(X) -> (AX, BX)` is the getAX_BX callback on band-major blocks (ncols, npw)."""
from __future__ import annotations
import numpy as np
from . import xg


def cheb_oracle1(xx, aa, bb, tol, nmax):
    """m_chebfi2.F90:1031-1064"""
    xred = (xx - (aa + bb) / 2) / (bb - aa) * 2
    yy = xred; yim1 = 1.0
    nn = nmax
    if 1 / (yy ** 2) < tol:
        nn = 1
    else:
        for ii in range(2, nmax):
            temp = yy
            yy = 2 * xred * yy - yim1
            yim1 = temp
            if 1 / (yy ** 2) < tol:
                nn = ii
                break
    return nn


def cheb_poly1(xx, nn, aa, bb):
    """m_chebfi2.F90:1084-1106"""
    xred = (xx - (aa + bb) / 2) / (bb - aa) * 2
    yy = xred; yim1 = 1.0
    for _ in range(2, nn + 1):
        temp = yy
        yy = 2 * xred * yy - yim1
        yim1 = temp
    return yy


def ndeg_from_residu(space, me_g0, ax, bx, div, occ, lm, lp, ndeg_max, tolerance, ndeg_filter, nbdbuf, oracle,
                     oracle_factor, oracle_min_occ):
    """chebfi_set_ndeg_from_residu, m_chebfi2.F90:1131-1210 (one band group)"""
    ncols = ax.shape[0]
    res = xg.colwise_norm2(space, xg.colwise_cymax(div, bx, ax), me_g0)
    if nbdbuf == -101:
        res = res * occ
    nb = nbdbuf if nbdbuf > 0 else 0
    nd = 0
    for i in range(ncols):
        t1 = res[i] < tolerance
        t2 = (i + 1) > ncols - nb
        t3 = nbdbuf == -101 and occ[i] < oracle_min_occ
        if t1 or t2 or t3:
            n = 0
        else:
            n_tol = cheb_oracle1(div[i], lm, lp, tolerance / res[i], 1000)
            if oracle == 1:
                n = min(ndeg_max, n_tol, ndeg_filter)
            elif oracle == 2:
                n = min(ndeg_max, n_tol, cheb_oracle1(div[i], lm, lp, oracle_factor, 15))
            else:
                raise ValueError("Wrong value for chebfi%oracle")
        nd = max(nd, n)
    return nd


def chebfi_run(apply_h, x0, space, me_g0, ecut, nline, tolerance=1e-20, occ=None, nbdbuf=0, oracle=0,
               oracle_factor=1e-2, oracle_min_occ=1e-8, info=None, get_bm1x=None):
    """Returns (eigenvalues, residuals, X).  x0: (nband, npw) complex.  get_bm1x(AX) -> S^-1 AX for PAW (chebfi%paw)."""
    def get_ax_bx(x):
        ax, bx = apply_h(x)
        ax = np.array(ax, copy=True); bx = np.array(bx, copy=True)
        xg.zero_im_g0(space, ax, me_g0); xg.zero_im_g0(space, bx, me_g0)      # :580-581
        return ax, bx

    x = np.array(x0, dtype=np.complex128, copy=True)
    lambda_plus = ecut
    ax, bx = get_ax_bx(x)
    # chebfi_rayleighRitzQuotients :761-810
    r1 = xg.colwise_dot(space, x, ax, me_g0); r2 = xg.colwise_dot(space, x, bx, me_g0)
    div = np.real(r1 / r2)
    maxeig, mineig = float(div.max()), float(div.min())
    lambda_minus = maxeig
    ndeg_max = cheb_oracle1(mineig, lambda_minus, lambda_plus, 1e-16, 40)
    ndeg = min(ndeg_max, nline)
    if oracle > 0:
        ndeg = ndeg_from_residu(space, me_g0, ax, bx, div, occ, lambda_minus, lambda_plus, ndeg_max, tolerance, nline,
                                nbdbuf, oracle, oracle_factor, oracle_min_occ)
    center = (lambda_plus + lambda_minus) * 0.5
    radius = (lambda_plus - lambda_minus) * 0.5
    one_over_r = 1 / radius; two_over_r = 2 / radius
    x_prev = None
    for ideg in range(ndeg):
        # chebfi_computeNextOrderChebfiPolynom :837-896, same operation order as the reference (NC: X_next = copy of AX)
        x_next = ax.copy() if get_bm1x is None else np.array(get_bm1x(ax), copy=True)
        x *= center
        x_next += -1.0 * x
        x *= 1 / center
        if ideg == 0:
            x_next *= one_over_r
        else:
            x_next *= two_over_r
            x_next += -1.0 * x_prev
        x_prev, x = x, x_next                                                  # chebfi_swapInnerBuffers
        ax, bx = get_ax_bx(x)
    # chebfi_ampfactor :944-995
    for j in range(x.shape[0]):
        amp = cheb_poly1(div[j], ndeg, lambda_minus, lambda_plus)
        if abs(amp) < 1e-3:
            amp = 1e-3
        x[j] *= 1 / amp; ax[j] *= 1 / amp; bx[j] *= 1 / amp
    w, x, ax, bx, _ = xg.rayleigh_ritz(space, x, ax, bx, me_g0, solve_ax_bx=True)   # :705
    resid = xg.colwise_norm2(space, xg.colwise_cymax(w, bx, ax), me_g0)        # :709-716 (NC: B X = X)
    if info is not None:
        info.update(ndeg=ndeg, lambda_minus=lambda_minus, lambda_plus=lambda_plus, div=div)
    return w, resid, x
