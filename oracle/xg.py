"""xgBlock algebra and xg_RayleighRitz (oracle; test infrastructure only).

Restates src/45_xgTools/m_xg.F90 and src/45_xgTools/m_xg_ortho_RR.F90:251-571 (VAR_X branch) with NumPy.
Blocks are band-major complex arrays (ncols, rows) == the memory of cg(2, npw*nband).
  SPACE_CR conventions (istwf_k >= 2): a block is 2*rows reals per column; dot products carry a factor 2 and the G=0
  coefficient (row 0, held by the rank with me_g0 = 1) is counted once:
    xgBlock_gemm 't','n'      m_xg.F90:1802-1882   W = 2 A^T B - 2 (A0r B0r + A0i B0i) + A0r B0r
    colwiseDotProduct         m_xg.F90:4686-4700   2 sum(a b) - a(1) b(1)
    colwiseNorm2              m_xg.F90:4484-4491   2 sum(a a) - (a(1)^2 + a(2)^2)
"parity unpinned" at vector level (the reference stores no xgBlock vectors); pinned through the SCF eigenvalues a
ChebFi2 run built from these primitives reproduces (tests/test_chebfi_pins.py)."""
from __future__ import annotations
import numpy as np
import scipy.linalg as sla

SPACE_R, SPACE_C, SPACE_CR = 1, 2, 3


def _rv(a):
    """real view (ncols, 2*rows) of a complex block"""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a.view(np.float64).reshape(a.shape[0], -1)


def zero_im_g0(space, x, me_g0):
    """m_xg.F90:5851-5898 (in place)"""
    if space == SPACE_CR and me_g0 == 1:
        x[:, 0] = x[:, 0].real
    return x


def gram(space, a, b, me_g0=1):
    """xgBlock_gemm('t','n', 1, A, B, 0, W): W(ncols_a, ncols_b)"""
    if space == SPACE_C:
        return a.conj() @ b.T
    ar, br = _rv(a), _rv(b)
    w = 2.0 * (ar @ br.T)
    if space == SPACE_CR and me_g0 == 1:
        w += -2.0 * (ar[:, :2] @ br[:, :2].T) + ar[:, :1] @ br[:, :1].T
    return w


def colwise_dot(space, a, b, me_g0=1):
    if space == SPACE_C:
        return np.sum(a.conj() * b, axis=1)
    ar, br = _rv(a), _rv(b)
    d = 2.0 * np.sum(ar * br, axis=1)
    if me_g0 == 1:
        d -= ar[:, 0] * br[:, 0]
    return d


def colwise_norm2(space, a, me_g0=1):
    ar = _rv(a)
    if space == SPACE_C:
        return np.sum(ar * ar, axis=1)
    d = 2.0 * np.sum(ar * ar, axis=1)
    if me_g0 == 1:
        d -= ar[:, 0] ** 2 + ar[:, 1] ** 2
    return d


def colwise_cymax(da, b, w):
    """A = W - da(col) * B (m_xg.F90:3301-3413)"""
    return -np.asarray(da)[:, None] * b + w


def rayleigh_ritz(space, x, ax, bx, me_g0=1, solve_ax_bx=True):
    """xg_RayleighRitz VAR_X: returns (eigenvalues, X C, AX C, BX C).  hegvd(1,'v','u') / heevd('v','u')."""
    x = x.copy(); ax = ax.copy(); bx = bx.copy()
    zero_im_g0(space, x, me_g0); zero_im_g0(space, ax, me_g0); zero_im_g0(space, bx, me_g0)      # :376-380
    sub_a = gram(space, x, ax, me_g0)                                                              # :384
    if solve_ax_bx:
        sub_b = gram(space, x, bx, me_g0)                                                          # :388
        w, c = sla.eigh(sub_a, sub_b, lower=False)
    else:
        w, c = sla.eigh(sub_a, lower=False)
    rot = lambda blk: c.T @ blk                                                                    # X . Cwp, :524-531
    return w, rot(x), rot(ax), rot(bx), c
