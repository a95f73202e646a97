"""PAW inverse overlap (oracle; test infrastructure only): restates src/66_wfs/m_invovl.F90
  make_invovl   :469-776   inv_sij, inv_s_approx = (inv_sij + P_atom1^H P_atom1)^-1, gram_projs = P^H P
                           (istwf_k > 1: projectors * sqrt2, G=0 coefficient real and unscaled, :601-607)
  apply_invovl  :790-1039  S^-1 = 1 - P (s^-1 + P^H P)^-1 P^H through nonlop choice 0 / solve_inner / nonlop choice 7
  solve_inner   :1052-1152 preconditioned fixed-point iteration, same stopping logic
  apply_block   :1165-1231 block-diagonal (per atom) symmetric apply
"parity unpinned" by stored vectors; checked by the invariant S S^-1 psi = psi with S from gemm_nonlop(paw_opt=3)."""
from __future__ import annotations
import numpy as np
from .nonlop import nlmn_of_types, _unpack_sym, opernla, opernlb


class Invovl:
    pass


def make_invovl(P, sij, indlmn, nattyp, istwf_k, me_g0=1):
    """P: (nprojs, npw) complex from nonlop.prep_projectors; sij: (ntypat, lmn2) packed."""
    iv = Invovl()
    nl = nlmn_of_types(indlmn)
    if istwf_k == 1:
        gram = np.conj(P) @ P.T                                   # gram(i,j) = sum_G conj(P_i) P_j
    else:
        Q = P * np.sqrt(2.0)
        if istwf_k == 2 and me_g0 == 1:
            Q[:, 0] = P[:, 0].real
        Qr = np.ascontiguousarray(Q).view(np.float64).reshape(Q.shape[0], -1)
        gram = Qr @ Qr.T
    iv.gram = gram
    iv.inv_sij = []; iv.inv_s_approx = []
    shift = 0
    for t in range(indlmn.shape[0]):
        n = nl[t]
        inv = np.linalg.inv(_unpack_sym(sij[t], n))
        iv.inv_sij.append(inv)
        iv.inv_s_approx.append(np.linalg.inv(inv + gram[shift:shift + n, shift:shift + n]))
        shift += n * int(nattyp[t])
    iv.nl = nl; iv.nattyp = [int(x) for x in nattyp]
    return iv


def apply_block(iv, mats, x):
    """y(ndat, nprojs): per atom y = M_type x (ZHEMM/DSYMM 'L','U')"""
    y = np.zeros_like(x)
    shift = 0
    for t, n in enumerate(iv.nl):
        for _ in range(iv.nattyp[t]):
            y[:, shift:shift + n] = x[:, shift:shift + n] @ mats[t].T
            shift += n
    return y


def solve_inner(iv, proj, info=None):
    precision = 1e-16
    normprojs = np.sum(np.abs(proj) ** 2, axis=1)
    x = apply_block(iv, iv.inv_s_approx, proj)
    additional = -1; previous = 0.0; maxerr = 0.0
    for i in range(1, 31):
        resid = apply_block(iv, iv.inv_sij, x)
        ptp = x @ iv.gram.T
        resid = proj - resid - ptp
        errs = np.sum(np.abs(resid) ** 2, axis=1)
        maxerr = np.sqrt(np.max(errs / normprojs))
        if maxerr < precision or additional == 1:
            break
        elif maxerr < 1e-10 and additional == -1:
            rate = -np.log(1e-10) / i
            additional = int(np.ceil(-np.log(precision / 1e-10) / rate)) + 1
        elif additional > 0:
            if previous < maxerr:
                break
            additional -= 1
        previous = maxerr
        x = x + apply_block(iv, iv.inv_s_approx, resid)
    if info is not None:
        info.update(iters=i, maxerr=maxerr)
    return x, ptp


def apply_invovl(P, iv, cwavef, istwf_k, me_g0=1, info=None):
    """Returns (sm1cwavef, cprj) with cprj = proj - P^H P sm1proj as the reference leaves it in cwaveprj."""
    cwavef = np.atleast_2d(cwavef)
    proj = opernla(P, cwavef, istwf_k, me_g0)                     # nonlop choice 0
    x, ptp = solve_inner(iv, proj, info)
    sm1proj = -x; ptpsm1 = -ptp
    sm1 = opernlb(P, sm1proj, istwf_k)                            # nonlop choice 7 (no + vectin)
    return cwavef + sm1, proj + ptpsm1
