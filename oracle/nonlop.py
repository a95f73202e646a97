"""gemm_nonlop, choice=1 (and 0/7), signs=2 (oracle; test infrastructure only).

Restates  prep_projectors     src/66_nonlocal/m_gemm_nonlop_projectors.F90:792-1038
          opernla_gemm        src/66_nonlocal/m_opernla_gemm.F90:361-712
          opernlc_ylm_allwf   src/66_nonlocal/m_opernlc_ylm_allwf.F90:308-332 (NC), :336-447 (PAW real Dij),
                              :1253-1295 (Sij)
          opernlb_gemm        src/66_nonlocal/m_opernlb_gemm.F90:353-837
          gemm_nonlop         src/66_nonlocal/m_gemm_nonlop.F90:191-1242
NC branch PINNED through the tbase3_1 SCF numbers (oracle/scf.py, tests/test_scf_pins.py); PAW branches are parity
unpinned by stored data and checked by invariants in tests/.
Index conventions: indlmn is the Fortran indlmn(6, lmnmax, ntypat) stored here as numpy (ntypat, lmnmax, 6)
(same memory, C order); l = indlmn[t, ilmn, 0], iln = indlmn[t, ilmn, 4] (1-based), validity indlmn[t, ilmn, 2] > 0.
"""
from __future__ import annotations
import numpy as np

FOUR_PI = 4.0 * np.pi


def nlmn_of_types(indlmn):
    return [int(np.count_nonzero(indlmn[t, :, 2] > 0)) for t in range(indlmn.shape[0])]


def count_nprojs(indlmn, nattyp):
    """m_gemm_nonlop.F90:400-403."""
    return int(sum(n * int(na) for n, na in zip(nlmn_of_types(indlmn), nattyp)))


def prep_projectors(ffnl, ph3d, indlmn, nattyp, ucvol):
    """P(nprojs, npw) complex: P = 4 pi / sqrt(ucvol) * ffnl(:,1,ilmn,itypat) * (-i)^l * conj(ph3d(:, ia))
    ordered type-major, atom, lmn (m_gemm_nonlop_projectors.F90:870-1012).
      ffnl : (ntypat, lmnmax, dimffnl, npw)  == Fortran ffnl(npw, dimffnl, lmnmax, ntypat)
      ph3d : (natom, npw) complex             == Fortran ph3d(2, npw, matblk), atoms sorted by type."""
    ntypat = indlmn.shape[0]
    npw = ffnl.shape[-1]
    wt = FOUR_PI / np.sqrt(ucvol)
    nprojs = count_nprojs(indlmn, nattyp)
    P = np.zeros((nprojs, npw), dtype=np.complex128)
    shift = 0; ia3 = 0
    for t in range(ntypat):
        nlmn = nlmn_of_types(indlmn)[t]
        phase = np.array([(-1j) ** (int(indlmn[t, i, 0]) % 4) for i in range(nlmn)])
        for _ in range(int(nattyp[t])):
            P[shift:shift + nlmn] = wt * ffnl[t, :nlmn, 0, :] * phase[:, None] * np.conj(ph3d[ia3])[None, :]
            shift += nlmn; ia3 += 1
    return P


def opernla(P, vectin, istwf_k, me_g0=1):
    """gx(ndat, nprojs) = P^H psi.  istwf_k>=2: real result 2*(P_r^T psi_r + P_i^T psi_i) with Re halved and
    Im zeroed at G=0 when istwf_k==2 (m_opernla_gemm.F90:569-612, 641-689)."""
    vectin = np.atleast_2d(vectin)
    if istwf_k == 1:
        return vectin @ np.conj(P).T
    vr = vectin.real.copy(); vi = vectin.imag.copy()
    if istwf_k == 2 and me_g0 == 1:
        vr[:, 0] *= 0.5; vi[:, 0] = 0.0
    return 2.0 * (vr @ P.real.T + vi @ P.imag.T)


def _unpack_sym(packed, nlmn):
    """Packed upper-triangular j0lmn=j(j-1)/2 storage -> full symmetric (m_opernlc_ylm_allwf.F90:395-447)."""
    D = np.zeros((nlmn, nlmn))
    for j in range(nlmn):
        for i in range(j + 1):
            D[i, j] = D[j, i] = packed[j * (j + 1) // 2 + i]
    return D


def opernlc(gx, enl, sij, indlmn, nattyp, atindx1, paw_opt, lambda_=None):
    """gxfac, gxfac_sij from gx.
      paw_opt=0 (NC): gxfac = enl[itypat, iln] * gx            (enl == Fortran enl(dimenl1, ntypat), here (ntypat, dimenl1))
      paw_opt=1/4   : gxfac = D_ij(atom) gx  (enl here (natom, dimenl1) packed, indexed by atindx1 (0-based))
      paw_opt=2     : gxfac = (D_ij - lambda S_ij) gx
      paw_opt=3/4   : gxfac_sij = S_ij(type) gx (sij here (ntypat, dimenl1) packed)"""
    ndat, nprojs = gx.shape
    gxfac = np.zeros_like(gx); gxs = np.zeros_like(gx) if paw_opt in (3, 4) else None
    shift = 0; iatm = 0
    for t in range(indlmn.shape[0]):
        nlmn = nlmn_of_types(indlmn)[t]
        S = _unpack_sym(sij[t], nlmn) if paw_opt in (2, 3, 4) else None
        for ia in range(int(nattyp[t])):
            sl = slice(shift, shift + nlmn)
            if paw_opt == 0:
                iln = indlmn[t, :nlmn, 4].astype(int) - 1
                gxfac[:, sl] = enl[t, iln][None, :] * gx[:, sl]
            elif paw_opt in (1, 2, 4):
                D = _unpack_sym(enl[atindx1[iatm + ia]], nlmn)
                gxfac[:, sl] = gx[:, sl] @ D.T
                if paw_opt == 2:
                    gxfac[:, sl] -= np.asarray(lambda_)[:, None] * (gx[:, sl] @ S.T)
            if paw_opt in (3, 4):
                gxs[:, sl] = gx[:, sl] @ S.T
            shift += nlmn
        iatm += int(nattyp[t])
    return gxfac, gxs


def _hermitian_from_packed(packed, nlmn, cplex_enl, diag_real=True):
    """Full matrix M[j, i] the reference applies for one spin-DIAGONAL block (m_opernlc_ylm_allwf.F90:453-655):
    M[j,i] = conj(E[i,j]) for i < j, Re E[j,j] on the diagonal (the stored imaginary part is never read), E[j,i] for i > j,
    E = packed upper triangle, (re, im) pairs when cplex_enl = 2."""
    E = np.zeros((nlmn, nlmn), dtype=np.complex128)
    for j in range(nlmn):
        for i in range(j + 1):
            pk = j * (j + 1) // 2 + i
            E[i, j] = packed[2 * pk] + 1j * packed[2 * pk + 1] if cplex_enl == 2 else packed[pk]
    M = np.zeros_like(E)
    for j in range(nlmn):
        for i in range(nlmn):
            if i < j:
                M[j, i] = np.conj(E[i, j])
            elif i == j:
                M[j, i] = E[j, j].real if diag_real else E[j, j]
            else:
                M[j, i] = E[j, i]
    return M, E


def opernlc_general(gx, enl, sij, indlmn, nattyp, atindx1, paw_opt, nspinor=1, cplex_enl=1, lambda_=None):
    """opernlc_ylm_allwf with complex Hermitian D_ij (cplex_enl = 2, m_opernlc_ylm_allwf.F90:453-576) and / or spinor
    wavefunctions (nspinortot = 2: spin-diagonal blocks :578-655, off-diagonal blocks :660-737), complex gx (cplex = 2).
      gx  : (ndat, nspinor, nprojs) complex   == Fortran gx(2, nprojs, nspinor, ndat)
      enl : (nblk, natom, dimenl1) real       == Fortran enl(dimenl1, natom, nspinortot**2), blocks [uu, dd, ud, du],
            dimenl1 = cplex_enl * lmn2 ((re, im) pairs of the packed upper triangle when complex); nblk = 1 without spinors
      sij : (ntypat, lmn2) real packed
    Returns gxfac, gxfac_sij in the layout of gx.  "parity unpinned": no stored reference data reaches these branches; checked by
    reduction to the pinned real-D_ij path and by Hermiticity of the assembled operator (tests/test_oracle_invariants.py)."""
    gx = np.asarray(gx, dtype=np.complex128)
    ndat, nsp, nprojs = gx.shape
    assert nsp == nspinor
    enl = np.asarray(enl)
    if enl.ndim == 2:
        enl = enl[None]
    want_d = paw_opt in (1, 2, 4); want_s = paw_opt in (3, 4)
    gxfac = np.zeros_like(gx); gxs = np.zeros_like(gx) if want_s else None
    shift = 0; iatm = 0
    for t in range(indlmn.shape[0]):
        nlmn = nlmn_of_types(indlmn)[t]
        S = _unpack_sym(sij[t], nlmn) if paw_opt in (2, 3, 4) else None
        for ia in range(int(nattyp[t])):
            sl = slice(shift, shift + nlmn)
            ie = atindx1[iatm + ia]
            if want_d:
                for isp in range(nspinor):
                    M, _ = _hermitian_from_packed(enl[isp, ie], nlmn, cplex_enl)
                    for idat in range(ndat):
                        Md = M - (lambda_[idat] * S if paw_opt == 2 else 0.0)
                        gxfac[idat, isp, sl] += Md @ gx[idat, isp, sl]
                if nspinor == 2:
                    # :660-737  ispinor = 1: gxfac(j, dn) += sum_{i<=j} conj(E_ud[i,j]) gx(i, up); gxfac(j, up) += sum_{i>j} E_ud[j,i] gx(i, dn)
                    #           ispinor = 2: the same with up <-> dn and E_du
                    for isp in range(2):
                        jsp = 1 - isp
                        _, E = _hermitian_from_packed(enl[2 + isp, ie], nlmn, cplex_enl)
                        lower = np.conj(np.triu(E)).T                    # [j, i] = conj(E[i, j]) for i <= j
                        upper = np.triu(E, 1)                            # [j, i] = E[j, i] for i > j
                        for idat in range(ndat):
                            gxfac[idat, jsp, sl] += lower @ gx[idat, isp, sl]
                            gxfac[idat, isp, sl] += upper @ gx[idat, jsp, sl]
            if want_s:
                for isp in range(nspinor):
                    gxs[:, isp, sl] = gx[:, isp, sl] @ S.T
            shift += nlmn
        iatm += int(nattyp[t])
    return gxfac, gxs


def gemm_nonlop_general(P, vectin, enl, sij, indlmn, nattyp, atindx1, paw_opt, nspinor=1, cplex_enl=1, lambda_=None):
    """gemm_nonlop choice 1, signs 2, istwf_k = 1 with complex D_ij and / or spinors: vectin (ndat, nspinor, npw).
    Returns (vectout, svectout)."""
    v = np.asarray(vectin, dtype=np.complex128)
    ndat, nsp, npw = v.shape
    gx = (v.reshape(ndat * nsp, npw) @ np.conj(P).T).reshape(ndat, nsp, -1)
    gxfac, gxs = opernlc_general(gx, enl, sij, indlmn, nattyp, atindx1, paw_opt, nspinor, cplex_enl, lambda_)
    vectout = (gxfac.reshape(ndat * nsp, -1) @ P).reshape(ndat, nsp, npw) if paw_opt in (1, 2, 4) else None
    svectout = (gxs.reshape(ndat * nsp, -1) @ P).reshape(ndat, nsp, npw) + v if paw_opt in (3, 4) else None
    return vectout, svectout


def opernlb(P, gxfac, istwf_k):
    """vect(ndat, npw) = P . gxfac ; istwf_k>=2: (P_r z, P_i z) interleaved (m_opernlb_gemm.F90:700-706, 804-833)."""
    if istwf_k == 1:
        return gxfac @ P
    return (gxfac @ P.real) + 1j * (gxfac @ P.imag)


def gemm_nonlop(P, vectin, enl, sij, indlmn, nattyp, atindx1, istwf_k, choice=1, paw_opt=0, cpopt=-1,
                lambda_=None, projections=None, me_g0=1):
    """Returns (vectout, svectout, projections).  choice=1 signs=2; choice=0 only computes projections;
    choice=7: svectout = P . projections, s_projections=projections (m_gemm_nonlop.F90:855-861), without + vectin.
    cpopt>=2: projections are taken from the caller instead of being computed (m_gemm_nonlop.F90:719-734)."""
    vectin = np.atleast_2d(vectin)
    if cpopt >= 2:
        gx = np.asarray(projections)
    else:
        gx = opernla(P, vectin, istwf_k, me_g0)
    if choice == 0:
        return None, None, gx
    vectout = svectout = None
    if choice == 7:
        # s_projections = projections (m_gemm_nonlop.F90:855-861); vectin is NOT added for choice 7
        # (m_opernlb_gemm.F90:654: "if(choice /= 7 ...) svectout = svectout + vectin"); apply_invovl adds it itself
        # (m_invovl.F90:1031 sm1cwavef = cwavef + sm1cwavef)
        svectout = opernlb(P, gx, istwf_k)
        return None, svectout, gx
    gxfac, gxs = opernlc(gx, enl, sij, indlmn, nattyp, atindx1, paw_opt, lambda_)
    if paw_opt in (3, 4):
        svectout = opernlb(P, gxs, istwf_k) + vectin       # m_opernlb_gemm.F90:654-665
    if paw_opt in (0, 1, 2, 4):
        vectout = opernlb(P, gxfac, istwf_k)
    return vectout, svectout, gx


def nonlop_naive(ffnl, ph3d, indlmn, nattyp, atindx1, ucvol, vectin, enl, sij, paw_opt):
    """Independent statement used to cross-check gemm_nonlop for istwf_k=1:
    Vnl psi = sum_atoms sum_ij |p_i> D_ij <p_j|psi>, evaluated atom by atom with explicit loops."""
    vectin = np.atleast_2d(vectin)
    out = np.zeros_like(vectin); sout = vectin.copy()
    wt = FOUR_PI / np.sqrt(ucvol)
    ia3 = 0
    for t in range(indlmn.shape[0]):
        nlmn = nlmn_of_types(indlmn)[t]
        for ia in range(int(nattyp[t])):
            p = np.stack([wt * ffnl[t, i, 0, :] * (-1j) ** int(indlmn[t, i, 0]) * np.conj(ph3d[ia3]) for i in range(nlmn)])
            c = vectin @ np.conj(p).T                            # (ndat, nlmn)
            if paw_opt == 0:
                D = np.diag([enl[t, int(indlmn[t, i, 4]) - 1] for i in range(nlmn)])
            else:
                D = _unpack_sym(enl[atindx1[ia3]], nlmn)
            out += (c @ D.T) @ p
            if paw_opt in (3, 4):
                sout += (c @ _unpack_sym(sij[t], nlmn).T) @ p
            ia3 += 1
    return out, sout
