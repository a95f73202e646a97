"""getghc: <G|H|C> = local + kinetic + non-local (oracle; test infrastructure only).

Restates src/66_wfs/m_getghc.F90:182-1447 for nspinor=1, nvloc=1, k == k', no Fock / mGGA / nucdip:
  local part     :536-551   fourwf(option=2)
  non-local part :1042-1074 nonlop(choice=1, signs=2, paw_opt = usepaw ; sij_opt/=0 -> sij_opt+3)
  assembly       :1266-1280 ghc = ghc + kinpw*cwavef + gvnlxc  (0 where kinpw >= huge*1e-11; gsc zeroed too)
  type_calc=1 filter :1003-1031
PINNED (NC, istwf_k=1) on the reference's tbase3_1 SCF numbers through oracle/scf.py (see oracle/__init__.py)."""
from __future__ import annotations
import numpy as np
from .fourwf import fourwf
from .nonlop import gemm_nonlop
from .gsphere import KIN_FILTER


def getghc(cwavef, vlocal, kg, ngfft, kinpw, P, enl, sij, indlmn, nattyp, atindx1, istwf_k=1,
           usepaw=0, sij_opt=0, cpopt=-1, type_calc=0, lambda_=None, ghc_in=None, me_g0=1,
           filter_dilatmx_loc=True, workers=None, local_impl="full"):
    """Returns (ghc, gsc, gvnlxc, projections).  local_impl: "full" = plain 3-D FFT (the checker of the parity tests),
    "pad" = the reference's zero-padded passes + Gamma-point band pairing (fourwf_pad.py; same result, the CPU timing arm)."""
    cwavef = np.atleast_2d(cwavef)
    ghc = np.zeros_like(cwavef) if ghc_in is None else np.array(ghc_in, dtype=np.complex128, copy=True)
    gsc = None; gvnlxc = np.zeros_like(cwavef); proj = None
    if type_calc in (0, 1, 3):
        cplex = 2 if np.iscomplexobj(vlocal) else 1
        if local_impl == "pad":
            from .fourwf_pad import fourwf_option2_padded
            ghc = fourwf_option2_padded(cplex, vlocal, cwavef, kg, ngfft, istwf_k, me_g0=me_g0, workers=workers)
        else:
            ghc, _, _ = fourwf(cplex, vlocal, cwavef, None, kg, kg, ngfft, 2, istwf_k, me_g0=me_g0, workers=workers)
        if type_calc == 1 and filter_dilatmx_loc:
            ghc[:, kinpw > KIN_FILTER] = 0.0
    if type_calc in (0, 2):
        paw_opt = usepaw
        if sij_opt != 0:
            paw_opt = sij_opt + 3
        cpopt_here = cpopt if usepaw == 1 else -1
        gvnlxc, gsc, proj = gemm_nonlop(P, cwavef, enl, sij, indlmn, nattyp, atindx1, istwf_k, choice=1,
                                        paw_opt=paw_opt, cpopt=cpopt_here, lambda_=lambda_, me_g0=me_g0)
    if type_calc in (0, 2, 3):
        ok = kinpw < KIN_FILTER
        kin = np.where(ok, kinpw, 0.0)
        ghc = np.where(ok[None, :], ghc + kin[None, :] * cwavef + gvnlxc, 0.0)
        if sij_opt == 1 and gsc is not None:
            gsc = np.where(ok[None, :], gsc, 0.0)
    return ghc, gsc, gvnlxc, proj


def getghc_paw_general(cwavef, vlocal, kg, ngfft, kinpw, P, enl, sij, indlmn, nattyp, atindx1, nspinor=1, cplex_enl=1, sij_opt=1,
                       workers=None):
    """getghc, type_calc 0, istwf_k = 1, PAW with complex Hermitian D_ij and / or spinor wavefunctions (collinear local potential
    nvloc = 1 or the four-component nvloc = 4): local part as getghc / getghc_spinor, non-local part through
    nonlop.gemm_nonlop_general.  cwavef: (ndat, nspinor, npw).  Returns (ghc, gsc)."""
    from .nonlop import gemm_nonlop_general
    cw = np.asarray(cwavef, dtype=np.complex128)
    ndat, nsp, npw = cw.shape
    v = np.asarray(vlocal)
    f = lambda vv, c: fourwf(2 if np.iscomplexobj(vv) else 1, vv, np.ascontiguousarray(c), None, kg, kg, ngfft, 2, 1, workers=workers)[0]
    ghc = np.zeros_like(cw)
    if nsp == 1:
        ghc[:, 0] = f(v, cw[:, 0])
    elif v.ndim == 3:
        ghc[:, 0] = f(v, cw[:, 0]); ghc[:, 1] = f(v, cw[:, 1])
    else:
        ghc[:, 0] = f(v[0], cw[:, 0]) + f(v[2] + 1j * v[3], cw[:, 1])
        ghc[:, 1] = f(v[2] - 1j * v[3], cw[:, 0]) + f(v[1], cw[:, 1])
    gv, gs = gemm_nonlop_general(P, cw, enl, sij, indlmn, nattyp, atindx1, 4 if sij_opt == 1 else 1, nsp, cplex_enl)
    ok = kinpw < KIN_FILTER
    kin = np.where(ok, kinpw, 0.0)
    ghc = np.where(ok[None, None, :], ghc + kin[None, None, :] * cw + gv, 0.0)
    gsc = np.where(ok[None, None, :], gs, 0.0) if gs is not None else None
    return ghc, gsc


def getghc_spinor(cwavef, vlocal, kg, ngfft, kinpw, P, enl, indlmn, nattyp, atindx1, type_calc=0, workers=None):
    """nspinor = 2 (norm-conserving, istwf_k = 1, no spin-orbit): restates the spinor branches of m_getghc.F90
      nvloc = 1  :555-653   the same real potential on both spinor components
      nvloc = 4  :655-830   ghc_up = V11 psi_up + (V3 + i V4) psi_dn ; ghc_dn = (V3 - i V4) psi_up + V22 psi_dn
      kinetic assembly :1266-1280 and nonlop per spinor component (opernlc NC: ekb(:,:,ispinor) identical for both).
    cwavef: (ndat, 2, npw) == Fortran cwavef(2, npw*nspinor*ndat); vlocal: (n3,n2,n1) or (4,n3,n2,n1) [V11, V22, Re V12, Im V12]."""
    cw = np.asarray(cwavef)
    ndat, _, npw = cw.shape
    up, dn = np.ascontiguousarray(cw[:, 0]), np.ascontiguousarray(cw[:, 1])
    v = np.asarray(vlocal)
    ghc = np.zeros_like(cw); gv = np.zeros_like(cw)
    f = lambda vv, c: fourwf(2 if np.iscomplexobj(vv) else 1, vv, c, None, kg, kg, ngfft, 2, 1, workers=workers)[0]
    if type_calc in (0, 1, 3):
        if v.ndim == 3:
            ghc[:, 0] = f(v, up); ghc[:, 1] = f(v, dn)
        else:
            g1 = f(v[0], up); g2 = f(v[1], dn)
            g3 = f(v[2] - 1j * v[3], up); g4 = f(v[2] + 1j * v[3], dn)
            ghc[:, 0] = g1 + g4; ghc[:, 1] = g3 + g2
        if type_calc == 1:
            ghc[:, :, kinpw > KIN_FILTER] = 0.0
    if type_calc in (0, 2):
        flat = cw.reshape(ndat * 2, npw)
        gvf, _, _ = gemm_nonlop(P, flat, enl, None, indlmn, nattyp, atindx1, 1, choice=1, paw_opt=0, cpopt=-1)
        gv = gvf.reshape(ndat, 2, npw)
    if type_calc in (0, 2, 3):
        ok = kinpw < KIN_FILTER
        kin = np.where(ok, kinpw, 0.0)
        ghc = np.where(ok[None, None, :], ghc + kin[None, None, :] * cw + gv, 0.0)
    return ghc, gv
