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
           filter_dilatmx_loc=True, workers=None):
    """Returns (ghc, gsc, gvnlxc, projections)."""
    cwavef = np.atleast_2d(cwavef)
    ghc = np.zeros_like(cwavef) if ghc_in is None else np.array(ghc_in, dtype=np.complex128, copy=True)
    gsc = None; gvnlxc = np.zeros_like(cwavef); proj = None
    if type_calc in (0, 1, 3):
        cplex = 2 if np.iscomplexobj(vlocal) else 1
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
