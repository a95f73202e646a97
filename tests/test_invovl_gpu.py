"""GPU parity: PAW inverse overlap (make_invovl / apply_invovl) and ChebFi2-PAW through the C-ABI vs the oracle."""
import numpy as np
import pytest
from oracle import nonlop as onl, invovl as oiv, getghc as ogh, chebfi as och, xg as oxg
from problems import make_problem, rel_err_per_band
import abinit_b200 as ab
from abinit_b200 import xg

pytestmark = pytest.mark.gpu


def _ham(p):
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, p.sij)
    h.load_k(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    return h


@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0)), (3, (.5, 0, 0))])
@pytest.mark.parametrize("ndat", [1, 6])
def test_apply_invovl(lib, istwf_k, kpt, ndat):
    p = make_problem(7.0, 8.5, kpt, istwf_k, ndat=ndat, natom_per_type=(1, 2), lmax_per_type=(2, 1), usepaw=1)
    h = _ham(p)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    iv = oiv.make_invovl(P, p.sij, p.indlmn, p.nattyp, istwf_k)
    ref, ref_cprj = oiv.apply_invovl(P, iv, p.cwavef, istwf_k)
    out = np.zeros_like(p.cwavef)
    cplex = 2 if istwf_k == 1 else 1
    cprj = np.zeros((ndat, h.nprojs, cplex))
    ab.apply_invovl(h, p.cwavef, out, cprj, p.npw, ndat)
    assert rel_err_per_band(out, ref) < 1e-11
    got_cprj = cprj[..., 0] + 1j * cprj[..., 1] if cplex == 2 else cprj[..., 0]
    assert np.max(np.abs(got_cprj - ref_cprj)) < 1e-10 * max(1.0, np.max(np.abs(ref_cprj)))
    # invariant: S (S^-1 psi) = psi with S from the library's own gemm_nonlop (choice 1, paw_opt 3)
    back = np.zeros_like(out)
    ab.nonlop(1, -1, None, None, h, 0, None, None, ndat, 1, 3, 2, back, 0, out, None)
    assert rel_err_per_band(back, p.cwavef) < 1e-12
    h.destroy()


@pytest.mark.parametrize("istwf_k,kpt", [(1, (-.25, .5, 0)), (2, (0, 0, 0))])
def test_chebfiwf2_paw_vs_oracle(lib, istwf_k, kpt):
    """ChebFi2-PAW: generalised problem H x = e S x, getBm1X = apply_invovl, BX = S X from getghc(sij_opt=1)."""
    nband = 8
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(1, 2), lmax_per_type=(1, 1), usepaw=1,
                     filter_shell=False)
    h = _ham(p)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    iv = oiv.make_invovl(P, p.sij, p.indlmn, p.nattyp, istwf_k)
    space, me_g0 = (xg.SPACE_C, -1) if istwf_k == 1 else (xg.SPACE_CR, 1)

    def apply_h(c):
        ghc, gsc, _, _ = ogh.getghc(c, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                                    istwf_k=istwf_k, usepaw=1, sij_opt=1)
        return ghc, gsc
    bm1 = lambda c: oiv.apply_invovl(P, iv, c, istwf_k)[0]
    x_ref = p.cwavef.copy(); cg = p.cwavef.copy()
    eig = np.zeros(nband); resid = np.zeros(nband)
    for it in range(3):
        w_ref, r_ref, x_ref = och.chebfi_run(apply_h, x_ref, space, me_g0, p.ecut, nline=4, get_bm1x=bm1)
        xg.chebfiwf2(cg, eig, None, None, h, nband, p.npw, 1, resid, 1e-16, p.ecut, 4, bandpp=3)
        assert np.max(np.abs(eig - w_ref)) < 1e-9 * max(1.0, np.max(np.abs(w_ref))), (it, eig - w_ref)
        assert np.max(np.abs(resid - r_ref) / (np.abs(r_ref) + 1e-12)) < 1e-5, (it, resid, r_ref)
    # S-orthonormality of the returned block: X^H S X = 1
    _, sx, _ = onl.gemm_nonlop(P, cg, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k, choice=1, paw_opt=3)
    assert np.max(np.abs(oxg.gram(space, cg, sx, me_g0) - np.eye(nband))) < 1e-9
    h.destroy()
