"""GPU parity: the eigensolver drivers on nspinor = 2 blocks (norm-conserving, istwf_k = 1; rows = npw*nspinor as in
m_chebfiwf.F90:227 / m_lobpcgwf.F90:192): chebfiwf2 and lobpcgwf2 (one and several blocks) through the C-ABI vs the oracle
solvers driven by the oracle's spinor getghc (collinear nvloc = 1 and non-collinear nvloc = 4 potentials)."""
import numpy as np
import pytest
from oracle import xg as oxg, chebfi as och, lobpcg as olb, getghc as ogh, nonlop as onl
from problems import make_problem
import abinit_b200 as ab
from abinit_b200 import xg

pytestmark = pytest.mark.gpu


def _setup(nband, nvloc):
    p = make_problem(7.0, (8.0, 9.0, 7.5), (.1, .2, .3), 1, ndat=2 * nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False)
    n1, n2, n3 = p.ngfft
    if nvloc == 1:
        vl = p.vlocal
    else:
        i3, i2, i1 = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
        vl = np.ascontiguousarray(np.stack([p.vlocal, p.vlocal + 0.2 * np.cos(2 * np.pi * i1 / n1), 0.15 * np.sin(2 * np.pi * i2 / n2),
                                            0.1 * np.cos(2 * np.pi * (i3 / n3 - i1 / n1))]))
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 0, p.ucvol)
    h.set_nspinor(2)
    h.load_spin_nvloc(vl, nvloc)
    h.load_enl(p.enl, None)
    h.load_k(1, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    rows = 2 * p.npw

    def apply_h(c):                      # blocks (ncols, npw*nspinor): spinor components of a band are consecutive
        out, _ = ogh.getghc_spinor(c.reshape(-1, 2, p.npw), vl, p.kg, p.ngfft, p.kinpw, P, p.enl, p.indlmn, p.nattyp, p.atindx1 - 1)
        return out.reshape(-1, rows), c.copy()

    def enl_ref(c):
        _, gv = ogh.getghc_spinor(c.reshape(-1, 2, p.npw), vl, p.kg, p.ngfft, p.kinpw, P, p.enl, p.indlmn, p.nattyp, p.atindx1 - 1,
                                  type_calc=2)
        return np.real(np.sum(np.conj(c) * gv.reshape(-1, rows), axis=1))
    x0 = np.ascontiguousarray(p.cwavef.reshape(nband, rows))
    return p, h, apply_h, enl_ref, x0


@pytest.mark.parametrize("nvloc", [1, 4])
def test_chebfiwf2_nspinor2_vs_oracle(lib, nvloc):
    nband = 8
    p, h, apply_h, enl_ref, x0 = _setup(nband, nvloc)
    x_ref = x0.copy(); cg = x0.copy()
    eig = np.zeros(nband); resid = np.zeros(nband); enl = np.zeros(nband)
    for it in range(3):
        w_ref, r_ref, x_ref = och.chebfi_run(apply_h, x_ref, oxg.SPACE_C, -1, p.ecut, nline=5, tolerance=1e-16)
        xg.chebfiwf2(cg, eig, None, enl, h, nband, p.npw, 2, resid, 1e-16, p.ecut, 5, bandpp=3)
        assert np.max(np.abs(eig - w_ref)) < 1e-9 * max(1.0, np.max(np.abs(w_ref))), (it, eig - w_ref)
        assert np.max(np.abs(resid - r_ref) / (np.abs(r_ref) + 1e-12)) < 1e-5, (it, resid, r_ref)
        ov = np.abs(np.diag(oxg.gram(oxg.SPACE_C, x_ref, cg, -1)))
        gaps = np.min(np.abs(np.subtract.outer(w_ref, w_ref)) + np.eye(nband), axis=1)
        assert np.max(np.abs(ov[gaps > 1e-4] - 1.0)) < 1e-7, (it, ov)
        assert np.max(np.abs(enl - enl_ref(cg))) < 1e-10
    h.destroy()


@pytest.mark.parametrize("nvloc,nblock", [(1, 1), (4, 1), (4, 2)])
def test_lobpcgwf2_nspinor2_vs_oracle(lib, nvloc, nblock):
    nband = 6
    p, h, apply_h, enl_ref, x0 = _setup(nband, nvloc)
    pcon = np.tile(olb.build_pcon(p.kinpw), 2)                   # xgBlock_apply_diag(W, pcond, nspinor)
    x_ref = x0.copy(); cg = x0.copy()
    eig = np.zeros(nband); resid = np.zeros(nband); enl = np.zeros(nband)
    for it in range(3):
        w_ref, r_ref, x_ref = olb.lobpcg_run(apply_h, x_ref, pcon, oxg.SPACE_C, -1, nline=4, tolerance=1e-30, nblock=nblock)
        xg.lobpcgwf2(cg, eig, None, enl, h, nband, p.npw, 2, resid, 1e-30, 4, nblock_lobpcg=nblock)
        assert np.max(np.abs(eig - w_ref)) < 1e-9 * max(1.0, np.max(np.abs(w_ref))), (it, eig - w_ref)
        assert np.max(np.abs(resid - r_ref) / (np.abs(r_ref) + 1e-13)) < 1e-4, (it, resid, r_ref)
        assert np.max(np.abs(enl - enl_ref(cg))) < 1e-10
    bx = cg
    assert np.max(np.abs(oxg.gram(oxg.SPACE_C, cg, bx, -1) - np.eye(nband))) < 1e-9
    h.destroy()
