"""GPU parity: PAW with complex Hermitian D_ij (cplex_dij = 2) and with spinor wavefunctions (nspinor = 2: the four D_ij blocks
up-up, dn-dn, up-dn, dn-up; m_opernlc_ylm_allwf.F90:453-737) through the C-ABI getghc and gemm_nonlop, against the oracle
restatement (oracle/nonlop.py opernlc_general; reduces to the pinned real path, Hermitian -- tests/test_oracle_invariants.py)."""
import numpy as np
import pytest
from oracle import getghc as ogh, nonlop as onl
from problems import make_problem, rel_err_per_band
import abinit_b200 as ab
from abinit_b200 import api

pytestmark = pytest.mark.gpu


def _problem(nspinor, ndat=5, seed=3):
    p = make_problem(8.0, (8.0, 8.5, 7.5), (0.2, -0.1, 0.3), 1, ndat=ndat, seed=seed, natom_per_type=(2, 1), lmax_per_type=(1, 2), usepaw=1)
    rng = np.random.default_rng(seed)
    lmn2 = p.lmnmax * (p.lmnmax + 1) // 2
    nblk = 4 if nspinor == 2 else 1
    enl = 0.4 * rng.standard_normal((nblk, p.natom, 2 * lmn2))
    if nblk == 4:
        for j in range(p.lmnmax):
            pk = j * (j + 1) // 2 + j
            enl[3, :, 2 * pk] = enl[2, :, 2 * pk]; enl[3, :, 2 * pk + 1] = -enl[2, :, 2 * pk + 1]
    c = rng.standard_normal((ndat, nspinor, p.npw)) + 1j * rng.standard_normal((ndat, nspinor, p.npw))
    c /= (1.0 + np.minimum(p.kinpw, 1e3))[None, None, :]
    return p, np.ascontiguousarray(enl), np.ascontiguousarray(c)


def _vlocal(p, nvloc):
    if nvloc == 1:
        return p.vlocal
    n1, n2, n3 = p.ngfft
    i3, i2, i1 = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
    return np.ascontiguousarray(np.stack([p.vlocal, p.vlocal + 0.2 * np.cos(2 * np.pi * i1 / n1), 0.15 * np.sin(2 * np.pi * i2 / n2),
                                          0.1 * np.cos(2 * np.pi * (i3 / n3 - i1 / n1))]))


@pytest.mark.parametrize("nspinor,nvloc,sij_opt", [(1, 1, 1), (1, 1, 0), (2, 1, 1), (2, 4, 1), (2, 4, 0)])
def test_getghc_complex_dij_and_spinors(lib, nspinor, nvloc, sij_opt):
    p, enl, c = _problem(nspinor)
    vl = _vlocal(p, nvloc)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    rg, rs = ogh.getghc_paw_general(c, vl, p.kg, p.ngfft, p.kinpw, P, enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, nspinor=nspinor,
                                    cplex_enl=2, sij_opt=sij_opt)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 1, p.ucvol)
    h.set_nspinor(nspinor)
    if nvloc == 1:
        h.load_spin(vl, 1)
    else:
        h.load_spin_nvloc(vl, 4)
    h.load_enl(enl if nspinor == 2 else enl[0], p.sij)
    h.load_k(1, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    flat = np.ascontiguousarray(c.reshape(p.ndat * nspinor, p.npw))
    ghc = np.zeros_like(flat); gsc = np.zeros_like(flat)
    ab.getghc(-1, flat, None, ghc, gsc if sij_opt else None, h, None, None, None, p.ndat, sij_opt=sij_opt)
    assert rel_err_per_band(ghc, rg.reshape(flat.shape)) < 1e-11
    if sij_opt:
        assert rel_err_per_band(gsc, rs.reshape(flat.shape)) < 1e-11
    # the sentinel shell is exactly zero in both outputs
    dead = p.kinpw > 1e290
    assert np.all(ghc[:, dead] == 0) and (not sij_opt or np.all(gsc[:, dead] == 0))
    h.destroy()


def test_real_dij_handed_over_as_complex_matches_the_real_kernel(lib):
    """cplex_dij = 2 with zero imaginary parts == the real packed kernel, bit for bit up to rounding (same GEMMs)."""
    p, enl, c = _problem(1)
    real_d = np.ascontiguousarray(enl[0, :, 0::2])
    e2 = np.zeros_like(enl[0]); e2[:, 0::2] = real_d
    outs = []
    for e in (real_d, e2):
        h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 1, p.ucvol)
        h.load_spin(p.vlocal, 1); h.load_enl(e, p.sij); h.load_k(1, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
        flat = np.ascontiguousarray(c[:, 0])
        ghc = np.zeros_like(flat); gsc = np.zeros_like(flat)
        ab.getghc(-1, flat, None, ghc, gsc, h, None, None, None, p.ndat, sij_opt=1)
        outs.append((ghc, gsc)); h.destroy()
    assert rel_err_per_band(outs[1][0], outs[0][0]) < 1e-13 and rel_err_per_band(outs[1][1], outs[0][1]) < 1e-13


@pytest.mark.parametrize("nspinor", [1, 2])
def test_gemm_nonlop_entry_complex_dij_and_spinors(lib, nspinor):
    """The gemm_nonlop C-ABI entry itself (enl(dimenl1, natom, nspinortot**2), vectin(2, npw*nspinor*ndat)), paw_opt 4."""
    p, enl, c = _problem(nspinor, ndat=4, seed=9)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    ro, rs = onl.gemm_nonlop_general(P, c, enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, 4, nspinor=nspinor, cplex_enl=2)
    api.prep_projectors(1, p.npw, p.indlmn, p.nattyp, 1, p.ucvol, p.ffnl, p.ph3d)
    api.set_gemm_nonlop_ikpt(1)
    flat = np.ascontiguousarray(c.reshape(p.ndat * nspinor, p.npw))
    vout = np.zeros_like(flat); sout = np.zeros_like(flat)
    api.gemm_nonlop(p.atindx1, 1, -1, None, enl if nspinor == 2 else enl[0], p.indlmn, 1, None, p.natom, p.nattyp, p.ndat, p.npw, p.npw,
                    nspinor, p.ntypat, 4, p.sij, sout, flat, vout)
    assert rel_err_per_band(vout, ro.reshape(flat.shape)) < 1e-11
    assert rel_err_per_band(sout, rs.reshape(flat.shape)) < 1e-11


def test_apply_invovl_and_chebfiwf2_paw_spinors_vs_oracle(lib):
    """PAW with nspinor = 2 through the solver side: apply_invovl on npw*nspinor blocks (S has no spin structure: every spinor
    component is one column, m_invovl.F90:851-958) and ChebFi2-PAW (getBm1X = apply_invovl, B = S) against the oracle's chebfi_run
    driven by getghc_paw_general and the oracle's apply_invovl."""
    from oracle import xg as oxg, chebfi as och, invovl as oiv
    from abinit_b200 import xg
    nband = 6
    p, enl, c = _problem(2, ndat=nband, seed=5)
    vl = _vlocal(p, 4)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    iv = oiv.make_invovl(P, p.sij, p.indlmn, p.nattyp, 1)
    rows = 2 * p.npw

    def apply_h(x):
        g, s = ogh.getghc_paw_general(x.reshape(-1, 2, p.npw), vl, p.kg, p.ngfft, p.kinpw, P, enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                                      nspinor=2, cplex_enl=2, sij_opt=1)
        return g.reshape(-1, rows), s.reshape(-1, rows)

    def bm1(ax):
        s, _ = oiv.apply_invovl(P, iv, np.ascontiguousarray(ax.reshape(-1, p.npw)), 1)
        return s.reshape(-1, rows)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 1, p.ucvol)
    h.set_nspinor(2)
    h.load_spin_nvloc(vl, 4)
    h.load_enl(enl, p.sij)
    h.load_k(1, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    x0 = np.ascontiguousarray(c.reshape(nband, rows))
    # apply_invovl: C-ABI with the reference's argument list (nspinor = 2) vs the oracle, and S^-1 S = 1
    sm1 = np.zeros_like(x0)
    ab.apply_invovl(h, x0, sm1, None, p.npw, nband, nspinor=2)
    assert rel_err_per_band(sm1, bm1(x0)) < 1e-10
    _, sx = apply_h(sm1)
    live = p.kinpw < 1e290
    assert np.max(np.abs(sx.reshape(-1, p.npw)[:, live] - x0.reshape(-1, p.npw)[:, live])) < 1e-9 * np.max(np.abs(x0))
    x_ref = x0.copy(); cg = x0.copy()
    eig = np.zeros(nband); resid = np.zeros(nband)
    for it in range(2):
        w_ref, r_ref, x_ref = och.chebfi_run(apply_h, x_ref, oxg.SPACE_C, -1, p.ecut, nline=4, tolerance=1e-16, get_bm1x=bm1)
        xg.chebfiwf2(cg, eig, None, None, h, nband, p.npw, 2, resid, 1e-16, p.ecut, 4, bandpp=3)
        assert np.max(np.abs(eig - w_ref)) < 1e-9 * max(1.0, np.max(np.abs(w_ref))), (it, eig - w_ref)
        assert np.max(np.abs(resid - r_ref) / (np.abs(r_ref) + 1e-12)) < 1e-5, (it, resid, r_ref)
    h.destroy()
