"""GPU parity: gemm_nonlop (choice 0/1/7, signs=2) through the C-ABI vs the oracle, 1e-11 relative per band."""
import numpy as np
import pytest
from oracle import nonlop as onl
from problems import make_problem, rel_err_per_band
import abinit_b200 as ab
from abinit_b200 import api

pytestmark = pytest.mark.gpu
TOL = 1e-11


def _setup(lib, p, ikpt=1):
    api.prep_projectors(ikpt, p.npw, p.indlmn, p.nattyp, p.istwf_k, p.ucvol, p.ffnl, p.ph3d)
    api.set_gemm_nonlop_ikpt(ikpt)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    return P


def _apply(p, choice, paw_opt, cpopt=-1, projections=None, lambda_=None):
    cplex = 2 if p.istwf_k == 1 else 1
    nprojs = onl.count_nprojs(p.indlmn, p.nattyp)
    vout = np.zeros((p.ndat, p.npw), dtype=np.complex128)
    sout = np.zeros((p.ndat, p.npw), dtype=np.complex128)
    proj = np.zeros((p.ndat, nprojs, cplex)) if projections is None else projections
    api.gemm_nonlop(p.atindx1, choice, cpopt, proj, p.enl, p.indlmn, p.istwf_k, lambda_, p.natom, p.nattyp, p.ndat,
                    p.npw, p.npw, 1, p.ntypat, paw_opt, p.sij, sout, p.cwavef, vout)
    return vout, sout, proj


def _proj_as_complex(proj, cplex):
    return proj[..., 0] + 1j * proj[..., 1] if cplex == 2 else proj[..., 0]


@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0)), (3, (.5, 0, 0)), (9, (.5, .5, .5))])
@pytest.mark.parametrize("ndat", [1, 5, 12])
def test_nc_choice1(lib, istwf_k, kpt, ndat):
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=ndat, natom_per_type=(2, 1), lmax_per_type=(1, 2))
    P = _setup(lib, p)
    vout, _, proj = _apply(p, 1, 0, cpopt=0)
    rv, _, rgx = onl.gemm_nonlop(P, p.cwavef, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k, 1, 0)
    assert rel_err_per_band(vout, rv) < TOL
    assert rel_err_per_band(_proj_as_complex(proj, 2 if istwf_k == 1 else 1), rgx) < TOL


@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0))])
@pytest.mark.parametrize("paw_opt", [1, 2, 3, 4])
def test_paw(lib, istwf_k, kpt, paw_opt):
    p = make_problem(7.0, 8.5, kpt, istwf_k, ndat=6, natom_per_type=(1, 3), lmax_per_type=(2, 1), usepaw=1)
    P = _setup(lib, p)
    lam = np.linspace(-0.3, 0.4, p.ndat)
    vout, sout, _ = _apply(p, 1, paw_opt, lambda_=lam)
    rv, rs, _ = onl.gemm_nonlop(P, p.cwavef, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k, 1, paw_opt,
                                lambda_=lam)
    if rv is not None:
        assert rel_err_per_band(vout, rv) < TOL
    if rs is not None:
        assert rel_err_per_band(sout, rs) < TOL


def test_choice0_choice7_and_cpopt2(lib):
    p = make_problem(7.0, 8.5, (.1, .2, .3), 1, ndat=4, natom_per_type=(2,), lmax_per_type=(2,), usepaw=1)
    P = _setup(lib, p)
    _, _, proj = _apply(p, 0, 4, cpopt=0)                                   # projections only
    rgx = onl.opernla(P, p.cwavef, 1)
    assert rel_err_per_band(_proj_as_complex(proj, 2), rgx) < TOL
    _, sout7, _ = _apply(p, 7, 3)
    _, rs7, _ = onl.gemm_nonlop(P, p.cwavef, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, 1, 7, 3)
    assert rel_err_per_band(sout7, rs7) < TOL
    # cpopt=2: <p|c> taken from the caller's buffer (m_gemm_nonlop.F90:719-734): feed a *modified* buffer
    proj2 = np.ascontiguousarray(proj * 1.5)
    vout, sout, _ = _apply(p, 1, 4, cpopt=2, projections=proj2)
    rv, rs, _ = onl.gemm_nonlop(P, p.cwavef, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, 1, 1, 4, cpopt=2,
                                projections=_proj_as_complex(proj2, 2))
    assert rel_err_per_band(vout, rv) < TOL and rel_err_per_band(sout, rs) < TOL


def test_naive_per_atom_statement(lib):
    """gemm_nonlop == sum_a sum_ij |p_i> D_ij <p_j|psi> evaluated atom by atom (SURVEY 8c invariant (ii))."""
    p = make_problem(6.0, 8.0, (.25, 0, .1), 1, ndat=3, natom_per_type=(2, 2), lmax_per_type=(1, 1), usepaw=1)
    _setup(lib, p)
    vout, sout, _ = _apply(p, 1, 4)
    rv, rs = onl.nonlop_naive(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.atindx1 - 1, p.ucvol, p.cwavef, p.enl, p.sij, 4)
    assert rel_err_per_band(vout, rv) < TOL and rel_err_per_band(sout, rs) < TOL


def test_explicit_projectors_larger(lib):
    """Random explicit P (bench-style), sizes that exercise several M/N tiles, K tails, split-K and every CTA tile
    width (128/64/32 effective columns) of the real and complex TN/NN kernels."""
    rng = np.random.default_rng(3)
    for istwf_k, npw, nprojs, ndat in ((2, 4097, 300, 70), (1, 2051, 259, 33), (2, 3001, 200, 40), (1, 1500, 130, 20),
                                       (1, 1501, 131, 10), (2, 1777, 515, 9), (1, 900, 1030, 130)):
        p = make_problem(3.0, 6.0, (0, 0, 0) if istwf_k == 2 else (.1, .2, .3), istwf_k, ndat=1)
        P = (rng.standard_normal((nprojs, npw)) + 1j * rng.standard_normal((nprojs, npw))) / np.sqrt(npw)
        c = rng.standard_normal((ndat, npw)) + 1j * rng.standard_normal((ndat, npw))
        if istwf_k == 2:
            P[:, 0] = P[:, 0].real; c[:, 0] = c[:, 0].real
        indlmn = np.zeros((1, 1, 6), dtype=np.int32); indlmn[0, 0] = (0, 0, 1, 1, 1, 1)
        nattyp = np.array([nprojs], dtype=np.int32); atindx1 = np.arange(1, nprojs + 1, dtype=np.int32)
        enl = rng.standard_normal((1, 1))
        api.set_projectors(2, npw, nprojs, istwf_k, np.ascontiguousarray(P))
        api.set_gemm_nonlop_ikpt(2)
        vout = np.zeros((ndat, npw), dtype=np.complex128)
        api.gemm_nonlop(atindx1, 1, -1, None, enl, indlmn, istwf_k, None, nprojs, nattyp, ndat, npw, npw, 1, 1, 0, None,
                        None, np.ascontiguousarray(c), vout)
        rv, _, _ = onl.gemm_nonlop(P, c, enl, None, indlmn, nattyp, atindx1 - 1, istwf_k, 1, 0)
        assert rel_err_per_band(vout, rv) < TOL


def test_mkffnl_on_device_matches_oracle(lib):
    """mkffnl (ider=0, useylm=1) on the device vs the oracle's restatement (pinned through the tbase3_1 SCF): psp8 form-factor
    splines of the Si-2 fixture at both k-points; the device-resident result feeds load_k directly."""
    import os
    torch = pytest.importorskip("torch")
    from oracle import scf
    from oracle.psp8 import ClampedSpline
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "si2_tbase3.npz"))
    s = scf.setup_from_fixture(fx)
    qg = np.array(fx["qgrid"]); tab = np.array(fx["ffspl_tab"]); yp = np.array(fx["ffspl_yp"])
    nln, mq = tab.shape
    ffspl = np.zeros((1, nln, 2, mq))
    for i in range(nln):
        cs = ClampedSpline(qg, tab[i], yp[i, 0], yp[i, 1])
        ffspl[0, i, 0] = tab[i]; ffspl[0, i, 1] = cs.cs(qg, 2)          # value and second derivative tables = ffspl(:,1:2,iln,1)
    indlmn = s.indlmn                                                  # (1, lmnmax, 6)
    lmnmax = indlmn.shape[1]; lmax = int(indlmn[0, :, 0].max())
    for ik, kpt in enumerate(s.kpts):
        kg = s.kg[ik]; npw = kg.shape[1]
        cart = s.gprimd @ (kg.astype(float) + np.asarray(kpt)[:, None])
        ylm = np.ascontiguousarray(scf.real_ylm(cart, lmax).T)         # (mpsang^2, npw) == Fortran ylm(npw, mpsang^2)
        ffnl_dev = torch.zeros((1, lmnmax, 1, npw), dtype=torch.float64, device="cuda")
        api.mkffnl(nln, 1, s.ekb, ffnl_dev, np.ascontiguousarray(ffspl), None, s.gprimd, 0, 0, indlmn, np.ascontiguousarray(kg.T.astype(np.int32)),
                   None, kpt, lmnmax, nln, lmax + 1, mq, 0, npw, 1, None, qg, None, 0, 1, ylm)
        ref = s.ffnl[ik]                                               # (1, lmnmax, 1, npw) from oracle/scf.mkffnl
        got = ffnl_dev.cpu().numpy()
        assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))
        # the device array goes straight into the Hamiltonian handle: same H psi as with the oracle's host ffnl
        h = ab.Hamiltonian(s.ngfft, s.xred.shape[1], 1, lmnmax, indlmn, s.nattyp, s.atindx1 + 1, 0, s.ucvol)
        h.load_spin(np.ascontiguousarray(s.vpsp), 1); h.load_enl(s.ekb, None)
        h.load_k_xred(1, np.ascontiguousarray(kg.T), s.kinpw[ik], ffnl_dev, kpt, np.ascontiguousarray(s.xred.T))
        c = np.ascontiguousarray(np.eye(npw, dtype=np.complex128)[:6])
        out = np.zeros_like(c)
        ab.getghc(-1, c, None, out, None, h, None, None, None, 6)
        ref_h = scf.apply_h_oracle(s)(ik, s.vpsp, c)
        assert rel_err_per_band(out, ref_h) < TOL
        h.destroy()


@pytest.mark.parametrize("mpsang", [1, 3, 4])
def test_initylmg_on_device_matches_oracle(lib, mpsang):
    """initylmg (optder=0, one k-point) on the device vs the oracle restatement of m_initylmg.F90 / ass_leg_pol, incl. the
    k+G = 0 point and points on the z axis; the spherical-harmonic addition theorem as an independent check."""
    from oracle import scf
    rng = np.random.default_rng(3)
    npw = 3001
    kg = rng.integers(-9, 10, size=(3, npw)); kg[:, 0] = 0; kg[:, 1] = (0, 0, 4); kg[:, 2] = (0, 0, -3)
    gprimd = np.diag([0.11, 0.09, 0.13]) + 0.01 * rng.standard_normal((3, 3))
    for kpt in ((0.0, 0.0, 0.0), (0.1, -0.2, 0.3)):
        ref = scf.initylmg_k(kg, kpt, gprimd, mpsang)
        ylm = np.zeros((mpsang * mpsang, npw))
        api.initylmg_k(gprimd, np.ascontiguousarray(kg.T.astype(np.int32)), kpt, mpsang, npw, ylm)
        assert np.max(np.abs(ylm - ref)) < 1e-13
        for l in range(mpsang):
            s = np.sum(ylm[l * l:(l + 1) ** 2] ** 2, axis=0)
            nz = np.linalg.norm(gprimd @ (kg + np.asarray(kpt)[:, None]), axis=0) > 1e-10
            assert np.allclose(s[nz], (2 * l + 1) / (4 * np.pi), atol=1e-13)


@pytest.mark.parametrize("istwf_k,kpt", [(2, (0, 0, 0)), (1, (.1, .2, .3))])
@pytest.mark.parametrize("ndat", [3, 10, 19, 38, 76, 100])
def test_ragged_band_blocks_with_many_projectors(lib, istwf_k, kpt, ndat):
    """Band blocks that do not fill their column tile, on a projector count large enough (>= 2048) for the ragged GEMM variants
    (column blocks spread over the SM sub-partitions, DMMAs of empty 8-column fragments skipped): choice 1 NC and PAW paw_opt 4."""
    for usepaw in (0, 1):
        p = make_problem(8.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=ndat, natom_per_type=(70, 60), lmax_per_type=(2, 2), usepaw=usepaw, seed=ndat)
        assert onl.count_nprojs(p.indlmn, p.nattyp) >= 2048
        P = _setup(lib, p)
        paw_opt = 4 if usepaw else 0
        rv, rs, rgx = onl.gemm_nonlop(P, p.cwavef, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k, 1, paw_opt)
        for knob in (0, 47, 63):               # no ragged variant / the default set / every variant (developer knob nonlop_rag)
            api.set_tuning("nonlop_rag", knob)
            try:
                vout, sout, proj = _apply(p, 1, paw_opt, cpopt=0)
            finally:
                api.set_tuning("nonlop_rag", 47)
            assert rel_err_per_band(vout, rv) < TOL, knob
            assert rel_err_per_band(_proj_as_complex(proj, 2 if istwf_k == 1 else 1), rgx) < TOL, knob
            if usepaw:
                assert rel_err_per_band(sout, rs) < TOL, knob
