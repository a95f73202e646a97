"""GPU parity: fused getghc through the C-ABI vs the oracle; invariants at larger sizes."""
import numpy as np
import pytest
from oracle import getghc as ogh, nonlop as onl, gsphere as g
from problems import make_problem, rel_err_per_band
import abinit_b200 as ab

pytestmark = pytest.mark.gpu
TOL = 1e-11


def _ham(p):
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, p.cplex)
    h.load_enl(p.enl, p.sij)
    h.load_k(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    return h


def _oracle(p, **kw):
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    return ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                      istwf_k=p.istwf_k, usepaw=p.usepaw, **kw)


@pytest.mark.parametrize("istwf_k,kpt", [(1, (-.25, .5, 0)), (2, (0, 0, 0)), (7, (.5, .5, 0))])
@pytest.mark.parametrize("type_calc", [0, 1, 2, 3])
def test_nc_type_calc(lib, istwf_k, kpt, type_calc):
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=5, natom_per_type=(2,), lmax_per_type=(1,))
    h = _ham(p)
    rng = np.random.default_rng(5)
    ghc0 = rng.standard_normal((p.ndat, p.npw)) + 1j * rng.standard_normal((p.ndat, p.npw))
    ghc = ghc0.copy(); gv = np.zeros_like(ghc)
    ab.getghc(-1, p.cwavef, None, ghc, None, h, gv, None, None, p.ndat, type_calc=type_calc)
    r_ghc, _, r_gv, _ = _oracle(p, type_calc=type_calc, ghc_in=ghc0)
    assert rel_err_per_band(ghc, r_ghc) < TOL
    if type_calc in (0, 2):
        assert rel_err_per_band(gv, r_gv) < TOL
    # sentinel filter reproduced bit-exactly (m_getghc.F90:1272-1277)
    if type_calc != 1:
        assert np.all(ghc[:, p.kinpw >= g.KIN_FILTER] == 0.0)
    h.destroy()


@pytest.mark.parametrize("sij_opt", [0, 1, -1])
@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0))])
def test_paw_sij_opt(lib, sij_opt, istwf_k, kpt):
    p = make_problem(7.0, 8.5, kpt, istwf_k, ndat=4, natom_per_type=(1, 2), lmax_per_type=(2, 1), usepaw=1)
    h = _ham(p)
    lam = np.linspace(-0.2, 0.3, p.ndat)
    ghc = np.zeros((p.ndat, p.npw), dtype=np.complex128); gsc = np.zeros_like(ghc); gv = np.zeros_like(ghc)
    cplex = 2 if istwf_k == 1 else 1
    prj = np.zeros((p.ndat, h.nprojs, cplex))
    ab.getghc(0, p.cwavef, prj, ghc, gsc, h, gv, lam, None, p.ndat, sij_opt=sij_opt)
    r_ghc, r_gsc, r_gv, r_prj = _oracle(p, sij_opt=sij_opt, cpopt=0, lambda_=lam)
    assert rel_err_per_band(ghc, r_ghc) < TOL
    assert rel_err_per_band(gv, r_gv) < TOL
    if sij_opt == 1:
        assert rel_err_per_band(gsc, r_gsc) < TOL
        assert np.all(gsc[:, p.kinpw >= g.KIN_FILTER] == 0.0)
    prjc = prj[..., 0] + 1j * prj[..., 1] if cplex == 2 else prj[..., 0]
    assert rel_err_per_band(prjc, r_prj) < TOL
    h.destroy()


@pytest.mark.parametrize("ndat", [2, 5])
@pytest.mark.parametrize("usepaw", [0, 1])
def test_gamma_two_bands_per_transform(lib, ndat, usepaw):
    """istwf_k=2 on a plane-stage box: two bands ride one complex transform (double_rfft_trick,
    m_getghc.F90:1999-2171); even and odd band counts, every type_calc, PAW gsc zeroing, and the unpacked path."""
    from abinit_b200 import api
    p = make_problem(7.0, 9.0, (0, 0, 0), 2, ndat=ndat, ngfft=(36, 40, 40), natom_per_type=(2, 1), lmax_per_type=(1, 2),
                     usepaw=usepaw)
    h = _ham(p)
    for pack in (1, 0):
        api.set_tuning("pack2", pack)
        try:
            for type_calc in (0, 1, 3):
                ghc = np.zeros((p.ndat, p.npw), dtype=np.complex128); gsc = np.zeros_like(ghc)
                sij_opt = 1 if (usepaw and type_calc == 0) else 0
                ab.getghc(-1, p.cwavef, None, ghc, gsc if sij_opt else None, h, None, None, None, p.ndat, sij_opt=sij_opt,
                          type_calc=type_calc)
                r_ghc, r_gsc, _, _ = _oracle(p, type_calc=type_calc, sij_opt=sij_opt)
                assert rel_err_per_band(ghc, r_ghc) < TOL
                assert np.all(ghc[:, 0].imag == 0.0)                       # Im c(G=0) = 0 (m_ompgpu_fourwf.F90:536-543)
                if type_calc != 1:
                    assert np.all(ghc[:, p.kinpw >= g.KIN_FILTER] == 0.0)
                if sij_opt:
                    assert rel_err_per_band(gsc, r_gsc) < TOL
        finally:
            api.set_tuning("pack2", 1)
    h.destroy()


def test_gvnlxc_absent_and_generic_fourwf(lib):
    """gvnlxc of size<=1 -> internal temporary (m_getghc.F90:320-331); forcing the generic fourwf gives the same."""
    p = make_problem(7.0, 8.5, (.1, .2, .3), 1, ndat=3)
    h = _ham(p)
    a = np.zeros((p.ndat, p.npw), dtype=np.complex128); b = np.zeros_like(a)
    ab.getghc(-1, p.cwavef, None, a, None, h, None, None, None, p.ndat)
    ab.api.L().abi_b200_fourwf_set_impl(1)
    ab.getghc(-1, p.cwavef, None, b, None, h, None, None, None, p.ndat)
    ab.api.L().abi_b200_fourwf_set_impl(0)
    r_ghc, _, _, _ = _oracle(p)
    assert rel_err_per_band(a, r_ghc) < TOL and rel_err_per_band(b, r_ghc) < TOL
    h.destroy()


def test_hermiticity_medium(lib):
    """<phi|H psi> = <H phi|psi> at a size the oracle would take long on (size-independent property)."""
    p = make_problem(14.0, 14.0, (.1, .2, .3), 1, ndat=8, natom_per_type=(6, 4), lmax_per_type=(2, 1),
                     filter_shell=False)
    h = _ham(p)
    ghc = np.zeros((p.ndat, p.npw), dtype=np.complex128)
    ab.getghc(-1, p.cwavef, None, ghc, None, h, None, None, None, p.ndat)
    A = np.conj(p.cwavef) @ ghc.T
    assert np.abs(A - A.conj().T).max() < 1e-11 * np.abs(A).max()
    h.destroy()


def test_istwfk2_equals_istwfk1_on_completed_sphere(lib):
    """Gamma point: the istwf_k=2 result equals the istwf_k=1 result on the time-reversal-completed sphere
    (SURVEY 8c invariant (ii)); same operator (problems.py draws it from a seed independent of npw)."""
    p2 = make_problem(7.0, 8.5, (0, 0, 0), 2, ndat=3, filter_shell=False)
    p1 = make_problem(7.0, 8.5, (0, 0, 0), 1, ndat=3, filter_shell=False)
    lut = {tuple(k): i for i, k in enumerate(p2.kgF.tolist())}
    c1 = np.zeros((3, p1.npw), dtype=np.complex128)
    for i, k in enumerate(p1.kgF.tolist()):
        if tuple(k) in lut:
            c1[:, i] = p2.cwavef[:, lut[tuple(k)]]
        else:
            c1[:, i] = np.conj(p2.cwavef[:, lut[tuple(-x for x in k)]])     # c(-G) = conj(c(G))
    p1.cwavef = np.ascontiguousarray(c1)
    h1, h2 = _ham(p1), _ham(p2)
    g1 = np.zeros((3, p1.npw), dtype=np.complex128); g2 = np.zeros((3, p2.npw), dtype=np.complex128)
    ab.getghc(-1, p1.cwavef, None, g1, None, h1, None, None, None, 3)
    ab.getghc(-1, p2.cwavef, None, g2, None, h2, None, None, None, 3)
    sel = np.array([i for i, k in enumerate(p1.kgF.tolist()) if tuple(k) in lut])
    tgt = np.array([lut[tuple(p1.kgF[i].tolist())] for i in sel])
    assert rel_err_per_band(g1[:, sel], g2[:, tgt]) < 1e-11
    h1.destroy(); h2.destroy()


@pytest.mark.parametrize("istwf_k,kpt,usepaw", [(1, (-.25, .5, 0), 0), (2, (0, 0, 0), 1), (6, (0, .5, 0), 0)])
def test_load_k_xred_builds_ph3d_on_device(lib, istwf_k, kpt, usepaw):
    """load_k_xred (ph1d3d fused into prep_projectors on the device) == load_k with the host ph3d, and == the oracle."""
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=4, natom_per_type=(2, 1), lmax_per_type=(2, 1), usepaw=usepaw)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, p.sij)
    h.load_k_xred(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.kpt, np.ascontiguousarray(p.xred.T), me_g0=1)
    a = np.zeros((p.ndat, p.npw), dtype=np.complex128)
    ab.getghc(-1, p.cwavef, None, a, None, h, None, None, None, p.ndat)
    h2 = _ham(p)
    b = np.zeros_like(a)
    ab.getghc(-1, p.cwavef, None, b, None, h2, None, None, None, p.ndat)
    r_ghc, _, _, _ = _oracle(p)
    assert rel_err_per_band(a, b) < 1e-12
    assert rel_err_per_band(a, r_ghc) < TOL
    h.destroy(); h2.destroy()


@pytest.mark.parametrize("nvloc", [1, 4])
@pytest.mark.parametrize("type_calc", [0, 1, 2, 3])
def test_nspinor2_nc(lib, nvloc, type_calc):
    """nspinor = 2, norm-conserving: collinear potential on both spinor components (nvloc=1) and the non-collinear 2x2
    potential (nvloc=4: V11, V22, Re V12, Im V12), every type_calc, host and device blocks, fused and generic fourwf."""
    ndat = 3
    p = make_problem(7.0, (8.0, 9.0, 7.5), (.1, .2, .3), 1, ndat=2 * ndat, natom_per_type=(2,), lmax_per_type=(1,))
    n1, n2, n3 = p.ngfft
    cw = np.ascontiguousarray(p.cwavef.reshape(ndat, 2, p.npw))
    rng = np.random.default_rng(21)
    if nvloc == 1:
        vl = p.vlocal
    else:
        i3, i2, i1 = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
        vl = np.stack([p.vlocal, p.vlocal + 0.2 * np.cos(2 * np.pi * i1 / n1), 0.15 * np.sin(2 * np.pi * i2 / n2),
                       0.1 * np.cos(2 * np.pi * (i3 / n3 - i1 / n1))])
        vl = np.ascontiguousarray(vl)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    ghc0 = rng.standard_normal(cw.shape) + 1j * rng.standard_normal(cw.shape)
    ref, ref_gv = ogh.getghc_spinor(cw, vl, p.kg, p.ngfft, p.kinpw, P, p.enl, p.indlmn, p.nattyp, p.atindx1 - 1, type_calc=type_calc)
    if type_calc == 2:
        # type_calc = 2 ADDS the non-local + kinetic part to the caller's ghc (m_getghc.F90:152)
        ok = p.kinpw < g.KIN_FILTER
        ref = np.where(ok[None, None, :], ghc0 + np.where(ok, p.kinpw, 0.0)[None, None, :] * cw + ref_gv, 0.0)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 0, p.ucvol)
    h.set_nspinor(2)
    h.load_spin_nvloc(vl, nvloc)
    h.load_enl(p.enl, None)
    h.load_k(1, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    for impl in (0, 1):
        ab.api.L().abi_b200_fourwf_set_impl(impl)
        out = ghc0.copy(); gv = np.zeros_like(out)
        ab.getghc(-1, cw, None, out, None, h, gv, None, None, ndat, type_calc=type_calc)
        assert rel_err_per_band(out.reshape(2 * ndat, -1), ref.reshape(2 * ndat, -1)) < TOL, (impl,)
        if type_calc in (0, 2):
            assert rel_err_per_band(gv.reshape(2 * ndat, -1), ref_gv.reshape(2 * ndat, -1)) < TOL
    ab.api.L().abi_b200_fourwf_set_impl(0)
    h.destroy()
