"""CPU suite: the oracle against every value the reference's own tests pin for this path."""
import os
import numpy as np
import pytest
from oracle import gsphere as g, fourwf as ofw, nonlop as onl, getghc as ogh
from problems import make_problem, rel_err_per_band

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_getng_and_kpgsph_tbase3():
    """tests/tutorial/Refs/tbase3_1.abo: ngfft 24 24 24 (:65), mpw 525 (:34), npw avg 520.500/520.494 (:181) for
    Si fcc a=10.18, ecut 12, k = (-1/4,1/2,0) w=3/4 and (-1/4,0,0) w=1/4."""
    a = 10.18
    rprimd = a * np.array([[0, .5, .5], [.5, 0, .5], [.5, .5, 0]]).T
    _, gmet, ucvol = g.metric(rprimd)
    assert g.getng(2.0, 12.0, gmet) == (24, 24, 24)
    n1 = g.kpgsph(12.0, gmet, (-.25, .5, 0)).shape[1]
    n2 = g.kpgsph(12.0, gmet, (-.25, 0, 0)).shape[1]
    assert (n1, n2) == (519, 525) and max(n1, n2) == 525
    assert abs(0.75 * n1 + 0.25 * n2 - 520.500) < 1e-9
    assert abs(np.exp(0.75 * np.log(n1) + 0.25 * np.log(n2)) - 520.494) < 5e-4


def test_getng_fftprof_box():
    """tests/unitary/Refs/tfourwf_01.stdout:36-40: ecut 30, 20 Bohr cube, k=(.1,.2,.3) -> FFT mesh 100 100 100."""
    _, gmet, _ = g.metric(np.eye(3) * 20.0)
    assert g.getng(2.0, 30.0, gmet, (.1, .2, .3)) == (100, 100, 100)


def test_si512_shape():
    """Si-512 (cubic 40.72, ecut 20, Gamma) -> 180^3, npw 288113 / 144057."""
    _, gmet, _ = g.metric(np.eye(3) * 40.72)
    assert g.getng(2.0, 20.0, gmet) == (180, 180, 180)
    assert g.kpgsph(20.0, gmet, (0, 0, 0), 1).shape[1] == 288113
    assert g.kpgsph(20.0, gmet, (0, 0, 0), 2).shape[1] == 144057


def test_kpgsph_ordering_and_half_sphere():
    _, gmet, _ = g.metric(np.eye(3) * 7.0)
    kg = g.kpgsph(6.0, gmet, (0, 0, 0), 1)
    assert tuple(kg[:, 0]) == (0, 0, 0)
    # ig1 innermost in the order 0..max,min..-1
    row0 = kg[:, (kg[1] == 0) & (kg[2] == 0)][0]
    m = row0.max()
    assert list(row0) == list(range(0, m + 1)) + list(range(-m, 0))
    kg2 = g.kpgsph(6.0, gmet, (0, 0, 0), 2)
    assert 2 * kg2.shape[1] - 1 == kg.shape[1]
    s = {tuple(x) for x in kg2.T.tolist()}
    assert all((tuple(-np.array(x)) not in s) or x == (0, 0, 0) for x in s)


@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0)), (3, (.5, 0, 0)), (6, (0, .5, 0)), (9, (.5, .5, .5))])
def test_fourwf_fftprof_closed_form(istwf_k, kpt):
    """src/70_gw/m_fft_prof.F90:873,936-960 vectors: c(G)=exp(-(2pi)^2 G.gmet.G), V=cos(2pi g0.r), g0=(1,-1,2)
    => out(G) = 1/2 [c(G-g0)+c(G+g0)]; tolerance = 10 x the reference's cross-library spread (tfourwf_01.stdout:129)."""
    _, gmet, _ = g.metric(np.eye(3) * 10.0)
    ng = g.getng(2.0, 8.0, gmet, kpt)
    kg = g.kpgsph(8.0, gmet, kpt, istwf_k)
    gsq = (2 * np.pi) ** 2 * np.einsum("ip,ij,jp->p", kg, gmet, kg.astype(float))
    c = np.exp(-gsq)[None, :].astype(complex)
    g0 = np.array([1, -1, 2]); n1, n2, n3 = ng
    i3, i2, i1 = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
    V = np.cos(2 * np.pi * (g0[0] * i1 / n1 + g0[1] * i2 / n2 + g0[2] * i3 / n3))
    out, _, _ = ofw.fourwf(1, V, c, None, kg, kg, ng, 2, istwf_k)
    full = ofw.sphere_to_box(c, kg, ng, istwf_k)[0]
    exp = 0.5 * (np.roll(full, (g0[2], g0[1], g0[0]), (0, 1, 2)) + np.roll(full, (-g0[2], -g0[1], -g0[0]), (0, 1, 2)))
    w1, w2, w3 = ofw._wrap(kg, ng)
    ref = exp[w3, w2, w1]
    if istwf_k == 2:
        ref[0] = ref[0].real
    assert np.abs(out[0] - ref).max() < 3.4e-15


def test_fourwf_roundtrip_and_density():
    p = make_problem(6.0, 8.0, (.1, .2, .3), 1, ndat=2)
    _, ur, _ = ofw.fourwf(1, None, p.cwavef, None, p.kg, p.kg, p.ngfft, 0)
    back, _, _ = ofw.fourwf(1, None, None, ur, p.kg, p.kg, p.ngfft, 3)
    assert rel_err_per_band(back, p.cwavef) < 1e-13
    # option 1 with unit weights integrates to N * sum |c|^2 (Parseval with the un-normalised G->r transform)
    rho0 = np.zeros(p.ngfft[::-1])
    _, _, rho = ofw.fourwf(1, rho0, p.cwavef, None, p.kg, p.kg, p.ngfft, 1, weight_r=1.0, weight_i=1.0)
    assert abs(rho.sum() / rho.size - np.sum(np.abs(p.cwavef) ** 2)) < 1e-12


def test_mkkin_sentinel():
    _, gmet, _ = g.metric(np.eye(3) * 8.0)
    kg = g.kpgsph(6.0, gmet, (0, 0, 0), 1)
    kin = g.mkkin(5.0, 0.0, 1.0, gmet, kg, (0, 0, 0))     # smaller ecut -> outer shell filtered
    assert np.all(kin[kin > 1e290] == g.KIN_SENTINEL) and np.any(kin > 1e290)
    assert g.KIN_SENTINEL >= g.KIN_FILTER


@pytest.mark.parametrize("usepaw,paw_opt", [(0, 0), (1, 4)])
def test_gemm_nonlop_invariants(usepaw, paw_opt):
    """gemm_nonlop is unpinned by stored vectors -> invariants: per-atom naive statement, Hermiticity of V_nl and S."""
    p = make_problem(6.0, 8.0, (.25, 0, .1), 1, ndat=4, natom_per_type=(2, 2), lmax_per_type=(1, 2), usepaw=usepaw)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    v, s, _ = onl.gemm_nonlop(P, p.cwavef, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, 1, 1, paw_opt)
    nv, ns = onl.nonlop_naive(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.atindx1 - 1, p.ucvol, p.cwavef, p.enl, p.sij, paw_opt)
    assert rel_err_per_band(v, nv) < 1e-13
    A = np.conj(p.cwavef) @ v.T
    assert np.abs(A - A.conj().T).max() < 1e-13 * max(1.0, np.abs(A).max())
    if s is not None:
        assert rel_err_per_band(s, ns) < 1e-13


def test_getghc_hermitian_and_gamma_consistency():
    p = make_problem(6.0, 8.0, (.1, .2, .3), 1, ndat=4, filter_shell=False)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    ghc, _, _, _ = ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1)
    A = np.conj(p.cwavef) @ ghc.T
    assert np.abs(A - A.conj().T).max() < 1e-13 * np.abs(A).max()
    # Gamma: istwf_k=2 == istwf_k=1 on the completed sphere
    p2 = make_problem(6.0, 8.0, (0, 0, 0), 2, ndat=2, filter_shell=False)
    p1 = make_problem(6.0, 8.0, (0, 0, 0), 1, ndat=2, filter_shell=False)
    lut = {tuple(k): i for i, k in enumerate(p2.kgF.tolist())}
    c1 = np.array([[p2.cwavef[b, lut[tuple(k)]] if tuple(k) in lut else np.conj(p2.cwavef[b, lut[tuple(-x for x in k)]])
                    for k in p1.kgF.tolist()] for b in range(2)])
    res = []
    for q, c in ((p1, c1), (p2, p2.cwavef)):
        Pq = onl.prep_projectors(q.ffnl, q.ph3d, q.indlmn, q.nattyp, q.ucvol)
        res.append(ogh.getghc(c, q.vlocal, q.kg, q.ngfft, q.kinpw, Pq, q.enl, q.sij, q.indlmn, q.nattyp, q.atindx1 - 1,
                              istwf_k=q.istwf_k)[0])
    sel = np.array([i for i, k in enumerate(p1.kgF.tolist()) if tuple(k) in lut])
    tgt = np.array([lut[tuple(p1.kgF[i].tolist())] for i in sel])
    assert rel_err_per_band(res[0][:, sel], res[1][:, tgt]) < 1e-13


@pytest.mark.parametrize("name", ["nc_k_istwfk1", "nc_gamma_istwfk2", "paw_k_istwfk1", "paw_half_istwfk5"])
def test_oracle_matches_golden(name):
    """Committed fixtures (tests/golden/make_golden.py) guard the oracle against silent drift."""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import problem_of
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    p = problem_of(name)
    assert np.array_equal(p.kg, gold["kg"]) and tuple(gold["ngfft"]) == tuple(p.ngfft)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    ghc, gsc, gv, prj = ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp,
                                   p.atindx1 - 1, istwf_k=p.istwf_k, usepaw=p.usepaw, sij_opt=1 if p.usepaw else 0)
    assert rel_err_per_band(ghc, gold["ghc"]) < 1e-13
    assert rel_err_per_band(gv, gold["gvnlxc"]) < 1e-13
    if p.usepaw:
        assert rel_err_per_band(gsc, gold["gsc"]) < 1e-13
