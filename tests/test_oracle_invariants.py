"""CPU: invariants of the oracle restatements that have no stored reference vectors (PAW inverse overlap, spinor getghc,
int8 slicing study): they are what the GPU parity tests lean on, so they are checked here without a GPU."""
import numpy as np
import pytest
from oracle import nonlop as onl, invovl as oiv, getghc as ogh
from problems import make_problem


@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0)), (3, (.5, 0, 0))])
def test_invovl_is_the_inverse_of_S(istwf_k, kpt):
    p = make_problem(6.0, 8.0, kpt, istwf_k, ndat=3, natom_per_type=(1, 2), lmax_per_type=(2, 1), usepaw=1)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    iv = oiv.make_invovl(P, p.sij, p.indlmn, p.nattyp, istwf_k)
    info = {}
    s1, _ = oiv.apply_invovl(P, iv, p.cwavef, istwf_k, info=info)
    _, back, _ = onl.gemm_nonlop(P, s1, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k, choice=1, paw_opt=3)
    assert np.max(np.abs(back - p.cwavef)) < 1e-13
    assert info["iters"] <= 30 and info["maxerr"] < 1e-12


def test_spinor_getghc_reduces_to_scalar_case_and_is_hermitian():
    ndat = 2
    p = make_problem(6.0, (7.0, 8.0, 7.5), (.1, .2, .3), 1, ndat=2 * ndat, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    cw = np.ascontiguousarray(p.cwavef.reshape(ndat, 2, p.npw))
    n1, n2, n3 = p.ngfft
    # nvloc = 4 with V11 = V22 and V12 = 0 is the collinear case, which is the scalar getghc on every spinor component
    v4 = np.stack([p.vlocal, p.vlocal, np.zeros_like(p.vlocal), np.zeros_like(p.vlocal)])
    a, _ = ogh.getghc_spinor(cw, v4, p.kg, p.ngfft, p.kinpw, P, p.enl, p.indlmn, p.nattyp, p.atindx1 - 1)
    b, _ = ogh.getghc_spinor(cw, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.indlmn, p.nattyp, p.atindx1 - 1)
    ref, _, _, _ = ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, None, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k=1)
    assert np.max(np.abs(a - b)) < 1e-13 and np.max(np.abs(b.reshape(2 * ndat, -1) - ref)) < 1e-13
    # a genuinely non-collinear potential: H stays Hermitian on spinors (V12 enters as V3 + i V4 / V3 - i V4)
    i3, i2, i1 = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
    v4 = np.stack([p.vlocal, p.vlocal + 0.2 * np.cos(2 * np.pi * i1 / n1), 0.15 * np.sin(2 * np.pi * i2 / n2),
                   0.1 * np.cos(2 * np.pi * (i3 / n3 - i1 / n1))])
    h, _ = ogh.getghc_spinor(cw, v4, p.kg, p.ngfft, p.kinpw, P, p.enl, p.indlmn, p.nattyp, p.atindx1 - 1)
    A = np.conj(cw.reshape(ndat, -1)) @ h.reshape(ndat, -1).T
    assert np.max(np.abs(A - A.conj().T)) < 1e-12 * np.max(np.abs(A))


def test_int8_slicing_is_exact_to_the_advertised_level():
    """The slicing scheme of csrc/ozaki.cu emulated in NumPy (tools/ozaki_study.py): 7 slices of 7 bits reproduce a GEMM with
    decaying columns to < 1e-12 relative, 6 slices do not reach 1e-11."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("ozaki_study", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "ozaki_study.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    rng = np.random.default_rng(0)
    K, M, N = 4000, 24, 12
    a = rng.standard_normal((K, M)) / np.sqrt(K)
    b = rng.standard_normal((K, N)) / (1.0 + np.repeat(np.sort(rng.uniform(0, 20, K // 2)), 2))[:, None]
    ref = a.T @ b
    err = {}
    for nsl in (6, 7):
        got, nprod = m.ozaki_gemm(a, b, nsl, 7)
        err[nsl] = np.max(np.linalg.norm(got - ref, axis=0) / np.linalg.norm(ref, axis=0))
        assert nprod == nsl * (nsl + 1) // 2
    assert err[7] < 1e-12 and err[6] > err[7]


@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0)), (3, (.5, 0, 0)), (6, (0, .5, 0)), (9, (.5, .5, .5))])
@pytest.mark.parametrize("cplex,ndat", [(1, 1), (1, 4), (1, 5), (2, 3)])
def test_padded_fourwf_equals_full_fft(istwf_k, kpt, cplex, ndat):
    """oracle/fourwf_pad.py (the reference's zero-padded passes, fftw3_fftpad.finc:14-196, + the Gamma-point band pairing of
    m_getghc.F90:1999-2171; used by bench.py's CPU arm) gives the result of the plain full-box oracle fourwf, option 2."""
    from oracle import gsphere as g, fourwf as ofw
    from oracle.fourwf_pad import fourwf_option2_padded
    rng = np.random.default_rng(5)
    _, gmet, _ = g.metric(np.diag([7., 8., 9.]))
    ng = g.getng(2.0, 6.0, gmet, kpt)
    kg = g.kpgsph(6.0, gmet, kpt, istwf_k)
    npw = kg.shape[1]
    c = rng.standard_normal((ndat, npw)) + 1j * rng.standard_normal((ndat, npw))
    if istwf_k == 2:
        c[:, 0] = c[:, 0].real
    n1, n2, n3 = ng
    v = rng.standard_normal((n3, n2, n1)) + (1j * rng.standard_normal((n3, n2, n1)) if cplex == 2 else 0)
    ref, _, _ = ofw.fourwf(cplex, v, c, None, kg, kg, ng, 2, istwf_k)
    out = fourwf_option2_padded(cplex, v, c, kg, ng, istwf_k, chunk=2)
    assert np.abs(out - ref).max() < 1e-13 * np.abs(ref).max()


def _cplx_paw_problem(nspinor, seed=21):
    from problems import make_problem
    p = make_problem(7.0, 8.0, (0.2, -0.1, 0.3), 1, ndat=3, seed=seed, natom_per_type=(2, 1), lmax_per_type=(1, 2), usepaw=1)
    rng = np.random.default_rng(seed)
    lmn2 = p.lmnmax * (p.lmnmax + 1) // 2
    nblk = 4 if nspinor == 2 else 1
    enl = 0.4 * rng.standard_normal((nblk, p.natom, 2 * lmn2))
    if nblk == 4:
        # the packed upper triangles of D^{ud} and D^{du} are independent data except on the diagonal, D^{du}_jj = conj(D^{ud}_jj)
        for j in range(p.lmnmax):
            pk = j * (j + 1) // 2 + j
            enl[3, :, 2 * pk] = enl[2, :, 2 * pk]; enl[3, :, 2 * pk + 1] = -enl[2, :, 2 * pk + 1]
    c = rng.standard_normal((p.ndat, nspinor, p.npw)) + 1j * rng.standard_normal((p.ndat, nspinor, p.npw))
    return p, enl, c


def test_complex_dij_reduces_to_the_pinned_real_path():
    """cplex_dij = 2 with zero imaginary parts (and spinors with empty off-diagonal blocks and equal diagonal blocks) must give the
    real packed-symmetric result of nonlop.opernlc, which the tw90_1 SCF pins (m_opernlc_ylm_allwf.F90:336-447 vs :453-737)."""
    from oracle import nonlop as onl
    p, enl, c = _cplx_paw_problem(2)
    lmn2 = p.lmnmax * (p.lmnmax + 1) // 2
    real_d = enl[0, :, 0::2].copy()
    e = np.zeros_like(enl); e[0, :, 0::2] = real_d; e[1, :, 0::2] = real_d
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    vo, so = onl.gemm_nonlop_general(P, c, e, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, 4, nspinor=2, cplex_enl=2)
    for isp in range(2):
        ro, rs, _ = onl.gemm_nonlop(P, c[:, isp], real_d, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, 1, choice=1, paw_opt=4)
        assert np.abs(vo[:, isp] - ro).max() < 1e-13 * np.abs(ro).max()
        assert np.abs(so[:, isp] - rs).max() < 1e-13 * np.abs(rs).max()


def test_complex_and_spinor_dij_operator_is_hermitian():
    """<a| Vnl b> = conj(<b| Vnl a>) for complex Hermitian D_ij (nspinor 1) and for the four spinor blocks (nspinor 2)."""
    from oracle import nonlop as onl
    for nspinor in (1, 2):
        p, enl, c = _cplx_paw_problem(nspinor, seed=30 + nspinor)
        P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
        vo, _ = onl.gemm_nonlop_general(P, c, enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, 1, nspinor=nspinor, cplex_enl=2)
        a = c.reshape(p.ndat, -1); hb = vo.reshape(p.ndat, -1)
        g = a.conj() @ hb.T
        assert np.abs(g - g.conj().T).max() < 1e-12 * np.abs(g).max()
