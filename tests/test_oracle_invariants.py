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
