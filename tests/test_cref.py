"""CPU: the C++/OpenMP restatement of the reference's CPU getghc (oracle/cref, the timed arm of bench.py) pinned on the NumPy
oracle, which is itself pinned on the reference's stored SCF results (oracle/__init__.py).  Tolerance 1e-13 relative per band."""
import numpy as np
import pytest
from problems import make_problem, rel_err_per_band
from oracle import getghc as ogh, nonlop as onl, fourwf as ofw, cref


@pytest.mark.parametrize("n", [180, 100, 96, 84, 45, 30, 28, 24, 16])
def test_fft_engine_matches_numpy(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((37, n)) + 1j * rng.standard_normal((37, n))
    assert np.abs(cref.fft1d(x, -1) - np.fft.fft(x, axis=1)).max() < 1e-13 * n
    assert np.abs(cref.fft1d(x, +1) - np.fft.ifft(x, axis=1) * n).max() < 1e-13 * n


@pytest.mark.parametrize("istwf_k,kpt,ndat", [(1, (0.1, 0.2, 0.3), 3), (2, (0.0, 0.0, 0.0), 5), (2, (0.0, 0.0, 0.0), 4)])
def test_fourwf_option2_matches_oracle(istwf_k, kpt, ndat):
    p = make_problem(9.0, (7.0, 8.0, 9.5), kpt, istwf_k, ndat=ndat, seed=5)
    ref, _, _ = ofw.fourwf(1, p.vlocal, p.cwavef, None, p.kg, p.kg, p.ngfft, 2, istwf_k)
    out = cref.fourwf_option2(p.vlocal, p.cwavef, p.kgF, p.ngfft, istwf_k)
    assert rel_err_per_band(out, ref) < 1e-13


@pytest.mark.parametrize("istwf_k,kpt,usepaw,ndat", [(1, (0.25, -0.125, 0.5), 0, 4), (2, (0.0, 0.0, 0.0), 0, 5),
                                                     (1, (0.25, -0.125, 0.5), 1, 3), (2, (0.0, 0.0, 0.0), 1, 4)])
def test_getghc_matches_oracle(istwf_k, kpt, usepaw, ndat):
    p = make_problem(8.0, 8.5, kpt, istwf_k, ndat=ndat, seed=11, natom_per_type=(2, 1), lmax_per_type=(2, 1), usepaw=usepaw)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    sij_opt = 1 if usepaw else 0
    ref = ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                     istwf_k=istwf_k, usepaw=usepaw, sij_opt=sij_opt)
    op = cref.Operator.from_oracle_arrays(p.vlocal, p.kgF, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                                          istwf_k, usepaw)
    for threads in (1, 3):
        cref.set_threads(threads)
        ghc, gsc = op.getghc(p.cwavef, sij_opt=sij_opt)
        assert rel_err_per_band(ghc, ref[0]) < 1e-13
        if usepaw:
            assert rel_err_per_band(gsc, ref[1]) < 1e-13
        # the sentinel shell is exactly zero
        assert np.all(ghc[:, p.kinpw > 1e290] == 0)
