"""GPU: the EXPERIMENTAL int8-sliced gemm_nonlop (csrc/ozaki.cu, opt-in) reproduces the FP64 DMMA path and the oracle to
the north-star tolerance (real istwf_k >= 2 and complex istwf_k = 1 contractions, NC getghc, PAW with S)."""
import numpy as np
import pytest
from oracle import getghc as ogh, nonlop as onl
from problems import make_problem, rel_err_per_band
import abinit_b200 as ab
from abinit_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ozaki_on(lib):
    api.set_tuning("nonlop_ozaki", 1)
    yield
    api.set_tuning("nonlop_ozaki", 0)


@pytest.mark.parametrize("istwf_k,kpt,usepaw,ndat", [(2, (0, 0, 0), 0, 8), (2, (0, 0, 0), 0, 5), (3, (.5, 0, 0), 0, 4), (2, (0, 0, 0), 1, 6),
                                                     (1, (.1, .2, .3), 0, 8), (1, (-.25, .5, 0), 0, 3), (1, (.1, .2, .3), 1, 5)])
def test_getghc_ozaki_vs_oracle(lib, ozaki_on, istwf_k, kpt, usepaw, ndat):
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=ndat, natom_per_type=(2, 1), lmax_per_type=(2, 1), usepaw=usepaw)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, p.sij); h.load_k(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    ghc = np.zeros((ndat, p.npw), dtype=np.complex128); gsc = np.zeros_like(ghc) if usepaw else None; gv = np.zeros_like(ghc)
    k0 = ab.kernel_launches()
    ab.getghc(-1, p.cwavef, None, ghc, gsc, h, gv, None, None, ndat, sij_opt=1 if usepaw else 0)
    r_ghc, r_gsc, r_gv, _ = ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                                       istwf_k=istwf_k, usepaw=usepaw, sij_opt=1 if usepaw else 0)
    assert rel_err_per_band(ghc, r_ghc) < 1e-11
    assert rel_err_per_band(gv, r_gv) < 1e-11
    if usepaw:
        assert rel_err_per_band(gsc, r_gsc) < 1e-11
    # the FP64 path on the same handle agrees as well (and is what runs by default)
    api.set_tuning("nonlop_ozaki", 0)
    ghc2 = np.zeros_like(ghc)
    ab.getghc(-1, p.cwavef, None, ghc2, None if not usepaw else np.zeros_like(ghc), h, None, None, None, ndat, sij_opt=1 if usepaw else 0)
    api.set_tuning("nonlop_ozaki", 1)
    assert rel_err_per_band(ghc, ghc2) < 1e-11
    h.destroy()


def test_igemm_tc_bit_exact_against_naive_kernel():
    """The hand-written tcgen05 int8 GEMM (csrc/igemm_tc.cuh) against a naive kernel: regular, ragged and split-K shapes,
    every tile variant (tools/igemm_lab.cu, built by __graft_entry__.build())."""
    import os, subprocess
    lab = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "igemm_lab")
    if not os.path.exists(lab):
        pytest.skip("tools/igemm_lab not built (python -c 'import __graft_entry__ as g; g.build()')")
    for variant in ("0", "1", "2", "3"):
        r = subprocess.run([lab, "small", variant], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        lines = [l for l in r.stdout.splitlines() if l.startswith("M=")]
        assert len(lines) == 8 and all(": ok" in l for l in lines), r.stdout
