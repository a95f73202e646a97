"""GPU parity against the committed golden fixtures (tests/golden/*.npz), through the C-ABI."""
import os
import sys
import numpy as np
import pytest
import abinit_b200 as ab
from problems import rel_err_per_band

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)


@pytest.mark.parametrize("name", ["nc_k_istwfk1", "nc_gamma_istwfk2", "paw_k_istwfk1", "paw_half_istwfk5"])
def test_getghc_matches_golden(lib, name):
    from make_golden import problem_of
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    p = problem_of(name)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, 1); h.load_enl(p.enl, p.sij); h.load_k(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d)
    ghc = np.zeros((p.ndat, p.npw), dtype=np.complex128); gsc = np.zeros_like(ghc); gv = np.zeros_like(ghc)
    ab.getghc(-1, p.cwavef, None, ghc, gsc if p.usepaw else None, h, gv, None, None, p.ndat, sij_opt=1 if p.usepaw else 0)
    assert rel_err_per_band(ghc, gold["ghc"]) < 1e-11
    assert rel_err_per_band(gv, gold["gvnlxc"]) < 1e-11
    if p.usepaw:
        assert rel_err_per_band(gsc, gold["gsc"]) < 1e-11
    out = np.zeros_like(ghc)
    n1, n2, n3 = p.ngfft
    ab.fourwf(1, p.vlocal, p.cwavef, out, None, None, None, p.istwf_k, p.kgF, p.kgF, max(p.ngfft), None, p.ndat, p.ngfft,
              p.npw, p.npw, n1, n2, n3, 2)
    assert rel_err_per_band(out, gold["fourwf_opt2"]) < 1e-11
    h.destroy()
