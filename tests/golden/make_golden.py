"""Generates tests/golden/*.npz from the oracle (run in the build container: python tests/golden/make_golden.py).
The reference is Fortran and cannot be executed here, so these vectors freeze the *pinned oracle's* outputs on small
seeded cases; both the oracle (CPU suite) and the CUDA path (GPU suite) are compared against them."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", ".."))
from problems import make_problem
from oracle import getghc as ogh, nonlop as onl, fourwf as ofw

CASES = {
    "nc_k_istwfk1": dict(ecut=6.0, L=(8.0, 9.0, 7.5), kpt=(-0.25, 0.5, 0.0), istwf_k=1, ndat=3, usepaw=0),
    "nc_gamma_istwfk2": dict(ecut=6.0, L=8.0, kpt=(0.0, 0.0, 0.0), istwf_k=2, ndat=3, usepaw=0),
    "paw_k_istwfk1": dict(ecut=6.0, L=8.5, kpt=(0.1, 0.2, 0.3), istwf_k=1, ndat=2, usepaw=1),
    "paw_half_istwfk5": dict(ecut=6.0, L=8.5, kpt=(0.5, 0.0, 0.5), istwf_k=5, ndat=2, usepaw=1),
}


def problem_of(name):
    c = CASES[name]
    return make_problem(c["ecut"], c["L"], c["kpt"], c["istwf_k"], ndat=c["ndat"], usepaw=c["usepaw"],
                        natom_per_type=(2, 1), lmax_per_type=(1, 2), seed=4242)


def main():
    for name in CASES:
        p = problem_of(name)
        P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
        sij_opt = 1 if p.usepaw else 0
        ghc, gsc, gv, prj = ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp,
                                       p.atindx1 - 1, istwf_k=p.istwf_k, usepaw=p.usepaw, sij_opt=sij_opt)
        loc, _, _ = ofw.fourwf(1, p.vlocal, p.cwavef, None, p.kg, p.kg, p.ngfft, 2, p.istwf_k)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kg=p.kg, ngfft=np.array(p.ngfft), cwavef=p.cwavef,
                            ghc=ghc, gsc=gsc if gsc is not None else np.zeros(0), gvnlxc=gv, proj=prj, fourwf_opt2=loc)
        print(name, p.ngfft, p.npw, float(np.abs(ghc).max()))


if __name__ == "__main__":
    main()
