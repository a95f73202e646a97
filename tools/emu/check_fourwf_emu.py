"""Developer check: fused/generic fourwf index logic in the single-thread emulation vs the oracle."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from abinit_b200 import lib as blib, api
from oracle import gsphere as g, fourwf as ofw
emu = blib.load_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libabinit_b200_emu.so"))
api._use_library(emu)
rng = np.random.default_rng(1)
def run(ecut, L, kpt, istwfk, ndat, cplex, option, impl, ng=None):
    _, gmet, _ = g.metric(np.diag(L) if np.ndim(L) else np.eye(3) * L)
    ng = ng or g.getng(2.0, ecut, gmet, kpt)
    kg = g.kpgsph(ecut, gmet, kpt, istwfk)
    npw = kg.shape[1]
    n1, n2, n3 = ng
    cg = rng.standard_normal((ndat, npw)) + 1j * rng.standard_normal((ndat, npw))
    if istwfk == 2: cg[:, 0] = cg[:, 0].real
    V = rng.standard_normal((n3, n2, n1)) + (1j * rng.standard_normal((n3, n2, n1)) if cplex == 2 else 0)
    kgF = np.ascontiguousarray(kg.T)
    out = np.zeros((ndat, npw), dtype=complex)
    fofr = np.zeros((ndat, n3, n2, n1), dtype=complex)
    den = np.ascontiguousarray(V.real if cplex == 1 else V)
    if option == 1: den = np.abs(den) + 0.0
    if option == 3:
        fofr[:] = rng.standard_normal(fofr.shape) + 1j * rng.standard_normal(fofr.shape)
    ref_out, ref_r, ref_den = ofw.fourwf(cplex, den.copy(), cg, fofr.copy(), kg, kg, ng, option, istwfk, weight_r=0.7, weight_i=0.3)
    den2 = den.copy()
    api.fourwf(cplex, den2, cg, out, fofr, None, None, istwfk, kgF, kgF, max(ng), None, ndat, ng, npw, npw, n1, n2, n3, option,
               weight_r=0.7, weight_i=0.3, impl=impl)
    if option in (2, 3): err = np.abs(out - ref_out).max() / np.abs(ref_out).max()
    elif option == 0: err = np.abs(fofr - ref_r).max() / np.abs(ref_r).max()
    else: err = np.abs(den2 - ref_den).max() / np.abs(ref_den).max()
    print(f"ng={ng} npw={npw} istwfk={istwfk} ndat={ndat} cplex={cplex} option={option} impl={impl}: rel err {err:.2e}")
    assert err < 1e-12
if __name__ == "__main__":
    api.init(0)
    if os.environ.get("HALF_CFG"):
        api.set_tuning("half_cfg", int(os.environ["HALF_CFG"]))   # developer variants of the half-support plane stage
    for impl in (1, 2):
        run(6.0, 8.0, (.1, .2, .3), 1, 2, 1, 2, impl)
        run(6.0, 8.0, (0, 0, 0), 2, 3, 1, 2, impl)
        run(5.0, (7., 8., 9.), (.5, 0, 0), 3, 1, 1, 2, impl)
        run(5.0, (7., 8., 9.), (.1, 0, .3), 1, 2, 2, 2, impl)
        run(5.0, 8., (.5, .5, .5), 9, 2, 1, 2, impl)
        run(4.0, 9., (0, .5, .5), 8, 1, 1, 2, impl, ng=(28, 35, 21))
    # plane-stage cases (n2 == n3 in the two-pass list), incl. non-cubic n1, Gamma, k=1/2 cases, complex V
    run(5.0, 8., (.1, .2, .3), 1, 2, 1, 2, 2, ng=(24, 24, 24))
    run(5.0, 8., (0, 0, 0), 2, 3, 1, 2, 2, ng=(27, 30, 30))      # packed Gamma path, odd ndat
    run(5.0, 8., (0, 0, 0), 2, 4, 1, 2, 2, ng=(24, 24, 24))
    run(5.0, 8., (0, 0, 0), 2, 1, 1, 2, 2, ng=(24, 24, 24))
    run(5.0, 8., (.5, .5, .5), 9, 2, 1, 2, 2, ng=(32, 36, 36))
    run(5.0, 8., (.1, 0, .3), 1, 2, 2, 2, 2, ng=(25, 40, 40))
    run(7.0, 9., (0, .5, 0), 6, 1, 1, 2, 2, ng=(45, 45, 45))
    # split plane stage (n2 != n3, both in the two-pass table): generic k, packed Gamma, complex V, k=1/2 spheres, option 1
    run(5.0, 8., (.1, .2, .3), 1, 2, 1, 2, 2, ng=(24, 30, 36))
    run(5.0, 8., (0, 0, 0), 2, 3, 1, 2, 2, ng=(27, 36, 30))
    run(5.0, (7., 8., 9.), (.1, 0, .3), 1, 2, 2, 2, 2, ng=(25, 40, 48))
    run(5.0, 8., (.5, .5, .5), 9, 2, 1, 2, 2, ng=(32, 36, 40))
    # option 1 (the emulation runs blockDim = 1, so k_rho_weights only fills transform 0: one transform per call here)
    run(5.0, 8., (.1, .2, .3), 1, 1, 1, 1, 2, ng=(24, 30, 36))
    run(5.0, 8., (0, 0, 0), 2, 2, 1, 1, 2, ng=(27, 36, 30))
    run(5.0, 8., (.5, .5, .5), 9, 1, 1, 1, 2, ng=(32, 36, 40))
    run(3.0, 7., (.1, .2, .3), 1, 1, 1, 2, 2, ng=(20, 60, 60))
    for option in (0, 1, 3):
        run(6.0, 8.0, (.1, .2, .3), 1, 2, 1, option, 0)
        run(6.0, 8.0, (0, 0, 0), 2, 2, 1, option, 0)
    print("emu fourwf OK")
