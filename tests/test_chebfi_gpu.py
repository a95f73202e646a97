"""GPU parity: ChebFi2 (chebfiwf2) through the C-ABI vs the oracle restatement of m_chebfi2.F90, and a full SCF of the
reference's tbase3_1 system solved with the CUDA ChebFi2 instead of a dense diagonalisation."""
import os
import numpy as np
import pytest
from oracle import scf, xg as oxg, chebfi as och, getghc as ogh, nonlop as onl
from problems import make_problem
import abinit_b200 as ab
from abinit_b200 import xg

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(__file__), "golden", "si2_tbase3.npz")
R = scf.REF_TBASE3_1


def _ham(p):
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, p.sij)
    h.load_k(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    return h


@pytest.mark.parametrize("istwf_k,kpt", [(1, (-.25, .5, 0)), (2, (0, 0, 0))])
@pytest.mark.parametrize("oracle_opt", [0, 1])
def test_chebfiwf2_vs_oracle(lib, istwf_k, kpt, oracle_opt):
    """One and several ChebFi2 calls from the same start block: eigenvalues, residuals, enl_out and the spanned space."""
    nband = 10
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False)
    h = _ham(p)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    space, me_g0 = (xg.SPACE_C, -1) if istwf_k == 1 else (xg.SPACE_CR, 1)

    def apply_h(c):
        out, _, _, _ = ogh.getghc(c, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, None, p.indlmn, p.nattyp, p.atindx1 - 1,
                                  istwf_k=istwf_k)
        return out, c.copy()
    occ = np.where(np.arange(nband) < 6, 1.0, 0.0)
    kw = dict(nline=5, tolerance=1e-16, occ=occ, nbdbuf=2 if oracle_opt else 0, oracle=oracle_opt)
    x_ref = p.cwavef.copy(); cg = p.cwavef.copy()
    eig = np.zeros(nband); resid = np.zeros(nband); enl = np.zeros(nband)
    for it in range(3):
        info = {}
        w_ref, r_ref, x_ref = och.chebfi_run(apply_h, x_ref, space, me_g0, p.ecut, info=info, **kw)
        xg.chebfiwf2(cg, eig, occ, enl, h, nband, p.npw, 1, resid, 1e-16, p.ecut, 5, nbdbuf=kw["nbdbuf"],
                     chebfi_oracle=oracle_opt, bandpp=4)
        assert np.max(np.abs(eig - w_ref)) < 1e-9 * max(1.0, np.max(np.abs(w_ref))), (it, eig - w_ref)
        assert np.max(np.abs(resid - r_ref) / (np.abs(r_ref) + 1e-12)) < 1e-5, (it, resid, r_ref)
        # same vectors up to a sign / phase (well separated eigenvalues of the synthetic operator)
        ov = np.abs(np.diag(oxg.gram(space, x_ref, cg, me_g0)))
        gaps = np.min(np.abs(np.subtract.outer(w_ref, w_ref)) + np.eye(nband), axis=1)
        ok = gaps > 1e-4
        assert np.max(np.abs(ov[ok] - 1.0)) < 1e-7, (it, ov)
        gv, _, _ = onl.gemm_nonlop(P, cg, p.enl, None, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k, choice=1, paw_opt=0,
                                   cpopt=-1, me_g0=1)
        assert np.max(np.abs(enl - np.real(oxg.colwise_dot(space, cg, gv, me_g0)))) < 1e-10
    h.destroy()


def test_chebfiwf2_device_resident_and_bandpp_independent(lib):
    torch = pytest.importorskip("torch")
    nband = 12
    p = make_problem(7.0, 8.0, (0, 0, 0), 2, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False)
    h = _ham(p)
    res = []
    for bandpp, on_dev in ((nband, False), (5, True), (2, True)):
        cg = p.cwavef.copy()
        eig = np.zeros(nband); resid = np.zeros(nband)
        if on_dev:
            d = torch.from_numpy(cg).cuda()
            xg.chebfiwf2(d, eig, None, None, h, nband, p.npw, 1, resid, 1e-16, p.ecut, 4, bandpp=bandpp)
            cg = d.cpu().numpy()
        else:
            xg.chebfiwf2(cg, eig, None, None, h, nband, p.npw, 1, resid, 1e-16, p.ecut, 4, bandpp=bandpp)
        res.append((eig, resid, cg))
    for e, r, c in res[1:]:
        assert np.max(np.abs(e - res[0][0])) < 1e-12
        assert np.max(np.abs(np.abs(c) - np.abs(res[0][2]))) < 1e-9
    h.destroy()


def test_scf_with_cuda_chebfi2_reaches_reference_etotal(lib):
    """tbase3_1 solved with ChebFi2 on the GPU (8 bands, nline 6, 2 calls per SCF step): total energy within 1e-8 Ha of the
    dense-diagonalisation oracle SCF and of the reference's stored etotal; eigenvalues within 1e-8 Ha (north_star)."""
    s = scf.setup_from_fixture(np.load(FIX))
    nband = 8
    hams = []
    for ik in range(len(s.kpts)):
        h = ab.Hamiltonian(s.ngfft, s.xred.shape[1], 1, s.indlmn.shape[1], s.indlmn, s.nattyp, s.atindx1 + 1, 0, s.ucvol)
        h.load_enl(s.ekb, None)
        hams.append(h)
    rng = np.random.default_rng(5)
    cgs = []
    for ik in range(len(s.kpts)):
        npw = s.kg[ik].shape[1]
        c = (rng.standard_normal((nband, npw)) + 1j * rng.standard_normal((nband, npw))) / (1.0 + s.kinpw[ik])[None, :]
        cgs.append(np.ascontiguousarray(c))

    def solver(ik, vloc):
        h = hams[ik]
        h.load_spin(np.ascontiguousarray(vloc, dtype=np.float64), 1)
        h.load_k(1, np.ascontiguousarray(s.kg[ik].T), s.kinpw[ik], s.ffnl[ik], s.ph3d[ik])
        npw = s.kg[ik].shape[1]
        eig = np.zeros(nband); resid = np.zeros(nband); enl = np.zeros(nband)
        for _ in range(2):
            xg.chebfiwf2(cgs[ik], eig, None, enl, h, nband, npw, 1, resid, 1e-22, s.ecut, 6)
        return eig, cgs[ik], enl
    l0 = ab.kernel_launches()
    res = scf.total_energy_scf(s, None, eigensolver=solver, nband=5, maxit=80)
    assert ab.kernel_launches() > l0
    ref = scf.total_energy_scf(s, scf.apply_h_oracle(s))
    assert abs(res["energies"]["total"] - ref["energies"]["total"]) < 1e-8
    assert abs(res["energies"]["total"] - R["total"]) < 1e-8
    for a, b in zip(res["eig"], ref["eig"]):
        assert np.max(np.abs(a[:4] - b[:4])) < 1e-8
    assert abs(res["energies"]["non_local_psp"] - ref["energies"]["non_local_psp"]) < 1e-7
    for h in hams:
        h.destroy()


@pytest.mark.parametrize("istwf_k,kpt,usepaw", [(1, (-.25, .5, 0), 0), (2, (0, 0, 0), 0), (2, (0, 0, 0), 1)])
def test_chebfiwf2_paral_on_one_rank_equals_chebfiwf2(lib, istwf_k, kpt, usepaw):
    """The library's band-parallel driver (abi_b200_chebfiwf2_paral_, NCCL inside the library) on a one-rank communicator: no
    collective is issued and the result is the serial chebfiwf2's (the 2-GPU comparison is tests/test_chebfi_mgpu.py); the
    transposer on one rank is the identity."""
    torch = pytest.importorskip("torch")
    nband = 9
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False, usepaw=usepaw)
    h = _ham(p)
    cg1 = p.cwavef.copy(); eig1 = np.zeros(nband); res1 = np.zeros(nband)
    xg.chebfiwf2(cg1, eig1, None, None, h, nband, p.npw, 1, res1, 1e-16, p.ecut, 5, bandpp=4)
    xg.comm_init_rank(bytes(128), 1, 0)
    for on_dev in (False, True):
        cg = p.cwavef.copy(); eig = np.zeros(nband); res = np.zeros(nband)
        if on_dev:
            d = torch.from_numpy(cg).cuda()
            xg.chebfiwf2_paral(d, eig, None, None, res, h, nband, nband, p.npw, 1, 1e-16, p.ecut, 5, bandpp=4)
            cg = d.cpu().numpy()
        else:
            xg.chebfiwf2_paral(cg, eig, None, None, res, h, nband, nband, p.npw, 1, 1e-16, p.ecut, 5, bandpp=4)
        assert np.max(np.abs(eig - eig1)) < 1e-10
        assert np.max(np.abs(res - res1) / (np.abs(res1) + 1e-12)) < 1e-5
        assert np.max(np.abs(np.abs(cg) - np.abs(cg1))) < 1e-8
    # oracle-driven degree (chebfi_oracle = 1, band buffer) and enl_out (NC) through the band-parallel entry == the serial entry
    occ = np.where(np.arange(nband) < 6, 1.0, 0.0)
    cga = p.cwavef.copy(); eiga = np.zeros(nband); resa = np.zeros(nband); enla = np.zeros(nband)
    xg.chebfiwf2(cga, eiga, occ, None if usepaw else enla, h, nband, p.npw, 1, resa, 1e-16, p.ecut, 5, nbdbuf=2, chebfi_oracle=1, bandpp=4)
    cgb = p.cwavef.copy(); eigb = np.zeros(nband); resb = np.zeros(nband); enlb = np.zeros(nband)
    xg.chebfiwf2_paral(cgb, eigb, occ, None if usepaw else enlb, resb, h, nband, nband, p.npw, 1, 1e-16, p.ecut, 5, nbdbuf=2, chebfi_oracle=1,
                       bandpp=4)
    assert np.max(np.abs(eigb - eiga)) < 1e-10 and np.max(np.abs(np.abs(cgb) - np.abs(cga))) < 1e-8
    assert usepaw or np.max(np.abs(enlb - enla)) < 1e-10
    a = torch.randn((nband, p.npw, 2), dtype=torch.float64, device="cuda"); b = torch.zeros_like(a); c = torch.zeros_like(a)
    xg.xg_transpose(True, a, b, p.npw, nband)
    xg.xg_transpose(False, c, b, p.npw, nband)
    ab.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)
    xg.comm_destroy()
    h.destroy()
