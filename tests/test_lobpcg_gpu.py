"""GPU parity: LOBPCG (lobpcgwf2) and its building blocks (xg_Borthonormalize, XW / XWP Rayleigh-Ritz) through the C-ABI vs the
oracle restatement of m_lobpcg2.F90 / m_xg_ortho_RR.F90."""
import numpy as np
import pytest
from oracle import xg as oxg, lobpcg as olb, getghc as ogh, nonlop as onl
from problems import make_problem
import abinit_b200 as ab
from abinit_b200 import xg

pytestmark = pytest.mark.gpu


def _ham(p):
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, p.sij)
    h.load_k(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    return h


@pytest.mark.parametrize("istwf_k,kpt,usepaw", [(1, (-.25, .5, 0), 0), (2, (0, 0, 0), 0), (1, (.1, .2, .3), 1), (2, (0, 0, 0), 1)])
@pytest.mark.parametrize("nband", [6, 9])
def test_lobpcgwf2_vs_oracle(lib, istwf_k, kpt, usepaw, nband):
    """Several LOBPCG calls from the same start block (odd and even block sizes): eigenvalues, residuals, spanned vectors."""
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), usepaw=usepaw,
                     filter_shell=False)
    h = _ham(p)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    space, me_g0 = (xg.SPACE_C, -1) if istwf_k == 1 else (xg.SPACE_CR, 1)

    def apply_h(c):
        ghc, gsc, _, _ = ogh.getghc(c, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                                    istwf_k=istwf_k, usepaw=usepaw, sij_opt=1 if usepaw else 0)
        return ghc, (gsc if usepaw else c.copy())
    pcon = olb.build_pcon(p.kinpw)
    x_ref = p.cwavef.copy(); cg = p.cwavef.copy()
    eig = np.zeros(nband); resid = np.zeros(nband); enl = np.zeros(nband)
    for it in range(3):
        w_ref, r_ref, x_ref = olb.lobpcg_run(apply_h, x_ref, pcon, space, me_g0, nline=4, tolerance=1e-30)
        xg.lobpcgwf2(cg, eig, None, None if usepaw else enl, h, nband, p.npw, 1, resid, 1e-30, 4, bandpp=4)
        assert np.max(np.abs(eig - w_ref)) < 1e-9 * max(1.0, np.max(np.abs(w_ref))), (it, eig - w_ref)
        assert np.max(np.abs(resid - r_ref) / (np.abs(r_ref) + 1e-13)) < 1e-4, (it, resid, r_ref)
        _, bx = apply_h(cg)
        ov = np.abs(np.diag(oxg.gram(space, x_ref, bx, me_g0)))
        gaps = np.min(np.abs(np.subtract.outer(w_ref, w_ref)) + np.eye(nband), axis=1)
        assert np.max(np.abs(ov[gaps > 1e-4] - 1.0)) < 1e-6, (it, ov)
    # converging: residuals fall monotonically over the calls and the block is B-orthonormal
    _, bx = apply_h(cg)
    assert np.max(np.abs(oxg.gram(space, cg, bx, me_g0) - np.eye(nband))) < 1e-9
    h.destroy()


@pytest.mark.parametrize("istwf_k,kpt,usepaw", [(1, (-.25, .5, 0), 0), (2, (0, 0, 0), 0), (1, (.1, .2, .3), 1), (2, (0, 0, 0), 1)])
@pytest.mark.parametrize("nband,nblock", [(6, 2), (6, 3), (8, 2)])
def test_lobpcgwf2_multiblock_vs_oracle(lib, istwf_k, kpt, usepaw, nband, nblock):
    """nblock_lobpcg > 1 (m_lobpcg2.F90:456-695: lobpcg_orthoXwrtBlocks of X and W against the finished blocks, AX / BX transfer,
    final Borthonormalize + Rayleigh-Ritz over all bands :744-751) against the oracle restatement, incl. nbdbuf windows."""
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), usepaw=usepaw,
                     filter_shell=False)
    h = _ham(p)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    space, me_g0 = (xg.SPACE_C, -1) if istwf_k == 1 else (xg.SPACE_CR, 1)

    def apply_h(c):
        ghc, gsc, _, _ = ogh.getghc(c, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1,
                                    istwf_k=istwf_k, usepaw=usepaw, sij_opt=1 if usepaw else 0)
        return ghc, (gsc if usepaw else c.copy())
    pcon = olb.build_pcon(p.kinpw)
    x_ref = p.cwavef.copy(); cg = p.cwavef.copy()
    eig = np.zeros(nband); resid = np.zeros(nband); enl = np.zeros(nband)
    for it, nbdbuf in enumerate((0, 2, 0)):
        w_ref, r_ref, x_ref = olb.lobpcg_run(apply_h, x_ref, pcon, space, me_g0, nline=4, tolerance=1e-30, nblock=nblock, nbdbuf=nbdbuf)
        xg.lobpcgwf2(cg, eig, None, None if usepaw else enl, h, nband, p.npw, 1, resid, 1e-30, 4, nblock_lobpcg=nblock, nbdbuf=nbdbuf)
        assert np.max(np.abs(eig - w_ref)) < 1e-9 * max(1.0, np.max(np.abs(w_ref))), (it, eig - w_ref)
        assert np.max(np.abs(resid - r_ref) / (np.abs(r_ref) + 1e-13)) < 1e-4, (it, resid, r_ref)
        _, bx = apply_h(cg)
        ov = np.abs(np.diag(oxg.gram(space, x_ref, bx, me_g0)))
        gaps = np.min(np.abs(np.subtract.outer(w_ref, w_ref)) + np.eye(nband), axis=1)
        assert np.max(np.abs(ov[gaps > 1e-4] - 1.0)) < 1e-6, (it, ov)
    _, bx = apply_h(cg)
    assert np.max(np.abs(oxg.gram(space, cg, bx, me_g0) - np.eye(nband))) < 1e-9
    h.destroy()


def test_lobpcg_multiblock_converges_to_dense_eigenvalues(lib):
    """Band-by-band-like blocks (blockdim 2) reach the dense eigenvalues of the oracle's H(G,G') as the one-block run does."""
    nband = 8
    p = make_problem(5.0, 7.0, (.1, .2, .3), 1, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False)
    h = _ham(p)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    eye = np.eye(p.npw, dtype=np.complex128)
    hm, _, _, _ = ogh.getghc(eye, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, None, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k=1)
    hm = hm.T
    wd = np.linalg.eigvalsh(0.5 * (hm + hm.conj().T))[:nband]
    cg = p.cwavef.copy(); eig = np.zeros(nband); resid = np.zeros(nband)
    for _ in range(12):
        xg.lobpcgwf2(cg, eig, None, None, h, nband, p.npw, 1, resid, 1e-24, 5, nblock_lobpcg=4)
    assert np.max(np.abs(eig[:6] - wd[:6])) < 1e-9
    h.destroy()


def test_lobpcg_converges_to_dense_eigenvalues(lib):
    """LOBPCG on the GPU reaches the dense eigenvalues of the oracle's H(G,G') (NC, istwf_k=1)."""
    nband = 8
    p = make_problem(5.0, 7.0, (.1, .2, .3), 1, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False)
    h = _ham(p)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    eye = np.eye(p.npw, dtype=np.complex128)
    hm, _, _, _ = ogh.getghc(eye, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, None, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k=1)
    hm = hm.T
    wd = np.linalg.eigvalsh(0.5 * (hm + hm.conj().T))[:nband]
    cg = p.cwavef.copy(); eig = np.zeros(nband); resid = np.zeros(nband)
    for _ in range(8):
        xg.lobpcgwf2(cg, eig, None, None, h, nband, p.npw, 1, resid, 1e-24, 5)
    assert np.max(np.abs(eig[:6] - wd[:6])) < 1e-9
    assert np.max(resid[:6]) < 1e-16
    h.destroy()


@pytest.mark.parametrize("istwf_k,kpt,usepaw", [(1, (-.25, .5, 0), 0), (2, (0, 0, 0), 0), (2, (0, 0, 0), 1)])
def test_lobpcgwf2_paral_on_one_rank_equals_lobpcgwf2(lib, istwf_k, kpt, usepaw):
    """The library's band-parallel LOBPCG driver (abi_b200_lobpcgwf2_paral_) on a one-rank communicator against the serial
    lobpcgwf2 (one block); the 2-GPU comparison is tests/test_chebfi_mgpu.py."""
    nband = 7
    p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), usepaw=usepaw,
                     filter_shell=False)
    h = _ham(p)
    cg1 = p.cwavef.copy(); eig1 = np.zeros(nband); res1 = np.zeros(nband)
    xg.lobpcgwf2(cg1, eig1, None, None, h, nband, p.npw, 1, res1, 1e-30, 3, bandpp=4)
    xg.comm_init_rank(bytes(128), 1, 0)
    cg = p.cwavef.copy(); eig = np.zeros(nband); res = np.zeros(nband)
    xg.lobpcgwf2_paral(cg, eig, res, h, nband, nband, p.npw, 1, 1e-30, 3, bandpp=4)
    xg.comm_destroy()
    assert np.max(np.abs(eig - eig1)) < 1e-9
    assert np.max(np.abs(res - res1) / (np.abs(res1) + 1e-12)) < 1e-4
    assert np.max(np.abs(np.abs(cg) - np.abs(cg1))) < 1e-7
    h.destroy()
