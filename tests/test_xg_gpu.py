"""GPU parity: xgBlock algebra, Rayleigh-Ritz and the nonlop signs=1 energy through the C-ABI vs the oracle."""
import numpy as np
import pytest
from oracle import xg as oxg, nonlop as onl
from problems import make_problem
import abinit_b200 as ab
from abinit_b200 import xg

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _blocks(rng, ncols, rows, g0real):
    a = rng.standard_normal((ncols, rows)) + 1j * rng.standard_normal((ncols, rows))
    if g0real:
        a[:, 0] = a[:, 0].real
    return a


@pytest.mark.parametrize("space,me_g0", [(xg.SPACE_CR, 1), (xg.SPACE_CR, 0), (xg.SPACE_C, -1)])
@pytest.mark.parametrize("rows,na,nb", [(519, 5, 5), (4001, 37, 130), (20000, 131, 64), (7, 3, 2)])
def test_gram(lib, space, me_g0, rows, na, nb):
    rng = np.random.default_rng(rows + na)
    a = _blocks(rng, na, rows, me_g0 == 1); b = _blocks(rng, nb, rows, me_g0 == 1)
    ref = oxg.gram(space, a, b, me_g0)                                   # (na, nb)
    da, db = _dev(a), _dev(b)
    ldw = na + 3
    w = torch.zeros((nb, ldw), dtype=torch.complex128 if space == xg.SPACE_C else torch.float64, device="cuda")
    xg.xg_gram(space, rows, na, nb, da, rows, db, rows, w, ldw, me_g0)
    got = w.cpu().numpy()[:, :na].T
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref)) * np.sqrt(rows)
    assert np.all(w.cpu().numpy()[:, na:] == 0)                          # padding untouched


@pytest.mark.parametrize("space", [xg.SPACE_CR, xg.SPACE_C])
@pytest.mark.parametrize("rows,k,nout", [(519, 5, 5), (30011, 37, 37), (9000, 130, 70), (300, 131, 131)])
def test_rotate(lib, space, rows, k, nout):
    rng = np.random.default_rng(rows + k)
    x = _blocks(rng, k, rows, False)
    ldc = (k + 1) & ~1
    if space == xg.SPACE_C:
        c = rng.standard_normal((k, nout)) + 1j * rng.standard_normal((k, nout))
    else:
        c = rng.standard_normal((k, nout))
    cm = np.zeros((nout, ldc), dtype=c.dtype); cm[:, :k] = c.T           # column-major, K-padded
    dx, dc = _dev(x), _dev(cm)
    xg.xg_rotate(space, rows, k, nout, dx, rows, dc, ldc)
    ref = c.T @ x
    got = dx.cpu().numpy()[:nout]
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref)) * np.sqrt(k)
    if nout < k:
        assert np.array_equal(dx.cpu().numpy()[nout:], x[nout:])


@pytest.mark.parametrize("space,me_g0", [(xg.SPACE_CR, 1), (xg.SPACE_CR, 0), (xg.SPACE_C, -1)])
def test_colwise(lib, space, me_g0):
    rng = np.random.default_rng(7)
    rows, n = 3571, 9
    a = _blocks(rng, n, rows, False); b = _blocks(rng, n, rows, False); w = _blocks(rng, n, rows, False)
    da_, db_, dw_ = _dev(a), _dev(b), _dev(w)
    out = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    xg.xg_colwise("dot", space, rows, n, da_, rows, db_, rows, out=out, me_g0=me_g0)
    ref = oxg.colwise_dot(space, a, b, me_g0)
    got = out.cpu().numpy()
    got = got.view(np.complex128) if space == xg.SPACE_C else got[:n]
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))
    xg.xg_colwise("norm2", space, rows, n, da_, rows, out=out, me_g0=me_g0)
    assert np.max(np.abs(out.cpu().numpy()[:n] - oxg.colwise_norm2(space, a, me_g0))) < 1e-11
    lam = rng.standard_normal(n); dl = _dev(lam)
    xg.xg_colwise("cymax", space, rows, n, da_, rows, db_, rows, dw_, rows, da=dl)
    assert np.max(np.abs(da_.cpu().numpy() - oxg.colwise_cymax(lam, b, w))) < 1e-13
    xg.xg_colwise("scale", space, rows, n, db_, rows, da=dl)
    assert np.max(np.abs(db_.cpu().numpy() - lam[:, None] * b)) < 1e-13
    xg.xg_colwise("zero_im_g0", space, rows, n, dw_, rows, me_g0=me_g0)
    wz = w.copy(); oxg.zero_im_g0(space, wz, me_g0)
    assert np.array_equal(dw_.cpu().numpy(), wz)


@pytest.mark.parametrize("space,me_g0", [(xg.SPACE_CR, 1), (xg.SPACE_C, -1)])
@pytest.mark.parametrize("n", [6, 37])
def test_rayleigh_ritz(lib, space, me_g0, n):
    """Eigenvalues vs the oracle (LAPACK hegvd); eigenvectors through invariants (signs / phases are solver-specific)."""
    rng = np.random.default_rng(11 + n)
    rows = 400
    x = _blocks(rng, n, rows, me_g0 == 1)
    if space == xg.SPACE_CR:
        # a real-symmetric operator in the SPACE_CR metric: diagonal in G plus a low-rank real coupling
        d = rng.uniform(0, 5, rows); u = _blocks(rng, 3, rows, True)
        def op(c):
            return d[None, :] * c + (oxg.gram(xg.SPACE_CR, c, u, 1) @ u)
    else:
        hm = _blocks(rng, rows, rows, False); hm = hm + hm.conj().T
        def op(c):
            return c @ hm.T
    ax = op(x); bx = x.copy()
    w_ref, x_ref, ax_ref, _, _ = oxg.rayleigh_ritz(space, x, ax, bx, me_g0)
    dx, dax = _dev(x), _dev(ax)
    eig = np.zeros(n)
    info = xg.xg_RayleighRitz(dx, dax, None, eig, space, rows, n, me_g0=me_g0)
    assert info == 0
    assert np.max(np.abs(eig - w_ref)) < 1e-10 * max(1.0, np.max(np.abs(w_ref)))
    xr, axr = dx.cpu().numpy(), dax.cpu().numpy()
    assert np.max(np.abs(oxg.gram(space, xr, xr, me_g0) - np.eye(n))) < 1e-10
    assert np.max(np.abs(oxg.gram(space, xr, axr, me_g0) - np.diag(eig))) < 1e-9 * max(1.0, np.max(np.abs(eig)))
    assert np.max(np.abs(axr - op(xr))) < 1e-9 * np.max(np.abs(axr))
    # same vectors as the oracle's up to a sign / phase per (non-degenerate) column
    ov = np.abs(np.diag(oxg.gram(space, x_ref, xr, me_g0)))
    assert np.max(np.abs(ov - 1.0)) < 1e-8
    # separate BX block (PAW-style call) gives the same result
    dx2, dax2, dbx2 = _dev(x), _dev(ax), _dev(bx)
    eig2 = np.zeros(n)
    assert xg.xg_RayleighRitz(dx2, dax2, dbx2, eig2, space, rows, n, me_g0=me_g0) == 0
    assert np.max(np.abs(eig2 - eig)) < 1e-12 * max(1.0, np.max(np.abs(eig)))
    assert np.max(np.abs(dbx2.cpu().numpy() - dx2.cpu().numpy())) < 1e-12


@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0))])
@pytest.mark.parametrize("usepaw", [0, 1])
def test_nonlop_signs1_enlout(lib, istwf_k, kpt, usepaw):
    """nonlop(choice=1, signs=1): enlout = <psi|Vnl|psi> = Re <psi | gvnlxc> (oracle signs=2 result, contracted)."""
    p = make_problem(7.0, 8.5, kpt, istwf_k, ndat=6, natom_per_type=(1, 2), lmax_per_type=(2, 1), usepaw=usepaw)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, p.usepaw, p.ucvol)
    h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, p.sij); h.load_k(p.istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    gv, _, _ = onl.gemm_nonlop(P, p.cwavef, p.enl, p.sij, p.indlmn, p.nattyp, p.atindx1 - 1, istwf_k, choice=1,
                               paw_opt=usepaw, cpopt=-1, me_g0=1)
    space, me_g0 = (xg.SPACE_C, -1) if istwf_k == 1 else (xg.SPACE_CR, 1)
    ref = np.real(oxg.colwise_dot(space, p.cwavef, gv, me_g0))
    enl = np.zeros(p.ndat)
    ab.nonlop(1, -1, None, enl, h, 0, None, None, p.ndat, 1, usepaw, 1, None, 0, p.cwavef, None)
    assert np.max(np.abs(enl - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
    # signs=2 through the same dispatcher == gemm_nonlop
    out = np.zeros_like(p.cwavef)
    ab.nonlop(1, -1, None, None, h, 0, None, None, p.ndat, 1, usepaw, 2, None, 0, p.cwavef, out)
    assert np.max(np.abs(out - gv)) < 1e-11 * np.max(np.abs(gv))
    h.destroy()
