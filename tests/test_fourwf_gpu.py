"""GPU parity: fourwf through the C-ABI vs the oracle (FP64, tolerance 1e-11 relative per band, north-star)."""
import numpy as np
import pytest
from oracle import fourwf as ofw, gsphere as g
from problems import make_problem, rel_err_per_band

pytestmark = pytest.mark.gpu
TOL = 1e-11   # BASELINE.json north_star: "match the reference's CPU getghc ... to 1e-11 relative in FP64"


def _run(lib, p, option, impl, weight_r=0.7, weight_i=0.3):
    n1, n2, n3 = p.ngfft
    rng = np.random.default_rng(7)
    out = np.zeros((p.ndat, p.npw), dtype=np.complex128)
    fofr = np.zeros((p.ndat, n3, n2, n1), dtype=np.complex128)
    den = p.vlocal.copy()
    if option == 1:
        den = np.ascontiguousarray(np.abs(p.vlocal.real))
    if option == 3:
        fofr[:] = rng.standard_normal(fofr.shape) + 1j * rng.standard_normal(fofr.shape)
    ref_out, ref_r, ref_den = ofw.fourwf(p.cplex, den.copy(), p.cwavef, fofr.copy(), p.kg, p.kg, p.ngfft, option,
                                         p.istwf_k, weight_r=weight_r, weight_i=weight_i)
    lib.fourwf(p.cplex, den, p.cwavef, out, fofr, None, None, p.istwf_k, p.kgF, p.kgF, max(p.ngfft), None, p.ndat,
               p.ngfft, p.npw, p.npw, n1, n2, n3, option, weight_r=weight_r, weight_i=weight_i, impl=impl)
    if option in (2, 3):
        return rel_err_per_band(out, ref_out)
    if option == 0:
        return rel_err_per_band(fofr, ref_r)
    return rel_err_per_band(den[None], ref_den[None])


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0)), (3, (.5, 0, 0)), (4, (0, 0, .5)),
                                          (5, (.5, 0, .5)), (6, (0, .5, 0)), (7, (.5, .5, 0)), (8, (0, .5, .5)),
                                          (9, (.5, .5, .5))])
def test_option2_all_istwfk(lib, impl, istwf_k, kpt):
    p = make_problem(6.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=3)
    assert _run(lib, p, 2, impl) < TOL


@pytest.mark.parametrize("impl", [1, 2])
def test_option2_complex_potential(lib, impl):
    p = make_problem(6.0, 8.0, (.1, .2, .3), 1, ndat=2, cplex=2)
    assert _run(lib, p, 2, impl) < TOL


@pytest.mark.parametrize("option", [0, 1, 3])
@pytest.mark.parametrize("istwf_k,kpt", [(1, (.1, .2, .3)), (2, (0, 0, 0))])
def test_options_0_1_3(lib, option, istwf_k, kpt):
    p = make_problem(6.0, 8.0, kpt, istwf_k, ndat=3)
    assert _run(lib, p, option, 0) < TOL


@pytest.mark.parametrize("ngfft", [(28, 35, 21), (24, 24, 24), (45, 50, 54), (56, 60, 64), (25, 27, 32), (96, 96, 96)])
def test_option2_mixed_radix_boxes(lib, ngfft):
    """radices 2,3,5,7 (7-smooth sizes are explicit ngfft cases) in every pass position."""
    p = make_problem(4.0, 9.0, (.1, 0, .3), 1, ndat=2, ngfft=ngfft)
    for impl in (1, 2):
        assert _run(lib, p, 2, impl) < TOL


def _plane_problem(n, kpt, istwf_k, ndat, n1=None, cplex=1):
    """Sphere diameter ~ n/2 (boxcut 2) so that the pruned passes see realistic occupancies; n1 < n shrinks the cell
    along x to keep the oracle's full-box FFT small at the largest n."""
    L = 10.0
    gmax = n / 4.0 - 0.75
    ecut = 0.5 * (gmax * 2 * np.pi / L) ** 2
    n1 = n1 or n
    return make_problem(ecut, (L * n1 / n, L, L), kpt, istwf_k, ndat=ndat, ngfft=(n1, n, n), cplex=cplex)


@pytest.mark.parametrize("n", [24, 30, 32, 36, 40, 45, 48, 50, 54, 56, 60, 64, 72, 75, 80, 81, 84, 90, 96, 100, 108, 112,
                               120, 128, 135, 144, 150, 160, 168, 180, 192, 196, 225, 240, 256])
def test_option2_plane_stage_every_length(lib, n):
    """Every FFT length of the register-resident two-pass plane stage (plane_stage.cuh), both warp layouts."""
    from abinit_b200 import api
    big = n > 128
    p = _plane_problem(n, (.1, .2, .3), 1, ndat=1 if big else 2, n1=(n if not big else 36))
    for cfg in (1, 2):
        api.set_tuning("plane_cfg", cfg)
        try:
            assert _run(lib, p, 2, 2) < TOL
        finally:
            api.set_tuning("plane_cfg", 0)


@pytest.mark.parametrize("istwf_k,kpt", [(2, (0, 0, 0)), (3, (.5, 0, 0)), (5, (.5, 0, .5)), (8, (0, .5, .5)), (9, (.5, .5, .5))])
def test_option2_plane_stage_time_reversal(lib, istwf_k, kpt):
    p = _plane_problem(48, kpt, istwf_k, ndat=3, n1=45)
    assert _run(lib, p, 2, 2) < TOL
    from abinit_b200 import api
    api.set_tuning("plane", 0)          # same box through the cluster/L2 plane kernel (fallback path)
    try:
        assert _run(lib, p, 2, 2) < TOL
    finally:
        api.set_tuning("plane", 1)


def test_option2_plane_stage_complex_potential_many_bands(lib):
    p = _plane_problem(60, (.1, .2, .3), 1, ndat=7, cplex=2)
    assert _run(lib, p, 2, 2) < TOL


def _split_problem(n1, n2, n3, kpt, istwf_k, ndat, cplex=1):
    """Orthorhombic cell whose boxcut-2 box is (n1, n2, n3) with n2 != n3: the split plane stage (three kernels, S in HBM)."""
    L = 10.0
    gmax = n2 / 4.0 - 0.75
    ecut = 0.5 * (gmax * 2 * np.pi / L) ** 2
    return make_problem(ecut, (L * n1 / n2, L, L * n3 / n2), kpt, istwf_k, ndat=ndat, ngfft=(n1, n2, n3), cplex=cplex)


@pytest.mark.parametrize("istwf_k,kpt,ndat", [(1, (.1, .2, .3), 3), (2, (0, 0, 0), 5), (2, (0, 0, 0), 4), (3, (.5, 0, 0), 2),
                                               (6, (0, .5, 0), 2), (9, (.5, .5, .5), 3)])
@pytest.mark.parametrize("ngfft", [(45, 48, 60), (36, 60, 40), (30, 72, 96)])
def test_option2_split_plane_stage_non_cubic(lib, ngfft, istwf_k, kpt, ndat):
    """n2 != n3 with both lengths in the two-pass table: y / z / y^-1 as three register-resident kernels (plane_stage.cuh,
    k_fw_plane_split) instead of the cluster kernel; all time-reversal spheres, packed Gamma pairs (odd and even ndat)."""
    p = _split_problem(*ngfft, kpt, istwf_k, ndat)
    from abinit_b200 import api
    api.profile_enable(True)
    try:
        assert _run(lib, p, 2, 2) < TOL
        prof = api.profile_collect()
    finally:
        api.profile_enable(False)
    assert "fourwf_plane_stage" in prof and "fourwf_plane_cluster" not in prof     # the split path really ran
    api.set_tuning("plane", 0)                      # same box through the cluster/L2 plane kernel
    try:
        assert _run(lib, p, 2, 2) < TOL
    finally:
        api.set_tuning("plane", 1)


def test_option2_split_plane_stage_complex_potential(lib):
    p = _split_problem(40, 50, 64, (.1, .2, .3), 1, ndat=5, cplex=2)
    assert _run(lib, p, 2, 2) < TOL


@pytest.mark.parametrize("istwf_k,kpt,ndat", [(1, (.1, .2, .3), 3), (2, (0, 0, 0), 5), (2, (0, 0, 0), 1), (9, (.5, .5, .5), 2)])
def test_option1_split_plane_stage_density(lib, istwf_k, kpt, ndat):
    """fourwf option 1 on a non-cubic box: y kernel + z-with-density-reduction kernel of the split plane stage."""
    p = _split_problem(45, 48, 60, kpt, istwf_k, ndat)
    n1, n2, n3 = p.ngfft
    rng = np.random.default_rng(12)
    wr = rng.uniform(0.1, 2.0, ndat); wi = rng.uniform(0.1, 2.0, ndat)
    den0 = np.ascontiguousarray(np.abs(p.vlocal.real))
    _, _, ref = ofw.fourwf(1, den0.copy(), p.cwavef, None, p.kg, p.kg, p.ngfft, 1, p.istwf_k, weight_r=wr, weight_i=wi)
    from abinit_b200 import api
    for impl in (0, 1):
        den = den0.copy()
        api.profile_enable(True)
        try:
            lib.fourwf(1, den, p.cwavef, None, None, None, None, p.istwf_k, p.kgF, p.kgF, max(p.ngfft), None, ndat, p.ngfft, p.npw,
                       p.npw, n1, n2, n3, 1, weight_array_r=wr, weight_array_i=wi, impl=impl)
            prof = api.profile_collect()
        finally:
            api.profile_enable(False)
        assert np.max(np.abs(den - ref)) < 1e-11 * np.max(np.abs(ref)), impl
        assert ("fourwf_plane_rho" in prof) == (impl == 0)       # impl 0: fused path (y kernel + z-with-density kernel)


@pytest.mark.parametrize("ecut,kpt,istwf_k,ngfft", [(0.01, (0, 0, 0), 2, (24, 24, 24)), (0.01, (0, 0, 0), 1, (24, 24, 24)),
                                                     (0.01, (0, 0, 0), 2, (24, 30, 36)), (0.3, (.5, .5, .5), 9, (24, 24, 24)),
                                                     (0.3, (.1, .2, .3), 1, (24, 24, 30))])
def test_tiny_spheres(lib, ecut, kpt, istwf_k, ngfft):
    """Degenerate spheres (npw = 1: G = 0 only; npw = 4) on the fused, split and generic paths, options 2 and 1: one occupied
    line / plane, a single band and an odd band count at Gamma (half-filled packed transform)."""
    for ndat in (1, 3):
        p = make_problem(ecut, 8.0, kpt, istwf_k, ndat=ndat, ngfft=ngfft)
        assert p.npw <= 4
        for impl in (1, 2):
            assert _run(lib, p, 2, impl) < TOL
        assert _run(lib, p, 1, 0) < TOL


@pytest.mark.parametrize("ndat", [1, 5, 16])
def test_option2_ndat_and_clusters(lib, ndat, monkeypatch):
    p = make_problem(8.0, 10.0, (0, 0, 0), 1, ndat=ndat)
    assert _run(lib, p, 2, 2) < TOL


def test_fftprof_known_answer(lib):
    """The reference's own unit-test vectors (src/70_gw/m_fft_prof.F90:873,936-960): c(G)=exp(-(2pi)^2 G.gmet.G),
    V=cos(2pi g0.r), g0=(1,-1,2) => out(G) = 1/2 [c(G-g0) + c(G+g0)] (zero outside the sphere).
    Tolerance: the cross-library spread stored in tests/unitary/Refs/tfourwf_01.stdout:129 (3.4e-16) x 10."""
    ecut, L, kpt = 10.0, 12.0, (.1, .2, .3)
    _, gmet, _ = g.metric(np.eye(3) * L)
    ng = g.getng(2.0, ecut, gmet, kpt)
    kg = g.kpgsph(ecut, gmet, kpt, 1)
    npw = kg.shape[1]
    gsq = (2 * np.pi) ** 2 * np.einsum("ip,ij,jp->p", kg, gmet, kg.astype(float))
    c = np.ascontiguousarray(np.exp(-gsq)[None, :].astype(np.complex128))
    g0 = np.array([1, -1, 2])
    n1, n2, n3 = ng
    i3, i2, i1 = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
    V = np.ascontiguousarray(np.cos(2 * np.pi * (g0[0] * i1 / n1 + g0[1] * i2 / n2 + g0[2] * i3 / n3)))
    full = ofw.sphere_to_box(c, kg, ng, 1)[0]
    exp = 0.5 * (np.roll(full, (g0[2], g0[1], g0[0]), (0, 1, 2)) + np.roll(full, (-g0[2], -g0[1], -g0[0]), (0, 1, 2)))
    w1, w2, w3 = ofw._wrap(kg, ng)
    ref = exp[w3, w2, w1]
    kgF = np.ascontiguousarray(kg.T)
    for impl in (1, 2):
        out = np.zeros((1, npw), dtype=np.complex128)
        lib.fourwf(1, V, c, out, None, None, None, 1, kgF, kgF, max(ng), None, 1, ng, npw, npw, n1, n2, n3, 2, impl=impl)
        assert np.abs(out[0] - ref).max() < 3.4e-15


def test_device_pointers_and_roundtrip(lib):
    """Device-resident arguments are used in place (m_getghc.F90:378-394 contract); option 0 then option 3 is the
    identity on the sphere (fftbox/fftu round-trip tests, src/53_ffts/m_fft.F90:1445-1656)."""
    import torch
    p = make_problem(8.0, 10.0, (.1, .2, .3), 1, ndat=4)
    n1, n2, n3 = p.ngfft
    d_c = torch.from_numpy(p.cwavef.view(np.float64).reshape(p.ndat, p.npw, 2)).cuda()
    d_r = torch.zeros((p.ndat, n3, n2, n1, 2), dtype=torch.float64, device="cuda")
    d_o = torch.zeros_like(d_c)
    lib.fourwf(1, None, d_c, None, d_r, None, None, 1, p.kgF, p.kgF, max(p.ngfft), None, p.ndat, p.ngfft, p.npw, p.npw,
               n1, n2, n3, 0)
    lib.fourwf(1, None, None, d_o, d_r, None, None, 1, p.kgF, p.kgF, max(p.ngfft), None, p.ndat, p.ngfft, p.npw, p.npw,
               n1, n2, n3, 3)
    torch.cuda.synchronize()
    assert rel_err_per_band(d_o.cpu().numpy(), d_c.cpu().numpy()) < TOL


@pytest.mark.parametrize("istwf_k,kpt,ndat", [(1, (.1, .2, .3), 3), (2, (0, 0, 0), 5), (2, (0, 0, 0), 4), (2, (0, 0, 0), 1),
                                               (3, (.5, 0, 0), 2), (9, (.5, .5, .5), 3)])
def test_option1_fused_density(lib, istwf_k, kpt, ndat):
    """fourwf option 1 on the fused zero-padded path (plane stage + density reduction; two bands per transform at Gamma)
    vs the oracle and vs the generic full-box kernels, per-band weights."""
    p = _plane_problem(48, kpt, istwf_k, ndat=ndat, n1=45)
    n1, n2, n3 = p.ngfft
    rng = np.random.default_rng(11)
    wr = rng.uniform(0.1, 2.0, ndat); wi = rng.uniform(0.1, 2.0, ndat)
    den0 = np.ascontiguousarray(np.abs(p.vlocal.real))
    _, _, ref = ofw.fourwf(1, den0.copy(), p.cwavef, None, p.kg, p.kg, p.ngfft, 1, p.istwf_k, weight_r=wr, weight_i=wi)
    res = {}
    for impl in (0, 1):
        den = den0.copy()
        k0 = lib.kernel_launches()
        lib.fourwf(1, den, p.cwavef, None, None, None, None, p.istwf_k, p.kgF, p.kgF, max(p.ngfft), None, ndat, p.ngfft, p.npw,
                   p.npw, n1, n2, n3, 1, weight_array_r=wr, weight_array_i=wi, impl=impl)
        res[impl] = (den, lib.kernel_launches() - k0)
        assert np.max(np.abs(den - ref)) < 1e-11 * np.max(np.abs(ref)), impl
    # the fused path really ran: 3 kernels per band chunk + weights + transpose-add, fewer than the generic path's passes
    assert res[0][1] <= 5


def test_option1_fused_every_length_sample(lib):
    for n in (24, 45, 60, 96, 180):
        p = _plane_problem(n, (.1, .2, .3), 1, ndat=2, n1=(n if n <= 128 else 36))
        assert _run(lib, p, 1, 0) < TOL
