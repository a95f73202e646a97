"""G-sphere, FFT-box and kinetic-energy set-up (oracle; test infrastructure only).

Restates, in NumPy, the reference routines
  kpgsph          src/52_fft_mpi_noabirule/m_fftcore.F90:3911-4250
  bound / getng   src/52_fft_mpi_noabirule/m_fftcore.F90:509-635, 691-1214
  sphereboundary  src/52_fft_mpi_noabirule/m_fftcore.F90:1253-1471
  mkkin           src/56_recipspace/m_kg.F90:325-460
  ph1d3d          src/56_recipspace/m_kg.F90:644-700
"""
from __future__ import annotations
import numpy as np

PI = np.pi
TOL10 = 1.0e-10
TOL12 = 1.0e-12
HUGE = np.finfo(np.float64).max          # Fortran huge(0.0_dp)
KIN_SENTINEL = HUGE * 1.0e-10            # m_kg.F90:422-429
KIN_FILTER = HUGE * 1.0e-11              # m_getghc.F90:1272


def metric(rprimd: np.ndarray):
    """rprimd[:, i] = i-th real-space primitive vector (Bohr). Returns gprimd, gmet, ucvol
    (src/41_geometry/m_geometry.F90 metric/matr3inv conventions: gprimd = inv(rprimd)^T)."""
    rprimd = np.asarray(rprimd, dtype=np.float64)
    gprimd = np.linalg.inv(rprimd).T
    gmet = gprimd.T @ gprimd
    ucvol = abs(np.linalg.det(rprimd))
    return gprimd, gmet, ucvol


def _std_order(nmax: int, ngrid: int) -> np.ndarray:
    """0 1 2 ... nmax nmin ... -1  (m_fftcore.F90:4005-4013)."""
    v = np.arange(ngrid)
    return np.where(v > nmax, v - ngrid, v)


def kpgsph(ecut: float, gmet: np.ndarray, kpt, istwf_k: int = 1) -> np.ndarray:
    """Reduced coordinates kg(3, npw) of the plane waves with 1/2 (2 pi)^2 |k+G|^2 <= ecut.

    Ordering and half-sphere rules follow m_fftcore.F90:4040-4092 (exchn2n3d=0, no MPI-FFT):
    ig3 outer, ig2, ig1 inner, each in the order 0..max,min..-1; for istwf_k>=2 only ig2>=0 is kept,
    for istwf_k in 2..5 the ig2=0 row is dropped when ig3<0, and for istwf_k in {2,3} the
    (ig2=0, ig3=0) line keeps ig1>=0 only.  Returns int32 array of shape (3, npw) (Fortran kg(3,npw))."""
    if not 1 <= istwf_k <= 9:
        raise ValueError("istwf_k must be between 1 and 9")
    gmet = np.asarray(gmet, dtype=np.float64)
    kpt = np.asarray(kpt, dtype=np.float64)
    gscut = 0.5 * ecut * (1.0 / PI) ** 2
    minor = np.array([gmet[1, 1] * gmet[2, 2] - gmet[1, 2] ** 2,
                      gmet[0, 0] * gmet[2, 2] - gmet[0, 2] ** 2,
                      gmet[1, 1] * gmet[0, 0] - gmet[0, 1] ** 2])
    numer = np.array([
        gmet[0, 1] ** 2 * gmet[2, 2] - 2.0 * gmet[0, 1] * gmet[0, 2] * gmet[1, 2] + gmet[0, 2] ** 2 * gmet[1, 1],
        gmet[1, 2] ** 2 * gmet[0, 0] - 2.0 * gmet[0, 1] * gmet[0, 2] * gmet[1, 2] + gmet[1, 0] ** 2 * gmet[2, 2],
        gmet[2, 1] ** 2 * gmet[0, 0] - 2.0 * gmet[0, 1] * gmet[0, 2] * gmet[1, 2] + gmet[0, 2] ** 2 * gmet[1, 1]])
    nmax = np.zeros(3, dtype=int); nmin = np.zeros(3, dtype=int); ngrid = np.zeros(3, dtype=int)
    for ii in range(3):
        xx = gmet[ii, ii] * minor[ii] - numer[ii]
        kmax = np.sqrt(gscut * minor[ii] / xx)
        nmax[ii] = int(np.floor(kmax - kpt[ii] + TOL10))
        nmin[ii] = int(np.ceil(-kmax - kpt[ii] - TOL10))
        ngrid[ii] = nmax[ii] - nmin[ii] + 1
    ig1arr = _std_order(nmax[0], ngrid[0]); kg1 = kpt[0] + ig1arr
    ig2pmax = ngrid[1] if istwf_k < 2 else nmax[1] + 1
    ig2arr = _std_order(nmax[1], ngrid[1])[:ig2pmax]
    ig3arr = _std_order(nmax[2], ngrid[2])
    out = []
    for ig3p in range(ngrid[2]):
        ig3 = ig3arr[ig3p]; v3 = kpt[2] + ig3
        ig2pmin = 1 if (2 <= istwf_k <= 5 and ig3 < 0) else 0
        for ig2p in range(ig2pmin, ig2pmax):
            ig2 = ig2arr[ig2p]; v2 = kpt[1] + ig2
            gs_part = gmet[1, 1] * v2 * v2 + gmet[2, 2] * v3 * v3 + 2.0 * gmet[1, 2] * v2 * v3
            gs_fact = 2.0 * (gmet[0, 1] * v2 + gmet[2, 0] * v3)
            ig1pmax = ngrid[0]
            if istwf_k in (2, 3) and ig3p == 0 and ig2p == 0:
                ig1pmax = nmax[0] + 1
            v1 = kg1[:ig1pmax]
            gmin = gs_part + v1 * (gs_fact + v1 * gmet[0, 0])
            sel = np.nonzero(gmin <= gscut)[0]
            if sel.size:
                blk = np.empty((3, sel.size), dtype=np.int32)
                blk[0] = ig1arr[sel]; blk[1] = ig2; blk[2] = ig3
                out.append(blk)
    if not out:
        return np.zeros((3, 0), dtype=np.int32)
    return np.ascontiguousarray(np.concatenate(out, axis=1))


def _dsq(i1, i2, i3, gmet, kpt):
    a = kpt[0] + i1; b = kpt[1] + i2; c = kpt[2] + i3
    return (gmet[0, 0] * a * a + gmet[1, 1] * b * b + gmet[2, 2] * c * c
            + 2.0 * (gmet[0, 1] * a * b + gmet[1, 2] * b * c + gmet[2, 0] * c * a))


def bound(gmet, kpt, ngfft):
    """Smallest |k+G|^2 on the faces of the FFT box and the face ('plane' 1..3) where it occurs
    (m_fftcore.F90:509-635)."""
    n = [int(x) // 2 for x in ngfft[:3]]
    best = float(_dsq(n[0], -n[1], -n[2], gmet, kpt)) + 0.01
    plane = 0
    r = [np.arange(-n[i], n[i] + 1, dtype=np.float64) for i in range(3)]
    # the reference scans plane 1, then 2, then 3 with strict '<' -> the first strict minimum wins
    for pl in (1, 2, 3):
        if pl == 1:
            A, B = np.meshgrid(r[1], r[2], indexing="ij")
            vals = np.minimum(_dsq(float(n[0]), A, B, gmet, kpt), _dsq(-float(n[0]), A, B, gmet, kpt))
        elif pl == 2:
            A, B = np.meshgrid(r[0], r[2], indexing="ij")
            vals = np.minimum(_dsq(A, float(n[1]), B, gmet, kpt), _dsq(A, -float(n[1]), B, gmet, kpt))
        else:
            A, B = np.meshgrid(r[0], r[1], indexing="ij")
            vals = np.minimum(_dsq(A, B, float(n[2]), gmet, kpt), _dsq(A, B, -float(n[2]), gmet, kpt))
        m = float(vals.min())
        if m < best:
            best = m; plane = pl
    return best, plane


def fft_sizes(limit: int = 4096, primes=(2, 3, 5)):
    """Allowed 1-D FFT lengths: 2^a 3^b 5^c (powers of 7 and 11 are '#if 0'-ed, m_fftcore.F90:720-731)."""
    s = {1}
    for p in primes:
        new = set()
        for v in s:
            w = v
            while w <= limit:
                new.add(w); w *= p
        s = new
    return sorted(s)


def getng(boxcutmin: float, ecut: float, gmet, kpt=(0.0, 0.0, 0.0), sizes=None):
    """Smallest allowed FFT box containing the boxcutmin-scaled sphere (m_fftcore.F90:880-907).
    Symmetry-commensurability post-processing (m_fftcore.F90:915-1100) is out of scope (nsym=1)."""
    srch = fft_sizes() if sizes is None else list(sizes)
    ngfft = [2, 2, 2]
    target = 0.5 * boxcutmin ** 2 * ecut / PI ** 2
    while True:
        dsqmin, plane = bound(np.asarray(gmet), np.asarray(kpt, dtype=np.float64), ngfft)
        if dsqmin >= target:
            break
        p = plane - 1
        for ii in range(len(srch) - 1):
            if srch[ii] >= ngfft[p]:
                ngfft[p] = srch[ii + 1]
                break
        else:
            raise RuntimeError("ngfft is bigger than allowed value")
    return tuple(ngfft)


def mkkin(ecut: float, ecutsm: float, effmass_free: float, gmet, kg, kpt) -> np.ndarray:
    """kinpw(npw) = 1/2 (2 pi)^2 |k+G|^2 / effmass, with the ecutsm smoothing and the huge*1e-10
    sentinel above ecut (m_kg.F90:386-455, order 0, no vector potential)."""
    gmet = np.asarray(gmet); kg = np.asarray(kg); kpt = np.asarray(kpt, dtype=np.float64)
    htpisq = 0.5 * (2.0 * PI) ** 2
    g = kg.astype(np.float64) + kpt[:, None]
    kpg2 = htpisq * (gmet[0, 0] * g[0] ** 2 + gmet[1, 1] * g[1] ** 2 + gmet[2, 2] * g[2] ** 2
                     + 2.0 * (g[0] * gmet[0, 1] * g[1] + g[0] * gmet[0, 2] * g[2] + g[1] * gmet[1, 2] * g[2]))
    kin = kpg2.copy()
    ecutsm_inv = 1.0 / ecutsm if ecutsm > 1.0e-20 else 0.0
    hi = kpg2 > ecut - ecutsm
    if np.any(hi):
        top = hi & (kpg2 > ecut - TOL12)
        mid = hi & ~top
        kin[top] = KIN_SENTINEL
        if np.any(mid):
            xx = np.maximum((ecut - kpg2[mid]) * ecutsm_inv, 1.0e-20)
            fsm = 1.0 / (xx ** 2 * (3 + xx * (1 + xx * (-6 + 3 * xx))))
            kin[mid] = kpg2[mid] * fsm
    ok = kin < KIN_SENTINEL
    kin[ok] = kin[ok] / effmass_free
    return kin


def ph3d(kg, kpt, xred) -> np.ndarray:
    """ph3d(natom, npw) = exp(+2 pi i (k+G).xred_a)  (m_kg.F90:644-700, ph1d3d; atoms must already be
    sorted by type, i.e. xred is the 'atindx'-ordered array)."""
    kg = np.asarray(kg); xred = np.asarray(xred, dtype=np.float64)
    kpg = kg.astype(np.float64) + np.asarray(kpt, dtype=np.float64)[:, None]     # (3, npw)
    arg = 2.0 * PI * (xred.T @ kpg)                                            # (natom, npw)
    return np.exp(1j * arg)


def sphereboundary(kg, istwf_k: int, mgfft: int) -> np.ndarray:
    """gbound(2*mgfft+8, 2) as built by sphereboundary (m_fftcore.F90:1253-1471).

    Column 1 (G->r): [g3min, g3max, then for each plane index (g2min,g2max) ..]; the product path does not
    consume gbound (it derives its line tables from kg), so only the layout needed by callers that
    size/forward the array is restated: entries (1,2) = min/max of the slowest index, then per-plane
    (min,max) pairs of the second index for the first-direction pass, and the mirrored table in column 2.
    """
    kg = np.asarray(kg)
    gb = np.zeros((2 * mgfft + 8, 2), dtype=np.int32)
    if kg.shape[1] == 0:
        return gb
    for col, (a, b) in enumerate(((2, 1), (1, 0))):   # (slow, fast) index pairs: (g3;g2) then (g2;g1)
        ga = kg[a].astype(int); gbv = kg[b].astype(int)
        if istwf_k >= 2:
            ga = np.concatenate([ga, -ga]); gbv = np.concatenate([gbv, -gbv])
        amin, amax = int(ga.min()), int(ga.max())
        gb[0, col] = amin; gb[1, col] = amax
        idx = 2
        order = list(range(0, amax + 1)) + list(range(amin, 0))
        for v in order:
            sel = gbv[ga == v]
            if idx + 1 >= gb.shape[0]:
                break
            if sel.size:
                gb[idx, col] = int(sel.min()); gb[idx + 1, col] = int(sel.max())
            else:
                gb[idx, col] = 0; gb[idx + 1, col] = -1
            idx += 2
    return gb
