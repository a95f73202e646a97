"""psp8 (ONCVPSP) reader and the radial transforms ABINIT applies to it (oracle; test infrastructure only).

Restates  psp8in   src/64_psp/m_psp8.F90:98-487       (file layout, indlmn, ekb)
          psp8cc   src/64_psp/m_psp8.F90:509-617      (model core charge on the n1xccc grid, 4-point Lagrange)
          psp8lo   src/64_psp/m_psptk.F90:1143-1262   (epsatm, q^2 V_loc(q) by corrected-trapezoid sine transforms)
          psp8nl   src/64_psp/m_psptk.F90:1310-1477   (f_ln(q) = int j_l(2 pi q r) [r p_ln(r)] r dr)
          ctrap    shared/common/src/28_numeric_noabirule/m_numeric_tools.F90:2752-2826
          q grids  src/64_psp/m_pspini.F90:188-208    (qmax = 1.2 sqrt(gsqcut), mqgrid points)
The reference fits cubic splines to the q-grid tables (yp1/ypn from 5-point end formulas) and evaluates them at |G|
(m_mklocl.F90, m_mkffnl.F90); `ClampedSpline` is that spline (the same piecewise cubic, evaluated by SciPy).
Only used by tests/golden/make_si2_fixture.py (which reads the psp file from the reference tree) and the SCF pin tests.
"""
from __future__ import annotations
import numpy as np
from scipy.interpolate import CubicSpline
from scipy.special import spherical_jn


def _f(x):
    return float(x.replace("D", "E").replace("d", "e"))


class Psp8:
    pass


def read_psp8(path, useylm=1):
    """m_psp8.F90:98-343.  Returns header scalars, rad, vloc, projectors vpspll(mmax, lnmax), ekb(lnmax), the core-charge
    block ff(mmax,5) and indlmn(lmnmax,6) in the (l, m, n, lm, ln, spin) order of the reference (useylm=1)."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    p = Psp8()
    p.title = lines[0]
    t = lines[1].split(); p.zatom, p.zion = _f(t[0]), _f(t[1])
    t = lines[2].split(); p.pspcod, p.pspxc, p.lmax, p.lloc, p.mmax = int(t[0]), int(t[1]), int(t[2]), int(t[3]), int(t[4])
    assert p.pspcod == 8
    t = lines[3].split(); p.rchrg, p.fchrg, p.qchrg = _f(t[0]), _f(t[1]), _f(t[2])
    nproj = [int(x) for x in lines[4].split()[:p.lmax + 1]]
    p.nproj = nproj
    ext = int(lines[5].split()[0])
    assert ext in (0, 1), "spin-orbit psp8 not handled by the oracle"
    pos = 6
    mmax = p.mmax
    rad = np.zeros(mmax); vloc = None
    ekb = []; cols = []
    for l in range(p.lmax + 1):
        if nproj[l] > 0:
            t = lines[pos].split(); pos += 1
            assert int(t[0]) == l
            ekb += [_f(x) for x in t[1:1 + nproj[l]]]
            blk = np.array([[_f(x) for x in lines[pos + i].split()[1:2 + nproj[l]]] for i in range(mmax)]); pos += mmax
            rad = blk[:, 0]
            for n in range(nproj[l]):
                cols.append(blk[:, 1 + n])
        elif l == p.lloc:
            assert int(lines[pos].split()[0]) == l; pos += 1
            blk = np.array([[_f(x) for x in lines[pos + i].split()[1:3]] for i in range(mmax)]); pos += mmax
            rad = blk[:, 0]; vloc = blk[:, 1]
    if p.lloc > p.lmax:
        assert int(lines[pos].split()[0]) == p.lloc; pos += 1
        blk = np.array([[_f(x) for x in lines[pos + i].split()[1:3]] for i in range(mmax)]); pos += mmax
        rad = blk[:, 0]; vloc = blk[:, 1]
    p.rad = rad; p.vloc = vloc; p.vpspll = np.array(cols).T; p.ekb = np.array(ekb)
    amesh = rad[1] - rad[0]
    assert rad[0] == 0.0 and np.max(np.abs(np.diff(rad) - amesh)) < 1e-8
    p.amesh = amesh
    if p.fchrg > 1e-15:
        p.ffcore = np.array([[_f(x) for x in lines[pos + i].split()[2:7]] for i in range(mmax)]); pos += mmax
    else:
        p.ffcore = None
    # indlmn(6, lmnmax): l, m, n, lm, ln, spin  (m_psp8.F90:229-250)
    ind = []; iln = 0
    for l in range(p.lmax + 1):
        for kk in range(1, nproj[l] + 1):
            iln += 1
            for mm in range(1, 2 * l * useylm + 2):
                ind.append((l, mm - l * useylm - 1, kk, l * l + (1 - useylm) * l + mm, iln, 1))
    p.indlmn = np.array(ind, dtype=np.int32)
    p.lnmax = iln
    return p


def ctrap(ff, hh):
    """Corrected trapezoid rule, imax >= 10 branch (m_numeric_tools.F90:2768-2783).  ff may be (..., imax)."""
    ff = np.asarray(ff)
    n = ff.shape[-1]
    assert n >= 10
    w = np.array([23.75, 95.10, 55.20, 79.30, 70.65]) / 72.0
    endpt = np.sum(w * (ff[..., :5] + ff[..., :-6:-1]), axis=-1)
    return (np.sum(ff[..., 5:n - 5], axis=-1) + endpt) * hh


def _end_derivs(y, h):
    yp1 = (-50.0 * y[0] + 96.0 * y[1] - 72.0 * y[2] + 32.0 * y[3] - 6.0 * y[4]) / (24.0 * h)
    ypn = (6.0 * y[-5] - 32.0 * y[-4] + 72.0 * y[-3] - 96.0 * y[-2] + 50.0 * y[-1]) / (24.0 * h)
    return yp1, ypn


class ClampedSpline:
    """The reference's `spline` + `splfit` pair: cubic spline through (x, y) with prescribed end slopes."""

    def __init__(self, x, y, yp1, ypn):
        self.cs = CubicSpline(x, y, bc_type=((1, yp1), (1, ypn)))

    def __call__(self, xx):
        return self.cs(xx)


def qgrid(gsqcut, mqgrid=3001):
    """m_pspini.F90:188-208."""
    qmax = 1.2 * np.sqrt(gsqcut)
    return np.arange(mqgrid) * (qmax / (mqgrid - 1))


def psp8lo(p, qg):
    """epsatm and the spline of q^2 V_loc(q) (m_psptk.F90:1143-1262; mesh_mult must be 1 for the oracle)."""
    rad, vloc, zion, amesh = p.rad, p.vloc, p.zion, p.amesh
    rvlpz = rad * vloc + zion
    epsatm = 4.0 * np.pi * ctrap(rad * rvlpz, amesh)
    amesh_new = 2.0 * np.pi / (200 * qg[-1])
    assert int(amesh / amesh_new) + 1 == 1, "radial mesh refinement (mesh_mult>1) not restated"
    q2vq = np.empty_like(qg)
    q2vq[0] = -zion / np.pi
    arg = 2.0 * np.pi * qg[1:, None] * rad[None, :]
    q2vq[1:] = q2vq[0] + 2.0 * qg[1:] * ctrap(np.sin(arg) * rvlpz[None, :], amesh)
    yp1, ypn = _end_derivs(q2vq, qg[1] - qg[0])
    return epsatm, ClampedSpline(qg, q2vq, yp1, ypn), q2vq


def psp8nl(p, qg):
    """ffspl[iln](q) (m_psptk.F90:1310-1477; mesh_mult must be 1).  Returns a list of splines, one per (l,n) channel."""
    rad, amesh = p.rad, p.amesh
    amesh_new = 2.0 * np.pi / (200 * qg[-1])
    assert int(amesh / amesh_new) + 1 == 1
    v = np.where(np.abs(p.vpspll) > 1e-10, p.vpspll, 0.0)
    nz = np.nonzero(np.any(np.abs(p.vpspll) > 1e-10, axis=1))[0]
    mv = int(nz[-1]) + 1
    ls = []
    seen = 0
    for row in p.indlmn:
        if row[4] > seen:
            seen = row[4]; ls.append(int(row[0]))
    x = 2.0 * np.pi * qg[:, None] * rad[None, :mv]
    out = []
    for iln, l in enumerate(ls):
        tab = ctrap(spherical_jn(l, x) * (v[:mv, iln] * rad[:mv])[None, :], amesh)
        yp1, ypn = _end_derivs(tab, qg[1] - qg[0])
        out.append(ClampedSpline(qg, tab, yp1, ypn))
    return out


def psp8cc(p, n1xccc=2501):
    """xccc1d(n1xccc, 0:2): core density and its first two derivatives in units of the scaled radius r/rchrg
    (m_psp8.F90:586-611: 4-point Lagrange interpolation from the file grid, 1/4pi normalisation)."""
    ff, amesh, rchrg, mmax = p.ffcore, p.amesh, p.rchrg, p.mmax
    xx = np.arange(n1xccc) * rchrg / (n1xccc - 1)
    irad = np.clip((xx / amesh).astype(int) + 1, 2, mmax - 2)           # 1-based
    xp = (xx - p.rad[irad - 1]) / amesh
    c1 = -xp * (xp - 1) * (xp - 2) / 6.0
    c2 = (xp + 1) * (xp - 1) * (xp - 2) / 2.0
    c3 = -xp * (xp + 1) * (xp - 2) / 2.0
    c4 = xp * (xp + 1) * (xp - 1) / 6.0
    out = np.zeros((n1xccc, 5))
    for jj in range(5):
        t = c1 * ff[irad - 2, jj] + c2 * ff[irad - 1, jj] + c3 * ff[irad, jj] + c4 * ff[irad + 1, jj]
        out[:, jj] = t * rchrg ** jj / (4.0 * np.pi)
    return out
