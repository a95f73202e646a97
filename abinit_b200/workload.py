"""Synthetic workloads of the named BASELINE.json shapes (host-side set-up only; no compute path here).

The G-sphere enumeration below exists so that bench.py / smoke() can build realistic (kg, kinpw) tables without
importing the test oracle.  It follows the ordering rules of kpgsph (src/52_fft_mpi_noabirule/m_fftcore.F90:
4040-4092: ig3 outer, ig2, ig1 inner, each 0..max,min..-1, half sphere for istwf_k=2) for orthorhombic cells.
"""
from __future__ import annotations
import numpy as np

HUGE = np.finfo(np.float64).max


def gsphere_orthorhombic(ecut: float, lengths, kpt=(0.0, 0.0, 0.0), istwf_k: int = 1):
    """Returns kg (npw, 3) int32 in Fortran memory order kg(3,npw), and kinpw (npw) = 1/2 (2pi)^2 |k+G|^2."""
    L = np.broadcast_to(np.asarray(lengths, dtype=np.float64), (3,))
    kpt = np.asarray(kpt, dtype=np.float64)
    gm = 1.0 / L ** 2                                    # diagonal metric
    gscut = 0.5 * ecut / np.pi ** 2
    nmax = np.floor(np.sqrt(gscut / gm) - kpt + 1e-10).astype(int)
    nmin = np.ceil(-np.sqrt(gscut / gm) - kpt - 1e-10).astype(int)

    def order(a, b):
        return np.concatenate([np.arange(0, b + 1), np.arange(a, 0)])
    g1, g2, g3 = order(nmin[0], nmax[0]), order(nmin[1], nmax[1]), order(nmin[2], nmax[2])
    if istwf_k >= 2:
        g2 = np.arange(0, nmax[1] + 1)
    G3, G2, G1 = np.meshgrid(g3, g2, g1, indexing="ij")
    q = gm[0] * (G1 + kpt[0]) ** 2 + gm[1] * (G2 + kpt[1]) ** 2 + gm[2] * (G3 + kpt[2]) ** 2
    keep = q <= gscut
    if 2 <= istwf_k <= 5:
        keep &= ~((G2 == 0) & (G3 < 0))
    if istwf_k in (2, 3):
        keep &= ~((G2 == 0) & (G3 == 0) & (G1 < 0))
    kg = np.stack([G1[keep], G2[keep], G3[keep]], axis=1).astype(np.int32)
    kin = 0.5 * (2 * np.pi) ** 2 * q[keep]
    return np.ascontiguousarray(kg), np.ascontiguousarray(kin)


def smooth_potential(ngfft, seed=0):
    """Sum of 8 random small-G cosines, amplitude 0.5 Ha, mean -0.3 Ha; shape (n3, n2, n1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n1, n2, n3 = ngfft
    x1 = np.arange(n1) / n1; x2 = np.arange(n2) / n2; x3 = np.arange(n3) / n3
    v = np.full((n3, n2, n1), -0.3)
    for _ in range(8):
        gv = rng.integers(-2, 3, size=3); ph = rng.uniform(0, 2 * np.pi)
        v += 0.0625 * np.cos(2 * np.pi * (gv[0] * x1[None, None, :] + gv[1] * x2[None, :, None] + gv[2] * x3[:, None, None]) + ph)
    return np.ascontiguousarray(v)


# name -> (ecut, cubic cell length, ngfft, natom, nlmn per atom, (l, n) channel list)
CONFIGS = {
    # Si 512-atom supercell NC, Gamma-only, ecut 20 Ha: box 180^3, npw 288113 (istwfk 1) / 144057 (istwfk 2),
    # nprojs = 512 * 18 (psp8 'nproj 2 2 2': l = 0,1,2 with 2 projectors each)
    "si512": dict(ecut=20.0, L=40.72, ngfft=(180, 180, 180), natom=512, lmax=2, nproj_per_l=2),
    # Au 108-atom-like shape (gemm_nonlop dominated): box 96^3
    "au108": dict(ecut=20.0, L=23.13, ngfft=(96, 96, 96), natom=108, lmax=2, nproj_per_l=2),
    # Si 2-atom-like tiny shape: box 24^3
    "si2": dict(ecut=12.0, L=7.2, ngfft=(24, 24, 24), natom=2, lmax=2, nproj_per_l=2),
    # mid-size synthetic sweep point
    "sweep96": dict(ecut=14.0, L=20.0, ngfft=(96, 96, 96), natom=64, lmax=2, nproj_per_l=2),
}


def nc_indlmn(lmax: int, nproj_per_l: int):
    """indlmn(6, lmnmax, 1) for one NC type in useylm order (l, n, m); returns (indlmn[1,lmnmax,6], lnmax)."""
    rows = []; iln = 0
    for l in range(lmax + 1):
        for n in range(nproj_per_l):
            iln += 1
            for m in range(-l, l + 1):
                rows.append((l, m, n + 1, l * l + l + m + 1, iln, 1))
    return np.asarray(rows, dtype=np.int32)[None], iln
