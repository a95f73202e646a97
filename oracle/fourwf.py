"""fourwf: sphere <-> FFT box transforms with V_loc application (oracle; test infrastructure only).

Restates  fourwf                src/53_ffts/m_fft.F90:2290-2940   (options 0,1,2,3)
          sphere                src/52_fft_mpi_noabirule/m_fftcore.F90:1532-1866 (iflag=+1 scatter, TR completion)
          cg_box2gsph           src/44_abitools/m_cgtools.F90:2227-2311 (plain gather x 1/N)
          cg_vlocpsi            src/44_abitools/m_cgtools.F90:2410-2491
          cg_addtorho           src/44_abitools/m_cgtools.F90:2338-2384
cross-checked with the offload statement of the same maths, src/46_ghc_omp/m_ompgpu_fourwf.F90:308-560.
Sign convention: G->r is e^{+i 2 pi G.r} un-normalised ("FFT_INVERSE"), r->G is e^{-i..} times 1/(n1 n2 n3).
Array conventions here: C-ordered numpy views of the Fortran arrays, i.e.
  cg   (ndat, npw) complex128          == Fortran fofgin(2, npw*ndat)
  box  (ndat, n3, n2, n1) complex128   == Fortran fofr(2, n1, n2, n3*ndat)   (n4,n5,n6 == n1,n2,n3)
  vloc (n3, n2, n1) float64 (cplex=1) or complex128 (cplex=2) == Fortran denpot(cplex*n1, n2, n3)
  kg   (3, npw) int32 (numpy shape) -- same memory as Fortran kg_k(3,npw) when transposed; see abinit_b200.
"""
from __future__ import annotations
import numpy as np
import scipy.fft as sfft


def _wrap(kg, ngfft):
    n1, n2, n3 = ngfft
    i1 = np.where(kg[0] < 0, kg[0] + n1, kg[0]).astype(np.int64)
    i2 = np.where(kg[1] < 0, kg[1] + n2, kg[1]).astype(np.int64)
    i3 = np.where(kg[2] < 0, kg[2] + n3, kg[2]).astype(np.int64)
    return i1, i2, i3


def inverse_indices(i1, i2, i3, ngfft, istwf_k):
    """Time-reversed partner of each box index (0-based), m_fftcore.F90:1561-1592 /
    m_ompgpu_fourwf.F90:311-326: shift_inv = n for the directions where k is 0, n-1 where k is 1/2."""
    n1, n2, n3 = ngfft
    s1 = n1 if istwf_k in (2, 4, 6, 8) else n1 - 1
    s2 = n2 if 2 <= istwf_k <= 5 else n2 - 1
    s3 = n3 if istwf_k in (2, 3, 6, 7) else n3 - 1
    return np.mod(s1 - i1, n1), np.mod(s2 - i2, n2), np.mod(s3 - i3, n3)


def sphere_to_box(cg, kg, ngfft, istwf_k=1, me_g0=1):
    """sphere(iflag=1): zero box, insert cg; for istwf_k>=2 also insert conj(cg) at the time-reversed point
    and force Im=0 at G=0 when istwf_k==2 (m_fftcore.F90:1624-1650)."""
    cg = np.atleast_2d(cg)
    ndat, npw = cg.shape
    n1, n2, n3 = ngfft
    box = np.zeros((ndat, n3, n2, n1), dtype=np.complex128)
    i1, i2, i3 = _wrap(kg, ngfft)
    box[:, i3, i2, i1] = cg
    if istwf_k >= 2:
        lo = 0
        if istwf_k == 2 and me_g0 == 1:
            box[:, 0, 0, 0] = cg[:, 0].real
            lo = 1
        j1, j2, j3 = inverse_indices(i1[lo:], i2[lo:], i3[lo:], ngfft, istwf_k)
        box[:, j3, j2, j1] = np.conj(cg[:, lo:])
    return box


def box_to_sphere(box, kg, ngfft, istwf_k=1, me_g0=1):
    """Plain gather times 1/N (cg_box2gsph, m_cgtools.F90:2227-2311; offload twin m_ompgpu_fourwf.F90:531-560):
    no time-reversal averaging on the way out; Im forced to 0 at G=0 for istwf_k==2."""
    n1, n2, n3 = ngfft
    i1, i2, i3 = _wrap(kg, ngfft)
    xnorm = 1.0 / float(n1 * n2 * n3)
    out = box[:, i3, i2, i1] * xnorm
    if istwf_k == 2 and me_g0 == 1:
        out[:, 0] = out[:, 0].real
    return out


def fourwf(cplex, denpot, fofgin, fofr, kg_kin, kg_kout, ngfft, option, istwf_k=1,
           weight_r=1.0, weight_i=1.0, me_g0=1, workers=None):
    """Returns (fofgout, fofr, denpot) with the entries the reference would have written for `option`:
      0: fofr = FFT^-1[fofgin]                      (m_fft.F90:2201)
      1: denpot += w_r Re(psi(r))^2 + w_i Im(psi(r))^2 over ndat  (m_fft.F90:2633-2653, cg_addtorho)
      2: fofgout = gather( FFT[ V * FFT^-1[scatter(fofgin)] ] ) / N
      3: fofgout = gather( FFT[fofr] ) / N
    weight_r / weight_i may be scalars or arrays of ndat (weight_array_r/i of the GPU entry point)."""
    n1, n2, n3 = ngfft
    fofgout = None
    if option not in (0, 1, 2, 3):
        raise ValueError("Only option=0, 1, 2 or 3 are allowed presently.")
    if option == 1 and cplex != 1:
        raise ValueError("With the option number 1, cplex must be 1")
    if option == 2 and cplex not in (1, 2):
        raise ValueError("With the option number 2, cplex must be 1 or 2")
    if option != 3:
        box = sphere_to_box(fofgin, kg_kin, ngfft, istwf_k, me_g0)
        ur = sfft.ifftn(box, axes=(1, 2, 3), norm="forward", workers=workers)   # e^{+i}, no scaling
    else:
        ur = np.asarray(fofr).reshape(-1, n3, n2, n1)
    if option == 0:
        return None, ur, denpot
    if option == 1:
        ndat = ur.shape[0]
        wr = np.broadcast_to(np.asarray(weight_r, dtype=np.float64), (ndat,))
        wi = np.broadcast_to(np.asarray(weight_i, dtype=np.float64), (ndat,))
        den = np.array(denpot, dtype=np.float64, copy=True)
        for idat in range(ndat):
            den += wr[idat] * ur[idat].real ** 2 + wi[idat] * ur[idat].imag ** 2
        return None, ur, den
    if option == 2:
        ur = ur * denpot[None]      # real or complex V (cg_vlocpsi)
    ug = sfft.fftn(ur, axes=(1, 2, 3), norm="backward", workers=workers)       # e^{-i}, unscaled
    fofgout = box_to_sphere(ug, kg_kout, ngfft, istwf_k, me_g0)
    return fofgout, ur, denpot
