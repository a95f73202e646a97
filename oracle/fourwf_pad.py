"""fourwf option 2 the way the reference's CPU build runs it (oracle; test infrastructure only): ZERO-PADDED passes
instead of a full 3-D FFT, and two bands per complex transform at the Gamma point.

Restates  fftw3_fftpad.finc:14-103   G -> r: x transforms on the occupied (i2, i3) lines only, y transforms on the occupied z
                                     planes only, z transforms on every column (src/52_fft_mpi_noabirule/fftw3_fftpad.finc,
                                     driven by sphereboundary's gbound, m_fftcore.F90:1253-1471)
          fftw3_fftpad.finc:105-196  r -> G: the same three passes in reverse order, pruned by the OUTPUT sphere
          m_getghc.F90:1999-2171     cwavef_double_rfft_trick_pack / _unpack (istwf_k = 2, real potential): E = C + i D on
                                     the completed sphere, H C = (F(G) + conj F(-G))/2, H D = (F(G) - conj F(-G))/(2i).
                                     (The reference applies this packing with paral_kgb = 1; its sequential FFT back ends get
                                     the same factor of two from the real-psi trick of sg_fftrisc / fftw3_fftrisc, so the
                                     packed transform is the fair stand-in for the CPU timing of bench.py.)
The result is identical to oracle/fourwf.py:fourwf(option=2) to rounding (tests/test_oracle_invariants.py); this module only
exists so that the CPU baseline of bench.py does the work the reference does, not the work of an unpruned 3-D FFT."""
from __future__ import annotations
import numpy as np
import scipy.fft as sfft
from .fourwf import _wrap, inverse_indices


def _lines_and_planes(i2, i3, n2):
    key = np.unique(i3 * n2 + i2)
    return key // n2, key % n2, np.unique(i3)          # (i3 of lines, i2 of lines, occupied z planes)


def _runs(idx):
    """sorted index list -> contiguous runs as slices (the occupied z planes of a sphere are one or two runs: views, no copies)"""
    cuts = np.flatnonzero(np.diff(idx) != 1) + 1
    return [slice(int(r[0]), int(r[-1]) + 1) for r in np.split(idx, cuts)]


def _g_to_r(box, key, zpl, workers):
    """box (nt, n3, n2, n1) holding the sphere coefficients and zeros -> psi(r), e^{+i}, unscaled.  key = i3*n2 + i2 of the lines."""
    nt, n3, n2, n1 = box.shape
    flat = box.reshape(nt, n3 * n2, n1)
    flat[:, key, :] = sfft.ifft(flat[:, key, :], axis=-1, norm="forward", workers=workers)              # x on occupied lines
    for sl in _runs(zpl):
        box[:, sl] = sfft.ifft(box[:, sl], axis=2, norm="forward", workers=workers)                     # y on occupied planes
    return sfft.ifft(box, axis=1, norm="forward", overwrite_x=True, workers=workers)                   # z on every column


def _r_to_g(box, key, zpl, workers):
    """psi(r) -> unscaled e^{-i} transform, valid on the lines `key` only; returns the (nt, nlines, n1) line block."""
    nt, n3, n2, n1 = box.shape
    box = sfft.fft(box, axis=1, norm="backward", overwrite_x=True, workers=workers)                    # z on every column
    for sl in _runs(zpl):
        box[:, sl] = sfft.fft(box[:, sl], axis=2, norm="backward", workers=workers)                     # y on the output planes
    return sfft.fft(box.reshape(nt, n3 * n2, n1)[:, key, :], axis=-1, norm="backward", workers=workers)  # x on the output lines


def _times_v(box, v, workers):
    """cg_vlocpsi, one transform per thread (NumPy's multiply is single-threaded; the reference's loop is OpenMP-parallel)"""
    if not workers or workers <= 1 or box.shape[0] == 1:
        box *= v[None]
        return
    from concurrent.futures import ThreadPoolExecutor

    def one(t):
        np.multiply(box[t], v, out=box[t])
    with ThreadPoolExecutor(max_workers=min(int(workers), box.shape[0])) as ex:
        list(ex.map(one, range(box.shape[0])))


def fourwf_option2_padded(cplex, denpot, fofgin, kg, ngfft, istwf_k=1, me_g0=1, workers=None, chunk=16):
    """fofgout(ndat, npw) = gather(FFT[V * FFT^-1[scatter(fofgin)]]) / N for identical in/out spheres (getghc's call)."""
    cg = np.atleast_2d(fofgin)
    ndat, npw = cg.shape
    n1, n2, n3 = ngfft
    i1, i2, i3 = _wrap(kg, ngfft)
    xnorm = 1.0 / float(n1 * n2 * n3)
    out = np.empty((ndat, npw), dtype=np.complex128)
    pack = (istwf_k == 2 and cplex == 1)
    if istwf_k >= 2:
        lo = 1 if (istwf_k == 2 and me_g0 == 1) else 0
        j1, j2, j3 = inverse_indices(i1[lo:], i2[lo:], i3[lo:], ngfft, istwf_k)
        a2 = np.concatenate([i2, j2]); a3 = np.concatenate([i3, j3])
    else:
        lo = 0; a2, a3 = i2, i3
    l3, l2, zpl = _lines_and_planes(a2, a3, n2)              # input lines / planes (completed sphere)
    key = l3 * n2 + l2
    if pack:
        okey, ozpl = key, zpl                                # the packed transform needs F on the completed sphere
    else:
        o3, o2, ozpl = _lines_and_planes(i2, i3, n2)         # plain gather on the stored half sphere
        okey = o3 * n2 + o2
    opos = np.searchsorted(okey, i3 * n2 + i2)               # line of every output coefficient
    if pack:
        mpos = np.searchsorted(okey, j3 * n2 + j2)           # line of -G for coefficients lo..npw-1
    step = 2 * chunk if pack else chunk
    for b0 in range(0, ndat, step):
        c = cg[b0:b0 + step]
        nb = c.shape[0]
        if pack:
            nt = (nb + 1) // 2
            cc = c[0::2]
            dd = np.zeros_like(cc); dd[:nb // 2] = c[1::2]
            box = np.zeros((nt, n3, n2, n1), dtype=np.complex128)
            if lo:
                cc = cc.copy(); dd = dd.copy()
                cc[:, 0] = cc[:, 0].real; dd[:, 0] = dd[:, 0].real            # Im c(G=0) = 0 (m_fftcore.F90:1632-1638)
            box[:, i3, i2, i1] = cc + 1j * dd
            box[:, j3, j2, j1] = np.conj(cc[:, lo:]) + 1j * np.conj(dd[:, lo:])
        else:
            nt = nb
            box = np.zeros((nt, n3, n2, n1), dtype=np.complex128)
            box[:, i3, i2, i1] = c
            if istwf_k >= 2:
                if lo:
                    box[:, 0, 0, 0] = c[:, 0].real
                box[:, j3, j2, j1] = np.conj(c[:, lo:])
        box = _g_to_r(box, key, zpl, workers)
        _times_v(box, denpot, workers)                                            # cg_vlocpsi
        lines = _r_to_g(box, okey, ozpl, workers)
        del box
        f = lines[:, opos, i1] * xnorm
        if pack:
            fm = np.empty_like(f)
            fm[:, lo:] = lines[:, mpos, j1] * xnorm
            if lo:
                fm[:, 0] = f[:, 0]
            hc = 0.5 * (f + np.conj(fm)); hd = -0.5j * (f - np.conj(fm))
            if lo:
                hc[:, 0] = f[:, 0].real; hd[:, 0] = f[:, 0].imag
            out[b0:b0 + nb:2] = hc
            out[b0 + 1:b0 + nb:2] = hd[:nb // 2]
        else:
            if istwf_k == 2 and me_g0 == 1:
                f[:, 0] = f[:, 0].real
            out[b0:b0 + nb] = f
    return out
