"""ctypes front end of the C++/OpenMP restatement of the reference's CPU getghc (oracle/cref/getghc_ref.cpp).

TEST / MEASUREMENT INFRASTRUCTURE ONLY, like the rest of ``oracle/``: imported by ``tests/`` and by the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  The shared object is built with g++ -fopenmp into ``oracle/_build/`` (git-ignored,
travels to the GPU box) by ``build()``; DGEMM is the host OpenBLAS that SciPy bundles, handed to the C++ side as a function
pointer taken from ``scipy.linalg.cython_blas``."""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "getghc_ref.cpp")
_OUT_DIR = os.path.join(os.path.dirname(_HERE), "_build")
_OUT = os.path.join(_OUT_DIR, "libgetghc_ref.so")
_LIB = None
# AVX2 + FMA: every x86-64 server CPU of the last decade; -march=native would tie the object to the build container's CPU
CXXFLAGS = ["-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-march=x86-64-v3", "-ffp-contract=fast", "-fno-math-errno"]


def _compiler() -> str:
    # the distribution g++ carries libgomp; a CXX from the environment may not (the image's /opt/gcc does not)
    for c in ("/usr/bin/g++", os.environ.get("CXX", ""), "g++"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "g++"


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        subprocess.check_call([_compiler()] + CXXFLAGS + ["-o", _OUT, _SRC])
    return _OUT


def _dgemm_ptr() -> int:
    from scipy.linalg import cython_blas
    cap = cython_blas.__pyx_capi__["dgemm"]
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap))


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        vp = C.c_void_p
        _LIB.cref_fft1d.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        _LIB.cref_fourwf_opt2.argtypes = [C.c_int] * 6 + [vp] * 4
        _LIB.cref_fourwf_opt2.restype = C.c_int
        _LIB.cref_getghc.argtypes = ([C.c_int] * 6 + [vp] * 4 + [C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int,
                                                               vp, vp, vp, vp])
        _LIB.cref_getghc.restype = C.c_int
        _LIB.cref_max_threads.restype = C.c_int
        _LIB.cref_set_threads.argtypes = [C.c_int]
    return _LIB


def set_threads(n: int) -> None:
    """OpenMP threads of the FFT passes and OpenBLAS threads of the DGEMMs."""
    lib().cref_set_threads(int(n))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=int(n), user_api="blas")
    except Exception:
        pass


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def fft1d(x, sign):
    """x (howmany, n) complex -> unscaled transform with e^{sign i}."""
    a = np.ascontiguousarray(x, dtype=np.complex128).copy()
    lib().cref_fft1d(_p(a), a.shape[1], a.shape[0], int(sign))
    return a


def fourwf_option2(vlocal, cwavef, kgF, ngfft, istwf_k=1):
    """kgF: (npw, 3) int32 (Fortran kg_k(3, npw) memory)."""
    cw = np.ascontiguousarray(np.atleast_2d(cwavef), dtype=np.complex128)
    kg = np.ascontiguousarray(kgF, dtype=np.int32)
    v = np.ascontiguousarray(vlocal, dtype=np.float64)
    out = np.empty_like(cw)
    n1, n2, n3 = ngfft
    rc = lib().cref_fourwf_opt2(cw.shape[0], cw.shape[1], int(istwf_k), n1, n2, n3, _p(kg), _p(cw), _p(v), _p(out))
    assert rc == 0, "cref_fourwf_opt2: unsupported istwf_k"
    return out


class Operator:
    """The arrays of one k-point in the layout the C++ side wants (built once, like gemm_nonlop's prep_projectors)."""

    def __init__(self, vlocal, kgF, ngfft, kinpw, Pr, Pi, istwf_k, ekb_proj=None, blk_nlmn=None, dij=None, sij=None):
        self.v = np.ascontiguousarray(vlocal, dtype=np.float64)
        self.kg = np.ascontiguousarray(kgF, dtype=np.int32)
        self.ngfft = tuple(int(n) for n in ngfft)
        self.kin = np.ascontiguousarray(kinpw, dtype=np.float64)
        self.Pr = np.ascontiguousarray(Pr, dtype=np.float64); self.Pi = np.ascontiguousarray(Pi, dtype=np.float64)
        self.istwf_k = int(istwf_k)
        self.paw = 0 if ekb_proj is not None else 1
        self.ekb = None if ekb_proj is None else np.ascontiguousarray(ekb_proj, dtype=np.float64)
        self.blk = None if blk_nlmn is None else np.ascontiguousarray(blk_nlmn, dtype=np.int32)
        self.dij = None if dij is None else np.ascontiguousarray(dij, dtype=np.float64)
        self.sij = None if sij is None else np.ascontiguousarray(sij, dtype=np.float64)
        self.dgemm = _dgemm_ptr()
        self.timings = np.zeros(4)

    @classmethod
    def from_oracle_arrays(cls, vlocal, kgF, ngfft, kinpw, P, enl, sij, indlmn, nattyp, atindx1, istwf_k, usepaw):
        """Same inputs as oracle.getghc.getghc (P complex (nprojs, npw), enl / sij as described in oracle.nonlop.opernlc)."""
        from ..nonlop import nlmn_of_types
        nl = nlmn_of_types(indlmn)
        if not usepaw:
            ekb = np.concatenate([np.tile(np.asarray(enl)[t, np.asarray(indlmn)[t, :nl[t], 4].astype(int) - 1], int(nattyp[t]))
                                  for t in range(len(nl))])
            return cls(vlocal, kgF, ngfft, kinpw, P.real, P.imag, istwf_k, ekb_proj=ekb)
        blk, d, s = [], [], []
        iatm = 0
        for t in range(len(nl)):
            for ia in range(int(nattyp[t])):
                blk.append(nl[t]); d.append(np.asarray(enl)[atindx1[iatm + ia]])
                if sij is not None:
                    s.append(np.asarray(sij)[t])
            iatm += int(nattyp[t])
        return cls(vlocal, kgF, ngfft, kinpw, P.real, P.imag, istwf_k, blk_nlmn=blk, dij=np.stack(d), sij=np.stack(s) if s else None)

    def getghc(self, cwavef, sij_opt=0):
        cw = np.ascontiguousarray(np.atleast_2d(cwavef), dtype=np.complex128)
        ndat, npw = cw.shape
        ghc = np.empty_like(cw)
        gsc = np.empty_like(cw) if sij_opt == 1 else None
        n1, n2, n3 = self.ngfft
        nprojs = self.Pr.shape[0]
        dim1 = 0 if self.dij is None else self.dij.shape[1]
        rc = lib().cref_getghc(ndat, npw, self.istwf_k, n1, n2, n3, _p(self.kg), _p(cw), _p(self.v), _p(self.kin), nprojs,
                               _p(self.Pr), _p(self.Pi), self.paw, _p(self.ekb), 0 if self.blk is None else len(self.blk),
                               _p(self.blk), _p(self.dij), _p(self.sij), dim1, int(sij_opt), _p(ghc), _p(gsc), C.c_void_p(self.dgemm),
                               _p(self.timings))
        assert rc == 0, "cref_getghc: unsupported istwf_k"
        return ghc, gsc
