"""Host-side mirror of the reference interfaces for the getghc hot path.

Function names, argument names, argument meaning and error behaviour follow the Fortran routines they stand
for, so that tests read like the reference's own call sites:

  fourwf        src/53_ffts/m_fft.F90:2290      (same positional argument list)
  gemm_nonlop   src/66_nonlocal/m_gemm_nonlop.F90:191 (arguments read by choice in {0,1,7}, signs=2)
  getghc        src/66_wfs/m_getghc.F90:182     (gs_ham -> Hamiltonian handle)
  Hamiltonian   src/66_nonlocal/m_hamiltonian.F90:99-467  init / load_spin / load_k life cycle

Arrays are passed exactly as Fortran lays them out in memory: complex data as float64 ``(..., 2)`` or
complex128 arrays, ``kg_k(3,npw)`` as a C-contiguous int32 array of shape ``(npw, 3)``, ``vlocal(n4,n5,n6)``
as shape ``(n6, n5, n4)``.  Each array may be a NumPy array (host) or a CUDA ``torch.Tensor`` (device): device
arrays are used in place, host arrays are staged by the library (include/abinit_b200.h, "Pointer residency").
Output arguments are written in place, as in Fortran.  Everything runs in libabinit_b200.so; there is no CPU
path here.
"""
from __future__ import annotations
import ctypes as C
import numpy as np
from . import lib as _lib

_L = None


def _use_library(handle):
    """Developer hook (tools/emu only): bind the API to an explicitly loaded library object."""
    global _L
    _L = handle


def L():
    global _L
    if _L is None:
        _L = _lib.load_library()
    return _L


def _ptr(a, dtype=None, name="array"):
    """Raw address of a NumPy array or CUDA torch tensor (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{name} must be contiguous (Fortran memory layout, see abinit_b200.api docstring)")
        if dtype is not None and a.dtype not in dtype:
            raise TypeError(f"{name} must have dtype in {dtype}, got {a.dtype}")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        return a.data_ptr()
    raise TypeError(f"{name}: expected numpy.ndarray or torch.Tensor, got {type(a)}")


_F = (np.dtype(np.float64), np.dtype(np.complex128))
_I = (np.dtype(np.int32),)


def _iref(v):
    return C.byref(C.c_int(int(v)))


def _dref(v):
    return C.byref(C.c_double(float(v)))


def init(rank: int = 0):
    L().abi_b200_init(int(rank))


def finalize():
    L().abi_b200_finalize()


def synchronize():
    L().abi_b200_synchronize()


def set_stream(cuda_stream_ptr):
    L().abi_b200_set_stream(cuda_stream_ptr)


def set_async(flag: bool):
    L().abi_b200_set_async(1 if flag else 0)


def kernel_launches() -> int:
    return int(L().abi_b200_kernel_launches())


def set_tuning(name: str, value: int):
    """Tuning knobs of the fused fourwf path (include/abinit_b200.h: abi_b200_fourwf_set_tuning)."""
    L().abi_b200_fourwf_set_tuning(name.encode(), int(value))


def profile_enable(on: bool):
    L().abi_b200_profile_enable(1 if on else 0)


def probe_fp64_peak() -> dict:
    """FP64 pipe peak of the current device measured now: {"dfma": TFLOP/s, "dmma": TFLOP/s} (abi_b200_probe_fp64_peak)."""
    a = C.c_double(); b = C.c_double()
    L().abi_b200_probe_fp64_peak(C.byref(a), C.byref(b))
    return {"dfma": a.value, "dmma": b.value}


def profile_collect() -> dict:
    """{kernel class: (total ms, launches)} since profile_enable(True)."""
    names = C.create_string_buffer(4096)
    ms = (C.c_double * 64)(); cnt = (C.c_longlong * 64)()
    n = L().abi_b200_profile_collect(names, 4096, ms, cnt, 64)
    keys = [k for k in names.value.decode().split(";") if k]
    return {keys[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}


def fourwf(cplex, denpot, fofgin, fofgout, fofr, gboundin, gboundout, istwf_k, kg_kin, kg_kout, mgfft, mpi_enreg,
           ndat, ngfft, npwin, npwout, n4, n5, n6, option, tim_fourwf=0, weight_r=1.0, weight_i=1.0,
           weight_array_r=None, weight_array_i=None, me_g0=1, impl=0):
    """fourwf (src/53_ffts/m_fft.F90:2290-2310 argument list).  ``impl``: 0 auto, 1 generic full-box kernels,
    2 fused zero-padded kernels (option 2 only)."""
    lib = L()
    ngfft_a = np.zeros(18, dtype=np.int32)
    ngfft_a[:len(ngfft)] = np.asarray(ngfft, dtype=np.int32)[:18]
    wr = np.ascontiguousarray(np.broadcast_to(np.asarray(weight_r if weight_array_r is None else weight_array_r,
                                                         dtype=np.float64), (ndat,)))
    wi = np.ascontiguousarray(np.broadcast_to(np.asarray(weight_i if weight_array_i is None else weight_array_i,
                                                         dtype=np.float64), (ndat,)))
    lib.abi_b200_set_me_g0(int(me_g0))
    lib.abi_b200_fourwf_set_impl(int(impl))
    lib.abi_b200_fourwf_(_iref(cplex), _ptr(denpot, _F, "denpot"), _ptr(fofgin, _F, "fofgin"),
                         _ptr(fofgout, _F, "fofgout"), _ptr(fofr, _F, "fofr"), _ptr(gboundin, _I, "gboundin"),
                         _ptr(gboundout, _I, "gboundout"), _iref(istwf_k), _ptr(kg_kin, _I, "kg_kin"),
                         _ptr(kg_kout, _I, "kg_kout"), _iref(mgfft), None, _iref(ndat), ngfft_a.ctypes.data,
                         _iref(npwin), _iref(npwout), _iref(n4), _iref(n5), _iref(n6), _iref(option), _iref(0),
                         _iref(tim_fourwf), wr.ctypes.data, wi.ctypes.data)
    lib.abi_b200_fourwf_set_impl(0)


class Hamiltonian:
    """gs_hamiltonian_type life cycle (m_hamiltonian.F90): init -> load_spin -> load_k -> getghc."""

    def __init__(self, ngfft, natom, ntypat, lmnmax, indlmn, nattyp, atindx1, usepaw, ucvol):
        lib = L()
        self._ngfft = np.zeros(18, dtype=np.int32); self._ngfft[:3] = np.asarray(ngfft[:3], dtype=np.int32)
        self._ngfft[3:6] = self._ngfft[:3]
        self.indlmn = np.ascontiguousarray(indlmn, dtype=np.int32)     # (ntypat, lmnmax, 6) == Fortran (6,lmnmax,ntypat)
        self.nattyp = np.ascontiguousarray(nattyp, dtype=np.int32)
        self.atindx1 = np.ascontiguousarray(atindx1, dtype=np.int32)   # 1-based, as in Fortran
        self.usepaw = int(usepaw)
        self.h = lib.abi_b200_ham_create(self._ngfft.ctypes.data, int(natom), int(ntypat), int(lmnmax),
                                         self.indlmn.ctypes.data, self.nattyp.ctypes.data, self.atindx1.ctypes.data,
                                         int(usepaw), float(ucvol))
        self.npw = 0; self.istwf_k = 1; self.me_g0 = 1

    def load_spin(self, vlocal, cplex=1):
        n1, n2, n3 = (int(x) for x in self._ngfft[:3])
        L().abi_b200_ham_load_spin(self.h, _ptr(vlocal, _F, "vlocal"), int(cplex), n1, n2, n3)

    def set_nspinor(self, nspinor):
        L().abi_b200_ham_set_nspinor(self.h, int(nspinor))

    def load_spin_nvloc(self, vlocal, nvloc):
        """vlocal(n4,n5,n6,nvloc) == shape (nvloc, n6, n5, n4); nvloc = 4: [V11, V22, Re V12, Im V12]."""
        n1, n2, n3 = (int(x) for x in self._ngfft[:3])
        L().abi_b200_ham_load_spin_nvloc(self.h, _ptr(vlocal, _F, "vlocal"), int(nvloc), n1, n2, n3)

    def load_enl(self, enl, sij=None):
        """enl (dimenl2, dimenl1) == Fortran enl(dimenl1, dimenl2), or (nspinortot**2, dimenl2, dimenl1) for the four spinor blocks
        [up-up, dn-dn, up-dn, dn-up] of a PAW spinor Hamiltonian; PAW dimenl1 = cplex_dij * lmn2 ((re, im) pairs when complex)."""
        enl = np.ascontiguousarray(enl, dtype=np.float64)
        sij_a = None if sij is None else np.ascontiguousarray(sij, dtype=np.float64)
        nblk = 1 if enl.ndim == 2 else int(enl.shape[0])
        L().abi_b200_ham_load_enl_spinor(self.h, enl.ctypes.data, int(enl.shape[-1]), int(enl.shape[-2]), nblk,
                                         None if sij_a is None else sij_a.ctypes.data)

    def load_k(self, istwf_k, kg_k, kinpw, ffnl=None, ph3d=None, me_g0=1):
        kg_k = np.ascontiguousarray(kg_k, dtype=np.int32)
        self.npw = int(kg_k.shape[0]); self.istwf_k = int(istwf_k); self.me_g0 = int(me_g0)
        kin = np.ascontiguousarray(kinpw, dtype=np.float64)
        dimffnl = 0 if ffnl is None else int(ffnl.shape[-2])           # (ntypat, lmnmax, dimffnl, npw)
        matblk = 0 if ph3d is None else int(ph3d.shape[0])             # (matblk, npw) complex / (matblk, npw, 2)
        L().abi_b200_ham_load_k(self.h, int(istwf_k), self.npw, kg_k.ctypes.data, kin.ctypes.data,
                                _ptr(ffnl, _F, "ffnl"), dimffnl, _ptr(ph3d, _F, "ph3d"), matblk, int(me_g0))

    def load_k_xred(self, istwf_k, kg_k, kinpw, ffnl, kpt, xred, me_g0=1):
        """load_k with ph3d built on the device from kpt(3) and xred (natom, 3) == Fortran xred(3,natom), type-sorted."""
        kg_k = np.ascontiguousarray(kg_k, dtype=np.int32)
        self.npw = int(kg_k.shape[0]); self.istwf_k = int(istwf_k); self.me_g0 = int(me_g0)
        kin = np.ascontiguousarray(kinpw, dtype=np.float64)
        kp = np.ascontiguousarray(kpt, dtype=np.float64); xr = np.ascontiguousarray(xred, dtype=np.float64)
        L().abi_b200_ham_load_k_xred(self.h, int(istwf_k), self.npw, kg_k.ctypes.data, kin.ctypes.data, _ptr(ffnl, _F, "ffnl"),
                                     int(ffnl.shape[-2]), kp.ctypes.data, xr.ctypes.data, int(me_g0))

    def set_projectors(self, projs, nprojs):
        L().abi_b200_ham_set_projectors(self.h, _ptr(projs, _F, "projs"), int(nprojs))

    @property
    def nprojs(self):
        return int(L().abi_b200_ham_nprojs(self.h))

    def destroy(self):
        if self.h:
            L().abi_b200_ham_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def getghc(cpopt, cwavef, cwaveprj, ghc, gsc, gs_ham: Hamiltonian, gvnlxc, lambda_, mpi_enreg, ndat, prtvol=0,
           sij_opt=0, tim_getghc=0, type_calc=0):
    """getghc (src/66_wfs/m_getghc.F90:182-202 argument list; cwaveprj is the flattened projections buffer)."""
    lam = np.ascontiguousarray(np.broadcast_to(np.asarray(0.0 if lambda_ is None else lambda_, dtype=np.float64), (ndat,)))
    hp = C.c_void_p(gs_ham.h)
    L().abi_b200_getghc_(_iref(cpopt), _ptr(cwavef, _F, "cwavef"), _ptr(cwaveprj, _F, "cwaveprj"), _ptr(ghc, _F, "ghc"),
                         _ptr(gsc, _F, "gsc"), C.byref(hp), _ptr(gvnlxc, _F, "gvnlxc"), lam.ctypes.data, _iref(ndat),
                         _iref(prtvol), _iref(sij_opt), _iref(tim_getghc), _iref(type_calc))


def getghc_batch(hams, cwavefs, ghcs, gscs=None, ndat=1, sij_opt=0, type_calc=0, use_graphs=True):
    """nk independent getghc calls on nk handles (one per (k, spin) pair) with device-resident blocks, overlapped on
    concurrent streams and replayed from CUDA graphs (abi_b200_getghc_batch_; the (k, spin) loop of m_vtorho.F90:789-1045).
    Identical results to a loop over getghc; gscs may be None (sij_opt = 0)."""
    nk = len(hams)
    Hp = (C.c_void_p * nk)(*[h.h for h in hams])
    Cp = (C.c_void_p * nk)(*[_ptr(c, _F, "cwavef") for c in cwavefs])
    Gp = (C.c_void_p * nk)(*[_ptr(g, _F, "ghc") for g in ghcs])
    Sp = (C.c_void_p * nk)(*[_ptr(g, _F, "gsc") for g in gscs]) if gscs is not None else None
    L().abi_b200_getghc_batch_(_iref(nk), Hp, Cp, Gp, Sp, _iref(ndat), _iref(sij_opt), _iref(type_calc), _iref(1 if use_graphs else 0))


def graphs_clear():
    L().abi_b200_graphs_clear()


_nonlop_slot_npw = {}


def prep_projectors(ikpt, npw, indlmn, nattyp, istwf_k, ucvol, ffnl, ph3d):
    """prep_projectors (m_gemm_nonlop_projectors.F90:792): builds P on the device for k-point slot ikpt (1-based)."""
    indlmn = np.ascontiguousarray(indlmn, dtype=np.int32)
    nattyp = np.ascontiguousarray(nattyp, dtype=np.int32)
    ntypat, lmnmax = indlmn.shape[0], indlmn.shape[1]
    dimffnl = int(ffnl.shape[-2]); matblk = int(ph3d.shape[0])
    L().abi_b200_prep_projectors_(_iref(ikpt), _iref(npw), _iref(lmnmax), _iref(ntypat), indlmn.ctypes.data,
                                  nattyp.ctypes.data, _iref(istwf_k), _dref(ucvol), _ptr(ffnl, _F, "ffnl"),
                                  _ptr(ph3d, _F, "ph3d"), _iref(dimffnl), _iref(matblk))


def set_projectors(ikpt, npw, nprojs, istwf_k, projs):
    L().abi_b200_set_projectors_(_iref(ikpt), _iref(npw), _iref(nprojs), _iref(istwf_k), _ptr(projs, _F, "projs"))


def set_gemm_nonlop_ikpt(ikpt):
    L().abi_b200_set_gemm_nonlop_ikpt_(_iref(ikpt))


def gemm_nonlop(atindx1, choice, cpopt, vectproj, enl, indlmn, istwf_k, lambda_, natom, nattyp, ndat, npwin, npwout,
                nspinor, ntypat, paw_opt, sij, svectout, vectin, vectout, signs=2, nnlout=1, useylm=1, dimekbq=1):
    """gemm_nonlop (src/66_nonlocal/m_gemm_nonlop.F90:191) for the current k-point slot (set_gemm_nonlop_ikpt)."""
    indlmn = np.ascontiguousarray(indlmn, dtype=np.int32)
    nattyp = np.ascontiguousarray(nattyp, dtype=np.int32)
    atindx1 = np.ascontiguousarray(atindx1, dtype=np.int32)
    enl = np.ascontiguousarray(enl, dtype=np.float64)
    sij_a = None if sij is None else np.ascontiguousarray(sij, dtype=np.float64)
    lam = np.ascontiguousarray(np.broadcast_to(np.asarray(0.0 if lambda_ is None else lambda_, dtype=np.float64), (ndat,)))
    lmnmax = indlmn.shape[1]
    L().abi_b200_gemm_nonlop_(atindx1.ctypes.data, _iref(choice), _iref(cpopt), _ptr(vectproj, _F, "vectproj"),
                              _iref(enl.shape[-1]), _iref(enl.shape[-2]), _iref(dimekbq), enl.ctypes.data,
                              indlmn.ctypes.data, _iref(istwf_k), lam.ctypes.data, _iref(lmnmax), _iref(natom),
                              nattyp.ctypes.data, _iref(ndat), _iref(nnlout), _iref(npwin), _iref(npwout),
                              _iref(nspinor), _iref(nspinor), _iref(ntypat), _iref(paw_opt),
                              None if sij_a is None else sij_a.ctypes.data, _ptr(svectout, _F, "svectout"),
                              _iref(useylm), _ptr(vectin, _F, "vectin"), _ptr(vectout, _F, "vectout"), _iref(signs))


def nonlop(choice, cpopt, cprjin, enlout, hamk: Hamiltonian, idir, lambda_, mpi_enreg, ndat, nnlout, paw_opt, signs,
           svectout, tim_nonlop, vectin, vectout):
    """nonlop (src/66_nonlocal/m_nonlop.F90:336 argument list) on the gemm_nonlop route; signs=1, choice=1 fills
    enlout(ndat) with <psi|Vnl|psi> (the call of m_chebfiwf.F90:296)."""
    lam = np.ascontiguousarray(np.broadcast_to(np.asarray(0.0 if lambda_ is None else lambda_, dtype=np.float64), (ndat,)))
    hp = C.c_void_p(hamk.h)
    L().abi_b200_nonlop_(_iref(choice), _iref(cpopt), _ptr(cprjin, _F, "cprjin"), _ptr(enlout, _F, "enlout"), C.byref(hp),
                         _iref(idir), lam.ctypes.data, _iref(ndat), _iref(nnlout), _iref(paw_opt), _iref(signs),
                         _ptr(svectout, _F, "svectout"), _iref(tim_nonlop), _ptr(vectin, _F, "vectin"),
                         _ptr(vectout, _F, "vectout"))


def make_invovl(ham: Hamiltonian):
    """make_invovl (src/66_wfs/m_invovl.F90:469) for the k-point loaded in ham."""
    hp = C.c_void_p(ham.h)
    L().abi_b200_make_invovl_(C.byref(hp))


def apply_invovl(ham: Hamiltonian, cwavef, sm1cwavef, cwaveprj, npw, ndat, mpi_enreg=None, nspinor=1, block_sliced=0):
    """apply_invovl (src/66_wfs/m_invovl.F90:790 argument list): sm1cwavef = S^-1 cwavef (PAW)."""
    hp = C.c_void_p(ham.h)
    L().abi_b200_apply_invovl_(C.byref(hp), _ptr(cwavef, _F, "cwavef"), _ptr(sm1cwavef, _F, "sm1cwavef"),
                               _ptr(cwaveprj, _F, "cwaveprj"), _iref(npw), _iref(ndat), _iref(nspinor), _iref(block_sliced))


def mkffnl(dimekb, dimffnl, ekb, ffnl, ffspl, gmet, gprimd, ider, idir, indlmn, kg, kpg, kpt, lmnmax, lnmax, mpsang, mqgrid, nkpg, npw,
           ntypat, pspso, qgrid, rmet, usepaw, useylm, ylm, ylm_gr=None):
    """mkffnl (src/66_nonlocal/m_mkffnl.F90:238 argument list), ider=0 / idir=0 / dimffnl=1 / useylm=1, on the device.
    ffnl (out): (ntypat, lmnmax, 1, npw); ffspl: (ntypat, lnmax, 2, mqgrid); ylm: (mpsang**2, npw); kg: (npw, 3) int32."""
    ind = np.ascontiguousarray(indlmn, dtype=np.int32); qg = np.ascontiguousarray(qgrid, dtype=np.float64)
    kp = np.ascontiguousarray(kpt, dtype=np.float64); gp = np.ascontiguousarray(np.asarray(gprimd, dtype=np.float64).T)  # Fortran order
    ek = None if ekb is None else np.ascontiguousarray(ekb, dtype=np.float64)
    so = np.zeros(int(ntypat), dtype=np.int32) if pspso is None else np.ascontiguousarray(pspso, dtype=np.int32)
    L().abi_b200_mkffnl_(_iref(dimekb), _iref(dimffnl), None if ek is None else ek.ctypes.data, _ptr(ffnl, _F, "ffnl"),
                         _ptr(ffspl, _F, "ffspl"), None, gp.ctypes.data, _iref(ider), _iref(idir), ind.ctypes.data, _ptr(kg, _I, "kg"), None,
                         kp.ctypes.data, _iref(lmnmax), _iref(lnmax), _iref(mpsang), _iref(mqgrid), _iref(nkpg), _iref(npw), _iref(ntypat),
                         so.ctypes.data, qg.ctypes.data, None, _iref(usepaw), _iref(useylm), _ptr(ylm, _F, "ylm"), None)


def initylmg_k(gprimd, kg, kpt, mpsang, npw, ylm):
    """initylmg (src/56_recipspace/m_initylmg.F90:94) for one k-point, optder=0: ylm (out) of shape (mpsang**2, npw)."""
    gp = np.ascontiguousarray(np.asarray(gprimd, dtype=np.float64).T); kp = np.ascontiguousarray(kpt, dtype=np.float64)
    L().abi_b200_initylmg_k_(gp.ctypes.data, _ptr(kg, _I, "kg"), kp.ctypes.data, _iref(mpsang), _iref(npw), _ptr(ylm, _F, "ylm"))
