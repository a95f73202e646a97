"""ctypes binding of the C-ABI declared in include/abinit_b200.h (the same symbols a Fortran
``iso_c_binding`` interface block binds)."""
from __future__ import annotations
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LibraryNotBuilt(RuntimeError):
    pass


def library_path() -> str:
    return os.path.join(_HERE, "libabinit_b200.so")


# every symbol include/abinit_b200.h declares (checked by tests/test_abi_symbols.py)
SYMBOLS = [
    "abi_b200_init", "abi_b200_finalize", "abi_b200_set_stream", "abi_b200_set_async", "abi_b200_synchronize",
    "abi_b200_kernel_launches", "abi_b200_version", "abi_b200_profile_enable", "abi_b200_profile_collect", "abi_b200_probe_fp64_peak",
    "abi_b200_fourwf_", "abi_b200_alloc_fourwf_", "abi_b200_free_fourwf_", "gpu_fourwf_", "alloc_gpu_fourwf_",
    "free_gpu_fourwf_", "abi_b200_set_me_g0", "abi_b200_fourwf_set_impl", "abi_b200_fourwf_set_tuning", "abi_b200_fourwf_counter",
    "abi_b200_init_gemm_nonlop_", "abi_b200_destroy_gemm_nonlop_", "abi_b200_prep_projectors_",
    "abi_b200_set_projectors_", "abi_b200_set_gemm_nonlop_ikpt_", "abi_b200_initylmg_k_", "abi_b200_mkffnl_", "abi_b200_gemm_nonlop_",
    "abi_b200_nonlop_counter",
    "abi_b200_ham_create", "abi_b200_ham_destroy", "abi_b200_ham_load_spin", "abi_b200_ham_set_nspinor", "abi_b200_ham_load_spin_nvloc", "abi_b200_ham_load_enl", "abi_b200_ham_load_enl_spinor",
    "abi_b200_ham_load_k", "abi_b200_ham_load_k_xred", "abi_b200_ham_set_projectors", "abi_b200_ham_nprojs", "abi_b200_getghc_", "abi_b200_getghc_batch_", "abi_b200_graphs_clear",
    "abi_b200_nonlop_",
    "abi_b200_xg_gram_", "abi_b200_xg_rotate_", "abi_b200_xg_hegvd_", "abi_b200_xg_gemm_nn_", "abi_b200_xg_chol_inverse_", "abi_b200_xg_colwise_", "abi_b200_xg_rayleigh_ritz_",
    "abi_b200_chebfiwf2_", "abi_b200_lobpcgwf2_", "abi_b200_chebfi_rq_", "abi_b200_chebfi_core_", "abi_b200_cheb_oracle1_", "abi_b200_cheb_poly1_",
    "abi_b200_make_invovl_", "abi_b200_apply_invovl_",
    "abi_b200_comm_get_unique_id_", "abi_b200_comm_init_rank_", "abi_b200_comm_adopt_", "abi_b200_comm_destroy_", "abi_b200_xg_transpose_",
    "abi_b200_chebfiwf2_paral_", "abi_b200_lobpcgwf2_paral_",
]


def load_library(path: str | None = None) -> C.CDLL:
    """Load libabinit_b200.so.  Raises LibraryNotBuilt (never falls back to a CPU path)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or library_path()
    if not os.path.exists(p):
        raise LibraryNotBuilt(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). abinit_b200 has no CPU fallback.")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    vp, ip, dp = C.c_void_p, C.c_void_p, C.c_void_p   # all array arguments are passed as raw addresses
    lib.abi_b200_init.argtypes = [C.c_int]
    lib.abi_b200_set_stream.argtypes = [vp]
    lib.abi_b200_set_async.argtypes = [C.c_int]
    lib.abi_b200_kernel_launches.restype = C.c_longlong
    lib.abi_b200_fourwf_counter.restype = C.c_longlong
    lib.abi_b200_profile_enable.argtypes = [C.c_int]
    lib.abi_b200_profile_collect.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    lib.abi_b200_profile_collect.restype = C.c_int
    lib.abi_b200_version.restype = C.c_char_p
    lib.abi_b200_probe_fp64_peak.argtypes = [vp, vp]
    lib.abi_b200_set_me_g0.argtypes = [C.c_int]
    lib.abi_b200_fourwf_set_impl.argtypes = [C.c_int]
    lib.abi_b200_fourwf_set_tuning.argtypes = [C.c_char_p, C.c_int]
    for name in ("abi_b200_fourwf_", "gpu_fourwf_"):
        getattr(lib, name).argtypes = [vp] * 24
    lib.abi_b200_alloc_fourwf_.argtypes = [vp] * 4
    if hasattr(lib, "abi_b200_gemm_nonlop_"):
        lib.abi_b200_nonlop_counter.restype = C.c_longlong
        lib.abi_b200_init_gemm_nonlop_.argtypes = [vp]
        lib.abi_b200_prep_projectors_.argtypes = [vp] * 12
        lib.abi_b200_set_projectors_.argtypes = [vp] * 5
        lib.abi_b200_set_gemm_nonlop_ikpt_.argtypes = [vp]
        lib.abi_b200_gemm_nonlop_.argtypes = [vp] * 28
        lib.abi_b200_mkffnl_.argtypes = [vp] * 27
        lib.abi_b200_initylmg_k_.argtypes = [vp] * 6
        lib.abi_b200_ham_create.restype = vp
        lib.abi_b200_ham_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_double]
        lib.abi_b200_ham_destroy.argtypes = [vp]
        lib.abi_b200_ham_load_spin.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.abi_b200_ham_set_nspinor.argtypes = [vp, C.c_int]
        lib.abi_b200_ham_load_spin_nvloc.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.abi_b200_ham_load_enl.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        lib.abi_b200_ham_load_enl_spinor.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
        lib.abi_b200_ham_load_k.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int]
        lib.abi_b200_ham_load_k_xred.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int]
        lib.abi_b200_ham_set_projectors.argtypes = [vp, vp, C.c_int]
        lib.abi_b200_ham_nprojs.argtypes = [vp]
        lib.abi_b200_ham_nprojs.restype = C.c_int
        lib.abi_b200_getghc_.argtypes = [vp] * 13
        lib.abi_b200_getghc_batch_.argtypes = [vp] * 9
        lib.abi_b200_xg_gram_.argtypes = [vp] * 11
        lib.abi_b200_nonlop_.argtypes = [vp] * 15
        lib.abi_b200_xg_rotate_.argtypes = [vp] * 8
        lib.abi_b200_xg_hegvd_.argtypes = [vp] * 8
        lib.abi_b200_xg_gemm_nn_.argtypes = [vp] * 11
        lib.abi_b200_xg_chol_inverse_.argtypes = [vp] * 5
        lib.abi_b200_xg_colwise_.argtypes = [vp] * 13
        lib.abi_b200_xg_rayleigh_ritz_.argtypes = [vp] * 13
        lib.abi_b200_chebfiwf2_.argtypes = [vp] * 18
        lib.abi_b200_lobpcgwf2_.argtypes = [vp] * 15
        lib.abi_b200_chebfi_rq_.argtypes = [vp] * 9
        lib.abi_b200_chebfi_core_.argtypes = [vp] * 12
        lib.abi_b200_cheb_oracle1_.argtypes = [vp] * 5
        lib.abi_b200_cheb_oracle1_.restype = C.c_int
        lib.abi_b200_cheb_poly1_.argtypes = [vp] * 4
        lib.abi_b200_cheb_poly1_.restype = C.c_double
        lib.abi_b200_make_invovl_.argtypes = [vp]
        lib.abi_b200_apply_invovl_.argtypes = [vp] * 8
        lib.abi_b200_comm_get_unique_id_.argtypes = [vp]
        lib.abi_b200_comm_init_rank_.argtypes = [vp] * 3
        lib.abi_b200_comm_adopt_.argtypes = [vp] * 3
        lib.abi_b200_xg_transpose_.argtypes = [vp] * 5
        lib.abi_b200_chebfiwf2_paral_.argtypes = [vp] * 18
        lib.abi_b200_lobpcgwf2_paral_.argtypes = [vp] * 11
    if path is None:
        _LIB = lib
    return lib
