"""abinit_b200: B200-native (sm_100a) getghc hot path -- fourwf + gemm_nonlop + kinetic assembly.

Host-side mirror of the reference interfaces (``fourwf``, ``gemm_nonlop``/``nonlop``, ``getghc`` and the
``gs_hamiltonian_type`` life cycle) above the C-ABI library ``libabinit_b200.so`` built from ``csrc/``.
There is no CPU fallback: importing the compute entry points without the CUDA library raises.
"""
from .lib import load_library, library_path, LibraryNotBuilt  # noqa: F401
from .api import (fourwf, gemm_nonlop, getghc, Hamiltonian, init, finalize, synchronize,  # noqa: F401
                  kernel_launches, set_stream, set_async, nonlop, make_invovl, apply_invovl)
from . import xg  # noqa: F401
from .xg import chebfiwf2, lobpcgwf2, xg_RayleighRitz  # noqa: F401

__all__ = ["fourwf", "gemm_nonlop", "getghc", "Hamiltonian", "init", "finalize", "synchronize",
           "kernel_launches", "set_stream", "set_async", "nonlop", "make_invovl", "apply_invovl", "xg", "chebfiwf2", "lobpcgwf2", "xg_RayleighRitz", "load_library", "library_path", "LibraryNotBuilt"]
