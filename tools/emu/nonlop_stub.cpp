// emu build only: the DMMA GEMM kernels cannot be emulated with one thread per block.
namespace abi {
void nonlop_release_all() {}
void ozaki_set_enabled(int) {}
void nonlop_set_rag(int) {}
void xg_release_workspace() {}
void chebfi_release_workspace() {}
}
