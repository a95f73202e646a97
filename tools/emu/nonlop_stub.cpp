// emu build only: the DMMA GEMM kernels cannot be emulated with one thread per block.
namespace abi { void nonlop_release_all() {} }
