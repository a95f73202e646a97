// Explicit instantiations of the half-support x passes (split over several units to compile in parallel).
#include "x_stage_impl.cuh"
namespace abi {
extern const int kXhGLHost = kXhGL;
template void xh_launch<9, 10>(int, XhParams&, cudaStream_t);
template void xh_launch<4, 3>(int, XhParams&, cudaStream_t);
template void xh_launch<5, 3>(int, XhParams&, cudaStream_t);
template void xh_launch<4, 4>(int, XhParams&, cudaStream_t);
}  // namespace abi
