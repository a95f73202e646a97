// Explicit instantiations of the half-support x passes (split over several units to compile in parallel).
#include "x_stage_impl.cuh"
namespace abi {
template void xh_launch<9, 2>(int, XhParams&, cudaStream_t);
template void xh_launch<4, 5>(int, XhParams&, cudaStream_t);
template void xh_launch<8, 3>(int, XhParams&, cudaStream_t);
template void xh_launch<4, 8>(int, XhParams&, cudaStream_t);
template void xh_launch<8, 8>(int, XhParams&, cudaStream_t);
}  // namespace abi
