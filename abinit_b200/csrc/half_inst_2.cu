// Explicit instantiations of the half-support plane stage (split over several units to compile in parallel).
#include "half_stage_impl.cuh"
namespace abi {
template void half_launch_n<4, 5, 8>(int, HalfParams&, cudaStream_t);
template void half_launch_n<8, 3, 8>(int, HalfParams&, cudaStream_t);
template void half_launch_n<4, 8, 4>(int, HalfParams&, cudaStream_t);
template void half_launch_n<8, 8, 4>(int, HalfParams&, cudaStream_t);
}  // namespace abi
