// Explicit instantiations of the half-support plane stage (split over several units to compile in parallel).
#include "half_stage_impl.cuh"
namespace abi {
template void half_launch_n<5, 3, 10>(int, HalfParams&, cudaStream_t);
template void half_launch_n<4, 4, 8>(int, HalfParams&, cudaStream_t);
template void half_launch_n<9, 2, 9>(int, HalfParams&, cudaStream_t);
}  // namespace abi
