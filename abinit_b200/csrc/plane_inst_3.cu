// Explicit instantiations of the fourwf plane stage (split over several units to compile in parallel).
#include "plane_stage_impl.cuh"
namespace abi {
template void plane_launch_n<5, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<5, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<5, 15>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<6, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<6, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<6, 8>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<6, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<6, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<6, 9>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<9, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<9, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<9, 9>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<10, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<10, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<10, 12>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<10, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<10, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<10, 15>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<12, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<12, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<12, 16>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<14, 14>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<14, 14>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<14, 14>(int, PlaneParams&, cudaStream_t);
}  // namespace abi
