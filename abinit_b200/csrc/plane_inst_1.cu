// Explicit instantiations of the fourwf plane stage (split over several units to compile in parallel).
#include "plane_stage_impl.cuh"
namespace abi {
template void plane_launch_n<4, 6>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<4, 6>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<4, 6>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<5, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<5, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<5, 9>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<5, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<5, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<5, 10>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<7, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<7, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<7, 12>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<8, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<8, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<8, 10>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<8, 14>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<8, 14>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<8, 14>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<9, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<9, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<9, 15>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<12, 14>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<12, 14>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<12, 14>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<15, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<15, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<15, 16>(int, PlaneParams&, cudaStream_t);
}  // namespace abi
