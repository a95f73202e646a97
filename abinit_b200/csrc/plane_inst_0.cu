// Explicit instantiations of the fourwf plane stage (split over several units to compile in parallel).
#include "plane_stage_impl.cuh"
namespace abi {
template void plane_launch_n<4, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<4, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<4, 8>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<6, 6>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<6, 6>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<6, 6>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<7, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<7, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<7, 8>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<8, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<8, 9>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<8, 9>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<8, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<8, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<8, 16>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<9, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<9, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<9, 10>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<9, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<9, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<9, 12>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<10, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<10, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<10, 16>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<16, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<16, 16>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<16, 16>(int, PlaneParams&, cudaStream_t);
}  // namespace abi
