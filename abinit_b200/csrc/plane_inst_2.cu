// Explicit instantiations of the fourwf plane stage (split over several units to compile in parallel).
#include "plane_stage_impl.cuh"
namespace abi {
template void plane_launch_n<5, 6>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<5, 6>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<5, 6>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<5, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<5, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<5, 8>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<6, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<6, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<6, 10>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<8, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<8, 8>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<8, 8>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<8, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<8, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<8, 12>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<10, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<10, 10>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<10, 10>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<12, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<12, 12>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<12, 12>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<12, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<12, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<12, 15>(int, PlaneParams&, cudaStream_t);
template void plane_launch_n<15, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_rho_n<15, 15>(PlaneParams&, cudaStream_t);
template void plane_launch_split_n<15, 15>(int, PlaneParams&, cudaStream_t);
}  // namespace abi
