// Launcher templates of the fourwf plane stage (included by plane_stage.cu and the plane_inst_*.cu instantiation units).
#pragma once
#include "plane_stage.cuh"
#include "fourwf.cuh"
#include "context.cuh"
#include <algorithm>

namespace abi {
void* plane_scratch_get(size_t bytes);

// L2 budget for the S planes of all resident CTAs (B200: 126 MB L2; leave room for the W1 / V_loc streams)
constexpr size_t kScratchL2Budget = (size_t)80 << 20;
template <int R1, int R2, int G, int WARPS>
void launch_cfg(PlaneParams& P, cudaStream_t st) {
  using F = PlaneFft<R1, R2, G>;
  auto kern = k_fw_plane<R1, R2, G, WARPS>;
  const size_t smem = sizeof(double2) * ((size_t)F::TWSIZE + F::ZOFF + (size_t)WARPS * F::ESIZE);
  int cps = 1;
#ifndef ABI_EMU
  static bool attr_done = false;
  if (!attr_done) { CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; }
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, WARPS * 32, smem));
  ABI_CHECK(cps >= 1, "plane stage: kernel does not fit on an SM");
#endif
  const size_t sbytes = sizeof(double2) * (size_t)P.nU * P.n2;
  const FourwfTuning& tune = fourwf_tuning();
  // The S planes are meant to stay in L2, but losing the second CTA of an SM costs far more than S spilling to HBM does
  // (measured on B200: the split path with S entirely in HBM is 5 % slower than the fused kernel at 180^3, one CTA per SM
  // costs 17 % at 192^3: 3.28 -> 2.81 ms per 64 bands) -> the L2 budget only trims a third or fourth CTA.
  int by_l2 = (int)std::max<size_t>(2, kScratchL2Budget / (sbytes * kNumSM));
  cps = std::min(cps, by_l2);
  if (tune.plane_ctas_per_sm > 0) cps = std::min(cps, tune.plane_ctas_per_sm);
  long long grid = std::min<long long>(P.nunits, (long long)kNumSM * cps);
#ifdef ABI_EMU
  grid = std::min<long long>(grid, 3);
#endif
  P.S = (double2*)plane_scratch_get(sbytes * (size_t)grid);
  ABI_LAUNCH(kern, dim3((unsigned)grid), dim3(WARPS * 32), smem, st, P);
}

template <int R1, int R2, int G, int WARPS>
void launch_cfg_rho(PlaneParams& P, cudaStream_t st) {
  using F = PlaneFft<R1, R2, G>;
  auto kern = k_fw_plane_rho<R1, R2, G, WARPS>;
  const size_t smem = sizeof(double2) * ((size_t)F::TWSIZE + F::ZOFF + (size_t)WARPS * F::ESIZE);
  int cps = 1;
#ifndef ABI_EMU
  static bool attr_done = false;
  if (!attr_done) { CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; }
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, WARPS * 32, smem));
  ABI_CHECK(cps >= 1, "plane stage: kernel does not fit on an SM");
#endif
  const size_t sbytes = sizeof(double2) * (size_t)P.nU * P.n2;
  cps = std::min(cps, (int)std::max<size_t>(2, kScratchL2Budget / (sbytes * kNumSM)));
  long long grid = std::min<long long>(P.nunits, (long long)kNumSM * cps);
#ifdef ABI_EMU
  grid = std::min<long long>(grid, 3);
#endif
  P.S = (double2*)plane_scratch_get(sbytes * (size_t)grid);
  ABI_LAUNCH(kern, dim3((unsigned)grid), dim3(WARPS * 32), smem, st, P);
}

template <int R1, int R2>
void plane_launch_rho_n(PlaneParams& P, cudaStream_t st) { launch_cfg_rho<R1, R2, 4, 8>(P, st); }

// one kernel of the split path (n2 != n3); P.S is already set by plane_stage_launch*
template <int R1, int R2, int KIND>
void launch_split_kind(PlaneParams& P, cudaStream_t st) {
  constexpr int G = 4, WARPS = 8;
  using F = PlaneFft<R1, R2, G>;
  auto kern = k_fw_plane_split<R1, R2, G, WARPS, KIND>;
  const size_t smem = sizeof(double2) * ((size_t)F::TWSIZE + F::ZOFF + (size_t)WARPS * F::ESIZE);
  int cps = 1;
#ifndef ABI_EMU
  static bool attr_done = false;
  if (!attr_done) { CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; }
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, WARPS * 32, smem));
  ABI_CHECK(cps >= 1, "plane stage: kernel does not fit on an SM");
#endif
  long long grid = std::min<long long>(P.nunits, (long long)kNumSM * cps);
#ifdef ABI_EMU
  grid = std::min<long long>(grid, 3);
#endif
  ABI_LAUNCH(kern, dim3((unsigned)grid), dim3(WARPS * 32), smem, st, P);
}

template <int R1, int R2>
void plane_launch_split_n(int kind, PlaneParams& P, cudaStream_t st) {
  if (kind == 0) launch_split_kind<R1, R2, 0>(P, st);
  else if (kind == 1) launch_split_kind<R1, R2, 1>(P, st);
  else if (kind == 2) launch_split_kind<R1, R2, 2>(P, st);
  else launch_split_kind<R1, R2, 3>(P, st);
}

template <int R1, int R2>
void plane_launch_n(PlaneParams& P, cudaStream_t st) {
  const int cfg = fourwf_tuning().plane_cfg;
  // measured on B200 (Si-512 box, 64 bands): (4 columns x 8 warps, 2 CTAs/SM) 4.24 ms, (4 x 16, 1 CTA/SM) 4.51 ms,
  // (8 x 4, 2 CTAs/SM) 5.15 ms -- the stage is latency bound, so warps per SM win over lane efficiency
  if (cfg == 3) launch_cfg<R1, R2, 4, 16>(P, st);
  else if (cfg == 1) launch_cfg<R1, R2, 8, 4>(P, st);
  else launch_cfg<R1, R2, 4, 8>(P, st);
}


}  // namespace abi
