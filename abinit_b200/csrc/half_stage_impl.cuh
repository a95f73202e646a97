// Launcher templates of the half-support plane stage (included by the half_inst_*.cu instantiation units).
#pragma once
#include "half_stage.cuh"
#include "fourwf.cuh"
#include "context.cuh"
#include <algorithm>

namespace abi {
void* half_scratch_get(size_t bytes, uint64_t layout_key, cudaStream_t st);

// kind 0: fused option 2, kind 1: fused option 1 (density)
template <int A, int B, int G, int WARPS, int MINB, int KIND>
void half_launch_cfg(HalfParams& P, cudaStream_t st) {
  auto kern = k_hw_plane<A, B, G, WARPS, MINB, KIND>;
  const size_t smem = half_smem_bytes<A, B, G>(WARPS, P.nU);
  ABI_CHECK(smem <= kMaxSmemPerCta, "half-support plane stage: too many occupied z planes for the shared-memory tables");
  int cps = 1;
#ifndef ABI_EMU
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, WARPS * 32, smem));
  ABI_CHECK(cps >= 1, "half-support plane stage: kernel does not fit on an SM");
#endif
  const FourwfTuning& tune = fourwf_tuning();
  cps = std::min(cps, MINB);
  if (tune.plane_ctas_per_sm > 0) cps = std::min(cps, tune.plane_ctas_per_sm);
  long long grid = std::min<long long>(P.nunits, (long long)kNumSM * cps);
#ifdef ABI_EMU
  grid = std::min<long long>(grid, 3);
#endif
  const size_t sbytes = sizeof(double2) * (size_t)P.ng2 * HalfFft<A, B, G>::GSTR;
  // rows without a plane must read as zero: the scratch is cleared whenever the plan (layout key) changes
  P.S = (double2*)half_scratch_get(sbytes * (size_t)grid, P.layout_key, st);
  ABI_LAUNCH(kern, dim3((unsigned)grid), dim3(WARPS * 32), smem, st, P);
}

template <int A, int B, int G>
void half_launch_n(int kind, HalfParams& P, cudaStream_t st) {
  using F = HalfFft<A, B, G>;
  static_assert(F::ROWS * G * 16 + F::ESIZE * 16 <= 14 * 1024, "per-warp shared memory too large for 16 warps per SM");
  const int cfg = fourwf_tuning().half_cfg;
  // warps that can be busy at once in a phase: line batches (y) or column batches (z) of ONE plane.  Small boxes have only a
  // few, so they run as small CTAs, several per SM (many planes in flight), instead of one 16-warp CTA per SM.
  constexpr int kBusy = F::NBY > F::NG ? F::NBY : F::NG;
  if (kBusy <= 4) {
    if (kind == 0) half_launch_cfg<A, B, G, 4, 6, 0>(P, st);
    else half_launch_cfg<A, B, G, 4, 6, 1>(P, st);
  } else if (kBusy <= 8 || cfg == 0 || kind != 0) {
    if (kind == 0) half_launch_cfg<A, B, G, 8, 2, 0>(P, st);
    else half_launch_cfg<A, B, G, 8, 2, 1>(P, st);
  } else {
    half_launch_cfg<A, B, G, 16, 1, 0>(P, st);
  }
}

}  // namespace abi
