// Launcher template of the half-support x passes (included by the x_inst_*.cu instantiation units).
#pragma once
#include "x_stage.cuh"
#include "context.cuh"
#include <algorithm>

namespace abi {

constexpr int kXhGL = 8;        // lines per warp batch: runs of 8 * 16 = 128 bytes at every i1 of W1 / W1o (4-line batches: 12 warps per SM but 64-byte runs, K3 0.66 -> 0.93 ms)

template <int A, int B, int DIR>
void xh_launch_dir(XhParams& P, cudaStream_t st) {
  constexpr int GL = kXhGL;
  // warps per CTA: as many as two CTAs of an SM can hold
  constexpr int WARPS = (int)((110 * 1024 - 16 * (XHalf<A, B, GL>::TW_SLOTS + XHalf<A, B, GL>::int_slots())) / (16 * XHalf<A, B, GL>::WSIZE)) < 1
                            ? 1 : ((int)((110 * 1024 - 16 * (XHalf<A, B, GL>::TW_SLOTS + XHalf<A, B, GL>::int_slots())) / (16 * XHalf<A, B, GL>::WSIZE)) > 8
                                   ? 8 : (int)((110 * 1024 - 16 * (XHalf<A, B, GL>::TW_SLOTS + XHalf<A, B, GL>::int_slots())) / (16 * XHalf<A, B, GL>::WSIZE)));
  auto kern = k_xh<A, B, GL, WARPS, DIR>;
  const size_t smem = xh_smem_bytes<A, B, GL>(WARPS);
  int cps = 1;
#ifndef ABI_EMU
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, WARPS * 32, smem));
  ABI_CHECK(cps >= 1, "half-support x pass: kernel does not fit on an SM");
#endif
  const long long nunits = (long long)P.nbatch * P.nb;
  long long grid = std::min<long long>((nunits + WARPS - 1) / WARPS, (long long)kNumSM * cps);
#ifdef ABI_EMU
  grid = std::min<long long>(nunits, 3);
#endif
  ABI_LAUNCH(kern, dim3((unsigned)std::max<long long>(grid, 1)), dim3(WARPS * 32), smem, st, P);
}

template <int A, int B>
void xh_launch(int dir, XhParams& P, cudaStream_t st) {
  if (dir == 0) xh_launch_dir<A, B, 0>(P, st);
  else xh_launch_dir<A, B, 1>(P, st);
}

}  // namespace abi
