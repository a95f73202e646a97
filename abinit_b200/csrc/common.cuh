// Common definitions for the abinit_b200 CUDA library (sm_100a only).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>

#ifdef ABI_EMU
// Developer-only single-thread emulation of the kernels' index logic (tools/emu). Never shipped, never
// loaded by the abinit_b200 package: see tools/emu/README.md.
#include "emu_shim.h"
#else
#include <cuda_runtime.h>
#endif

#define ABI_HD __host__ __device__ __forceinline__
#define ABI_DEV __device__ __forceinline__

// Error convention of the reference GPU code (src/46_manage_cuda/gpu_fourwf.cu:196-207,
// shared/common/src/17_gpu_toolbox/cuda_api_error_check.h): print a message and abort -- no status codes,
// no silent CPU fallback.
[[noreturn]] inline void abi_b200_abort(const char* file, int line, const char* msg) {
  fprintf(stderr, "\n--- !ERROR\nsrc_file: %s\nsrc_line: %d\nmessage: |\n    abinit_b200: %s\n...\n", file, line, msg);
  fflush(stderr);
  abort();
}
#define ABI_ERROR(msg) abi_b200_abort(__FILE__, __LINE__, (msg))
#define ABI_CHECK(cond, msg) do { if (!(cond)) abi_b200_abort(__FILE__, __LINE__, (msg)); } while (0)
#define CUDA_CHECK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    char b__[512]; snprintf(b__, sizeof b__, "CUDA error '%s' in %s", cudaGetErrorString(e__), #call); \
    abi_b200_abort(__FILE__, __LINE__, b__); } } while (0)

#ifdef ABI_EMU
#define ABI_LAUNCH(kernel, grid, block, smem, stream, ...) \
  abi_emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define ABI_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(abi_emu::smem_pool())
#else
#define ABI_LAUNCH(kernel, grid, block, smem, stream, ...) do { \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); CUDA_CHECK(cudaGetLastError()); } while (0)
#define ABI_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char abi_dyn_smem__[]; \
  type* name = reinterpret_cast<type*>(abi_dyn_smem__)
#endif

namespace abi {
constexpr int kNumSM = 148;            // B200
constexpr int kMaxFftLen = 1024;       // pos field of packed table entries is 10 bits
constexpr size_t kMaxSmemPerCta = 227 * 1024;

template <typename T> ABI_HD T ceil_div(T a, T b) { return (a + b - 1) / b; }

#ifndef ABI_EMU
ABI_DEV double2 ldcg2(const double2* p) { return __ldcg(p); }
ABI_DEV void stcg2(double2* p, double2 v) { __stcg(p, v); }
ABI_DEV double ldg1(const double* p) { return __ldg(p); }
ABI_DEV double2 ldg2(const double2* p) { return __ldg(p); }
#else
inline double2 ldcg2(const double2* p) { return *p; }
inline void stcg2(double2* p, double2 v) { *p = v; }
inline double ldg1(const double* p) { return *p; }
inline double2 ldg2(const double2* p) { return *p; }
#endif
}  // namespace abi
