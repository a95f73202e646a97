// Developer tool: single-thread host emulation of CUDA kernels' index logic (blockDim == 1).
// Lets the host-side planners and the non-tensor kernels (scatter / FFT passes / gather / epilogues) be
// debugged in a container without a GPU.  NOT a product path: abinit_b200 never loads a library built with
// ABI_EMU and the GPU tests never touch it.
#pragma once
#include <functional>
#include <vector>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x)
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct int2 { int x, y; };
struct short4 { short x, y, z, w; };
struct int4 { int x, y, z, w; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef int cudaError_t; typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return 0; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return 0; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
namespace abi_emu {
inline dim3& tIdx() { static dim3 v; return v; }
inline dim3& bIdx() { static dim3 v; return v; }
inline dim3& bDim() { static dim3 v; return v; }
inline dim3& gDim() { static dim3 v; return v; }
inline unsigned char* smem_pool() { static std::vector<unsigned char> pool(1 << 20); return pool.data(); }
inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  (void)block; (void)smem;
  gDim() = grid; bDim() = dim3(1, 1, 1); tIdx() = dim3(0, 0, 0);
  for (unsigned z = 0; z < grid.z; z++) for (unsigned y = 0; y < grid.y; y++) for (unsigned x = 0; x < grid.x; x++) {
    bIdx() = dim3(x, y, z); bIdx().x = x; bIdx().y = y; bIdx().z = z;
    body();
  }
}
}  // namespace abi_emu
#define threadIdx (abi_emu::tIdx())
#define blockIdx (abi_emu::bIdx())
#define blockDim (abi_emu::bDim())
#define gridDim (abi_emu::gDim())
inline void __syncthreads() {}
#include <algorithm>
using std::min; using std::max;
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline int __double2hiint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b >> 32); }
inline int __double2loint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
inline double __hiloint2double(int hi, int lo) { long long b = ((long long)(unsigned)hi << 32) | (unsigned)lo; double x; memcpy(&x, &b, 8); return x; }
