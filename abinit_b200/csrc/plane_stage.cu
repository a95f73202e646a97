// Dispatch table of the fourwf plane stage (see plane_stage.cuh for the algorithm; plane_inst_*.cu hold the kernels).
#include "plane_stage.cuh"
#include "context.cuh"

namespace abi {

template <int R1, int R2> void plane_launch_n(PlaneParams& P, cudaStream_t st);   // plane_stage_impl.cuh
template <int R1, int R2> void plane_launch_rho_n(PlaneParams& P, cudaStream_t st);

namespace {
struct Scratch {
  void* p = nullptr; size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) { if (p) CUDA_CHECK(cudaFree(p)); CUDA_CHECK(cudaMalloc(&p, bytes)); cap = bytes; }
    return p;
  }
} g_scratch;

typedef void (*LaunchFn)(PlaneParams&, cudaStream_t);
struct Entry { int n; LaunchFn fn; LaunchFn fn_rho; };
#define PLANE_ENTRY(R1, R2) {R1 * R2, &plane_launch_n<R1, R2>, &plane_launch_rho_n<R1, R2>}
const Entry kEntries[] = {
    PLANE_ENTRY(4, 6), PLANE_ENTRY(5, 6), PLANE_ENTRY(4, 8), PLANE_ENTRY(6, 6), PLANE_ENTRY(5, 8), PLANE_ENTRY(5, 9), PLANE_ENTRY(6, 8), PLANE_ENTRY(5, 10), PLANE_ENTRY(6, 9), PLANE_ENTRY(7, 8), PLANE_ENTRY(6, 10), PLANE_ENTRY(8, 8), PLANE_ENTRY(8, 9), PLANE_ENTRY(5, 15), PLANE_ENTRY(8, 10), PLANE_ENTRY(9, 9), PLANE_ENTRY(7, 12), PLANE_ENTRY(9, 10), PLANE_ENTRY(8, 12), PLANE_ENTRY(10, 10),
    PLANE_ENTRY(9, 12), PLANE_ENTRY(8, 14), PLANE_ENTRY(10, 12), PLANE_ENTRY(8, 16), PLANE_ENTRY(9, 15), PLANE_ENTRY(12, 12), PLANE_ENTRY(10, 15), PLANE_ENTRY(10, 16), PLANE_ENTRY(12, 14), PLANE_ENTRY(12, 15), PLANE_ENTRY(12, 16), PLANE_ENTRY(14, 14), PLANE_ENTRY(15, 15), PLANE_ENTRY(15, 16), PLANE_ENTRY(16, 16),
};
const Entry* find_entry(int n) {
  for (const Entry& e : kEntries) if (e.n == n) return &e;
  return nullptr;
}
}  // namespace

void* plane_scratch_get(size_t bytes) { return g_scratch.get(bytes); }

bool plane_stage_supported(int n) { return find_entry(n) != nullptr; }

void plane_stage_launch(int n, PlaneParams& P, cudaStream_t st) {
  const Entry* e = find_entry(n);
  ABI_CHECK(e != nullptr, "plane stage: unsupported FFT length");
  e->fn(P, st);
}

void plane_stage_launch_rho(int n, PlaneParams& P, cudaStream_t st) {
  const Entry* e = find_entry(n);
  ABI_CHECK(e != nullptr, "plane stage: unsupported FFT length");
  e->fn_rho(P, st);
}

void plane_stage_release() { if (g_scratch.p) cudaFree(g_scratch.p); g_scratch.p = nullptr; g_scratch.cap = 0; }

}  // namespace abi
