// Dispatch table of the fourwf plane stage (see plane_stage.cuh for the algorithm; plane_inst_*.cu hold the kernels).
#include "plane_stage.cuh"
#include "fourwf.cuh"
#include "context.cuh"

namespace abi {

template <int R1, int R2> void plane_launch_n(PlaneParams& P, cudaStream_t st);   // plane_stage_impl.cuh
template <int R1, int R2> void plane_launch_rho_n(PlaneParams& P, cudaStream_t st);
template <int R1, int R2> void plane_launch_split_n(int kind, PlaneParams& P, cudaStream_t st);

namespace {
struct Scratch {
  void* p = nullptr; size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) { if (p) CUDA_CHECK(cudaFree(p)); CUDA_CHECK(cudaMalloc(&p, bytes)); cap = bytes; }
    return p;
  }
} g_scratch_all[kMaxLanes];
#define g_scratch g_scratch_all[ctx().lane]

typedef void (*LaunchFn)(PlaneParams&, cudaStream_t);
typedef void (*SplitFn)(int, PlaneParams&, cudaStream_t);
struct Entry { int n; LaunchFn fn; LaunchFn fn_rho; SplitFn fn_split; };
#define PLANE_ENTRY(R1, R2) {R1 * R2, &plane_launch_n<R1, R2>, &plane_launch_rho_n<R1, R2>, &plane_launch_split_n<R1, R2>}
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

// split path (n2 != n3): S planes of every unit in global memory
static double2* split_scratch(const PlaneParams& P) {
  return (double2*)g_scratch.get(sizeof(double2) * (size_t)P.nunits * P.nU * P.n2);
}

void plane_stage_launch(PlaneParams& P, cudaStream_t st) {
  const Entry* e2 = find_entry(P.n2);
  const Entry* e3 = find_entry(P.n3);
  ABI_CHECK(e2 != nullptr && e3 != nullptr, "plane stage: unsupported FFT length");
  if (P.n2 == P.n3 && !fourwf_tuning().plane_split) { e2->fn(P, st); return; }
  P.S = split_scratch(P);
  { ProfScope ps("fourwf_plane_split_y"); e2->fn_split(0, P, st); }          // y on the occupied z planes
  { ProfScope ps("fourwf_plane_split_z"); e3->fn_split(1, P, st); }          // z, * V_loc, z^-1 on every column
  { ProfScope ps("fourwf_plane_split_yinv"); e2->fn_split(2, P, st); }       // y^-1 onto the output lines
}

void plane_stage_launch_rho(PlaneParams& P, cudaStream_t st) {
  const Entry* e2 = find_entry(P.n2);
  const Entry* e3 = find_entry(P.n3);
  ABI_CHECK(e2 != nullptr && e3 != nullptr, "plane stage: unsupported FFT length");
  if (P.n2 == P.n3) { e2->fn_rho(P, st); return; }
  P.S = split_scratch(P);
  e2->fn_split(0, P, st);
  e3->fn_split(3, P, st);          // z + density accumulation
}

void plane_stage_release() { for (auto& g : g_scratch_all) { if (g.p) cudaFree(g.p); g.p = nullptr; g.cap = 0; } }

}  // namespace abi
