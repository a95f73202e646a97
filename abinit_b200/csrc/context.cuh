// Process-wide state of the library: device, stream, staging buffers for host-resident arguments.
#pragma once
#include "common.cuh"
#include <vector>

namespace abi {

struct Staging {             // grow-only device buffer used to mirror one host argument
  void* d = nullptr; size_t cap = 0;
  void* get(size_t bytes);
  void release();
};

constexpr int kMaxLanes = 32;   // concurrent streams of the batched small-system path (getghc_batch); lane 0 = the library stream

struct Context {
  bool initialized = false;
  int device = 0;
  cudaStream_t stream = 0;
  cudaStream_t copy_stream = 0;   // host<->device staging pipeline of getghc (H2D / D2H overlapped with compute)
  bool pipeline = true;
  int pipe_chunks = 16;           // band chunks (H2D) / row slabs (D2H) of the pipelined host-array getghc (B200: e2e 2835 at 8, 2889 at 16 and 24)
  bool own_stream = false;
  bool async = false;
  int me_g0 = 1;
  int fourwf_impl = 0;          // 0 auto, 1 generic, 2 fused
  long long fourwf_counter = 0; // m_fft.F90:2333-2336
  long long nonlop_counter = 0; // m_nonlop.F90:389-392
  Staging stage[16];
  // lanes: every internal workspace exists once per lane, so calls issued on different lanes may overlap on the device
  int lane = 0;
  cudaStream_t lane_stream[kMaxLanes] = {};
  bool force_scratch_clear = false;   // set while a call is captured into a CUDA graph (see half_scratch_get)
};
Context& ctx();
void ensure_init();

bool is_device_ptr(const void* p);

// Mirror of one argument on the device: if `p` is a host pointer, a staging buffer is used (copied in when
// `in`), and copy_back() copies results to the host; if `p` is already a device pointer it is used as is.
struct DevArg {
  void* host = nullptr; void* dev = nullptr; size_t bytes = 0; bool staged = false;
  DevArg() {}
  DevArg(int slot, const void* p, size_t bytes, bool in);
  void copy_back();
  template <typename T> T* as() const { return reinterpret_cast<T*>(dev); }
};

// Optional per-kernel-class device timers (CUDA events on the library stream) for bench.py's roofline object.
struct ProfScope {
  int id; int slot = -1;
  explicit ProfScope(const char* name);
  ~ProfScope();
};
// NVTX range on the calling thread, named like the reference's regions (src/44_abitools/m_nvtx_data.F90:157-240: "GETGHC",
// "LOCPOT", "NLOCPOT", "KINETIC", "CHEBFI2", "RAYLRITZ", ...; pushed at the same places, m_getghc.F90:264,400,1042,1162) so that an
// Nsight Systems timeline of this library lines up with one of the reference.  Every ProfScope (one per kernel class) is a range too.
struct NvtxRange {
  explicit NvtxRange(const char* name);
  ~NvtxRange();
};
void prof_enable(bool on);
// FP64 pipe peak of the current device measured now (register-resident DFMA / DMMA m8n8k4 loops, best of 3): TFLOP/s
void probe_fp64_peak(double* dfma_tflops, double* dmma_tflops);
int prof_collect(char* names, int names_cap, double* ms, long long* counts, int cap);   // syncs; returns #classes

}  // namespace abi
