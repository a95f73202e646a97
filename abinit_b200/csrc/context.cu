#include "context.cuh"
#include "fourwf.cuh"
#include <string>
#ifndef ABI_EMU
#include <nvtx3/nvToolsExt.h>
#endif

namespace abi {

void* Staging::get(size_t bytes) {
  if (bytes > cap) {
    if (d) CUDA_CHECK(cudaFree(d));
    size_t want = bytes + bytes / 8 + 256;
    CUDA_CHECK(cudaMalloc(&d, want));
    cap = want;
  }
  return d;
}
void Staging::release() { if (d) cudaFree(d); d = nullptr; cap = 0; }

Context& ctx() { static Context c; return c; }

void ensure_init() {
  Context& c = ctx();
  if (c.initialized) return;
#ifndef ABI_EMU
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    ABI_ERROR("no CUDA device visible: abinit_b200 has no CPU fallback (north-star: sm_100a only)");
  CUDA_CHECK(cudaGetDevice(&c.device));
  if (!c.stream) { CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking)); c.own_stream = true; }
  if (!c.copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  if (const char* e = getenv("ABI_B200_PIPELINE")) c.pipeline = atoi(e) != 0;
  if (const char* e = getenv("ABI_B200_PIPE_CHUNKS")) c.pipe_chunks = atoi(e) > 0 ? atoi(e) : c.pipe_chunks;
#endif
  c.initialized = true;
}

bool is_device_ptr(const void* p) {
#ifdef ABI_EMU
  (void)p; return false;
#else
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
#endif
}

DevArg::DevArg(int slot, const void* p, size_t nbytes, bool in) {
  host = const_cast<void*>(p); bytes = nbytes;
  if (p == nullptr || nbytes == 0) { dev = nullptr; return; }
  if (is_device_ptr(p)) { dev = host; staged = false; return; }
  Context& c = ctx();
  dev = c.stage[slot].get(nbytes);
  staged = true;
  if (in) CUDA_CHECK(cudaMemcpyAsync(dev, p, nbytes, cudaMemcpyHostToDevice, c.stream));
}
void DevArg::copy_back() {
  if (staged && dev) CUDA_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx().stream));
}

// ---- profiler ----
#ifndef ABI_EMU
struct ProfRec { int id; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<std::string> g_prof_names;
static std::vector<double> g_prof_ms;
static std::vector<long long> g_prof_cnt;
static std::vector<ProfRec> g_prof_open;
static std::vector<cudaEvent_t> g_prof_pool;
static int prof_id(const char* name) {
  for (size_t i = 0; i < g_prof_names.size(); i++) if (g_prof_names[i] == name) return (int)i;
  g_prof_names.push_back(name); g_prof_ms.push_back(0.0); g_prof_cnt.push_back(0);
  return (int)g_prof_names.size() - 1;
}
static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); return e;
}
static void prof_drain() {
  if (g_prof_open.empty()) return;
  CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
  for (auto& r : g_prof_open) {
    float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    g_prof_ms[r.id] += ms; g_prof_cnt[r.id]++;
    g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
  }
  g_prof_open.clear();
}
NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

ProfScope::ProfScope(const char* name) : id(-1) {
  nvtxRangePushA(name);
  if (!g_prof_on) return;
  if (g_prof_open.size() >= 4096) prof_drain();
  id = prof_id(name);
  ProfRec r; r.id = id; r.a = prof_event(); r.b = prof_event();
  CUDA_CHECK(cudaEventRecord(r.a, ctx().stream));
  g_prof_open.push_back(r); slot = (int)g_prof_open.size() - 1;
}
ProfScope::~ProfScope() {
  nvtxRangePop();
  if (id < 0 || slot < 0 || slot >= (int)g_prof_open.size()) return;
  cudaEventRecord(g_prof_open[slot].b, ctx().stream);
}

// ---- FP64 pipe probe (bench.py measures its roofline denominator in the run it reports) ----
__global__ void __launch_bounds__(256) k_probe_dfma(double* out, int iters) {
  double a[16]; const double x = threadIdx.x * 1e-9 + 1.0, y = 0.999999;
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = i * 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0; for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_probe_dmma(double* out, int iters) {
  double c[8][2]; const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
void probe_fp64_peak(double* dfma_tflops, double* dmma_tflops) {
  ensure_init();
  cudaStream_t st = ctx().stream;
  const int grid = kNumSM * 4, block = 256, iters = 20000;
  double* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof(double) * grid * block));
  cudaEvent_t e0, e1; CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  double best[2] = {0.0, 0.0};
  for (int kind = 0; kind < 2; kind++)
    for (int rep = 0; rep < 4; rep++) {
      CUDA_CHECK(cudaEventRecord(e0, st));
      if (kind == 0) k_probe_dfma<<<grid, block, 0, st>>>(d, iters); else k_probe_dmma<<<grid, block, 0, st>>>(d, iters);
      CUDA_CHECK(cudaEventRecord(e1, st));
      CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      // DFMA: 16 FMAs per thread and iteration; DMMA m8n8k4: 8 x 8 x 4 FMAs per warp instruction, 8 instructions per iteration
      const double flops = kind == 0 ? 2.0 * 16 * (double)iters * grid * block : 2.0 * 256 * 8 * (double)iters * grid * (block / 32);
      if (rep > 0) best[kind] = std::max(best[kind], flops / (ms * 1e-3) / 1e12);
    }
  CUDA_CHECK(cudaEventDestroy(e0)); CUDA_CHECK(cudaEventDestroy(e1)); CUDA_CHECK(cudaFree(d));
  *dfma_tflops = best[0]; *dmma_tflops = best[1];
}
void prof_enable(bool on) {
  if (!on) prof_drain();
  g_prof_on = on;
  if (on) { for (auto& v : g_prof_ms) v = 0.0; for (auto& v : g_prof_cnt) v = 0; }
}
int prof_collect(char* names, int names_cap, double* ms, long long* counts, int cap) {
  prof_drain();
  int n = 0; std::string all;
  for (size_t i = 0; i < g_prof_names.size() && n < cap; i++) {
    if (g_prof_cnt[i] == 0) continue;
    ms[n] = g_prof_ms[i]; counts[n] = g_prof_cnt[i]; all += g_prof_names[i]; all += ";"; n++;
  }
  snprintf(names, names_cap, "%s", all.c_str());
  return n;
}
#else
NvtxRange::NvtxRange(const char*) {}
NvtxRange::~NvtxRange() {}
void probe_fp64_peak(double* a, double* b) { *a = 0.0; *b = 0.0; }
ProfScope::ProfScope(const char*) : id(-1) {}
ProfScope::~ProfScope() {}
void prof_enable(bool) {}
int prof_collect(char*, int, double*, long long*, int) { return 0; }
#endif

}  // namespace abi
