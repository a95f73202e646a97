#include "context.cuh"
#include "fourwf.cuh"

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

}  // namespace abi
