// C-ABI entry points: lifecycle + fourwf.  See include/abinit_b200.h for the reference interfaces replaced.
#include "../../include/abinit_b200.h"
#include "context.cuh"
#include "fourwf.cuh"
#include "nonlop.cuh"
#include <string>
#include <algorithm>

using namespace abi;

namespace abi {
void fourwf_release_workspace();
void nonlop_release_all();
static VlocDev g_vloc_call;     // V_loc staged by the plain fourwf entry point (one per call, reused buffer)
}

namespace abi { void xg_release_workspace(); void chebfi_release_workspace(); }

extern "C" {

const char* abi_b200_version(void) { return "abinit_b200 0.1 (sm_100a; getghc = fourwf + gemm_nonlop)"; }

void abi_b200_init(int rank) {
#ifndef ABI_EMU
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) ABI_ERROR("no CUDA device visible: abinit_b200 has no CPU fallback");
  CUDA_CHECK(cudaSetDevice(rank % ndev));   // m_initcuda.F90:326-333
#else
  (void)rank;
#endif
  ensure_init();
}

void abi_b200_finalize(void) {
  Context& c = ctx();
  if (!c.initialized) return;
#ifndef ABI_EMU
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
#endif
  fourwf_clear_plans();
  fft_tables_clear();
  fourwf_release_workspace();
  nonlop_release_all();
  xg_release_workspace();
  chebfi_release_workspace();
  for (auto& s : c.stage) s.release();
  if (g_vloc_call.d_v) { cudaFree(g_vloc_call.d_v); cudaFree(g_vloc_call.d_vT); g_vloc_call = VlocDev(); }
#ifndef ABI_EMU
  if (c.own_stream) { cudaStreamDestroy(c.stream); c.stream = 0; c.own_stream = false; }
  if (c.copy_stream) { cudaStreamDestroy(c.copy_stream); c.copy_stream = 0; }
#endif
  c.initialized = false;
}

void abi_b200_set_stream(void* s) {
  Context& c = ctx();
#ifndef ABI_EMU
  if (c.own_stream && c.stream) { cudaStreamSynchronize(c.stream); cudaStreamDestroy(c.stream); c.own_stream = false; }
  c.stream = (cudaStream_t)s;
  if (!s && c.initialized) { CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking)); c.own_stream = true; }
#else
  (void)s; (void)c;
#endif
}
void abi_b200_set_async(int flag) { ctx().async = flag != 0; }
void abi_b200_synchronize(void) { ensure_init(); CUDA_CHECK(cudaStreamSynchronize(ctx().stream)); }
long long abi_b200_kernel_launches(void) { return g_kernel_launches; }
void abi_b200_profile_enable(int on) { ensure_init(); prof_enable(on != 0); }
int abi_b200_profile_collect(char* names, int names_cap, double* ms, long long* counts, int cap) {
  ensure_init(); return prof_collect(names, names_cap, ms, counts, cap);
}
void abi_b200_probe_fp64_peak(double* dfma_tflops, double* dmma_tflops) { probe_fp64_peak(dfma_tflops, dmma_tflops); }
void abi_b200_set_me_g0(int me_g0) { ctx().me_g0 = me_g0; }
void abi_b200_fourwf_set_impl(int impl) { ctx().fourwf_impl = impl; }
long long abi_b200_fourwf_counter(void) { return ctx().fourwf_counter; }
void abi_b200_fourwf_set_tuning(const char* name, int value) {
  FourwfTuning& t = fourwf_tuning();
  const std::string k(name ? name : "");
  if (k == "plane") t.plane = value;
  else if (k == "plane_cfg") t.plane_cfg = value;
  else if (k == "pack2") t.pack2 = value;
  else if (k == "half") t.half = value;
  else if (k == "half_cfg") t.half_cfg = value;
  else if (k == "half_skip") t.half_skip = value;
  else if (k == "xhalf") t.xhalf = value;
  else if (k == "xh_order") t.xh_order = value;
  else if (k == "pipeline") ctx().pipeline = value != 0;
  else if (k == "nonlop_ozaki") ozaki_set_enabled(value);     // EXPERIMENTAL int8-sliced gemm_nonlop (ozaki.cu), default off
  else if (k == "nonlop_rag") nonlop_set_rag(value);
  else if (k == "pipe_chunks") ctx().pipe_chunks = std::max(1, value);
  else if (k == "plane_ctas_per_sm") t.plane_ctas_per_sm = value;
  else if (k == "plane_split") t.plane_split = value;
  else if (k == "cluster") t.cluster = value;
  else if (k == "lines_x") t.lines_x = value;
  else if (k == "smem_kb_mid") t.smem_kb_mid = value;
  else if (k == "band_chunk") t.band_chunk = value;
  else ABI_ERROR("abi_b200_fourwf_set_tuning: unknown knob");
}

void abi_b200_alloc_fourwf_(int* ngfft, int* ndat, int* npwin, int* npwout) {
  (void)ndat; (void)npwin; (void)npwout;
  ensure_init();
  for (int i = 0; i < 3; i++) fft_tables(ngfft[i]);    // build twiddle/permutation tables ahead of the first call
}
void abi_b200_free_fourwf_(void) { fourwf_clear_plans(); fourwf_release_workspace(); }

void abi_b200_fourwf_(int* cplex, double* denpot, double* fofgin, double* fofgout, double* fofr, int* gboundin,
                      int* gboundout, int* istwf_k, int* kg_kin, int* kg_kout, int* mgfft, void* mpi_enreg, int* ndat,
                      int* ngfft, int* npwin, int* npwout, int* n4, int* n5, int* n6, int* option, int* paral_kgb,
                      int* tim_fourwf, double* weight_r, double* weight_i) {
  (void)gboundin; (void)gboundout; (void)mgfft; (void)mpi_enreg; (void)paral_kgb; (void)tim_fourwf;
  ensure_init();
  NvtxRange nvtx("FOURWF");                               // NVTX_FOURWF
  Context& c = ctx();
  const int n1 = ngfft[0], n2 = ngfft[1], n3 = ngfft[2];
  const int opt = *option, nd = *ndat;
  // m_fft.F90:2333-2336
  c.fourwf_counter += nd; if (opt == 2) c.fourwf_counter += nd;
  if (n1 != *n4 || n2 != *n5 || n3 != *n6) {
    char b[256];
    snprintf(b, sizeof b, "FFT SIZE ERROR: when gpu mode is on the fft grid must not be augmented (n1,n2,n3)=(%d,%d,%d) whereas (n4,n5,n6)=(%d,%d,%d)",
             n1, n2, n3, *n4, *n5, *n6);
    ABI_ERROR(b);
  }
  ABI_CHECK(opt >= 0 && opt <= 3, "Only option=0, 1, 2 or 3 are allowed presently.");
  ABI_CHECK(!(opt == 1 && *cplex != 1), "With the option number 1, cplex must be 1");
  ABI_CHECK(!(opt == 2 && *cplex != 1 && *cplex != 2), "With the option number 2, cplex must be 1 or 2");
  ABI_CHECK(!(*cplex == 2 && *istwf_k != 1), "cplex=2 only allowed for istwf_k=1");
  ABI_CHECK(nd >= 1, "ndat must be >= 1");
  ABI_CHECK(!is_device_ptr(kg_kin) && !is_device_ptr(kg_kout), "kg_kin/kg_kout must be host arrays (they define the plan)");
  const size_t N = (size_t)n1 * n2 * n3;
  // option 3 only uses the output sphere; give the planner a valid input sphere
  const int* kin = (opt == 3) ? kg_kout : kg_kin;
  const int npin = (opt == 3) ? *npwout : *npwin;
  const int* kout = (opt == 0 || opt == 1) ? kin : kg_kout;
  const int npout = (opt == 0 || opt == 1) ? npin : *npwout;
  FourwfPlan* pl = fourwf_get_plan(kin, npin, kout, npout, ngfft, *istwf_k, c.me_g0);

  DevArg a_in, a_out, a_fofr, a_den, a_wr, a_wi;
  if (opt != 3) a_in = DevArg(0, fofgin, sizeof(double) * 2 * (size_t)npin * nd, true);
  if (opt == 2 || opt == 3) a_out = DevArg(1, fofgout, sizeof(double) * 2 * (size_t)npout * nd, false);
  bool fused = (opt == 2) && pl->fused_ok && c.fourwf_impl != 1;
  if (c.fourwf_impl == 2 && opt == 2) ABI_CHECK(pl->fused_ok, "fused fourwf requested but not available for this box");
  if (opt == 0 || opt == 3) a_fofr = DevArg(2, fofr, sizeof(double) * 2 * N * nd, opt == 3);
  if (opt == 1 || (opt == 2 && !fused)) a_den = DevArg(3, denpot, sizeof(double) * (*cplex) * N, true);
  if (opt == 1) {
    a_wr = DevArg(4, weight_r, sizeof(double) * nd, true);
    a_wi = DevArg(5, weight_i, sizeof(double) * nd, true);
  }
  if (opt == 1 && c.fourwf_impl != 1 && fourwf_fused_opt1_available(*pl)) {
    // density accumulation on the fused zero-padded path (fofr is not produced, as in the reference's GPU paths)
    fourwf_fused_opt1(*pl, a_in.as<double2>(), a_den.as<double>(), nd, a_wr.as<double>(), a_wi.as<double>(), c.stream);
  } else if (fused) {
    vloc_upload(g_vloc_call, denpot, is_device_ptr(denpot), *cplex, n1, n2, n3, c.stream);
    FourwfEpilogue epi;
    fourwf_fused_opt2(*pl, g_vloc_call, a_in.as<double2>(), a_out.as<double2>(), nd, epi, c.stream);
  } else {
    fourwf_generic(*pl, opt, *cplex, a_den.as<double>(), a_in.as<double2>(), a_out.as<double2>(),
                   a_fofr.as<double2>(), nd, a_wr.as<double>(), a_wi.as<double>(), c.stream);
  }
  if (opt == 2 || opt == 3) a_out.copy_back();
  if (opt == 0) a_fofr.copy_back();
  if (opt == 1) a_den.copy_back();
  const bool any_staged = a_in.staged || a_out.staged || a_fofr.staged || a_den.staged;
  if (!c.async || any_staged) CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void gpu_fourwf_(int* cplex, double* denpot, double* fofgin, double* fofgout, double* fofr, int* gboundin,
                 int* gboundout, int* istwf_k, int* kg_kin, int* kg_kout, int* mgfft, void* mpi_enreg, int* ndat,
                 int* ngfft, int* npwin, int* npwout, int* n4, int* n5, int* n6, int* option, int* paral_kgb,
                 int* tim_fourwf, double* weight_r, double* weight_i) {
  abi_b200_fourwf_(cplex, denpot, fofgin, fofgout, fofr, gboundin, gboundout, istwf_k, kg_kin, kg_kout, mgfft,
                   mpi_enreg, ndat, ngfft, npwin, npwout, n4, n5, n6, option, paral_kgb, tim_fourwf, weight_r,
                   weight_i);
}
void alloc_gpu_fourwf_(int* ngfft, int* ndat, int* npwin, int* npwout) { abi_b200_alloc_fourwf_(ngfft, ndat, npwin, npwout); }
void free_gpu_fourwf_(void) { abi_b200_free_fourwf_(); }

}  // extern "C"
