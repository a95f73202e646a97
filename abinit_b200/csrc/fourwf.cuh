// fourwf on sm_100a: plans, tables and kernel entry points (internal header).
// Reference semantics: src/53_ffts/m_fft.F90:2290-2940 (fourwf), src/46_ghc_omp/m_ompgpu_fourwf.F90:179-585.
#pragma once
#include "common.cuh"
#include "fft_engine.cuh"
#include <vector>
#include <memory>

namespace abi {

// device-resident tables for one FFT length
struct FftTables {
  Fft1d plan;                                  // device pointers inside
  std::vector<unsigned short> pos_of_idx;      // host copies (used by the planners)
  std::vector<unsigned short> idx_of_pos;
};
const FftTables& fft_tables(int n);            // cached; aborts if n has a prime factor > 7 or n > kMaxFftLen
bool fft_length_supported(int n);
void fft_tables_clear();

// Epilogue fused into the last pass of option 2 when called from getghc (m_getghc.F90:1266-1280):
//   out = (kinpw < huge*1e-11) ? fourwf + kinpw*cwavef + gvnlxc : 0
struct FourwfEpilogue {
  int mode = 0;                  // 0: plain fourwf; 1: + kinetic + nonlocal with filter; 2: filter only (type_calc=1)
  const double* kinpw = nullptr; // device, npw_out
  const double2* cwavef = nullptr;  // device, ndat*npw_out
  const double2* gvnlxc = nullptr;  // device or null
  double2* gsc = nullptr;        // device or null: zeroed where filtered (sij_opt==1)
};

// fused getghc assembly of one output coefficient (m_getghc.F90:1266-1280, type_calc=1 filter :1003-1031)
ABI_DEV bool fw_epilogue(const FourwfEpilogue& epi, double kin_filter, int ipw, size_t o, double2& v) {
  if (epi.mode == 1) {
    const double k = epi.kinpw[ipw];
    if (k < kin_filter) {
      const double2 c = epi.cwavef[o];
      v.x = v.x + k * c.x; v.y = v.y + k * c.y;
      if (epi.gvnlxc) { const double2 g = epi.gvnlxc[o]; v.x += g.x; v.y += g.y; }
    } else {
      if (epi.gsc) epi.gsc[o] = make_double2(0.0, 0.0);
      return false;
    }
  } else if (epi.mode == 2) {
    if (epi.kinpw[ipw] > kin_filter) return false;
  }
  return true;
}


struct XhTabs;
void xh_tabs_free(XhTabs* t);

struct FourwfPlan {
  int n1 = 0, n2 = 0, n3 = 0, istwf_k = 1, me_g0 = 1;
  int npw_in = 0, npw_out = 0;
  bool same_sphere = false;      // kg_kin == kg_kout (lets the Gamma-point path pack two bands per transform)
  uint64_t key = 0;
  // ---- generic path (options 0,1,3 and fallback shapes) ----
  int nent_in = 0;
  int* d_in_src = nullptr;       // ipw | conj<<31 | zero_imag<<30
  int* d_in_box = nullptr;       // linear box index
  int* d_out_box = nullptr;      // npw_out
  // ---- fused option-2 path ----
  bool fused_ok = false;
  int nlin = 0, nlout = 0, nU = 0;
  int2* d_in_ent = nullptr;      // sorted by in-line: {src, (line<<10)|pos1}
  int* d_lin_estart = nullptr;   // nlin+1
  int* d_lin_u = nullptr;        // nlin
  unsigned short* d_lin_pos2 = nullptr;  // nlin: DIT input position of the line's i2
  int* d_inpl_start = nullptr;   // nU+1
  int2* d_out_ent = nullptr;     // sorted by out-line: {ipw, (line<<10)|pos1}
  int* d_lout_estart = nullptr;  // nlout+1
  int* d_lout_u = nullptr;
  unsigned short* d_lout_pos2 = nullptr;
  int* d_outpl_start = nullptr;  // nU+1
  unsigned short* d_u_i3 = nullptr;      // nU
  unsigned char* d_u_flags = nullptr;    // bit0: has input lines, bit1: has output lines
  // ---- plane-stage tables (plane_stage.cuh): the i2 of the rows of a plane, and the occupied i3, are each at most two
  // contiguous runs [a,a+la) U [b,b+lb) (true for every convex G-sphere, shifted or time-reversal completed)
  bool plane_ok = false;
  int za = 0, zla = 0, zb = 0, zlb = 0;
  int* d_pin_start = nullptr; short4* d_pin_runs = nullptr;      // nU each
  int* d_pout_start = nullptr; short4* d_pout_runs = nullptr;
  // ---- half-support plane stage (half_stage.cuh): every row / the occupied z planes are [0, la) U [n - lb, n), la, lb <= n/2
  bool half_ok_in = false, half_ok_out = false;   // input rows + occupied z planes / output rows
  int h_za = 0, h_zla = 0, h_zb = 0, h_zlb = 0;   // occupied z planes [za, za + zla) U [zb, zb + zlb), zb >= n3 / 2
  int y_amb_in = 0, y_amb_out = 0;
  int4* d_hin_rows = nullptr; int4* d_hout_rows = nullptr;       // nU each: {first line, a | la << 16, (b - n2/2) | lb << 16, 0}
  // per-configuration z tables, built by the launcher at first use (mutable cache; the plan is otherwise immutable)
  mutable int h_cfg_key = -1; mutable bool h_z_ok = false;
  mutable int* d_hz_ovoff = nullptr; mutable int* d_hz_sign = nullptr; mutable int* d_hu_row = nullptr;
  mutable std::vector<void*> owned_lazy;
  // host copies of the line / entry tables (the half-support x kernels derive their own tables from them at first use)
  std::vector<int2> h_in_ent, h_out_ent; std::vector<int> h_in_estart, h_out_estart; std::vector<int2> h_lin_i2i3;
  mutable XhTabs* xh = nullptr;
  std::vector<int> h_kg_in, h_kg_out;   // the spheres the plan was built for (compared on a cache hit; h_kg_out empty = same array)
  std::vector<void*> owned;      // device allocations to free
  void release();
  FourwfPlan() = default;
  FourwfPlan(const FourwfPlan&) = delete;
  FourwfPlan& operator=(const FourwfPlan&) = delete;
  ~FourwfPlan();
};

// Build (or fetch from the cache) the plan for (kg_in, kg_out). kg arrays are HOST pointers, Fortran layout kg(3,npw).
FourwfPlan* fourwf_get_plan(const int* kg_in, int npw_in, const int* kg_out, int npw_out, const int* ngfft,
                            int istwf_k, int me_g0);
// same, returning a reference the caller keeps for as long as it uses the plan (Hamiltonian handles): the plan then survives
// fourwf_clear_plans / free_gpu_fourwf_ / cache trimming
std::shared_ptr<FourwfPlan> fourwf_get_plan_shared(const int* kg_in, int npw_in, const int* kg_out, int npw_out, const int* ngfft,
                                                   int istwf_k, int me_g0);
void fourwf_clear_plans();

// V_loc on the device: natural layout [i3][i2][i1] (cplex doubles per point) and the x-slowest transpose
// [i1][i3][i2] used by the fused z pass (built once per load_spin).
struct VlocDev {
  int cplex = 1, n1 = 0, n2 = 0, n3 = 0;
  double* d_v = nullptr;
  double* d_vT = nullptr;
  uint64_t stamp = 0;
  // V_loc in the register order of the half-support z pass (half_stage.cuh), built at first use per (stamp, configuration)
  mutable double* d_vP = nullptr; mutable size_t vP_cap = 0; mutable uint64_t vP_stamp = ~0ULL; mutable int vP_key = -1;
};
void vloc_upload(VlocDev& v, const double* denpot, bool on_device, int cplex, int n1, int n2, int n3,
                 cudaStream_t st);

// All pointers below are DEVICE pointers.
void fourwf_generic(const FourwfPlan& pl, int option, int cplex, double* d_denpot, const double2* d_fofgin,
                    double2* d_fofgout, double2* d_fofr, int ndat, const double* d_wr, const double* d_wi,
                    cudaStream_t st);
void fourwf_fused_opt2(const FourwfPlan& pl, const VlocDev& v, const double2* d_fofgin, double2* d_fofgout,
                       int ndat, const FourwfEpilogue& epi, cudaStream_t st);

// fused option 1 (density accumulation) on the zero-padded plane stage; d_wr/d_wi: DEVICE arrays of ndat weights
bool fourwf_fused_opt1_available(const FourwfPlan& pl);
void fourwf_fused_opt1(const FourwfPlan& pl, const double2* d_fofgin, double* d_denpot, int ndat, const double* d_wr,
                       const double* d_wi, cudaStream_t st);

// tuning knobs (env ABI_B200_* override), reported by bench.py
struct FourwfTuning {
  int cluster = 0; int lines_x = 32; int smem_kb_mid = 0; int band_chunk = 0;
  int plane = 1;               // 1: register-resident two-pass plane stage (plane_stage.cuh) when the box allows it
  int plane_cfg = 0;           // 0 auto, 1: (G=8, 4 warps), 2: (G=4, 8 warps)
  int plane_ctas_per_sm = 0;   // 0: occupancy / L2-budget limited
  int plane_split = 0;         // 1: run the split (three-kernel) plane stage for cubic boxes too (developer comparison)
  int half = 1;                // 1: half-support plane stage (half_stage.cuh) when the sphere fits in half of the box axes
  int xhalf = 3;               // half-support x passes (x_stage.cuh) when the plan allows them: bit 0 forward (K1), bit 1 backward (K3)
  int xh_order = 1;            // x passes: bit 0 = batch index fastest (neighbouring warps touch neighbouring runs), bit 1 = evict_first on the W1o strip
  int half_skip = 0;           // developer timing aid (phase mask), see HalfParams::dbg_skip
  int half_cfg = 1;            // 0: 8 warps x 2 CTAs/SM, 1: 16 warps x 1 CTA/SM (measured on B200, Si-512: 3.55 vs 3.44 ms)
  int pack2 = 1;               // istwf_k=2: two bands per complex transform (double_rfft_trick, m_getghc.F90:1999-2171)
};
FourwfTuning& fourwf_tuning();

// launch counter for bench.py's gpu_launches
extern long long g_kernel_launches;

}  // namespace abi
