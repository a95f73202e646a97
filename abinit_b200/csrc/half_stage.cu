// Host side of the half-support plane stage (half_stage.cuh): configuration table, per-plan z tables, V_loc permutation.
#include "half_stage.cuh"
#include "fourwf.cuh"
#include "context.cuh"
#include <map>
#include <vector>

namespace abi {

template <int A, int B, int G> void half_launch_n(int kind, HalfParams& P, cudaStream_t st);   // half_stage_impl.cuh

namespace {
typedef void (*HalfFn)(int, HalfParams&, cudaStream_t);
struct Entry { HalfCfg cfg; HalfFn fn; };
#define HALF_ENTRY(A, B, G) {{2 * (A) * (B), A, B, G}, &half_launch_n<A, B, G>}
const Entry kEntries[] = {
    HALF_ENTRY(4, 3, 8), HALF_ENTRY(5, 3, 10), HALF_ENTRY(4, 4, 8), HALF_ENTRY(9, 2, 9), HALF_ENTRY(4, 5, 8), HALF_ENTRY(8, 3, 8),
    HALF_ENTRY(4, 8, 4), HALF_ENTRY(8, 8, 4), HALF_ENTRY(9, 10, 3),
};
const Entry* find_entry(int n) {
  for (const Entry& e : kEntries) if (e.cfg.n == n) return &e;
  return nullptr;
}
int h_gcd_rt(int a, int b) { return b == 0 ? a : h_gcd_rt(b, a % b); }
int h_inv_rt(int a, int m) { for (int i = 0; i < m; i++) if ((a * i) % m == 1 % m) return i; return 0; }
int rin_rt(const HalfCfg& c, int t, int j) { return h_gcd_rt(c.A, c.B) == 1 ? (c.B * t + c.A * j) % (c.A * c.B) : j + c.B * t; }
int kout_rt(const HalfCfg& c, int k1, int k2) {
  if (h_gcd_rt(c.A, c.B) != 1) return k1 + c.A * k2;
  return (k1 * c.B * h_inv_rt(c.B % c.A, c.A) + k2 * c.A * h_inv_rt(c.A % c.B, c.B)) % (c.A * c.B);
}
template <typename T> T* upload(const std::vector<T>& v, std::vector<void*>& owned) {
  T* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  owned.push_back(d);
  return d;
}
struct HScratch { void* p = nullptr; size_t cap = 0; uint64_t key = 0; } g_hscratch_all[kMaxLanes];
#define g_hscratch g_hscratch_all[ctx().lane]

// permutation tables of one (n2, n3) pair: grid index of every slot of the vP / rhoP layout
struct PermTab { int* d_i2_of_cid = nullptr; int* d_i3_of_slot = nullptr; };
std::map<std::pair<int, int>, PermTab>& perm_cache() { static std::map<std::pair<int, int>, PermTab> c; return c; }
std::vector<void*>& perm_owned() { static std::vector<void*> v; return v; }

const PermTab& perm_tables(const HalfCfg& c2, const HalfCfg& c3) {
  auto key = std::make_pair(c2.n, c3.n);
  auto it = perm_cache().find(key);
  if (it != perm_cache().end()) return it->second;
  const int G = c2.G, ng2 = (c2.n + G - 1) / G;
  std::vector<int> i2_of_cid((size_t)ng2 * G, -1), i3_of_slot(c3.n);
  for (int k2 = 0; k2 < c2.B; k2++) for (int h = 0; h < 2; h++) for (int k1 = 0; k1 < c2.A; k1++)
    i2_of_cid[k2 * 2 * c2.A + h * c2.A + k1] = 2 * kout_rt(c2, k1, k2) + h;
  for (int k2 = 0; k2 < c3.B; k2++) for (int h = 0; h < 2; h++) for (int k1 = 0; k1 < c3.A; k1++)
    i3_of_slot[k2 * 2 * c3.A + h * c3.A + k1] = 2 * kout_rt(c3, k1, k2) + h;
  PermTab t;
  t.d_i2_of_cid = upload(i2_of_cid, perm_owned());
  t.d_i3_of_slot = upload(i3_of_slot, perm_owned());
  return perm_cache()[key] = t;
}

// vP[i1][g][k2][c][hk1] = vT[i1][i3(k2, hk1)][i2(g G + c)]   (cplex doubles per point)
template <int CPLEX>
__global__ void k_vloc_permute(const double* __restrict__ vT, double* __restrict__ vP, const int* __restrict__ i2_of_cid,
                               const int* __restrict__ i3_of_slot, int n1, int n2, int n3, int ng2, int G, int A2x2, int B3) {
  const long long per_i1 = (long long)ng2 * B3 * G * A2x2;
  const long long total = per_i1 * n1;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int i1 = (int)(o / per_i1);
    long long r = o - (long long)i1 * per_i1;
    const int hk1 = (int)(r % A2x2); r /= A2x2;
    const int c = (int)(r % G); r /= G;
    const int k2 = (int)(r % B3);
    const int g = (int)(r / B3);
    const int i2 = i2_of_cid[g * G + c];
    const int i3 = i3_of_slot[k2 * A2x2 + hk1];
    if (CPLEX == 1) vP[o] = i2 >= 0 ? vT[((size_t)i1 * n3 + i3) * n2 + i2] : 0.0;
    else {
      vP[2 * o] = i2 >= 0 ? vT[2 * (((size_t)i1 * n3 + i3) * n2 + i2)] : 0.0;
      vP[2 * o + 1] = i2 >= 0 ? vT[2 * (((size_t)i1 * n3 + i3) * n2 + i2) + 1] : 0.0;
    }
  }
}

// denpot[i3][i2][i1] += rhoP[i1][g][k2][c][hk1]
__global__ void k_rho_unpermute_add(const double* __restrict__ rhoP, double* __restrict__ rho, const int* __restrict__ i2_of_cid,
                                    const int* __restrict__ i3_of_slot, int n1, int n2, int n3, int ng2, int G, int A2x2, int B3) {
  const long long per_i1 = (long long)ng2 * B3 * G * A2x2;
  const long long total = per_i1 * n1;
  // i1 fastest across threads: the writes of a warp are contiguous in denpot, the reads stride by one permuted plane
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int i1 = (int)(o % n1);
    long long r = o / n1;
    const int hk1 = (int)(r % A2x2); r /= A2x2;
    const int c = (int)(r % G); r /= G;
    const int k2 = (int)(r % B3);
    const int g = (int)(r / B3);
    const int i2 = i2_of_cid[g * G + c];
    if (i2 < 0) continue;
    const int i3 = i3_of_slot[k2 * A2x2 + hk1];
    const long long src = (long long)i1 * per_i1 + (((long long)g * B3 + k2) * G + c) * A2x2 + hk1;
    rho[((size_t)i3 * n2 + i2) * n1 + i1] += rhoP[src];
  }
}

// z tables of the plan for one configuration (cached in the plan)
bool ensure_z_tables(const FourwfPlan& pl, const HalfCfg& c3, int G) {
  const int key = c3.n * 1000 + c3.A * 32 + G;
  if (pl.h_cfg_key == key) return pl.h_z_ok;
  for (void* p : pl.owned_lazy) cudaFree(p);
  pl.owned_lazy.clear();
  const int M = c3.A * c3.B, za = pl.h_za, zla = pl.h_zla, zbm = pl.h_zb - M, zlb = pl.h_zlb;
  std::vector<int> sign(M, 0), ovrow(M, -1), rowu(M + kHalfOV, -1);
  int nov = 0, nplanes = 0;
  for (int t = 0; t < c3.A; t++) for (int j = 0; j < c3.B; j++) {
    const int q = t * c3.B + j, r = rin_rt(c3, t, j);
    const bool lo = r >= za && r < za + zla, hi = r >= zbm && r < zbm + zlb;
    const int u_lo = r - za, u_hi = zla + r - zbm;
    if (lo) { sign[q] = 1; rowu[q] = u_lo; nplanes++; }
    if (hi) {
      if (lo) { ovrow[q] = M + nov; if (nov < kHalfOV) rowu[M + nov] = u_hi; nov++; }
      else { sign[q] = -1; rowu[q] = u_hi; }
      nplanes++;
    }
  }
  pl.h_cfg_key = key;
  pl.h_z_ok = nplanes == pl.nU && nov <= kHalfOV;
  if (!pl.h_z_ok) return false;
  pl.d_hz_sign = upload(sign, pl.owned_lazy);
  pl.d_hz_ovoff = upload(ovrow, pl.owned_lazy);
  pl.d_hu_row = upload(rowu, pl.owned_lazy);
  return true;
}

void fill_params(const FourwfPlan& pl, const HalfCfg& c2, const HalfCfg& c3, const HalfLaunch& L, HalfParams& P) {
  ABI_CHECK(ensure_z_tables(pl, c3, c2.G), "half-support plane stage: unsupported occupied z planes");
  P.n1 = pl.n1; P.n2 = pl.n2; P.n3 = pl.n3; P.nb = L.nb; P.nU = pl.nU; P.cplex = 1;
  P.nlin = L.nlin; P.nlout = L.nlout; P.nunits = (long long)L.nb * pl.n1;
  P.W1 = L.W1; P.W1o = L.W1o; P.S = nullptr; P.vP = nullptr;
  P.tw2 = fft_tables(pl.n2).plan.tw; P.tw3 = fft_tables(pl.n3).plan.tw;
  P.in_rows = pl.d_hin_rows; P.out_rows = L.out_is_in ? pl.d_hin_rows : pl.d_hout_rows;
  P.y_amb_in = pl.y_amb_in; P.y_amb_out = L.out_is_in ? pl.y_amb_in : pl.y_amb_out;
  P.z_sign = pl.d_hz_sign; P.z_ovrow = pl.d_hz_ovoff; P.row_u = pl.d_hu_row;
  P.layout_key = pl.key ^ ((unsigned long long)pl.h_cfg_key << 40) ^ 0x9e3779b97f4a7c15ULL;
  P.ng2 = (pl.n2 + c2.G - 1) / c2.G;
  P.rhoP = L.rhoP; P.wxy = L.wxy;
  P.dbg_skip = fourwf_tuning().half_skip;
}
}  // namespace

const HalfCfg* half_stage_cfg(int n) { const Entry* e = find_entry(n); return e ? &e->cfg : nullptr; }

void* half_scratch_get(size_t bytes, uint64_t layout_key, cudaStream_t st) {
  HScratch& h = g_hscratch;
  if (bytes > h.cap) { if (h.p) CUDA_CHECK(cudaFree(h.p)); CUDA_CHECK(cudaMalloc(&h.p, bytes)); h.cap = bytes; h.key = 0; }
  // (a call being captured into a CUDA graph always clears: at replay time another plan may have used the buffer in between)
  if (h.key != layout_key || ctx().force_scratch_clear) { CUDA_CHECK(cudaMemsetAsync(h.p, 0, h.cap, st)); h.key = layout_key; }   // the whole buffer: later launches of this plan may use more of it
  return h.p;
}

bool half_stage_usable(const FourwfPlan& pl, bool out_is_in) {
  if (!(pl.fused_ok && pl.half_ok_in && (out_is_in || pl.half_ok_out) && fourwf_tuning().half && pl.n2 == pl.n3)) return false;
  const Entry* e = find_entry(pl.n2);
  return e != nullptr && ensure_z_tables(pl, e->cfg, e->cfg.G);
}

void half_stage_launch(const FourwfPlan& pl, const VlocDev& v, const HalfLaunch& L, cudaStream_t st) {
  const Entry* e2 = find_entry(pl.n2);
  const Entry* e3 = find_entry(pl.n3);
  ABI_CHECK(e2 != nullptr && e3 != nullptr && pl.half_ok_in, "half-support plane stage: unsupported plan");
  const HalfCfg &c2 = e2->cfg, &c3 = e3->cfg;
  HalfParams P;
  fill_params(pl, c2, c3, L, P);
  P.cplex = v.cplex;
  // V_loc in z-pass register order, rebuilt when the potential or the configuration changed
  const int vkey = (c2.n * 1000 + c3.n) * 8 + v.cplex;
  const size_t per_i1 = (size_t)P.ng2 * c3.B * c2.G * 2 * c3.A;
  const size_t vbytes = sizeof(double) * v.cplex * per_i1 * pl.n1;
  if (v.vP_stamp != v.stamp || v.vP_key != vkey || v.d_vP == nullptr) {
    if (vbytes > v.vP_cap) { if (v.d_vP) CUDA_CHECK(cudaFree(v.d_vP)); CUDA_CHECK(cudaMalloc(&v.d_vP, vbytes)); v.vP_cap = vbytes; }
    const PermTab& pt = perm_tables(c2, c3);
    if (v.cplex == 1) ABI_LAUNCH(k_vloc_permute<1>, dim3(kNumSM * 8), dim3(256), 0, st, v.d_vT, v.d_vP, pt.d_i2_of_cid, pt.d_i3_of_slot,
                                 pl.n1, pl.n2, pl.n3, P.ng2, c2.G, 2 * c3.A, c3.B);
    else ABI_LAUNCH(k_vloc_permute<2>, dim3(kNumSM * 8), dim3(256), 0, st, v.d_vT, v.d_vP, pt.d_i2_of_cid, pt.d_i3_of_slot,
                    pl.n1, pl.n2, pl.n3, P.ng2, c2.G, 2 * c3.A, c3.B);
    g_kernel_launches++;
    v.vP_stamp = v.stamp; v.vP_key = vkey;
  }
  P.vP = v.d_vP;
  e2->fn(0, P, st);
}

size_t half_rho_elems(const FourwfPlan& pl) {
  const HalfCfg* c2 = half_stage_cfg(pl.n2); const HalfCfg* c3 = half_stage_cfg(pl.n3);
  ABI_CHECK(c2 && c3, "half-support plane stage: unsupported plan");
  const int ng2 = (pl.n2 + c2->G - 1) / c2->G;
  return (size_t)pl.n1 * ng2 * c3->B * c2->G * 2 * c3->A;
}

void half_stage_launch_rho(const FourwfPlan& pl, const HalfLaunch& L, cudaStream_t st) {
  const Entry* e2 = find_entry(pl.n2);
  const Entry* e3 = find_entry(pl.n3);
  ABI_CHECK(e2 != nullptr && e3 != nullptr && pl.half_ok_in, "half-support plane stage: unsupported plan");
  HalfParams P;
  fill_params(pl, e2->cfg, e3->cfg, L, P);
  e2->fn(1, P, st);
}

void half_rho_unpermute_add(const FourwfPlan& pl, const double* rhoP, double* denpot, cudaStream_t st) {
  const HalfCfg* c2 = half_stage_cfg(pl.n2); const HalfCfg* c3 = half_stage_cfg(pl.n3);
  const PermTab& pt = perm_tables(*c2, *c3);
  const int ng2 = (pl.n2 + c2->G - 1) / c2->G;
  ABI_LAUNCH(k_rho_unpermute_add, dim3(kNumSM * 8), dim3(256), 0, st, rhoP, denpot, pt.d_i2_of_cid, pt.d_i3_of_slot, pl.n1, pl.n2,
             pl.n3, ng2, c2->G, 2 * c3->A, c3->B);
  g_kernel_launches++;
}

void half_stage_release() {
  for (auto& g : g_hscratch_all) { if (g.p) cudaFree(g.p); g = HScratch(); }
  for (void* p : perm_owned()) cudaFree(p);
  perm_owned().clear(); perm_cache().clear();
}

}  // namespace abi
