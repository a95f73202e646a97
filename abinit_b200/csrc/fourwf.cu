// fourwf for sm_100a: sphere<->box transforms with V_loc application, hand-written FP64 FFTs.
//
// Reference semantics (not code): src/53_ffts/m_fft.F90:2290-2940 (fourwf),
// src/52_fft_mpi_noabirule/m_fftcore.F90:1532-1866 (sphere), src/44_abitools/m_cgtools.F90:2227-2491
// (cg_box2gsph, cg_vlocpsi, cg_addtorho), src/46_ghc_omp/m_ompgpu_fourwf.F90:179-585 (offload twin).
//
// Two implementations live here:
//  (A) generic: scatter into the full box, three batched 1-D passes, point-wise kernel, three passes, gather.
//      Used for options 0, 1, 3 (secondary options) and as an on-device cross-check of (B).
//  (B) fused option 2 (the getghc hot path), zero-padded / pruned exactly where the reference's
//      fftw3_fftpad.finc:14-196 prunes:
//        K1  sphere lines -> x FFT (only the C (i2,i3) lines that hold plane waves)      -> W1[b][i1][line]
//        K2  per (band, i1) yz-plane: y FFT on occupied z planes, z FFT on every column, * V_loc,
//            inverse z, inverse y -- the plane never leaves the chip/L2 (cluster of CTAs + L2 scratch)
//        K3  x FFT^-1 on the output lines, gather to the sphere, * 1/N, fused kinetic/non-local assembly
//      HBM traffic per band: read psi, write W1, read W1, write W1', read W1', write H psi (+V via L2).
#include "fourwf.cuh"
#include "plane_stage.cuh"
#include "half_stage.cuh"
#include "x_stage.cuh"
#include "context.cuh"
#include <algorithm>
#include <cstring>
#include <map>
#include <unordered_map>
#include <memory>
#ifndef ABI_EMU
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#endif

namespace abi {

long long g_kernel_launches = 0;

FourwfTuning& fourwf_tuning() {
  static FourwfTuning t;
  static bool init = false;
  if (!init) {
    init = true;
    if (const char* e = getenv("ABI_B200_FOURWF_CLUSTER")) t.cluster = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_LINES_X")) t.lines_x = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_SMEM_KB")) t.smem_kb_mid = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_BAND_CHUNK")) t.band_chunk = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_PLANE")) t.plane = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_PLANE_CFG")) t.plane_cfg = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_PLANE_CTAS")) t.plane_ctas_per_sm = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_PACK2")) t.pack2 = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_HALF_CFG")) t.half_cfg = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_HALF")) t.half = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_XHALF")) t.xhalf = atoi(e);
    if (const char* e = getenv("ABI_B200_FOURWF_HALF_CFG")) t.half_cfg = atoi(e);
  }
  return t;
}

// ---------------------------------------------------------------------------------------------------------
// FFT tables
// ---------------------------------------------------------------------------------------------------------
static bool factorize(int n, std::vector<int>& fac) {
  fac.clear();
  if (n < 1) return false;
  int m = n;
  int c2 = 0, c3 = 0, c5 = 0, c7 = 0;
  while (m % 2 == 0) { m /= 2; c2++; }
  while (m % 3 == 0) { m /= 3; c3++; }
  while (m % 5 == 0) { m /= 5; c5++; }
  while (m % 7 == 0) { m /= 7; c7++; }
  if (m != 1) return false;
  // few, large radices: 9 = 3x3, 8 = 2^3, 6 = 2x3, 4 = 2^2
  while (c3 >= 2) { fac.push_back(9); c3 -= 2; }
  while (c2 >= 3) { fac.push_back(8); c2 -= 3; }
  while (c7 >= 1) { fac.push_back(7); c7--; }
  while (c5 >= 1) { fac.push_back(5); c5--; }
  if (c2 >= 1 && c3 >= 1) { fac.push_back(6); c2--; c3--; }
  while (c2 >= 2) { fac.push_back(4); c2 -= 2; }
  while (c3 >= 1) { fac.push_back(3); c3--; }
  while (c2 >= 1) { fac.push_back(2); c2--; }
  if (fac.empty()) fac.push_back(1);
  return fac.size() <= 8;
}

bool fft_length_supported(int n) {
  std::vector<int> f;
  return n >= 1 && n <= kMaxFftLen && factorize(n, f);
}

static std::map<int, std::unique_ptr<FftTables>>& table_cache() {
  static std::map<int, std::unique_ptr<FftTables>> c;
  return c;
}

static void twiddle(int j, int n, double& c, double& s) {
  // exp(-2 pi i j/n) with octant reduction in long double
  long double x = (long double)j / (long double)n;  // turns in [0,1)
  const long double twopi = 6.283185307179586476925286766559005768L;
  int oct = (int)floorl(x * 8.0L);
  long double cc, ss;
  switch (oct) {
    case 0: cc = cosl(twopi * x); ss = sinl(twopi * x); break;
    case 1: { long double y = 0.25L - x; cc = sinl(twopi * y); ss = cosl(twopi * y); } break;
    case 2: { long double y = x - 0.25L; cc = -sinl(twopi * y); ss = cosl(twopi * y); } break;
    case 3: { long double y = 0.5L - x; cc = -cosl(twopi * y); ss = sinl(twopi * y); } break;
    case 4: { long double y = x - 0.5L; cc = -cosl(twopi * y); ss = -sinl(twopi * y); } break;
    case 5: { long double y = 0.75L - x; cc = -sinl(twopi * y); ss = -cosl(twopi * y); } break;
    case 6: { long double y = x - 0.75L; cc = sinl(twopi * y); ss = -cosl(twopi * y); } break;
    default: { long double y = 1.0L - x; cc = cosl(twopi * y); ss = -sinl(twopi * y); } break;
  }
  c = (double)cc; s = -(double)ss;
}

const FftTables& fft_tables(int n) {
  auto& cache = table_cache();
  auto it = cache.find(n);
  if (it != cache.end()) return *it->second;
  std::vector<int> fac;
  if (n > kMaxFftLen || !factorize(n, fac)) {
    char b[256];
    snprintf(b, sizeof b, "FFT length %d not supported (need 2^a 3^b 5^c 7^d <= %d)", n, kMaxFftLen);
    ABI_ERROR(b);
  }
  auto t = std::make_unique<FftTables>();
  t->plan.n = n;
  t->plan.nfac = (int)fac.size();
  for (int i = 0; i < 8; i++) t->plan.radix[i] = i < (int)fac.size() ? fac[i] : 1;
  // digit-reversal: p(k; n) = (k mod r) * m + p(k div r; m)
  t->pos_of_idx.resize(n);
  t->idx_of_pos.resize(n);
  for (int k = 0; k < n; k++) {
    int kk = k, blk = n, pos = 0;
    for (int f : fac) {
      int m = blk / f;
      pos += (kk % f) * m;
      kk /= f;
      blk = m;
    }
    t->pos_of_idx[k] = (unsigned short)pos;
    t->idx_of_pos[pos] = (unsigned short)k;
  }
  std::vector<double2> tw(n);
  for (int j = 0; j < n; j++) twiddle(j, n, tw[j].x, tw[j].y);
  double2* d_tw; unsigned short *d_p, *d_i;
  CUDA_CHECK(cudaMalloc(&d_tw, sizeof(double2) * n));
  CUDA_CHECK(cudaMalloc(&d_p, sizeof(unsigned short) * n));
  CUDA_CHECK(cudaMalloc(&d_i, sizeof(unsigned short) * n));
  CUDA_CHECK(cudaMemcpy(d_tw, tw.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(d_p, t->pos_of_idx.data(), sizeof(unsigned short) * n, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(d_i, t->idx_of_pos.data(), sizeof(unsigned short) * n, cudaMemcpyHostToDevice));
  t->plan.tw = d_tw; t->plan.pos_of_idx = d_p; t->plan.idx_of_pos = d_i;
  auto& ref = *t;
  cache[n] = std::move(t);
  return ref;
}

void fft_tables_clear() {
  for (auto& kv : table_cache()) {
    cudaFree((void*)kv.second->plan.tw);
    cudaFree((void*)kv.second->plan.pos_of_idx);
    cudaFree((void*)kv.second->plan.idx_of_pos);
  }
  table_cache().clear();
}

// ---------------------------------------------------------------------------------------------------------
// Workspace (static device buffers re-allocated when sizes grow, like gpu_fourwf.cu:104-131,210-214)
// ---------------------------------------------------------------------------------------------------------
struct Workspace {
  void* p = nullptr; size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) {
      if (p) CUDA_CHECK(cudaFree(p));
      size_t want = bytes + bytes / 8;
      CUDA_CHECK(cudaMalloc(&p, want));
      cap = want;
    }
    return p;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
static Workspace g_ws_all[kMaxLanes][4];
#define g_ws g_ws_all[ctx().lane]
void fourwf_release_workspace() { for (auto& l : g_ws_all) for (auto& w : l) w.release(); plane_stage_release(); half_stage_release(); }

// ---------------------------------------------------------------------------------------------------------
// Planner
// ---------------------------------------------------------------------------------------------------------
void FourwfPlan::release() {
  for (void* p : owned) cudaFree(p);
  owned.clear();
  for (void* p : owned_lazy) cudaFree(p);
  owned_lazy.clear();
  if (xh) { xh_tabs_free(xh); xh = nullptr; }
}

template <typename T> static T* to_device(const std::vector<T>& v, std::vector<void*>& owned) {
  T* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  owned.push_back(d);
  return d;
}

static uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ULL) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
  return h;
}

// Plans are shared: the cache holds one reference, every Hamiltonian handle that loaded the k-point holds another
// (fourwf_get_plan_shared), so clearing or trimming the cache never frees tables a handle still uses; the device tables go
// away with the last reference (FourwfPlan::~FourwfPlan).
static std::unordered_map<uint64_t, std::shared_ptr<FourwfPlan>>& plan_cache() {
  static std::unordered_map<uint64_t, std::shared_ptr<FourwfPlan>> c;
  return c;
}
FourwfPlan::~FourwfPlan() { release(); }
void fourwf_clear_plans() { plan_cache().clear(); }

static inline int wrapi(int g, int n) { return g < 0 ? g + n : g; }

FourwfPlan* fourwf_get_plan(const int* kg_in, int npw_in, const int* kg_out, int npw_out, const int* ngfft,
                            int istwf_k, int me_g0) {
  return fourwf_get_plan_shared(kg_in, npw_in, kg_out, npw_out, ngfft, istwf_k, me_g0).get();
}

std::shared_ptr<FourwfPlan> fourwf_get_plan_shared(const int* kg_in, int npw_in, const int* kg_out, int npw_out, const int* ngfft,
                                                   int istwf_k, int me_g0) {
  const int n1 = ngfft[0], n2 = ngfft[1], n3 = ngfft[2];
  uint64_t key = fnv1a(kg_in, sizeof(int) * 3 * (size_t)npw_in);
  if (kg_out != kg_in) key = fnv1a(kg_out, sizeof(int) * 3 * (size_t)npw_out, key);
  int meta[8] = {n1, n2, n3, istwf_k, me_g0, npw_in, npw_out, kg_out == kg_in};
  key = fnv1a(meta, sizeof meta, key);
  auto& cache = plan_cache();
  auto it = cache.find(key);
  if (it != cache.end()) {
    // the 64-bit hash only finds the candidate: the spheres themselves must match
    FourwfPlan& c = *it->second;
    const bool same = c.npw_in == npw_in && c.npw_out == npw_out && c.h_kg_in.size() == (size_t)3 * npw_in &&
                      memcmp(c.h_kg_in.data(), kg_in, sizeof(int) * 3 * (size_t)npw_in) == 0 &&
                      (c.h_kg_out.empty() ? kg_out == kg_in || memcmp(kg_in, kg_out, sizeof(int) * 3 * (size_t)npw_in) == 0
                                          : (c.h_kg_out.size() == (size_t)3 * npw_out && memcmp(c.h_kg_out.data(), kg_out, sizeof(int) * 3 * (size_t)npw_out) == 0));
    if (same) return it->second;
    cache.erase(it);                               // hash collision: rebuild (the old plan lives on in the handles that hold it)
  }
  if (cache.size() > 64) {                         // bound device memory held by stale k-points: drop the plans nobody holds
    for (auto q = cache.begin(); q != cache.end();) { if (q->second.use_count() == 1) q = cache.erase(q); else ++q; }
  }

  ABI_CHECK(istwf_k >= 1 && istwf_k <= 9, "istwf_k must be between 1 and 9");
  auto pl = std::make_shared<FourwfPlan>();
  pl->h_kg_in.assign(kg_in, kg_in + 3 * (size_t)npw_in);
  if (kg_out != kg_in) pl->h_kg_out.assign(kg_out, kg_out + 3 * (size_t)npw_out);
  pl->n1 = n1; pl->n2 = n2; pl->n3 = n3; pl->istwf_k = istwf_k; pl->me_g0 = me_g0;
  pl->npw_in = npw_in; pl->npw_out = npw_out; pl->key = key;
  pl->same_sphere = (kg_out == kg_in) || (npw_in == npw_out && memcmp(kg_in, kg_out, sizeof(int) * 3 * (size_t)npw_in) == 0);

  // ---- input entries: direct plane waves then time-reversed images (m_fftcore.F90:1598-1650) ----
  struct Ent { int src; int i1, i2, i3; };
  std::vector<Ent> ents;
  ents.reserve((size_t)npw_in * (istwf_k >= 2 ? 2 : 1));
  const long long N = (long long)n1 * n2 * n3;
  ABI_CHECK(N < (1LL << 31), "FFT box too large for 32-bit indexing");
  for (int ipw = 0; ipw < npw_in; ipw++) {
    int i1 = wrapi(kg_in[3 * ipw + 0], n1), i2 = wrapi(kg_in[3 * ipw + 1], n2), i3 = wrapi(kg_in[3 * ipw + 2], n3);
    ABI_CHECK(i1 >= 0 && i1 < n1 && i2 >= 0 && i2 < n2 && i3 >= 0 && i3 < n3, "kg_kin outside the FFT box");
    int src = ipw;
    if (istwf_k == 2 && me_g0 == 1 && ipw == 0) src |= (1 << 30);   // Im forced to zero at G=0
    ents.push_back({src, i1, i2, i3});
  }
  if (istwf_k >= 2) {
    const int s1 = (istwf_k == 2 || istwf_k == 4 || istwf_k == 6 || istwf_k == 8) ? n1 : n1 - 1;
    const int s2 = (istwf_k >= 2 && istwf_k <= 5) ? n2 : n2 - 1;
    const int s3 = (istwf_k == 2 || istwf_k == 3 || istwf_k == 6 || istwf_k == 7) ? n3 : n3 - 1;
    const int npwmin = (istwf_k == 2 && me_g0 == 1) ? 1 : 0;
    for (int ipw = npwmin; ipw < npw_in; ipw++) {
      int i1 = wrapi(kg_in[3 * ipw + 0], n1), i2 = wrapi(kg_in[3 * ipw + 1], n2), i3 = wrapi(kg_in[3 * ipw + 2], n3);
      int j1 = ((s1 - i1) % n1 + n1) % n1, j2 = ((s2 - i2) % n2 + n2) % n2, j3 = ((s3 - i3) % n3 + n3) % n3;
      ents.push_back({ipw | (int)(1u << 31), j1, j2, j3});
    }
  }
  // later writes win on collisions (reference order: direct, G=0 fix, then images)
  {
    std::unordered_map<int, int> last;
    last.reserve(ents.size() * 2);
    for (int e = 0; e < (int)ents.size(); e++) last[ents[e].i1 + n1 * (ents[e].i2 + n2 * ents[e].i3)] = e;
    if (last.size() != ents.size()) {
      std::vector<Ent> kept;
      for (int e = 0; e < (int)ents.size(); e++)
        if (last[ents[e].i1 + n1 * (ents[e].i2 + n2 * ents[e].i3)] == e) kept.push_back(ents[e]);
      ents.swap(kept);
    }
  }
  pl->nent_in = (int)ents.size();
  {
    std::vector<int> src(ents.size()), box(ents.size());
    for (size_t e = 0; e < ents.size(); e++) { src[e] = ents[e].src; box[e] = ents[e].i1 + n1 * (ents[e].i2 + n2 * ents[e].i3); }
    pl->d_in_src = to_device(src, pl->owned);
    pl->d_in_box = to_device(box, pl->owned);
    std::vector<int> obox(npw_out);
    for (int ipw = 0; ipw < npw_out; ipw++) {
      int i1 = wrapi(kg_out[3 * ipw + 0], n1), i2 = wrapi(kg_out[3 * ipw + 1], n2), i3 = wrapi(kg_out[3 * ipw + 2], n3);
      ABI_CHECK(i1 >= 0 && i1 < n1 && i2 >= 0 && i2 < n2 && i3 >= 0 && i3 < n3, "kg_kout outside the FFT box");
      obox[ipw] = i1 + n1 * (i2 + n2 * i3);
    }
    pl->d_out_box = to_device(obox, pl->owned);
  }

  // ---- fused-path tables ----
  pl->fused_ok = fft_length_supported(n1) && fft_length_supported(n2) && fft_length_supported(n3) &&
                 npw_in > 0 && npw_out > 0;
  if (pl->fused_ok) {
    const FftTables& t1 = fft_tables(n1);
    const FftTables& t2 = fft_tables(n2);
    // union of occupied z planes
    std::vector<unsigned char> zflag(n3, 0);
    for (auto& e : ents) zflag[e.i3] |= 1;
    std::vector<Ent> oents(npw_out);
    for (int ipw = 0; ipw < npw_out; ipw++) {
      oents[ipw] = {ipw, wrapi(kg_out[3 * ipw + 0], n1), wrapi(kg_out[3 * ipw + 1], n2), wrapi(kg_out[3 * ipw + 2], n3)};
      zflag[oents[ipw].i3] |= 2;
    }
    std::vector<int> u_of_i3(n3, -1);
    std::vector<unsigned short> u_i3; std::vector<unsigned char> u_flags;
    for (int i3 = 0; i3 < n3; i3++) if (zflag[i3]) { u_of_i3[i3] = (int)u_i3.size(); u_i3.push_back((unsigned short)i3); u_flags.push_back(zflag[i3]); }
    pl->nU = (int)u_i3.size();
    pl->d_u_i3 = to_device(u_i3, pl->owned);
    pl->d_u_flags = to_device(u_flags, pl->owned);

    // n2 == n3: fused plane kernel; n2 != n3 (both lengths in the two-pass table): the split plane stage (plane_stage.cuh)
    bool plane_ok = plane_stage_supported(n2) && plane_stage_supported(n3) && n2 < 32768 && n3 < 32768;
    // at most two contiguous runs of a sorted index list: {a, la, b, lb}; false if it needs more
    auto two_runs = [](const std::vector<int>& r, int out[4]) {
      out[0] = out[1] = out[2] = out[3] = 0;
      if (r.empty()) return true;
      size_t k = 1;
      while (k < r.size() && r[k] == r[k - 1] + 1) k++;
      out[0] = r[0]; out[1] = (int)k;
      if (k == r.size()) return true;
      out[2] = r[k]; out[3] = (int)(r.size() - k);
      for (size_t q = k + 1; q < r.size(); q++) if (r[q] != r[q - 1] + 1) return false;
      return true;
    };
    {
      std::vector<int> zs(u_i3.begin(), u_i3.end());
      int zr[4];
      if (!two_runs(zs, zr)) plane_ok = false;
      pl->za = zr[0]; pl->zla = zr[1]; pl->zb = zr[2]; pl->zlb = zr[3];
    }
    // half-support form of a sorted index list on an axis of length n = 2m: one run [a, a + la) inside [0, m) and one run
    // [b, b + lb) inside [m, n) (either may be empty)
    auto half_runs = [](const std::vector<int>& r, int n, int& a, int& la, int& b, int& lb) {
      a = la = lb = 0; b = n / 2;
      if (n % 2) return false;
      const int m = n / 2;
      for (int i : r) { if (i < m) la++; else lb++; }
      if (la > 0) a = r[0];
      if (lb > 0) b = r[la];
      for (int q = 0; q < la; q++) if (r[q] != a + q) return false;
      for (int q = 0; q < lb; q++) if (r[la + q] != b + q) return false;
      return true;
    };
    bool half_ok_in = true, half_ok_out = true;
    {
      std::vector<int> zs(u_i3.begin(), u_i3.end());
      if (!half_runs(zs, n3, pl->h_za, pl->h_zla, pl->h_zb, pl->h_zlb)) half_ok_in = false;
    }
    auto build = [&](std::vector<Ent>& es, bool dit_positions, int& nlines, int2*& d_ent, int*& d_estart, int*& d_lu,
                     unsigned short*& d_lpos2, int*& d_plstart, int*& d_pstart, short4*& d_pruns, int4*& d_hrows, int& y_amb, bool& half_ok) {
      // sort by (plane u, i2, then original order) -> lines of one plane are contiguous
      std::vector<int> order(es.size());
      for (size_t i = 0; i < es.size(); i++) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        int ka = u_of_i3[es[a].i3] * n2 + es[a].i2, kb = u_of_i3[es[b].i3] * n2 + es[b].i2;
        return ka < kb;
      });
      std::vector<int2> ent(es.size());
      std::vector<int> estart, lu, plstart(pl->nU + 1, 0);
      std::vector<unsigned short> lpos2;
      int prev = -1, line = -1;
      for (size_t k = 0; k < order.size(); k++) {
        const Ent& e = es[order[k]];
        int kk = u_of_i3[e.i3] * n2 + e.i2;
        if (kk != prev) {
          prev = kk; line++;
          estart.push_back((int)k);
          if (dit_positions) pl->h_lin_i2i3.push_back(make_int2(e.i2, e.i3));
          lu.push_back(u_of_i3[e.i3]);
          lpos2.push_back(t2.pos_of_idx[e.i2]);
          plstart[u_of_i3[e.i3] + 1]++;
        }
        ABI_CHECK(line < (1 << 21), "too many FFT lines for the packed table format");
        ent[k].x = e.src;
        ent[k].y = (line << 10) | (int)t1.pos_of_idx[e.i1];
      }
      estart.push_back((int)order.size());
      for (int u = 0; u < pl->nU; u++) plstart[u + 1] += plstart[u];
      nlines = line + 1;
      // plane-stage row descriptors: the i2 of the lines of plane u (ascending) as at most two runs
      {
        std::vector<int> pstart(pl->nU, 0); std::vector<short4> pruns(pl->nU);
        std::vector<std::vector<int>> rows(pl->nU);
        int pv = -1;
        for (size_t k = 0; k < order.size(); k++) {
          const Ent& e = es[order[k]];
          int kk = u_of_i3[e.i3] * n2 + e.i2;
          if (kk != pv) { pv = kk; rows[u_of_i3[e.i3]].push_back(e.i2); }
        }
        for (int u = 0; u < pl->nU; u++) {
          pstart[u] = plstart[u];
          int rr[4];
          if (!two_runs(rows[u], rr)) plane_ok = false;
          pruns[u].x = (short)rr[0]; pruns[u].y = (short)rr[1]; pruns[u].z = (short)rr[2]; pruns[u].w = (short)rr[3];
        }
        d_pstart = to_device(pstart, pl->owned); d_pruns = to_device(pruns, pl->owned);
        std::vector<int4> hrows(pl->nU);
        y_amb = 0;
        for (int u = 0; u < pl->nU; u++) {
          int a = 0, la = 0, b = 0, lb = 0;
          if (!half_runs(rows[u], n2, a, la, b, lb)) half_ok = false;
          const int bm = b - n2 / 2;                       // the high run in r = i2 - m coordinates
          hrows[u].x = plstart[u]; hrows[u].y = a | (la << 16); hrows[u].z = bm | (lb << 16); hrows[u].w = 0;
          if (la > 0 && lb > 0 && a < bm + lb && bm < a + la) y_amb = 1;
          if (la + lb > n2 / 2 + 2) half_ok = false;     // staging buffer of the plane stage: rows of at most m + kHalfOV entries
        }
        d_hrows = to_device(hrows, pl->owned);
      }
      d_ent = to_device(ent, pl->owned);
      d_estart = to_device(estart, pl->owned);
      if (dit_positions) { pl->h_in_ent = ent; pl->h_in_estart = estart; } else { pl->h_out_ent = ent; pl->h_out_estart = estart; }
      d_lu = to_device(lu, pl->owned);
      d_lpos2 = to_device(lpos2, pl->owned);
      d_plstart = to_device(plstart, pl->owned);
      (void)dit_positions;
    };
    build(ents, true, pl->nlin, pl->d_in_ent, pl->d_lin_estart, pl->d_lin_u, pl->d_lin_pos2, pl->d_inpl_start,
          pl->d_pin_start, pl->d_pin_runs, pl->d_hin_rows, pl->y_amb_in, half_ok_in);
    build(oents, false, pl->nlout, pl->d_out_ent, pl->d_lout_estart, pl->d_lout_u, pl->d_lout_pos2, pl->d_outpl_start,
          pl->d_pout_start, pl->d_pout_runs, pl->d_hout_rows, pl->y_amb_out, half_ok_out);
    pl->plane_ok = plane_ok;
    pl->half_ok_in = half_ok_in && pl->nU > 0;
    pl->half_ok_out = half_ok_out;
  }
  cache[key] = pl;
  return pl;
}

// ---------------------------------------------------------------------------------------------------------
// V_loc upload + transpose
// ---------------------------------------------------------------------------------------------------------
template <int CPLEX>
__global__ void k_vloc_transpose(const double* __restrict__ v, double* __restrict__ vT, int n1, int n2, int n3) {
  // vT[i1][i3][i2] = v[i3][i2][i1]; one thread per output element, i2 fastest on the write side
  const long long total = (long long)n1 * n2 * n3;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    int i2 = (int)(o % n2);
    long long r = o / n2;
    int i3 = (int)(r % n3);
    int i1 = (int)(r / n3);
    long long src = i1 + (long long)n1 * (i2 + (long long)n2 * i3);
    if (CPLEX == 1) vT[o] = v[src];
    else { vT[2 * o] = v[2 * src]; vT[2 * o + 1] = v[2 * src + 1]; }
  }
}

void vloc_upload(VlocDev& v, const double* denpot, bool on_device, int cplex, int n1, int n2, int n3, cudaStream_t st) {
  const size_t n = (size_t)cplex * n1 * n2 * n3;
  if (v.n1 * v.n2 * v.n3 * v.cplex != (int)n || v.d_v == nullptr) {
    if (v.d_v) cudaFree(v.d_v);
    if (v.d_vT) cudaFree(v.d_vT);
    CUDA_CHECK(cudaMalloc(&v.d_v, sizeof(double) * n));
    CUDA_CHECK(cudaMalloc(&v.d_vT, sizeof(double) * n));
  }
  v.cplex = cplex; v.n1 = n1; v.n2 = n2; v.n3 = n3;
  CUDA_CHECK(cudaMemcpyAsync(v.d_v, denpot, sizeof(double) * n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  if (cplex == 1) ABI_LAUNCH(k_vloc_transpose<1>, dim3(kNumSM * 8), dim3(256), 0, st, v.d_v, v.d_vT, n1, n2, n3);
  else ABI_LAUNCH(k_vloc_transpose<2>, dim3(kNumSM * 8), dim3(256), 0, st, v.d_v, v.d_vT, n1, n2, n3);
  g_kernel_launches++;
  v.stamp++;
}

// ---------------------------------------------------------------------------------------------------------
// (A) generic kernels
// ---------------------------------------------------------------------------------------------------------
__global__ void k_scatter(const double2* __restrict__ cg, double2* __restrict__ box, const int* __restrict__ src,
                          const int* __restrict__ boxidx, int nent, int npw, long long nbox) {
  const int b = blockIdx.y;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nent; e += gridDim.x * blockDim.x) {
    int s = src[e];
    double2 v = cg[(size_t)b * npw + (s & 0x3fffffff)];
    if (s < 0) v.y = -v.y;
    if (s & (1 << 30)) v.y = 0.0;
    box[(size_t)b * nbox + boxidx[e]] = v;
  }
}

__global__ void k_gather(const double2* __restrict__ box, double2* __restrict__ out, const int* __restrict__ boxidx,
                         int npw, long long nbox, double xnorm, int zero_im_g0) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npw; i += gridDim.x * blockDim.x) {
    double2 v = box[(size_t)b * nbox + boxidx[i]];
    v.x *= xnorm; v.y *= xnorm;
    if (zero_im_g0 && i == 0) v.y = 0.0;
    out[(size_t)b * npw + i] = v;
  }
}

template <int CPLEX>
__global__ void k_vlocpsi(double2* __restrict__ box, const double* __restrict__ v, long long nbox) {
  const int b = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nbox; i += (long long)gridDim.x * blockDim.x) {
    double2 p = box[(size_t)b * nbox + i];
    if (CPLEX == 1) { double w = v[i]; p.x *= w; p.y *= w; }
    else { double2 w = make_double2(v[2 * i], v[2 * i + 1]); p = cmul(p, w); }
    box[(size_t)b * nbox + i] = p;
  }
}

__global__ void k_addtorho(const double2* __restrict__ box, double* __restrict__ rho, const double* __restrict__ wr,
                           const double* __restrict__ wi, int ndat, long long nbox) {
  // m_ompgpu_fourwf.F90:455-470: band loop innermost, fixed order -> deterministic
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nbox; i += (long long)gridDim.x * blockDim.x) {
    double acc = rho[i];
    for (int b = 0; b < ndat; b++) {
      double2 p = box[(size_t)b * nbox + i];
      acc += p.x * p.x * wr[b];
      acc += p.y * p.y * wi[b];
    }
    rho[i] = acc;
  }
}

// batched in-place 1-D FFT along one axis of [nb][n3][n2][n1]
template <int SIGN>
__global__ void k_fft_axis(double2* __restrict__ data, Fft1d pl, long long nlines_total, int lines_per_cta,
                           long long inner, long long outer_stride, long long estride) {
  ABI_DYN_SMEM(double2, sm);
  const int n = pl.n;
  const int ls = n | 1;
  double2* tw = sm;
  double2* buf = sm + n;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int j = tid; j < n; j += nthr) tw[j] = pl.tw[j];
  const long long l0 = (long long)blockIdx.x * lines_per_cta;
  const int nl = (int)min((long long)lines_per_cta, nlines_total - l0);
  if (nl <= 0) return;
  if (estride == 1) {
    for (int w = tid; w < nl * n; w += nthr) {
      int l = w / n, p = w - l * n;
      long long L = l0 + l;
      buf[l * ls + p] = data[(L / inner) * outer_stride + (L % inner) + p];
    }
  } else {
    for (int w = tid; w < nl * n; w += nthr) {
      int p = w / nl, l = w - p * nl;
      long long L = l0 + l;
      buf[l * ls + p] = data[(L / inner) * outer_stride + (L % inner) + (long long)p * estride];
    }
  }
  __syncthreads();
  fft_lines_dif<SIGN>(buf, ls, nl, pl, tw, tid, nthr);
  if (estride == 1) {
    for (int w = tid; w < nl * n; w += nthr) {
      int l = w / n, k = w - l * n;
      long long L = l0 + l;
      data[(L / inner) * outer_stride + (L % inner) + k] = buf[l * ls + pl.pos_of_idx[k]];
    }
  } else {
    for (int w = tid; w < nl * n; w += nthr) {
      int k = w / nl, l = w - k * nl;
      long long L = l0 + l;
      data[(L / inner) * outer_stride + (L % inner) + (long long)k * estride] = buf[l * ls + pl.pos_of_idx[k]];
    }
  }
}

template <int SIGN>
static void fft3d_generic(double2* d_box, int nb, int n1, int n2, int n3, cudaStream_t st) {
  const long long N = (long long)n1 * n2 * n3;
  for (int axis = 0; axis < 3; axis++) {
    const int n = axis == 0 ? n1 : axis == 1 ? n2 : n3;
    const FftTables& t = fft_tables(n);
    long long nlines = (long long)nb * N / n;
    long long inner, ostride, estride;
    if (axis == 0) { inner = 1; ostride = n1; estride = 1; }
    else if (axis == 1) { inner = n1; ostride = (long long)n1 * n2; estride = n1; }
    else { inner = (long long)n1 * n2; ostride = N; estride = (long long)n1 * n2; }
    int ls = n | 1;
    int lpc = (int)std::max<long long>(1, std::min<long long>(32, (96 * 1024 - 16 * n) / (16LL * ls)));
    // a batch must not straddle an 'outer' boundary for strided axes (lines of one batch share the outer index)
    if (axis != 0) { while (inner % lpc != 0) lpc--; }
    size_t smem = sizeof(double2) * ((size_t)n + (size_t)lpc * ls);
#ifndef ABI_EMU
    CUDA_CHECK(cudaFuncSetAttribute(k_fft_axis<SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
    long long nblk = ceil_div<long long>(nlines, lpc);
    ABI_LAUNCH(k_fft_axis<SIGN>, dim3((unsigned)nblk), dim3(256), smem, st, d_box, t.plan, nlines, lpc, inner, ostride, estride);
    g_kernel_launches++;
  }
}

void fourwf_generic(const FourwfPlan& pl, int option, int cplex, double* d_denpot, const double2* d_fofgin,
                    double2* d_fofgout, double2* d_fofr, int ndat, const double* d_wr, const double* d_wi,
                    cudaStream_t st) {
  const int n1 = pl.n1, n2 = pl.n2, n3 = pl.n3;
  const long long N = (long long)n1 * n2 * n3;
  ABI_CHECK(option >= 0 && option <= 3, "Only option=0, 1, 2 or 3 are allowed presently.");
  ABI_CHECK(!(option == 1 && cplex != 1), "With the option number 1, cplex must be 1");
  ABI_CHECK(!(option == 2 && cplex != 1 && cplex != 2), "With the option number 2, cplex must be 1 or 2");
  double2* box = d_fofr;
  if (box == nullptr) box = (double2*)g_ws[0].get(sizeof(double2) * N * ndat);
  const int zero_im = (pl.istwf_k == 2 && pl.me_g0 == 1) ? 1 : 0;
  if (option != 3) {
    CUDA_CHECK(cudaMemsetAsync(box, 0, sizeof(double2) * N * ndat, st));
    ABI_LAUNCH(k_scatter, dim3(std::max(1, std::min(1024, ceil_div(pl.nent_in, 256))), ndat), dim3(256), 0, st,
               d_fofgin, box, pl.d_in_src, pl.d_in_box, pl.nent_in, pl.npw_in, N);
    g_kernel_launches++;
    fft3d_generic<+1>(box, ndat, n1, n2, n3, st);
  }
  if (option == 0) return;
  if (option == 1) {
    ABI_LAUNCH(k_addtorho, dim3(kNumSM * 8), dim3(256), 0, st, box, d_denpot, d_wr, d_wi, ndat, N);
    g_kernel_launches++;
    return;
  }
  double2* work = box;
  if (option == 3) {
    // the reference transforms out of place and leaves fofr untouched (m_ompgpu_fourwf.F90:268-271)
    work = (double2*)g_ws[0].get(sizeof(double2) * N * ndat);
    CUDA_CHECK(cudaMemcpyAsync(work, d_fofr, sizeof(double2) * N * ndat, cudaMemcpyDeviceToDevice, st));
  }
  if (option == 2) {
    if (cplex == 1) ABI_LAUNCH(k_vlocpsi<1>, dim3(kNumSM * 8, ndat), dim3(256), 0, st, work, d_denpot, N);
    else ABI_LAUNCH(k_vlocpsi<2>, dim3(kNumSM * 8, ndat), dim3(256), 0, st, work, d_denpot, N);
    g_kernel_launches++;
  }
  fft3d_generic<-1>(work, ndat, n1, n2, n3, st);
  ABI_LAUNCH(k_gather, dim3(std::max(1, std::min(1024, ceil_div(pl.npw_out, 256))), ndat), dim3(256), 0, st, work,
             d_fofgout, pl.d_out_box, pl.npw_out, N, 1.0 / (double)N, zero_im);
  g_kernel_launches++;
}

// ---------------------------------------------------------------------------------------------------------
// (B) fused option 2
// ---------------------------------------------------------------------------------------------------------
// K1: sphere lines -> zero-padded x FFT (e^{+i}) -> W1[b][i1][line]
__global__ void __launch_bounds__(256)
k_fw_x_forward(const double2* __restrict__ cg, double2* __restrict__ W1, Fft1d pl, const int2* __restrict__ ent,
               const int* __restrict__ estart, int nlines, int lines_per_cta, int npw, int pack_ndat) {
  ABI_DYN_SMEM(double2, sm);
  const int n = pl.n, ls = n | 1;
  double2* tw = sm;
  double2* buf = sm + n;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * lines_per_cta;
  const int nl = min(lines_per_cta, nlines - l0);
  for (int j = tid; j < n; j += nthr) tw[j] = pl.tw[j];
  for (int w = tid; w < nl * ls; w += nthr) buf[w] = make_double2(0.0, 0.0);
  __syncthreads();
  const int e0 = estart[l0], e1 = estart[l0 + nl];
  if (pack_ndat == 0) {
    const double2* cgb = cg + (size_t)b * npw;
    for (int e = e0 + tid; e < e1; e += nthr) {
      const int2 en = ent[e];
      double2 v = cgb[en.x & 0x3fffffff];
      if (en.x < 0) v.y = -v.y;
      if (en.x & (1 << 30)) v.y = 0.0;
      buf[((en.y >> 10) - l0) * ls + (en.y & 1023)] = v;
    }
  } else {
    // Gamma point, two real-in-r bands per complex transform (the reference's double_rfft_trick,
    // src/66_wfs/m_getghc.F90:1999-2066): E(G) = C(G) + i D(G), E(-G) = conj(C(G)) + i conj(D(G))
    const double2* c0 = cg + (size_t)(2 * b) * npw;
    const bool has_d = 2 * b + 1 < pack_ndat;
    const double2* c1 = c0 + npw;
    for (int e = e0 + tid; e < e1; e += nthr) {
      const int2 en = ent[e];
      const int ipw = en.x & 0x3fffffff;
      double2 c = c0[ipw];
      double2 d = has_d ? c1[ipw] : make_double2(0.0, 0.0);
      if (en.x < 0) { c.y = -c.y; d.y = -d.y; }
      if (en.x & (1 << 30)) { c.y = 0.0; d.y = 0.0; }
      buf[((en.y >> 10) - l0) * ls + (en.y & 1023)] = make_double2(c.x - d.y, c.y + d.x);
    }
  }
  __syncthreads();
  fft_lines_dit<+1>(buf, ls, nl, pl, tw, tid, nthr);
  double2* out = W1 + (size_t)b * n * nlines + l0;
  {
    // w = tid + k nthr -> (i1, l) = (w / nl, w % nl) by increments: one division per thread instead of one per element
    int i1 = tid / nl, l = tid - i1 * nl;
    const int di = nthr / nl, dl = nthr - di * nl;
#pragma unroll 4
    for (int w = tid; w < nl * n; w += nthr) {
      out[(size_t)i1 * nlines + l] = buf[l * ls + i1];
      i1 += di; l += dl;
      if (l >= nl) { l -= nl; i1++; }
    }
  }
}

// K3: W1out[b][i1][line] -> x FFT (e^{-i}) -> gather to the sphere * xnorm (+ fused getghc assembly)
__global__ void __launch_bounds__(256)
k_fw_x_backward(const double2* __restrict__ W1o, double2* __restrict__ outg, Fft1d pl, const int2* __restrict__ ent,
                const int* __restrict__ estart, int nlines, int lines_per_cta, int npw, double xnorm,
                int zero_im_g0, FourwfEpilogue epi, double kin_filter) {
  ABI_DYN_SMEM(double2, sm);
  const int n = pl.n, ls = n | 1;
  double2* tw = sm;
  double2* buf = sm + n;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * lines_per_cta;
  const int nl = min(lines_per_cta, nlines - l0);
  for (int j = tid; j < n; j += nthr) tw[j] = pl.tw[j];
  const double2* in = W1o + (size_t)b * n * nlines + l0;
  {
    // (i1, l) = (w / nl, w % nl) by increments; 8 independent 16-byte loads in flight per thread
    int i1 = tid / nl, l = tid - i1 * nl;
    const int di = nthr / nl, dl = nthr - di * nl;
#pragma unroll 8
    for (int w = tid; w < nl * n; w += nthr) {
      buf[l * ls + i1] = ldg2(in + (size_t)i1 * nlines + l);
      i1 += di; l += dl;
      if (l >= nl) { l -= nl; i1++; }
    }
  }
  __syncthreads();
  fft_lines_dif<-1>(buf, ls, nl, pl, tw, tid, nthr);
  const int e0 = estart[l0], e1 = estart[l0 + nl];
  for (int e = e0 + tid; e < e1; e += nthr) {
    const int2 en = ent[e];
    const int ipw = en.x;
    double2 v = buf[((en.y >> 10) - l0) * ls + (en.y & 1023)];
    v.x *= xnorm; v.y *= xnorm;
    if (zero_im_g0 && ipw == 0) v.y = 0.0;
    const size_t o = (size_t)b * npw + ipw;
    if (!fw_epilogue(epi, kin_filter, ipw, o, v)) v = make_double2(0.0, 0.0);
    outg[o] = v;
  }
}

// K3 for the packed Gamma-point transform: F = H C + i H D on the full sphere;
//   H C(G) = [F(G) + conj(F(-G))]/2,  H D(G) = [F(G) - conj(F(-G))]/(2i)   (cwavef_double_rfft_trick_unpack,
// src/66_wfs/m_getghc.F90:2097-2171).  G and -G sit on different lines (different CTAs), so each of the two
// contributions is added with one red.global per component onto a zero-initialised output: two addends commute
// exactly, the result is deterministic.  The getghc assembly rides on the direct (G) contribution.
__global__ void __launch_bounds__(256)
k_fw_x_backward_packed(const double2* __restrict__ W1o, double2* __restrict__ outg, Fft1d pl, const int2* __restrict__ ent,
                       const int* __restrict__ estart, int nlines, int lines_per_cta, int npw, int ndat, double xnorm,
                       FourwfEpilogue epi, double kin_filter) {
  ABI_DYN_SMEM(double2, sm);
  const int n = pl.n, ls = n | 1;
  double2* tw = sm;
  double2* buf = sm + n;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * lines_per_cta;
  const int nl = min(lines_per_cta, nlines - l0);
  for (int j = tid; j < n; j += nthr) tw[j] = pl.tw[j];
  const double2* in = W1o + (size_t)b * n * nlines + l0;
  {
    // (i1, l) = (w / nl, w % nl) by increments; 8 independent 16-byte loads in flight per thread
    int i1 = tid / nl, l = tid - i1 * nl;
    const int di = nthr / nl, dl = nthr - di * nl;
#pragma unroll 8
    for (int w = tid; w < nl * n; w += nthr) {
      buf[l * ls + i1] = ldg2(in + (size_t)i1 * nlines + l);
      i1 += di; l += dl;
      if (l >= nl) { l -= nl; i1++; }
    }
  }
  __syncthreads();
  fft_lines_dif<-1>(buf, ls, nl, pl, tw, tid, nthr);
  const int e0 = estart[l0], e1 = estart[l0 + nl];
  const bool has_d = 2 * b + 1 < ndat;
  const double h = 0.5 * xnorm;
  for (int e = e0 + tid; e < e1; e += nthr) {
    const int2 en = ent[e];
    const int ipw = en.x & 0x3fffffff;
    const double2 f = buf[((en.y >> 10) - l0) * ls + (en.y & 1023)];
    const size_t oc = (size_t)(2 * b) * npw + ipw, od = oc + npw;
    double2 vc, vd;
    if (en.x & (1 << 30)) {               // G = 0: H C(0) = Re F(0), H D(0) = Im F(0), both real; no image entry
      vc = make_double2(f.x * xnorm, 0.0); vd = make_double2(f.y * xnorm, 0.0);
    } else if (en.x >= 0) {               // direct entry: + F/2 and -i F/2
      vc = make_double2(f.x * h, f.y * h); vd = make_double2(f.y * h, -f.x * h);
    } else {                              // image entry holds F(-G): + conj(F)/2 and + i conj(F)/2
      vc = make_double2(f.x * h, -f.y * h); vd = make_double2(f.y * h, f.x * h);
    }
    if (en.x >= 0) {                      // assembly terms once per coefficient
      if (!fw_epilogue(epi, kin_filter, ipw, oc, vc)) vc = make_double2(0.0, 0.0);
      if (has_d && !fw_epilogue(epi, kin_filter, ipw, od, vd)) vd = make_double2(0.0, 0.0);
    } else if (epi.mode != 0) {
      const double k = epi.kinpw[ipw];
      if ((epi.mode == 1 && !(k < kin_filter)) || (epi.mode == 2 && k > kin_filter)) { vc = make_double2(0.0, 0.0); vd = vc; }
    }
#ifndef ABI_EMU
    atomicAdd(&outg[oc].x, vc.x); atomicAdd(&outg[oc].y, vc.y);
    if (has_d) { atomicAdd(&outg[od].x, vd.x); atomicAdd(&outg[od].y, vd.y); }
#else
    outg[oc].x += vc.x; outg[oc].y += vc.y;
    if (has_d) { outg[od].x += vd.x; outg[od].y += vd.y; }
#endif
  }
}

struct MidParams {
  int n1, n2, n3, nb, nlin, nlout, nU, cplex, lb, csize, nclusters;
  const double2* W1; double2* W1o; double2* scratch; const double* vT;
  const int* inpl_start; const int* lin_u; const unsigned short* lin_pos2;
  const int* outpl_start; const int* lout_u; const unsigned short* lout_pos2;
  const unsigned short* u_i3; const unsigned char* u_flags;
  Fft1d p2, p3;
};

template <bool CLUSTER>
ABI_DEV void mid_sync() {
#ifndef ABI_EMU
  if (CLUSTER) cg::this_cluster().sync();
  else __syncthreads();
#endif
}

// K2: one (band, i1) yz-plane per cluster iteration
template <bool CLUSTER>
__global__ void __launch_bounds__(256)
k_fw_mid(MidParams P) {
  ABI_DYN_SMEM(double2, sm);
  const int n2 = P.n2, n3 = P.n3, ls2 = n2 | 1, ls3 = n3 | 1;
  double2* tw2 = sm;
  double2* tw3 = tw2 + n2;
  double2* buf = tw3 + n3;
  unsigned short* idx3 = reinterpret_cast<unsigned short*>(buf + (size_t)P.lb * max(ls2, ls3));
  unsigned short* ui3 = idx3 + n3;
  unsigned char* ufl = reinterpret_cast<unsigned char*>(ui3 + P.nU);
  const int tid = threadIdx.x, nthr = blockDim.x;
  int rank = 0, cid = blockIdx.x;
#ifndef ABI_EMU
  if (CLUSTER) { rank = (int)cg::this_cluster().block_rank(); cid = blockIdx.x / P.csize; }
#endif
  const int cs = CLUSTER ? P.csize : 1;
  for (int j = tid; j < n2; j += nthr) tw2[j] = P.p2.tw[j];
  for (int j = tid; j < n3; j += nthr) { tw3[j] = P.p3.tw[j]; idx3[j] = P.p3.idx_of_pos[j]; }
  for (int j = tid; j < P.nU; j += nthr) { ui3[j] = P.u_i3[j]; ufl[j] = P.u_flags[j]; }
  __syncthreads();
  const int nU = P.nU, lb = P.lb;
  const int u_lo = (int)((long long)rank * nU / cs), u_hi = (int)((long long)(rank + 1) * nU / cs);
  const int j_lo = (int)((long long)rank * n2 / cs), j_hi = (int)((long long)(rank + 1) * n2 / cs);
  const long long nunits = (long long)P.nb * P.n1;
  int it = 0;
  for (long long unit = cid; unit < nunits; unit += P.nclusters, it++) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    double2* S = P.scratch + ((size_t)cid * 2 + (it & 1)) * (size_t)nU * n2;
    const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
    double2* w1o = P.W1o + ((size_t)b * P.n1 + i1) * P.nlout;
    // ---- y pass (e^{+i}) on my share of occupied z planes ----
    for (int u0 = u_lo; u0 < u_hi; u0 += lb) {
      const int nl = min(lb, u_hi - u0);
      for (int w = tid; w < nl * ls2; w += nthr) buf[w] = make_double2(0.0, 0.0);
      __syncthreads();
      const int q0 = P.inpl_start[u0], q1 = P.inpl_start[u0 + nl];
      for (int q = q0 + tid; q < q1; q += nthr) buf[(P.lin_u[q] - u0) * ls2 + P.lin_pos2[q]] = w1[q];
      __syncthreads();
      fft_lines_dit<+1>(buf, ls2, nl, P.p2, tw2, tid, nthr);
      for (int w = tid; w < nl * n2; w += nthr) {
        const int l = w / n2, i2 = w - l * n2;
        stcg2(&S[(size_t)(u0 + l) * n2 + i2], buf[l * ls2 + i2]);
      }
      __syncthreads();
    }
    mid_sync<CLUSTER>();
    // ---- z pass, V_loc, inverse z on my share of y columns ----
    for (int j0 = j_lo; j0 < j_hi; j0 += lb) {
      const int nl = min(lb, j_hi - j0);
      for (int w = tid; w < nl * ls3; w += nthr) buf[w] = make_double2(0.0, 0.0);
      __syncthreads();
      for (int w = tid; w < nl * nU; w += nthr) {
        const int u = w / nl, l = w - u * nl;
        if (ufl[u] & 1) buf[l * ls3 + ui3[u]] = ldcg2(&S[(size_t)u * n2 + j0 + l]);
      }
      __syncthreads();
      fft_lines_dif<+1>(buf, ls3, nl, P.p3, tw3, tid, nthr);
      if (P.cplex == 1) {
        const double* v = P.vT + (size_t)i1 * n3 * n2 + j0;
        for (int w = tid; w < nl * n3; w += nthr) {
          const int p = w / nl, l = w - p * nl;
          const double vv = v[(size_t)idx3[p] * n2 + l];
          double2 x = buf[l * ls3 + p];
          x.x *= vv; x.y *= vv;
          buf[l * ls3 + p] = x;
        }
      } else {
        const double2* v = reinterpret_cast<const double2*>(P.vT) + (size_t)i1 * n3 * n2 + j0;
        for (int w = tid; w < nl * n3; w += nthr) {
          const int p = w / nl, l = w - p * nl;
          buf[l * ls3 + p] = cmul(buf[l * ls3 + p], v[(size_t)idx3[p] * n2 + l]);
        }
      }
      __syncthreads();
      fft_lines_dit<-1>(buf, ls3, nl, P.p3, tw3, tid, nthr);
      for (int w = tid; w < nl * nU; w += nthr) {
        const int u = w / nl, l = w - u * nl;
        if (ufl[u] & 2) stcg2(&S[(size_t)u * n2 + j0 + l], buf[l * ls3 + ui3[u]]);
      }
      __syncthreads();
    }
    mid_sync<CLUSTER>();
    // ---- inverse y pass (e^{-i}) on my share of planes, keep the output lines ----
    for (int u0 = u_lo; u0 < u_hi; u0 += lb) {
      const int nl = min(lb, u_hi - u0);
      const int q0 = P.outpl_start[u0], q1 = P.outpl_start[u0 + nl];
      if (q1 > q0) {
        for (int w = tid; w < nl * n2; w += nthr) {
          const int l = w / n2, i2 = w - l * n2;
          buf[l * ls2 + i2] = ldcg2(&S[(size_t)(u0 + l) * n2 + i2]);
        }
        __syncthreads();
        fft_lines_dif<-1>(buf, ls2, nl, P.p2, tw2, tid, nthr);
        for (int q = q0 + tid; q < q1; q += nthr) w1o[q] = buf[(P.lout_u[q] - u0) * ls2 + P.lout_pos2[q]];
        __syncthreads();
      }
    }
  }
}

static size_t mid_smem_bytes(int n2, int n3, int nU, int lb) {
  const int ls = std::max(n2 | 1, n3 | 1);
  size_t s = sizeof(double2) * ((size_t)n2 + n3 + (size_t)lb * ls);
  s += sizeof(unsigned short) * ((size_t)n3 + nU) + nU;
  return (s + 15) & ~(size_t)15;
}

void fourwf_fused_opt2(const FourwfPlan& pl, const VlocDev& v, const double2* d_fofgin, double2* d_fofgout, int ndat,
                       const FourwfEpilogue& epi, cudaStream_t st) {
  ABI_CHECK(pl.fused_ok, "fused fourwf path not available for this FFT box");
  ABI_CHECK(v.n1 == pl.n1 && v.n2 == pl.n2 && v.n3 == pl.n3, "vlocal dimensions do not match ngfft (augmented grids are rejected)");
  const int n1 = pl.n1, n2 = pl.n2, n3 = pl.n3;
  const FftTables& t1 = fft_tables(n1);
  const FftTables& t2 = fft_tables(n2);
  const FftTables& t3 = fft_tables(n3);
  FourwfTuning& tune = fourwf_tuning();

  // ---- K2 geometry ----
  int cs = tune.cluster;
#ifdef ABI_EMU
  cs = 1;
#else
  if (cs <= 0) cs = 4;
#endif
  size_t smem_budget = (size_t)(tune.smem_kb_mid > 0 ? tune.smem_kb_mid : 100) * 1024;
  int lb = 1;
  {
    const int maxl = std::max(ceil_div(pl.nU, cs), ceil_div(n2, cs));
    while (lb < maxl && mid_smem_bytes(n2, n3, pl.nU, lb + 1) <= smem_budget) lb++;
  }
  const size_t smem_mid = mid_smem_bytes(n2, n3, pl.nU, lb);
  ABI_CHECK(smem_mid <= kMaxSmemPerCta, "FFT box too large for the shared-memory plane stage");
  const int ctas_per_sm = std::max<int>(1, std::min<int>(2, (int)(kMaxSmemPerCta / smem_mid)));
  int nclusters = std::max(1, (kNumSM * ctas_per_sm) / cs);

  // Gamma point: two bands per complex transform (needs a real potential and identical in/out spheres)
  const bool pack2 = tune.pack2 && pl.istwf_k == 2 && v.cplex == 1 && pl.same_sphere && ndat >= 2 && pl.plane_ok && tune.plane;
  const int nlout_eff = pack2 ? pl.nlin : pl.nlout;       // packed: the output lines are the (completed) input lines
  const int ntrans = pack2 ? (ndat + 1) / 2 : ndat;       // transforms to run

  const bool use_half = pl.plane_ok && tune.plane && !tune.plane_split && half_stage_usable(pl, pack2);
  // ---- band chunking bounds the workspace (W1, W1', scratch) ----
  const bool split_plane = pl.plane_ok && tune.plane && (n2 != n3 || tune.plane_split);          // S planes of the chunk live in global memory
  const size_t per_band = sizeof(double2) * (size_t)n1 * ((size_t)pl.nlin + nlout_eff + (split_plane ? (size_t)pl.nU * n2 : 0));
  int chunk = tune.band_chunk > 0 ? tune.band_chunk : (int)std::max<size_t>(1, ((size_t)3 << 30) / per_band);
  chunk = std::min(chunk, ntrans);
  double2* W1 = (double2*)g_ws[1].get(sizeof(double2) * (size_t)n1 * pl.nlin * chunk);
  double2* W1o = (double2*)g_ws[2].get(sizeof(double2) * (size_t)n1 * nlout_eff * chunk);
  double2* scratch = (pl.plane_ok && tune.plane) ? nullptr
                                                 : (double2*)g_ws[3].get(sizeof(double2) * (size_t)nclusters * 2 * pl.nU * n2);

  int lx = std::max(1, tune.lines_x);
  while (lx > 1 && sizeof(double2) * ((size_t)n1 + (size_t)lx * (n1 | 1)) > 110 * 1024) lx--;
  const size_t smem_x = sizeof(double2) * ((size_t)n1 + (size_t)lx * (n1 | 1));
#ifndef ABI_EMU
  CUDA_CHECK(cudaFuncSetAttribute(k_fw_x_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
  CUDA_CHECK(cudaFuncSetAttribute(k_fw_x_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
  CUDA_CHECK(cudaFuncSetAttribute(k_fw_x_backward_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
  CUDA_CHECK(cudaFuncSetAttribute(k_fw_mid<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mid));
  CUDA_CHECK(cudaFuncSetAttribute(k_fw_mid<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mid));
#endif
  const double xnorm = 1.0 / ((double)n1 * n2 * n3);
  const int zero_im = (pl.istwf_k == 2 && pl.me_g0 == 1) ? 1 : 0;
  const double kin_filter = 1.7976931348623157e308 * 1.0e-11;   // huge(0d0)*1d-11, m_getghc.F90:1272

  const bool use_xh = x_stage_usable(pl, pack2);
  if (pack2 && !(use_xh && (tune.xhalf & 2))) CUDA_CHECK(cudaMemsetAsync(d_fofgout, 0, sizeof(double2) * (size_t)pl.npw_out * ndat, st));
  for (int t0 = 0; t0 < ntrans; t0 += chunk) {
    const int nb = std::min(chunk, ntrans - t0);            // transforms in this chunk
    const int b0 = pack2 ? 2 * t0 : t0;                     // first band of the chunk
    const int nbands = std::min(ndat - b0, pack2 ? 2 * nb : nb);
    { ProfScope ps("fourwf_x_forward");
    if (use_xh && (tune.xhalf & 1)) x_stage_forward(pl, d_fofgin + (size_t)b0 * pl.npw_in, W1, nb, pack2 ? nbands : 0, st);
    else ABI_LAUNCH(k_fw_x_forward, dim3(ceil_div(pl.nlin, lx), nb), dim3(256), smem_x, st,
               d_fofgin + (size_t)b0 * pl.npw_in, W1, t1.plan, pl.d_in_ent, pl.d_lin_estart, pl.nlin, lx, pl.npw_in,
               pack2 ? nbands : 0); }
    MidParams P;
    P.n1 = n1; P.n2 = n2; P.n3 = n3; P.nb = nb; P.nlin = pl.nlin; P.nlout = pl.nlout; P.nU = pl.nU;
    P.cplex = v.cplex; P.lb = lb; P.csize = cs; P.nclusters = nclusters;
    P.W1 = W1; P.W1o = W1o; P.scratch = scratch; P.vT = v.d_vT;
    P.inpl_start = pl.d_inpl_start; P.lin_u = pl.d_lin_u; P.lin_pos2 = pl.d_lin_pos2;
    P.outpl_start = pl.d_outpl_start; P.lout_u = pl.d_lout_u; P.lout_pos2 = pl.d_lout_pos2;
    P.u_i3 = pl.d_u_i3; P.u_flags = pl.d_u_flags; P.p2 = t2.plan; P.p3 = t3.plan;
    if (getenv("ABI_B200_DEBUG")) fprintf(stderr, "abinit_b200: fourwf fused: plane_ok=%d tune.plane=%d half=%d (half_ok=%d/%d amb=%d/%d) nU=%d z=[%d,+%d)U[%d,+%d)\n", (int)pl.plane_ok, tune.plane, (int)use_half, (int)pl.half_ok_in, (int)pl.half_ok_out, pl.y_amb_in, pl.y_amb_out, pl.nU, pl.za, pl.zla, pl.zb, pl.zlb);
    if (use_half) {
      ProfScope ps("fourwf_plane_stage");
      HalfLaunch L;
      L.nb = nb; L.W1 = W1; L.W1o = W1o; L.nlin = pl.nlin; L.nlout = nlout_eff; L.out_is_in = pack2;
      half_stage_launch(pl, v, L, st);
    } else if (pl.plane_ok && tune.plane) {
      ProfScope ps("fourwf_plane_stage");
      PlaneParams Q;
      Q.n1 = n1; Q.n2 = n2; Q.n3 = n3; Q.nb = nb; Q.nU = pl.nU; Q.cplex = v.cplex; Q.za = pl.za; Q.zla = pl.zla; Q.zb = pl.zb; Q.zlb = pl.zlb;
      Q.nlin = pl.nlin; Q.nlout = nlout_eff; Q.nunits = (long long)nb * n1;
      Q.W1 = W1; Q.W1o = W1o; Q.S = nullptr; Q.vT = v.d_vT; Q.tw = t2.plan.tw; Q.tw3 = t3.plan.tw;
      Q.in_start = pl.d_pin_start; Q.in_runs = pl.d_pin_runs;
      Q.out_start = pack2 ? pl.d_pin_start : pl.d_pout_start; Q.out_runs = pack2 ? pl.d_pin_runs : pl.d_pout_runs;
      plane_stage_launch(Q, st);
    } else
    { ProfScope ps("fourwf_plane_cluster");
#ifdef ABI_EMU
    ABI_LAUNCH(k_fw_mid<false>, dim3(nclusters), dim3(256), smem_mid, st, P);
#else
    if (cs == 1) {
      ABI_LAUNCH(k_fw_mid<false>, dim3(nclusters), dim3(256), smem_mid, st, P);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nclusters * cs); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem_mid; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_fw_mid<true>, P));
    }
#endif
    }
    FourwfEpilogue e = epi;
    if (e.cwavef) e.cwavef += (size_t)b0 * pl.npw_out;
    if (e.gvnlxc) e.gvnlxc += (size_t)b0 * pl.npw_out;
    if (e.gsc) e.gsc += (size_t)b0 * pl.npw_out;
    { ProfScope ps("fourwf_x_backward");
    if (use_xh && (tune.xhalf & 2)) {
      x_stage_backward(pl, W1o, d_fofgout + (size_t)b0 * pl.npw_out, nb, pack2 ? nbands : 0, xnorm, zero_im, e, kin_filter, st);
    } else if (pack2) {
      ABI_LAUNCH(k_fw_x_backward_packed, dim3(ceil_div(pl.nlin, lx), nb), dim3(256), smem_x, st, W1o,
                 d_fofgout + (size_t)b0 * pl.npw_out, t1.plan, pl.d_in_ent, pl.d_lin_estart, pl.nlin, lx, pl.npw_out, nbands,
                 xnorm, e, kin_filter);
    } else {
      ABI_LAUNCH(k_fw_x_backward, dim3(ceil_div(pl.nlout, lx), nb), dim3(256), smem_x, st, W1o,
                 d_fofgout + (size_t)b0 * pl.npw_out, t1.plan, pl.d_out_ent, pl.d_lout_estart, pl.nlout, lx, pl.npw_out,
                 xnorm, zero_im, e, kin_filter);
    } }
    g_kernel_launches += split_plane ? 5 : 3;
  }
}

// ---------------------------------------------------------------------------------------------------------
// (C) fused option 1: rho(r) += weight_r Re(psi(r))^2 + weight_i Im(psi(r))^2  (m_fft.F90:2633-2653)
//     K1 (x pass on the occupied lines) -> plane stage (y pass, z pass, density reduction into rhoT[i1][i3][i2]) ->
//     one transpose-add into the caller's denpot.  No full box per band ever exists.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_rho_weights(double2* __restrict__ wxy, const double* __restrict__ wr, const double* __restrict__ wi, int ntrans,
                              int b0, int ndat, int pack2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntrans) return;
  if (pack2) {
    // E = C + i D: Re E(r) = C(r), Im E(r) = D(r), both real in r at Gamma (their own imaginary parts vanish)
    const int a = b0 + 2 * t, b = a + 1;
    wxy[t] = make_double2(wr[a], b < ndat ? wr[b] : 0.0);
  } else {
    wxy[t] = make_double2(wr[b0 + t], wi[b0 + t]);
  }
}

__global__ void k_rho_untranspose_add(const double* __restrict__ rhoT, double* __restrict__ rho, int n1, int n2, int n3) {
  // rho[i3][i2][i1] += rhoT[i1][i3][i2] through a 32x33 shared tile over (i1, i2) for fixed i3
#ifndef ABI_EMU
  __shared__ double tile[32][33];
  const int i3 = blockIdx.z;
  const int i1b = blockIdx.x * 32, i2b = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i1 = i1b + r, i2 = i2b + threadIdx.x;
    tile[r][threadIdx.x] = (i1 < n1 && i2 < n2) ? rhoT[((size_t)i1 * n3 + i3) * n2 + i2] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i2 = i2b + r, i1 = i1b + threadIdx.x;
    if (i1 < n1 && i2 < n2) rho[((size_t)i3 * n2 + i2) * n1 + i1] += tile[threadIdx.x][r];
  }
#else
  for (int i1 = 0; i1 < n1; i1++) for (int i3 = 0; i3 < n3; i3++) for (int i2 = 0; i2 < n2; i2++)
    rho[((size_t)i3 * n2 + i2) * n1 + i1] += rhoT[((size_t)i1 * n3 + i3) * n2 + i2];
#endif
}

bool fourwf_fused_opt1_available(const FourwfPlan& pl) {
  return pl.fused_ok && pl.plane_ok && fourwf_tuning().plane && plane_stage_supported(pl.n2) && plane_stage_supported(pl.n3);
}

void fourwf_fused_opt1(const FourwfPlan& pl, const double2* d_fofgin, double* d_denpot, int ndat, const double* d_wr,
                       const double* d_wi, cudaStream_t st) {
  ABI_CHECK(fourwf_fused_opt1_available(pl), "fused fourwf option 1 not available for this FFT box");
  const int n1 = pl.n1, n2 = pl.n2, n3 = pl.n3;
  const size_t N = (size_t)n1 * n2 * n3;
  const FftTables& t1 = fft_tables(n1);
  const FftTables& t2 = fft_tables(n2);
  const FftTables& t3 = fft_tables(n3);
  FourwfTuning& tune = fourwf_tuning();
  const bool pack2 = tune.pack2 && pl.istwf_k == 2 && ndat >= 2;
  const int ntrans = pack2 ? (ndat + 1) / 2 : ndat;
  const size_t per_band = sizeof(double2) * (size_t)n1 * ((size_t)pl.nlin + (n2 != n3 ? (size_t)pl.nU * n2 : 0));
  int chunk = tune.band_chunk > 0 ? tune.band_chunk : (int)std::max<size_t>(1, ((size_t)3 << 30) / per_band);
  chunk = std::min(chunk, ntrans);
  double2* W1 = (double2*)g_ws[1].get(sizeof(double2) * (size_t)n1 * pl.nlin * chunk);
  // half-support plane stage: the density is accumulated in the register order of its z pass (rhoP) and un-permuted at the end
  const bool use_half = half_stage_usable(pl, true);
  const bool use_xh = x_stage_usable(pl, pl.istwf_k == 2 && pl.same_sphere) && (tune.xhalf & 1);
  const size_t nrho = use_half ? half_rho_elems(pl) : N;
  // rhoT / rhoP (nrho doubles) followed by the per-transform weights
  double2* wxy = (double2*)g_ws[2].get(sizeof(double2) * (size_t)ntrans + sizeof(double) * nrho);
  double* rhoT = reinterpret_cast<double*>(wxy + ntrans);
  CUDA_CHECK(cudaMemsetAsync(rhoT, 0, sizeof(double) * nrho, st));
  int lx = std::max(1, tune.lines_x);
  while (lx > 1 && sizeof(double2) * ((size_t)n1 + (size_t)lx * (n1 | 1)) > 110 * 1024) lx--;
  const size_t smem_x = sizeof(double2) * ((size_t)n1 + (size_t)lx * (n1 | 1));
#ifndef ABI_EMU
  CUDA_CHECK(cudaFuncSetAttribute(k_fw_x_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
#endif
  for (int t0 = 0; t0 < ntrans; t0 += chunk) {
    const int nb = std::min(chunk, ntrans - t0);
    const int b0 = pack2 ? 2 * t0 : t0;
    const int nbands = std::min(ndat - b0, pack2 ? 2 * nb : nb);
    ABI_LAUNCH(k_rho_weights, dim3(ceil_div(nb, 128)), dim3(128), 0, st, wxy, d_wr, d_wi, nb, b0, ndat, pack2 ? 1 : 0);
    { ProfScope ps("fourwf_x_forward");
    if (use_xh) x_stage_forward(pl, d_fofgin + (size_t)b0 * pl.npw_in, W1, nb, pack2 ? nbands : 0, st);
    else ABI_LAUNCH(k_fw_x_forward, dim3(ceil_div(pl.nlin, lx), nb), dim3(256), smem_x, st, d_fofgin + (size_t)b0 * pl.npw_in, W1,
               t1.plan, pl.d_in_ent, pl.d_lin_estart, pl.nlin, lx, pl.npw_in, pack2 ? nbands : 0); }
    { ProfScope ps("fourwf_plane_rho");
    if (use_half) {
      HalfLaunch L;
      L.nb = nb; L.W1 = W1; L.W1o = nullptr; L.nlin = pl.nlin; L.nlout = pl.nlin; L.out_is_in = true; L.rhoP = rhoT; L.wxy = wxy;
      half_stage_launch_rho(pl, L, st);
    } else {
    PlaneParams Q;
    Q.n1 = n1; Q.n2 = n2; Q.n3 = n3; Q.nb = nb; Q.nU = pl.nU; Q.cplex = 1; Q.za = pl.za; Q.zla = pl.zla; Q.zb = pl.zb; Q.zlb = pl.zlb;
    Q.nlin = pl.nlin; Q.nlout = pl.nlin; Q.nunits = (long long)nb * n1;
    Q.W1 = W1; Q.W1o = nullptr; Q.S = nullptr; Q.vT = nullptr; Q.tw = t2.plan.tw; Q.tw3 = t3.plan.tw;
    Q.in_start = pl.d_pin_start; Q.in_runs = pl.d_pin_runs; Q.out_start = pl.d_pin_start; Q.out_runs = pl.d_pin_runs;
    Q.rhoT = rhoT; Q.wxy = wxy;
    plane_stage_launch_rho(Q, st); } }
    g_kernel_launches += (n2 != n3) ? 4 : 3;
  }
  if (use_half) {
    half_rho_unpermute_add(pl, rhoT, d_denpot, st);
  } else {
#ifndef ABI_EMU
  k_rho_untranspose_add<<<dim3(ceil_div(n1, 32), ceil_div(n2, 32), n3), dim3(32, 8), 0, st>>>(rhoT, d_denpot, n1, n2, n3);
  CUDA_CHECK(cudaGetLastError());
#else
  k_rho_untranspose_add(rhoT, d_denpot, n1, n2, n3);
#endif
  g_kernel_launches++;
  }
}

}  // namespace abi
