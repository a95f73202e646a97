// Host side of the half-support x passes (x_stage.cuh): per-plan slot / batch tables and the launchers.
#include "x_stage.cuh"
#include "context.cuh"
#include <map>
#include <vector>
#include <unordered_map>

namespace abi {

template <int A, int B> void xh_launch(int dir, XhParams& P, cudaStream_t st);   // x_stage_impl.cuh
extern const int kXhGLHost;

struct XhSet {                 // tables of one line set (input lines, or output lines of the unpacked path)
  bool ok = false;
  int nlines = 0, nbatch1 = 0, nbatch3 = 0;
  int* d_sign = nullptr; int* d_ovslot = nullptr;
  int2* d_ent = nullptr;       // K1 entries (input set only)
  int4* d_batches = nullptr; int2* d_oent = nullptr; int* d_bstart = nullptr;    // K3
};
struct XhTabs {
  bool built = false;
  XhSet in_set;                // K1, and K3 of the packed Gamma path (mirror-paired batches)
  XhSet in_plain;              // K3 on the input lines without pairing is never needed; kept empty
  XhSet out_set;               // K3 of the plain path (output lines)
  std::vector<void*> owned;
};
void xh_tabs_free(XhTabs* t) {
  if (!t) return;
  for (void* p : t->owned) cudaFree(p);
  delete t;
}

namespace {
typedef void (*XhFn)(int, XhParams&, cudaStream_t);
struct Entry { int n, A, B; XhFn fn; };
#define XH_ENTRY(A, B) {2 * (A) * (B), A, B, &xh_launch<A, B>}
const Entry kEntries[] = {
    XH_ENTRY(4, 3), XH_ENTRY(5, 3), XH_ENTRY(4, 4), XH_ENTRY(9, 2), XH_ENTRY(4, 5), XH_ENTRY(8, 3), XH_ENTRY(4, 8), XH_ENTRY(8, 8),
    XH_ENTRY(9, 10),
};
const Entry* find_entry(int n) {
  for (const Entry& e : kEntries) if (e.n == n) return &e;
  return nullptr;
}
int gcd_rt(int a, int b) { return b == 0 ? a : gcd_rt(b, a % b); }
int rin_rt(int A, int B, int t, int j) { return gcd_rt(A, B) == 1 ? (B * t + A * j) % (A * B) : j + B * t; }
template <typename T> T* upload(const std::vector<T>& v, std::vector<void*>& owned) {
  T* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  owned.push_back(d);
  return d;
}

// slot of every entry of a line set: q(r) for i1 = r (or i1 = r + m where r never occurs as a low index), m + k for the high
// partner of the k-th index r that occurs both ways
bool assign_slots(const FourwfPlan& pl, const Entry& e, const std::vector<int2>& ent, std::vector<int>& slot, std::vector<int>& sign,
                  std::vector<int>& ovslot) {
  const int n1 = pl.n1, M = n1 / 2;
  const FftTables& t1 = fft_tables(n1);
  std::vector<char> has_lo(M, 0), has_hi(M, 0);
  std::vector<int> i1s(ent.size());
  for (size_t k = 0; k < ent.size(); k++) {
    const int i1 = t1.idx_of_pos[ent[k].y & 1023];
    i1s[k] = i1;
    if (i1 < M) has_lo[i1] = 1; else has_hi[i1 - M] = 1;
  }
  std::vector<int> q_of_r(M, 0), amb(M, -1);
  for (int t = 0; t < e.A; t++) for (int j = 0; j < e.B; j++) q_of_r[rin_rt(e.A, e.B, t, j)] = t * e.B + j;
  int nov = 0;
  sign.assign(M, 1); ovslot.assign(M, -1);
  for (int r = 0; r < M; r++) {
    if (has_lo[r] && has_hi[r]) { if (nov >= kHalfOV) return false; amb[r] = M + nov; ovslot[q_of_r[r]] = M + nov; nov++; }
    else if (has_hi[r]) sign[q_of_r[r]] = -1;
  }
  slot.resize(ent.size());
  for (size_t k = 0; k < ent.size(); k++) {
    const int i1 = i1s[k];
    slot[k] = i1 < M ? q_of_r[i1] : (amb[i1 - M] >= 0 ? amb[i1 - M] : q_of_r[i1 - M]);
  }
  return true;
}

void build_set(const FourwfPlan& pl, const Entry& e, XhTabs& tabs, XhSet& s, const std::vector<int2>& ent, const std::vector<int>& estart,
               bool with_k1, bool paired) {
  const int GL = kXhGLHost;
  s.nlines = (int)estart.size() - 1;
  std::vector<int> slot, sign, ovslot;
  if (s.nlines <= 0 || !assign_slots(pl, e, ent, slot, sign, ovslot)) { s.ok = false; return; }
  s.d_sign = upload(sign, tabs.owned); s.d_ovslot = upload(ovslot, tabs.owned);
  if (with_k1) {
    std::vector<int2> e2(ent.size());
    for (size_t k = 0; k < ent.size(); k++) e2[k] = make_int2(ent[k].x, ((ent[k].y >> 10) << 10) | slot[k]);
    s.d_ent = upload(e2, tabs.owned);
    s.nbatch1 = (s.nlines + GL - 1) / GL;
  }
  // ---- K3 batches ----
  std::vector<int4> batches; std::vector<int2> oent; std::vector<int> bstart;
  auto push_int4 = [](int a, int b, int c, int d) { int4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; };
  if (!paired) {
    for (int l0 = 0; l0 < s.nlines; l0 += GL) {
      const int ka = std::min(GL, s.nlines - l0);
      bstart.push_back((int)oent.size());
      batches.push_back(push_int4(l0, ka, 0, 0));
      for (int k = estart[l0]; k < estart[l0 + ka]; k++)
        oent.push_back(make_int2(ent[k].x & 0x3fffffff, slot[k] | (((ent[k].y >> 10) - l0) << 8)));
    }
  } else {
    // mirror line of (i2, i3) is (-i2, -i3); image entries carry bit 31 of src
    const int n2 = pl.n2, n3 = pl.n3;
    std::unordered_map<long long, int> line_of;
    for (int l = 0; l < s.nlines; l++) line_of[(long long)pl.h_lin_i2i3[l].y * n2 + pl.h_lin_i2i3[l].x] = l;
    std::vector<int> mirror(s.nlines, -1);
    for (int l = 0; l < s.nlines; l++) {
      const int i2 = (n2 - pl.h_lin_i2i3[l].x) % n2, i3 = (n3 - pl.h_lin_i2i3[l].y) % n3;
      auto it = line_of.find((long long)i3 * n2 + i2);
      if (it != line_of.end()) mirror[l] = it->second;
    }
    // image entry (line, slot) of every plane wave
    std::vector<int> img_line(pl.npw_in, -1), img_slot(pl.npw_in, 0);
    for (size_t k = 0; k < ent.size(); k++) if (ent[k].x < 0) { const int ipw = ent[k].x & 0x3fffffff; img_line[ipw] = ent[k].y >> 10; img_slot[ipw] = slot[k]; }
    std::vector<char> done(s.nlines, 0);
    const int half = GL / 2;
    for (int l = 0; l < s.nlines; l++) {
      if (done[l]) continue;
      // run of consecutive lines of one plane whose mirrors run backwards
      int ka = 1;
      done[l] = 1;
      const int m0 = mirror[l];
      if (m0 >= 0 && m0 != l) done[m0] = 1;
      while (ka < half && l + ka < s.nlines && !done[l + ka] && pl.h_lin_i2i3[l + ka].y == pl.h_lin_i2i3[l].y && m0 >= 0 && m0 != l &&
             mirror[l + ka] == m0 - ka && mirror[l + ka] != l + ka) {
        done[l + ka] = 1; done[mirror[l + ka]] = 1; ka++;
      }
      const bool has_b = m0 >= 0 && m0 != l;
      const int b0 = has_b ? m0 - (ka - 1) : 0, kb = has_b ? ka : 0;
      bstart.push_back((int)oent.size());
      batches.push_back(push_int4(l, ka, b0, kb));
      auto local = [&](int line) { return (line >= l && line < l + ka) ? line - l : ka + (line - b0); };
      auto emit = [&](int line) {
        for (int k = estart[line]; k < estart[line + 1]; k++) {
          if (ent[k].x < 0) continue;                         // image entries are read through their direct partner
          const int ipw = ent[k].x & 0x3fffffff;
          int y = slot[k] | (local(line) << 8);
          if (!(ent[k].x & (1 << 30))) {
            ABI_CHECK(img_line[ipw] >= 0, "half-support x pass: plane wave without a time-reversed image");
            const int il = img_line[ipw];
            ABI_CHECK((il >= l && il < l + ka) || (kb > 0 && il >= b0 && il < b0 + kb), "half-support x pass: image line outside the batch");
            y |= (img_slot[ipw] << 12) | (local(il) << 20) | (1 << 24);
          }
          oent.push_back(make_int2(ipw | (ent[k].x & (1 << 30)), y));
        }
      };
      for (int q = 0; q < ka; q++) emit(l + q);
      for (int q = 0; q < kb; q++) emit(b0 + q);
    }
  }
  bstart.push_back((int)oent.size());
  s.nbatch3 = (int)batches.size();
  s.d_batches = upload(batches, tabs.owned); s.d_oent = upload(oent, tabs.owned); s.d_bstart = upload(bstart, tabs.owned);
  s.ok = true;
}

XhTabs& tabs_of(const FourwfPlan& pl) {
  if (!pl.xh) pl.xh = new XhTabs();
  XhTabs& t = *pl.xh;
  if (!t.built) {
    t.built = true;
    const Entry* e = find_entry(pl.n1);
    if (e != nullptr && pl.n1 % 2 == 0 && !pl.h_in_estart.empty()) {
      const bool gamma = pl.istwf_k == 2 && pl.same_sphere;
      build_set(pl, *e, t, t.in_set, pl.h_in_ent, pl.h_in_estart, true, gamma);
      if (!pl.h_out_estart.empty()) build_set(pl, *e, t, t.out_set, pl.h_out_ent, pl.h_out_estart, false, false);
    }
  }
  return t;
}
}  // namespace

bool x_stage_usable(const FourwfPlan& pl, bool packed) {
  if (!fourwf_tuning().xhalf || !pl.fused_ok || find_entry(pl.n1) == nullptr) return false;
  XhTabs& t = tabs_of(pl);
  return t.in_set.ok && (packed ? (pl.istwf_k == 2 && pl.same_sphere) : t.out_set.ok);
}

static void fill_common(const FourwfPlan& pl, const XhSet& s, XhParams& P, int nb, int pack_ndat) {
  P.n1 = pl.n1; P.nb = nb; P.nlines = s.nlines; P.pack_ndat = pack_ndat;
  P.tw1 = fft_tables(pl.n1).plan.tw; P.x_sign = s.d_sign; P.x_ovslot = s.d_ovslot;
  P.cg = nullptr; P.out = nullptr; P.W1in = nullptr; P.W1 = nullptr; P.ent = nullptr; P.estart = nullptr;
  P.batches = nullptr; P.oent = nullptr; P.bstart = nullptr; P.xnorm = 1.0; P.kin_filter = 0.0; P.zero_im_g0 = 0; P.order = fourwf_tuning().xh_order;
}

void x_stage_forward(const FourwfPlan& pl, const double2* cg, double2* W1, int nb, int pack_ndat, cudaStream_t st) {
  const Entry* e = find_entry(pl.n1);
  XhTabs& t = tabs_of(pl);
  ABI_CHECK(e != nullptr && t.in_set.ok, "half-support x pass: unsupported plan");
  XhParams P;
  fill_common(pl, t.in_set, P, nb, pack_ndat);
  P.npw = pl.npw_in; P.nbatch = t.in_set.nbatch1; P.cg = cg; P.W1 = W1; P.ent = t.in_set.d_ent; P.estart = pl.d_lin_estart;
  e->fn(0, P, st);
}

void x_stage_backward(const FourwfPlan& pl, const double2* W1o, double2* out, int nb, int pack_ndat, double xnorm, int zero_im_g0,
                      const FourwfEpilogue& epi, double kin_filter, cudaStream_t st) {
  const Entry* e = find_entry(pl.n1);
  XhTabs& t = tabs_of(pl);
  const XhSet& s = pack_ndat > 0 ? t.in_set : t.out_set;
  ABI_CHECK(e != nullptr && s.ok, "half-support x pass: unsupported plan");
  XhParams P;
  fill_common(pl, s, P, nb, pack_ndat);
  P.npw = pl.npw_out; P.nbatch = s.nbatch3; P.W1in = W1o; P.out = out; P.batches = s.d_batches; P.oent = s.d_oent; P.bstart = s.d_bstart;
  P.xnorm = xnorm; P.kin_filter = kin_filter; P.zero_im_g0 = zero_im_g0; P.epi = epi;
  e->fn(1, P, st);
}

}  // namespace abi
