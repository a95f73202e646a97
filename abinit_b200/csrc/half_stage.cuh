// Plane stage of the fused fourwf, "half-support" engine (round 2).
//
// A boxcut >= 2 FFT box holds the G-sphere in at most half of every axis (+1 point): along an axis of length n = 2m the
// non-zero input indices are [0, la) U [n - lb, n) with la, lb <= m.  For such a line the radix-2 decimation-in-frequency
// split costs no butterflies:
//      r in [0, m):   v[r] = x[r]  (r < la)   or   x[r + m]  (r >= m - lb)            (both only on a one-point overlap)
//      X[2k]   = FFT_m( v )[k]
//      X[2k+1] = FFT_m( s_r w^r v )[k],      w = e^{+2 pi i / n},  s_r = +1 (low run) / -1 (high run)
// and on the way back only the 'la + lb' wanted outputs are formed,  x[r] = E[r] + s_r w^-r O[r]  (E, O = the two inverse
// m-point transforms).  A length-n zero-padded transform is therefore TWO m-point transforms plus m twiddle products, and the
// m-point transforms are done as two register-resident passes m = A * B exchanged once through a warp-private shared-memory
// buffer -- Good-Thomas (no inter-pass twiddles) when gcd(A, B) = 1, Cooley-Tukey otherwise:
//      pass 1  item (line, j):        A inputs v[rin(t, j)]  -> DFT_A of the even and of the odd half (2A values)
//      pass 2  item (line, half, k1): B values over j        -> DFT_B  -> outputs at index 2 kout(k1, k2) + half
// Work items of a pass are spread over the lanes of ONE warp; a warp owns a batch of G lines (y phases) or G columns (z phase)
// and never waits for another warp inside a phase.
//
// The (nU x n2) plane S between the y and z phases lives in a per-CTA scratch that is meant to stay in L2 (evict_last policy on
// its accesses, evict_first on the W1 / W1o streams).  Its layout is chosen for the z phase:  S[g][row][c]  with column batches
// g of G columns (column id cid = k2 * 2A + half * A + k1 of the y transform, g = cid / G, c = cid % G) and the occupied z
// planes stored in the order the z pass 1 reads them (t-major, j-minor): a warp's z batch is ONE contiguous block.  V_loc is
// pre-permuted to the order the z pass 2 holds the grid points in registers:  vP[i1][g][k2][c][half * A + k1].
//
// Reference semantics: the zero-padded passes of src/52_fft_mpi_noabirule/fftw3_fftpad.finc:14-196 (forward: x on nlinex lines,
// y on n_zplanes planes, z on all columns; reverse order on the way back), V_loc application src/44_abitools/m_cgtools.F90:2410-2491,
// density accumulation :2338-2384.
#pragma once
#include "plane_stage.cuh"
#include "hdft.cuh"

namespace abi {

struct HalfParams {
  int n1, n2, n3, nb, nU, cplex;
  int nlin, nlout;                    // lines per (band, i1) plane of W1 / W1o
  long long nunits;                   // nb * n1
  const double2* W1; double2* W1o;    // [b][i1][line]
  double2* S;                         // fused: [gridDim.x][ng2][ROWS][G]; split: [nunits][ng2][ROWS][G]
  const double* vP;                   // V_loc permuted [i1][ng2][B3][G][2 A3] (cplex doubles per point)
  const double2* tw2;                 // exp(-2 pi i q / n2), q < n2
  const double2* tw3;                 // exp(-2 pi i q / n3)
  // y direction: per occupied plane u the compact row of W1 / W1o = the la entries i2 in [a, a + la) (a + la <= m2) then the
  // lb entries i2 in [b, b + lb) (b >= m2):  rows[u] = {first line of the plane, a | la << 16, (b - m2) | lb << 16, 0}
  const int4* in_rows; const int4* out_rows;
  int y_amb_in, y_amb_out;            // 1 if some row holds both i2 = r and i2 = r + m2 for some r (overlap of the two runs)
  // z direction, in pass-1 order q = t * B + j (r = rin(t, j)): S row q holds the occupied plane behind v[r]; where both i3 = r and
  // i3 = r + m3 are occupied the high partner sits in one of the OV extra rows m3 + idx
  const int* z_sign;                  // [m3] +1 / -1: s_r of the plane in row q (high planes carry -1); 0: no plane (row stays zero)
  const int* z_ovrow;                 // [m3] extra row (>= m3) of the high partner, or -1
  const int* row_u;                   // [m3 + kHalfOV] the occupied plane u stored in S row q, or -1
  int ng2;                            // column batches: ceil(n2 / G)
  int dbg_skip = 0;                   // developer timing aid: bit 0 / 1 / 2 skips the y / z / y^-1 phase (results are then wrong)
  unsigned long long layout_key = 0;  // identifies (plan, configuration): the scratch is cleared when it changes
  // option 1 (density accumulation): rhoP[i1][ng2][B3][G][2 A3] += wxy[b].x Re(psi)^2 + wxy[b].y Im(psi)^2
  double* rhoP = nullptr; const double2* wxy = nullptr;
};

// ---- cache-policy helpers: S stays in L2 (evict_last), the W1 / W1o streams pass through (evict_first) ----
#ifndef ABI_EMU
ABI_DEV unsigned long long policy_evict_last() {
  unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
ABI_DEV unsigned long long policy_evict_first() {
  unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
ABI_DEV unsigned long long policy_evict_normal() {
  unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p;
}
ABI_DEV double2 ld_keep(const double2* a, unsigned long long pol) {
  double2 v; asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol)); return v;
}
ABI_DEV void st_keep(double2* a, double2 v, unsigned long long pol) {
  asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
ABI_DEV void st_stream(double2* a, double2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
// 16-byte asynchronous global -> shared copy (LDGSTS), L1 bypassed, with an L2 policy
ABI_DEV void cp_async16(double2* sdst, const double2* gsrc, unsigned long long pol) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "l"(pol) : "memory");
}
ABI_DEV void cp_async16_plain(double2* sdst, const double2* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
ABI_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ABI_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
inline unsigned long long policy_evict_last() { return 0; }
inline unsigned long long policy_evict_first() { return 0; }
inline unsigned long long policy_evict_normal() { return 0; }
inline double2 ld_keep(const double2* a, unsigned long long) { return *a; }
inline void st_keep(double2* a, double2 v, unsigned long long) { *a = v; }
inline void st_stream(double2* a, double2 v, unsigned long long) { *a = v; }
inline void cp_async16(double2* sdst, const double2* gsrc, unsigned long long) { *sdst = *gsrc; }
inline void cp_async16_plain(double2* sdst, const double2* gsrc) { *sdst = *gsrc; }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
#endif

ABI_HD constexpr int h_gcd(int a, int b) { return b == 0 ? a : h_gcd(b, a % b); }
ABI_HD constexpr int h_inv_mod(int a, int m) { int r = 0; for (int i = 0; i < m; i++) if ((a * i) % m == 1 % m) r = i; return r; }

// index maps of the two-pass m-point transform, m = A * B (host + device, also used by the planner / V_loc permutation)
template <int A, int B> struct HalfMap {
  static constexpr int M = A * B;
  static constexpr bool PFA = h_gcd(A, B) == 1;
  ABI_HD static constexpr int rin(int t, int j) { return PFA ? (B * t + A * j) % M : j + B * t; }
  ABI_HD static constexpr int kout(int k1, int k2) {
    return PFA ? (k1 * B * h_inv_mod(B % A, A) + k2 * A * h_inv_mod(A % B, B)) % M : k1 + A * k2;
  }
};

constexpr int kHalfOV = 2;   // extra S rows for planes / row entries present at both r and r + m (exact boxcut-2 boxes: one)

template <int A, int B, int G>
struct HalfFft {
  static_assert((2 * A) % G == 0, "the column-batch width must divide 2A (compile-time S strides)");
  using Map = HalfMap<A, B>;
  static constexpr int M = A * B, N = 2 * M;
  static constexpr bool PFA = Map::PFA;
  static constexpr int ROWS = M + kHalfOV;               // rows of one column batch of S
  static constexpr int ZK = (B * G) | 1;                 // odd stride between the (half, k1) slabs of the exchange buffer
  static constexpr int ESIZE = 2 * A * ZK;               // exchange buffer, double2 per warp
  static constexpr int STG = ROWS * G;                   // staging buffer (prefetched W1 rows / S block), double2 per warp
  static constexpr int WSIZE = ESIZE + STG + 1;          // + one slot that always holds zero (reads of absent row entries)
  static constexpr int NG = (N + G - 1) / G;             // column batches of a length-N axis
  static constexpr int GSTR = ROWS * G;                  // double2 between consecutive column batches of S
  // shared-memory tables (double2 slots): Ty[M] | Tz[M] | ctwA[M] | ctwB[M] (Cooley-Tukey only) | ints
  static constexpr int TW_SLOTS = 2 * M + (PFA ? 0 : 2 * M);
  ABI_HD static constexpr int int_slots() { return (B + M + ROWS + 3) / 4; }   // zmask[B], zov[M], rowu[ROWS]

  struct Tables {
    const double2* Ty; const double2* Tz; const double2* ctwA; const double2* ctwB;
    const int* zmask; const int* zov; const int* rowu;
  };

  // tid/nthr: the whole CTA fills the tables once (ends without a barrier: the caller synchronises)
  ABI_DEV static Tables load_tables(double2* sm, const HalfParams& P, bool ytab, bool ztab, int tid, int nthr) {
    double2* Ty = sm; double2* Tz = sm + M; double2* cA = sm + 2 * M; double2* cB = cA + M;
    int* zmask = reinterpret_cast<int*>(sm + TW_SLOTS); int* zov = zmask + B; int* rowu = zov + M;
    for (int q = tid; q < M; q += nthr) {
      const int t = q / B, j = q - t * B;
      const int r = Map::rin(t, j);
      if (ytab) Ty[q] = P.tw2[r];
      if (ztab) {
        double2 w = P.tw3[r];
        if (P.z_sign[q] < 0) { w.x = -w.x; w.y = -w.y; }
        Tz[q] = w; zov[q] = P.z_ovrow[q];
      }
      if (!PFA) {
        // inter-pass twiddle w_m^(j k1): element q = k1 * B + j of ctwA, j * A + k1 of ctwB
        const int k1 = q / B, jj = q - k1 * B;
        const double2 w = (ytab ? P.tw2 : P.tw3)[2 * ((jj * k1) % M)];
        cA[q] = w; cB[jj * A + k1] = w;
      }
    }
    // per j: bit t = row (t, j) holds a plane; bit 16 + t = it also has a high partner in an extra row
    if (ztab) for (int j = tid; j < B; j += nthr) {
      int m = 0;
      for (int t = 0; t < A; t++) { if (P.z_sign[t * B + j] != 0) m |= 1 << t; if (P.z_ovrow[t * B + j] >= 0) m |= 1 << (16 + t); }
      zmask[j] = m;
    }
    for (int q = tid; q < ROWS; q += nthr) rowu[q] = P.row_u[q];
    Tables T; T.Ty = Ty; T.Tz = Tz; T.ctwA = cA; T.ctwB = cB; T.zmask = zmask; T.zov = zov; T.rowu = rowu;
    return T;
  }

  // ---- pass 1 (forward) on registers: x = even-half input, xo = odd-half input already multiplied by s_r w^r ----
  ABI_DEV static void fwd1_store(double2* x, double2* xo, double2* e, const Tables& T, int j) {
    HDft<A, +1>::run(x);
    HDft<A, +1>::run(xo);
    if (!PFA) {
#pragma unroll
      for (int k1 = 1; k1 < A; k1++) { const double2 w = T.ctwA[k1 * B + j]; x[k1] = cmulc(x[k1], w); xo[k1] = cmulc(xo[k1], w); }
    }
#pragma unroll
    for (int k1 = 0; k1 < A; k1++) { e[k1 * ZK] = x[k1]; e[(A + k1) * ZK] = xo[k1]; }
  }
  // ---- pass 1' (inverse) on registers: reads both halves of item (line, j), leaves E[t], O[t] of r = rin(t, j) ----
  ABI_DEV static void inv1_load(double2* ye, double2* yo, const double2* e) {
#pragma unroll
    for (int k1 = 0; k1 < A; k1++) { ye[k1] = e[k1 * ZK]; yo[k1] = e[(A + k1) * ZK]; }
    HDft<A, -1>::run(ye);
    HDft<A, -1>::run(yo);
  }
  // ye + yo * tw as four fused multiply-adds per element
  ABI_DEV static double2 comb(double2 ye, double2 yo, double2 tw) {
    return make_double2(fma(yo.x, tw.x, fma(-yo.y, tw.y, ye.x)), fma(yo.x, tw.y, fma(yo.y, tw.x, ye.y)));
  }

  // ---------------- phase Y: compact rows of W1 -> S ----------------
  // A line batch is G consecutive ROWS of S (rows q0 .. q0 + G - 1 in pass-1 order of the z transform; rowu[q] = the occupied
  // plane stored in row q, or -1): the y-side stores / loads of S then touch G adjacent rows of every column batch, i.e.
  // runs of G * G * 16 bytes instead of G * 16.  The W1 rows of a batch are fetched separately (they are not neighbours in W1).
  static constexpr int RS = M + kHalfOV;                 // staging slots per line
  static constexpr int NBY = (ROWS + G - 1) / G;         // line batches of a plane

  ABI_DEV static void y_prefetch(const HalfParams& P, const Tables& T, const double2* __restrict__ w1, double2* stg, int b,
                                 unsigned long long pstream) {
    if (b >= NBY) return;
#pragma unroll
    for (int line = 0; line < G; line++) {
      const int q = b * G + line;
      const int u = q < ROWS ? T.rowu[q] : -1;
      if (u >= 0) {
        const int4 row = P.in_rows[u];
        const int len = (row.y >> 16) + (row.z >> 16);
        ABI_FOR_LANES {
          for (int p = lane; p < len; p += 32) cp_async16(stg + line * RS + p, w1 + row.x + p, pstream);
        }
      }
    }
    cp_async_commit();
  }

  ABI_DEV static void y_pass1(const HalfParams& P, const Tables& T, double2* E, const double2* stg, int b) {
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, line = w - j * G;
        const int q = b * G + line;
        const int u = (w < G * B && q < ROWS) ? T.rowu[q] : -1;
        if (u >= 0) {
          const int4 row = P.in_rows[u];
          const int a = row.y & 0xffff, la = row.y >> 16, bm = row.z & 0xffff, lb = row.z >> 16;
          const int olo = line * RS - a, ohi = line * RS + (la - bm);   // stg[olo + r] = x[r], stg[ohi + r] = x[r + m]
          double2 x[A], xo[A];
          int r = PFA ? (A * j) % M : j;
#pragma unroll
          for (int t = 0; t < A; t++) {
            const bool lo = (unsigned)(r - a) < (unsigned)la, hi = (unsigned)(r - bm) < (unsigned)lb;
            // absent entries read the zero slot; the odd half of a high-run entry carries s_r = -1 (sign bit flipped on the way)
            const double2 v = stg[lo ? olo + r : (hi ? ohi + r : STG)];
            const int flip = lo ? 0 : (int)0x80000000;
            double2 vo = make_double2(__hiloint2double(__double2hiint(v.x) ^ flip, __double2loint(v.x)),
                                      __hiloint2double(__double2hiint(v.y) ^ flip, __double2loint(v.y)));
            x[t] = v;
            if (P.y_amb_in && lo && hi) { const double2 h = stg[ohi + r]; vo = csub(v, h); x[t] = cadd(v, h); }
            xo[t] = cmulc(vo, T.Ty[t * B + j]);
            r += B; if (PFA && r >= M) r -= M;
          }
          fwd1_store(x, xo, E + j * G + line, T, j);
        }
      }
    }
    ABI_SYNCWARP();
  }

  ABI_DEV static void y_pass2(const Tables& T, double2* __restrict__ S, const double2* E, int b, unsigned long long pkeep) {
    for (int w0 = 0; w0 < 2 * A * G; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w / (2 * A), hk1 = w - line * (2 * A);
        const int q = b * G + line;
        if (w < 2 * A * G && q < ROWS && T.rowu[q] >= 0) {
          const double2* e = E + hk1 * ZK + line;
          double2 v[B];
#pragma unroll
          for (int j = 0; j < B; j++) v[j] = e[j * G];
          HDft<B, +1>::run(v);
          // column id k2 * 2A + hk1 -> batch g = g0 + k2 * (2A / G), c = hk1 % G
          double2* dst = S + q * G + (hk1 / G) * GSTR + (hk1 % G);
#pragma unroll
          for (int k2 = 0; k2 < B; k2++) st_keep(dst + k2 * ((2 * A / G) * GSTR), v[k2], pkeep);
        }
      }
    }
    ABI_SYNCWARP();
  }

  // all line batches of one plane that belong to this warp; the first batch must already be in flight (y_prefetch)
  ABI_DEV static void phase_y(const HalfParams& P, const Tables& T, const double2* __restrict__ w1, double2* __restrict__ S,
                              double2* E, double2* stg, int warp, int nwarps, unsigned long long pkeep, unsigned long long pstream) {
    for (int b = warp; b < NBY; b += nwarps) {
      cp_async_wait_all();
      ABI_SYNCWARP();
      y_pass1(P, T, E, stg, b);
      y_prefetch(P, T, w1, stg, b + nwarps, pstream);   // flies during pass 2
      y_pass2(T, S, E, b, pkeep);
    }
  }

  // ---------------- phase Z: one column batch of S -> z FFT, * V_loc, z FFT^-1 -> S (in place) ----------------
  ABI_DEV static void z_pass1(const Tables& T, const double2* __restrict__ Sg, double2* E, int nl, unsigned long long pkeep) {
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, c = w - j * G;
        if (w < G * B && c < nl) {
          const double2* src = Sg + j * G + c;
          double2 x[A], xo[A];
#pragma unroll
          for (int t = 0; t < A; t++) x[t] = ld_keep(src + t * (B * G), pkeep);
          const int zm = T.zmask[j];
          if (zm >> 16) {                     // a plane pair (r, r + m) in this lane's column: rare, one or two lanes of a warp
#pragma unroll
            for (int t = 0; t < A; t++) {
              double2 vo = x[t];
              if ((zm >> (16 + t)) & 1) { const double2 h = ld_keep(Sg + T.zov[t * B + j] * G + c, pkeep); vo = csub(x[t], h); x[t] = cadd(x[t], h); }
              xo[t] = cmulc(vo, T.Tz[t * B + j]);
            }
          } else {
#pragma unroll
            for (int t = 0; t < A; t++) xo[t] = cmulc(x[t], T.Tz[t * B + j]);
          }
          fwd1_store(x, xo, E + j * G + c, T, j);
        }
      }
    }
    ABI_SYNCWARP();
  }

  template <bool RHO>
  ABI_DEV static void z_batch(const HalfParams& P, const Tables& T, double2* __restrict__ S, const double* __restrict__ vunit,
                              double* __restrict__ runit, double2 wxy, double2* E, int g, unsigned long long pkeep) {
    const int nl = min(G, P.n2 - g * G);
    double2* Sg = S + (size_t)g * GSTR;
    z_pass1(T, Sg, E, nl, pkeep);
    for (int w0 = 0; w0 < 2 * A * G; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int c = w / (2 * A), hk1 = w - c * (2 * A);
        if (w < 2 * A * G && c < nl) {
          double2* e = E + hk1 * ZK + c;
          if (RHO) {
            double2 v[B];
#pragma unroll
            for (int j = 0; j < B; j++) v[j] = e[j * G];
            HDft<B, +1>::run(v);
            double* rp = runit + (size_t)g * (B * G * 2 * A) + c * (2 * A) + hk1;
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) ABI_RED_ADD(rp + k2 * (G * 2 * A), wxy.x * v[k2].x * v[k2].x + wxy.y * v[k2].y * v[k2].y);
          } else {
            // V_loc of this lane's B grid points first: the loads fly while the exchange buffer is read and transformed
            double2 vv[B];
            if (P.cplex == 1) {
              const double* vp = vunit + (size_t)g * (B * G * 2 * A) + c * (2 * A) + hk1;
#pragma unroll
              for (int k2 = 0; k2 < B; k2++) vv[k2] = make_double2(ldg1(vp + k2 * (G * 2 * A)), 0.0);
            } else {
              const double2* vp = reinterpret_cast<const double2*>(vunit) + (size_t)g * (B * G * 2 * A) + c * (2 * A) + hk1;
#pragma unroll
              for (int k2 = 0; k2 < B; k2++) vv[k2] = ldg2(vp + k2 * (G * 2 * A));
            }
            double2 v[B];
#pragma unroll
            for (int j = 0; j < B; j++) v[j] = e[j * G];
            HDft<B, +1>::run(v);
            if (P.cplex == 1) {
#pragma unroll
              for (int k2 = 0; k2 < B; k2++) { v[k2].x *= vv[k2].x; v[k2].y *= vv[k2].x; }
            } else {
#pragma unroll
              for (int k2 = 0; k2 < B; k2++) v[k2] = cmul(v[k2], vv[k2]);
            }
            HDft<B, -1>::run(v);
            const int k1 = hk1 >= A ? hk1 - A : hk1;
            e[0] = v[0];
#pragma unroll
            for (int j = 1; j < B; j++) e[j * G] = PFA ? v[j] : cmul(v[j], T.ctwB[j * A + k1]);
          }
        }
      }
    }
    ABI_SYNCWARP();
    if (RHO) return;
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, c = w - j * G;
        if (w < G * B && c < nl) {
          double2 ye[A], yo[A];
          inv1_load(ye, yo, E + j * G + c);
          double2* dst = Sg + j * G + c;
          const int zm = T.zmask[j];
#pragma unroll
          for (int t = 0; t < A; t++) {
            const double2 tw = T.Tz[t * B + j];
            if ((zm >> t) & 1) st_keep(dst + t * (B * G), comb(ye[t], yo[t], tw), pkeep);
            if ((zm >> (16 + t)) & 1) st_keep(Sg + T.zov[t * B + j] * G + c, comb(ye[t], yo[t], make_double2(-tw.x, -tw.y)), pkeep);
          }
        }
      }
    }
    ABI_SYNCWARP();
  }

  template <bool RHO>
  ABI_DEV static void phase_z(const HalfParams& P, const Tables& T, double2* __restrict__ S, const double* __restrict__ vunit,
                              double* __restrict__ runit, double2 wxy, double2* E, int warp, int nwarps, unsigned long long pkeep) {
    for (int g = warp; g < P.ng2; g += nwarps) z_batch<RHO>(P, T, S, vunit, runit, wxy, E, g, pkeep);
  }

  // ---------------- phase Y': S -> y FFT^-1 -> compact output rows of W1o ----------------
  ABI_DEV static void yinv_batch(const HalfParams& P, const Tables& T, const double2* __restrict__ S, double2* __restrict__ w1o,
                                 double2* E, double2* stg, int b, unsigned long long pkeep, unsigned long long pstream) {
    for (int w0 = 0; w0 < 2 * A * G; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w / (2 * A), hk1 = w - line * (2 * A);
        const int q = b * G + line;
        if (w < 2 * A * G && q < ROWS && T.rowu[q] >= 0) {
          const double2* src = S + q * G + (hk1 / G) * GSTR + (hk1 % G);
          double2 v[B];
#pragma unroll
          for (int k2 = 0; k2 < B; k2++) v[k2] = ld_keep(src + k2 * ((2 * A / G) * GSTR), pkeep);
          HDft<B, -1>::run(v);
          const int k1 = hk1 >= A ? hk1 - A : hk1;
          double2* e = E + hk1 * ZK + line;
          e[0] = v[0];
#pragma unroll
          for (int j = 1; j < B; j++) e[j * G] = PFA ? v[j] : cmul(v[j], T.ctwB[j * A + k1]);
        }
      }
    }
    ABI_SYNCWARP();
    // pass 1': the wanted outputs of each line go to the staging buffer in compact row order ...
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, line = w - j * G;
        const int q = b * G + line;
        const int u = (w < G * B && q < ROWS) ? T.rowu[q] : -1;
        if (u >= 0) {
          const int4 row = P.out_rows[u];
          const int a = row.y & 0xffff, la = row.y >> 16, bm = row.z & 0xffff, lb = row.z >> 16;
          double2* dlo = stg + line * RS - a;
          double2* dhi = stg + line * RS + (la - bm);
          double2 ye[A], yo[A];
          inv1_load(ye, yo, E + j * G + line);
          int r = PFA ? (A * j) % M : j;
#pragma unroll
          for (int t = 0; t < A; t++) {
            const double2 tw = T.Ty[t * B + j];
            if ((unsigned)(r - a) < (unsigned)la) dlo[r] = comb(ye[t], yo[t], tw);
            if ((unsigned)(r - bm) < (unsigned)lb) dhi[r] = comb(ye[t], yo[t], make_double2(-tw.x, -tw.y));
            r += B; if (PFA && r >= M) r -= M;
          }
        }
      }
    }
    ABI_SYNCWARP();
    // ... and leave as whole rows (coalesced 16-byte stores)
#pragma unroll
    for (int line = 0; line < G; line++) {
      const int q = b * G + line;
      const int u = q < ROWS ? T.rowu[q] : -1;
      if (u >= 0) {
        const int4 row = P.out_rows[u];
        const int len = (row.y >> 16) + (row.z >> 16);
        ABI_FOR_LANES {
          for (int p = lane; p < len; p += 32) st_stream(w1o + row.x + p, stg[line * RS + p], pstream);
        }
      }
    }
    ABI_SYNCWARP();
  }
};

// dynamic shared memory of the kernels below (bytes)
template <int A, int B, int G> ABI_HD constexpr size_t half_smem_bytes(int warps, int nU) {
  using F = HalfFft<A, B, G>;
  (void)nU; return sizeof(double2) * ((size_t)F::TW_SLOTS + F::int_slots() + (size_t)warps * F::WSIZE);
}

// one CTA = one (transform, i1) plane at a time; warps take line / column batches round-robin inside each phase
// KIND 0: option 2 (y, z * V_loc, y^-1), KIND 1: option 1 (y, z, density accumulation)
template <int A, int B, int G, int WARPS, int MINB, int KIND>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_hw_plane(HalfParams P) {
  using F = HalfFft<A, B, G>;
  ABI_DYN_SMEM(double2, sm);
#ifdef ABI_EMU
  const int warp = 0, nwarps = 1, tid = 0, nthr = 1;
#else
  const int warp = threadIdx.x >> 5, nwarps = WARPS, tid = threadIdx.x, nthr = WARPS * 32;
#endif
  const typename F::Tables T = F::load_tables(sm, P, true, true, tid, nthr);
  double2* E = sm + F::TW_SLOTS + F::int_slots() + (size_t)warp * F::WSIZE;
  double2* stg = E + F::ESIZE;
  stg[F::STG] = make_double2(0.0, 0.0);                  // the zero slot (every lane writes the same value)
  double2* S = P.S + (size_t)blockIdx.x * P.ng2 * F::GSTR;
  const unsigned long long pkeep = policy_evict_last(), pstream = policy_evict_first();
  const size_t vplane = (size_t)P.ng2 * (B * G * 2 * A);
  __syncthreads();
  if (blockIdx.x < P.nunits) {
    const int i1 = (int)(blockIdx.x / P.nb), b = (int)(blockIdx.x - (long long)i1 * P.nb);
    F::y_prefetch(P, T, P.W1 + ((size_t)b * P.n1 + i1) * P.nlin, stg, warp, pstream);
  }
  for (long long unit = blockIdx.x; unit < P.nunits; unit += gridDim.x) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
    if (!(P.dbg_skip & 1)) F::phase_y(P, T, w1, S, E, stg, warp, nwarps, pkeep, pstream);
    __syncthreads();
    if (P.dbg_skip & 2) {
    } else if (KIND == 0) {
      F::template phase_z<false>(P, T, S, P.vP + (size_t)P.cplex * i1 * vplane, nullptr, make_double2(0.0, 0.0), E, warp, nwarps, pkeep);
    } else {
      F::template phase_z<true>(P, T, S, nullptr, P.rhoP + (size_t)i1 * vplane, P.wxy[b], E, warp, nwarps, pkeep);
    }
    __syncthreads();
    if (KIND == 0 && !(P.dbg_skip & 4)) {
      double2* w1o = P.W1o + ((size_t)b * P.n1 + i1) * P.nlout;
      for (int bb = warp; bb < F::NBY; bb += nwarps) F::yinv_batch(P, T, S, w1o, E, stg, bb, pkeep, pstream);
    }
    // the staging buffer is idle again: fetch this warp's first line batch of the next plane while the others finish
    const long long next = unit + gridDim.x;
    if (next < P.nunits) {
      const int i1n = (int)(next / P.nb), bn = (int)(next - (long long)i1n * P.nb);
      F::y_prefetch(P, T, P.W1 + ((size_t)bn * P.n1 + i1n) * P.nlin, stg, warp, pstream);
    }
    if (KIND == 0) __syncthreads();
  }
  cp_async_wait_all();
}

// ---- host interface (half_stage.cu) ----
struct FourwfPlan; struct VlocDev;
struct HalfCfg { int n, A, B, G; };
// two-pass configuration of a length-n axis (n even, n / 2 = A * B), or nullptr when the half-support engine has no kernel for n
const HalfCfg* half_stage_cfg(int n);
// true when the fused half-support plane kernel can run this plan (cubic yz plane, half-support sphere, kernels instantiated)
bool half_stage_usable(const FourwfPlan& pl, bool out_is_in);
struct HalfLaunch {
  int nb = 0; const double2* W1 = nullptr; double2* W1o = nullptr; int nlin = 0, nlout = 0;
  bool out_is_in = false;          // packed Gamma path: the output rows are the (completed) input rows
  double* rhoP = nullptr; const double2* wxy = nullptr;   // option 1
};
void half_stage_launch(const FourwfPlan& pl, const VlocDev& v, const HalfLaunch& L, cudaStream_t st);
// option 1: rhoP (permuted, n1 * half_rho_plane doubles, zero-initialised by the caller) and the final un-permute-add into denpot
size_t half_rho_elems(const FourwfPlan& pl);
void half_stage_launch_rho(const FourwfPlan& pl, const HalfLaunch& L, cudaStream_t st);
void half_rho_unpermute_add(const FourwfPlan& pl, const double* rhoP, double* denpot, cudaStream_t st);
void half_stage_release();

}  // namespace abi
