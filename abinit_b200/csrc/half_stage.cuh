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

namespace abi {

struct HalfParams {
  int n1, n2, n3, nb, nU, cplex;
  int nlin, nlout;                    // lines per (band, i1) plane of W1 / W1o
  long long nunits;                   // nb * n1
  const double2* W1; double2* W1o;    // [b][i1][line]
  double2* S;                         // fused: [gridDim.x][ng2][nU][G]; split: [nunits][ng2][nU][G]
  const double* vP;                   // V_loc permuted [i1][ng2][B3][G][2 A3] (cplex doubles per point)
  const double2* tw2;                 // exp(-2 pi i q / n2), q < n2
  const double2* tw3;                 // exp(-2 pi i q / n3)
  // y direction: per occupied plane u the compact row of W1 / W1o = the la entries i2 in [a, a + la) (a + la <= m2) then the
  // lb entries i2 in [b, b + lb) (b >= m2):  rows[u] = {first line of the plane, a | la << 16, (b - m2) | lb << 16, 0}
  const int4* in_rows; const int4* out_rows;
  int y_amb_in, y_amb_out;            // 1 if some row holds both i2 = r and i2 = r + m2 for some r (overlap of the two runs)
  const int* z_rowoff;                // [m3] in pass-1 order (t * B + j): row * G of the plane holding v[r], or -1
  const int* z_ovoff;                 // [m3] row * G of the high partner where both r and r + m3 are occupied, else -1
  const int* z_sign;                  // [m3] +1 / -1: s_r of the plane behind z_rowoff
  const int* u_row;                   // [nU] row * G of plane u
  int z_has_ov;
  int ng2;                            // column batches: ceil(n2 / G)
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
ABI_DEV double2 ld_keep(const double2* a, unsigned long long pol) {
  double2 v; asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol)); return v;
}
ABI_DEV void st_keep(double2* a, double2 v, unsigned long long pol) {
  asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
ABI_DEV double2 ld_stream(const double2* a, unsigned long long pol) {
  double2 v; asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol)); return v;
}
ABI_DEV void st_stream(double2* a, double2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
#else
inline unsigned long long policy_evict_last() { return 0; }
inline unsigned long long policy_evict_first() { return 0; }
inline double2 ld_keep(const double2* a, unsigned long long) { return *a; }
inline void st_keep(double2* a, double2 v, unsigned long long) { *a = v; }
inline double2 ld_stream(const double2* a, unsigned long long) { return *a; }
inline void st_stream(double2* a, double2 v, unsigned long long) { *a = v; }
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

template <int A, int B, int G>
struct HalfFft {
  using Map = HalfMap<A, B>;
  static constexpr int M = A * B, N = 2 * M;
  static constexpr bool PFA = Map::PFA;
  static constexpr int ZK = (B * G) | 1;                 // odd stride between the (half, k1) slabs of the exchange buffer
  static constexpr int ESIZE = 2 * A * ZK;               // double2 per warp
  static constexpr int NG = (N + G - 1) / G;             // column batches of a length-N axis
  // shared-memory tables (double2 slots): Ty[M] | Tz[M] | ctwA[M] | ctwB[M] (Cooley-Tukey only) | ints
  static constexpr int TW_SLOTS = 2 * M + (PFA ? 0 : 2 * M);
  ABI_HD static constexpr int int_slots(int nU) { return (2 * M + nU + 3) / 4; }   // z_rowoff, z_ovoff, u_row

  struct Tables {
    const double2* Ty; const double2* Tz; const double2* ctwA; const double2* ctwB;
    const int* zrow; const int* zov; const int* urow;
  };

  // tid/nthr: the whole CTA fills the tables once
  ABI_DEV static Tables load_tables(double2* sm, const HalfParams& P, bool ytab, bool ztab, int tid, int nthr) {
    double2* Ty = sm; double2* Tz = sm + M; double2* cA = sm + 2 * M; double2* cB = cA + M;
    int* zrow = reinterpret_cast<int*>(sm + TW_SLOTS); int* zov = zrow + M; int* urow = zov + M;
    for (int q = tid; q < M; q += nthr) {
      const int t = q / B, j = q - t * B;
      const int r = Map::rin(t, j);
      if (ytab) Ty[q] = P.tw2[r];
      if (ztab) {
        double2 w = P.tw3[r];
        if (P.z_sign[q] < 0) { w.x = -w.x; w.y = -w.y; }
        Tz[q] = w; zrow[q] = P.z_rowoff[q]; zov[q] = P.z_ovoff[q];
      }
      if (!PFA) {
        // inter-pass twiddle w_m^(j k1): element q = k1 * B + j of ctwA, j * A + k1 of ctwB
        const int k1 = q / B, jj = q - k1 * B;
        const double2 w = (ytab ? P.tw2 : P.tw3)[2 * ((jj * k1) % M)];
        cA[q] = w; cB[jj * A + k1] = w;
      }
    }
    for (int u = tid; u < P.nU; u += nthr) urow[u] = P.u_row[u];
    Tables T; T.Ty = Ty; T.Tz = Tz; T.ctwA = cA; T.ctwB = cB; T.zrow = zrow; T.zov = zov; T.urow = urow;
    return T;
  }

  // ---- pass 1 (forward) on registers: x = even-half input, xo = odd-half input already multiplied by s_r w^r ----
  ABI_DEV static void fwd1_store(double2* x, double2* xo, double2* e, const Tables& T, int j) {
    Dft<A, +1>::run(x);
    Dft<A, +1>::run(xo);
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
    Dft<A, -1>::run(ye);
    Dft<A, -1>::run(yo);
  }

  // ---------------- phase Y: compact rows of W1 -> S ----------------
  ABI_DEV static void phase_y(const HalfParams& P, const Tables& T, const double2* __restrict__ w1, double2* __restrict__ S,
                              double2* E, int u0, unsigned long long pkeep, unsigned long long pstream) {
    const int nl = min(G, P.nU - u0);
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, line = w - j * G;
        if (w < G * B && line < nl) {
          const int4 row = P.in_rows[u0 + line];
          const int a = row.y & 0xffff, la = row.y >> 16, bm = row.z & 0xffff, lb = row.z >> 16;
          const double2* slo = w1 + row.x - a;                 // slo[r] = x[r]      for r in [a, a + la)
          const double2* shi = w1 + row.x + (la - bm);         // shi[r] = x[r + m]  for r in [bm, bm + lb)
          double2 x[A], xo[A];
          int r = PFA ? (A * j) % M : j;
#pragma unroll
          for (int t = 0; t < A; t++) {
            const bool lo = (unsigned)(r - a) < (unsigned)la, hi = (unsigned)(r - bm) < (unsigned)lb;
            double2 v = make_double2(0.0, 0.0);
            if (lo) v = ld_stream(slo + r, pstream);
            else if (hi) v = ld_stream(shi + r, pstream);
            double2 tw = T.Ty[t * B + j];
            if (!lo) { tw.x = -tw.x; tw.y = -tw.y; }
            double2 vo = v;
            if (P.y_amb_in && lo && hi) { const double2 h = ld_stream(shi + r, pstream); vo = csub(v, h); v = cadd(v, h); }
            x[t] = v; xo[t] = cmulc(vo, tw);
            r += B; if (PFA && r >= M) r -= M;
          }
          fwd1_store(x, xo, E + j * G + line, T, j);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < 2 * A * G; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w / (2 * A), hk1 = w - line * (2 * A);
        if (w < 2 * A * G && line < nl) {
          const double2* e = E + hk1 * ZK + line;
          double2 v[B];
#pragma unroll
          for (int j = 0; j < B; j++) v[j] = e[j * G];
          Dft<B, +1>::run(v);
          double2* dst = S + T.urow[u0 + line];
          int g = hk1 / G, c = hk1 - g * G;
#pragma unroll
          for (int k2 = 0; k2 < B; k2++) {
            st_keep(dst + (size_t)g * (P.nU * G) + c, v[k2], pkeep);
            g += (2 * A) / G; c += (2 * A) % G;
            if ((2 * A) % G != 0 && c >= G) { c -= G; g++; }
          }
        }
      }
    }
    ABI_SYNCWARP();
  }

  // ---------------- phase Z: one column batch of S -> z FFT, * V_loc, z FFT^-1 -> S (in place) ----------------
  ABI_DEV static void z_pass1(const HalfParams& P, const Tables& T, const double2* __restrict__ Sg, double2* E, int nl,
                              unsigned long long pkeep) {
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, c = w - j * G;
        if (w < G * B && c < nl) {
          const double2* src = Sg + c;
          double2 x[A], xo[A];
#pragma unroll
          for (int t = 0; t < A; t++) {
            const int o = T.zrow[t * B + j];
            x[t] = (o >= 0) ? ld_keep(src + o, pkeep) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int t = 0; t < A; t++) {
            double2 vo = x[t];
            if (P.z_has_ov) {
              const int o2 = T.zov[t * B + j];
              if (o2 >= 0) { const double2 h = ld_keep(src + o2, pkeep); vo = csub(x[t], h); x[t] = cadd(x[t], h); }
            }
            xo[t] = cmulc(vo, T.Tz[t * B + j]);
          }
          fwd1_store(x, xo, E + j * G + c, T, j);
        }
      }
    }
    ABI_SYNCWARP();
  }

  ABI_DEV static void phase_z(const HalfParams& P, const Tables& T, double2* __restrict__ S, const double* __restrict__ vunit,
                              double2* E, int g, unsigned long long pkeep) {
    const int nl = min(G, P.n2 - g * G);
    double2* Sg = S + (size_t)g * (P.nU * G);
    z_pass1(P, T, Sg, E, nl, pkeep);
    for (int w0 = 0; w0 < 2 * A * G; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int c = w / (2 * A), hk1 = w - c * (2 * A);
        if (w < 2 * A * G && c < nl) {
          double2* e = E + hk1 * ZK + c;
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
          Dft<B, +1>::run(v);
          if (P.cplex == 1) {
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) { v[k2].x *= vv[k2].x; v[k2].y *= vv[k2].x; }
          } else {
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) v[k2] = cmul(v[k2], vv[k2]);
          }
          Dft<B, -1>::run(v);
          const int k1 = hk1 >= A ? hk1 - A : hk1;
          e[0] = v[0];
#pragma unroll
          for (int j = 1; j < B; j++) e[j * G] = PFA ? v[j] : cmul(v[j], T.ctwB[j * A + k1]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, c = w - j * G;
        if (w < G * B && c < nl) {
          double2 ye[A], yo[A];
          inv1_load(ye, yo, E + j * G + c);
          double2* dst = Sg + c;
#pragma unroll
          for (int t = 0; t < A; t++) {
            const int o = T.zrow[t * B + j];
            const double2 to = cmul(yo[t], T.Tz[t * B + j]);
            if (o >= 0) st_keep(dst + o, cadd(ye[t], to), pkeep);
            if (P.z_has_ov) {
              const int o2 = T.zov[t * B + j];
              if (o2 >= 0) st_keep(dst + o2, csub(ye[t], to), pkeep);
            }
          }
        }
      }
    }
    ABI_SYNCWARP();
  }

  // ---------------- phase Z (option 1): z FFT of one column batch -> rhoP += w |psi(r)|^2 ----------------
  ABI_DEV static void phase_z_rho(const HalfParams& P, const Tables& T, const double2* __restrict__ S, double* __restrict__ runit,
                                  double2 wxy, double2* E, int g, unsigned long long pkeep) {
    const int nl = min(G, P.n2 - g * G);
    const double2* Sg = S + (size_t)g * (P.nU * G);
    z_pass1(P, T, Sg, E, nl, pkeep);
    for (int w0 = 0; w0 < 2 * A * G; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int c = w / (2 * A), hk1 = w - c * (2 * A);
        if (w < 2 * A * G && c < nl) {
          const double2* e = E + hk1 * ZK + c;
          double2 v[B];
#pragma unroll
          for (int j = 0; j < B; j++) v[j] = e[j * G];
          Dft<B, +1>::run(v);
          double* rp = runit + (size_t)g * (B * G * 2 * A) + c * (2 * A) + hk1;
#pragma unroll
          for (int k2 = 0; k2 < B; k2++) ABI_RED_ADD(rp + k2 * (G * 2 * A), wxy.x * v[k2].x * v[k2].x + wxy.y * v[k2].y * v[k2].y);
        }
      }
    }
    ABI_SYNCWARP();
  }

  // ---------------- phase Y': S -> y FFT^-1 -> compact output rows of W1o ----------------
  ABI_DEV static void phase_yinv(const HalfParams& P, const Tables& T, const double2* __restrict__ S, double2* __restrict__ w1o,
                                 double2* E, int u0, unsigned long long pkeep, unsigned long long pstream) {
    const int nl = min(G, P.nU - u0);
    for (int w0 = 0; w0 < 2 * A * G; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w / (2 * A), hk1 = w - line * (2 * A);
        if (w < 2 * A * G && line < nl) {
          const double2* src = S + T.urow[u0 + line];
          double2 v[B];
          int g = hk1 / G, c = hk1 - g * G;
#pragma unroll
          for (int k2 = 0; k2 < B; k2++) {
            v[k2] = ld_keep(src + (size_t)g * (P.nU * G) + c, pkeep);
            g += (2 * A) / G; c += (2 * A) % G;
            if ((2 * A) % G != 0 && c >= G) { c -= G; g++; }
          }
          Dft<B, -1>::run(v);
          const int k1 = hk1 >= A ? hk1 - A : hk1;
          double2* e = E + hk1 * ZK + line;
          e[0] = v[0];
#pragma unroll
          for (int j = 1; j < B; j++) e[j * G] = PFA ? v[j] : cmul(v[j], T.ctwB[j * A + k1]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < G * B; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / G, line = w - j * G;
        if (w < G * B && line < nl) {
          const int4 row = P.out_rows[u0 + line];
          const int a = row.y & 0xffff, la = row.y >> 16, bm = row.z & 0xffff, lb = row.z >> 16;
          double2* dlo = w1o + row.x - a;
          double2* dhi = w1o + row.x + (la - bm);
          double2 ye[A], yo[A];
          inv1_load(ye, yo, E + j * G + line);
          int r = PFA ? (A * j) % M : j;
#pragma unroll
          for (int t = 0; t < A; t++) {
            const double2 to = cmul(yo[t], T.Ty[t * B + j]);
            if ((unsigned)(r - a) < (unsigned)la) st_stream(dlo + r, cadd(ye[t], to), pstream);
            if ((unsigned)(r - bm) < (unsigned)lb) st_stream(dhi + r, csub(ye[t], to), pstream);
            r += B; if (PFA && r >= M) r -= M;
          }
        }
      }
    }
    ABI_SYNCWARP();
  }
};

// dynamic shared memory of the kernels below (bytes)
template <int A, int B, int G> ABI_HD constexpr size_t half_smem_bytes(int warps, int nU) {
  using F = HalfFft<A, B, G>;
  return sizeof(double2) * ((size_t)F::TW_SLOTS + F::int_slots(nU) + (size_t)warps * F::ESIZE);
}

// one CTA = one (transform, i1) plane at a time; warps take line / column batches round-robin inside each phase
template <int A, int B, int G, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_hw_plane(HalfParams P) {
  using F = HalfFft<A, B, G>;
  ABI_DYN_SMEM(double2, sm);
#ifdef ABI_EMU
  const int warp = 0, nwarps = 1, tid = 0, nthr = 1;
#else
  const int warp = threadIdx.x >> 5, nwarps = WARPS, tid = threadIdx.x, nthr = WARPS * 32;
#endif
  const typename F::Tables T = F::load_tables(sm, P, true, true, tid, nthr);
  double2* E = sm + F::TW_SLOTS + F::int_slots(P.nU) + (size_t)warp * F::ESIZE;
  double2* S = P.S + (size_t)blockIdx.x * P.ng2 * P.nU * G;
  const unsigned long long pkeep = policy_evict_last(), pstream = policy_evict_first();
  const size_t vplane = (size_t)P.cplex * P.ng2 * (B * G * 2 * A);
  __syncthreads();
  for (long long unit = blockIdx.x; unit < P.nunits; unit += gridDim.x) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
    double2* w1o = P.W1o + ((size_t)b * P.n1 + i1) * P.nlout;
    const double* vunit = P.vP + (size_t)i1 * vplane;
    for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_y(P, T, w1, S, E, u0, pkeep, pstream);
    __syncthreads();
    for (int g = warp; g < P.ng2; g += nwarps) F::phase_z(P, T, S, vunit, E, g, pkeep);
    __syncthreads();
    for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_yinv(P, T, S, w1o, E, u0, pkeep, pstream);
    __syncthreads();
  }
}

// option 1: y FFT, z FFT, density accumulation (no way back)
template <int A, int B, int G, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_hw_plane_rho(HalfParams P) {
  using F = HalfFft<A, B, G>;
  ABI_DYN_SMEM(double2, sm);
#ifdef ABI_EMU
  const int warp = 0, nwarps = 1, tid = 0, nthr = 1;
#else
  const int warp = threadIdx.x >> 5, nwarps = WARPS, tid = threadIdx.x, nthr = WARPS * 32;
#endif
  const typename F::Tables T = F::load_tables(sm, P, true, true, tid, nthr);
  double2* E = sm + F::TW_SLOTS + F::int_slots(P.nU) + (size_t)warp * F::ESIZE;
  double2* S = P.S + (size_t)blockIdx.x * P.ng2 * P.nU * G;
  const unsigned long long pkeep = policy_evict_last(), pstream = policy_evict_first();
  const size_t rplane = (size_t)P.ng2 * (B * G * 2 * A);
  __syncthreads();
  for (long long unit = blockIdx.x; unit < P.nunits; unit += gridDim.x) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
    double* runit = P.rhoP + (size_t)i1 * rplane;
    const double2 wxy = P.wxy[b];
    for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_y(P, T, w1, S, E, u0, pkeep, pstream);
    __syncthreads();
    for (int g = warp; g < P.ng2; g += nwarps) F::phase_z_rho(P, T, S, runit, wxy, E, g, pkeep);
    __syncthreads();
  }
}

// Split plane stage for n2 != n3: the same phases as three kernels, each templated on ONE half-length, with the S planes of
// all units of the chunk in global memory.  kind: 0 = y, 1 = z (* V_loc), 2 = y^-1, 3 = z + density accumulation.
// G is the column-batch width of S and must be the same for the y and z kernels of one launch (the host picks it).
template <int A, int B, int G, int WARPS, int KIND>
__global__ void __launch_bounds__(WARPS * 32, 2) k_hw_plane_split(HalfParams P) {
  using F = HalfFft<A, B, G>;
  ABI_DYN_SMEM(double2, sm);
#ifdef ABI_EMU
  const int warp = 0, nwarps = 1, tid = 0, nthr = 1;
#else
  const int warp = threadIdx.x >> 5, nwarps = WARPS, tid = threadIdx.x, nthr = WARPS * 32;
#endif
  constexpr bool ZK_ = (KIND == 1 || KIND == 3);
  const typename F::Tables T = F::load_tables(sm, P, !ZK_, ZK_, tid, nthr);
  double2* E = sm + F::TW_SLOTS + F::int_slots(P.nU) + (size_t)warp * F::ESIZE;
  const unsigned long long pkeep = policy_evict_last(), pstream = policy_evict_first();
  const size_t vplane = (size_t)(KIND == 1 ? P.cplex : 1) * P.ng2 * (B * G * 2 * A);
  __syncthreads();
  for (long long unit = blockIdx.x; unit < P.nunits; unit += gridDim.x) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    double2* S = P.S + (size_t)unit * P.ng2 * P.nU * G;
    if (KIND == 0) {
      const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
      for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_y(P, T, w1, S, E, u0, pkeep, pstream);
    } else if (KIND == 1) {
      const double* vunit = P.vP + (size_t)i1 * vplane;
      for (int g = warp; g < P.ng2; g += nwarps) F::phase_z(P, T, S, vunit, E, g, pkeep);
    } else if (KIND == 2) {
      double2* w1o = P.W1o + ((size_t)b * P.n1 + i1) * P.nlout;
      for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_yinv(P, T, S, w1o, E, u0, pkeep, pstream);
    } else {
      double* runit = P.rhoP + (size_t)i1 * vplane;
      const double2 wxy = P.wxy[b];
      for (int g = warp; g < P.ng2; g += nwarps) F::phase_z_rho(P, T, S, runit, wxy, E, g, pkeep);
    }
  }
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
