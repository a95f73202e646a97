// x passes of the fused fourwf on the half-support two-pass engine (see half_stage.cuh for the transform itself).
//
//   K1  k_xh_forward : sphere coefficients of GL lines -> zero-padded x FFT (e^{+i}) -> W1[b][i1][line]
//   K3  k_xh_backward: W1o[b][i1][line] -> x FFT^-1 on the wanted outputs -> sphere (* 1/N, Gamma-point unpack, getghc assembly)
//
// One WARP owns a batch of GL lines from the sphere to W1 (or back): no block barrier, the exchange between the two radix passes
// goes through a warp-private shared-memory buffer.  Lanes run over (item, line) with the LINE fastest, so that every access to
// W1 / W1o is a run of GL consecutive 16-byte words at one i1 (the layout the plane stage wants), and the coefficients of a
// batch of consecutive lines are one contiguous piece of the wavefunction array.
//
// Gamma point, two bands per transform (cwavef_double_rfft_trick_pack/unpack, src/66_wfs/m_getghc.F90:1999-2171): F = H C + i H D
// on the completed sphere; the line of -G is the mirror (-i2, -i3) of the line of G.  A K3 batch therefore holds GL/2 lines AND
// their GL/2 mirror lines: both F(G) and F(-G) are in the warp's staging buffer and every output coefficient is written once
// (no zero-fill of the output, no atomics).
#pragma once
#include "half_stage.cuh"
#include "fourwf.cuh"

namespace abi {

struct XhParams {
  int n1, nb, npw, nlines;            // nlines: lines per transform of W1 / W1o
  int nbatch;                         // line batches per transform
  int pack_ndat;                      // > 0: Gamma-point packing, number of bands covered by the nb transforms
  const double2* cg; double2* out;    // sphere arrays [band][npw]
  const double2* W1in; double2* W1;   // K3 input / K1 output [b][i1][line]
  const double2* tw1;                 // exp(-2 pi i q / n1)
  const int* x_sign;                  // [M] +1 / -1 per slot q = t * B + j (sign of the odd-half twiddle)
  const int* x_ovslot;                // [M] extra slot (>= M) of the high partner where both i1 = r and r + m occur, else -1
  // K1: entries sorted by line, {src, (line << 10) | slot}; estart[line]
  const int2* ent; const int* estart;
  // K3: batches {a0, ka, b0, kb} (lines [a0, a0 + ka) then [b0, b0 + kb)); entries sorted by batch
  //     {ipw | g0flag << 30, slotD | lineD << 8 | slotI << 12 | lineI << 20 | has_image << 24}; bstart[batch]
  const int4* batches; const int2* oent; const int* bstart;
  double xnorm; double kin_filter; int zero_im_g0;
  int order;                          // unit order: 0 transform fastest, 1 batch fastest (neighbouring warps touch neighbouring 128-byte runs)
  FourwfEpilogue epi;
};

template <int A, int B, int GL>
struct XHalf {
  using Map = HalfMap<A, B>;
  static constexpr int M = A * B, N = 2 * M;
  static constexpr bool PFA = Map::PFA;
  static constexpr int RS = (M + kHalfOV) | 1;           // staging slots per line (odd: lines fall in different banks)
  static constexpr int ZK = (B * GL) | 1;
  static constexpr int ESIZE = 2 * A * ZK;
  static constexpr int STG = GL * RS;
  static constexpr int WSIZE = ESIZE + STG;
  static constexpr int TW_SLOTS = M + (PFA ? 0 : 2 * M);
  ABI_HD static constexpr int int_slots() { return (B + M + 2 * M + 3) / 4; }   // xmask[B], xov[M], i1row[2M]

  struct Tables { const double2* Tx; const double2* ctwA; const double2* ctwB; const int* xmask; const int* xov; const int* i1row; };

  ABI_DEV static Tables load_tables(double2* sm, const XhParams& P, int tid, int nthr) {
    double2* Tx = sm; double2* cA = sm + M; double2* cB = cA + M;
    int* xmask = reinterpret_cast<int*>(sm + TW_SLOTS); int* xov = xmask + B; int* i1row = xov + M;
    // row = (h * A + k1) * B + k2 of the exchange buffer <-> i1 = 2 kout(k1, k2) + h
    for (int q = tid; q < 2 * M; q += nthr) {
      const int hk1 = q / B, k2 = q - hk1 * B, h = hk1 >= A ? 1 : 0;
      i1row[q] = 2 * Map::kout(hk1 - h * A, k2) + h;
    }
    for (int q = tid; q < M; q += nthr) {
      const int t = q / B, j = q - t * B;
      double2 w = P.tw1[Map::rin(t, j)];
      if (P.x_sign[q] < 0) { w.x = -w.x; w.y = -w.y; }
      Tx[q] = w; xov[q] = P.x_ovslot[q];
      if (!PFA) {
        const int k1 = q / B, jj = q - k1 * B;
        const double2 c = P.tw1[2 * ((jj * k1) % M)];
        cA[q] = c; cB[jj * A + k1] = c;
      }
    }
    for (int j = tid; j < B; j += nthr) {
      int m = 0;
      for (int t = 0; t < A; t++) if (P.x_ovslot[t * B + j] >= 0) m |= 1 << t;
      xmask[j] = m;
    }
    Tables T; T.Tx = Tx; T.ctwA = cA; T.ctwB = cB; T.xmask = xmask; T.xov = xov; T.i1row = i1row;
    return T;
  }

  // ---------------- K1: one batch of lines, sphere -> W1 ----------------
  // The coefficients of a batch are gathered into the staging buffer ONE UNIT AHEAD: the loads of the next batch are issued
  // inside pass 2 of the current one (UG entries per lane and pass-2 iteration, their table entries one iteration earlier), so
  // that a warp never sits on the table -> coefficient -> shared-memory chain with nothing else to do (ncu r02e: 56 % of the
  // warp stalls were long-scoreboard waits with 6 warps per SM).
  static constexpr int UG = 4;
  struct Gather {                       // where the coefficients of one unit come from
    const double2* c0; const double2* c1; bool pack, has_d; int l0, e0, e1;
  };
  ABI_DEV static Gather gather_of(const XhParams& P, int b, int batch) {
    Gather g;
    g.pack = P.pack_ndat != 0;
    g.c0 = P.cg + (size_t)(g.pack ? 2 * b : b) * P.npw; g.c1 = g.c0 + P.npw;
    g.has_d = g.pack && 2 * b + 1 < P.pack_ndat;
    g.l0 = batch * GL;
    const int nl = min(GL, P.nlines - g.l0);
    g.e0 = P.estart[g.l0]; g.e1 = P.estart[g.l0 + nl];
    return g;
  }
  ABI_DEV static void g_load(const Gather& g, int2 en, double2& c, double2& d) {
    const int ipw = en.x & 0x3fffffff;
    c = g.c0[ipw];
    d = g.has_d ? g.c1[ipw] : make_double2(0.0, 0.0);
  }
  // E(G) = C(G) + i D(G), E(-G) = conj(C(G)) + i conj(D(G)) (image entries: bit 31; G = 0: bit 30, imaginary parts dropped)
  ABI_DEV static void g_store(const Gather& g, double2* stg, int2 en, double2 c, double2 d) {
    if (en.x < 0) { c.y = -c.y; d.y = -d.y; }
    if (en.x & (1 << 30)) { c.y = 0.0; d.y = 0.0; }
    stg[((en.y >> 10) - g.l0) * RS + (en.y & 1023)] = g.pack ? make_double2(c.x - d.y, c.y + d.x) : c;
  }
  ABI_DEV static void zero_stg(double2* stg) {
    ABI_FOR_LANES {
      for (int q = lane; q < STG; q += 32) stg[q] = make_double2(0.0, 0.0);
    }
    ABI_SYNCWARP();
  }
  // un-pipelined gather (first unit of a warp)
  ABI_DEV static void gather_plain(const XhParams& P, double2* stg, int b, int batch) {
    zero_stg(stg);
    const Gather g = gather_of(P, b, batch);
    ABI_FOR_LANES {
      for (int e = g.e0 + lane; e < g.e1; e += 32) {
        const int2 en = P.ent[e];
        double2 c, d;
        g_load(g, en, c, d);
        g_store(g, stg, en, c, d);
      }
    }
    ABI_SYNCWARP();
  }

  // stg holds the coefficients of (b, batch) on entry and those of (bn, batchn) on exit (when have_next)
  ABI_DEV static void forward(const XhParams& P, const Tables& T, double2* E, double2* stg, int b, int batch, bool have_next, int bn,
                              int batchn) {
    const int l0 = batch * GL, nl = min(GL, P.nlines - l0);
    Gather g;
    if (have_next) g = gather_of(P, bn, batchn);         // estart loads fly during pass 1
    else { g.e0 = g.e1 = 0; g.l0 = 0; g.pack = false; g.has_d = false; g.c0 = g.c1 = nullptr; }
    for (int w0 = 0; w0 < B * GL; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / GL, line = w - j * GL;
        if (w < B * GL && line < nl) {
          const double2* src = stg + line * RS + j;
          double2 x[A], xo[A];
#pragma unroll
          for (int t = 0; t < A; t++) x[t] = src[t * B];
          const int xm = T.xmask[j];
          if (xm) {
#pragma unroll
            for (int t = 0; t < A; t++) {
              double2 vo = x[t];
              if ((xm >> t) & 1) { const double2 h = stg[line * RS + T.xov[t * B + j]]; vo = csub(x[t], h); x[t] = cadd(x[t], h); }
              xo[t] = cmulc(vo, T.Tx[t * B + j]);
            }
          } else {
#pragma unroll
            for (int t = 0; t < A; t++) xo[t] = cmulc(x[t], T.Tx[t * B + j]);
          }
          HDft<A, +1>::run(x);
          HDft<A, +1>::run(xo);
          if (!PFA) {
#pragma unroll
            for (int k1 = 1; k1 < A; k1++) { const double2 c = T.ctwA[k1 * B + j]; x[k1] = cmulc(x[k1], c); xo[k1] = cmulc(xo[k1], c); }
          }
          double2* e = E + j * GL + line;
#pragma unroll
          for (int k1 = 0; k1 < A; k1++) { e[k1 * ZK] = x[k1]; e[(A + k1) * ZK] = xo[k1]; }
        }
      }
    }
    ABI_SYNCWARP();
    if (have_next) zero_stg(stg);                        // the staging buffer is free: it receives the next unit during pass 2
    double2* outb = P.W1 + (size_t)b * P.n1 * P.nlines + l0;
    {
      ABI_FOR_LANES {
        int2 en[UG];
        int ecur = g.e0 + lane;
#pragma unroll
        for (int u = 0; u < UG; u++) en[u] = ecur + 32 * u < g.e1 ? P.ent[ecur + 32 * u] : make_int2(0, -1);
#pragma unroll
        for (int w0 = 0; w0 < 2 * A * GL; w0 += 32) {
          // coefficients of this iteration's entries and the table entries of the next iteration: all in flight during the DFT
          double2 cv[UG], dv[UG]; int2 enn[UG];
#pragma unroll
          for (int u = 0; u < UG; u++) { cv[u] = dv[u] = make_double2(0.0, 0.0); if (en[u].y >= 0) g_load(g, en[u], cv[u], dv[u]); }
          ecur += 32 * UG;
#pragma unroll
          for (int u = 0; u < UG; u++) enn[u] = ecur + 32 * u < g.e1 ? P.ent[ecur + 32 * u] : make_int2(0, -1);
          const int w = w0 + lane;
          const int hk1 = w / GL, line = w - hk1 * GL;
          if (w < 2 * A * GL && line < nl) {
            const double2* e = E + hk1 * ZK + line;
            double2 v[B];
#pragma unroll
            for (int j = 0; j < B; j++) v[j] = e[j * GL];
            HDft<B, +1>::run(v);
            const int h = hk1 >= A ? 1 : 0, k1 = hk1 - h * A;
            // i1 = 2 kout(k1, k2) + h; kout(k1, k2) = (kout(k1, 0) + kout(0, k2)) mod M for both index maps
            const int kb = Map::kout(0, 0) + (PFA ? (k1 * (B * h_inv_mod(B % A, A))) % M : k1);
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) {
              int kk = kb + Map::kout(0, k2);
              if (PFA && kk >= M) kk -= M;
              outb[(size_t)(2 * kk + h) * P.nlines + line] = v[k2];
            }
          }
#pragma unroll
          for (int u = 0; u < UG; u++) { if (en[u].y >= 0) g_store(g, stg, en[u], cv[u], dv[u]); en[u] = enn[u]; }
        }
        // long batches: what the pass-2 iterations did not cover
        for (;;) {
          bool any = false;
#pragma unroll
          for (int u = 0; u < UG; u++) if (en[u].y >= 0) { double2 c, d; g_load(g, en[u], c, d); g_store(g, stg, en[u], c, d); any = true; }
          if (!any) break;
          ecur += 32 * UG;
#pragma unroll
          for (int u = 0; u < UG; u++) en[u] = ecur + 32 * u < g.e1 ? P.ent[ecur + 32 * u] : make_int2(0, -1);
        }
      }
    }
    ABI_SYNCWARP();
  }

  // ---------------- K3: one batch of lines (+ mirror lines), W1o -> sphere ----------------
  // The (2M x GL) strip of W1o of a batch is copied into the exchange buffer with cp.async (every 16-byte word straight to the
  // slot pass 2' reads it from: the pass then works in place), ONE UNIT AHEAD: the copy of the next strip is issued as soon as
  // pass 1' has emptied the buffer and flies during the scatter / getghc assembly of the current unit.
  ABI_DEV static void prefetch_strip(const XhParams& P, const Tables& T, double2* E, int b, int batch, unsigned long long pol) {
    const int4 bd = P.batches[batch];
    const int nl = bd.y + bd.w;
    const double2* inb = P.W1in + (size_t)b * P.n1 * P.nlines;
    ABI_FOR_LANES {
#pragma unroll 5
      for (int idx = lane; idx < 2 * M * GL; idx += 32) {
        const int row = idx / GL, line = idx - row * GL;           // row = hk1 * B + k2
        if (line < nl) {
          const int gl = line < bd.y ? bd.x + line : bd.z + (line - bd.y);
          const int hk1 = row / B, k2 = row - hk1 * B;
          // default L2 policy: a 64-byte half-run that straddles a sector shares it with the neighbouring batch, which must find it
          // in L2 (evict_first here: 1.8 GB of DRAM reads per launch instead of 1.0)
          cp_async16(E + hk1 * ZK + k2 * GL + line, inb + (size_t)T.i1row[row] * P.nlines + gl, pol);
        }
      }
    }
    cp_async_commit();
  }

  ABI_DEV static void backward(const XhParams& P, const Tables& T, double2* E, double2* stg, int b, int batch, bool have_next, int bn,
                               int batchn, unsigned long long pol) {
    const int4 bd = P.batches[batch];
    const int nl = bd.y + bd.w;
    cp_async_wait_all();
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < 2 * A * GL; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int hk1 = w / GL, line = w - hk1 * GL;
        if (w < 2 * A * GL && line < nl) {
          const int h = hk1 >= A ? 1 : 0, k1 = hk1 - h * A;
          double2* e = E + hk1 * ZK + line;
          double2 v[B];
#pragma unroll
          for (int k2 = 0; k2 < B; k2++) v[k2] = e[k2 * GL];
          HDft<B, -1>::run(v);
          e[0] = v[0];
#pragma unroll
          for (int j = 1; j < B; j++) e[j * GL] = PFA ? v[j] : cmul(v[j], T.ctwB[j * A + k1]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < B * GL; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int j = w / GL, line = w - j * GL;
        if (w < B * GL && line < nl) {
          const double2* e = E + j * GL + line;
          double2 ye[A], yo[A];
#pragma unroll
          for (int k1 = 0; k1 < A; k1++) { ye[k1] = e[k1 * ZK]; yo[k1] = e[(A + k1) * ZK]; }
          HDft<A, -1>::run(ye);
          HDft<A, -1>::run(yo);
          double2* dst = stg + line * RS + j;
          const int xm = T.xmask[j];
#pragma unroll
          for (int t = 0; t < A; t++) {
            const double2 tw = T.Tx[t * B + j];
            dst[t * B] = make_double2(fma(yo[t].x, tw.x, fma(-yo[t].y, tw.y, ye[t].x)), fma(yo[t].x, tw.y, fma(yo[t].y, tw.x, ye[t].y)));
            if ((xm >> t) & 1)
              stg[line * RS + T.xov[t * B + j]] =
                  make_double2(fma(-yo[t].x, tw.x, fma(yo[t].y, tw.y, ye[t].x)), fma(-yo[t].x, tw.y, fma(-yo[t].y, tw.x, ye[t].y)));
          }
        }
      }
    }
    ABI_SYNCWARP();
    if (have_next) prefetch_strip(P, T, E, bn, batchn, pol);
    const int e0 = P.bstart[batch], e1 = P.bstart[batch + 1];
    {
      ABI_FOR_LANES {
        if (P.pack_ndat == 0) {
          for (int e = e0 + lane; e < e1; e += 32) {
            const int2 en = P.oent[e];
            const int ipw = en.x & 0x3fffffff;
            double2 v = stg[((en.y >> 8) & 15) * RS + (en.y & 255)];
            v.x *= P.xnorm; v.y *= P.xnorm;
            if (P.zero_im_g0 && ipw == 0) v.y = 0.0;
            const size_t o = (size_t)b * P.npw + ipw;
            if (!fw_epilogue(P.epi, P.kin_filter, ipw, o, v)) v = make_double2(0.0, 0.0);
            P.out[o] = v;
          }
        } else {
          // H C(G) = [F(G) + conj(F(-G))]/2,  H D(G) = [F(G) - conj(F(-G))]/(2i); G = 0: H C = Re F, H D = Im F (both real)
          // U entries per lane and iteration: every operand of the getghc assembly (kinpw, psi, gvnlxc of both bands) is
          // requested before the first one is used, so a lane has up to 5 U independent loads in flight
          constexpr int U = 4;
          const bool has_d = 2 * b + 1 < P.pack_ndat;
          const double hn = 0.5 * P.xnorm;
          const int mode = P.epi.mode;
          const bool has_g = P.epi.gvnlxc != nullptr;
          const size_t ob = (size_t)(2 * b) * P.npw;
          for (int eb = e0 + lane; eb < e1; eb += 32 * U) {
            int2 en[U]; double kin[U]; double2 pc[U], pd[U], gc[U], gd[U];
#pragma unroll
            for (int u = 0; u < U; u++) { const int e = eb + 32 * u; en[u] = e < e1 ? P.oent[e] : make_int2(-1, 0); }
#pragma unroll
            for (int u = 0; u < U; u++) {
              const int ipw = en[u].x & 0x3fffffff;
              kin[u] = 0.0; pc[u] = pd[u] = gc[u] = gd[u] = make_double2(0.0, 0.0);
              if (en[u].x >= 0 && mode != 0) {
                kin[u] = P.epi.kinpw[ipw];
                if (mode == 1) {
                  pc[u] = P.epi.cwavef[ob + ipw];
                  if (has_d) pd[u] = P.epi.cwavef[ob + P.npw + ipw];
                  if (has_g) { gc[u] = P.epi.gvnlxc[ob + ipw]; if (has_d) gd[u] = P.epi.gvnlxc[ob + P.npw + ipw]; }
                }
              }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
              if (en[u].x < 0) continue;
              const int ipw = en[u].x & 0x3fffffff;
              const double2 f = stg[((en[u].y >> 8) & 15) * RS + (en[u].y & 255)];
              double2 vc, vd;
              if (en[u].x & (1 << 30)) {
                vc = make_double2(f.x * P.xnorm, 0.0); vd = make_double2(f.y * P.xnorm, 0.0);
              } else {
                const double2 g = stg[((en[u].y >> 20) & 15) * RS + ((en[u].y >> 12) & 255)];     // F(-G)
                vc = make_double2((f.x + g.x) * hn, (f.y - g.y) * hn);
                vd = make_double2((f.y + g.y) * hn, (g.x - f.x) * hn);
              }
              const size_t oc = ob + ipw, od = oc + P.npw;
              // getghc assembly (m_getghc.F90:1266-1280; type_calc = 1 filter :1003-1031), same arithmetic as fw_epilogue
              bool keep = true;
              if (mode == 1) {
                if (kin[u] < P.kin_filter) {
                  vc.x = vc.x + kin[u] * pc[u].x; vc.y = vc.y + kin[u] * pc[u].y;
                  vd.x = vd.x + kin[u] * pd[u].x; vd.y = vd.y + kin[u] * pd[u].y;
                  if (has_g) { vc.x += gc[u].x; vc.y += gc[u].y; vd.x += gd[u].x; vd.y += gd[u].y; }
                } else {
                  keep = false;
                  if (P.epi.gsc) { P.epi.gsc[oc] = make_double2(0.0, 0.0); if (has_d) P.epi.gsc[od] = make_double2(0.0, 0.0); }
                }
              } else if (mode == 2) {
                if (kin[u] > P.kin_filter) keep = false;
              }
              if (!keep) { vc = make_double2(0.0, 0.0); vd = vc; }
              P.out[oc] = vc;
              if (has_d) P.out[od] = vd;
            }
          }
        }
      }
    }
    ABI_SYNCWARP();
  }
};

template <int A, int B, int GL> ABI_HD constexpr size_t xh_smem_bytes(int warps) {
  using F = XHalf<A, B, GL>;
  return sizeof(double2) * ((size_t)F::TW_SLOTS + F::int_slots() + (size_t)warps * F::WSIZE);
}

// DIR 0: K1 (forward), 1: K3 (backward).  Units (batch, transform) are dealt to the warps of the grid round-robin, the transform
// index fastest (the entry tables of a batch are then reused from L1/L2 by the next nb units).
template <int A, int B, int GL, int WARPS, int DIR>
__global__ void __launch_bounds__(WARPS * 32) k_xh(XhParams P) {
  using F = XHalf<A, B, GL>;
  ABI_DYN_SMEM(double2, sm);
#ifdef ABI_EMU
  const int warp = 0, tid = 0, nthr = 1;
  const long long gw = blockIdx.x, nw = gridDim.x;
#else
  const int warp = threadIdx.x >> 5, tid = threadIdx.x, nthr = WARPS * 32;
  const long long gw = (long long)blockIdx.x * WARPS + warp, nw = (long long)gridDim.x * WARPS;
#endif
  const typename F::Tables T = F::load_tables(sm, P, tid, nthr);
  double2* E = sm + F::TW_SLOTS + F::int_slots() + (size_t)warp * F::WSIZE;
  double2* stg = E + F::ESIZE;
  __syncthreads();
  const long long nunits = (long long)P.nbatch * P.nb;
  auto decode = [&](long long unit, int& b, int& batch) {
    if ((P.order & 1) == 0) { batch = (int)(unit / P.nb); b = (int)(unit - (long long)batch * P.nb); }
    else { b = (int)(unit / P.nbatch); batch = (int)(unit - (long long)b * P.nbatch); }
  };
  const unsigned long long pol = (P.order & 2) ? policy_evict_first() : policy_evict_normal();
  int b = 0, batch = 0;
  if (gw < nunits) {
    decode(gw, b, batch);
    if (DIR == 0) F::gather_plain(P, stg, b, batch);
    else F::prefetch_strip(P, T, E, b, batch, pol);
  }
  for (long long unit = gw; unit < nunits; unit += nw) {
    const bool have_next = unit + nw < nunits;
    int bn = 0, batchn = 0;
    if (have_next) decode(unit + nw, bn, batchn);
    if (DIR == 0) F::forward(P, T, E, stg, b, batch, have_next, bn, batchn);
    else F::backward(P, T, E, stg, b, batch, have_next, bn, batchn, pol);
    b = bn; batch = batchn;
  }
  cp_async_wait_all();
}

// ---- host interface (x_stage.cu) ----
// true when the half-support x kernels can run this plan (n1 even with an instantiated configuration, at most kHalfOV indices r
// that occur both as i1 = r and i1 = r + n1/2); 'packed': Gamma-point two-band transforms (output lines = input lines)
bool x_stage_usable(const FourwfPlan& pl, bool packed);
void x_stage_forward(const FourwfPlan& pl, const double2* cg, double2* W1, int nb, int pack_ndat, cudaStream_t st);
void x_stage_backward(const FourwfPlan& pl, const double2* W1o, double2* out, int nb, int pack_ndat, double xnorm, int zero_im_g0,
                      const FourwfEpilogue& epi, double kin_filter, cudaStream_t st);

}  // namespace abi
