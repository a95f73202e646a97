// Plane stage of the fused fourwf (option 2): for one (band, i1) yz-plane
//     y FFT (e^{+i}) on the occupied z planes -> z FFT on every column -> * V_loc -> z FFT^-1 -> y FFT^-1
// with every 1-D transform done as TWO register-resident radix passes (n = R1*R2, R1,R2 <= 16) exchanged once
// through shared memory by a single warp (no block barrier inside a phase):
//
//   forward  X[k1 + R1 k2] = sum_j W_R2^{j k2} { W_n^{j k1} [ sum_t x[j + R2 t] W_R1^{t k1} ] }      (A then B)
//   inverse  x[j + R2 t]   = sum_k1 W_R1^{-t k1} { W_n^{-j k1} [ sum_k2 X[k1 + R1 k2] W_R2^{-j k2} ] } (B' then A')
//
// so the z pass is  A -> smem -> B, *V, B' -> smem -> A'  (the spectrum never leaves registers around V_loc) and each
// element crosses shared memory twice per transform instead of twice per radix pass.  Global traffic of a phase is
// register <-> L2 directly, with the lane order chosen per phase so that a warp touches whole 128-byte lines:
//   y phases: lanes run over (line u, j) with j fastest   (rows of the compact disc / of S are i2-contiguous)
//   z phase : lanes run over (j, column i2) with i2 fastest (G consecutive columns of S / V_loc)
// The (nU x n2) plane S between the phases lives in a per-CTA L2 scratch (stcg/ldcg); it is never written to HBM
// on purpose (38-77 MB for the whole grid, below the 126 MB L2).
//
// Reference semantics: the zero-padded passes of src/52_fft_mpi_noabirule/fftw3_fftpad.finc:14-196 and the
// cache-blocked per-plane variant fftw3_fftrisc.finc; V_loc application src/44_abitools/m_cgtools.F90:2410-2491.
#pragma once
#include "fft_engine.cuh"

namespace abi {

// Good-Thomas prime-factor DFT for coprime A, B: no internal twiddles; index maps are compile-time constants.
template <int A, int B, int SIGN> struct DftPFA {
  static constexpr int N = A * B;
  ABI_HD static constexpr int inv_mod(int a, int m) { int r = 1; for (int i = 1; i < m; i++) if ((a * i) % m == 1) r = i; return r; }
  ABI_HD static void run(double2* x) {
    constexpr int bi = inv_mod(B % A, A), ai = inv_mod(A % B, B);
    double2 y[B][A];
#pragma unroll
    for (int n2 = 0; n2 < B; n2++) {
      double2 col[A];
#pragma unroll
      for (int n1 = 0; n1 < A; n1++) col[n1] = x[(B * n1 + A * n2) % N];
      Dft<A, SIGN>::run(col);
#pragma unroll
      for (int k1 = 0; k1 < A; k1++) y[n2][k1] = col[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < A; k1++) {
      double2 row[B];
#pragma unroll
      for (int n2 = 0; n2 < B; n2++) row[n2] = y[n2][k1];
      Dft<B, SIGN>::run(row);
#pragma unroll
      for (int k2 = 0; k2 < B; k2++) x[(B * bi * k1 + A * ai * k2) % N] = row[k2];
    }
  }
};
template <int SIGN> struct Dft<10, SIGN> : DftPFA<2, 5, SIGN> {};
template <int SIGN> struct Dft<12, SIGN> : DftPFA<3, 4, SIGN> {};
template <int SIGN> struct Dft<14, SIGN> : DftPFA<2, 7, SIGN> {};
template <int SIGN> struct Dft<15, SIGN> : DftPFA<3, 5, SIGN> {};

struct PlaneParams {
  int n1, n2, n3, nb, nU, cplex;
  int za, zla, zb, zlb;             // occupied z planes: i3 in [za,za+zla) -> u = i3-za ; i3 in [zb,zb+zlb) -> u = zla + i3-zb
  int nlin, nlout;                  // lines per plane of W1 / W1o
  long long nunits;                 // nb * n1
  const double2* W1; double2* W1o;  // [b][i1][line]
  double2* S;                       // [gridDim.x][nU][n2] L2 scratch
  const double* vT;                 // V_loc as [i1][i3][i2] (cplex doubles per point)
  const double2* tw;                // exp(-2 pi i j / n2)
  const double2* tw3 = nullptr;     // exp(-2 pi i j / n3) (split path, n2 != n3)
  // per occupied plane u: the rows of W1 / W1o hold i2 in [a, a+la) then [b, b+lb), starting at line start[u];
  // runs[u] = {a, la, b, lb}
  const int* in_start; const short4* in_runs;
  const int* out_start; const short4* out_runs;
  // option 1 (density accumulation, k_fw_plane_rho): rhoT[i1][i3][i2] += wxy[b].x Re(psi)^2 + wxy[b].y Im(psi)^2
  double* rhoT = nullptr; const double2* wxy = nullptr;
};

#ifdef ABI_EMU
#define ABI_FOR_LANES for (int lane = 0; lane < 32; lane++)
#define ABI_SYNCWARP()
#define ABI_RED_ADD(p, v) (*(p) += (v))
#else
#define ABI_RED_ADD(p, v) atomicAdd((p), (v))
#define ABI_FOR_LANES const int lane = threadIdx.x & 31;
#define ABI_SYNCWARP() __syncwarp()
#endif

template <int R1, int R2, int G>
struct PlaneFft {
  static constexpr int N = R1 * R2;
  // y-mode exchange buffer: E[line][k1][j], j fastest, k1 stride odd (conflict-free reads at fixed j)
  static constexpr int YK = (R2 % 2) ? R2 : R2 + 1;
  static constexpr int YL = R1 * YK;
  // z-mode exchange buffer: E[k1][j][line], line fastest; for G=4 two k1 share a 128-byte wavefront -> stride = 4 mod 8
  static constexpr int ZK = (G >= 8) ? R2 * G : ((R2 * G) % 8 == 4 ? R2 * G : R2 * G + 4);
  static constexpr int ESIZE = (G * YL > R1 * ZK) ? G * YL : R1 * ZK;      // double2 per warp
  static constexpr int ZOFF = (N * 4 + 15) / 16;                           // double2 slots of the i3 -> plane offset table
  // Inter-pass twiddles w^(j k1), w = exp(-2 pi i / n), as TWO shared-memory tables so that a warp's lanes always read
  // consecutive 16-byte words (no bank conflicts, offsets are compile-time immediates, no index multiply):
  //   twA[k1 * R2 + j]  -- forward steps: lanes run over j at a fixed k1 per instruction
  //   twB[j * R1 + k1]  -- inverse steps: lanes run over k1 at a fixed j per instruction
  static constexpr int TWSIZE = 2 * N;                                     // double2 slots of both tables
  ABI_DEV static void load_tw(double2* tws, const double2* __restrict__ twg, int tid, int nthr) {
    for (int q = tid; q < N; q += nthr) {
      const int k1 = q / R2, j = q - k1 * R2;
      const double2 w = twg[j * k1];
      tws[q] = w; tws[N + j * R1 + k1] = w;
    }
  }
  ABI_DEV static double2 twf(const double2* tw, int k1, int j, double2 v) { return cmulc(v, tw[k1 * R2 + j]); }     // * e^{+2 pi i j k1/n}
  ABI_DEV static double2 twb(const double2* tw, int k1, int j, double2 v) { return cmul(v, tw[N + j * R1 + k1]); }  // * e^{-2 pi i j k1/n}

  // ---------------- phase Y: compact disc rows of W1 -> S[u][i2] ----------------
  ABI_DEV static void phase_y(const PlaneParams& P, const double2* __restrict__ w1, double2* __restrict__ S, double2* E,
                              const double2* tw, int u0) {
    const int nl = min(G, P.nU - u0);
    for (int w0 = 0; w0 < nl * R2; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        if (w < nl * R2) {
          const int line = w / R2, j = w - line * R2, u = u0 + line;
          const short4 rr = P.in_runs[u];
          const double2* src = w1 + P.in_start[u];
          double2 x[R1];
#pragma unroll
          for (int t = 0; t < R1; t++) {
            const int i2 = j + R2 * t;
            x[t] = make_double2(0.0, 0.0);
            if ((unsigned)(i2 - rr.x) < (unsigned)rr.y) x[t] = src[i2 - rr.x];
            else if ((unsigned)(i2 - rr.z) < (unsigned)rr.w) x[t] = src[rr.y + i2 - rr.z];
          }
          Dft<R1, +1>::run(x);
          double2* e = E + line * YL + j;
          e[0] = x[0];
#pragma unroll
          for (int k1 = 1; k1 < R1; k1++) e[k1 * YK] = twf(tw, k1, j, x[k1]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < nl * R1; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        if (w < nl * R1) {
          const int line = w / R1, k1 = w - line * R1, u = u0 + line;
          const double2* e = E + line * YL + k1 * YK;
          double2 v[R2];
#pragma unroll
          for (int j = 0; j < R2; j++) v[j] = e[j];
          Dft<R2, +1>::run(v);
          double2* dst = S + (size_t)u * P.n2 + k1;
#pragma unroll
          for (int k2 = 0; k2 < R2; k2++) stcg2(dst + R1 * k2, v[k2]);
        }
      }
    }
    ABI_SYNCWARP();
  }

  // ---------------- phase Z: columns of S -> z FFT, * V_loc, z FFT^-1 -> S (in place) ----------------
  ABI_DEV static int u_of_i3(const PlaneParams& P, int i3) {
    return (unsigned)(i3 - P.za) < (unsigned)P.zla ? i3 - P.za : ((unsigned)(i3 - P.zb) < (unsigned)P.zlb ? P.zla + i3 - P.zb : -1);
  }

  ABI_DEV static void phase_z(const PlaneParams& P, double2* __restrict__ S, const double* __restrict__ vplane, double2* E,
                              const double2* tw, const int* __restrict__ zoff, int c0) {
    const int nl = min(G, P.n2 - c0);
    const int n2 = P.n2;
    for (int w0 = 0; w0 < G * R2; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w % G, j = w / G;
        if (w < G * R2 && line < nl) {
          const double2* src = S + c0 + line;
          double2 x[R1];
#pragma unroll
          for (int t = 0; t < R1; t++) {
            const int o = zoff[j + R2 * t];                      // u * n2 of the occupied plane, -1 for a zero-padded one
            x[t] = (o >= 0) ? ldcg2(src + o) : make_double2(0.0, 0.0);
          }
          Dft<R1, +1>::run(x);
          double2* e = E + j * G + line;
          e[0] = x[0];
#pragma unroll
          for (int k1 = 1; k1 < R1; k1++) e[k1 * ZK] = twf(tw, k1, j, x[k1]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < G * R1; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w % G, k1 = w / G;
        if (w < G * R1 && line < nl) {
          double2* e = E + k1 * ZK + line;
          // V_loc of this lane's R2 grid points first: the loads fly while the exchange buffer is read and transformed
          double2 vv[R2];
          if (P.cplex == 1) {
            const double* vp = vplane + (size_t)k1 * n2 + c0 + line;
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) vv[k2] = make_double2(ldg1(vp + (size_t)(R1 * k2) * n2), 0.0);
          } else {
            const double2* vp = reinterpret_cast<const double2*>(vplane) + (size_t)k1 * n2 + c0 + line;
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) vv[k2] = ldg2(vp + (size_t)(R1 * k2) * n2);
          }
          double2 v[R2];
#pragma unroll
          for (int j = 0; j < R2; j++) v[j] = e[j * G];
          Dft<R2, +1>::run(v);
          if (P.cplex == 1) {
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) { v[k2].x *= vv[k2].x; v[k2].y *= vv[k2].x; }
          } else {
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) v[k2] = cmul(v[k2], vv[k2]);
          }
          Dft<R2, -1>::run(v);
          e[0] = v[0];
#pragma unroll
          for (int j = 1; j < R2; j++) e[j * G] = twb(tw, k1, j, v[j]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < G * R2; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w % G, j = w / G;
        if (w < G * R2 && line < nl) {
          const double2* e = E + j * G + line;
          double2 y[R1];
#pragma unroll
          for (int k1 = 0; k1 < R1; k1++) y[k1] = e[k1 * ZK];
          Dft<R1, -1>::run(y);
          double2* dst = S + c0 + line;
#pragma unroll
          for (int t = 0; t < R1; t++) {
            const int o = zoff[j + R2 * t];
            if (o >= 0) stcg2(dst + o, y[t]);
          }
        }
      }
    }
    ABI_SYNCWARP();
  }

  // ---------------- phase Z (option 1): columns of S -> z FFT -> rhoT += w |psi(r)|^2 ----------------
  // fourwf option 1 (src/53_ffts/m_fft.F90:2633-2653, cg_addtorho src/44_abitools/m_cgtools.F90:2338-2384) on the same
  // register-resident z transform; the plane of rhoT is i2-contiguous so a warp's reductions fill whole sectors.
  ABI_DEV static void phase_z_rho(const PlaneParams& P, const double2* __restrict__ S, double* __restrict__ rplane, double2 wxy,
                                  double2* E, const double2* tw, const int* __restrict__ zoff, int c0) {
    const int nl = min(G, P.n2 - c0);
    const int n2 = P.n2;
    for (int w0 = 0; w0 < G * R2; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w % G, j = w / G;
        if (w < G * R2 && line < nl) {
          const double2* src = S + c0 + line;
          double2 x[R1];
#pragma unroll
          for (int t = 0; t < R1; t++) {
            const int o = zoff[j + R2 * t];                      // u * n2 of the occupied plane, -1 for a zero-padded one
            x[t] = (o >= 0) ? ldcg2(src + o) : make_double2(0.0, 0.0);
          }
          Dft<R1, +1>::run(x);
          double2* e = E + j * G + line;
          e[0] = x[0];
#pragma unroll
          for (int k1 = 1; k1 < R1; k1++) e[k1 * ZK] = twf(tw, k1, j, x[k1]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < G * R1; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        const int line = w % G, k1 = w / G;
        if (w < G * R1 && line < nl) {
          const double2* e = E + k1 * ZK + line;
          double2 v[R2];
#pragma unroll
          for (int j = 0; j < R2; j++) v[j] = e[j * G];
          Dft<R2, +1>::run(v);
          double* rp = rplane + (size_t)k1 * n2 + c0 + line;
#pragma unroll
          for (int k2 = 0; k2 < R2; k2++) ABI_RED_ADD(rp + (size_t)(R1 * k2) * n2, wxy.x * v[k2].x * v[k2].x + wxy.y * v[k2].y * v[k2].y);
        }
      }
    }
    ABI_SYNCWARP();
  }

  // ---------------- phase Y': S[u][i2] -> y FFT^-1 -> compact output rows of W1o ----------------
  ABI_DEV static void phase_yinv(const PlaneParams& P, const double2* __restrict__ S, double2* __restrict__ w1o, double2* E,
                                 const double2* tw, int u0) {
    const int nl = min(G, P.nU - u0);
    for (int w0 = 0; w0 < nl * R1; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        if (w < nl * R1) {
          const int line = w / R1, k1 = w - line * R1, u = u0 + line;
          const double2* src = S + (size_t)u * P.n2 + k1;
          double2 v[R2];
#pragma unroll
          for (int k2 = 0; k2 < R2; k2++) v[k2] = ldcg2(src + R1 * k2);
          Dft<R2, -1>::run(v);
          double2* e = E + line * YL + k1 * YK;
          e[0] = v[0];
#pragma unroll
          for (int j = 1; j < R2; j++) e[j] = twb(tw, k1, j, v[j]);
        }
      }
    }
    ABI_SYNCWARP();
    for (int w0 = 0; w0 < nl * R2; w0 += 32) {
      ABI_FOR_LANES {
        const int w = w0 + lane;
        if (w < nl * R2) {
          const int line = w / R2, j = w - line * R2, u = u0 + line;
          const short4 rr = P.out_runs[u];
          const double2* e = E + line * YL + j;
          double2 y[R1];
#pragma unroll
          for (int k1 = 0; k1 < R1; k1++) y[k1] = e[k1 * YK];
          Dft<R1, -1>::run(y);
          double2* dst = w1o + P.out_start[u];
#pragma unroll
          for (int t = 0; t < R1; t++) {
            const int i2 = j + R2 * t;
            if ((unsigned)(i2 - rr.x) < (unsigned)rr.y) dst[i2 - rr.x] = y[t];
            else if ((unsigned)(i2 - rr.z) < (unsigned)rr.w) dst[rr.y + i2 - rr.z] = y[t];
          }
        }
      }
    }
    ABI_SYNCWARP();
  }
};

// one CTA = one (band, i1) plane at a time; warps take line batches round-robin inside each phase
template <int R1, int R2, int G, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS >= 16 ? 1 : 2)) k_fw_plane(PlaneParams P) {
  using F = PlaneFft<R1, R2, G>;
  ABI_DYN_SMEM(double2, sm);
  double2* tw = sm;                                    // twA | twB (PlaneFft::load_tw)
#ifdef ABI_EMU
  const int warp = 0, nwarps = 1;
  F::load_tw(tw, P.tw, 0, 1);
#else
  const int warp = threadIdx.x >> 5, nwarps = WARPS;
  F::load_tw(tw, P.tw, threadIdx.x, WARPS * 32);
#endif
  int* zoff = reinterpret_cast<int*>(sm + F::TWSIZE);  // N ints: i3 -> u(i3) * n2, or -1 (built once per CTA)
#ifdef ABI_EMU
  for (int i3 = 0; i3 < F::N; i3++) { const int u = F::u_of_i3(P, i3); zoff[i3] = u >= 0 ? u * P.n2 : -1; }
#else
  for (int i3 = threadIdx.x; i3 < F::N; i3 += WARPS * 32) { const int u = F::u_of_i3(P, i3); zoff[i3] = u >= 0 ? u * P.n2 : -1; }
#endif
  double2* E = sm + F::TWSIZE + F::ZOFF + (size_t)warp * F::ESIZE;
  double2* S = P.S + (size_t)blockIdx.x * P.nU * P.n2;
  __syncthreads();
  for (long long unit = blockIdx.x; unit < P.nunits; unit += gridDim.x) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
    double2* w1o = P.W1o + ((size_t)b * P.n1 + i1) * P.nlout;
    const double* vplane = P.vT + (size_t)P.cplex * i1 * P.n3 * P.n2;
    for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_y(P, w1, S, E, tw, u0);
    __syncthreads();
    for (int c0 = warp * G; c0 < P.n2; c0 += nwarps * G) F::phase_z(P, S, vplane, E, tw, zoff, c0);
    __syncthreads();
    for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_yinv(P, S, w1o, E, tw, u0);
    __syncthreads();
  }
}

// option 1: one CTA = one (band or band pair, i1) plane: y FFT, z FFT, density accumulation (no way back)
template <int R1, int R2, int G, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS >= 16 ? 1 : 2)) k_fw_plane_rho(PlaneParams P) {
  using F = PlaneFft<R1, R2, G>;
  ABI_DYN_SMEM(double2, sm);
  double2* tw = sm;
#ifdef ABI_EMU
  const int warp = 0, nwarps = 1;
  F::load_tw(tw, P.tw, 0, 1);
#else
  const int warp = threadIdx.x >> 5, nwarps = WARPS;
  F::load_tw(tw, P.tw, threadIdx.x, WARPS * 32);
#endif
  int* zoff = reinterpret_cast<int*>(sm + F::TWSIZE);  // N ints: i3 -> u(i3) * n2, or -1 (built once per CTA)
#ifdef ABI_EMU
  for (int i3 = 0; i3 < F::N; i3++) { const int u = F::u_of_i3(P, i3); zoff[i3] = u >= 0 ? u * P.n2 : -1; }
#else
  for (int i3 = threadIdx.x; i3 < F::N; i3 += WARPS * 32) { const int u = F::u_of_i3(P, i3); zoff[i3] = u >= 0 ? u * P.n2 : -1; }
#endif
  double2* E = sm + F::TWSIZE + F::ZOFF + (size_t)warp * F::ESIZE;
  double2* S = P.S + (size_t)blockIdx.x * P.nU * P.n2;
  __syncthreads();
  for (long long unit = blockIdx.x; unit < P.nunits; unit += gridDim.x) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
    double* rplane = P.rhoT + (size_t)i1 * P.n3 * P.n2;
    const double2 wxy = P.wxy[b];
    for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_y(P, w1, S, E, tw, u0);
    __syncthreads();
    for (int c0 = warp * G; c0 < P.n2; c0 += nwarps * G) F::phase_z_rho(P, S, rplane, wxy, E, tw, zoff, c0);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Split plane stage for n2 != n3 (non-cubic boxes): the same phases as three kernels, each templated on ONE length, with the
// S planes of all (band, i1) units of the chunk in global memory (S[unit][nU][n2]) between them.  Costs two more round
// trips of S than the fused kernel but keeps every transform in registers; kind: 0 = y, 1 = z (* V_loc), 2 = y^-1,
// 3 = z + density accumulation (option 1).
// ---------------------------------------------------------------------------------------------------------
template <int R1, int R2, int G, int WARPS, int KIND>
__global__ void __launch_bounds__(WARPS * 32, 2) k_fw_plane_split(PlaneParams P) {
  using F = PlaneFft<R1, R2, G>;
  ABI_DYN_SMEM(double2, sm);
  double2* tw = sm;                                    // twA | twB of the length this kernel transforms
  const double2* twg = (KIND == 1 || KIND == 3) ? P.tw3 : P.tw;
#ifdef ABI_EMU
  const int warp = 0, nwarps = 1;
  F::load_tw(tw, twg, 0, 1);
#else
  const int warp = threadIdx.x >> 5, nwarps = WARPS;
  F::load_tw(tw, twg, threadIdx.x, WARPS * 32);
#endif
  int* zoff = reinterpret_cast<int*>(sm + F::TWSIZE);
  if (KIND == 1 || KIND == 3) {
#ifdef ABI_EMU
    for (int i3 = 0; i3 < F::N; i3++) { const int u = F::u_of_i3(P, i3); zoff[i3] = u >= 0 ? u * P.n2 : -1; }
#else
    for (int i3 = threadIdx.x; i3 < F::N; i3 += WARPS * 32) { const int u = F::u_of_i3(P, i3); zoff[i3] = u >= 0 ? u * P.n2 : -1; }
#endif
  }
  double2* E = sm + F::TWSIZE + F::ZOFF + (size_t)warp * F::ESIZE;
  __syncthreads();
  for (long long unit = blockIdx.x; unit < P.nunits; unit += gridDim.x) {
    const int i1 = (int)(unit / P.nb), b = (int)(unit - (long long)i1 * P.nb);
    double2* S = P.S + (size_t)unit * P.nU * P.n2;
    if (KIND == 0) {
      const double2* w1 = P.W1 + ((size_t)b * P.n1 + i1) * P.nlin;
      for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_y(P, w1, S, E, tw, u0);
    } else if (KIND == 1) {
      const double* vplane = P.vT + (size_t)P.cplex * i1 * P.n3 * P.n2;
      for (int c0 = warp * G; c0 < P.n2; c0 += nwarps * G) F::phase_z(P, S, vplane, E, tw, zoff, c0);
    } else if (KIND == 2) {
      double2* w1o = P.W1o + ((size_t)b * P.n1 + i1) * P.nlout;
      for (int u0 = warp * G; u0 < P.nU; u0 += nwarps * G) F::phase_yinv(P, S, w1o, E, tw, u0);
    } else {
      double* rplane = P.rhoT + (size_t)i1 * P.n3 * P.n2;
      const double2 wxy = P.wxy[b];
      for (int c0 = warp * G; c0 < P.n2; c0 += nwarps * G) F::phase_z_rho(P, S, rplane, wxy, E, tw, zoff, c0);
    }
  }
}

// host interface (plane_stage.cu)
bool plane_stage_supported(int n);
// launches the plane stage for the (n2, n3) of P; P.S is provided by the launcher: the per-CTA L2 scratch of the fused kernel
// when n2 == n3, the S planes of all units (global memory) and the three split kernels otherwise
void plane_stage_launch(PlaneParams& P, cudaStream_t st);
// same for option 1 (P.rhoT / P.wxy set): k_fw_plane_rho, or the y and z+density kernels of the split path
void plane_stage_launch_rho(PlaneParams& P, cudaStream_t st);
void plane_stage_release();

}  // namespace abi
