// Shared-memory batched 1-D complex FP64 FFT engine for sm_100a (hand-written; no cuFFT).
//
// A CTA holds `nlines` lines of length n in shared memory (line stride `lstride` double2, odd so that the
// 16-byte accesses of 8 consecutive lines fall in 8 different bank groups) and runs in-place mixed-radix
// passes over all of them.  Work items are (line, butterfly) pairs with the LINE index fastest across
// threads: a warp touches 32 different lines at the same in-line offset -> conflict-free shared-memory
// access for every pass and a warp-uniform (broadcast) twiddle load.
//
//   DIF passes: natural order in  -> digit-reversed order out   (factors applied first..last)
//   DIT passes: digit-reversed in -> natural order out          (factors applied last..first)
// The permutation is never executed: loads/stores that feed a transform place/read elements through the
// pos<->idx tables of the plan (the scatter/gather of fourwf does an indexed copy anyway), and the z pass of
// the fused V_loc stage goes natural -DIF-> reversed, multiplies V at permuted addresses, -DIT-> natural.
//
// Radix butterflies 2,3,4,5,7 (direct) and 6,8,9,16 (in-register Cooley-Tukey) cover every length
// 2^a 3^b 5^c 7^d in at most 4 passes up to n=1024.
#pragma once
#include "common.cuh"

namespace abi {

#include "roots.inc"

ABI_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
ABI_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
ABI_HD double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
ABI_HD double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
template <int SIGN> ABI_HD double2 mul_si(double2 a) {  // a * (SIGN * i)
  return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
// a * exp(SIGN * 2 pi i * k / R) with compile-time (R, k)
template <int R, int K, int SIGN> ABI_HD double2 mul_root(double2 a) {
  constexpr int k = ((K % R) + R) % R;
  if constexpr (k == 0) return a;
  else if constexpr (4 * k == R) return mul_si<SIGN>(a);
  else if constexpr (2 * k == R) return make_double2(-a.x, -a.y);
  else if constexpr (4 * k == 3 * R) return mul_si<-SIGN>(a);
  else {
    constexpr double c = root_c<R>(k);
    constexpr double s = SIGN * root_s<R>(k);
    return make_double2(fma(a.x, c, -a.y * s), fma(a.x, s, a.y * c));
  }
}

template <int R, int SIGN> struct Dft;

template <int SIGN> struct Dft<2, SIGN> {
  ABI_HD static void run(double2* x) {
    double2 a = x[0], b = x[1];
    x[0] = cadd(a, b); x[1] = csub(a, b);
  }
};
template <int SIGN> struct Dft<4, SIGN> {
  ABI_HD static void run(double2* x) {
    double2 a = cadd(x[0], x[2]), b = csub(x[0], x[2]);
    double2 c = cadd(x[1], x[3]), d = mul_si<SIGN>(csub(x[1], x[3]));
    x[0] = cadd(a, c); x[1] = cadd(b, d); x[2] = csub(a, c); x[3] = csub(b, d);
  }
};
// odd prime radix: X[k], X[R-k] = P_k +- SIGN*i*Q_k with P_k = x0 + sum_t cos(2 pi k t/R) (x_t + x_{R-t}),
// Q_k = sum_t sin(2 pi k t/R) (x_t - x_{R-t})
template <int R, int SIGN> struct DftPrime {
  ABI_HD static void run(double2* x) {
    constexpr int H = (R - 1) / 2;
    double2 a[H], b[H];
    double2 s0 = x[0];
#pragma unroll
    for (int t = 1; t <= H; t++) {
      a[t - 1] = cadd(x[t], x[R - t]); b[t - 1] = csub(x[t], x[R - t]);
      s0 = cadd(s0, a[t - 1]);
    }
    double2 x0 = x[0];
    x[0] = s0;
#pragma unroll
    for (int k = 1; k <= H; k++) {
      double2 p = x0, q = make_double2(0.0, 0.0);
#pragma unroll
      for (int t = 1; t <= H; t++) {
        const double c = root_c<R>((k * t) % R), s = root_s<R>((k * t) % R);
        p.x = fma(c, a[t - 1].x, p.x); p.y = fma(c, a[t - 1].y, p.y);
        q.x = fma(s, b[t - 1].x, q.x); q.y = fma(s, b[t - 1].y, q.y);
      }
      double2 iq = mul_si<SIGN>(q);
      x[k] = cadd(p, iq); x[R - k] = csub(p, iq);
    }
  }
};
template <int SIGN> struct Dft<3, SIGN> : DftPrime<3, SIGN> {};
template <int SIGN> struct Dft<5, SIGN> : DftPrime<5, SIGN> {};
template <int SIGN> struct Dft<7, SIGN> : DftPrime<7, SIGN> {};

// composite R = A*B, decimation in frequency inside registers:
//   y_{k0}[j0] = (sum_{j1} x[j1*B+j0] w_A^{j1 k0}) * w_R^{j0 k0};  X[A*k1+k0] = sum_{j0} y_{k0}[j0] w_B^{j0 k1}
template <int A, int B, int SIGN> struct DftComposite {
  template <int J0, int K0> ABI_HD static void tw_k0(double2 (*y)[B], const double2* col) {
    y[K0][J0] = mul_root<A * B, J0 * K0, SIGN>(col[K0]);
    if constexpr (K0 + 1 < A) tw_k0<J0, K0 + 1>(y, col);
  }
  template <int J0> ABI_HD static void cols(double2 (*y)[B], const double2* x) {
    double2 col[A];
#pragma unroll
    for (int j1 = 0; j1 < A; j1++) col[j1] = x[j1 * B + J0];
    Dft<A, SIGN>::run(col);
    tw_k0<J0, 0>(y, col);
    if constexpr (J0 + 1 < B) cols<J0 + 1>(y, x);
  }
  ABI_HD static void run(double2* x) {
    double2 y[A][B];
    cols<0>(y, x);
#pragma unroll
    for (int k0 = 0; k0 < A; k0++) {
      Dft<B, SIGN>::run(y[k0]);
#pragma unroll
      for (int k1 = 0; k1 < B; k1++) x[A * k1 + k0] = y[k0][k1];
    }
  }
};
template <int SIGN> struct Dft<6, SIGN> : DftComposite<2, 3, SIGN> {};
template <int SIGN> struct Dft<8, SIGN> : DftComposite<2, 4, SIGN> {};
template <int SIGN> struct Dft<9, SIGN> : DftComposite<3, 3, SIGN> {};
template <int SIGN> struct Dft<16, SIGN> : DftComposite<4, 4, SIGN> {};

// ---------------------------------------------------------------------------------------------------------
// Plan for one length (host builds it; kernels receive it by value / from constant memory)
struct Fft1d {
  int n;
  int nfac;
  int radix[8];            // DIF order
  const double2* tw;       // device: tw[j] = exp(-2 pi i j / n), j < n
  const unsigned short* pos_of_idx;  // device: position (after DIF / before DIT) of natural index
  const unsigned short* idx_of_pos;  // device: inverse table
};

// one in-place pass over nlines lines; DIT=false: butterfly then twiddle; DIT=true: twiddle then butterfly
template <int R, int SIGN, bool DIT>
ABI_DEV void fft_pass(double2* buf, int lstride, int nlines, int n, int blk, const double2* tw,
                             int tid, int nthr) {
  const int m = blk / R;
  const int tstep = n / blk;
  const int total = (n / R) * nlines;
  for (int w = tid; w < total; w += nthr) {
    const int q = w / nlines;
    const int line = w - q * nlines;
    const int bi = q / m;
    const int j = q - bi * m;
    double2* p = buf + (size_t)line * lstride + bi * blk + j;
    double2 x[R];
#pragma unroll
    for (int t = 0; t < R; t++) x[t] = p[t * m];
    if (DIT && j != 0) {
#pragma unroll
      for (int t = 1; t < R; t++) {
        const double2 wv = tw[t * j * tstep];
        x[t] = SIGN > 0 ? cmulc(x[t], wv) : cmul(x[t], wv);
      }
    }
    Dft<R, SIGN>::run(x);
    if (!DIT && j != 0) {
#pragma unroll
      for (int t = 1; t < R; t++) {
        const double2 wv = tw[t * j * tstep];
        x[t] = SIGN > 0 ? cmulc(x[t], wv) : cmul(x[t], wv);
      }
    }
#pragma unroll
    for (int t = 0; t < R; t++) p[t * m] = x[t];
  }
}

template <int SIGN, bool DIT>
ABI_DEV void fft_pass_any(int r, double2* buf, int lstride, int nlines, int n, int blk,
                                 const double2* tw, int tid, int nthr) {
  switch (r) {
    case 2: fft_pass<2, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    case 3: fft_pass<3, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    case 4: fft_pass<4, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    case 5: fft_pass<5, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    case 6: fft_pass<6, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    case 7: fft_pass<7, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    case 8: fft_pass<8, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    case 9: fft_pass<9, SIGN, DIT>(buf, lstride, nlines, n, blk, tw, tid, nthr); break;
    default: break;
  }
}

// natural -> digit-reversed. Ends with a block barrier.
template <int SIGN>
ABI_DEV void fft_lines_dif(double2* buf, int lstride, int nlines, const Fft1d& pl, const double2* tw,
                                  int tid, int nthr) {
  int blk = pl.n;
  for (int f = 0; f < pl.nfac; f++) {
    const int r = pl.radix[f];
    fft_pass_any<SIGN, false>(r, buf, lstride, nlines, pl.n, blk, tw, tid, nthr);
    blk /= r;
    __syncthreads();
  }
}
// digit-reversed -> natural. Ends with a block barrier.
template <int SIGN>
ABI_DEV void fft_lines_dit(double2* buf, int lstride, int nlines, const Fft1d& pl, const double2* tw,
                                  int tid, int nthr) {
  int blk = 1;
  for (int f = pl.nfac - 1; f >= 0; f--) {
    const int r = pl.radix[f];
    blk *= r;
    fft_pass_any<SIGN, true>(r, buf, lstride, nlines, pl.n, blk, tw, tid, nthr);
    __syncthreads();
  }
}

}  // namespace abi
