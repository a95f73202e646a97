// Register-resident DFT codelets of the half-support plane stage (half_stage.cuh), lengths 2-10, 12, 14, 15, 16.
//
// Differences from the Dft<> codelets of fft_engine.cuh (which stay in use by the x passes and the older plane stage):
//  * the sine term of the odd-prime butterflies is folded into the two output additions (X_k, X_{R-k} = P -+ s Q as fused
//    multiply-adds), radix 3 costs 12 FP64 instructions instead of 14;
//  * every irrational constant is read from ONE __constant__ table: ptxas then feeds DFMA from uniform registers filled by
//    LDCU.128 (two constants per instruction) instead of two UMOV per use of a literal;
//  * 6 = 2 x 3 is done as a Good-Thomas transform (no internal twiddles).
// SIGN = +1: e^{+2 pi i jk/R} (G -> r), SIGN = -1: the inverse kernel (unnormalised).
#pragma once
#include "fft_engine.cuh"

namespace abi {

#ifdef ABI_EMU
#define ABI_CONST_TABLE static const
#else
#define ABI_CONST_TABLE static __constant__
#endif
// cos / sin of the angles the codelets need
ABI_CONST_TABLE double kHC[20] = {
    0.8660254037844386,    // 0  sin(2pi/3)
    0.30901699437494745,   // 1  cos(2pi/5)
    -0.8090169943749475,   // 2  cos(4pi/5)
    0.9510565162951535,    // 3  sin(2pi/5)
    0.5877852522924731,    // 4  sin(4pi/5)
    0.7071067811865476,    // 5  sqrt(1/2)
    0.766044443118978,     // 6  cos(2pi/9)
    0.6427876096865394,    // 7  sin(2pi/9)
    0.17364817766693036,   // 8  cos(4pi/9)
    0.984807753012208,     // 9  sin(4pi/9)
    -0.9396926207859084,   // 10 cos(8pi/9)
    0.3420201433256687,    // 11 sin(8pi/9)
    0.6234898018587335,    // 12 cos(2pi/7)
    -0.2225209339563144,   // 13 cos(4pi/7)
    -0.9009688679024191,   // 14 cos(6pi/7)
    0.7818314824680298,    // 15 sin(2pi/7)
    0.9749279121818236,    // 16 sin(4pi/7)
    0.4338837391175581,    // 17 sin(6pi/7)
    0.9238795325112867,    // 18 cos(2pi/16)
    0.3826834323650898,    // 19 sin(2pi/16)
};

// a * (c + i SIGN s) with c, s from the table
template <int SIGN> ABI_DEV double2 hmul_cs(double2 a, double c, double s) {
  const double ss = SIGN > 0 ? s : -s;
  return make_double2(fma(a.x, c, -a.y * ss), fma(a.x, ss, a.y * c));
}
// p -+ i SIGN q  (q scaled by s inside the fused multiply-adds): lo = p + i SIGN s b, hi = p - i SIGN s b
template <int SIGN> ABI_DEV void hpm_is(double2 p, double2 b, double s, double2& lo, double2& hi) {
  const double ss = SIGN > 0 ? s : -s;
  lo = make_double2(fma(-ss, b.y, p.x), fma(ss, b.x, p.y));
  hi = make_double2(fma(ss, b.y, p.x), fma(-ss, b.x, p.y));
}

template <int R, int SIGN> struct HDft;

template <int SIGN> struct HDft<2, SIGN> {
  ABI_DEV static void run(double2* x) { const double2 a = x[0], b = x[1]; x[0] = cadd(a, b); x[1] = csub(a, b); }
};
template <int SIGN> struct HDft<3, SIGN> {
  ABI_DEV static void run(double2* x) {
    const double2 a = cadd(x[1], x[2]), b = csub(x[1], x[2]);
    const double2 p = make_double2(fma(-0.5, a.x, x[0].x), fma(-0.5, a.y, x[0].y));
    x[0] = cadd(x[0], a);
    hpm_is<SIGN>(p, b, kHC[0], x[1], x[2]);
  }
};
template <int SIGN> struct HDft<4, SIGN> {
  ABI_DEV static void run(double2* x) {
    const double2 a = cadd(x[0], x[2]), b = csub(x[0], x[2]);
    const double2 c = cadd(x[1], x[3]), d = mul_si<SIGN>(csub(x[1], x[3]));
    x[0] = cadd(a, c); x[1] = cadd(b, d); x[2] = csub(a, c); x[3] = csub(b, d);
  }
};
template <int SIGN> struct HDft<5, SIGN> {
  ABI_DEV static void run(double2* x) {
    const double c1 = kHC[1], c2 = kHC[2], s1 = kHC[3], s2 = kHC[4];
    const double2 a1 = cadd(x[1], x[4]), b1 = csub(x[1], x[4]), a2 = cadd(x[2], x[3]), b2 = csub(x[2], x[3]);
    const double2 x0 = x[0];
    x[0] = cadd(cadd(x0, a1), a2);
    const double2 p1 = make_double2(fma(c2, a2.x, fma(c1, a1.x, x0.x)), fma(c2, a2.y, fma(c1, a1.y, x0.y)));
    const double2 p2 = make_double2(fma(c1, a2.x, fma(c2, a1.x, x0.x)), fma(c1, a2.y, fma(c2, a1.y, x0.y)));
    const double2 q1 = make_double2(fma(s2, b2.x, s1 * b1.x), fma(s2, b2.y, s1 * b1.y));
    const double2 q2 = make_double2(fma(-s1, b2.x, s2 * b1.x), fma(-s1, b2.y, s2 * b1.y));
    const double2 i1 = mul_si<SIGN>(q1), i2 = mul_si<SIGN>(q2);
    x[1] = cadd(p1, i1); x[4] = csub(p1, i1); x[2] = cadd(p2, i2); x[3] = csub(p2, i2);
  }
};
template <int SIGN> struct HDft<7, SIGN> {
  ABI_DEV static void run(double2* x) {
    const double c1 = kHC[12], c2 = kHC[13], c3 = kHC[14], s1 = kHC[15], s2 = kHC[16], s3 = kHC[17];
    const double2 a1 = cadd(x[1], x[6]), b1 = csub(x[1], x[6]), a2 = cadd(x[2], x[5]), b2 = csub(x[2], x[5]);
    const double2 a3 = cadd(x[3], x[4]), b3 = csub(x[3], x[4]);
    const double2 x0 = x[0];
    x[0] = cadd(cadd(x0, a1), cadd(a2, a3));
    // k = 1: (c1, c2, c3 | s1, s2, s3); k = 2: (c2, c3, c1 | s2, -s3, -s1); k = 3: (c3, c1, c2 | s3, -s1, s2)
    const double2 p1 = make_double2(fma(c3, a3.x, fma(c2, a2.x, fma(c1, a1.x, x0.x))), fma(c3, a3.y, fma(c2, a2.y, fma(c1, a1.y, x0.y))));
    const double2 p2 = make_double2(fma(c1, a3.x, fma(c3, a2.x, fma(c2, a1.x, x0.x))), fma(c1, a3.y, fma(c3, a2.y, fma(c2, a1.y, x0.y))));
    const double2 p3 = make_double2(fma(c2, a3.x, fma(c1, a2.x, fma(c3, a1.x, x0.x))), fma(c2, a3.y, fma(c1, a2.y, fma(c3, a1.y, x0.y))));
    const double2 q1 = make_double2(fma(s3, b3.x, fma(s2, b2.x, s1 * b1.x)), fma(s3, b3.y, fma(s2, b2.y, s1 * b1.y)));
    const double2 q2 = make_double2(fma(-s1, b3.x, fma(-s3, b2.x, s2 * b1.x)), fma(-s1, b3.y, fma(-s3, b2.y, s2 * b1.y)));
    const double2 q3 = make_double2(fma(s2, b3.x, fma(-s1, b2.x, s3 * b1.x)), fma(s2, b3.y, fma(-s1, b2.y, s3 * b1.y)));
    const double2 i1 = mul_si<SIGN>(q1), i2 = mul_si<SIGN>(q2), i3 = mul_si<SIGN>(q3);
    x[1] = cadd(p1, i1); x[6] = csub(p1, i1); x[2] = cadd(p2, i2); x[5] = csub(p2, i2); x[3] = cadd(p3, i3); x[4] = csub(p3, i3);
  }
};
// 8 = 2 x 4 Cooley-Tukey: y_k0[j0] = (x[j0] +- x[j0 + 4]) w8^(j0 k0); X[2 k1 + k0] = DFT4(y_k0)[k1]
template <int SIGN> struct HDft<8, SIGN> {
  ABI_DEV static void run(double2* x) {
    const double h = kHC[5];
    double2 e[4], o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { e[j] = cadd(x[j], x[j + 4]); o[j] = csub(x[j], x[j + 4]); }
    // o[j] *= w8^j: (1 + i s)/sqrt2, i s, (-1 + i s)/sqrt2   (s = SIGN)
    { const double2 t = o[1]; o[1] = SIGN > 0 ? make_double2((t.x - t.y) * h, (t.x + t.y) * h) : make_double2((t.x + t.y) * h, (t.y - t.x) * h); }
    o[2] = mul_si<SIGN>(o[2]);
    { const double2 t = o[3]; o[3] = SIGN > 0 ? make_double2(-(t.x + t.y) * h, (t.x - t.y) * h) : make_double2((t.y - t.x) * h, -(t.x + t.y) * h); }
    HDft<4, SIGN>::run(e); HDft<4, SIGN>::run(o);
#pragma unroll
    for (int k = 0; k < 4; k++) { x[2 * k] = e[k]; x[2 * k + 1] = o[k]; }
  }
};
// 9 = 3 x 3 Cooley-Tukey: y_k0[j0] = DFT3_j1(x[3 j1 + j0])[k0] w9^(j0 k0); X[3 k1 + k0] = DFT3_j0(y_k0)[k1]
template <int SIGN> struct HDft<9, SIGN> {
  ABI_DEV static void run(double2* x) {
    double2 y[3][3];
#pragma unroll
    for (int j0 = 0; j0 < 3; j0++) {
      double2 col[3] = {x[j0], x[3 + j0], x[6 + j0]};
      HDft<3, SIGN>::run(col);
      y[0][j0] = col[0];
      if (j0 == 0) { y[1][0] = col[1]; y[2][0] = col[2]; }
      else if (j0 == 1) { y[1][1] = hmul_cs<SIGN>(col[1], kHC[6], kHC[7]); y[2][1] = hmul_cs<SIGN>(col[2], kHC[8], kHC[9]); }
      else { y[1][2] = hmul_cs<SIGN>(col[1], kHC[8], kHC[9]); y[2][2] = hmul_cs<SIGN>(col[2], kHC[10], kHC[11]); }
    }
#pragma unroll
    for (int k0 = 0; k0 < 3; k0++) {
      HDft<3, SIGN>::run(y[k0]);
#pragma unroll
      for (int k1 = 0; k1 < 3; k1++) x[3 * k1 + k0] = y[k0][k1];
    }
  }
};
// 16 = 4 x 4 Cooley-Tukey: y_k0[j0] = DFT4_j1(x[4 j1 + j0])[k0] w16^(j0 k0); X[4 k1 + k0] = DFT4_j0(y_k0)[k1]
template <int SIGN> struct HDft<16, SIGN> {
  template <int E> ABI_DEV static double2 tw(double2 a) {      // a * w16^E, E = j0 k0 in {0,1,2,3,4,6,9}
    const double c = kHC[18], s = kHC[19], h = kHC[5];
    if (E == 0) return a;
    if (E == 1) return hmul_cs<SIGN>(a, c, s);
    if (E == 2) return hmul_cs<SIGN>(a, h, h);
    if (E == 3) return hmul_cs<SIGN>(a, s, c);
    if (E == 4) return mul_si<SIGN>(a);
    if (E == 6) return hmul_cs<SIGN>(a, -h, h);
    return hmul_cs<SIGN>(a, -c, -s);                              // E == 9
  }
  ABI_DEV static void run(double2* x) {
    double2 y[4][4];
#pragma unroll
    for (int j0 = 0; j0 < 4; j0++) {
      double2 col[4] = {x[j0], x[4 + j0], x[8 + j0], x[12 + j0]};
      HDft<4, SIGN>::run(col);
      y[0][j0] = col[0];
      if (j0 == 0) { y[1][0] = col[1]; y[2][0] = col[2]; y[3][0] = col[3]; }
      else if (j0 == 1) { y[1][1] = tw<1>(col[1]); y[2][1] = tw<2>(col[2]); y[3][1] = tw<3>(col[3]); }
      else if (j0 == 2) { y[1][2] = tw<2>(col[1]); y[2][2] = tw<4>(col[2]); y[3][2] = tw<6>(col[3]); }
      else { y[1][3] = tw<3>(col[1]); y[2][3] = tw<6>(col[2]); y[3][3] = tw<9>(col[3]); }
    }
#pragma unroll
    for (int k0 = 0; k0 < 4; k0++) {
      HDft<4, SIGN>::run(y[k0]);
#pragma unroll
      for (int k1 = 0; k1 < 4; k1++) x[4 * k1 + k0] = y[k0][k1];
    }
  }
};

// Good-Thomas prime-factor DFT for coprime A, B on the HDft codelets: no internal twiddles, compile-time index maps
template <int A, int B, int SIGN> struct HDftPFA {
  static constexpr int N = A * B;
  ABI_HD static constexpr int inv_mod(int a, int m) { int r = 1; for (int i = 1; i < m; i++) if ((a * i) % m == 1) r = i; return r; }
  ABI_DEV static void run(double2* x) {
    constexpr int bi = inv_mod(B % A, A), ai = inv_mod(A % B, B);
    double2 y[B][A];
#pragma unroll
    for (int n2 = 0; n2 < B; n2++) {
      double2 col[A];
#pragma unroll
      for (int n1 = 0; n1 < A; n1++) col[n1] = x[(B * n1 + A * n2) % N];
      HDft<A, SIGN>::run(col);
#pragma unroll
      for (int k1 = 0; k1 < A; k1++) y[n2][k1] = col[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < A; k1++) {
      double2 row[B];
#pragma unroll
      for (int n2 = 0; n2 < B; n2++) row[n2] = y[n2][k1];
      HDft<B, SIGN>::run(row);
#pragma unroll
      for (int k2 = 0; k2 < B; k2++) x[(B * bi * k1 + A * ai * k2) % N] = row[k2];
    }
  }
};
template <int SIGN> struct HDft<6, SIGN> : HDftPFA<2, 3, SIGN> {};
template <int SIGN> struct HDft<10, SIGN> : HDftPFA<2, 5, SIGN> {};
template <int SIGN> struct HDft<12, SIGN> : HDftPFA<3, 4, SIGN> {};
template <int SIGN> struct HDft<14, SIGN> : HDftPFA<2, 7, SIGN> {};
template <int SIGN> struct HDft<15, SIGN> : HDftPFA<3, 5, SIGN> {};

}  // namespace abi
