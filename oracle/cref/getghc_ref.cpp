// C++/OpenMP restatement of the reference's CPU getghc for the bench arm (TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing
// under abinit_b200/ may link or load this; see oracle/__init__.py).  The Fortran reference cannot be built in this image
// (no Fortran compiler, no FFTW, no MPI), so this file restates the algorithm of its CPU build in C++ with OpenMP:
//
//   local part    zero-padded 3-D FFT passes   src/52_fft_mpi_noabirule/fftw3_fftpad.finc:14-103 (G -> r: x transforms on the
//                 occupied (i2,i3) lines, y transforms on the occupied z planes, z transforms on every column) and :105-196
//                 (r -> G, reverse order, pruned by the output sphere); sphere <-> box  src/52_fft_mpi_noabirule/m_fftcore.F90:1532-1866;
//                 V(r) psi(r)  src/44_abitools/m_cgtools.F90:2410-2491; two bands per complex transform at the Gamma point
//                 (cwavef_double_rfft_trick_pack / _unpack, src/66_wfs/m_getghc.F90:1999-2171)
//   non-local     gemm_nonlop with P_r / P_i held separately and real DGEMMs  src/66_nonlocal/m_gemm_nonlop.F90:191-1242,
//                 m_opernla_gemm.F90:569-689, m_opernlc_ylm_allwf.F90:308-447 and :1253-1295, m_opernlb_gemm.F90:700-833
//   assembly      ghc = vloc psi + kinpw psi + vnl psi with the huge * 1e-11 filter  src/66_wfs/m_getghc.F90:1266-1280
//   threading     OpenMP inside every pass (the reference threads its FFT passes and loops over bands the same way,
//                 multithreaded_getghc m_getghc.F90:2443-2542) + the threads of the BLAS it is given
//
// The 1-D transforms are a Stockham mixed-radix (2, 3, 4, 5, generic odd) engine working on batches of lines held as
// structure-of-arrays so that the butterfly loops vectorise over the batch (the role FFTW3's SIMD codelets play in the
// reference); DGEMM is the host BLAS handed in as a function pointer (OpenBLAS, the BLAS the reference links to).
// Pinned against oracle/getghc.py (NumPy) to 1e-13 by tests/test_cref.py; the NumPy oracle itself is pinned on the reference's
// stored SCF results (oracle/__init__.py).
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using cplx = std::complex<double>;
constexpr int HC = 16;   // lines per batch of the SoA work buffers
typedef double V __attribute__((vector_size(32), aligned(8)));   // 4 lines per SIMD register (AVX2); GCC vector extension
constexpr int VL = 4;

struct Fft1D {
  int n = 0;
  std::vector<int> radix;
  std::vector<double> wr, wi;                       // exp(+2 pi i q / n)
  std::vector<std::vector<double>> pr, pi;          // per pass: twiddles exp(+2 pi i q k / (p R)) at [k * R + q], k < p
  explicit Fft1D(int n_) : n(n_) {
    int m = n;
    for (int r : {4, 2, 3, 5}) while (m % r == 0) { radix.push_back(r); m /= r; }
    for (int r = 7; m > 1; r += 2) while (m % r == 0) { radix.push_back(r); m /= r; }
    wr.resize(n); wi.resize(n);
    for (int q = 0; q < n; q++) { wr[q] = std::cos(2.0 * M_PI * q / n); wi[q] = std::sin(2.0 * M_PI * q / n); }
    int p = 1;
    for (int r : radix) {
      std::vector<double> a((size_t)p * r), b((size_t)p * r);
      const int step = n / (p * r);
      for (int k = 0; k < p; k++) for (int q = 0; q < r; q++) { const int idx = (int)(((long long)q * k * step) % n); a[k * r + q] = wr[idx]; b[k * r + q] = wi[idx]; }
      pr.push_back(a); pi.push_back(b);
      p *= r;
    }
  }
};

// one Stockham pass of radix R on HC interleaved lines: x[(i + q t)][h] -> y[(j + q p)][h]
template <int R>
inline void pass_fixed(const Fft1D& f, int ipass, int p, int sign, const double* __restrict xr, const double* __restrict xi,
                       double* __restrict yr, double* __restrict yi) {
  const int n = f.n, t = n / R;
  const double* tr = f.pr[ipass].data(); const double* ti = f.pi[ipass].data();
  const double sg = (double)sign;
  for (int i = 0; i < t; i++) {
    const int k = i % p, j = (i - k) * R + k;
    double twr[R], twi[R];
    for (int q = 0; q < R; q++) { twr[q] = tr[k * R + q]; twi[q] = sg * ti[k * R + q]; }
    const double* ar[R]; const double* ai[R]; double* br[R]; double* bi[R];
    for (int q = 0; q < R; q++) {
      ar[q] = xr + (size_t)(i + q * t) * HC; ai[q] = xi + (size_t)(i + q * t) * HC;
      br[q] = yr + (size_t)(j + q * p) * HC; bi[q] = yi + (size_t)(j + q * p) * HC;
    }
    const bool unit = (k == 0);                      // every twiddle is 1 (always true in the first pass)
    for (int h = 0; h < HC; h += VL) {
      V ur[R], ui[R];
      for (int q = 0; q < R; q++) {
        const V a = *reinterpret_cast<const V*>(ar[q] + h), b = *reinterpret_cast<const V*>(ai[q] + h);
        if (unit || q == 0) { ur[q] = a; ui[q] = b; }
        else { ur[q] = a * twr[q] - b * twi[q]; ui[q] = a * twi[q] + b * twr[q]; }
      }
      V yr_[R], yi_[R];
      if (R == 2) {
        yr_[0] = ur[0] + ur[1]; yi_[0] = ui[0] + ui[1];
        yr_[1] = ur[0] - ur[1]; yi_[1] = ui[0] - ui[1];
      } else if (R == 4) {
        const V s0r = ur[0] + ur[2], s0i = ui[0] + ui[2], d0r = ur[0] - ur[2], d0i = ui[0] - ui[2];
        const V s1r = ur[1] + ur[3], s1i = ui[1] + ui[3], d1r = ur[1] - ur[3], d1i = ui[1] - ui[3];
        const V jr = -sg * d1i, ji = sg * d1r;       // sign * i * d1
        yr_[0] = s0r + s1r; yi_[0] = s0i + s1i;
        yr_[2] = s0r - s1r; yi_[2] = s0i - s1i;
        yr_[1] = d0r + jr; yi_[1] = d0i + ji;
        yr_[3] = d0r - jr; yi_[3] = d0i - ji;
      } else if (R == 3) {
        const double c = -0.5, s3 = sg * 0.86602540378443864676;
        const V sr = ur[1] + ur[2], si = ui[1] + ui[2], dr = ur[1] - ur[2], di = ui[1] - ui[2];
        const V mr = ur[0] + c * sr, mi = ui[0] + c * si;
        yr_[0] = ur[0] + sr; yi_[0] = ui[0] + si;
        yr_[1] = mr - s3 * di; yi_[1] = mi + s3 * dr;
        yr_[2] = mr + s3 * di; yi_[2] = mi - s3 * dr;
      } else if (R == 5) {
        const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
        const double s1 = sg * 0.95105651629515357212, s2 = sg * 0.58778525229247312917;
        const V a1r = ur[1] + ur[4], a1i = ui[1] + ui[4], b1r = ur[1] - ur[4], b1i = ui[1] - ui[4];
        const V a2r = ur[2] + ur[3], a2i = ui[2] + ui[3], b2r = ur[2] - ur[3], b2i = ui[2] - ui[3];
        yr_[0] = ur[0] + a1r + a2r; yi_[0] = ui[0] + a1i + a2i;
        const V m1r = ur[0] + c1 * a1r + c2 * a2r, m1i = ui[0] + c1 * a1i + c2 * a2i;
        const V m2r = ur[0] + c2 * a1r + c1 * a2r, m2i = ui[0] + c2 * a1i + c1 * a2i;
        const V n1r = s1 * b1r + s2 * b2r, n1i = s1 * b1i + s2 * b2i;
        const V n2r = s2 * b1r - s1 * b2r, n2i = s2 * b1i - s1 * b2i;
        yr_[1] = m1r - n1i; yi_[1] = m1i + n1r;
        yr_[4] = m1r + n1i; yi_[4] = m1i - n1r;
        yr_[2] = m2r - n2i; yi_[2] = m2i + n2r;
        yr_[3] = m2r + n2i; yi_[3] = m2i - n2r;
      }
      for (int q = 0; q < R; q++) { *reinterpret_cast<V*>(br[q] + h) = yr_[q]; *reinterpret_cast<V*>(bi[q] + h) = yi_[q]; }
    }
  }
}

// generic odd radix (7, 11, ...): O(R^2) per butterfly
inline void pass_generic(const Fft1D& f, int R, int p, int sign, const double* __restrict xr, const double* __restrict xi,
                         double* __restrict yr, double* __restrict yi) {
  const int n = f.n, t = n / R, step = n / (p * R), rs = n / R;
  std::vector<double> ur(R * HC), ui(R * HC);
  for (int i = 0; i < t; i++) {
    const int k = i % p, j = (i - k) * R + k;
    for (int q = 0; q < R; q++) {
      const int idx = (int)(((long long)q * k * step) % n);
      const double cr = f.wr[idx], ci = sign * f.wi[idx];
      for (int h = 0; h < HC; h++) {
        const double a = xr[(size_t)(i + q * t) * HC + h], b = xi[(size_t)(i + q * t) * HC + h];
        ur[q * HC + h] = a * cr - b * ci; ui[q * HC + h] = a * ci + b * cr;
      }
    }
    for (int o = 0; o < R; o++) {
      double* br = yr + (size_t)(j + o * p) * HC; double* bi = yi + (size_t)(j + o * p) * HC;
      for (int h = 0; h < HC; h++) { br[h] = 0.0; bi[h] = 0.0; }
      for (int q = 0; q < R; q++) {
        const int idx = (int)(((long long)o * q * rs) % n);
        const double cr = f.wr[idx], ci = sign * f.wi[idx];
        for (int h = 0; h < HC; h++) { br[h] += ur[q * HC + h] * cr - ui[q * HC + h] * ci; bi[h] += ur[q * HC + h] * ci + ui[q * HC + h] * cr; }
      }
    }
  }
}

// all passes of one batch; the result ends in (xr, xi) (the pointers are swapped along the way)
inline void run_passes(const Fft1D& f, int sign, double*& xr, double*& xi, double*& yr, double*& yi) {
  int p = 1, ip = 0;
  for (int r : f.radix) {
    switch (r) {
      case 2: pass_fixed<2>(f, ip, p, sign, xr, xi, yr, yi); break;
      case 3: pass_fixed<3>(f, ip, p, sign, xr, xi, yr, yi); break;
      case 4: pass_fixed<4>(f, ip, p, sign, xr, xi, yr, yi); break;
      case 5: pass_fixed<5>(f, ip, p, sign, xr, xi, yr, yi); break;
      default: pass_generic(f, r, p, sign, xr, xi, yr, yi);
    }
    std::swap(xr, yr); std::swap(xi, yi);
    p *= r; ip++;
  }
}

struct Work {
  std::vector<double> buf; double *ar, *ai, *br, *bi;
  explicit Work(int n) : buf(4 * (size_t)n * HC) { ar = buf.data(); ai = ar + (size_t)n * HC; br = ai + (size_t)n * HC; bi = br + (size_t)n * HC; }
};

// copy-in / copy-out of one batch of lines data[off[l0 + h] + e * stride]; ein / eout: the elements to read / write (nullptr: all;
// the others are zero on input / not wanted on output)
inline void load_batch(const cplx* data, const int64_t* off, int nh, int64_t stride, int n, const int* el, int nel, double* ar, double* ai) {
  if (el != nullptr || nh < HC) std::memset(ar, 0, sizeof(double) * 2 * (size_t)n * HC);   // ar and ai are adjacent
  if (stride == 1) {
    for (int h = 0; h < nh; h++) {
      const cplx* src = data + off[h];
      if (el) for (int q = 0; q < nel; q++) { const int e = el[q]; ar[(size_t)e * HC + h] = src[e].real(); ai[(size_t)e * HC + h] = src[e].imag(); }
      else for (int e = 0; e < n; e++) { ar[(size_t)e * HC + h] = src[e].real(); ai[(size_t)e * HC + h] = src[e].imag(); }
    }
  } else {
    const int ne = el ? nel : n;
    for (int q = 0; q < ne; q++) {
      const int e = el ? el[q] : q;
      for (int h = 0; h < nh; h++) { const cplx v = data[off[h] + e * stride]; ar[(size_t)e * HC + h] = v.real(); ai[(size_t)e * HC + h] = v.imag(); }
    }
  }
}
inline void store_batch(cplx* data, const int64_t* off, int nh, int64_t stride, int n, const int* el, int nel, const double* xr, const double* xi) {
  if (stride == 1) {
    for (int h = 0; h < nh; h++) {
      cplx* dst = data + off[h];
      if (el) for (int q = 0; q < nel; q++) { const int e = el[q]; dst[e] = cplx(xr[(size_t)e * HC + h], xi[(size_t)e * HC + h]); }
      else for (int e = 0; e < n; e++) dst[e] = cplx(xr[(size_t)e * HC + h], xi[(size_t)e * HC + h]);
    }
  } else {
    const int ne = el ? nel : n;
    for (int q = 0; q < ne; q++) {
      const int e = el ? el[q] : q;
      for (int h = 0; h < nh; h++) data[off[h] + e * stride] = cplx(xr[(size_t)e * HC + h], xi[(size_t)e * HC + h]);
    }
  }
}

// in-place transforms of the lines data[off[l] + e * stride], e < n, l < nl (sign = +1: e^{+i}, -1: e^{-i}; unscaled)
void fft_lines(const Fft1D& f, cplx* data, const int64_t* off, int64_t nl, int64_t stride, int sign,
               const std::vector<int>* ein = nullptr, const std::vector<int>* eout = nullptr) {
  const int n = f.n;
#pragma omp parallel
  {
    Work w(n);
#pragma omp for schedule(static)
    for (int64_t l0 = 0; l0 < nl; l0 += HC) {
      const int nh = (int)std::min<int64_t>(HC, nl - l0);
      load_batch(data, off + l0, nh, stride, n, ein ? ein->data() : nullptr, ein ? (int)ein->size() : 0, w.ar, w.ai);
      double *xr = w.ar, *xi = w.ai, *yr = w.br, *yi = w.bi;
      run_passes(f, sign, xr, xi, yr, yi);
      store_batch(data, off + l0, nh, stride, n, eout ? eout->data() : nullptr, eout ? (int)eout->size() : 0, xr, xi);
    }
  }
}

// z columns: transform e^{+i} of the occupied input planes, V(r) psi(r) on the batch while it sits in the work buffers (the way
// sg_fftrisc applies the potential between its z transforms, src/52_fft_mpi_noabirule/m_sgfft.F90), transform back, store the
// output planes only.  vlocal is indexed like the box.
void fft_z_times_v(const Fft1D& f, cplx* data, const int64_t* off, int64_t nl, int64_t stride, const double* vlocal,
                   const std::vector<int>& ein, const std::vector<int>& eout) {
  const int n = f.n;
#pragma omp parallel
  {
    Work w(n);
#pragma omp for schedule(static)
    for (int64_t l0 = 0; l0 < nl; l0 += HC) {
      const int nh = (int)std::min<int64_t>(HC, nl - l0);
      load_batch(data, off + l0, nh, stride, n, ein.data(), (int)ein.size(), w.ar, w.ai);
      double *xr = w.ar, *xi = w.ai, *yr = w.br, *yi = w.bi;
      run_passes(f, +1, xr, xi, yr, yi);
      for (int e = 0; e < n; e++)
        for (int h = 0; h < nh; h++) { const double v = vlocal[off[l0 + h] + e * stride]; xr[(size_t)e * HC + h] *= v; xi[(size_t)e * HC + h] *= v; }
      run_passes(f, -1, xr, xi, yr, yi);
      store_batch(data, off + l0, nh, stride, n, eout.data(), (int)eout.size(), xr, xi);
    }
  }
}

struct Sphere {
  int n1, n2, n3, npw, istwf_k, lo;
  std::vector<int> i1, i2, i3, j1, j2, j3;          // wrapped box indices of G and of -G (time-reversal images, entries lo..)
  std::vector<int64_t> in_lines, out_lines;         // box offsets (i3 * n2 + i2) * n1 of the occupied x lines
  std::vector<int64_t> in_y, out_y;                 // offsets of the y lines of the occupied z planes (i3 * n2 * n1 + i1)
  std::vector<int64_t> z_all;                       // offsets of every z column (i2 * n1 + i1)
  std::vector<int> in_planes, out_planes;           // occupied z planes
};

int wrap(int g, int n) { int r = g % n; return r < 0 ? r + n : r; }

void build_sphere(Sphere& s, const int* kg, bool packed) {
  const int n1 = s.n1, n2 = s.n2, n3 = s.n3, npw = s.npw;
  s.i1.resize(npw); s.i2.resize(npw); s.i3.resize(npw);
  for (int p = 0; p < npw; p++) { s.i1[p] = wrap(kg[3 * p], n1); s.i2[p] = wrap(kg[3 * p + 1], n2); s.i3[p] = wrap(kg[3 * p + 2], n3); }
  s.lo = 0;
  std::vector<char> lin_in((size_t)n2 * n3, 0), lin_out((size_t)n2 * n3, 0);
  for (int p = 0; p < npw; p++) { lin_in[(size_t)s.i3[p] * n2 + s.i2[p]] = 1; lin_out[(size_t)s.i3[p] * n2 + s.i2[p]] = 1; }
  if (s.istwf_k == 2) {
    s.lo = 1;                                        // G = 0 is the first plane wave and has no image (me_g0 = 1)
    s.j1.assign(npw, 0); s.j2.assign(npw, 0); s.j3.assign(npw, 0);
    for (int p = s.lo; p < npw; p++) {
      s.j1[p] = wrap(-kg[3 * p], n1); s.j2[p] = wrap(-kg[3 * p + 1], n2); s.j3[p] = wrap(-kg[3 * p + 2], n3);
      lin_in[(size_t)s.j3[p] * n2 + s.j2[p]] = 1;
      if (packed) lin_out[(size_t)s.j3[p] * n2 + s.j2[p]] = 1;   // the packed transform needs F on the completed sphere
    }
  }
  auto lines_of = [&](const std::vector<char>& m, std::vector<int64_t>& lines, std::vector<int64_t>& ys, std::vector<int>& planes) {
    std::vector<char> zpl(n3, 0);
    for (int a = 0; a < n3; a++) for (int b = 0; b < n2; b++) if (m[(size_t)a * n2 + b]) { lines.push_back(((int64_t)a * n2 + b) * n1); zpl[a] = 1; }
    for (int a = 0; a < n3; a++) if (zpl[a]) { planes.push_back(a); for (int c = 0; c < n1; c++) ys.push_back((int64_t)a * n2 * n1 + c); }
  };
  lines_of(lin_in, s.in_lines, s.in_y, s.in_planes);
  lines_of(lin_out, s.out_lines, s.out_y, s.out_planes);
  s.z_all.resize((size_t)n1 * n2);
  for (int64_t q = 0; q < (int64_t)n1 * n2; q++) s.z_all[q] = q;
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// fourwf option 2 on the whole block: out(ndat, npw) = gather(FFT[V FFT^-1[scatter(in)]]) / N
void fourwf_opt2(const Sphere& s, const double* vlocal, const cplx* in, cplx* out, int ndat) {
  const int n1 = s.n1, n2 = s.n2, n3 = s.n3, npw = s.npw;
  const int64_t N = (int64_t)n1 * n2 * n3;
  const double xnorm = 1.0 / (double)N;
  const bool pack = s.istwf_k == 2;
  const Fft1D f1(n1), f2(n2), f3(n3);
  // uninitialised on purpose: every transform zeroes the planes it reads
  struct Free { void operator()(void* q) const { std::free(q); } };
  std::unique_ptr<void, Free> boxv(std::malloc(sizeof(cplx) * (size_t)N));
  cplx* box = static_cast<cplx*>(boxv.get());
  const int ntrans = pack ? (ndat + 1) / 2 : ndat;
  for (int t = 0; t < ntrans; t++) {
    // only the occupied z planes are ever read before they are written: zero those (the z pass skips the others)
    const int64_t plane = (int64_t)n1 * n2;
#pragma omp parallel for schedule(static)
    for (int64_t a = 0; a < (int64_t)s.in_planes.size(); a++) std::memset((void*)(box + s.in_planes[a] * plane), 0, sizeof(cplx) * plane);
    if (pack) {
      const cplx* c = in + (size_t)(2 * t) * npw;
      const bool has_d = 2 * t + 1 < ndat;
      const cplx* d = c + npw;
      // E(G) = C(G) + i D(G); E(-G) = conj(C(G)) + i conj(D(G)); Im c(G=0) = 0 (m_fftcore.F90:1632-1638)
#pragma omp parallel for schedule(static)
      for (int p = 0; p < npw; p++) {
        cplx cc = c[p], dd = has_d ? d[p] : cplx(0.0, 0.0);
        if (p < s.lo) { cc = cplx(cc.real(), 0.0); dd = cplx(dd.real(), 0.0); }
        box[((int64_t)s.i3[p] * n2 + s.i2[p]) * n1 + s.i1[p]] = cc + cplx(0.0, 1.0) * dd;
        if (p >= s.lo) box[((int64_t)s.j3[p] * n2 + s.j2[p]) * n1 + s.j1[p]] = std::conj(cc) + cplx(0.0, 1.0) * std::conj(dd);
      }
    } else {
      const cplx* c = in + (size_t)t * npw;
#pragma omp parallel for schedule(static)
      for (int p = 0; p < npw; p++) box[((int64_t)s.i3[p] * n2 + s.i2[p]) * n1 + s.i1[p]] = c[p];
    }
    fft_lines(f1, box, s.in_lines.data(), (int64_t)s.in_lines.size(), 1, +1);
    fft_lines(f2, box, s.in_y.data(), (int64_t)s.in_y.size(), n1, +1);
    fft_z_times_v(f3, box, s.z_all.data(), (int64_t)s.z_all.size(), (int64_t)n1 * n2, vlocal, s.in_planes, s.out_planes);
    fft_lines(f2, box, s.out_y.data(), (int64_t)s.out_y.size(), n1, -1);
    fft_lines(f1, box, s.out_lines.data(), (int64_t)s.out_lines.size(), 1, -1);
    if (pack) {
      cplx* hc = out + (size_t)(2 * t) * npw;
      const bool has_d = 2 * t + 1 < ndat;
      cplx* hd = hc + npw;
      // H C(G) = [F(G) + conj F(-G)] / 2,  H D(G) = [F(G) - conj F(-G)] / (2i); G = 0: Re F and Im F
#pragma omp parallel for schedule(static)
      for (int p = 0; p < npw; p++) {
        const cplx f = box[((int64_t)s.i3[p] * n2 + s.i2[p]) * n1 + s.i1[p]] * xnorm;
        if (p < s.lo) { hc[p] = cplx(f.real(), 0.0); if (has_d) hd[p] = cplx(f.imag(), 0.0); continue; }
        const cplx g = std::conj(box[((int64_t)s.j3[p] * n2 + s.j2[p]) * n1 + s.j1[p]] * xnorm);
        hc[p] = 0.5 * (f + g);
        if (has_d) hd[p] = cplx(0.0, -0.5) * (f - g);
      }
    } else {
      cplx* o = out + (size_t)t * npw;
#pragma omp parallel for schedule(static)
      for (int p = 0; p < npw; p++) o[p] = box[((int64_t)s.i3[p] * n2 + s.i2[p]) * n1 + s.i1[p]] * xnorm;
    }
  }
}

typedef void (*dgemm_t)(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*,
                        const double*, const int*, const double*, double*, const int*);

}  // namespace

extern "C" {

// 1-D transform check entry: data(n, howmany) interleaved complex, contiguous lines
void cref_fft1d(double* data, int n, int howmany, int sign) {
  const Fft1D f(n);
  std::vector<int64_t> off(howmany);
  for (int h = 0; h < howmany; h++) off[h] = (int64_t)h * n;
  fft_lines(f, reinterpret_cast<cplx*>(data), off.data(), howmany, 1, sign);
}

// fourwf option 2, cplex 1 (real V), istwf_k 1 or 2 (Gamma, G = 0 first, two bands per transform)
int cref_fourwf_opt2(int ndat, int npw, int istwf_k, int n1, int n2, int n3, const int* kg, const double* cwavef, const double* vlocal,
                     double* out) {
  if (istwf_k != 1 && istwf_k != 2) return 1;
  Sphere s; s.n1 = n1; s.n2 = n2; s.n3 = n3; s.npw = npw; s.istwf_k = istwf_k;
  build_sphere(s, kg, istwf_k == 2);
  fourwf_opt2(s, vlocal, reinterpret_cast<const cplx*>(cwavef), reinterpret_cast<cplx*>(out), ndat);
  return 0;
}

// getghc, type_calc 0, cpopt -1.  paw = 0: gxfac = ekb_proj * gx; paw = 1: per-block packed symmetric D_ij (dij[nblk][dimenl1]) and,
// with sij_opt = 1, S_ij (sij[nblk][dimenl1]) -> gsc.  Pr, Pi: [nprojs][npw].  timings[4]: fourwf, opernla, opernlc, opernlb + assembly.
int cref_getghc(int ndat, int npw, int istwf_k, int n1, int n2, int n3, const int* kg, const double* cwavef, const double* vlocal,
                const double* kinpw, int nprojs, const double* Pr, const double* Pi, int paw, const double* ekb_proj, int nblk,
                const int* blk_nlmn, const double* dij, const double* sij, int dimenl1, int sij_opt, double* ghc, double* gsc,
                void* dgemm_fn, double* timings) {
  if (istwf_k != 1 && istwf_k != 2) return 1;
  dgemm_t dgemm = reinterpret_cast<dgemm_t>(dgemm_fn);
  const double kin_filter = 1.7976931348623157e308 * 1.0e-11;   // huge(0d0) * 1d-11 (m_getghc.F90:1272)
  double t0 = now();
  Sphere s; s.n1 = n1; s.n2 = n2; s.n3 = n3; s.npw = npw; s.istwf_k = istwf_k;
  build_sphere(s, kg, istwf_k == 2);
  const cplx* cw = reinterpret_cast<const cplx*>(cwavef);
  cplx* gh = reinterpret_cast<cplx*>(ghc);
  fourwf_opt2(s, vlocal, cw, gh, ndat);
  double t1 = now();
  // ---- opernla: gx = P^H psi on split real / imaginary parts ----
  const size_t nw = (size_t)npw * ndat;
  std::vector<double> vr(nw), vi(nw);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)nw; q++) { vr[q] = cw[q].real(); vi[q] = cw[q].imag(); }
  const bool real_proj = istwf_k == 2;
  const int ncx = real_proj ? 1 : 2;
  std::vector<double> gx((size_t)ncx * nprojs * ndat), gxfac(gx.size()), gxs(sij_opt == 1 ? gx.size() : 0);
  double* gxr = gx.data(); double* gxi = gxr + (real_proj ? 0 : (size_t)nprojs * ndat);
  const double one = 1.0, zero = 0.0, two = 2.0, mone = -1.0;
  if (real_proj) {
    // Re halved and Im dropped at G = 0, then 2 (P_r^T psi_r + P_i^T psi_i)   (m_opernla_gemm.F90:569-612, 641-689)
    for (int b = 0; b < ndat; b++) { vr[(size_t)b * npw] *= 0.5; vi[(size_t)b * npw] = 0.0; }
    dgemm("T", "N", &nprojs, &ndat, &npw, &two, Pr, &npw, vr.data(), &npw, &zero, gxr, &nprojs);
    dgemm("T", "N", &nprojs, &ndat, &npw, &two, Pi, &npw, vi.data(), &npw, &one, gxr, &nprojs);
    for (int b = 0; b < ndat; b++) { vr[(size_t)b * npw] = cw[(size_t)b * npw].real(); vi[(size_t)b * npw] = cw[(size_t)b * npw].imag(); }
  } else {
    dgemm("T", "N", &nprojs, &ndat, &npw, &one, Pr, &npw, vr.data(), &npw, &zero, gxr, &nprojs);
    dgemm("T", "N", &nprojs, &ndat, &npw, &one, Pi, &npw, vi.data(), &npw, &one, gxr, &nprojs);
    dgemm("T", "N", &nprojs, &ndat, &npw, &one, Pr, &npw, vi.data(), &npw, &zero, gxi, &nprojs);
    dgemm("T", "N", &nprojs, &ndat, &npw, &mone, Pi, &npw, vr.data(), &npw, &one, gxi, &nprojs);
  }
  double t2 = now();
  // ---- opernlc ----
  if (!paw) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)gx.size(); q++) gxfac[q] = ekb_proj[q % nprojs] * gx[q];
  } else {
    std::vector<int> bstart(nblk + 1, 0);
    for (int a = 0; a < nblk; a++) bstart[a + 1] = bstart[a] + blk_nlmn[a];
#pragma omp parallel for schedule(dynamic, 4) collapse(2)
    for (int col = 0; col < ncx * ndat; col++)
      for (int a = 0; a < nblk; a++) {
        const int nl = blk_nlmn[a];
        const double* g = gx.data() + (size_t)col * nprojs + bstart[a];
        double* o = gxfac.data() + (size_t)col * nprojs + bstart[a];
        double* os = sij_opt == 1 ? gxs.data() + (size_t)col * nprojs + bstart[a] : nullptr;
        const double* D = dij + (size_t)a * dimenl1;
        const double* S = sij_opt == 1 ? sij + (size_t)a * dimenl1 : nullptr;
        for (int i = 0; i < nl; i++) {
          double acc = 0.0, accs = 0.0;
          for (int j = 0; j < nl; j++) {
            const int ij = i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j;
            acc += D[ij] * g[j];
            if (S) accs += S[ij] * g[j];
          }
          o[i] = acc;
          if (os) os[i] = accs;
        }
      }
  }
  double t3 = now();
  // ---- opernlb + assembly ----
  std::vector<double> wr(nw), wi(nw);
  auto back = [&](const std::vector<double>& z) {
    const double* zr = z.data(); const double* zi = zr + (real_proj ? 0 : (size_t)nprojs * ndat);
    dgemm("N", "N", &npw, &ndat, &nprojs, &one, Pr, &npw, zr, &nprojs, &zero, wr.data(), &npw);
    dgemm("N", "N", &npw, &ndat, &nprojs, &one, Pi, &npw, zr, &nprojs, &zero, wi.data(), &npw);
    if (!real_proj) {
      dgemm("N", "N", &npw, &ndat, &nprojs, &mone, Pi, &npw, zi, &nprojs, &one, wr.data(), &npw);
      dgemm("N", "N", &npw, &ndat, &nprojs, &one, Pr, &npw, zi, &nprojs, &one, wi.data(), &npw);
    }
  };
  back(gxfac);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)nw; q++) {
    const int p = (int)(q % npw);
    const double k = kinpw[p];
    gh[q] = k < kin_filter ? gh[q] + k * cw[q] + cplx(wr[q], wi[q]) : cplx(0.0, 0.0);
  }
  if (sij_opt == 1 && gsc != nullptr) {
    back(gxs);
    cplx* gs = reinterpret_cast<cplx*>(gsc);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)nw; q++) {
      const int p = (int)(q % npw);
      gs[q] = kinpw[p] < kin_filter ? cw[q] + cplx(wr[q], wi[q]) : cplx(0.0, 0.0);
    }
  }
  double t4 = now();
  if (timings) { timings[0] = t1 - t0; timings[1] = t2 - t1; timings[2] = t3 - t2; timings[3] = t4 - t3; }
  return 0;
}

int cref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void cref_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"
