// PAW inverse overlap S^-1 = 1 - P (s^-1 + P^H P)^-1 P^H (Woodbury), as used by ChebFi2's getBm1X.
// Reference semantics (not code): src/66_wfs/m_invovl.F90:469-776 (make_invovl), :790-1039 (apply_invovl),
// :1052-1152 (solve_inner: preconditioned fixed-point iteration on the projections), :1165-1231 (apply_block).
//
// Design: the projector Gram matrix P^H P is one DMMA GEMM on the resident projector matrix (the SPACE_CR Gram with the G=0
// convention reproduces the sqrt(2)-scaled half-sphere projectors of m_invovl.F90:601-607); the inner solve runs entirely on
// the device on the padded projection layout of nonlop.cu (one small GEMM + two fused block-diagonal kernels per
// iteration); only the ndat convergence numbers cross to the host per iteration.
#include "ham.cuh"
#include "xg.cuh"
#include "context.cuh"
#include <algorithm>
#include <complex>
#include <vector>

namespace abi {

#ifndef ABI_EMU
namespace {

typedef std::complex<double> cd;

// in-place inverse of a small dense matrix (Gauss-Jordan, partial pivoting); the reference uses xSYTRF/xSYTRI
// (m_invovl.F90:571-577): same matrix to rounding
void invert_small(std::vector<cd>& a, int n) {
  std::vector<cd> inv((size_t)n * n, cd(0.0));
  for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; c++) {
    int piv = c; double best = std::abs(a[(size_t)c * n + c]);
    for (int r = c + 1; r < n; r++) if (std::abs(a[(size_t)r * n + c]) > best) { best = std::abs(a[(size_t)r * n + c]); piv = r; }
    ABI_CHECK(best > 0.0, "make_invovl: singular overlap block");
    if (piv != c) for (int k = 0; k < n; k++) { std::swap(a[(size_t)c * n + k], a[(size_t)piv * n + k]); std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]); }
    const cd d = 1.0 / a[(size_t)c * n + c];
    for (int k = 0; k < n; k++) { a[(size_t)c * n + k] *= d; inv[(size_t)c * n + k] *= d; }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      const cd f = a[(size_t)r * n + c];
      if (f == cd(0.0)) continue;
      for (int k = 0; k < n; k++) { a[(size_t)r * n + k] -= f * a[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; }
    }
  }
  a.swap(inv);
}

// y (+)= M_type(atom) x per (sorted atom, band); mats: [ntypat][lmnmax][lmnmax] row-major complex (re,im) or real
// mode 0: y = M x ; 1: y += M x ; 2: y = -x_in - 0 (unused)
__global__ void k_block_apply(int cplx, int lmnmax, const double* __restrict__ mats, const int* __restrict__ atom_first,
                              const int* __restrict__ atom_typ, const double* __restrict__ x, double* __restrict__ y, long long ldg,
                              int accumulate) {
  const int a = blockIdx.x, n = blockIdx.y;
  const int first = atom_first[a], nlmn = atom_first[a + 1] - first;
  const double* M = mats + (size_t)cplx * lmnmax * lmnmax * atom_typ[a];
  const double* xv = x + (size_t)n * ldg + (size_t)cplx * first;
  double* yv = y + (size_t)n * ldg + (size_t)cplx * first;
  for (int i = threadIdx.x; i < nlmn; i += blockDim.x) {
    if (cplx == 1) {
      double s = 0.0;
      for (int j = 0; j < nlmn; j++) s += M[i * lmnmax + j] * xv[j];
      yv[i] = accumulate ? yv[i] + s : s;
    } else {
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < nlmn; j++) {
        const double mr = M[2 * (i * lmnmax + j)], mi = M[2 * (i * lmnmax + j) + 1];
        sr += mr * xv[2 * j] - mi * xv[2 * j + 1];
        si += mr * xv[2 * j + 1] + mi * xv[2 * j];
      }
      yv[2 * i] = accumulate ? yv[2 * i] + sr : sr;
      yv[2 * i + 1] = accumulate ? yv[2 * i + 1] + si : si;
    }
  }
}

// one CTA per band: resid = proj - t - ptp (when t != null), out[n] = sum resid^2 (or sum proj^2 when t == null)
__global__ void k_resid_norm(int nreal, const double* __restrict__ proj, const double* __restrict__ t, const double* __restrict__ ptp,
                             double* __restrict__ resid, long long ldg, double* __restrict__ out) {
  __shared__ double red[256];
  const int n = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < nreal; i += blockDim.x) {
    const size_t o = (size_t)n * ldg + i;
    double r = proj[o];
    if (t) { r = r - t[o] - ptp[o]; resid[o] = r; }
    s += r * r;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) { if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) out[n] = red[0];
}

// z = -x ; optional cprj(cplex, nprojs, ndat) = proj - ptp  (m_invovl.F90:938-939, 1004-1023)
__global__ void k_finish(int nreal, const double* __restrict__ x, double* __restrict__ z, const double* __restrict__ proj,
                         const double* __restrict__ ptp, double* __restrict__ cprj, long long ldg, int ndat) {
  const long long total = (long long)nreal * ndat;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / nreal; const int i = (int)(idx % nreal);
    const size_t o = (size_t)n * ldg + i;
    z[o] = -x[o];
    if (cprj) cprj[idx] = proj[o] - ptp[o];
  }
}

struct Ws { double* d = nullptr; size_t cap = 0;
  double* get(size_t n) { if (n > cap) { if (d) cudaFree(d); CUDA_CHECK(cudaMalloc(&d, sizeof(double) * n)); cap = n; } return d; }
  void release() { if (d) cudaFree(d); d = nullptr; cap = 0; } };
Ws g_iv[7];   // proj, x, resid, tmp, ptp, errs, z

}  // namespace

void invovl_release_workspace() { for (auto& w : g_iv) w.release(); }

void Invovl::release() {
  if (d_inv_sij) cudaFree(d_inv_sij);
  if (d_inv_s_approx) cudaFree(d_inv_s_approx);
  if (d_gram) cudaFree(d_gram);
  d_inv_sij = d_inv_s_approx = d_gram = nullptr; nprojs = -1;
}

// make_invovl (m_invovl.F90:469-776) from the projectors already resident in the handle
void make_invovl(abi_b200_ham* h, cudaStream_t st) {
  ABI_CHECK(h->usepaw == 1, "make_invovl: PAW only");
  ABI_CHECK(h->P.d_p != nullptr, "make_invovl: projectors not loaded (load_k)");
  ABI_CHECK(h->enl.d_sij != nullptr, "make_invovl: sij not loaded (load_enl)");
  Invovl& iv = h->invovl;
  iv.release();
  const int cplx = h->istwf_k == 1 ? 2 : 1, nprojs = h->atoms.nprojs, lmnmax = h->lmnmax, ntypat = h->ntypat;
  iv.cplx = cplx; iv.nprojs = nprojs; iv.lmnmax = lmnmax; iv.ntypat = ntypat;
  iv.ldgram = cplx == 2 ? nprojs : ((nprojs + 1) & ~1);
  // gram_projs = P^H P on the full sphere (:706-745)
  CUDA_CHECK(cudaMalloc(&iv.d_gram, sizeof(double) * cplx * (size_t)iv.ldgram * std::max(1, nprojs)));
  CUDA_CHECK(cudaMemsetAsync(iv.d_gram, 0, sizeof(double) * cplx * (size_t)iv.ldgram * std::max(1, nprojs), st));
  const int space = cplx == 2 ? SPACE_C : SPACE_CR;
  const int me_g0 = cplx == 2 ? -1 : ((h->istwf_k == 2 && h->me_g0 == 1) ? 1 : 0);
  xg_gram(space, h->npw, nprojs, nprojs, h->P.d_p, h->npw, h->P.d_p, h->npw, iv.d_gram, iv.ldgram, me_g0, st);
  // s_ij per type -> inv_sij ; inv_s_approx = (inv_sij + Gram of the first atom of the type)^-1 (:560-700)
  const int dimenl1 = h->enl.sij_dim1 > 0 ? h->enl.sij_dim1 : h->enl.dimenl1;       // S_ij is real whatever cplex_dij is
  ABI_CHECK(dimenl1 == lmnmax * (lmnmax + 1) / 2, "make_invovl: sij size not recognized (real packed sij only)");
  std::vector<double> sij((size_t)dimenl1 * ntypat);
  CUDA_CHECK(cudaMemcpyAsync(sij.data(), h->enl.d_sij, sizeof(double) * sij.size(), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<double> inv_sij((size_t)cplx * lmnmax * lmnmax * ntypat, 0.0), inv_app(inv_sij.size(), 0.0);
  int shift = 0;
  for (int t = 0; t < ntypat; t++) {
    const int nlmn = h->atoms.nlmn[t];
    std::vector<cd> m((size_t)nlmn * nlmn);
    for (int j = 0; j < nlmn; j++) for (int i = 0; i <= j; i++) { m[(size_t)i * nlmn + j] = m[(size_t)j * nlmn + i] = sij[(size_t)dimenl1 * t + j * (j + 1) / 2 + i]; }
    invert_small(m, nlmn);
    std::vector<cd> app = m;
    if (h->atoms.nattyp[t] > 0 && nlmn > 0) {
      // Gram block of the type's first atom, D2H
      std::vector<double> blk((size_t)cplx * nlmn * nlmn);
      CUDA_CHECK(cudaMemcpy2DAsync(blk.data(), sizeof(double) * cplx * nlmn, iv.d_gram + (size_t)cplx * ((size_t)shift * iv.ldgram + shift),
                                   sizeof(double) * cplx * iv.ldgram, sizeof(double) * cplx * nlmn, nlmn, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      for (int j = 0; j < nlmn; j++) for (int i = 0; i < nlmn; i++) {   // blk is column-major: element (i,j) at j*nlmn+i
        const cd g = cplx == 2 ? cd(blk[2 * ((size_t)j * nlmn + i)], blk[2 * ((size_t)j * nlmn + i) + 1]) : cd(blk[(size_t)j * nlmn + i], 0.0);
        app[(size_t)i * nlmn + j] += g;
      }
      invert_small(app, nlmn);
    }
    for (int i = 0; i < nlmn; i++) for (int j = 0; j < nlmn; j++) {
      const size_t o = (size_t)cplx * (((size_t)t * lmnmax + i) * lmnmax + j);
      inv_sij[o] = m[(size_t)i * nlmn + j].real(); inv_app[o] = app[(size_t)i * nlmn + j].real();
      if (cplx == 2) { inv_sij[o + 1] = m[(size_t)i * nlmn + j].imag(); inv_app[o + 1] = app[(size_t)i * nlmn + j].imag(); }
    }
    shift += nlmn * h->atoms.nattyp[t];
  }
  CUDA_CHECK(cudaMalloc(&iv.d_inv_sij, sizeof(double) * std::max<size_t>(1, inv_sij.size())));
  CUDA_CHECK(cudaMalloc(&iv.d_inv_s_approx, sizeof(double) * std::max<size_t>(1, inv_app.size())));
  CUDA_CHECK(cudaMemcpyAsync(iv.d_inv_sij, inv_sij.data(), sizeof(double) * inv_sij.size(), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(iv.d_inv_s_approx, inv_app.data(), sizeof(double) * inv_app.size(), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
}

// apply_invovl (m_invovl.F90:790-1039): sm1cwavef = S^-1 cwavef ; cprj (optional, (cplex,nprojs,ndat)) = <p|S^-1 c>-style
// projections as the reference leaves them in cwaveprj (proj - P^H P sm1proj).  Device pointers.
void apply_invovl_device(abi_b200_ham* h, const double* cwavef, double* sm1cwavef, double* cprj, int ndat, cudaStream_t st) {
  if (h->invovl.nprojs < 0) make_invovl(h, st);
  const Invovl& iv = h->invovl;
  const int cplx = iv.cplx, nprojs = iv.nprojs, natom = h->atoms.natom;
  const size_t nv = sizeof(double) * 2 * (size_t)h->npw * ndat;
  if (nprojs == 0 || ndat == 0) { if (ndat > 0 && sm1cwavef != cwavef) CUDA_CHECK(cudaMemcpyAsync(sm1cwavef, cwavef, nv, cudaMemcpyDeviceToDevice, st)); return; }
  const long long ldg = nonlop_ldg(h->P);
  const int nreal = cplx * nprojs;
  const size_t nb = (size_t)ldg * ndat;
  double *proj = g_iv[0].get(nb), *x = g_iv[1].get(nb), *resid = g_iv[2].get(nb), *tmp = g_iv[3].get(nb), *ptp = g_iv[4].get(nb);
  double* d_errs = g_iv[5].get((size_t)2 * ndat);
  double* z = g_iv[6].get(nb);
  CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * nb, st));          // pad elements stay finite (they multiply zero-filled rows)
  CUDA_CHECK(cudaMemsetAsync(ptp, 0, sizeof(double) * nb, st));
  // proj = <p|c> (nonlop choice 0, :887-897)
  nonlop_project(h->P, h->me_g0, cwavef, ndat, proj, st);
  // ---- solve_inner (:1052-1152)
  const dim3 gb(natom, ndat);
  std::vector<double> normprojs(ndat), errs(ndat);
  k_resid_norm<<<ndat, 256, 0, st>>>(nreal, proj, nullptr, nullptr, nullptr, ldg, d_errs);
  CUDA_CHECK(cudaMemcpyAsync(normprojs.data(), d_errs, sizeof(double) * ndat, cudaMemcpyDeviceToHost, st));
  k_block_apply<<<gb, 64, 0, st>>>(cplx, iv.lmnmax, iv.d_inv_s_approx, h->atoms.d_atom_first, h->atoms.d_atom_typ, proj, x, ldg, 0);
  g_kernel_launches += 2;
  const double precision = 1e-16;
  int additional = -1; double maxerr = 0.0, previous = 0.0;
  for (int i = 1; i <= 30; i++) {
    k_block_apply<<<gb, 64, 0, st>>>(cplx, iv.lmnmax, iv.d_inv_sij, h->atoms.d_atom_first, h->atoms.d_atom_typ, x, tmp, ldg, 0);
    if (cplx == 2) zgemm_nn(nprojs, ndat, nprojs, iv.d_gram, iv.ldgram, x, ldg / 2, ptp, ldg / 2, st);
    else dgemm_nn(nprojs, ndat, nprojs, iv.d_gram, iv.ldgram, x, ldg, ptp, ldg, st);
    k_resid_norm<<<ndat, 256, 0, st>>>(nreal, proj, tmp, ptp, resid, ldg, d_errs);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches += 2;
    CUDA_CHECK(cudaMemcpyAsync(errs.data(), d_errs, sizeof(double) * ndat, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    maxerr = 0.0;
    for (int n = 0; n < ndat; n++) maxerr = std::max(maxerr, errs[n] / normprojs[n]);
    maxerr = std::sqrt(maxerr);
    if (maxerr < precision || additional == 1) {
      break;
    } else if (maxerr < 1e-10 && additional == -1) {
      const double rate = -std::log(1e-10) / i;
      additional = (int)std::ceil(-std::log(precision / 1e-10) / rate) + 1;
    } else if (additional > 0) {
      if (previous < maxerr) break;
      additional--;
    }
    previous = maxerr;
    k_block_apply<<<gb, 64, 0, st>>>(cplx, iv.lmnmax, iv.d_inv_s_approx, h->atoms.d_atom_first, h->atoms.d_atom_typ, resid, x, ldg, 1);
    g_kernel_launches++;
  }
  if (maxerr >= precision && maxerr >= 1e-10)
    fprintf(stderr, "\n--- !WARNING\nmessage: |\n    In invovl, max error was %g after 30 iterations\n...\n", maxerr);
  // sm1proj = -x ; cprj = proj - P^H P x ; sm1cwavef = cwavef + P sm1proj (:938-1031)
  const int blocks = std::min(kNumSM * 8, (int)ceil_div<long long>((long long)nreal * ndat, 256));
  if (ldg != nreal) CUDA_CHECK(cudaMemsetAsync(z, 0, sizeof(double) * nb, st));
  k_finish<<<blocks, 256, 0, st>>>(nreal, x, z, proj, ptp, cprj, ldg, ndat);
  CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
  nonlop_expand(h->P, z, ndat, sm1cwavef, cwavef, st);
}

#else
void invovl_release_workspace() {}
void Invovl::release() {}
#endif

}  // namespace abi
