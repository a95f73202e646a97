// The flattened gs_hamiltonian_type (src/66_nonlocal/m_hamiltonian.F90:99-467) behind the opaque C handle.
#pragma once
#include "fourwf.cuh"
#include "nonlop.cuh"
#include <vector>

namespace abi {
// invovl_kpt_type (src/66_wfs/m_invovl.F90:120-150): per-k data of the PAW inverse overlap, device resident
struct Invovl {
  int nprojs = -1, cplx = 1, lmnmax = 0, ntypat = 0;
  long long ldgram = 0;               // leading dimension of gram_projs in its own elements
  double* d_inv_sij = nullptr;        // [ntypat][lmnmax][lmnmax] (cplx interleaved), row-major
  double* d_inv_s_approx = nullptr;   // same shape
  double* d_gram = nullptr;           // (cplx, ldgram, nprojs) column-major P^H P
  void release();
};
}  // namespace abi

namespace abi { inline unsigned& ham_epoch_counter() { static unsigned c = 0; return c; } }

struct abi_b200_ham {
  // (the members are abi:: types; the struct itself lives in the global namespace because the C header names it)
  int ngfft[18];
  int natom, ntypat, lmnmax, usepaw;
  double ucvol;
  abi::NonlopAtoms atoms;
  abi::NonlopEnl enl;
  abi::Projectors P;
  abi::VlocDev vloc;
  // nspinor = 2 (m_hamiltonian.F90 nspinor / nvloc): nvloc = 4 keeps V22 and the two complex off-diagonal potentials
  int nspinor = 1, nvloc = 1;
  abi::VlocDev vloc22, vloc_ud, vloc_du;   // V22 (real); (V3 + i V4) applied to psi_dn -> ghc_up; (V3 - i V4) applied to psi_up -> ghc_dn
  double* d_spin_tmp = nullptr; size_t spin_tmp_cap = 0;
  int istwf_k = 1, npw = 0, me_g0 = 1;
  std::vector<int> kg;
  double* d_kinpw = nullptr;
  std::shared_ptr<abi::FourwfPlan> plan_ref;   // keeps the plan alive whatever happens to the plan cache
  abi::FourwfPlan* plan = nullptr;
  double* d_gvnlxc = nullptr; size_t gvnlxc_cap = 0;
  // bumped by every load_* / set_*: invalidates the CUDA graphs captured for this handle.  Values are drawn from ONE process-wide
  // counter, so a handle allocated at the address of a destroyed one can never match the graph keys (handle, epoch, arrays) of its
  // predecessor.
  unsigned epoch = ++abi::ham_epoch_counter();
  abi::Invovl invovl;                 // built lazily by apply_invovl, dropped by load_k / load_enl / set_projectors
};

namespace abi {
void make_invovl(abi_b200_ham* h, cudaStream_t st);
void apply_invovl_device(abi_b200_ham* h, const double* cwavef, double* sm1cwavef, double* cprj, int ndat, cudaStream_t st);
void invovl_release_workspace();
}  // namespace abi
