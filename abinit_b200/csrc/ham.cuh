// The flattened gs_hamiltonian_type (src/66_nonlocal/m_hamiltonian.F90:99-467) behind the opaque C handle.
#pragma once
#include "fourwf.cuh"
#include "nonlop.cuh"
#include <vector>

struct abi_b200_ham {
  // (the members are abi:: types; the struct itself lives in the global namespace because the C header names it)
  int ngfft[18];
  int natom, ntypat, lmnmax, usepaw;
  double ucvol;
  abi::NonlopAtoms atoms;
  abi::NonlopEnl enl;
  abi::Projectors P;
  abi::VlocDev vloc;
  int istwf_k = 1, npw = 0, me_g0 = 1;
  std::vector<int> kg;
  double* d_kinpw = nullptr;
  abi::FourwfPlan* plan = nullptr;
  double* d_gvnlxc = nullptr; size_t gvnlxc_cap = 0;
};
