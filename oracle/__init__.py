"""CPU oracle for the getghc hot path (fourwf + gemm_nonlop + kinetic assembly).

TEST INFRASTRUCTURE ONLY.  This package is a NumPy/SciPy restatement of the reference's CPU
algorithm (ABINIT 10.6, Fortran; file:line citations in every function).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it -- and only as the checker or as the timed CPU baseline, never as the product path.  The product
(``abinit_b200``) never imports it and fails loudly when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * ``gsphere.getng`` / ``gsphere.kpgsph``  -- PINNED on values printed in the reference's own test outputs
    (tests/tutorial/Refs/tbase3_1.abo: ngfft 24^3, npw 519/525, mpw 525;
     tests/unitary/Refs/tfourwf_01.stdout: 100^3 box for ecut 30, 20 Bohr cube).
  * ``fourwf.fourwf``  -- PINNED on the reference's unit-test known-answer vectors
    (src/70_gw/m_fft_prof.F90:873,936-976: c(G)=exp(-(2pi)^2 G.gmet.G), V=cos(2pi g0.r), g0=(1,-1,2),
     closed form out(G)=1/2[c(G-g0)+c(G+g0)]), tolerance = the cross-library spread the reference stores
     (tests/unitary/Refs/tfourwf_01.stdout:129, 3.4e-16 abs).
  * ``nonlop.gemm_nonlop`` (NC path: prep_projectors normalisation / phases / (-i)^l, opernla/c/b) and
    ``getghc.getghc`` (local + kinetic + non-local assembly) -- PINNED on the reference's SCF golden numbers:
    oracle/scf.py runs a complete LDA ground state of the tutorial test tbase3_1 AROUND getghc (psp8 tables
    restated in oracle/psp8.py, epsatm 6.67004110 and the Ewald energy reproduced to every printed digit) and
    reproduces tests/tutorial/Refs/tbase3_1.abo: etotal -8.51873906424 Ha to 1.4e-10 Ha, kinetic / local_psp /
    non_local_psp / hartree / xc to < 3e-5 Ha (first order in the reference's own SCF residual) and the five printed
    eigenvalues at k=(-1/4,1/2,0) to print precision (tests/test_scf_pins.py).
  * istwf_k = 2 (Gamma point: time-reversal completion of the sphere, G = 0 conventions, real-projection gemm_nonlop,
    SPACE_CR xgBlock algebra, LOBPCG) -- PINNED on the tutorial test tbase1_1 (H2, Gamma only, istwfk 2, 30^3, npw 1503): the
    same SCF around getghc(istwf_k=2) reproduces tests/tutorial/Refs/tbase1_1.abo: etotal -1.11718434634432 Ha to 6e-12 Ha,
    the Ewald / psp-core terms to all digits and both printed eigenvalues (tests/test_scf_pins.py).
  * istwf_k = 3 and 7 (k = (1/2,0,0), (1/2,1/2,0)) -- PINNED with istwf_k = 2 on dataset 1 of tests/tutoplugs/Input/tw90_1.abi
    (Si-2, Gamma-centred 2x2x2 mesh, tolvrs 1e-10): etotal -8.42438318247138 Ha to 3e-12 Ha, components to 4e-7.
    The other representatives of the L and X stars of that mesh carry istwf_k 4, 5, 6, 8, 9 and give the same stored etotal
    (within 1e-9 Ha) after the density symmetrisation: every time-reversal mode is PINNED (tests/test_scf_pins.py).
  * nspinor = 2 (getghc_spinor, collinear nvloc = 1 branch) -- PINNED on the same tw90_1 SCF run in spinor form (no spin-orbit:
    every band becomes a degenerate pair of occupation 1; stored etotal within 1e-9 Ha, tests/test_scf_pins.py).  The
    non-collinear nvloc = 4 branch has no stored data without the magnetisation machinery: invariants only.
  * PAW application machinery (per-atom packed D_ij with off-diagonal terms, opernlc PAW branch, gsc assembly) -- PINNED through
    an exact rewriting of the norm-conserving tw90_1 operator (rotated projector pairs p' = R p, D' = R diag(ekb) R^T, S_ij = 0):
    the SCF through the PAW code path gives the stored etotal within 1e-9 Ha (tests/test_scf_pins.py).
  No stored per-vector dumps exist in the reference, so a non-trivial S_ij (paw_opt 3-4, apply_invovl) and cprj remain
  pinned by invariants only (naive per-atom sum, Hermiticity of H and S, S S^-1 = 1): "parity unpinned" for those branches.
"""
