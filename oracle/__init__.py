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
  * ``nonlop.gemm_nonlop`` and ``getghc.getghc`` -- PARITY UNPINNED at vector level: the reference holds no
    stand-alone golden vectors for them (only SCF-level observables that need a full Fortran build, which this
    environment cannot produce: no Fortran compiler).  They are checked by mathematical invariants only
    (naive per-atom sum, Hermiticity, istwfk=2 == istwfk=1 on the completed sphere).
"""
