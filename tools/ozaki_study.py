#!/usr/bin/env python
"""Feasibility study (CPU, NumPy): error-free int8 slicing (Ozaki scheme I) of the gemm_nonlop contractions.

gx = P^T psi (K = 2 npw long) and vect = P z (K = nprojs) are at the FP64 peak of the B200 with DMMA; the only way past it
is fewer FP64 flops.  Split every row of A and column of B against its own power-of-two scale into S signed slices of
`bits` bits; every slice product A_s^T B_t is EXACT in int32 (|a_s b_t| K < 2^31) and runs on the int8 tensor pipe
(tcgen05 kind::i8, ~4.5 Pop/s dense on B200 vs 0.037 PFLOP/s FP64); the FP64 result is sum_{s+t < S} 2^(e_a+e_b-bits(s+t+2)) (A_s^T B_t).
This script measures, on data shaped like the Si-512 workload (random unit-norm projectors, wavefunctions with the
1/(1+kin) spectral decay), the relative error of gx as a function of the number of slices -- i.e. how many int8 GEMMs the
1e-11 north-star tolerance costs.  It is a study tool: nothing in abinit_b200/ imports it."""
import numpy as np


def slices(x, axis, nsl, bits):
    """x = sum_s q_s * 2^(e - bits*(s+1)) along `axis`-wise scales; q_s integer in [-2^(bits-1), 2^(bits-1)]."""
    amax = np.max(np.abs(x), axis=axis, keepdims=True)
    e = np.ceil(np.log2(np.maximum(amax, 1e-300))) + 1          # |x| / 2^e < 1/2
    r = x / 2.0 ** e
    out = []
    for s in range(nsl):
        r = r * 2.0 ** bits
        q = np.rint(r)
        r = r - q
        out.append(q)                                            # integer-valued float64: products below stay exact (< 2^53)
    return out, e


def ozaki_gemm(a, b, nsl, bits):
    """a: (K, M) columns scaled per column; b: (K, N); returns a^T b emulated with slice products s + t < nsl."""
    qa, ea = slices(a, 0, nsl, bits)
    qb, eb = slices(b, 0, nsl, bits)
    K = a.shape[0]
    assert K * (2 ** (bits - 1)) ** 2 < 2 ** 31, "int32 accumulator would overflow: split K"
    acc = np.zeros((a.shape[1], b.shape[1]))
    nprod = 0
    for g in range(nsl - 1, -1, -1):                             # smallest terms first
        part = np.zeros((a.shape[1], b.shape[1]))
        for s in range(g + 1):
            part += qa[s].T @ qb[g - s]
            nprod += 1
        acc += part * 2.0 ** (-bits * (g + 2))
    return acc * 2.0 ** ea.T * 2.0 ** eb, nprod


def main():
    rng = np.random.default_rng(0)
    K, M, N = 2 * 20000, 96, 48                                  # reduced K (the error model scales like sqrt(K))
    kin = np.sort(rng.uniform(0, 20, K // 2)) ; kin = np.repeat(kin, 2)
    P = rng.standard_normal((K, M)) / np.sqrt(K / 2)
    psi = rng.standard_normal((K, N)) / (1.0 + kin)[:, None]
    psi /= np.linalg.norm(psi, axis=0, keepdims=True)
    ref = P.T @ psi
    print(f"K={K} M={M} N={N}  max|psi|/rms = {np.max(np.abs(psi)) / np.sqrt(np.mean(psi ** 2)):.1f}")
    print(f"{'bits':>4s} {'slices':>6s} {'int8 GEMMs':>10s} {'max rel err (per column norm)':>30s}")
    for bits in (6, 7):
        for nsl in range(4, 10):
            got, nprod = ozaki_gemm(P, psi, nsl, bits)
            err = np.max(np.linalg.norm(got - ref, axis=0) / np.linalg.norm(ref, axis=0))
            print(f"{bits:4d} {nsl:6d} {nprod:10d} {err:30.3e}")


if __name__ == "__main__":
    main()
