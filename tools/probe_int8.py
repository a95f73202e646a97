#!/usr/bin/env python
"""Probe: int8 x int8 -> int32 GEMM throughput (cuBLASLt through torch._int_mm) at the gemm_nonlop shapes of Si-512, next to
the FP64 DGEMM rate -- the denominator of the Ozaki-slicing study (tools/ozaki_study.py)."""
import time, torch
dev = torch.device("cuda", 0)
def bench(M, N, K, reps=5):
    a = torch.randint(-64, 64, (M, K), dtype=torch.int8, device=dev)
    b = torch.randint(-64, 64, (N, K), dtype=torch.int8, device=dev).t()      # (K, N) column-major
    for _ in range(2): c = torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): c = torch._int_mm(a, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"int8 M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TOP/s", flush=True)
for N in (128, 384, 896):
    bench(9216, N, 288128)        # opernla: P_s^T [psi_0 .. psi_t]
for N in (128, 384, 896):
    bench(288128, N, 9216)        # opernlb: Pt_s^T [z_0 .. z_t]
a = torch.randn((9216, 288114), dtype=torch.float64, device=dev); b = torch.randn((288114, 128), dtype=torch.float64, device=dev)
for _ in range(2): c = a @ b
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): c = a @ b
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
print(f"fp64 cuBLAS M=9216 N=128 K=288114: {ms:.3f} ms  {2.0 * 9216 * 128 * 288114 / ms / 1e9:.1f} TFLOP/s")
