"""Probe (not product): cuBLAS DGEMM throughput on this B200 = the FP64 'tensor' roofline denominator.
Writes gpurun_out/fp64_peak.json. Protocol mirrors MEASURED_PEAKS.json: best-of-10 burst and 4 s sustained."""
import json, time, torch, subprocess, os
dev = torch.device("cuda:0")
res = {"gpu": torch.cuda.get_device_name(0)}
def bench(m, n, k, reps=10, sustained=0.0):
    a = torch.randn(m, k, device=dev, dtype=torch.float64)
    b = torch.randn(k, n, device=dev, dtype=torch.float64)
    c = torch.empty(m, n, device=dev, dtype=torch.float64)
    for _ in range(3): torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out = {"burst_tflops": 2.0 * m * n * k / best / 1e9, "ms": best}
    if sustained > 0:
        t0 = time.time(); n_it = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < sustained:
            for _ in range(10): torch.matmul(a, b, out=c)
            n_it += 10
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        out["sustained_tflops"] = 2.0 * m * n * k * n_it / e0.elapsed_time(e1) / 1e9
    del a, b, c
    return out
res["dgemm_8192"] = bench(8192, 8192, 8192, sustained=4.0)
# gemm_nonlop shapes at Gamma for Si-512: K = 2*npw = 288114, nprojs = 9216, ndat = 128
res["opernla_TN_9216x128x288114"] = None
a = torch.randn(9216, 288114, device=dev, dtype=torch.float64)   # row-major == P^T with K contiguous
b = torch.randn(128, 288114, device=dev, dtype=torch.float64)
for name, f in (("opernla_9216x128x288114", lambda: torch.matmul(a, b.t())),
                ("opernlb_288114x128x9216", lambda: torch.matmul(a.t(), torch.ones(9216, 128, device=dev, dtype=torch.float64)))):
    for _ in range(2): f()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    res[name] = {"tflops": 2.0 * 9216 * 128 * 288114 / best / 1e9, "ms": best}
del a, b
q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
res["smi_after"] = q
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(res))
