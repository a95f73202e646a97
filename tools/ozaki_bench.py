#!/usr/bin/env python
"""OPT-IN path: Si-512 getghc with gemm_nonlop through int8 slice products (csrc/ozaki.cu + igemm_tc.cuh) vs the default FP64 DMMA path:
speed and agreement at full size.  python tools/ozaki_bench.py   (on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import json
import numpy as np, torch
import abinit_b200 as ab
from abinit_b200 import api, workload as wl

ab.init(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); api.set_stream(stream.cuda_stream)
cfg = wl.CONFIGS["si512"]
kg, kin = wl.gsphere_orthorhombic(cfg["ecut"], cfg["L"], (0, 0, 0), 2)
npw = kg.shape[0]; ndat = 128
indlmn, lnmax = wl.nc_indlmn(cfg["lmax"], cfg["nproj_per_l"]); nlmn = indlmn.shape[1]; natom = cfg["natom"]; nprojs = natom * nlmn
h = ab.Hamiltonian(cfg["ngfft"], natom, 1, nlmn, indlmn, np.array([natom], dtype=np.int32), np.arange(1, natom + 1, dtype=np.int32), 0, cfg["L"] ** 3)
h.load_spin(wl.smooth_potential(cfg["ngfft"], seed=5), 1)
h.load_enl(np.random.default_rng(1).standard_normal((1, lnmax)), None)
h.load_k(2, kg, kin, None, None, me_g0=1)
with torch.cuda.stream(stream):
    gen = torch.Generator(device=dev).manual_seed(4321)
    P = torch.randn((nprojs, npw, 2), generator=gen, device=dev, dtype=torch.float64) / np.sqrt(npw); P[:, 0, 1] = 0
    cw = torch.randn((ndat, npw, 2), generator=gen, device=dev, dtype=torch.float64) * torch.from_numpy(1 / (1 + kin)).to(dev)[None, :, None]
    cw[:, 0, 1] = 0
    out = {k: torch.zeros_like(cw) for k in (0, 1)}
stream.synchronize()
h.set_projectors(P, nprojs); del P; torch.cuda.empty_cache()
api.set_async(True)
res = {}
for mode in (0, 1):
    api.set_tuning("nonlop_ozaki", mode)
    for _ in range(3): ab.getghc(-1, cw, None, out[mode], None, h, None, None, None, ndat)
    stream.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): ab.getghc(-1, cw, None, out[mode], None, h, None, None, None, ndat)
    e1.record(stream); stream.synchronize()
    ms = e0.elapsed_time(e1) / 5
    api.profile_enable(True)
    for _ in range(2): ab.getghc(-1, cw, None, out[mode], None, h, None, None, None, ndat)
    prof = api.profile_collect(); api.profile_enable(False)
    res[mode] = (ms, {k: v[0] / 2 for k, v in prof.items()})
    print(f"ozaki={mode}: {ms:.2f} ms/step, {ndat / ms * 1e3:.0f} band-app/s;  " + ", ".join(f"{k} {v[0] / 2:.2f}" for k, v in prof.items()), flush=True)
d = out[1] - out[0]
rel = (torch.linalg.norm(d.reshape(ndat, -1), dim=1) / torch.linalg.norm(out[0].reshape(ndat, -1), dim=1)).max().item()
print(f"max per-band relative difference int8-sliced vs FP64 DMMA: {rel:.3e}")
print(f"memory in use: {torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9:.1f} GB")
if "--json" in sys.argv:
    print(json.dumps({"name": "int8-sliced gemm_nonlop (Ozaki scheme I: 7 slices x 7 bits = 28 exact int8 GEMMs per contraction), opt-in",
                      "value": ndat / (res[1][0] * 1e-3), "unit": "band-applications/s", "ms_per_step": res[1][0],
                      "fp64_path_ms_per_step_same_process": res[0][0], "max_rel_diff_vs_fp64_path": rel, "kernel_ms": res[1][1],
                      "int8_gemm": "hand-written tcgen05 kind::i8 kernel (csrc/igemm_tc.cuh: TMA 128B-swizzled K-major tiles, mbarrier ring, int32 accumulators in TMEM)",
                      "default": "off -- the default product path is the FP64 DMMA kernel; enable with ABI_B200_OZAKI=1",
                      "extra_memory_gb": 2 * 7 * nprojs * 2 * npw / 1e9}))
