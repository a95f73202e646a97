#!/usr/bin/env python
"""Developer tool: BASELINE configs[3] shape (Au-108-like PAW, box 96^3, Gamma handled as istwf_k = 1 like the reference's
tparal_bandpw_03 case, gemm_nonlop dominated) on one GPU, device-resident: getghc with gsc (sij_opt = 1: three GEMMs + the
packed D_ij / S_ij kernel) and one ChebFi2-PAW call (getghc + apply_invovl per degree + Rayleigh-Ritz with hegvd).
Two projector sets: 18 per atom (standard JTH, nprojs 1944) and 32 per atom (semicore-like, nprojs 3456).
   python tools/au108_paw_bench.py > gpurun_out/au108_paw.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import abinit_b200 as ab
from abinit_b200 import api, xg, workload as wl
import bench
import test_full_size_gpu as fs


def main():
    ab.init(0)
    hbm, _, fp64, _ = bench.peaks()
    for tag, lmax, ndat, nband in (("au108_jth18", 2, 128, 768), ("au108_semicore32", 3, 128, 768)):
        wl.CONFIGS[tag] = dict(wl.CONFIGS["au108"], lmax=lmax)
        cfg, h, cw, kg, kinpw, npw, nprojs = fs._setup(tag, 1, ndat, usepaw=1, seed=3, sentinel=False)
        ghc = torch.zeros_like(cw); gsc = torch.zeros_like(cw)
        torch.cuda.synchronize()
        api.set_async(True)
        step = lambda: ab.getghc(-1, cw, None, ghc, gsc, h, None, None, None, ndat, sij_opt=1)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream()
        steps = 10
        api.set_async(False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3 / steps
        api.profile_enable(True)
        for _ in range(steps):
            step()
        prof = api.profile_collect(); api.profile_enable(False)
        kms = {k: t / steps for k, (t, c) in prof.items()}
        t_nl = sum(v for k, v in kms.items() if k.startswith("dgemm"))
        f_nl = 3 * 8.0 * npw * nprojs                       # g = 3 (PAW with gsc), complex
        t_fw = sum(kms.get(k, 0.0) for k in ("fourwf_x_forward", "fourwf_plane_stage", "fourwf_x_backward"))
        out = {"case": tag, "ngfft": cfg["ngfft"], "npw": npw, "nprojs": nprojs, "ndat": ndat, "istwfk": 1, "usepaw": 1, "sij_opt": 1,
               "ms_per_step": ms, "band_app_per_s": ndat / (ms * 1e-3), "kernel_ms": {k: round(v, 4) for k, v in kms.items()},
               "gemm_nonlop": {"ms": t_nl, "TFLOPs": f_nl * ndat / (t_nl * 1e-3) / 1e12, "fp64_frac": f_nl * ndat / (t_nl * 1e-3) / 1e12 / fp64},
               "fourwf_us_per_band": 1e3 * t_fw / ndat}
        # one ChebFi2-PAW call on nband bands (second call timed: the first one orthonormalises the random block)
        gen = torch.Generator(device=cw.device).manual_seed(7)
        x = torch.randn((nband, npw, 2), generator=gen, device=cw.device, dtype=torch.float64)
        x *= torch.from_numpy(1.0 / (1.0 + kinpw)).to(cw.device)[None, :, None]
        eig = np.zeros(nband); resid = np.zeros(nband)
        torch.cuda.synchronize()
        xg.chebfiwf2(x, eig, None, None, h, nband, npw, 1, resid, 1e-16, cfg["ecut"], 4, bandpp=128)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        xg.chebfiwf2(x, eig, None, None, h, nband, npw, 1, resid, 1e-16, cfg["ecut"], 4, bandpp=128)
        torch.cuda.synchronize()
        out["chebfi2_paw_call"] = {"nband": nband, "nline": 4, "bandpp": 128, "seconds": time.perf_counter() - t0,
                                   "eig_min_max": [float(eig.min()), float(eig.max())], "resid_max": float(resid.max())}
        print(json.dumps(out), flush=True)
        h.destroy(); del cw, ghc, gsc, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
