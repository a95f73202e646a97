#!/usr/bin/env python
"""Developer tool: BASELINE configs[4] "synthetic getghc sweep" (SURVEY 8d): cubic FFT boxes 48^3..192^3 at boxcut 2,
npw ~ 0.065 N, nprojs / npw as in Si-512, band blocks 64 / 256, one GPU, device-resident arrays.  One JSON line per point with
the per-kernel-class device times of a getghc step and the two roofline fractions (fourwf: algorithmic bytes / HBM peak;
gemm_nonlop: flops / FP64 peak).  These are parity-test shapes, not bench lines (bench.py measures Si-512).
   python tools/sweep.py [--boxes 48,64,96] [--blocks 64,256] [--istwfk 2] > gpurun_out/sweep.jsonl"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--boxes", default="48,64,96,128,144,192")
    ap.add_argument("--blocks", default="64,256")
    ap.add_argument("--istwfk", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import bench
    import abinit_b200 as ab
    from abinit_b200 import api, workload as wl
    ab.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    api.set_stream(stream.cuda_stream)
    hbm, _, fp64, _ = bench.peaks()
    for n in [int(x) for x in a.boxes.split(",")]:
        # sphere of radius r index units in a box n >= 4 r + 1 (boxcut 2); ecut 20 Ha fixes the cell length
        r = (n - 1) / 4.0 - 0.25
        L = 2 * np.pi * r / np.sqrt(2 * 20.0)
        name = f"sweep{n}"
        npw_full = 4.0 / 3.0 * np.pi * r ** 3
        natom = max(8, int(round(0.032 * npw_full / 18)))
        wl.CONFIGS[name] = dict(ecut=20.0, L=float(L), ngfft=(n, n, n), natom=natom, lmax=2, nproj_per_l=2)
        args = argparse.Namespace(workload=name, istwfk=a.istwfk)
        w = bench.build_workload(args)
        npw, nprojs = w["npw"], w["nprojs"]
        ham = ab.Hamiltonian(w["ngfft"], w["natom"], 1, w["nlmn"], w["indlmn"], w["nattyp"], w["atindx1"], 0, w["ucvol"])
        ham.load_spin(w["vlocal"], 1); ham.load_enl(w["ekb"], None)
        ham.load_k(a.istwfk, w["kg"], w["kinpw"], None, None, me_g0=1)
        with torch.cuda.stream(stream):
            gen = torch.Generator(device=dev).manual_seed(4321)
            P = torch.randn((nprojs, npw, 2), generator=gen, device=dev, dtype=torch.float64) / np.sqrt(npw)
            if a.istwfk == 2:
                P[:, 0, 1] = 0.0
        stream.synchronize()
        ham.set_projectors(P, nprojs)
        del P
        for ndat in [int(x) for x in a.blocks.split(",")]:
            with torch.cuda.stream(stream):
                cw = torch.randn((ndat, npw, 2), generator=gen, device=dev, dtype=torch.float64)
                if a.istwfk == 2:
                    cw[:, 0, 1] = 0.0
                ghc = torch.zeros_like(cw)
            stream.synchronize()
            api.set_async(True)
            step = lambda: ab.getghc(-1, cw, None, ghc, None, ham, None, None, None, ndat)
            for _ in range(3):
                step()
            stream.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(a.steps):
                step()
            e1.record(stream); stream.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            api.profile_enable(True)
            for _ in range(a.steps):
                step()
            prof = api.profile_collect(); api.profile_enable(False)
            api.set_async(False)
            kms = {k: t / a.steps for k, (t, c) in prof.items()}
            b_fw, f_nl, C = bench.algorithmic_units(w, a.istwfk, ndat)
            t_fw = sum(kms.get(k, 0.0) for k in ("fourwf_x_forward", "fourwf_plane_stage", "fourwf_plane_cluster", "fourwf_x_backward"))
            t_nl = sum(kms.get(k, 0.0) for k in ("dgemm_tn_opernla", "dgemm_nn_opernlb"))
            out = {"box": n, "npw": npw, "nprojs": nprojs, "ndat": ndat, "istwfk": a.istwfk, "ms_per_step": ms,
                   "band_app_per_s": ndat / (ms * 1e-3), "kernel_ms": {k: round(v, 4) for k, v in kms.items()},
                   "fourwf": {"ms": t_fw, "us_per_band": 1e3 * t_fw / ndat, "bytes_per_band": b_fw,
                              "GBps": b_fw * ndat / (t_fw * 1e-3) / 1e9 if t_fw else None,
                              "hbm_frac": b_fw * ndat / (t_fw * 1e-3) / 1e9 / hbm if t_fw else None},
                   "gemm_nonlop": {"ms": t_nl, "flops_per_band": f_nl, "TFLOPs": f_nl * ndat / (t_nl * 1e-3) / 1e12 if t_nl else None,
                                   "fp64_frac": f_nl * ndat / (t_nl * 1e-3) / 1e12 / fp64 if t_nl else None}}
            print(json.dumps(out), flush=True)
            del cw, ghc
        ham.destroy()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
