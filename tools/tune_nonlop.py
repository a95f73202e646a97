#!/usr/bin/env python
"""Developer tool: time gemm_nonlop (device-resident arrays, explicit random projectors) through the C-ABI.
   python tools/tune_nonlop.py --npw 144057 --nprojs 9216 --ndat 128 --istwfk 2"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npw", type=int, default=144057); ap.add_argument("--nprojs", type=int, default=9216)
    ap.add_argument("--ndat", type=int, default=128); ap.add_argument("--istwfk", type=int, default=2)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import abinit_b200 as ab
    from abinit_b200 import api
    ab.init(0)
    npw, nprojs, ndat = a.npw, a.nprojs, a.ndat
    P = torch.randn((nprojs, npw, 2), device="cuda", dtype=torch.float64) / np.sqrt(npw)
    c = torch.randn((ndat, npw, 2), device="cuda", dtype=torch.float64)
    if a.istwfk == 2:
        P[:, 0, 1] = 0; c[:, 0, 1] = 0
    api.set_projectors(1, npw, nprojs, a.istwfk, P); api.set_gemm_nonlop_ikpt(1)
    del P
    out = torch.zeros_like(c)
    indlmn = np.zeros((1, 1, 6), dtype=np.int32); indlmn[0, 0] = (0, 0, 1, 1, 1, 1)
    nattyp = np.array([nprojs], dtype=np.int32); atindx1 = np.arange(1, nprojs + 1, dtype=np.int32)
    enl = np.ones((1, 1))

    def run():
        api.gemm_nonlop(atindx1, 1, -1, None, enl, indlmn, a.istwfk, None, nprojs, nattyp, ndat, npw, npw, 1, 1, 0, None, None, c, out)
    run(); torch.cuda.synchronize()
    api.profile_enable(True)
    for _ in range(a.reps):
        run()
    prof = api.profile_collect(); api.profile_enable(False)
    cplx = 2 if a.istwfk == 1 else 1
    fl = 2.0 * (2 * npw) * nprojs * ndat * cplx
    res = {k: {"ms": round(t / n, 3), "tflops": round(fl / (t / n * 1e-3) / 1e12, 2) if "dgemm" in k else None} for k, (t, n) in prof.items()}
    print(json.dumps({"npw": npw, "nprojs": nprojs, "ndat": ndat, "istwfk": a.istwfk, "kernels": res}))


if __name__ == "__main__":
    main()
