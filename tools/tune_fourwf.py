#!/usr/bin/env python
"""Developer tool: time fourwf option 2 (device-resident arrays) for one workload under several tunings.
   python tools/tune_fourwf.py --workload si512 --ndat 64 --istwfk 2 --set plane_cfg=1,2,3 --set plane_ctas_per_sm=0,1"""
import argparse, itertools, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="si512"); ap.add_argument("--ndat", type=int, default=64)
    ap.add_argument("--istwfk", type=int, default=2); ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--set", action="append", default=[])
    a = ap.parse_args()
    import torch
    import abinit_b200 as ab
    from abinit_b200 import api, workload as wl
    ab.init(0)
    cfg = wl.CONFIGS[a.workload]
    kg, kin = wl.gsphere_orthorhombic(cfg["ecut"], cfg["L"], (0, 0, 0), a.istwfk)
    npw = kg.shape[0]; ng = cfg["ngfft"]
    v = torch.from_numpy(wl.smooth_potential(ng, 1)).cuda()
    c = torch.randn((a.ndat, npw, 2), dtype=torch.float64, device="cuda")
    if a.istwfk == 2:
        c[:, 0, 1] = 0
    out = torch.zeros_like(c)
    knobs = [(s.split("=")[0], [int(x) for x in s.split("=")[1].split(",")]) for s in a.set]
    names = [k for k, _ in knobs]
    for combo in itertools.product(*[vals for _, vals in knobs]) if knobs else [()]:
        for k, val in zip(names, combo):
            api.set_tuning(k, val)
        def run():
            api.fourwf(1, v, c, out, None, None, None, a.istwfk, kg, kg, max(ng), None, a.ndat, ng, npw, npw, ng[0], ng[1], ng[2], 2, impl=2)
        run(); torch.cuda.synchronize()
        api.profile_enable(True)
        for _ in range(a.reps):
            run()
        prof = api.profile_collect(); api.profile_enable(False)
        ms = {k: round(t / a.reps, 3) for k, (t, n) in prof.items()}
        tot = sum(ms.values())
        print(json.dumps({"tuning": dict(zip(names, combo)), "ndat": a.ndat, "npw": npw, "ms": ms, "total_ms": round(tot, 3),
                          "us_per_band": round(1e3 * tot / a.ndat, 2)}), flush=True)


if __name__ == "__main__":
    main()
