# times fourwf option 2 on a non-cubic Si-512-like box through the split plane stage vs the cluster kernel
import sys, json, numpy as np, torch
sys.path.insert(0, '/root/repo')
import abinit_b200 as ab
from abinit_b200 import api, workload as wl
ab.init(0)
for ng, L in [((180, 180, 192), (40.72, 40.72, 43.4)), ((180, 180, 180), (40.72, 40.72, 40.72))]:
    kg, kin = wl.gsphere_orthorhombic(20.0, L, (0, 0, 0), 2)
    npw = kg.shape[0]; ndat = 64
    v = torch.from_numpy(wl.smooth_potential(ng, 1)).cuda()
    c = torch.randn((ndat, npw, 2), dtype=torch.float64, device="cuda"); c[:, 0, 1] = 0
    out = torch.zeros_like(c)
    for plane in (1, 0):
        api.set_tuning("plane", plane)
        run = lambda: api.fourwf(1, v, c, out, None, None, None, 2, kg, kg, max(ng), None, ndat, ng, npw, npw, ng[0], ng[1], ng[2], 2, impl=2)
        run(); torch.cuda.synchronize()
        api.profile_enable(True)
        for _ in range(3): run()
        prof = api.profile_collect(); api.profile_enable(False)
        ms = {k: round(t / 3, 3) for k, (t, n) in prof.items()}
        print(json.dumps({"ngfft": ng, "npw": npw, "ndat": ndat, "plane": plane, "ms": ms, "total_ms": round(sum(ms.values()), 3)}), flush=True)
    api.set_tuning("plane", 1)
# cubic box through the split kernels (developer comparison: what the L2-resident S of the fused kernel buys)
ng = (180, 180, 180); L = (40.72,) * 3
kg, kin = wl.gsphere_orthorhombic(20.0, L, (0, 0, 0), 2)
npw = kg.shape[0]; ndat = 64
v = torch.from_numpy(wl.smooth_potential(ng, 1)).cuda()
c = torch.randn((ndat, npw, 2), dtype=torch.float64, device="cuda"); c[:, 0, 1] = 0
out = torch.zeros_like(c)
for split in (0, 1):
    api.set_tuning("plane_split", split)
    run = lambda: api.fourwf(1, v, c, out, None, None, None, 2, kg, kg, max(ng), None, ndat, ng, npw, npw, ng[0], ng[1], ng[2], 2, impl=2)
    run(); torch.cuda.synchronize()
    api.profile_enable(True)
    for _ in range(3): run()
    prof = api.profile_collect(); api.profile_enable(False)
    ms = {k: round(t / 3, 3) for k, (t, n) in prof.items()}
    print(json.dumps({"ngfft": ng, "ndat": ndat, "plane_split": split, "ms": ms}), flush=True)
api.set_tuning("plane_split", 0)
