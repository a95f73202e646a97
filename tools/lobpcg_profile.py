import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import abinit_b200 as ab
from abinit_b200 import api, workload as wl, xg
ab.init(0)
dev = torch.device("cuda", 0)
cfg = wl.CONFIGS["si512"]
kg, kin = wl.gsphere_orthorhombic(cfg["ecut"], cfg["L"], (0, 0, 0), 2)
npw = kg.shape[0]; nband = 1100
indlmn, lnmax = wl.nc_indlmn(cfg["lmax"], cfg["nproj_per_l"]); nlmn = indlmn.shape[1]; natom = cfg["natom"]; nprojs = natom * nlmn
h = ab.Hamiltonian(cfg["ngfft"], natom, 1, nlmn, indlmn, np.array([natom], dtype=np.int32), np.arange(1, natom + 1, dtype=np.int32), 0, cfg["L"] ** 3)
h.load_spin(wl.smooth_potential(cfg["ngfft"], seed=5), 1); h.load_enl(np.random.default_rng(1).standard_normal((1, lnmax)), None)
h.load_k(2, kg, kin, None, None, me_g0=1)
gen = torch.Generator(device=dev).manual_seed(4321)
P = torch.randn((nprojs, npw, 2), generator=gen, device=dev, dtype=torch.float64) / np.sqrt(npw); P[:, 0, 1] = 0
cg = torch.randn((nband, npw, 2), generator=gen, device=dev, dtype=torch.float64) * torch.from_numpy(1 / (1 + kin)).to(dev)[None, :, None]; cg[:, 0, 1] = 0
torch.cuda.synchronize()
h.set_projectors(P, nprojs); del P; torch.cuda.empty_cache()
eig = np.zeros(nband); res = np.zeros(nband)
xg.lobpcgwf2(cg, eig, None, None, h, nband, npw, 1, res, 1e-30, 4, bandpp=128)
api.profile_enable(True)
t0 = time.perf_counter()
xg.lobpcgwf2(cg, eig, None, None, h, nband, npw, 1, res, 1e-30, 4, bandpp=128)
dt = time.perf_counter() - t0
prof = api.profile_collect()
print("total", dt)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]): print(f"  {k:24s} {v[0]:9.1f} ms  x{v[1]}")
