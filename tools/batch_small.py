#!/usr/bin/env python
"""Developer tool: Fe-2-like (k, spin) batch through getghc_batch vs one call after the other (launch-bound regime)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import abinit_b200 as ab
from abinit_b200 import api
from problems import make_problem

ab.init(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); api.set_stream(stream.cuda_stream)
ndat, nk = 24, 48
rng = np.random.default_rng(3)
hams, cws, ghcs, gscs = [], [], [], []
for ik in range(nk):
    k = tuple(rng.uniform(-0.5, 0.5, 3))
    p = make_problem(20.0, 5.42, k, 1, ndat=ndat, seed=10 + ik, natom_per_type=(2,), lmax_per_type=(2,), usepaw=1, ngfft=(30, 30, 30))
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 1, p.ucvol)
    h.load_spin(p.vlocal, 1); h.load_enl(p.enl, p.sij); h.load_k(1, p.kgF, p.kinpw, p.ffnl, p.ph3d)
    hams.append(h); cws.append(torch.from_numpy(p.cwavef).to(dev)); ghcs.append(torch.zeros_like(cws[-1])); gscs.append(torch.zeros_like(cws[-1]))
print("npw", p.npw, "nprojs", hams[0].nprojs, flush=True)
torch.cuda.synchronize()
api.set_async(True)
def plain():
    for h, c, g, s in zip(hams, cws, ghcs, gscs):
        ab.getghc(-1, c, None, g, s, h, None, None, None, ndat, sij_opt=1)
for name, fn in (("plain loop", plain), ("batch, lanes only", lambda: api.getghc_batch(hams, cws, ghcs, gscs, ndat=ndat, sij_opt=1, use_graphs=False)),
                 ("batch, lanes + graphs", lambda: api.getghc_batch(hams, cws, ghcs, gscs, ndat=ndat, sij_opt=1, use_graphs=True))):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 20
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {dt / n / nk * 1e6:.1f} us per (k,spin) call, {ndat * nk * n / dt:.0f} band-app/s", flush=True)
api.profile_enable(True)
for _ in range(5): plain()
prof = api.profile_collect(); api.profile_enable(False)
print("per call (us): " + ", ".join(f"{k} {v[0] / v[1] * 1e3:.1f}" for k, v in prof.items()), flush=True)
