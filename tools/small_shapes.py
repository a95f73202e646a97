#!/usr/bin/env python
"""Time getghc on the small BASELINE shapes (configs[0] Si-2 and configs[2] Fe-2-like, k-point sharded): launch-bound regime.
  python tools/small_shapes.py        (on the GPU box)"""
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
for name, ecut, L, kpt, istwf, ndat, natom, lmax, paw in (("si2 (24^3, npw~520, 8 bands, NC)", 12.0, 7.2, (-.25, .5, 0), 1, 8, (2,), (2,), 0),
                                                           ("fe2 (30^3, npw~700, 24 bands, PAW)", 20.0, 5.42, (.125, .25, .375), 1, 24, (2,), (2,), 1),
                                                           ("au-like (48^3, npw~7k, 64 bands, PAW)", 12.0, 14.0, (0, 0, 0), 2, 64, (16,), (2,), 1)):
    p = make_problem(ecut, L, kpt, istwf, ndat=ndat, natom_per_type=natom, lmax_per_type=lmax, usepaw=paw)
    h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, paw, p.ucvol)
    h.load_spin(p.vlocal, 1); h.load_enl(p.enl, p.sij); h.load_k(istwf, p.kgF, p.kinpw, p.ffnl, p.ph3d)
    cw = torch.from_numpy(p.cwavef).to(dev); ghc = torch.zeros_like(cw); gsc = torch.zeros_like(cw) if paw else None
    torch.cuda.synchronize()
    api.set_async(True)
    call = lambda: ab.getghc(-1, cw, None, ghc, gsc, h, None, None, None, ndat, sij_opt=1 if paw else 0)
    for _ in range(20): call()
    stream.synchronize()
    l0 = ab.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    n = 200
    for _ in range(n): call()
    e1.record(stream); stream.synchronize(); wall = time.perf_counter() - t0
    print(f"{name}: ngfft {p.ngfft} npw {p.npw} nprojs {h.nprojs}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/call device, {wall / n * 1e6:.1f} us/call host, "
          f"{(ab.kernel_launches() - l0) / n:.0f} launches/call, {ndat * n / wall:.0f} band-app/s", flush=True)
    api.profile_enable(True)
    for _ in range(50): call()
    prof = api.profile_collect(); api.profile_enable(False)
    print("    " + ", ".join(f"{k} {v[0] / 50 * 1e3:.1f} us" for k, v in prof.items()), flush=True)
    h.destroy()
