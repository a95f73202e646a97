#!/usr/bin/env python
"""bench.py -- getghc band-applications/s on synthetic wavefunctions/potentials of the Si-512 shape.

One "step" = one getghc call (type_calc=0: fourwf option 2 + gemm_nonlop choice 1 + kinetic assembly) on a block of
`ndat` bands at Gamma.  Weak scaling: every rank (one per GPU) applies H to its own band block with P, V_loc, kg and
kinpw replicated; there is no data-path collective in getghc.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload si512] [--ndat 128] [--istwfk 2]
  python bench.py --impl reference ...   # the reference algorithm (oracle port) on the host cores, bounded sample

Prints ONE JSON line (contract in the task statement): value = device-resident throughput; e2e = the same call with
HOST buffers (H2D/D2H inside the timed region); roofline = dominant kernel (DMMA GEMM of gemm_nonlop) against the
measured cuBLAS DGEMM peak of this pool's B200 (profiles/fp64_peak_r01.json -- MEASURED_PEAKS.json holds no FP64
figure), plus the fourwf HBM fraction as an extra; cpu_baseline = oracle port timed on the box's host cores.
"""
from __future__ import annotations
import argparse, json, os, subprocess, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="si512")
    ap.add_argument("--ndat", type=int, default=128)
    ap.add_argument("--istwfk", type=int, default=2)
    ap.add_argument("--cpu-bands", type=int, default=32, help="bands in the bounded CPU sample (one block: amortises the stream of P like bandpp does)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scf-step", action="store_true")
    ap.add_argument("--nonlop", default="fp64", choices=["fp64", "int8"],
                    help="gemm_nonlop arithmetic: fp64 = FP64 DMMA kernels (default, the product path); int8 = opt-in exact int8 slice "
                         "products on the tcgen05 kernel (DESIGN.md 3.5)")
    ap.add_argument("--nband", type=int, default=1100, help="bands of the ChebFi2 (SCF-step-equivalent) leg, sharded over the GPUs")
    ap.add_argument("--nline", type=int, default=4, help="Chebyshev filter degree of the ChebFi2 leg")
    return ap.parse_args()


def peaks():
    hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            hbm = float(json.load(open(p))["hbm_gbs"]); hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            pass
    fp64, fp64_src = 35.45, "profiles/fp64_peak_r01.json (cuBLAS DGEMM 8192^3 on this pool's B200, burst = 4 s sustained)"
    q = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    if os.path.exists(q):
        try:
            fp64 = float(json.load(open(q))["dgemm_8192"]["burst_tflops"])
        except Exception:
            pass
    return hbm, hbm_src, fp64, fp64_src


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (the clocks line of B200_PROFILING.md).  NVML is queried
    in-process every 10 ms (an `nvidia-smi -lms` child does not produce a row within a 0.2 s timed region); nvidia-smi is
    the fallback when pynvml is missing."""
    BITS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, gpu_index=0):
        self.idx = gpu_index; self.sm = []; self.mx = None; self.mask = 0; self.run = False; self.t = None; self.err = None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while self.run:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(get_reasons(h))
                except Exception:
                    pass
                time.sleep(0.01)
        except Exception as e:                                   # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True); self.t.start()

    def stop(self):
        self.run = False
        if self.t is not None:
            self.t.join(timeout=2)
        if not self.sm:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
                a, b = [float(x) for x in out.strip().split(",")[:2]]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "note": f"after the timed region ({self.err})"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable: " + str(self.err)], "samples": 0}
        reasons = sorted(k for k, b in self.BITS.items() if self.mask & b)
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(self.sm)}


def build_workload(args):
    from abinit_b200 import workload as wl
    cfg = wl.CONFIGS[args.workload]
    kg, kin = wl.gsphere_orthorhombic(cfg["ecut"], cfg["L"], (0.0, 0.0, 0.0), args.istwfk)
    npw = kg.shape[0]
    # outermost 0.5 % shell carries the huge*1e-10 sentinel (m_kg.F90:422-429) so the filter branch is live
    thr = np.quantile(kin, 0.995)
    kinpw = np.where(kin >= thr, wl.HUGE * 1e-10, kin)
    indlmn, lnmax = wl.nc_indlmn(cfg["lmax"], cfg["nproj_per_l"])
    nlmn = indlmn.shape[1]
    natom = cfg["natom"]
    w = dict(cfg=cfg, kg=kg, kinpw=np.ascontiguousarray(kinpw), kin_raw=kin, npw=npw, indlmn=indlmn, lnmax=lnmax, nlmn=nlmn,
             natom=natom, nprojs=natom * nlmn, ngfft=cfg["ngfft"], ucvol=float(cfg["L"]) ** 3,
             nattyp=np.array([natom], dtype=np.int32), atindx1=np.arange(1, natom + 1, dtype=np.int32),
             vlocal=wl.smooth_potential(cfg["ngfft"], seed=1234 + 1),
             ekb=np.ascontiguousarray(np.random.Generator(np.random.PCG64(1235)).standard_normal((1, lnmax))))
    return w


def algorithmic_units(w, istwfk, ndat):
    """bytes per band-application for fourwf and flops per band-application for gemm_nonlop."""
    n1, n2, n3 = w["ngfft"]
    kg = w["kg"]
    full = kg if istwfk == 1 else np.concatenate([kg, -kg])
    lines = np.unique(np.mod(full[:, 1], n2).astype(np.int64) * n3 + np.mod(full[:, 2], n3))
    C = int(lines.size)
    N = n1 * n2 * n3
    b_fw = 32.0 * w["npw"] + 64.0 * C * n1 + (8.0 * N + 20.0 * w["npw"]) / ndat
    g = 2
    f_nl = g * (8.0 if istwfk == 1 else 4.0) * w["npw"] * w["nprojs"]
    return b_fw, f_nl, C


def fourwf_flops_per_band(w, istwfk, C):
    """SURVEY 8d F_fw: pruned zero-padded 3-D FFT pair, nominal 5 n log2 n flops per 1-D transform (x on the C occupied lines,
    y on the occupied half of the z planes, z on every column) + the V_loc multiply; halved per band at Gamma with a real
    potential, where two bands ride one complex transform (cwavef_double_rfft_trick, m_getghc.F90:1999-2171)."""
    n1, n2, n3 = w["ngfft"]
    l2 = np.log2
    f = 2.0 * (C * 5.0 * n1 * l2(n1) + n1 * (n3 / 2.0) * 5.0 * n2 * l2(n2) + n1 * n2 * 5.0 * n3 * l2(n3)) + 2.0 * n1 * n2 * n3
    return f * (0.5 if istwfk == 2 else 1.0)


def run_reference(args):
    """The reference's CPU algorithm for the path (oracle port; the Fortran reference cannot be built here: no Fortran
    compiler), all host threads (OpenBLAS + pocketfft workers), each step a bounded sample of `cpu_bands` bands."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import getghc as ogh
    w = build_workload(args)
    nb = args.cpu_bands
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.time()
    gen = torch.Generator().manual_seed(4321)
    Pr = (torch.randn((w["nprojs"], w["npw"]), generator=gen, dtype=torch.float64) / np.sqrt(w["npw"])).numpy()
    Pi = (torch.randn((w["nprojs"], w["npw"]), generator=gen, dtype=torch.float64) / np.sqrt(w["npw"])).numpy()
    if args.istwfk == 2:
        Pi[:, 0] = 0.0

    class SplitP:       # P_r / P_i held separately like the reference does for istwf_k>1 (m_gemm_nonlop_projectors.F90)
        real = Pr; imag = Pi
    P = SplitP if args.istwfk >= 2 else (Pr + 1j * Pi)
    rng = np.random.Generator(np.random.PCG64(99))
    c = rng.standard_normal((nb, w["npw"])) + 1j * rng.standard_normal((nb, w["npw"]))
    if args.istwfk == 2:
        c[:, 0] = c[:, 0].real
    kg3 = np.ascontiguousarray(w["kg"].T)
    setup_s = time.time() - t0

    def step():
        ogh.getghc(c, w["vlocal"], kg3, w["ngfft"], w["kinpw"], P, w["ekb"], None, w["indlmn"], w["nattyp"],
                   w["atindx1"] - 1, istwf_k=args.istwfk, usepaw=0, workers=cores, local_impl="pad")
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t1 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t1
    val = nb * args.steps / dt
    out = {"metric": "getghc band-applications/s", "value": val, "unit": "band-applications/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None,
           "dtype": "f64" if args.nonlop == "fp64" else "f64 (fourwf) + exact int8 slices with int32 accumulation recombined in f64 (gemm_nonlop)",
           "data": "synthetic", "impl": "reference",
           "config": {"workload": f"{args.workload}: box {w['ngfft']}, npw {w['npw']}, nprojs {w['nprojs']}, istwfk {args.istwfk}",
                      "sample": f"{nb} bands per step"},
           "cpu_baseline": {"value": val, "unit": "band-applications/s", "cores": cores, "kind": "port",
                            "sample": f"{nb} bands x {args.steps} steps of the full-size operator (oracle NumPy/SciPy port: zero-padded pocketfft passes with Gamma-point band pairing + OpenBLAS GEMMs); set-up {setup_s:.0f} s untimed"},
           "e2e": {"value": val, "unit": "band-applications/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import abinit_b200 as ab
    from abinit_b200 import api
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ab.init(local)
    stream = torch.cuda.Stream(device=dev)
    api.set_stream(stream.cuda_stream)

    w = build_workload(args)
    ndat, npw, nprojs = args.ndat, w["npw"], w["nprojs"]
    ham = ab.Hamiltonian(w["ngfft"], w["natom"], 1, w["nlmn"], w["indlmn"], w["nattyp"], w["atindx1"], 0, w["ucvol"])
    ham.load_spin(w["vlocal"], 1)
    ham.load_enl(w["ekb"], None)
    ham.load_k(args.istwfk, w["kg"], w["kinpw"], None, None, me_g0=1)
    with torch.cuda.stream(stream):
        gen = torch.Generator(device=dev).manual_seed(4321 + rank)
        P = torch.randn((nprojs, npw, 2), generator=gen, device=dev, dtype=torch.float64) / np.sqrt(npw)
        if args.istwfk == 2:
            P[:, 0, 1] = 0.0
        cw = torch.randn((ndat, npw, 2), generator=gen, device=dev, dtype=torch.float64)
        if args.istwfk == 2:
            cw[:, 0, 1] = 0.0
        ghc = torch.zeros_like(cw)
    stream.synchronize()
    ham.set_projectors(P, nprojs)
    del P
    torch.cuda.empty_cache()

    def step_dev():
        ab.getghc(-1, cw, None, ghc, None, ham, None, None, None, ndat)

    def barrier():
        stream.synchronize(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    if args.nonlop == "int8":
        api.set_tuning("nonlop_ozaki", 1)
    api.set_async(True)
    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = ab.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ab.kernel_launches() - l0
    clocks = sampler.stop()
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * ndat * args.steps / (ms * 1e-3)

    # per-kernel-class device times (second pass, same work) for the roofline object
    api.profile_enable(True)
    for _ in range(args.steps):
        step_dev()
    prof = api.profile_collect()
    api.profile_enable(False)

    # end-to-end: host (pinned) buffers through the same public call
    e2e = None
    if not args.no_e2e:
        api.set_async(False)
        h_c = torch.empty((ndat, npw, 2), dtype=torch.float64).pin_memory(); h_c.copy_(cw.cpu())
        h_g = torch.empty((ndat, npw, 2), dtype=torch.float64).pin_memory()
        hc, hg = h_c.numpy(), h_g.numpy()
        for _ in range(2):
            ab.getghc(-1, hc, None, hg, None, ham, None, None, None, ndat)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ab.getghc(-1, hc, None, hg, None, ham, None, None, None, ndat)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * ndat * args.steps / float(dt.item()), "unit": "band-applications/s",
               "h2d_bytes_per_step": int(ndat * npw * 16), "d2h_bytes_per_step": int(ndat * npw * 16)}

    # SCF-step-equivalent (BASELINE metric "s/SCF step"): ONE ChebFi2 call on nband bands at this k-point, band-sharded over
    # the GPUs: (ndeg+1) getghc passes + Rayleigh-Ritz (all-to-all re-layout, Gram allreduce over NCCL, hegvd, rotations)
    scf_step = None
    if not args.no_scf_step:
        from abinit_b200 import parallel as par
        api.set_async(False)
        f, l = par.band_block(args.nband, world, rank)
        with torch.cuda.stream(stream):
            gen = torch.Generator(device=dev).manual_seed(777 + rank)
            damp = torch.from_numpy(1.0 / (1.0 + np.minimum(w["kinpw"], 1e6))).to(dev)
            cg = torch.randn((l - f, npw, 2), generator=gen, device=dev, dtype=torch.float64) * damp[None, :, None]
            if args.istwfk == 2:
                cg[:, 0, 1] = 0.0
            times = []
            for it in range(2):                                   # first call warms up (plans, cuSOLVER handle, workspaces)
                barrier()
                t0 = time.perf_counter()
                l1 = ab.kernel_launches()
                eig, res = par.chebfi_band_parallel(ham, cg, args.nband, float(w["cfg"]["ecut"]), args.nline, bandpp=ndat)
                barrier()
                times.append(time.perf_counter() - t0)
            dtm = torch.tensor([times[-1]], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(dtm, op=dist.ReduceOp.MAX)
        api.profile_enable(True)
        with torch.cuda.stream(stream):
            par.chebfi_band_parallel(ham, cg, args.nband, float(w["cfg"]["ecut"]), args.nline, bandpp=ndat)
        prof_scf = api.profile_collect()
        api.profile_enable(False)
        scf_step = {"value": float(dtm.item()), "unit": "s per ChebFi2 call (one k-point, SCF-step-equivalent)", "nband": args.nband,
                    "nline": args.nline, "bands_per_gpu": l - f, "launches": int(ab.kernel_launches() - l1),
                    "eig_min_max": [float(np.min(eig)), float(np.max(eig))], "resid_max": float(np.max(res)),
                    "timing": "host clock between barrier + device synchronise on both sides, max over ranks (the call holds host syncs)",
                    "kernel_ms": {k: v[0] for k, v in prof_scf.items()}}
        del cg
    # the same SCF-step-equivalent with LOBPCG (one block of all bands, nline LOBPCG iterations; single GPU only in this build)
    lobpcg_step = None
    if world == 1 and not args.no_scf_step:
        from abinit_b200 import xg as xgm
        api.set_async(False)
        with torch.cuda.stream(stream):
            gen = torch.Generator(device=dev).manual_seed(778)
            damp = torch.from_numpy(1.0 / (1.0 + np.minimum(w["kinpw"], 1e6))).to(dev)
            cgl = torch.randn((args.nband, npw, 2), generator=gen, device=dev, dtype=torch.float64) * damp[None, :, None]
            if args.istwfk == 2:
                cgl[:, 0, 1] = 0.0
            eigl = np.zeros(args.nband); resl = np.zeros(args.nband)
            tl = []
            for it in range(2):
                barrier(); t0 = time.perf_counter()
                xgm.lobpcgwf2(cgl, eigl, None, None, ham, args.nband, npw, 1, resl, 1e-30, args.nline, bandpp=ndat)
                barrier(); tl.append(time.perf_counter() - t0)
        lobpcg_step = {"value": tl[-1], "unit": "s per LOBPCG call (one k-point, one block of all bands)", "nband": args.nband,
                       "nline": args.nline, "eig_min_max": [float(eigl.min()), float(eigl.max())], "resid_max": float(resl.max())}
        del cgl
        torch.cuda.empty_cache()
    # density build (the step after the solver): fourwf option 1 on the same band block, fused path
    density = None
    if not args.no_scf_step:
        api.set_async(True)
        with torch.cuda.stream(stream):
            rho = torch.zeros(tuple(reversed(w["ngfft"])), device=dev, dtype=torch.float64)
            wts = np.full(ndat, 2.0 / w["ucvol"])
            n1, n2, n3 = w["ngfft"]
            def dens():
                api.fourwf(1, rho, cw, None, None, None, None, args.istwfk, w["kg"], w["kg"], max(w["ngfft"]), None, ndat, w["ngfft"],
                           npw, npw, n1, n2, n3, 1, weight_array_r=wts, weight_array_i=wts)
            for _ in range(3):
                dens()
            barrier()
            d0 = torch.cuda.Event(enable_timing=True); d1 = torch.cuda.Event(enable_timing=True)
            d0.record(stream)
            for _ in range(5):
                dens()
            d1.record(stream)
            barrier()
            dms = d0.elapsed_time(d1) / 5
        density = {"ms_per_block": dms, "bands": ndat, "bands_per_s": ndat / (dms * 1e-3), "kernel": "fourwf option 1, fused (x pass, plane stage with density reduction, transpose-add)"}
        del rho
    # sanity anchor against the only fourwf timing the reference stores: option 2, cplex 1, istwfk 1,
    # box 100^3 (ecut 30 Ha, 20 Bohr cube, k = (.1,.2,.3)): 3.8 ms per call with FFTW3 on one CPU core
    # (tests/unitary/Refs/tfourwf_01.stdout:117-129)
    anchor = None                                     # rank 0 only: NO collective in this block (local synchronisation)
    if rank == 0 and not args.no_scf_step:
        from abinit_b200 import workload as wl2
        kg_a, _ = wl2.gsphere_orthorhombic(30.0, 20.0, (0.1, 0.2, 0.3), 1)
        npw_a = kg_a.shape[0]
        api.set_async(True)
        with torch.cuda.stream(stream):
            va = torch.from_numpy(wl2.smooth_potential((100, 100, 100), seed=3)).to(dev)
            res_a = {}
            for nd_a in (1, 64):
                ca = torch.randn((nd_a, npw_a, 2), device=dev, dtype=torch.float64); oa = torch.zeros_like(ca)
                def fw():
                    api.fourwf(1, va, ca, oa, None, None, None, 1, kg_a, kg_a, 100, None, nd_a, (100, 100, 100), npw_a, npw_a,
                               100, 100, 100, 2)
                for _ in range(3):
                    fw()
                stream.synchronize()
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for _ in range(10):
                    fw()
                a1.record(stream)
                stream.synchronize()
                res_a[nd_a] = a0.elapsed_time(a1) / 10 / nd_a
        anchor = {"case": "tfourwf_01: fourwf option 2, box 100^3, npw %d, istwfk 1, device-resident" % npw_a,
                  "ms_per_band_ndat1": res_a[1], "ms_per_band_ndat64": res_a[64],
                  "reference_ms_per_band": 18.8, "reference_ms_per_band_goedecker112": 30.2,
                  "reference_source": "tests/unitary/Refs/tfourwf_01.stdout:117-129 (CPU-time per call as printed, m_fft_prof.F90:580: FFTW3 fftalg 312 / Goedecker fftalg 112, 1 CPU core)"}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    hbm, hbm_src, fp64, fp64_src = peaks()
    b_fw, f_nl, C = algorithmic_units(w, args.istwfk, ndat)
    def per_launch(name):
        t, c = prof.get(name, (0.0, 0))
        return (t / c) if c else None
    t_tn, t_nn = per_launch("dgemm_tn_opernla"), per_launch("dgemm_nn_opernlb")
    t_fw = sum((per_launch(k) or 0.0) * (prof.get(k, (0, 0))[1] / max(1, args.steps)) for k in
               ("fourwf_x_forward", "fourwf_plane_stage", "fourwf_x_backward"))
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
    except Exception:
        pass
    roof = None
    if t_nn:
        flops_launch = 0.5 * f_nl * ndat                    # one of the two GEMMs of a getghc step
        ach = flops_launch / (t_nn * 1e-3) / 1e12
        roof = {"kernel": "k_dgemm_nn (opernlb: vect = P . gxfac, DMMA m8n8k4)", "bound": "tensor", "achieved": ach, "peak": fp64,
                "unit": "TFLOP/s", "frac": ach / fp64, "traffic": (traffic.get("k_dgemm_nn") or {}).get("bytes"),
                "traffic_source": (traffic.get("k_dgemm_nn") or {}).get("source"), "peak_source": fp64_src,
                "flops_per_launch": flops_launch, "ms_per_launch": t_nn}
    extra = {}
    t_i8, n_i8 = prof.get("ozaki_igemm", (0.0, 0))
    if n_i8:
        # opt-in int8-sliced path: the dominant kernel is k_igemm_tc (tcgen05 kind::i8); algorithmic int8 ops of the 28 slice
        # products of both contractions over the summed launch time of a step
        bf16 = None
        try:
            bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            pass
        peak_i8 = 2.0 * bf16 if bf16 else 4500.0
        ops_step = 28.0 * f_nl * ndat                     # 28 slice products, each with the op count of its FP64 GEMM (f_nl covers both contractions)
        ach = ops_step / (t_i8 / args.steps * 1e-3) / 1e12
        roof = {"kernel": "k_igemm_tc (tcgen05.mma kind::i8, int32 accumulators in TMEM; 14 launches per step)", "bound": "tensor",
                "achieved": ach, "peak": peak_i8, "unit": "TOP/s (int8)", "frac": ach / peak_i8,
                "traffic": None, "peak_source": "2 x bf16_tflops of MEASURED_PEAKS.json (int8 dense = 2 x bf16 dense)" if bf16 else "nominal 4.5 Pop/s",
                "ops_per_step": ops_step, "ms_per_step": t_i8 / args.steps}
    if t_tn:
        ach = 0.5 * f_nl * ndat / (t_tn * 1e-3) / 1e12
        extra["roofline_opernla"] = {"kernel": "k_dgemm_tn (split-K P^T psi)", "bound": "tensor", "achieved": ach, "peak": fp64,
                                     "unit": "TFLOP/s", "frac": ach / fp64, "ms_per_launch": t_tn}
    if t_fw:
        ach = b_fw * ndat / (t_fw * 1e-3) / 1e9
        extra["roofline_fourwf"] = {"kernel": "fourwf option 2 (3 fused kernels)", "bound": "hbm", "achieved": ach, "peak": hbm,
                                    "unit": "GB/s", "frac": ach / hbm, "peak_source": hbm_src, "bytes_per_band": b_fw,
                                    "ms_per_step": t_fw, "lines_C": C}
        # fourwf in FP64 sits above the FP64 ridge of the chip (AI ~ 6-11 flop/B vs 5.4): the FP64 pipe is its other bound
        f_fw = fourwf_flops_per_band(w, args.istwfk, C)
        extra["roofline_fourwf"]["fp64"] = {"flops_per_band": f_fw, "achieved": f_fw * ndat / (t_fw * 1e-3) / 1e12, "peak": fp64,
                                            "unit": "TFLOP/s", "frac": f_fw * ndat / (t_fw * 1e-3) / 1e12 / fp64,
                                            "note": "nominal 5 n log2 n flops; DADD/DMUL issue at the DFMA rate, so the pipe is busier than this fraction"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                   "--workload", args.workload, "--istwfk", str(args.istwfk), "--cpu-bands", str(args.cpu_bands)]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:   # the baseline is reported, never a gate
            cpu = {"value": None, "unit": "band-applications/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    # EXPERIMENTAL (reported separately, never part of `value`): gemm_nonlop's two contractions through exact int8 slice
    # products (csrc/ozaki.cu: own slicing + FP64 recombination around cuBLASLt int8 GEMMs) instead of the FP64 DMMA kernels.
    # Runs in its own process after this one has released the device, so that nothing it does can touch the numbers above.
    experimental = None
    if world == 1 and not args.no_scf_step and args.nonlop == "fp64":
        try:
            ham.destroy(); ab.finalize(); del cw, ghc
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ozaki_bench.py"), "--json"], capture_output=True, text=True, timeout=600)
            experimental = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as ex:                                   # noqa: BLE001
            experimental = {"error": repr(ex)}
    out = {"metric": "getghc band-applications/s", "value": value, "unit": "band-applications/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None,
           "dtype": "f64" if args.nonlop == "fp64" else "f64 (fourwf) + exact int8 slices with int32 accumulation recombined in f64 (gemm_nonlop)",
           "data": "synthetic",
           "config": {"workload": f"{args.workload} (BASELINE configs[1] shape): FFT box {w['ngfft']}, Gamma, istwfk {args.istwfk}, "
                                  f"npw {npw}, nprojs {nprojs}, band block {ndat} per GPU, NC (paw_opt 0), type_calc 0, gemm_nonlop arithmetic {args.nonlop}",
                      "l2": "inputs larger than L2 (P = %.1f GB streamed twice per step)" % (16.0 * npw * nprojs / 1e9),
                      "parallelism": f"band blocks over {world} GPU(s), no data-path collective"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
           "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()}, "scf_step": scf_step, "lobpcg_step": lobpcg_step, "density_step": density, "fourwf_anchor": anchor,
           "experimental_int8_sliced": experimental}
    out.update(extra)
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
