#!/usr/bin/env python
"""bench.py -- getghc band-applications/s on synthetic wavefunctions/potentials of the BASELINE.json shapes.

One "step" = one pass of the hot path over one batch of synthetic input:
  --workload si512 (default, BASELINE configs[1], the configuration the metric is quoted on): ONE getghc call (fourwf option 2 +
             gemm_nonlop choice 1 + kinetic assembly) on a block of --ndat bands at Gamma, NC.  Weak scaling: every rank (one per
             GPU) applies H to its own band block with P, V_loc, kg, kinpw replicated; no data-path collective.
  --workload au108 (configs[3]): the same on the Au-108 PAW shape (box 96^3, istwf_k 1, getghc with gsc: three GEMMs + packed
             D_ij / S_ij), --au-lmax 2 (18 projectors per atom) or 3 (32, semicore-like).
  --workload fe2   (configs[2]): bcc Fe-2 spin-polarised PAW, the irreducible wedge of a 12x12x12 mesh (84 k-points) x 2 spins,
             24 bands each, (k, spin) pairs dealt round-robin to the ranks (m_vtorho.F90:855-862), one step = getghc_batch over the
             rank's pairs (concurrent lanes + CUDA graphs).  Strong scaling.
  --workload sweep (configs[4]): the SURVEY 8d grid (boxes 48^3-192^3 x band blocks 64-1024, an nprojs axis, istwfk 1 rows); one
             JSON line per point into --out, one summary line on stdout.
  python bench.py --impl reference ...   # the reference's CPU algorithm (C++/OpenMP restatement, oracle/cref) on the host cores

Prints ONE JSON line (contract in the task statement): value = device-resident throughput; e2e = the same call with HOST buffers
(H2D/D2H inside the timed region); roofline = dominant kernel against the FP64 peak MEASURED IN THIS RUN (cuBLAS DGEMM through
torch and the library's DFMA / DMMA pipe probe; the denominator is the largest of them), roofline_fourwf against the HBM peak of
MEASURED_PEAKS.json; parity = the timed configuration compared with the CPU restatement on 4 bands; cpu_baseline = the same
restatement timed on the box's host cores.
"""
from __future__ import annotations
import os
import sys


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# The CPU legs use every host core: OpenBLAS and libgomp read their thread counts from the environment when they are first
# loaded, and a launcher (torch.distributed.run) exports OMP_NUM_THREADS=1 -- fix the environment BEFORE NumPy is imported.
if "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(host_cores())

import argparse, json, subprocess, threading, time   # noqa: E402
import numpy as np                                  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "getghc band-applications/s"
UNIT = "band-applications/s"


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="si512")
    ap.add_argument("--ndat", type=int, default=None, help="band block per GPU (default 128; fe2: 24 bands per (k, spin))")
    ap.add_argument("--istwfk", type=int, default=None, help="default: 2 for si512 / sweep, 1 for au108 / fe2")
    ap.add_argument("--au-lmax", type=int, default=2, help="au108: 2 = 18 projectors per atom (nprojs 1944), 3 = 32 (nprojs 3456)")
    ap.add_argument("--scf-driver", choices=("native", "python"), default="native", help="band-parallel ChebFi2 of the scf_step leg: the library's own NCCL driver or the torch.distributed one")
    ap.add_argument("--cpu-bands", type=int, default=32, help="bands in the bounded CPU sample (one block: amortises the stream of P like bandpp does)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-scf-step", action="store_true")
    ap.add_argument("--nonlop", default="fp64", choices=["fp64", "int8"],
                    help="gemm_nonlop arithmetic: fp64 = FP64 DMMA kernels (default, the product path); int8 = opt-in exact int8 slice "
                         "products on the tcgen05 kernel (DESIGN.md 3.5)")
    ap.add_argument("--nband", type=int, default=1100, help="bands of the ChebFi2 (SCF-step-equivalent) leg, sharded over the GPUs")
    ap.add_argument("--nline", type=int, default=4, help="Chebyshev filter degree of the ChebFi2 leg")
    ap.add_argument("--out", default=None, help="sweep: file receiving one JSON line per point")
    ap.add_argument("--sweep-quick", action="store_true", help="sweep: boxes up to 128^3 and blocks 64 / 256 only")
    a = ap.parse_args(argv)
    if a.istwfk is None:
        a.istwfk = 1 if a.workload in ("au108", "fe2") else 2
    if a.ndat is None:
        a.ndat = 24 if a.workload == "fe2" else 128
    return a


# ------------------------------------------------------------------------------------------------------------------------------
# peaks
# ------------------------------------------------------------------------------------------------------------------------------
def hbm_peak():
    hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            hbm = float(json.load(open(p))["hbm_gbs"]); src = "MEASURED_PEAKS.json"
        except Exception:
            pass
    return hbm, src


_FP64 = None


def measure_fp64_peak(dev=None):
    """FP64 peak of THIS device in THIS process: cuBLAS DGEMM 8192^3 through torch (best of 5) and the library's register-resident
    DFMA / DMMA m8n8k4 probes.  The roofline denominator is the largest of the three, so that frac <= 1 by construction."""
    global _FP64
    if _FP64 is not None:
        return _FP64
    import torch
    from abinit_b200 import api
    out = {}
    try:
        n = 8192
        a = torch.randn((n, n), device=dev, dtype=torch.float64); b = torch.randn((n, n), device=dev, dtype=torch.float64)
        c = torch.empty_like(a)
        torch.matmul(a, b, out=c); torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["cublas_dgemm_8192"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        del a, b, c
        torch.cuda.empty_cache()
    except Exception as ex:                                       # noqa: BLE001
        out["cublas_dgemm_8192"] = None; out["cublas_error"] = repr(ex)
    pr = api.probe_fp64_peak()
    out["dfma_probe"] = pr["dfma"]; out["dmma_probe"] = pr["dmma"]
    vals = [v for v in (out.get("cublas_dgemm_8192"), out["dfma_probe"], out["dmma_probe"]) if v]
    out["peak"] = max(vals)
    out["source"] = "measured in this run on this device: max of cuBLAS DGEMM 8192^3 (torch.matmul), DFMA and DMMA m8n8k4 pipe probes (abi_b200_probe_fp64_peak)"
    _FP64 = out
    return out


def peaks():
    """(hbm GB/s, source, fp64 TFLOP/s, source) -- kept for tools/*.py; the FP64 figure is measured when a device is present."""
    hbm, hsrc = hbm_peak()
    try:
        f = measure_fp64_peak()
        return hbm, hsrc, f["peak"], f["source"]
    except Exception:
        return hbm, hsrc, 37.0, "fallback: pipe probe of round 1 (profiles/fp64_pipe_probe_r01.json)"


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


# ------------------------------------------------------------------------------------------------------------------------------
# workloads (host side)
# ------------------------------------------------------------------------------------------------------------------------------
def build_workload(args, usepaw=0):
    """Host arrays of one k-point of a named shape (abinit_b200.workload.CONFIGS): sphere, kinetic energies with a sentinel shell,
    potential, NC ekb or PAW D_ij / S_ij."""
    from abinit_b200 import workload as wl
    cfg = dict(wl.CONFIGS[args.workload])
    if args.workload == "au108":
        cfg["lmax"] = args.au_lmax
    kg, kin = wl.gsphere_orthorhombic(cfg["ecut"], cfg["L"], (0.0, 0.0, 0.0), args.istwfk)
    npw = kg.shape[0]
    # outermost 0.5 % shell carries the huge*1e-10 sentinel (m_kg.F90:422-429) so the filter branch is live
    thr = np.quantile(kin, 0.995)
    kinpw = np.where(kin >= thr, wl.HUGE * 1e-10, kin)
    indlmn, lnmax = wl.nc_indlmn(cfg["lmax"], cfg["nproj_per_l"])
    nlmn = indlmn.shape[1]
    natom = cfg["natom"]
    rng = np.random.Generator(np.random.PCG64(1235))
    w = dict(cfg=cfg, name=args.workload, kg=kg, kinpw=np.ascontiguousarray(kinpw), kin_raw=kin, npw=npw, indlmn=indlmn, lnmax=lnmax,
             nlmn=nlmn, natom=natom, nprojs=natom * nlmn, ngfft=cfg["ngfft"], ucvol=float(cfg["L"]) ** 3, usepaw=usepaw,
             nattyp=np.array([natom], dtype=np.int32), atindx1=np.arange(1, natom + 1, dtype=np.int32),
             vlocal=wl.smooth_potential(cfg["ngfft"], seed=1234 + 1),
             ekb=np.ascontiguousarray(rng.standard_normal((1, lnmax))))
    if usepaw:
        lmn2 = nlmn * (nlmn + 1) // 2
        w["dij"] = np.ascontiguousarray(0.3 * rng.standard_normal((natom, lmn2)))
        a = 0.1 * rng.standard_normal((nlmn, nlmn)); a = a @ a.T
        w["sij"] = np.ascontiguousarray(np.array([[a[i, j] for j in range(nlmn) for i in range(j + 1)]]))
    return w


def algorithmic_units(w, istwfk, ndat, g=2):
    """bytes per band-application for fourwf and flops per band-application for gemm_nonlop (g GEMMs: 2 NC, 3 PAW + gsc)."""
    n1, n2, n3 = w["ngfft"]
    kg = w["kg"]
    full = kg if istwfk == 1 else np.concatenate([kg, -kg])
    lines = np.unique(np.mod(full[:, 1], n2).astype(np.int64) * n3 + np.mod(full[:, 2], n3))
    C = int(lines.size)
    N = n1 * n2 * n3
    b_fw = 32.0 * w["npw"] + 64.0 * C * n1 + (8.0 * N + 20.0 * w["npw"]) / ndat
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


def cref_operator(w, istwfk, Pr, Pi):
    """The CPU restatement (oracle/cref) of this workload's operator.  Checker / CPU baseline only."""
    from oracle import cref
    cref.set_threads(host_cores())
    if w["usepaw"]:
        blk = np.full(w["natom"], w["nlmn"], dtype=np.int32)
        return cref.Operator(w["vlocal"], w["kg"], w["ngfft"], w["kinpw"], Pr, Pi, istwfk, blk_nlmn=blk, dij=w["dij"],
                             sij=np.repeat(w["sij"], w["natom"], axis=0))
    iln = w["indlmn"][0, :w["nlmn"], 4].astype(int) - 1
    ekb_proj = np.tile(w["ekb"][0, iln], w["natom"])
    return cref.Operator(w["vlocal"], w["kg"], w["ngfft"], w["kinpw"], Pr, Pi, istwfk, ekb_proj=ekb_proj)


def time_cref(op, c, sij_opt, steps, warmup=1):
    for _ in range(warmup):
        op.getghc(c, sij_opt=sij_opt)
    t0 = time.time()
    for _ in range(steps):
        op.getghc(c, sij_opt=sij_opt)
    return (time.time() - t0) / steps


CPU_SAMPLE_TEXT = ("{nb} bands x {steps} steps of the full-size operator (C++/OpenMP restatement oracle/cref: zero-padded Stockham FFT passes "
                   "vectorised over line batches, V(r) applied between the z transforms, Gamma-point band pairing, OpenBLAS DGEMMs on "
                   "P_r / P_i; {cores} threads)")


def run_reference(args):
    """The reference's CPU algorithm for the path on all host threads, each step a bounded sample of `cpu_bands` bands.  The Fortran
    reference cannot be built here (no Fortran compiler); oracle/cref is its C++/OpenMP restatement (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = host_cores()
    torch.set_num_threads(cores)
    if args.workload == "fe2":
        return run_reference_fe2(args, cores)
    usepaw = 1 if args.workload == "au108" else 0
    w = build_workload(args, usepaw)
    nb = args.cpu_bands
    t0 = time.time()
    gen = torch.Generator().manual_seed(4321)
    Pr = (torch.randn((w["nprojs"], w["npw"]), generator=gen, dtype=torch.float64) / np.sqrt(w["npw"])).numpy()
    Pi = (torch.randn((w["nprojs"], w["npw"]), generator=gen, dtype=torch.float64) / np.sqrt(w["npw"])).numpy()
    if args.istwfk == 2:
        Pi[:, 0] = 0.0
    rng = np.random.Generator(np.random.PCG64(99))
    c = rng.standard_normal((nb, w["npw"])) + 1j * rng.standard_normal((nb, w["npw"]))
    if args.istwfk == 2:
        c[:, 0] = c[:, 0].real
    op = cref_operator(w, args.istwfk, Pr, Pi)
    setup_s = time.time() - t0
    sij_opt = 1 if usepaw else 0
    dt = time_cref(op, c, sij_opt, args.steps, warmup=max(1, min(args.warmup, 1)))
    val = nb / dt
    cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": CPU_SAMPLE_TEXT.format(nb=nb, steps=args.steps, cores=cores) + f"; set-up {setup_s:.0f} s untimed",
           "seconds_by_part": dict(zip(("fourwf", "opernla", "opernlc", "opernlb_assembly"), [float(x) for x in op.timings]))}
    out = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "impl": "reference", "config": {"workload": workload_text(args, w, nb), "sample": f"{nb} bands per step on {cores} host threads"},
           "cpu_baseline": cpu, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


def workload_text(args, w, ndat):
    tag = {"si512": "BASELINE configs[1] shape", "au108": "BASELINE configs[3] shape"}.get(args.workload, "BASELINE configs[4] sweep point")
    kind = "PAW with S (paw_opt 4: ghc and gsc)" if w["usepaw"] else "NC (paw_opt 0)"
    return (f"{args.workload} ({tag}): FFT box {tuple(w['ngfft'])}, Gamma, istwfk {args.istwfk}, npw {w['npw']}, nprojs {w['nprojs']}, "
            f"band block {ndat} per GPU, {kind}, type_calc 0, gemm_nonlop arithmetic {args.nonlop}")


# ------------------------------------------------------------------------------------------------------------------------------
# fe2: many small (k, spin) Hamiltonians
# ------------------------------------------------------------------------------------------------------------------------------
FE2 = dict(ecut=20.0, L=5.42, ngfft=(24, 24, 24), natom=2, lmax=2, nproj_per_l=2, nband=24, mesh=12)


def fe2_kpoints():
    """Irreducible wedge (i >= j >= l) of the Gamma-centred 12x12x12 mesh of the cubic cell: 84 k-points."""
    m = FE2["mesh"]
    return [(i / m, j / m, l / m) for i in range(m // 2 + 1) for j in range(i + 1) for l in range(j + 1)]


def fe2_problem(ik, isppol, kpt, ndat, with_host_p=False):
    """Host arrays of one (k, spin) pair: sphere of k, the spin's potential and D_ij, random projectors of k."""
    from abinit_b200 import workload as wl
    kg, kin = wl.gsphere_orthorhombic(FE2["ecut"], FE2["L"], kpt, 1)
    npw = kg.shape[0]
    indlmn, lnmax = wl.nc_indlmn(FE2["lmax"], FE2["nproj_per_l"])
    nlmn = indlmn.shape[1]; natom = FE2["natom"]; lmn2 = nlmn * (nlmn + 1) // 2
    rs = np.random.Generator(np.random.PCG64(500 + isppol))      # per spin
    dij = np.ascontiguousarray(0.3 * rs.standard_normal((natom, lmn2)))
    r0 = np.random.Generator(np.random.PCG64(499))               # shared
    a = 0.1 * r0.standard_normal((nlmn, nlmn)); a = a @ a.T
    sij = np.ascontiguousarray(np.array([[a[i, j] for j in range(nlmn) for i in range(j + 1)]]))
    rk = np.random.Generator(np.random.PCG64(1000 + ik))         # per k
    P = rk.standard_normal((natom * nlmn, npw, 2)) / np.sqrt(npw)
    rc = np.random.Generator(np.random.PCG64(5000 + 2 * ik + isppol))
    c = rc.standard_normal((ndat, npw, 2)) / (1.0 + kin)[None, :, None]
    return dict(kg=kg, kinpw=np.ascontiguousarray(kin), npw=npw, indlmn=indlmn, nlmn=nlmn, natom=natom, nprojs=natom * nlmn,
                ngfft=FE2["ngfft"], ucvol=FE2["L"] ** 3, usepaw=1, nattyp=np.array([natom], dtype=np.int32),
                atindx1=np.arange(1, natom + 1, dtype=np.int32), vlocal=wl.smooth_potential(FE2["ngfft"], seed=40 + isppol),
                dij=dij, sij=sij, P=np.ascontiguousarray(P), c=np.ascontiguousarray(c))


def fe2_text(npairs, ndat, world):
    return (f"fe2 (BASELINE configs[2] shape): bcc Fe-2 spin-polarised PAW with S, box {FE2['ngfft']}, ecut {FE2['ecut']} Ha, 84 k-points "
            f"(wedge of a 12x12x12 mesh) x 2 spins = {npairs} (k, spin) pairs, {ndat} bands each, istwfk 1, nprojs 36, getghc with gsc, "
            f"pairs dealt round-robin to {world} GPU(s)")


def run_reference_fe2(args, cores):
    from oracle import cref
    cref.set_threads(cores)
    kpts = fe2_kpoints()
    pairs = [(ik, isp) for isp in range(2) for ik in range(len(kpts))]
    sample = pairs[::7]                                           # bounded sample: every 7th pair (24 of 168)
    ops = []
    for ik, isp in sample:
        q = fe2_problem(ik, isp, kpts[ik], args.ndat)
        op = cref_operator(q, 1, q["P"][..., 0], q["P"][..., 1])
        ops.append((op, q["c"][..., 0] + 1j * q["c"][..., 1]))
    for op, c in ops[:2]:
        op.getghc(c, sij_opt=1)
    t0 = time.time()
    for _ in range(args.steps):
        for op, c in ops:
            op.getghc(c, sij_opt=1)
    dt = (time.time() - t0) / args.steps
    val = len(ops) * args.ndat / dt
    cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{len(ops)} of the {len(pairs)} (k, spin) pairs x {args.ndat} bands per step (oracle/cref C++/OpenMP restatement, {cores} threads)"}
    out = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "impl": "reference", "config": {"workload": fe2_text(len(pairs), args.ndat, args.gpus), "sample": cpu["sample"]},
           "cpu_baseline": cpu, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


def run_fe2(args, env):
    import torch
    import abinit_b200 as ab
    from abinit_b200 import api, parallel as par
    rank, world, dev, stream, dist, barrier = env["rank"], env["world"], env["dev"], env["stream"], env["dist"], env["barrier"]
    ndat = args.ndat
    kpts = fe2_kpoints()
    nk = len(kpts)
    mine = par.my_kpoints(nk, 2, world, rank)
    npairs = 2 * nk
    hams, cws, ghcs, gscs, probs = [], [], [], [], []
    for ik, isp in mine:
        q = fe2_problem(ik, isp, kpts[ik], ndat)
        h = ab.Hamiltonian(q["ngfft"], q["natom"], 1, q["nlmn"], q["indlmn"], q["nattyp"], q["atindx1"], 1, q["ucvol"])
        h.load_spin(q["vlocal"], 1); h.load_enl(q["dij"], q["sij"]); h.load_k(1, q["kg"], q["kinpw"], None, None, me_g0=1)
        h.set_projectors(torch.from_numpy(q["P"]).to(dev), q["nprojs"])
        hams.append(h); probs.append(q)
    # the wavefunction blocks of all (k, spin) pairs of this rank live in ONE array with per-pair offsets, like the reference's
    # cg(2, mcg) with its icg offsets (src/79_seqpar_mpi/m_vtorho.F90:1035-1045): the blocks handed to getghc are views
    sizes = [int(q["c"].size) for q in probs]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cw_all = torch.empty(int(offs[-1]), dtype=torch.float64, device=dev); ghc_all = torch.zeros_like(cw_all); gsc_all = torch.zeros_like(cw_all)
    for i, q in enumerate(probs):
        cw_all[offs[i]:offs[i + 1]] = torch.from_numpy(q["c"].reshape(-1)).to(dev)
        shp = q["c"].shape
        cws.append(cw_all[offs[i]:offs[i + 1]].view(shp)); ghcs.append(ghc_all[offs[i]:offs[i + 1]].view(shp)); gscs.append(gsc_all[offs[i]:offs[i + 1]].view(shp))
    torch.cuda.synchronize()
    api.set_async(True)

    def step_dev():
        api.getghc_batch(hams, cws, ghcs, gscs, ndat=ndat, sij_opt=1, use_graphs=True)

    def step_plain():
        for h, c, g, s in zip(hams, cws, ghcs, gscs):
            ab.getghc(-1, c, None, g, s, h, None, None, None, ndat, sij_opt=1)

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    sampler = ClockSampler(env["local"]); sampler.start()
    l0 = ab.kernel_launches()
    ms = timed(step_dev, args.steps)
    launches = ab.kernel_launches() - l0
    clocks = sampler.stop()
    value = npairs * ndat * args.steps / (ms * 1e-3)
    for _ in range(3):
        step_plain()
    ms_plain = timed(step_plain, args.steps)
    # per-kernel classes from the plain loop (the graph replays are opaque to the in-library timers)
    api.profile_enable(True)
    for _ in range(args.steps):
        step_plain()
    prof = api.profile_collect(); api.profile_enable(False)
    # parity: every 6th pair of this rank against the CPU restatement, both outputs
    parity = None
    if not args.no_parity and rank == 0:
        step_dev()
        stream.synchronize(); torch.cuda.synchronize()           # rank 0 only: NO collective in this block
        errs = []
        for i in range(0, len(hams), 6):
            q = probs[i]
            op = cref_operator(q, 1, q["P"][..., 0], q["P"][..., 1])
            rg, rs = op.getghc(q["c"][..., 0] + 1j * q["c"][..., 1], sij_opt=1)
            g = ghcs[i].cpu().numpy(); s = gscs[i].cpu().numpy()
            for a, b in ((g, rg), (s, rs)):
                a = a[..., 0] + 1j * a[..., 1]
                errs.append(float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1))))
        parity = {"rel_err_vs_oracle": max(errs), "pairs_checked": len(range(0, len(hams), 6)), "bands": ndat, "tolerance": 1e-11,
                  "checker": "oracle/cref (C++/OpenMP restatement pinned on the NumPy oracle at 1e-13, tests/test_cref.py)"}
    # end to end: pinned host blocks -> device, batch, results back to the host
    e2e = None
    if not args.no_e2e:
        hc = cw_all.cpu().pin_memory()                   # host cg / ghc / gsc arrays of the whole sweep: one copy each way per step
        hg = torch.empty_like(hc).pin_memory(); hs = torch.empty_like(hc).pin_memory()

        def step_host():
            with torch.cuda.stream(stream):
                cw_all.copy_(hc, non_blocking=True)
                step_dev()
                hg.copy_(ghc_all, non_blocking=True)
                hs.copy_(gsc_all, non_blocking=True)
        for _ in range(3):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nbytes = sum(int(c.numel()) * 8 for c in cws)
        e2e = {"value": npairs * ndat * args.steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 2 * nbytes,
               "note": "bytes of this rank's pairs; the pairs of all ranks move concurrently"}
    if rank != 0:
        return None
    hbm, hbm_src = hbm_peak()
    w0 = dict(ngfft=FE2["ngfft"], kg=probs[0]["kg"], npw=probs[0]["npw"], nprojs=36)
    b_fw, f_nl, C = algorithmic_units(w0, 1, ndat, g=3)
    t_fw = sum(prof.get(k, (0.0, 0))[0] for k in ("fourwf_x_forward", "fourwf_plane_stage", "fourwf_plane_cluster", "fourwf_x_backward"))
    calls = max(1, len(hams) * args.steps)
    roof = {"kernel": "fourwf option 2 (3 kernels per call, timed in the un-batched loop)", "bound": "hbm",
            "achieved": b_fw * ndat * calls / (t_fw * 1e-3) / 1e9 if t_fw else None, "peak": hbm, "unit": "GB/s",
            "frac": b_fw * ndat * calls / (t_fw * 1e-3) / 1e9 / hbm if t_fw else None, "traffic": None, "peak_source": hbm_src,
            "note": "launch-latency regime: 6-9 kernels of a few microseconds per call; the batch path overlaps them on 8 lanes and replays CUDA graphs"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_subprocess(args)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": fe2_text(npairs, ndat, world), "l2": "L2-resident working set by nature (48 MB of blocks per step); launch-latency bound",
                      "parallelism": f"(k, spin) pairs round-robin over {world} GPU(s) (m_vtorho.F90:855-862), no data-path collective"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "parity": parity,
           "plain_loop": {"value": npairs * ndat * args.steps / (ms_plain * 1e-3), "unit": UNIT, "note": "one getghc call after the other on one stream"},
           "us_per_pair": 1e3 * ms / args.steps / max(1, len(hams)),
           "kernel_us_per_call_unbatched": {k: 1e3 * v[0] / max(1, v[1]) for k, v in prof.items()}}
    return out


def cpu_baseline_subprocess(args):
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", args.workload,
               "--istwfk", str(args.istwfk), "--cpu-bands", str(args.cpu_bands), "--ndat", str(args.ndat), "--au-lmax", str(args.au_lmax)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        return json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as ex:   # the baseline is reported, never a gate
        return {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": f"failed: {ex}"}


# ------------------------------------------------------------------------------------------------------------------------------
# one Hamiltonian, one band block per GPU (si512, au108, sweep points)
# ------------------------------------------------------------------------------------------------------------------------------
class Block:
    """Device side of one workload: the Hamiltonian handle, the band block and (rank 0, when a CPU leg follows) host copies of P."""

    def __init__(self, args, env, w, ndat, keep_host_p, seed=4321):
        import torch
        import abinit_b200 as ab
        dev, stream, rank = env["dev"], env["stream"], env["rank"]
        self.w, self.ndat = w, ndat
        npw, nprojs = w["npw"], w["nprojs"]
        self.ham = ab.Hamiltonian(w["ngfft"], w["natom"], 1, w["nlmn"], w["indlmn"], w["nattyp"], w["atindx1"], w["usepaw"], w["ucvol"])
        self.ham.load_spin(w["vlocal"], 1)
        if w["usepaw"]:
            self.ham.load_enl(w["dij"], w["sij"])
        else:
            self.ham.load_enl(w["ekb"], None)
        self.ham.load_k(args.istwfk, w["kg"], w["kinpw"], None, None, me_g0=1)
        with torch.cuda.stream(stream):
            # ONE operator for the whole job (the band-sharded ChebFi2 leg needs the same P on every rank), rank-specific bands
            gen = torch.Generator(device=dev).manual_seed(seed)
            P = torch.randn((nprojs, npw, 2), generator=gen, device=dev, dtype=torch.float64) / np.sqrt(npw)
            if args.istwfk == 2:
                P[:, 0, 1] = 0.0
            gen = torch.Generator(device=dev).manual_seed(seed + 1000 + rank)
            self.cw = torch.randn((ndat, npw, 2), generator=gen, device=dev, dtype=torch.float64)
            if args.istwfk == 2:
                self.cw[:, 0, 1] = 0.0
            self.ghc = torch.zeros_like(self.cw)
            self.gsc = torch.zeros_like(self.cw) if w["usepaw"] else None
        stream.synchronize()
        self.Pr = self.Pi = None
        if keep_host_p:
            self.Pr = P[..., 0].contiguous().cpu().numpy(); self.Pi = P[..., 1].contiguous().cpu().numpy()
        self.ham.set_projectors(P, nprojs)
        del P
        torch.cuda.empty_cache()
        self.sij_opt = 1 if w["usepaw"] else 0

    def step(self, cw=None, ghc=None, gsc=None):
        import abinit_b200 as ab
        ab.getghc(-1, self.cw if cw is None else cw, None, self.ghc if ghc is None else ghc,
                  (self.gsc if gsc is None else gsc) if self.sij_opt else None, self.ham, None, None, None, self.ndat, sij_opt=self.sij_opt)

    def destroy(self):
        self.ham.destroy()


def measure_block(args, env, blk, steps, warmup, want_e2e=True):
    """value (device-resident, CUDA events, max over ranks), per-kernel-class times, launches, clocks, e2e."""
    import torch
    import abinit_b200 as ab
    from abinit_b200 import api
    dev, stream, dist, barrier, world = env["dev"], env["stream"], env["dist"], env["barrier"], env["world"]
    ndat, npw = blk.ndat, blk.w["npw"]
    api.set_async(True)
    for _ in range(max(3, warmup)):
        blk.step()
    barrier()
    sampler = ClockSampler(env["local"]); sampler.start()
    l0 = ab.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        blk.step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ab.kernel_launches() - l0
    clocks = sampler.stop()
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    res = {"ms": ms, "value": world * ndat * steps / (ms * 1e-3), "launches": int(launches), "clocks": clocks}
    # per-kernel-class device times (second pass, same work) for the roofline object
    api.profile_enable(True)
    for _ in range(steps):
        blk.step()
    res["prof"] = api.profile_collect()
    api.profile_enable(False)
    res["e2e"] = None
    if want_e2e:
        api.set_async(False)
        ngs = 2 if blk.sij_opt else 1
        numa = numa_prefer_gpu_node(dev)              # pinned pages on the GPU's own NUMA node (round-1 e2e limit at N = 8)
        h_c = torch.empty((ndat, npw, 2), dtype=torch.float64).pin_memory(); h_c.copy_(blk.cw.cpu())
        h_g = torch.empty((ndat, npw, 2), dtype=torch.float64).pin_memory(); h_g.zero_()
        h_s = torch.empty((ndat, npw, 2), dtype=torch.float64).pin_memory() if blk.sij_opt else None
        if h_s is not None:
            h_s.zero_()
        numa_reset_policy()
        hc, hg, hs = h_c.numpy(), h_g.numpy(), (h_s.numpy() if h_s is not None else None)
        for _ in range(2):
            blk.step(hc, hg, hs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            blk.step(hc, hg, hs)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": world * ndat * steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(ndat * npw * 16),
                      "d2h_bytes_per_step": int(ngs * ndat * npw * 16), "host_numa": numa}
        api.set_async(True)
    return res


def _set_mempolicy(mode, node):
    """set_mempolicy(2) through libc.syscall (no libnuma in the image); returns True on success."""
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    if node is None:
        return libc.syscall(238, 0, None, 0) == 0                       # MPOL_DEFAULT
    mask = (ctypes.c_ulong * 16)()
    mask[node // 64] = 1 << (node % 64)
    return libc.syscall(238, mode, mask, 16 * 64 + 1) == 0


def numa_prefer_gpu_node(dev):
    """Host buffers of the e2e leg: prefer the NUMA node the GPU hangs off (sysfs numa_node of its PCI function) for the pages
    allocated next -- first touch happens right after the allocation.  Anything unexpected leaves the policy unchanged."""
    info = {"gpu_node": None, "policy": "default"}
    try:
        import torch
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        info["gpu_node"] = node
        if node >= 0 and os.path.isdir("/sys/devices/system/node/node%d" % node) and _set_mempolicy(1, node):   # MPOL_PREFERRED
            info["policy"] = "preferred"
    except Exception as e:                                                  # noqa: BLE001
        info["error"] = str(e)[:80]
    return info


def numa_reset_policy():
    try:
        _set_mempolicy(0, None)
    except Exception:                                                       # noqa: BLE001
        pass


def rooflines(args, blk, res, steps, fp64):
    """roofline (dominant kernel) + per-part rooflines from the in-library per-kernel-class timers."""
    w, ndat = blk.w, blk.ndat
    prof = res["prof"]
    hbm, hbm_src = hbm_peak()
    g = 3 if blk.sij_opt else 2
    b_fw, f_nl, C = algorithmic_units(w, args.istwfk, ndat, g=g)

    def per_launch(name):
        t, c = prof.get(name, (0.0, 0))
        return (t / c) if c else None
    t_tn, t_nn = per_launch("dgemm_tn_opernla"), per_launch("dgemm_nn_opernlb")
    t_fw = sum(prof.get(k, (0.0, 0))[0] for k in ("fourwf_x_forward", "fourwf_plane_stage", "fourwf_plane_cluster", "fourwf_x_backward")) / max(1, steps)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
    except Exception:
        pass
    flops_gemm = f_nl * ndat / g                               # one GEMM launch of a getghc step
    roof = None; extra = {}
    if t_nn:
        ach = flops_gemm / (t_nn * 1e-3) / 1e12
        roof = {"kernel": "k_dgemm_nn (opernlb: vect = P . gxfac, DMMA m8n8k4)", "bound": "tensor", "achieved": ach, "peak": fp64["peak"],
                "unit": "TFLOP/s", "frac": ach / fp64["peak"], "traffic": (traffic.get("k_dgemm_nn") or {}).get("bytes") if w["name"] == "si512" else None,
                "traffic_source": (traffic.get("k_dgemm_nn") or {}).get("source") if w["name"] == "si512" else None,
                "peak_source": fp64["source"], "peak_candidates": {k: fp64.get(k) for k in ("cublas_dgemm_8192", "dfma_probe", "dmma_probe")},
                "flops_per_launch": flops_gemm, "ms_per_launch": t_nn}
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
        ops_step = 28.0 * f_nl * ndat
        ach = ops_step / (t_i8 / steps * 1e-3) / 1e12
        roof = {"kernel": "k_igemm_tc (tcgen05.mma kind::i8, int32 accumulators in TMEM; 14 launches per step)", "bound": "tensor",
                "achieved": ach, "peak": peak_i8, "unit": "TOP/s (int8)", "frac": ach / peak_i8, "traffic": None,
                "peak_source": "2 x bf16_tflops of MEASURED_PEAKS.json (int8 dense = 2 x bf16 dense)" if bf16 else "nominal 4.5 Pop/s",
                "ops_per_step": ops_step, "ms_per_step": t_i8 / steps}
    if t_tn:
        ach = flops_gemm / (t_tn * 1e-3) / 1e12
        extra["roofline_opernla"] = {"kernel": "k_dgemm_tn (split-K P^T psi)", "bound": "tensor", "achieved": ach, "peak": fp64["peak"],
                                     "unit": "TFLOP/s", "frac": ach / fp64["peak"], "ms_per_launch": t_tn}
    if t_fw:
        ach = b_fw * ndat / (t_fw * 1e-3) / 1e9
        f_fw = fourwf_flops_per_band(w, args.istwfk, C)
        extra["roofline_fourwf"] = {
            "kernel": "fourwf option 2 (3 fused kernels)", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "peak_source": hbm_src, "bytes_per_band": b_fw, "ms_per_step": t_fw, "lines_C": C,
            # fourwf in FP64 sits above the FP64 ridge of the chip (AI ~ 6-11 flop/B vs 5.6): the FP64 pipe is its other bound
            "fp64": {"flops_per_band": f_fw, "achieved": f_fw * ndat / (t_fw * 1e-3) / 1e12, "peak": fp64["peak"], "unit": "TFLOP/s",
                     "frac": f_fw * ndat / (t_fw * 1e-3) / 1e12 / fp64["peak"],
                     "note": "nominal 5 n log2 n flops; DADD/DMUL issue at the DFMA rate, so the pipe is busier than this fraction"}}
    return roof, extra


def parity_check(args, blk, nb=4):
    """The timed configuration against the CPU restatement on the first nb bands (same P, same block)."""
    import torch
    op = cref_operator(blk.w, args.istwfk, blk.Pr, blk.Pi)
    blk.step(); torch.cuda.synchronize()
    c = blk.cw[:nb].cpu().numpy(); c = c[..., 0] + 1j * c[..., 1]
    rg, rs = op.getghc(c, sij_opt=blk.sij_opt)
    g = blk.ghc[:nb].cpu().numpy(); g = g[..., 0] + 1j * g[..., 1]
    err = float(np.max(np.linalg.norm(g - rg, axis=1) / np.linalg.norm(rg, axis=1)))
    out = {"rel_err_vs_oracle": err, "bands": nb, "tolerance": 1e-11,
           "checker": "oracle/cref (C++/OpenMP restatement pinned on the NumPy oracle at 1e-13, tests/test_cref.py; the NumPy oracle is pinned on the reference's stored SCF results)"}
    if blk.sij_opt:
        s = blk.gsc[:nb].cpu().numpy(); s = s[..., 0] + 1j * s[..., 1]
        out["rel_err_gsc"] = float(np.max(np.linalg.norm(s - rs, axis=1) / np.linalg.norm(rs, axis=1)))
    return out, op


def cpu_baseline_inprocess(args, blk, op, nb, steps=2):
    cores = host_cores()
    rng = np.random.Generator(np.random.PCG64(99))
    c = rng.standard_normal((nb, blk.w["npw"])) + 1j * rng.standard_normal((nb, blk.w["npw"]))
    if args.istwfk == 2:
        c[:, 0] = c[:, 0].real
    dt = time_cref(op, c, blk.sij_opt, steps, warmup=1)
    return {"value": nb / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": CPU_SAMPLE_TEXT.format(nb=nb, steps=steps, cores=cores),
            "seconds_by_part": dict(zip(("fourwf", "opernla", "opernlc", "opernlb_assembly"), [float(x) for x in op.timings]))}


def run_block_workload(args, env):
    """si512 / au108: the bench line of one (Hamiltonian, band block) workload."""
    import torch
    import abinit_b200 as ab
    from abinit_b200 import api
    rank, world, dev, stream, dist, barrier = env["rank"], env["world"], env["dev"], env["stream"], env["dist"], env["barrier"]
    usepaw = 1 if args.workload == "au108" else 0
    w = build_workload(args, usepaw)
    ndat, npw, nprojs = args.ndat, w["npw"], w["nprojs"]
    fp64 = measure_fp64_peak(dev)
    cpu_legs = rank == 0 and not (args.no_parity and (args.no_cpu_baseline or world > 1))
    blk = Block(args, env, w, ndat, keep_host_p=cpu_legs)
    if args.nonlop == "int8":
        api.set_tuning("nonlop_ozaki", 1)
    res = measure_block(args, env, blk, args.steps, args.warmup, want_e2e=not args.no_e2e)
    prof = res["prof"]
    cw, ghc, ham = blk.cw, blk.ghc, blk.ham

    # SCF-step-equivalent (BASELINE metric "s/SCF step"): ONE ChebFi2 call on nband bands at this k-point, band-sharded over
    # the GPUs: (ndeg+1) getghc passes + Rayleigh-Ritz (all-to-all re-layout, Gram allreduce over NCCL, hegvd, rotations).
    # The start block is a function of the GLOBAL band index only, so the eigenvalues must agree between runs on 1, 2, 4, 8 GPUs.
    scf_step = None
    extras = args.workload == "si512" and not args.no_scf_step
    if extras:
        from abinit_b200 import parallel as par
        api.set_async(False)
        # "native": abi_b200_chebfiwf2_paral_ (NCCL inside the library, collectives on the library stream); "python": the same
        # scheme strung together from the xg_* entries with torch.distributed (abinit_b200/parallel.py)
        cheb_paral = par.chebfi_band_parallel_native if args.scf_driver == "native" else par.chebfi_band_parallel
        f, l = par.band_block(args.nband, world, rank)
        with torch.cuda.stream(stream):
            damp = torch.from_numpy(1.0 / (1.0 + np.minimum(w["kinpw"], 1e6))).to(dev)
            cg = torch.empty((l - f, npw, 2), device=dev, dtype=torch.float64)
            for b in range(f, l):
                gen = torch.Generator(device=dev).manual_seed(777000 + b)
                cg[b - f] = torch.randn((npw, 2), generator=gen, device=dev, dtype=torch.float64)
            cg *= damp[None, :, None]
            if args.istwfk == 2:
                cg[:, 0, 1] = 0.0
            times = []
            for it in range(2):                                   # first call warms up (plans, cuSOLVER handle, workspaces)
                barrier()
                t0 = time.perf_counter()
                l1 = ab.kernel_launches()
                eig, resid = cheb_paral(ham, cg, args.nband, float(w["cfg"]["ecut"]), args.nline, bandpp=ndat)
                barrier()
                times.append(time.perf_counter() - t0)
            dtm = torch.tensor([times[-1]], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(dtm, op=dist.ReduceOp.MAX)
        api.profile_enable(True)
        with torch.cuda.stream(stream):
            cheb_paral(ham, cg, args.nband, float(w["cfg"]["ecut"]), args.nline, bandpp=ndat)
        prof_scf = api.profile_collect()
        api.profile_enable(False)
        eig = np.asarray(eig)
        scf_step = {"value": float(dtm.item()), "unit": "s per ChebFi2 call (one k-point, SCF-step-equivalent)", "nband": args.nband,
                    "nline": args.nline, "bands_per_gpu": l - f, "driver": args.scf_driver, "launches": int(ab.kernel_launches() - l1),
                    "eig_min_max": [float(np.min(eig)), float(np.max(eig))], "resid_max": float(np.max(resid)),
                    "cross_n_invariant": {"what": "eigenvalues after two ChebFi2 calls from a start block seeded per GLOBAL band index: identical input on every "
                                                  "GPU count, results must agree to 1e-8 Ha between the N = 1, 2, 4, 8 lines",
                                          "eig_sum": float(np.sum(eig)), "eig_first": [float(x) for x in eig[:4]], "eig_last": [float(x) for x in eig[-4:]]},
                    "timing": "host clock between barrier + device synchronise on both sides, max over ranks (the call holds host syncs)",
                    "kernel_ms": {k: v[0] for k, v in prof_scf.items()}}
        del cg
    # the same SCF-step-equivalent with LOBPCG (one block of all bands, nline LOBPCG iterations): lobpcgwf2 on one GPU, the library's
    # band-parallel driver (abi_b200_lobpcgwf2_paral_) on several; start block seeded per GLOBAL band index like the ChebFi2 leg
    lobpcg_step = None
    if extras:
        from abinit_b200 import xg as xgm, parallel as par
        api.set_async(False)
        f, l = par.band_block(args.nband, world, rank)
        with torch.cuda.stream(stream):
            damp = torch.from_numpy(1.0 / (1.0 + np.minimum(w["kinpw"], 1e6))).to(dev)
            cg0 = torch.empty((l - f, npw, 2), device=dev, dtype=torch.float64)
            for b in range(f, l):
                gen = torch.Generator(device=dev).manual_seed(778000 + b)
                cg0[b - f] = torch.randn((npw, 2), generator=gen, device=dev, dtype=torch.float64)
            cg0 *= damp[None, :, None]
            if args.istwfk == 2:
                cg0[:, 0, 1] = 0.0
            eigl = np.zeros(args.nband); resl = np.zeros(args.nband)
            tl = []
            for it in range(2):                                   # second call timed, both from the same start block
                cgl = cg0.clone()
                barrier(); t0 = time.perf_counter()
                if world == 1:
                    xgm.lobpcgwf2(cgl, eigl, None, None, ham, args.nband, npw, 1, resl, 1e-30, args.nline, bandpp=ndat)
                else:
                    eigl, resl = par.lobpcg_band_parallel_native(ham, cgl, args.nband, args.nline, bandpp=ndat)
                barrier(); tl.append(time.perf_counter() - t0)
            dtl = torch.tensor([tl[-1]], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(dtl, op=dist.ReduceOp.MAX)
        eigl = np.asarray(eigl); resl = np.asarray(resl)
        lobpcg_step = {"value": float(dtl.item()), "unit": "s per LOBPCG call (one k-point, one block of all bands)", "nband": args.nband,
                       "nline": args.nline, "bands_per_gpu": l - f, "driver": "lobpcgwf2" if world == 1 else "abi_b200_lobpcgwf2_paral_ (NCCL inside the library)",
                       "eig_min_max": [float(eigl.min()), float(eigl.max())], "resid_max": float(resl.max()),
                       "cross_n_invariant": {"eig_sum": float(np.sum(eigl)), "eig_first": [float(x) for x in eigl[:4]]}}
        del cgl, cg0
        torch.cuda.empty_cache()
    # density build (the step after the solver): fourwf option 1 on the same band block, fused path
    density = None
    if extras:
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
        density = {"ms_per_block": dms, "bands": ndat, "bands_per_s": ndat / (dms * 1e-3),
                   "kernel": "fourwf option 1, fused (x pass, plane stage with density reduction, transpose-add)"}
        del rho
    # sanity anchor against the only fourwf timing the reference stores: option 2, cplex 1, istwfk 1,
    # box 100^3 (ecut 30 Ha, 20 Bohr cube, k = (.1,.2,.3)): 18.8 ms per call with FFTW3 on one CPU core
    # (tests/unitary/Refs/tfourwf_01.stdout:117-129)
    anchor = None                                     # rank 0 only: NO collective in this block (local synchronisation)
    if rank == 0 and extras:
        from abinit_b200 import workload as wl2
        kg_a, _ = wl2.gsphere_orthorhombic(30.0, 20.0, (0.1, 0.2, 0.3), 1)
        npw_a = kg_a.shape[0]
        api.set_async(True)
        with torch.cuda.stream(stream):
            va_h = wl2.smooth_potential((100, 100, 100), seed=3)
            va = torch.from_numpy(va_h).to(dev)
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
        if not args.no_cpu_baseline and world == 1:
            try:                                                  # the same call through the CPU restatement on ONE thread
                from oracle import cref
                cref.set_threads(1)
                c1 = np.random.default_rng(0).standard_normal((2, npw_a)) + 1j * np.random.default_rng(1).standard_normal((2, npw_a))
                cref.fourwf_option2(va_h, c1, kg_a, (100, 100, 100), 1)
                t0 = time.time(); cref.fourwf_option2(va_h, c1, kg_a, (100, 100, 100), 1)
                anchor["cpu_port_ms_per_band_1thread"] = (time.time() - t0) / 2 * 1e3
                cref.set_threads(host_cores())
            except Exception as ex:                               # noqa: BLE001
                anchor["cpu_port_error"] = repr(ex)
    # parity of the timed configuration + the CPU baseline, both on the same host copy of P (rank 0)
    parity = None; cpu = None
    if rank == 0 and blk.Pr is not None:
        try:
            api.set_async(False)
            parity, op = parity_check(args, blk)
            if not args.no_cpu_baseline and world == 1:
                cpu = cpu_baseline_inprocess(args, blk, op, args.cpu_bands)
            del op
        except Exception as ex:                                   # noqa: BLE001  (reported, never a gate)
            parity = parity or {"rel_err_vs_oracle": None, "error": repr(ex)}
        blk.Pr = blk.Pi = None
    if rank != 0:
        return None
    roof, extra = rooflines(args, blk, res, args.steps, fp64)
    # EXPERIMENTAL (reported separately, never part of `value`): gemm_nonlop's two contractions through exact int8 slice
    # products (csrc/ozaki.cu) instead of the FP64 DMMA kernels.  Runs in its own process after this one has released the device.
    experimental = None
    if world == 1 and extras and args.nonlop == "fp64":
        try:
            blk.destroy(); ab.finalize(); del cw, ghc, blk
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ozaki_bench.py"), "--json"], capture_output=True, text=True, timeout=600)
            experimental = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as ex:                                   # noqa: BLE001
            experimental = {"error": repr(ex)}
    out = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": res["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64" if args.nonlop == "fp64" else "f64 (fourwf) + exact int8 slices with int32 accumulation recombined in f64 (gemm_nonlop)",
           "data": "synthetic",
           "config": {"workload": workload_text(args, w, ndat),
                      "l2": "inputs larger than L2 (P = %.1f GB streamed %d times per step)" % (16.0 * npw * nprojs / 1e9, 3 if usepaw else 2),
                      "parallelism": f"band blocks over {world} GPU(s), no data-path collective"},
           "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["launches"], "roofline": roof, "cpu_baseline": cpu, "parity": parity,
           "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()}, "scf_step": scf_step, "lobpcg_step": lobpcg_step,
           "density_step": density, "fourwf_anchor": anchor, "experimental_int8_sliced": experimental,
           "fp64_peak_measured": {k: fp64.get(k) for k in ("cublas_dgemm_8192", "dfma_probe", "dmma_probe", "peak")}}
    out.update(extra)
    return out


# ------------------------------------------------------------------------------------------------------------------------------
# sweep (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------------------------------------
def sweep_points(quick):
    boxes = (48, 64, 96, 128) if quick else (48, 64, 96, 128, 144, 192)
    blocks = (64, 256) if quick else (64, 256, 1024)
    pts = [dict(box=n, ndat=b, istwfk=2, nprojs=None) for n in boxes for b in blocks]
    pts += [dict(box=96, ndat=256, istwfk=2, nprojs=p) for p in (512, 2048, 8192, 16384)]       # nprojs axis
    pts += [dict(box=n, ndat=256, istwfk=1, nprojs=None) for n in ((64, 96) if quick else (64, 96, 128))]   # general-k storage
    return pts


def run_sweep(args, env):
    import torch
    from abinit_b200 import workload as wl
    rank, world = env["rank"], env["world"]
    fp64 = measure_fp64_peak(env["dev"])
    out_path = args.out or os.path.join(ROOT, "gpurun_out", f"sweep_{world}gpu.jsonl")
    lines = []
    t_all = 0.0
    blk = None
    steps = max(2, min(args.steps, 5))
    for pt in sweep_points(args.sweep_quick):
        n = pt["box"]
        # sphere of radius r index units in a box n >= 4 r + 1 (boxcut 2); ecut 20 Ha fixes the cell length
        r = (n - 1) / 4.0 - 0.25
        L = 2 * np.pi * r / np.sqrt(2 * 20.0)
        npw_full = 4.0 / 3.0 * np.pi * r ** 3
        nlmn = 18
        natom = max(8, int(round(0.032 * npw_full / nlmn))) if pt["nprojs"] is None else max(1, pt["nprojs"] // nlmn)
        name = f"sweep{n}_{natom}"
        wl.CONFIGS[name] = dict(ecut=20.0, L=float(L), ngfft=(n, n, n), natom=natom, lmax=2, nproj_per_l=2)
        a2 = argparse.Namespace(**vars(args)); a2.workload = name; a2.istwfk = pt["istwfk"]
        if blk is not None:
            blk.destroy(); blk = None
            torch.cuda.empty_cache()
        w = build_workload(a2, 0)
        blk = Block(a2, env, w, pt["ndat"], keep_host_p=False)
        res = measure_block(a2, env, blk, steps, 3, want_e2e=False)
        if rank == 0:
            roof, extra = rooflines(a2, blk, res, steps, fp64)
            kms = {k: v[0] / steps for k, v in res["prof"].items()}
            line = {"box": n, "npw": w["npw"], "nprojs": w["nprojs"], "ndat": pt["ndat"], "istwfk": pt["istwfk"], "n_gpus": world,
                    "ms_per_step": res["ms"] / steps, "band_app_per_s": res["value"], "kernel_ms": {k: round(v, 4) for k, v in kms.items()},
                    "fourwf": {k: (extra.get("roofline_fourwf") or {}).get(k) for k in ("ms_per_step", "bytes_per_band", "achieved", "frac")},
                    "gemm_nonlop_nn": {k: (roof or {}).get(k) for k in ("ms_per_launch", "achieved", "frac")},
                    "gemm_nonlop_tn": {k: (extra.get("roofline_opernla") or {}).get(k) for k in ("ms_per_launch", "achieved", "frac")},
                    "clocks": res["clocks"]}
            lines.append(line)
            t_all += res["ms"] / steps
    if blk is not None:
        blk.destroy()
    if rank != 0:
        return None
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        for ln in lines:
            fh.write(json.dumps(ln) + "\n")
    gm = float(np.exp(np.mean([np.log(ln["band_app_per_s"]) for ln in lines])))
    return {"metric": METRIC, "value": gm, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": t_all / max(1, len(lines)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"sweep (BASELINE configs[4], SURVEY 8d): {len(lines)} points, boxes 48^3-192^3 at boxcut 2, band blocks 64-1024 per GPU, nprojs axis "
                                   f"512-16384 at 96^3, istwfk 1 rows; value = geometric mean over the points; per-point lines in {os.path.relpath(out_path, ROOT)}",
                       "parallelism": f"band blocks over {world} GPU(s), no data-path collective"},
            "clocks": lines[-1]["clocks"], "e2e": None, "gpu_launches": None, "roofline": None, "cpu_baseline": None,
            "points": [{k: ln[k] for k in ("box", "npw", "nprojs", "ndat", "istwfk", "band_app_per_s")} for ln in lines],
            "fp64_peak_measured": {k: fp64.get(k) for k in ("cublas_dgemm_8192", "dfma_probe", "dmma_probe", "peak")}}


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

    def barrier():
        stream.synchronize(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
    env = dict(rank=rank, world=world, local=local, dev=dev, stream=stream, dist=dist, barrier=barrier)
    if args.workload == "fe2":
        out = run_fe2(args, env)
    elif args.workload == "sweep":
        out = run_sweep(args, env)
    else:
        out = run_block_workload(args, env)
    if rank == 0 and out is not None:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
