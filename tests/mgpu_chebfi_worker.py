"""Worker of tests/test_chebfi_mgpu.py (one process per GPU, launched with torch.distributed.run): band-parallel ChebFi2
(abinit_b200.parallel.chebfi_band_parallel: NCCL all-to-all re-layout + Gram allreduce) against the single-GPU chebfiwf2
on the same start block."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import abinit_b200 as ab                      # noqa: E402
from abinit_b200 import parallel as par, xg   # noqa: E402
from problems import make_problem             # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ab.init(local)
    ok = True
    for istwf_k, kpt, nband in ((2, (0, 0, 0), 13), (1, (-.25, .5, 0), 10)):
        p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False)
        h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 0, p.ucvol)
        h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, None); h.load_k(istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
        # single-GPU reference on every rank (same input -> same output)
        cg1 = p.cwavef.copy(); eig1 = np.zeros(nband); res1 = np.zeros(nband)
        xg.chebfiwf2(cg1, eig1, None, None, h, nband, p.npw, 1, res1, 1e-16, p.ecut, 5, bandpp=4)
        f, l = par.band_block(nband, world, rank)
        cg = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eig, res = par.chebfi_band_parallel(h, cg, nband, p.ecut, 5, bandpp=4)
        e_err = float(np.max(np.abs(eig - eig1))); e_ok = e_err < 1e-8      # north_star: eigenvalues within 1e-8 Ha
        r_ok = np.max(np.abs(res - res1[f:l]) / (np.abs(res1[f:l]) + 1e-12)) < 1e-5
        c = cg.cpu().numpy(); c = c[..., 0] + 1j * c[..., 1]
        v_ok = np.max(np.abs(np.abs(c) - np.abs(cg1[f:l]))) < 1e-8
        print(f"rank {rank} istwf_k {istwf_k}: eig {e_ok} ({e_err:.2e}) resid {r_ok} vec {v_ok}", flush=True)
        ok = ok and e_ok and r_ok and v_ok
        # the same through the library's own driver (NCCL inside libabinit_b200.so, abi_b200_chebfiwf2_paral_)
        cgn = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eign, resn = par.chebfi_band_parallel_native(h, cgn, nband, p.ecut, 5, bandpp=4)
        ne_err = float(np.max(np.abs(eign - eig1)))
        nr_ok = np.max(np.abs(resn - res1[f:l]) / (np.abs(res1[f:l]) + 1e-12)) < 1e-5
        cn = cgn.cpu().numpy(); cn = cn[..., 0] + 1j * cn[..., 1]
        nv_ok = np.max(np.abs(np.abs(cn) - np.abs(cg1[f:l]))) < 1e-8
        print(f"rank {rank} istwf_k {istwf_k}: native eig {ne_err:.2e} resid {nr_ok} vec {nv_ok}", flush=True)
        ok = ok and ne_err < 1e-8 and nr_ok and nv_ok
        # oracle-driven degree (chebfi_oracle = 1 with a band buffer): per-rank degrees, MAX over the ranks
        occ = np.where(np.arange(nband) < nband - 3, 1.0, 0.0)
        cgo1 = p.cwavef.copy(); eigo1 = np.zeros(nband); reso1 = np.zeros(nband)
        xg.chebfiwf2(cgo1, eigo1, occ, None, h, nband, p.npw, 1, reso1, 1e-16, p.ecut, 5, nbdbuf=2, chebfi_oracle=1, bandpp=4)
        cgo = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eigo, _ = par.chebfi_band_parallel_native(h, cgo, nband, p.ecut, 5, bandpp=4, occ=occ, nbdbuf=2, chebfi_oracle=1)
        oe_err = float(np.max(np.abs(eigo - eigo1)))
        print(f"rank {rank} istwf_k {istwf_k}: native oracle=1 eig {oe_err:.2e}", flush=True)
        ok = ok and oe_err < 1e-8
        # band-parallel LOBPCG (row-sharded linear algebra, Gram allreduce) vs the single-GPU lobpcgwf2
        cgl1 = p.cwavef.copy(); eigl1 = np.zeros(nband); resl1 = np.zeros(nband)
        xg.lobpcgwf2(cgl1, eigl1, None, None, h, nband, p.npw, 1, resl1, 1e-30, 3, bandpp=4)
        cgl = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eigl, resl = par.lobpcg_band_parallel(h, cgl, nband, 3, p.kinpw, bandpp=4)
        le_err = float(np.max(np.abs(eigl - eigl1))); lr_ok = np.max(np.abs(resl - resl1) / (np.abs(resl1) + 1e-12)) < 1e-4
        cl = cgl.cpu().numpy(); cl = cl[..., 0] + 1j * cl[..., 1]
        lv_ok = np.max(np.abs(np.abs(cl) - np.abs(cgl1[f:l]))) < 1e-7
        print(f"rank {rank} istwf_k {istwf_k}: lobpcg eig {le_err:.2e} resid {lr_ok} vec {lv_ok}", flush=True)
        ok = ok and le_err < 1e-8 and lr_ok and lv_ok
        cgn = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eign, resn = par.lobpcg_band_parallel_native(h, cgn, nband, 3, bandpp=4)
        ne_err = float(np.max(np.abs(eign - eigl1))); nr_ok = np.max(np.abs(resn - resl1) / (np.abs(resl1) + 1e-12)) < 1e-4
        cn = cgn.cpu().numpy(); cn = cn[..., 0] + 1j * cn[..., 1]
        nv_ok = np.max(np.abs(np.abs(cn) - np.abs(cgl1[f:l]))) < 1e-7
        print(f"rank {rank} istwf_k {istwf_k}: native lobpcg eig {ne_err:.2e} resid {nr_ok} vec {nv_ok}", flush=True)
        ok = ok and ne_err < 1e-8 and nr_ok and nv_ok
        h.destroy()
    # PAW (B = S): S X in the Rayleigh quotients, apply_invovl in the filter, X^H S X in the Rayleigh-Ritz step, BX transposed too
    for istwf_k, kpt, nband in ((2, (0, 0, 0), 11), (1, (-.25, .5, 0), 10)):
        p = make_problem(7.0, (8.0, 9.0, 7.5), kpt, istwf_k, ndat=nband, natom_per_type=(2,), lmax_per_type=(1,), filter_shell=False, usepaw=1)
        h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, 1, p.ucvol)
        h.load_spin(p.vlocal, p.cplex); h.load_enl(p.enl, p.sij); h.load_k(istwf_k, p.kgF, p.kinpw, p.ffnl, p.ph3d, me_g0=1)
        cg1 = p.cwavef.copy(); eig1 = np.zeros(nband); res1 = np.zeros(nband)
        xg.chebfiwf2(cg1, eig1, None, None, h, nband, p.npw, 1, res1, 1e-16, p.ecut, 5, bandpp=4)
        f, l = par.band_block(nband, world, rank)
        cg = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eig, res = par.chebfi_band_parallel(h, cg, nband, p.ecut, 5, bandpp=4)
        e_err = float(np.max(np.abs(eig - eig1))); e_ok = e_err < 1e-8
        r_ok = np.max(np.abs(res - res1[f:l]) / (np.abs(res1[f:l]) + 1e-12)) < 1e-5
        c = cg.cpu().numpy(); c = c[..., 0] + 1j * c[..., 1]
        v_ok = np.max(np.abs(np.abs(c) - np.abs(cg1[f:l]))) < 1e-8
        print(f"rank {rank} istwf_k {istwf_k} PAW: eig {e_ok} ({e_err:.2e}) resid {r_ok} vec {v_ok}", flush=True)
        ok = ok and e_ok and r_ok and v_ok
        cgn = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eign, resn = par.chebfi_band_parallel_native(h, cgn, nband, p.ecut, 5, bandpp=4)
        ne_err = float(np.max(np.abs(eign - eig1)))
        nr_ok = np.max(np.abs(resn - res1[f:l]) / (np.abs(res1[f:l]) + 1e-12)) < 1e-5
        cn = cgn.cpu().numpy(); cn = cn[..., 0] + 1j * cn[..., 1]
        nv_ok = np.max(np.abs(np.abs(cn) - np.abs(cg1[f:l]))) < 1e-8
        print(f"rank {rank} istwf_k {istwf_k} PAW: native eig {ne_err:.2e} resid {nr_ok} vec {nv_ok}", flush=True)
        ok = ok and ne_err < 1e-8 and nr_ok and nv_ok
        cgl1 = p.cwavef.copy(); eigl1 = np.zeros(nband); resl1 = np.zeros(nband)
        xg.lobpcgwf2(cgl1, eigl1, None, None, h, nband, p.npw, 1, resl1, 1e-30, 3, bandpp=4)
        cgl = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eigl, resl = par.lobpcg_band_parallel(h, cgl, nband, 3, p.kinpw, bandpp=4)
        le_err = float(np.max(np.abs(eigl - eigl1))); lr_ok = np.max(np.abs(resl - resl1) / (np.abs(resl1) + 1e-12)) < 1e-4
        cl = cgl.cpu().numpy(); cl = cl[..., 0] + 1j * cl[..., 1]
        lv_ok = np.max(np.abs(np.abs(cl) - np.abs(cgl1[f:l]))) < 1e-7
        print(f"rank {rank} istwf_k {istwf_k} PAW: lobpcg eig {le_err:.2e} resid {lr_ok} vec {lv_ok}", flush=True)
        ok = ok and le_err < 1e-8 and lr_ok and lv_ok
        cgn = torch.from_numpy(np.ascontiguousarray(p.cwavef[f:l]).view(np.float64).reshape(l - f, p.npw, 2)).cuda()
        eign, resn = par.lobpcg_band_parallel_native(h, cgn, nband, 3, bandpp=4)
        ne_err = float(np.max(np.abs(eign - eigl1))); nr_ok = np.max(np.abs(resn - resl1) / (np.abs(resl1) + 1e-12)) < 1e-4
        cn = cgn.cpu().numpy(); cn = cn[..., 0] + 1j * cn[..., 1]
        nv_ok = np.max(np.abs(np.abs(cn) - np.abs(cgl1[f:l]))) < 1e-7
        print(f"rank {rank} istwf_k {istwf_k} PAW: native lobpcg eig {ne_err:.2e} resid {nr_ok} vec {nv_ok}", flush=True)
        ok = ok and ne_err < 1e-8 and nr_ok and nv_ok
        h.destroy()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
