"""CPU suite: C-ABI symbols, host-side planners/sharding, multi-process (gloo, world_size 2) Gram allreduce."""
import ctypes
import os
import re
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built_lib():
    from abinit_b200 import build, lib
    path = lib.library_path()
    if not os.path.exists(path):
        build.build()
    return path


def test_library_exports_every_declared_symbol():
    """Every function include/abinit_b200.h declares is exported by libabinit_b200.so (no compute calls here)."""
    header = open(os.path.join(ROOT, "include", "abinit_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", header)) - {"defined"}
    declared = {d for d in declared if d.startswith(("abi_b200_", "gpu_fourwf_", "alloc_gpu_", "free_gpu_"))}
    assert len(declared) >= 30
    lib = ctypes.CDLL(_built_lib())
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from abinit_b200 import lib as blib
    assert set(blib.SYMBOLS) == declared
    lib.abi_b200_version.restype = ctypes.c_char_p
    assert b"abinit_b200" in lib.abi_b200_version()


def test_library_is_built_for_sm_100a_with_dmma():
    path = _built_lib()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_52" not in out
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in sass        # FP64 tensor-core path of gemm_nonlop
    assert "LDGSTS" in sass            # cp.async pipeline feeding it


def test_no_cuda_device_fails_loudly():
    """No CPU fallback: without a GPU the first compute entry point aborts with an ABINIT-style error document."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    code = ("import sys; sys.path.insert(0, %r); import abinit_b200 as ab; ab.init(0)" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "!ERROR" in r.stderr and "no CPU fallback" in r.stderr


def test_missing_library_raises(tmp_path):
    from abinit_b200 import lib as blib
    with pytest.raises(blib.LibraryNotBuilt):
        blib.load_library(str(tmp_path / "nope.so"))


def test_workload_sphere_matches_oracle():
    from abinit_b200 import workload as wl
    from oracle import gsphere as g
    for ist, k, L in ((1, (0, 0, 0), 12.0), (2, (0, 0, 0), 12.0), (1, (.1, .2, .3), 9.0), (3, (.5, 0, 0), 10.0)):
        kg, kin = wl.gsphere_orthorhombic(9.0, L, k, ist)
        _, gm, _ = g.metric(np.eye(3) * L)
        ref = g.kpgsph(9.0, gm, k, ist)
        assert np.array_equal(kg.T, ref)
        assert np.allclose(kin, g.mkkin(9.0, 0.0, 1.0, gm, ref, k), rtol=0, atol=1e-12)


def test_band_and_kpoint_sharding():
    from abinit_b200 import parallel as par
    for nband, n in ((1100, 8), (5, 2), (7, 8), (128, 1)):
        blocks = [par.band_block(nband, n, r) for r in range(n)]
        assert blocks[0][0] == 0 and blocks[-1][1] == nband
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(n - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert par.band_block(1100, 8, 0) == (0, 138) and par.band_block(1100, 8, 7) == (963, 1100)
    owned = [par.my_kpoints(72, 2, 8, r) for r in range(8)]
    flat = sorted(x for o in owned for x in o)
    assert flat == sorted((ik, isp) for isp in range(2) for ik in range(72))
    assert all(len(o) == 18 for o in owned)


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from abinit_b200 import parallel as par
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(11)
npw, nband = 301, 6
X = rng.standard_normal((nband, npw)) + 1j * rng.standard_normal((nband, npw))
AX = rng.standard_normal((nband, npw)) + 1j * rng.standard_normal((nband, npw))
lo, hi = par.row_shard(npw, world, rank)
# partial Gram on my plane-wave rows (xgBlock_gemm with comm=, m_xg.F90:1969-1974), then the allreduce
g1 = torch.from_numpy(np.conj(X[:, lo:hi]) @ AX[:, lo:hi].T)
g2 = torch.from_numpy(np.conj(X[:, lo:hi]) @ X[:, lo:hi].T)
par.gram_allreduce(g1, g2)
ok = np.allclose(g1.numpy(), np.conj(X) @ AX.T, atol=1e-12) and np.allclose(g2.numpy(), np.conj(X) @ X.T, atol=1e-12)
# band-block sharding: every rank applies its block, the union covers all bands exactly once
first, last = par.band_block(nband, world, rank)
cover = torch.zeros(nband, dtype=torch.float64); cover[first:last] = 1
dist.all_reduce(cover)
ok = ok and bool((cover == 1).all())
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_gram_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    rcs = [p.wait(timeout=180) for p in procs]
    assert rcs == [0, 0]


_WORKER_T = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from abinit_b200 import parallel as par
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(3)
ok = True
for npw, nband in ((301, 7), (64, 2), (1000, 33)):
    X = rng.standard_normal((nband, npw, 2))                       # the full block, same on every rank
    f, l = par.band_block(nband, world, rank)
    lo, hi = par.row_shard(npw, world, rank)
    cols = torch.from_numpy(np.ascontiguousarray(X[f:l]))
    rows = par.transpose_cols_to_rows(cols, nband, npw)            # xgTransposer STATE_COLSROWS -> STATE_LINALG
    ok = ok and rows.shape == (nband, hi - lo, 2) and np.array_equal(rows.numpy(), X[:, lo:hi])
    back = par.transpose_rows_to_cols(rows, nband, npw)            # and back
    ok = ok and np.array_equal(back.numpy(), X[f:l])
    # row-sharded SPACE_CR Gram partials (G=0 correction on the rank that owns row 0) sum to the full Gram
    from oracle import xg as oxg
    Xc = X[..., 0] + 1j * X[..., 1]
    part = oxg.gram(oxg.SPACE_CR, np.ascontiguousarray(Xc[:, lo:hi]), np.ascontiguousarray(Xc[:, lo:hi]), 1 if rank == 0 else 0)
    t = torch.from_numpy(part); dist.all_reduce(t)
    ok = ok and np.allclose(t.numpy(), oxg.gram(oxg.SPACE_CR, Xc, Xc, 1), atol=1e-10)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_xg_transposer_gloo_world2(tmp_path):
    """Band-sharded <-> row-sharded re-layout of the band-parallel ChebFi2 (abinit_b200.parallel), world size 2 on gloo."""
    script = tmp_path / "worker_t.py"
    script.write_text(_WORKER_T % {"root": ROOT})
    port = 31500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    rcs = [p.wait(timeout=180) for p in procs]
    assert rcs == [0, 0]
