"""Build libabinit_b200.so in-tree with nvcc for sm_100a (called by __graft_entry__.build())."""
from __future__ import annotations
import os, subprocess, sys, hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["fourwf.cu", "plane_stage.cu", "plane_inst_0.cu", "plane_inst_1.cu", "plane_inst_2.cu", "plane_inst_3.cu", "half_stage.cu", "half_inst_0.cu", "half_inst_1.cu", "half_inst_2.cu", "x_stage.cu", "x_inst_0.cu", "x_inst_1.cu", "nonlop.cu", "context.cu", "api_fourwf.cu", "api_nonlop.cu", "xg.cu", "api_xg.cu", "invovl.cu", "ozaki.cu", "comm.cu"]
HEADERS = ["common.cuh", "fft_engine.cuh", "plane_stage.cuh", "plane_stage_impl.cuh", "half_stage.cuh", "half_stage_impl.cuh", "hdft.cuh", "x_stage.cuh", "x_stage_impl.cuh", "roots.inc", "fourwf.cuh", "nonlop.cuh", "context.cuh", "ham.cuh", "xg.cuh", "comm.cuh",
           os.path.join("..", "..", "include", "abinit_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _stamp() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    out = os.path.join(HERE, "libabinit_b200.so")
    stamp_file = os.path.join(HERE, "build", "stamp.txt")
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    stamp = _stamp()
    if not force and os.path.exists(out) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
