#!/usr/bin/env python
"""Developer tool: cost of the gemm_nonlop GEMMs as a function of the band-block width (Si-512 shape): a ragged block must cost
what its bands cost (rounded up to 8 columns), not a whole 128-column tile."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import abinit_b200 as ab
from abinit_b200 import api
import bench

args = bench.parse(["--workload", "si512"])
ab.init(0)
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); api.set_stream(stream.cuda_stream)
env = dict(rank=0, world=1, local=0, dev=dev, stream=stream, dist=None, barrier=lambda: (stream.synchronize(), torch.cuda.synchronize()))
w = bench.build_workload(args, 0)
blk = bench.Block(args, env, w, 256, keep_host_p=False)
api.set_async(True)
for rag in (0, 47):
    api.set_tuning("nonlop_rag", rag)
    for nd in (128, 100, 76, 48, 40, 24, 19, 10, 4):
        blk.ndat = nd
        cw, ghc = blk.cw[:nd], blk.ghc[:nd]
        for _ in range(2):
            blk.step(cw, ghc)
        api.profile_enable(True)
        for _ in range(3):
            blk.step(cw, ghc)
        prof = api.profile_collect(); api.profile_enable(False)
        print(json.dumps({"nonlop_rag": rag, "ndat": nd, "ms": {k: round(v[0] / 3, 3) for k, v in prof.items() if k.startswith("dgemm")}}), flush=True)
