#!/usr/bin/env python
"""Summarise an Nsight Compute report for profiles/: headline raw metrics + SASS opcode mix + stall reasons.

  python tools/ncu_summary.py gpurun_out/ncu_k_fw_mid_r01.ncu-rep [> profiles/<name>.txt]
"""
import csv, io, subprocess, sys, collections

RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
            "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = run(["-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) >= 3:
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            print(f"== kernel {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for k in hdr:
                if k in RAW_KEYS or any(s in k for s in ("pipe_fp64", "dmma", "bank_conflict")):
                    print(f"  {k:86s} {d[k]:>18s} {u.get(k, '')}")
    src = run(["-i", rep, "--page", "source", "--csv"])
    lines = src.splitlines()
    start = next((i for i, l in enumerate(lines) if l.startswith('"Address"')), None)
    if start is None:
        return
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    ops = collections.Counter(); stall = collections.Counter(); samples = collections.Counter(); tot = 0
    for r in rd:
        try:
            n = int(r["Instructions Executed"])
        except Exception:
            continue
        s = r["Source"].split()
        op = s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "DMMA", "BAR", "LDGSTS")) else op.split(".")[0]
        ops[op] += n; tot += n
        samples[op] += int(r.get("# Samples", 0) or 0)
        for k, v in r.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    stall[k] += int(v)
                except Exception:
                    pass
    print(f"-- SASS opcode mix (warp-level instructions executed, total {tot}):")
    for op, n in ops.most_common(28):
        print(f"  {op:18s} {n:>14d} {100.0 * n / max(tot, 1):6.2f} %   samples {samples[op]}")
    st = sum(stall.values())
    print(f"-- warp stall samples (total {st}):")
    for k, v in stall.most_common(12):
        print(f"  {k:28s} {v:>10d} {100.0 * v / max(st, 1):6.2f} %")


if __name__ == "__main__":
    main()
