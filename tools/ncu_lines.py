#!/usr/bin/env python
"""Per-source-line view of an Nsight Compute report (needs -lineinfo + --import-source on):
executed warp instructions by class (fp64 / memory / other) and stall samples per CUDA source line.

  python tools/ncu_lines.py gpurun_out/x.ncu-rep [min_share_percent]"""
import csv, io, subprocess, sys, collections

def main():
    rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    files = []; cur = None; hdr = None
    agg = collections.OrderedDict()
    i = 0
    while i < len(lines):
        l = lines[i]
        if l.startswith('"File Path"'):
            cur = next(csv.reader([l]))[1]; i += 1; continue
        if l.startswith('"Function Name"'):
            i += 1; continue
        if l.startswith('"Line No"'):
            hdr = next(csv.reader([l])); i += 1; continue
        row = next(csv.reader([l]))
        if hdr is None or len(row) < len(hdr) - 2:
            i += 1; continue
        d = dict(zip(range(len(hdr)), row))
        lineno, src, addr, sass = row[0], row[1], row[2], row[3]
        if lineno != "":
            key = (cur.split("/")[-1], int(lineno), src.strip()[:100]); agg.setdefault(key, collections.Counter())
            curkey = key
            try: samples = int(row[hdr.index("# Samples")] or 0)
            except ValueError: samples = 0
            agg[key]["samples"] += samples
            for k in ("stall_long_sb", "stall_wait", "stall_barrier", "stall_short_sb", "stall_math", "stall_mio", "stall_lg", "stall_branch_resolving"):
                idx = hdr.index(k)
                try: agg[key][k] += int(row[idx] or 0)
                except ValueError: pass
        else:
            try: n = int(row[hdr.index("Instructions Executed")] or 0)
            except ValueError: n = 0
            s = sass.split()
            op = s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")
            base = op.split(".")[0]
            cls = "fp64" if base in ("DADD", "DFMA", "DMUL") else ("mem" if base in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LDL", "STL", "RED", "ATOMG") else "other")
            agg[curkey][cls] += n; agg[curkey]["inst"] += n; agg[curkey]["op_" + base] += n
        i += 1
    tot = sum(v["inst"] for v in agg.values()); tots = sum(v["samples"] for v in agg.values())
    print(f"total warp instructions {tot}, samples {tots}")
    print(f"{'file:line':28s} {'inst%':>6s} {'fp64':>10s} {'mem':>10s} {'other':>10s} {'smp%':>6s} {'long':>6s} {'wait':>6s} {'bar':>6s} {'short':>6s} {'math':>6s}  source / top other ops")
    for (f, ln, src), v in agg.items():
        if 100.0 * v["inst"] / max(tot, 1) < thr and 100.0 * v["samples"] / max(tots, 1) < thr:
            continue
        others = sorted(((k[3:], n) for k, n in v.items() if k.startswith("op_") and k[3:] not in ("DADD", "DFMA", "DMUL")), key=lambda x: -x[1])[:5]
        print(f"{f + ':' + str(ln):28s} {100.0 * v['inst'] / max(tot, 1):6.2f} {v['fp64']:10d} {v['mem']:10d} {v['other']:10d} {100.0 * v['samples'] / max(tots, 1):6.2f} "
              f"{v['stall_long_sb']:6d} {v['stall_wait']:6d} {v['stall_barrier']:6d} {v['stall_short_sb']:6d} {v['stall_math']:6d}  {src[:70]} | " + " ".join(f"{o}:{n // 1000000}M" for o, n in others))

if __name__ == "__main__":
    main()
