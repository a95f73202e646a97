#!/usr/bin/env python
"""Per-kernel share of a step from an ncu launch list (gpu__time_duration.sum pass): python tools/launch_shares.py launches.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(u, 1)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += ns
ours = {k: v for k, v in tot.items() if k.startswith(("abi::", "k_")) or "abi::" in k}
s = sum(v[1] for v in ours.values())
print(f"# launches of this library in the capture: {sum(v[0] for v in ours.values())}, total device time {s/1e6:.3f} ms (cold-cache, serialised)")
print(f"{'kernel':90s} {'launches':>8s} {'total ms':>10s} {'ms/launch':>10s} {'share':>7s}")
for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} {n:8d} {t/1e6:10.3f} {t/1e6/n:10.3f} {100*t/s:6.1f}%")
other = {k: v for k, v in tot.items() if k not in ours}
print(f"# other kernels (torch set-up: random fills, copies): {sum(v[0] for v in other.values())} launches, {sum(v[1] for v in other.values())/1e6:.3f} ms")
