#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean/total time, share.
   python tools/launch_summary.py gpurun_out/r01_launches.csv [substring filter ...]"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1].strip()
        full = r[ix["Kernel Name"]]
        if full.startswith("<unnamed>::") or full.startswith("void <unnamed>::") or "lgs_" in full:
            name = "lgs:" + name
        t = float(r[ix["Metric Value"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ix["Metric Unit"]], 1e-3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    ours = {k: v for k, v in agg.items() if k.startswith("lgs:")}
    tot = sum(v[1] for v in ours.values())
    print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches; library kernels {sum(v[0] for v in ours.values())} launches, {tot:.1f} us")
    print(f"{'kernel':40s} {'launches':>8s} {'mean us':>10s} {'share of lgs time':>18s}")
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[4:]:40s} {n:8d} {t / n:10.2f} {100 * t / tot:17.1f}%")
    other = {k: v for k, v in agg.items() if not k.startswith("lgs:")}
    print("# other (torch set-up / memsets):", {k[:40]: (n, round(t, 1)) for k, (n, t) in list(other.items())[:8]})


if __name__ == "__main__":
    main()
