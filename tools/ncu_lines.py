#!/usr/bin/env python
"""Per-CUDA-source-line summary of an ncu report (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]
Prints samples / executed warp-instructions / dominant stall reasons per source line."""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, lines = "", None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ix = {}
            for i, n in enumerate(hdr):
                ix.setdefault(n, i)
        elif hdr and len(r) == len(hdr) and r[0] != "":
            lines.append((fname, r))
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    S = lambda r, n: int(r[ix[n]] or 0)
    tot_s = sum(S(r, "# Samples") for _, r in lines)
    tot_i = sum(S(r, "Instructions Executed") for _, r in lines)
    print(f"kernel {pat}: samples {tot_s}  warp-instructions {tot_i}")
    agg = {}
    for n in stalls:
        agg[n] = sum(S(r, n) for _, r in lines)
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    lines.sort(key=lambda fr: -S(fr[1], "# Samples"))
    for f, r in lines[:top]:
        st = sorted(((S(r, n), n[6:]) for n in stalls), reverse=True)[:3]
        print(f"{S(r, '# Samples'):7d} {100.0 * S(r, '# Samples') / max(tot_s, 1):5.1f}% {S(r, 'Instructions Executed'):10d} "
              f"{f}:{r[0]:>4} {r[1].strip()[:110]}   {[f'{n}:{v}' for v, n in st if v]}")


if __name__ == "__main__":
    main()
