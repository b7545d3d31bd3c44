#!/usr/bin/env python
"""Per-kernel summary of an ncu --set full report:  python tools/ncu_summary.py <report.ncu-rep> > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
        ("smsp__cycles_active.avg", "smsp active cycles (avg)"), ("sm__cycles_elapsed.max", "sm cycles elapsed"),
        ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "registers"),
        ("launch__shared_mem_per_block_static", "static smem"), ("launch__shared_mem_per_block_dynamic", "dynamic smem"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("lts__t_bytes.sum", "L2 bytes"),
        ("smsp__sass_inst_executed_op_global_red.sum", "global RED instr"),
        ("l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "RED L1 accesses")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: ncu --set full --clock-control none (per launch; cold-cache, serialised replays)")
    for r in data:
        name = r[ix["Kernel Name"]]
        short = name.split("(")[0].split("::")[-1]
        print(f"\n## {short}")
        for key, label in WANT:
            if key in ix and r[ix[key]] != "":
                print(f"  {label:28s} {r[ix[key]]} {units[ix[key]]}")
        try:
            t = float(r[ix["gpu__time_duration.sum"]])
            tu = units[ix["gpu__time_duration.sum"]]
            t_s = t * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(tu.replace("second", "s").replace("usecond", "us"), 1e-6)
            conv = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            b = sum(float(r[ix[k]]) * conv.get(units[ix[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            print(f"  {'dram traffic':28s} {b / 1e6:.2f} MB  -> {b / t_s / 1e9:.0f} GB/s")
        except Exception as e:  # noqa
            pass


if __name__ == "__main__":
    main()
