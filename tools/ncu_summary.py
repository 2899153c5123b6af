#!/usr/bin/env python3
"""ncu_summary.py -- condense an .ncu-rep (read here with `ncu -i`, no GPU needed) into the small JSON
summaries kept under profiles/: one entry per captured launch with the counters the roofline argument uses.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.json --command "<the ncu command>"
"""
import argparse
import csv
import io
import json
import subprocess

KEEP = ["dram__bytes.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--command", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        d = dict(zip(head, r))
        e = {"Kernel Name": d.get("Kernel Name")}
        for k in KEEP:
            if k in d:
                e[k] = ("%s %s" % (d[k], units[head.index(k)])).strip()
        launches.append(e)
    json.dump({"command": a.command, "launches": launches}, open(a.out, "w"), indent=1)
    print("wrote %s: %d launches" % (a.out, len(launches)))


if __name__ == "__main__":
    main()
