#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep   > profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64",
        "sm__pipe_fp64_cycles_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "sm__pipe_tensor_cycles_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block",
        "smsp__warp_issue_stalled", "sm__inst_executed.sum", "l1tex__data_bank_conflicts")


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("unnamed>::", "")
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)")
    print(f"{'kernel':48s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:48]:48s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / v[0] / 1e3:10.2f} {v[1] / tot:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## {name.split('(')[0]}")
        for i, h in enumerate(hdr):
            if any(h.startswith(w) for w in WANT):
                print(f"{h:80s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
