#!/usr/bin/env python
"""Print selected metrics (substring match) from an .ncu-rep, one kernel per block."""
import csv, subprocess, sys
rep, pats = sys.argv[1], sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("## " + r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            print(f"  {h:95s} {r[i]:>16s} {units[i]}")
