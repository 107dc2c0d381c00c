#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of one kernel: joins ncu's SASS page
(`ncu -i rep --page source --csv`) with `nvdisasm --print-line-info` by instruction order.

  python scripts/ncu_lines.py <rep> <kernel-regex> <object.o> <mangled-substring> [top]
"""
import collections, csv, re, subprocess, sys, tempfile, os
rep, kre, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
inst = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break  # next kernel of the report
    if len(r) > iX and r[iN].isdigit():
        inst.append((r[iS].strip(), int(r[iN] or 0), int(r[iX] or 0)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for l in sass:
    if l.startswith(".text."):
        on = mangled in l
        continue
    if not on:
        continue
    m = re.search(r'//## File ".*", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
if len(lines) != len(inst):
    print("warning: %d SASS instructions in cubin vs %d in report" % (len(lines), len(inst)))
agg = collections.defaultdict(lambda: [0, 0])
for (src, ns, nx), ln in zip(inst, lines):
    agg[ln][0] += nx; agg[ln][1] += ns
tot_x = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot_x}, samples {tot_s}")
for ln, (nx, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"line {ln}: inst {nx} ({100*nx/max(tot_x,1):.1f}%)  samples {ns} ({100*ns/max(tot_s,1):.1f}%)")
