#!/usr/bin/env python
"""Instruction mix and the tensor-core / bulk-copy / barrier instructions of the hot kernels, from `cuobjdump -sass`
of the in-tree objects (ceres_slam_b200/csrc/build/*.o, sm_100a).  Usage: python scripts/sass_extract.py > profiles/rNN_sass_extracts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "ceres_slam_b200", "csrc", "build")
KERNELS = [  # (object, substring of the mangled name, label)
    ("kernels_grouped.o", "schur_grouped2_kernelILb0ELi2ELb0E", "schur_grouped2_kernel<false,2,false> (K2)"),
    ("kernels_grouped.o", "schur_grouped2_kernelILb0ELi2ELb1E", "schur_grouped2_kernel<false,2,true> (K2, ragged groups)"),
    ("kernels.o", "schur_wide_kernel", "schur_wide_kernel (K2w, window tile)"),
    ("kernels.o", "schur_wide_produce_kernel", "schur_wide_produce_kernel (K2w, Z per observation)"),
    ("kernels.o", "resjac_kernelILb1E", "resjac_kernel (K1)"),
    ("kernels_wband.o", "wband_panel_kernel", "wband_panel_kernel (K3e)"),
    ("kernels_wband.o", "wband_syrk_kernel", "wband_syrk_kernel (K3e)"),
    ("kernels_wband.o", "wband_backsolve_kernel", "wband_backsolve_kernel (K3e)"),
    ("kernels_dense.o", "dense_syrk_kernel", "dense_syrk_kernel (K3d)"),
    ("kernels_dense.o", "dense_panel_kernel", "dense_panel_kernel (K3d)"),
    ("kernels_band.o", "bcr_odd2_kernelILi54E", "bcr_odd2_kernel<54> (K3b)"),
    ("kernels_band.o", "band_leaf2_kernelILi9E", "band_leaf2_kernel<9> (K3b)"),
]
KINDS = ("DMMA", "UBLKCP", "SYNCS", "MUFU.RSQ64H", "BAR.", "RED.", "REDG", "LDGSTS", "STL", "LDL")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            res[cur].append(line.split("/*", 2)[1].split("*/", 1)[0] + " " + line.split("*/", 1)[1].split("/*")[0].strip())
    return res


def main():
    print("# SASS extracts (cuobjdump -sass of the in-tree objects, sm_100a), scripts/sass_extract.py")
    print("# per kernel: instruction mix (mnemonic counts) and the first lines of each tensor-core / bulk-copy / barrier kind")
    cache = {}
    for obj, key, label in KERNELS:
        fns = cache.setdefault(obj, functions(obj))
        name = next((n for n in fns if key in n), None)
        print(f"\n## {label}, {obj}")
        if not name:
            print("(not found)")
            continue
        ins = fns[name]
        mn = collections.Counter()
        for l in ins:
            toks = [t for t in l.split()[1:] if not t.startswith("@")]
            if toks:
                mn[toks[0].split(".")[0]] += 1
        print("Function :", name[:150])
        print("instructions:", len(ins))
        print("mix:", ", ".join(f"{k} {v}" for k, v in mn.most_common(18)))
        print("counts:", ", ".join(f"{k} {sum(1 for l in ins if k in l)}" for k in KINDS))
        for k in ("DMMA", "UBLKCP", "SYNCS", "MUFU.RSQ64H"):
            for l in [l for l in ins if k in l][:3]:
                print("/*%s*/ %s" % (l.split()[0], " ".join(l.split()[1:])))


if __name__ == "__main__":
    main()
