"""Ragged stereo tracks (variable length, drop-outs) vs the regular C5-style track at the same size.
Usage: python scripts/ragged_timing.py [n_poses] [mean] [max] [drop]"""
import json, sys
sys.path.insert(0, '.')
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
kw = {}
if len(sys.argv) > 2: kw["mean"] = float(sys.argv[2])
if len(sys.argv) > 3: kw["lmax"] = int(sys.argv[3])
if len(sys.argv) > 4: kw["drop"] = float(sys.argv[4])
print(json.dumps(bench.bench_ragged(n_poses=n, **kw), indent=1))
