"""Closed-loop full batch: dense Cholesky vs PCG fallback.  Usage: python scripts/loop_timing.py [n_poses]"""
import json, sys
sys.path.insert(0, '.')
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
print(json.dumps(bench.bench_loop_closure(n_poses=n), indent=1))
