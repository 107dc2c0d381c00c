"""RANSAC front end: one batched launch over the pose pairs of a 1 k-pose track (bench.py's figure)."""
import sys
sys.path.insert(0, '.')
import bench
print(bench.bench_ransac_front_end(), file=sys.stderr)
