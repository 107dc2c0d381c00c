"""The ragged 5 k-pose track of bench.py's c5_ragged alone, a few LM iterations (for ncu launch lists of K3e).
Usage: python scripts/wband_timing.py [n_poses] [iters]"""
import os, sys
sys.path.insert(0, '.')
os.environ.setdefault("CSLAM_DEBUG_SOLVER", "1")
import bench
from ceres_slam_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fixed = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
tr = syn.make_track(n, 100, 10, seed=42, ragged=dict(mean=8.0, max=30, drop=0.1))
p, _, _ = syn.build_problem(tr, max_num_iterations=10 ** 6, profile_kernels=1, **dict(bench.LM_EXACT, **fixed))
p.upload()
p.lm_begin()
p.lm_iterate(1, ignore_convergence=True)
p.reset_profile()
s0 = p.lm_iterate(0, ignore_convergence=True).device_ms
s = p.lm_iterate(iters, ignore_convergence=True)
print("ms per LM iteration", (s.device_ms - s0) / iters, {k: v[0] / max(1, v[1]) for k, v in p.profile().items() if v[1]})
