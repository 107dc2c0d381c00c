"""Config 3 (dataset_ba_phong joint solve, 2k poses x 200k vertices): per-kernel-class times of one
LM iteration on the GPU.  Usage: python scripts/c3_timing.py [n_poses] [iters]"""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from ceres_slam_b200 import synthetic as syn

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
t0 = time.perf_counter()
tr = syn.add_phong(syn.make_track(n_poses, 100, 10, seed=42), shared_textures=True)
print("generate %.1f s: %d poses, %d vertices, %d obs" % (time.perf_counter() - t0, tr["n_poses"], tr["n_points"], tr["obs_cam"].size), file=sys.stderr)
FIXED = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
p, st = syn.build_phong_problem(tr, bounds=True, max_num_iterations=iters, profile_kernels=1, **FIXED)
t0 = time.perf_counter(); p.upload(); t1 = time.perf_counter()
print("upload %.1f ms" % ((t1 - t0) * 1e3), file=sys.stderr)
p.lm_begin()
p.lm_iterate(2, True)
p.reset_profile()
t0 = time.perf_counter(); s = p.lm_iterate(iters, True); t1 = time.perf_counter()
print("LM %d iters: wall %.2f ms/iter, device %.2f ms total" % (iters, (t1 - t0) * 1e3 / iters, s.device_ms), file=sys.stderr)
for k, (ms, n) in p.profile().items():
    if n:
        print("  %-12s %8.3f ms / %3d launches-groups = %.3f ms" % (k, ms, n, ms / n), file=sys.stderr)
print(np.array2string(p.iteration_log()[:, [0, 1, 5, 6, 9]], precision=5), file=sys.stderr)
