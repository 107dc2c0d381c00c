"""Config 5 per-kernel-class step profile.  Usage: python scripts/c5_timing.py [scale] [key=value ...]"""
import sys, time
sys.path.insert(0, '.')
import bench
from ceres_slam_b200 import synthetic as syn
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
extra = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
tr = bench.c5_track(scale)
p, _, _ = syn.build_problem(tr, max_num_iterations=10 ** 6, profile_kernels=1, **dict(bench.LM_EXACT, **extra))
p.upload(); p.lm_begin(); p.lm_iterate(3, True); p.reset_profile()
t0 = time.perf_counter(); s0 = p.lm_iterate(0, True).device_ms; s = p.lm_iterate(6, True)
print("opts", extra, "ms/iter %.3f" % ((s.device_ms - s0) / 6), file=sys.stderr)
for k, (ms, n) in p.profile().items():
    if n:
        print("  %-12s %.3f ms" % (k, ms / n), file=sys.stderr)
log = p.iteration_log()
print("  cost", log[0, 1], "->", log[-1, 1], file=sys.stderr)
