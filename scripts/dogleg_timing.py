"""C5 at reduced scale with the DOGLEG strategy (also used for ncu captures of the dogleg_* kernels)."""
import sys
sys.path.insert(0, '.')
import bench
from ceres_slam_b200 import synthetic as syn
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
tr = bench.c5_track(scale)
for strat in (1, 0):
    p, _, _ = syn.build_problem(tr, max_num_iterations=10 ** 6, profile_kernels=1, trust_region_strategy=strat,
                                dogleg_type=1, **bench.LM_EXACT)
    p.upload(); p.lm_begin(); p.lm_iterate(3, True); p.reset_profile()
    s0 = p.lm_iterate(0, True).device_ms
    s = p.lm_iterate(6, True)
    print("strategy", strat, "ms/iter %.3f" % ((s.device_ms - s0) / 6), {k: round(v[0] / v[1], 3) for k, v in p.profile().items() if v[1]},
          "cost", p.iteration_log()[-1, 1], file=sys.stderr)
    p.close()
