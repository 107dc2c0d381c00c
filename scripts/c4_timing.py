"""Host phases of cslam_solve_batch on the C4 batch (256 windows), LM and SUBSPACE_DOGLEG (CSLAM_WINDOW_TIMING=1)."""
import os, sys, time
os.environ["CSLAM_WINDOW_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ceres_slam_b200 import synthetic as syn
from ceres_slam_b200.problem import solve_batch

tracks = [syn.make_track(100, 15, 10, seed=42 + t) for t in range(3)]
wins = [syn.window_of(tr, k1, k1 + 2) for tr in tracks for k1 in range(5, 95)][:256]
for extra in (dict(), dict(trust_region_strategy=1, dogleg_type=1)):
    kw = dict(max_num_iterations=6, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, **extra)
    for rep in range(4):
        probs = [syn.build_problem(w, **kw)[0] for w in wins]
        t0 = time.perf_counter()
        sums = solve_batch(probs)
        wall = time.perf_counter() - t0
        print(f"{extra or 'LM'} rep {rep}: wall {wall * 1e3:.3f} ms, kernel {sums[0].device_ms:.3f} ms, "
              f"iterations {sum(s.num_iterations for s in sums)}", flush=True)
        for p in probs:
            p.close()
