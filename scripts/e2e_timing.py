import sys, time, os
sys.path.insert(0,'.')
import bench
from ceres_slam_b200 import synthetic as syn
tr = bench.c5_track(1.0)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    p, poses, points = syn.build_problem(tr, max_num_iterations=10, **bench.LM_EXACT)
    t0 = time.perf_counter(); s = p.solve(); t1 = time.perf_counter()
    print("solve wall %.1f ms, device LM %.1f ms, iters %d" % ((t1-t0)*1e3, s.device_ms, s.num_iterations), file=sys.stderr)
    p.close()
