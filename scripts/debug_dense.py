"""Dense reduced solve vs the banded one on the same problem (debug aid)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from ceres_slam_b200 import synthetic as syn
shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (40, 8, 5)
tr = syn.make_track(*shape, seed=31)
for ds in (1, 0):
    p, poses, points = syn.build_problem(tr, dense_solver=ds, max_num_iterations=4, window_path=1)
    s = p.solve()
    print("dense_solver", ds, "term", s.termination_type, s.termination_reason, "iters", s.num_iterations)
    print(p.iteration_log()[:, [1, 4, 7, 8, 9]])
