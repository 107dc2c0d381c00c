"""K3c (bandpc_solver = 1) on the ragged 5 k-pose track: LM log and CG iteration counts (debug aid).
Env: CSLAM_NO_WIDE, CSLAM_EXACT_MIN_GROUP, CSLAM_BPC_DEBUG."""
import sys
sys.path.insert(0, '.')
import numpy as np
from ceres_slam_b200 import synthetic as syn
tr = syn.make_track(5000, 100, 10, seed=42, ragged=dict(mean=8.0, max=30, drop=0.1))
p, poses, points = syn.build_problem(tr, max_num_iterations=int(sys.argv[1]) if len(sys.argv) > 1 else 3, bandpc_solver=1, function_tolerance=0.0, parameter_tolerance=0.0,
                                     gradient_tolerance=0.0)
s = p.solve()
print("term", s.termination_type, s.termination_reason, "iters", s.num_iterations)
print(p.iteration_log()[:, [1, 4, 6, 7, 8, 9]])
