"""Wide-band Cholesky (K3e) vs the other exact solvers on the same problem (debug aid): iterates and solve times.
Usage: python scripts/debug_wband.py [n_poses] [per_frame] [max_len]"""
import os, sys
sys.path.insert(0, '.')
os.environ["CSLAM_DEBUG_SOLVER"] = "1"
import numpy as np
from ceres_slam_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
per = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lmax = int(sys.argv[3]) if len(sys.argv) > 3 else 20
tr = syn.make_track(n, per, 6, seed=38, ragged=dict(mean=7, max=lmax, drop=0.1))
ref = None
for name, kw in (("wband", dict(bandpc_solver=2)), ("other exact", dict(bandpc_solver=1)), ("pcg 1e-15", dict(bandpc_solver=-1, dense_solver=-1))):
    p, poses, points = syn.build_problem(tr, max_num_iterations=4, window_path=1, profile_kernels=1, function_tolerance=0.0,
                                         parameter_tolerance=0.0, gradient_tolerance=0.0, **kw)
    s = p.solve()
    prof = p.profile() if hasattr(p, "profile") else {}
    print(name, "term", s.termination_type, s.termination_reason, "iters", s.num_iterations, "final cost %.10e" % s.final_cost,
          "linear_solve", prof.get("linear_solve"))
    print(p.iteration_log()[:, [1, 4, 7, 8, 9]])
    if ref is None:
        ref = poses.copy()
    else:
        print("  poses vs wband: %.3e" % (np.abs(poses - ref).max() / np.abs(ref).max()))
