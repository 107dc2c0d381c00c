import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ceres_slam_b200 import synthetic as syn
from oracle import pybinding as orc
from ceres_slam_b200.problem import solve_batch
import test_gpu_parity as T
np.set_printoptions(linewidth=200, precision=10)
cases = T._window_cases()
kw = dict(T.FIXED, max_num_iterations=6)
gpu = [syn.build_problem(w, **dict(kw, **extra)) for w, extra in cases]
sums = solve_batch([g[0] for g in gpu])
for i, ((w, extra), (pg, poses_g, points_g), sg) in enumerate(zip(cases, gpu, sums)):
    po, poses_o, points_o = orc.build_problem(w, **dict(kw, **extra))
    so = po.solve()
    lg, lo = pg.iteration_log(), po.iteration_log()
    ok = lg.shape == lo.shape and np.allclose(lg[:, 1], lo[:, 1], rtol=1e-6) and sg.num_successful_steps == so.num_successful_steps
    print(i, w["n_poses"], w["n_points"], "OK" if ok else "MISMATCH", sg.num_successful_steps, so.num_successful_steps, sg.num_iterations, so.num_iterations)
    if not ok:
        print("gpu\n", lg[:, [0,1,4,5,6,8,9]]); print("oracle\n", lo[:, [0,1,4,5,6,8,9]])
