import numpy as np, sys
sys.path.insert(0, '.')
from ceres_slam_b200 import synthetic as syn
from oracle import pybinding as orc
np.set_printoptions(linewidth=200, precision=10)
FIXED = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
tr = syn.make_track(80, 15, 8, seed=9)
for pre in (0, 1):
    for iters in (1, 2, 3, 5):
        kw = dict(FIXED, max_num_iterations=iters, linear_solver=1, preconditioner=pre)
        pg, poses_g, points_g = syn.build_problem(tr, **kw)
        po, poses_o, points_o = orc.build_problem(tr, **kw)
        sg, so = pg.solve(), po.solve()
        lg, lo = pg.iteration_log(), po.iteration_log()
        print("pre", pre, "iters", iters, "pose err", np.abs(poses_g-poses_o).max(), "pt err", np.abs(points_g-points_o).max())
        print(" cost g", lg[:,1]); print(" cost o", lo[:,1]); print(" cg", lg[:,7], lo[:,7]); print(" rho g", lg[:,5]); print(" rho o", lo[:,5])
