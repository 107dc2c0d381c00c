import numpy as np, sys
sys.path.insert(0, '.')
from ceres_slam_b200 import synthetic as syn
tr = syn.add_sun(syn.make_track(300, 40, 8, seed=77))
kw = dict(max_num_iterations=6, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, sun=True)
for ls in (0, 1):
    res = []
    for rep in range(4):
        p, poses, points = syn.build_problem(tr, linear_solver=ls, **kw)
        s = p.solve()
        res.append((poses.copy(), points.copy(), p.iteration_log()))
    for rep in range(1, 4):
        print("ls", ls, "rep", rep, "poses equal", np.array_equal(res[0][0], res[rep][0]), "max diff", np.abs(res[0][0]-res[rep][0]).max(),
              "points", np.abs(res[0][1]-res[rep][1]).max(), "cg", res[rep][2][:,7])
