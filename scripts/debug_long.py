import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ceres_slam_b200 import synthetic as syn
import oracle.pybinding as orc

tr = syn.add_phong(syn.make_track(400, 2, 12, seed=9, ragged=dict(mean=30, max=70, drop=0.05)), shared_textures=True)
for dt, radius in ((1, 1e4), (0, 1e4), (1, 3.0), (0, 0.5)):
    kw = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, max_num_iterations=4,
              trust_region_strategy=1, dogleg_type=dt, initial_trust_region_radius=radius)
    pg, sg = syn.build_phong_problem(tr, bounds=True, **kw)
    po, so = orc.build_phong_problem(tr, bounds=True, num_threads=16, **kw)
    pg.solve(); po.solve()
    lg, lo = pg.iteration_log(), po.iteration_log()
    n = min(len(lg), len(lo))
    print(dt, radius, "cost rel", np.abs(lg[:n, 1] / lo[:n, 1] - 1), "step rel", np.abs(lg[1:n, 4] / lo[1:n, 4] - 1), "radius", lg[:n, 6], lo[:n, 6], flush=True)
