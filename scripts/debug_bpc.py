import sys; sys.path.insert(0,'.')
from ceres_slam_b200 import synthetic as syn
tr = syn.make_track(2300, 6, 6, seed=37, ragged=dict(mean=6, max=20, drop=0.1))
p,_,_ = syn.build_problem(tr, max_num_iterations=1)
p.solve()
