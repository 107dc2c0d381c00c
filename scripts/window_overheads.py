"""Where a sequential sliding window's wall time goes on the host side of the C ABI (configs 1/2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ceres_slam_b200 import synthetic as syn, initial_guess as ig
from ceres_slam_b200.problem import BAProblem

tr = syn.add_sun(syn.make_track(60, 15, 10, seed=42, per_obs_W=True, pix_sigma=0.25), sigma_deg=1.0)
w = syn.window_of(tr, 20, 22)
prior = (0, w["poses_gt"][0].copy(), np.eye(6) * 1e3)


def timeit(name, fn, reps=30):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print(f"{name:50s} {1e3 * (time.perf_counter() - t0) / reps:8.3f} ms")


rng = ig.state_ranges(tr["obs_cam"], tr["n_poses"])
pt = tr["obs_pt"].astype(np.int64)
kp, kc = ig.match_pair(pt[rng[20]:rng[21]], pt[rng[21]:rng[22]])
p0 = [ig.triangulate(tr["cam"], tr["uvd"][rng[20]:rng[21]][kp])]
p1 = [ig.triangulate(tr["cam"], tr["uvd"][rng[21]:rng[22]][kc])]
timeit("cslam_ransac_align, 1 pair", lambda: ig.ransac_align(p0, p1, tr["cam"]))


def build(**kw):
    return syn.build_problem(w, sun=True, prior=prior, huber=1.0, hold_first=False, max_num_iterations=10, **kw)


timeit("problem create + blocks (no solve)", lambda: build()[0].close())


def solve(**kw):
    p, _, _ = build(**kw)
    p.solve()
    p.close()


def solve_cov(**kw):
    p, _, _ = build(**kw)
    p.solve()
    p.covariance_block(1)
    p.close()


timeit("create + solve, LM (window kernel)", lambda: solve())
timeit("create + solve, LM (generic engine, window_path=1)", lambda: solve(window_path=1))
timeit("create + solve, SUBSPACE_DOGLEG", lambda: solve(trust_region_strategy=1, dogleg_type=1))
timeit("create + solve + covariance_block, LM", lambda: solve_cov())
timeit("create + solve + covariance_block, DOGLEG", lambda: solve_cov(trust_region_strategy=1, dogleg_type=1))
