"""TEST INFRASTRUCTURE — the reference drivers' window loops over the CPU oracle.

`dataset_vo(...)` and `dataset_vo_sun(...)` restate what tests/dataset_vo.cpp:87-138 and
tests/dataset_vo_sun.cpp:189-323 do around `solveWindow`: per window the RANSAC initial guess
(compute_initial_guess), residual blocks for the points it initialised, the solve, (sun: the
marginal covariance of the window's second pose, which becomes the prior of the next window),
`reset_points()`.  Every numerical step runs on the oracle (`oracle/pybinding.py`); the covariance
is `(J^T J)^-1` of the oracle's loss-corrected tangent-space Jacobians through a sparse LU — what
`ceres::Covariance` computes (dataset_vo_sun.cpp:159-183).

Used by the parity tests (the restated C++ drivers over the CUDA library must reproduce the SAME
window sequence) and by bench.py's CPU baselines for configs 1 and 2.
"""
import numpy as np

from ceres_slam_b200 import initial_guess as ig
from oracle import pybinding as orc


def _inverse_sqrt(cov):
    """Symmetric inverse square root (SelfAdjointEigenSolver::operatorInverseSqrt, dataset_vo.cpp:29-32)."""
    cov = np.asarray(cov, dtype=np.float64)
    w, V = np.linalg.eigh(0.5 * (cov + cov.T))
    return (V * (1.0 / np.sqrt(w))) @ V.T


def _window_blocks(track, initialized, k1, k2):
    rng = ig.state_ranges(track["obs_cam"], track["n_poses"])
    idx = np.arange(rng[k1], rng[k2])
    return idx[initialized[track["obs_pt"][idx]]]


def dataset_vo(track, var, first_pose, window, max_iters, problem_cls=None, guess=None, on_window=None):
    """track: obs_cam / obs_pt / uvd / cam / n_poses / n_points (rows grouped by state, as in the CSV).
    Returns the final poses (n, 12)."""
    problem_cls = problem_cls or orc.OracleProblem
    guess = guess or orc.compute_initial_guess
    n = track["n_poses"]
    poses = np.tile(np.asarray(first_pose, dtype=np.float64).reshape(1, 12), (n, 1))
    points = np.zeros((track["n_points"], 3))
    init = np.zeros(track["n_points"], dtype=bool)
    W = _inverse_sqrt(np.diag(var)).reshape(9)
    c = track["cam"]
    if window == 0 or window > n:
        window = n
    for k1 in range(0, n - window + 1):
        k2 = k1 + window
        guess(track, poses, points, init, k1=k1, k2=k2)
        sel = _window_blocks(track, init, k1, k2)
        ids, pt_local = np.unique(track["obs_pt"][sel], return_inverse=True)
        p = problem_cls(max_num_iterations=max_iters, use_nonmonotonic_steps=1)
        p.set_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["b"])
        const = np.zeros(window, dtype=np.uint8)
        const[0] = 1                                                         # dataset_vo.cpp:62
        pw = p.set_poses(poses[k1:k2].copy(), const)
        xw = p.set_points(points[ids].copy() if ids.size else np.zeros((1, 3)))
        if sel.size:
            p.add_stereo((track["obs_cam"][sel] - k1).astype(np.uint32), pt_local.astype(np.uint32), track["uvd"][sel], W)
        s = p.solve()
        poses[k1:k2] = pw
        if ids.size:
            points[ids] = xw
        if on_window:
            on_window(k1, s)
        p.close()
        init[:] = False                                                      # reset_points, :130
    return poses


def tangent_covariance_block(ev, cams, pts, n_poses, n_points, constant, cam, sun_cam=None, prior_cam=None):
    """[(J^T J)^-1]_{cam,cam} from per-block tangent Jacobians (stereo, sun, prior), sparse LU."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    free = np.flatnonzero(np.asarray(constant) == 0)
    col_of = -np.ones(n_poses, dtype=np.int64)
    col_of[free] = 6 * np.arange(free.size)
    off_l = 6 * free.size
    rows, cols, vals = [], [], []
    r0 = 0

    def add(c0, J):
        rr, cc = np.meshgrid(np.arange(J.shape[0]), np.arange(J.shape[1]), indexing="ij")
        rows.append((r0 + rr).ravel()); cols.append((c0 + cc).ravel()); vals.append(J.ravel())

    for i in range(len(cams)):
        k, j = int(cams[i]), int(pts[i])
        if col_of[k] >= 0:
            add(col_of[k], ev["Jpose_stereo"][i])
        add(off_l + 3 * j, ev["Jpoint_stereo"][i])
        r0 += 3
    for i, k in enumerate(sun_cam if sun_cam is not None else []):
        if col_of[int(k)] >= 0:
            add(col_of[int(k)], ev["J_sun"][i])
        r0 += 2
    for i, k in enumerate(prior_cam if prior_cam is not None else []):
        if col_of[int(k)] >= 0:
            add(col_of[int(k)], ev["J_prior"][i])
        r0 += 6
    ncol = off_l + 3 * n_points
    J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, ncol))
    H = (J.T @ J).tocsc()
    lu = spla.splu(H)
    E = np.zeros((ncol, 6))
    E[col_of[cam]:col_of[cam] + 6] = np.eye(6)
    return lu.solve(E)[col_of[cam]:col_of[cam] + 6]


def dataset_vo_sun(track, obs_covars, sun, first_pose, window, max_iters, use_sun, huber=0.0, az=1000.0, zen=1000.0,
                   dogleg=True, covariance=True, poses=None, pose_covars=None, on_window=None):
    """One pass of dataset_vo_sun (tests/dataset_vo_sun.cpp:271-311).  `obs_covars` (n_obs, 9) as in the CSV,
    `sun` = dict(dir_g (n, 3), obs (n, 3), covars (n, 4), has (n,)).  `poses` / `pose_covars` carry over
    from the first pass to the second like the dataset object's fields do.  Returns (poses, pose_covars)."""
    n = track["n_poses"]
    if poses is None:
        poses = np.tile(np.asarray(first_pose, dtype=np.float64).reshape(1, 12), (n, 1))
    if pose_covars is None:
        pose_covars = np.tile((np.eye(6) * 1e-12).reshape(1, 36), (n, 1))    # dataset_problem_sun.cpp:80
    points = np.zeros((track["n_points"], 3))
    init = np.zeros(track["n_points"], dtype=bool)
    c = track["cam"]
    if window == 0 or window > n:
        window = n
    for k1 in range(0, n - window + 1):
        k2 = k1 + window
        st = orc.compute_initial_guess(track, poses, points, init, k1=k1, k2=k2)
        if (st["n_inliers"] < 3).any():                                       # dataset_problem_sun.cpp:323-326
            poses[k2 - 1] = poses[k1]
            pose_covars[k2 - 1] = pose_covars[k1]
            init[:] = False
            continue
        sel = _window_blocks(track, init, k1, k2)
        ids, pt_local = np.unique(track["obs_pt"][sel], return_inverse=True)
        p = orc.OracleProblem(max_num_iterations=max_iters, use_nonmonotonic_steps=1,
                              trust_region_strategy=1 if dogleg else 0, dogleg_type=1)
        p.set_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["b"])
        pw = p.set_poses(poses[k1:k2].copy(), np.zeros(window, dtype=np.uint8))
        xw = p.set_points(points[ids].copy())
        # the reference indexes the per-observation covariances by POINT id (dataset_vo_sun.cpp:58)
        ci = np.where(track["obs_pt"][sel] < obs_covars.shape[0], track["obs_pt"][sel], sel)
        Wst = np.stack([_inverse_sqrt(obs_covars[i].reshape(3, 3)).reshape(9) for i in ci])
        cams = (track["obs_cam"][sel] - k1).astype(np.uint32)
        p.add_stereo(cams, pt_local.astype(np.uint32), track["uvd"][sel], Wst)
        sun_cam = []
        if use_sun:
            sun_cam = [k - k1 for k in range(k1, k2) if sun["has"][k]]
            if sun_cam:
                ks = np.array(sun_cam) + k1
                W2 = np.stack([_inverse_sqrt(sun["covars"][k].reshape(2, 2)).reshape(4) for k in ks])
                p.add_sun(np.array(sun_cam, dtype=np.uint32), sun["obs"][ks], sun["dir_g"][ks], W2, az, zen, huber)
        p.add_pose_prior(0, poses[k1].copy(), _inverse_sqrt(pose_covars[k1].reshape(6, 6)).reshape(36))
        s = p.solve()
        poses[k1:k2] = pw
        points[ids] = xw
        if covariance and k1 + 1 < n:
            ev = p.evaluate(apply_loss=True)
            try:
                cov = tangent_covariance_block(ev, cams, pt_local, window, ids.size, np.zeros(window), 1,
                                               sun_cam=sun_cam, prior_cam=[0])
                pose_covars[k1 + 1] = cov.reshape(36)
            except RuntimeError:
                pose_covars[k1 + 1] = pose_covars[k1]
        if on_window:
            on_window(k1, s)
        p.close()
        init[:] = False
    return poses, pose_covars
