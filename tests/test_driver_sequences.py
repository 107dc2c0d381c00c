"""The restated C++ drivers (host/dataset_vo_b200, host/dataset_vo_sun_b200) over the CUDA library
against the SAME window sequence run on the CPU oracle (oracle/driver_mirror.py): RANSAC initial
guess -> window problem -> solve (-> covariance -> prior of the next window) -> reset, window after
window, every window starting from what the previous one left.  Compared in memory / at the
drivers' full-precision output (17 digits), not at the reference's 4-digit CSV precision.

Tolerance: 1e-6 relative to the pose entries (BASELINE.json: poses after the solve within 1e-6);
the windows run to Ceres' default tolerances, so an iterate is only defined to ~function_tolerance
— each window is therefore also checked to end with the same iteration count.
"""
import os

import numpy as np
import pytest

from ceres_slam_b200 import synthetic as syn
from oracle import driver_mirror as dm
from test_gpu_parity import _poses_csv, _run_driver, _steady_track


def _poses12(T):
    return np.concatenate([T[:, :3, 3], T[:, :3, :3].reshape(-1, 9)], axis=1)


def _first_pose(tr):
    return tr["poses_gt"][0]


def test_oracle_window_sequence_tracks_ground_truth():
    """CPU: the mirror itself is a sane restatement (chained 2-pose windows stay on the track)."""
    tr = _steady_track(12, seed=17)
    var = 1.0 / np.diag(np.asarray(tr["W"]).reshape(3, 3)) ** 2
    iters = []
    poses = dm.dataset_vo(tr, var, _first_pose(tr), 2, 100, on_window=lambda k1, s: iters.append(s.num_iterations))
    assert len(iters) == 11
    assert np.abs(poses[:, :3] - tr["poses_gt"][:, :3]).max() < 0.1
    assert np.abs(poses[:, 3:] - tr["poses_gt"][:, 3:]).max() < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("window", [2, 4, 0])
def test_dataset_vo_driver_matches_oracle_sequence(product, tmp_path, window):
    tr = _steady_track(30, seed=17)
    csv = os.path.join(tmp_path, "track.csv")
    syn.write_track_csv(tr, csv)
    text = _run_driver("dataset_vo_b200", [csv, "--window", str(window), "--max-iters", "100"], tmp_path)
    Tg = _poses12(_poses_csv(os.path.join(tmp_path, "track_poses.csv"), 30))
    var = 1.0 / np.diag(np.asarray(tr["W"]).reshape(3, 3)) ** 2
    its = []
    To = dm.dataset_vo(tr, var, _first_pose(tr), window, 100, on_window=lambda k1, s: its.append(s.num_iterations))
    its_g = [int(l.split("Iterations:")[1].split(",")[0]) for l in text.splitlines() if "Iterations:" in l]
    assert len(its_g) == len(its)
    assert sum(a != b for a, b in zip(its, its_g)) <= max(1, len(its) // 10), (its, its_g)
    assert np.abs(Tg - To).max() <= 1e-6 * np.abs(To).max(), np.abs(Tg - To).max()


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["lm", "dogleg"])
def test_dataset_vo_sun_driver_matches_oracle_sequence(product, tmp_path, strategy):
    """Both passes of dataset_vo_sun incl. the covariance chain: pass 1 (VO), pass 2 (sun blocks with Huber
    loss, starting from pass 1's poses and covariances as the dataset object carries them)."""
    n = 25
    tr = syn.add_sun(_steady_track(n, seed=23, per_obs_W=True), sigma_deg=1.0)
    paths = [os.path.join(tmp_path, f) for f in ("track.csv", "sun_ref.csv", "sun_obs.csv")]
    syn.write_sun_csvs(tr, *paths)
    _run_driver("dataset_vo_sun_b200", paths + ["--window", "2", "--huber-param", "1.0", "--max-iters", "100",
                                                "--strategy", strategy], tmp_path)
    Tv = _poses12(_poses_csv(os.path.join(tmp_path, "track_poses.csv"), n))
    Ts = _poses12(_poses_csv(os.path.join(tmp_path, "track_obs_poses.csv"), n))
    # the mirror reads what the driver reads: covariances are the CSV's (inverse squares of the stiffness)
    W = np.asarray(tr["W"]).reshape(-1, 3, 3)
    cov = np.stack([0.5 * (c + c.T) for c in (np.linalg.inv(w @ w) for w in W)]).reshape(-1, 9)
    sW = tr["sun_W"].reshape(-1, 2, 2)
    sun = dict(dir_g=tr["sun_ref_g"], obs=tr["sun_obs_c"], covars=np.stack([np.linalg.inv(w @ w) for w in sW]).reshape(-1, 4),
               has=np.ones(n, dtype=bool))
    kw = dict(window=2, max_iters=100, dogleg=(strategy == "dogleg"))
    p1, c1 = dm.dataset_vo_sun(tr, cov, sun, _first_pose(tr), use_sun=False, **kw)
    assert np.abs(Tv - p1).max() <= 1e-6 * np.abs(p1).max(), np.abs(Tv - p1).max()
    p2, _ = dm.dataset_vo_sun(tr, cov, sun, _first_pose(tr), use_sun=True, huber=1.0, poses=p1.copy(), pose_covars=c1.copy(), **kw)
    assert np.abs(Ts - p2).max() <= 1e-6 * np.abs(p2).max(), np.abs(Ts - p2).max()
