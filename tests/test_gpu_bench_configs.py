"""Oracle parity ON THE BENCHMARKED CONFIGURATIONS (not only on small tracks):

* config 5 at 0.05 scale (1000 poses x 100 landmarks/frame x track length 10, ~1 M observations, the same
  generator call bench.py makes) through the paths the headline number takes — structure analysed on
  the device, grouped DMMA Schur kernel, leaves + block-cyclic-reduction separators — against the
  oracle for 5 LM iterations at 1e-6 (cost / radius / accept trajectory, poses, landmarks);
* config 2: windows cut from a real 1000-pose sun track (per-observation covariances, sun blocks with
  Huber loss, pose prior), and the restated dataset_vo_sun driver over the first 120 states of it
  against the oracle's window sequence;
* config 1: the dataset_vo driver over a 100-pose track (~150 landmarks/frame), window 2 and full batch.
"""
import os

import numpy as np
import pytest

import bench
from ceres_slam_b200 import capi, synthetic as syn
from oracle import driver_mirror as dm
from oracle import pybinding as orc
from test_gpu_parity import FIXED, check_lm, _poses_csv, _run_driver
from test_driver_sequences import _poses12

pytestmark = pytest.mark.gpu


def test_c5_scaled_default_paths_match_oracle(product):
    tr = bench.c5_track(0.05)
    assert tr["n_poses"] == 1000 and tr["obs_cam"].size > 900_000
    kw = dict(FIXED, max_num_iterations=5, linear_solver=0)
    os.environ["CSLAM_GPU_STRUCTURE_MIN"] = "0"      # device structure analysis (the default from 2^20 blocks)
    try:
        pg, poses_g, points_g = syn.build_problem(tr, band_separator_solver=2, **kw)
        sg = pg.solve()
    finally:
        del os.environ["CSLAM_GPU_STRUCTURE_MIN"]
    # (nearly) every landmark of this workload falls into a group of identical camera lists: the grouped (DMMA)
    # kernel ran on them, the generic kernel on the few whose list is unique after the visibility filter
    info = pg.analyze()
    assert info["n_landmarks"] > 90_000 and info["n_grouped_landmarks"] >= 0.99 * info["n_landmarks"] and info["n_groups"] > 900
    po, poses_o, points_o = orc.build_problem(tr, num_threads=os.cpu_count() or 8, **kw)
    so = po.solve()
    check_lm((pg, sg, poses_g, points_g), (po, so, poses_o, points_o))
    assert sg.final_cost < 0.05 * sg.initial_cost


def _sun_track_1k():
    return syn.add_sun(syn.make_track(1000, 15, 10, seed=42, per_obs_W=True, pix_sigma=0.25), sigma_deg=1.0)


@pytest.mark.parametrize("k1", [9, 400, 977])
def test_c2_windows_of_a_1k_pose_track_match_oracle(product, k1):
    tr = _sun_track_1k()
    w = syn.window_of(tr, k1, k1 + 2)
    prior = (0, w["poses_gt"][0].copy(), np.eye(6) * 1e3)
    for strat in (0, 1):
        kw = dict(FIXED, max_num_iterations=8, sun=True, prior=prior, huber=1.0, hold_first=False,
                  trust_region_strategy=strat, dogleg_type=1)
        pg, poses_g, points_g = syn.build_problem(w, **kw)
        po, poses_o, points_o = orc.build_problem(w, **kw)
        check_lm((pg, pg.solve(), poses_g, points_g), (po, po.solve(), poses_o, points_o))


def test_c2_driver_sequence_on_the_1k_pose_track(product, tmp_path):
    n = 120
    tr = _sun_track_1k()
    keep = (tr["obs_cam"] >= 9) & (tr["obs_cam"] < 9 + n)
    cut = dict(tr, n_poses=n, obs_cam=(tr["obs_cam"][keep] - 9).astype(np.uint32), obs_pt=tr["obs_pt"][keep].copy(),
               uvd=tr["uvd"][keep].copy(), W=tr["W"][keep].copy(), poses=tr["poses"][9:9 + n].copy(),
               poses_gt=tr["poses_gt"][9:9 + n].copy(), sun_cam=np.arange(n, dtype=np.uint32),
               sun_obs_c=tr["sun_obs_c"][9:9 + n].copy(), sun_ref_g=tr["sun_ref_g"][9:9 + n].copy(),
               sun_W=tr["sun_W"][9:9 + n].copy())
    paths = [os.path.join(tmp_path, f) for f in ("track.csv", "sun_ref.csv", "sun_obs.csv")]
    syn.write_sun_csvs(cut, *paths)
    _run_driver("dataset_vo_sun_b200", paths + ["--window", "2", "--huber-param", "1.0", "--max-iters", "100"], tmp_path)
    Ts = _poses12(_poses_csv(os.path.join(tmp_path, "track_obs_poses.csv"), n))
    W = np.asarray(cut["W"]).reshape(-1, 3, 3)
    cov = np.stack([0.5 * (c + c.T) for c in (np.linalg.inv(w @ w) for w in W)]).reshape(-1, 9)
    sW = cut["sun_W"].reshape(-1, 2, 2)
    sun = dict(dir_g=cut["sun_ref_g"], obs=cut["sun_obs_c"], covars=np.stack([np.linalg.inv(w @ w) for w in sW]).reshape(-1, 4),
               has=np.ones(n, dtype=bool))
    kw = dict(window=2, max_iters=100, dogleg=True)          # the reference's SUBSPACE_DOGLEG default
    p1, c1 = dm.dataset_vo_sun(cut, cov, sun, cut["poses_gt"][0], use_sun=False, **kw)
    p2, _ = dm.dataset_vo_sun(cut, cov, sun, cut["poses_gt"][0], use_sun=True, huber=1.0, poses=p1.copy(), pose_covars=c1.copy(), **kw)
    assert np.abs(Ts - p2).max() <= 1e-6 * np.abs(p2).max(), np.abs(Ts - p2).max()
    assert np.abs(p2[:, :3] - cut["poses_gt"][:, :3]).max() < 0.5


@pytest.mark.parametrize("window", [2, 0])
def test_c1_driver_on_the_100_pose_track(product, tmp_path, window):
    """BASELINE.json config 1: 100 poses, ~150 landmarks per frame; `--window 2` (99 sequential windows) and
    `--window 0` = full batch (dataset_vo.cpp:118-121), driver vs the oracle's sequence."""
    n = 100
    tr = syn.make_track(n + 18, 15, 10, seed=42, pix_sigma=0.25)
    keep = (tr["obs_cam"] >= 9) & (tr["obs_cam"] < 9 + n)
    cut = dict(tr, n_poses=n, obs_cam=(tr["obs_cam"][keep] - 9).astype(np.uint32), obs_pt=tr["obs_pt"][keep].copy(),
               uvd=tr["uvd"][keep].copy(), poses=tr["poses"][9:9 + n].copy(), poses_gt=tr["poses_gt"][9:9 + n].copy())
    csv = os.path.join(tmp_path, "track.csv")
    syn.write_track_csv(cut, csv)
    _run_driver("dataset_vo_b200", [csv, "--window", str(window), "--max-iters", "100"], tmp_path)
    Tg = _poses12(_poses_csv(os.path.join(tmp_path, "track_poses.csv"), n))
    var = 1.0 / np.diag(np.asarray(cut["W"]).reshape(3, 3)) ** 2
    To = dm.dataset_vo(cut, var, cut["poses_gt"][0], window, 100)
    assert np.abs(Tg - To).max() <= 1e-6 * np.abs(To).max(), np.abs(Tg - To).max()
    assert np.abs(To[:, :3] - cut["poses_gt"][:, :3]).max() < 0.3
