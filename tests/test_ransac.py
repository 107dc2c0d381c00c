"""Front end: the 3-point RANSAC of PointCloudAligner and compute_initial_guess.
CPU part: the restated random draws against the std:: classes of this image's compiler, the
alignment against LAPACK's SVD.  GPU part: the batched kernel against the oracle."""
import numpy as np
import pytest

from ceres_slam_b200 import capi, initial_guess as ig, synthetic as syn
from oracle import pybinding as orc


def _draws(oracle, n, count, variant):
    a, b = np.zeros(count, dtype=np.uint32), np.zeros(count, dtype=np.uint32)
    oracle.ransac_draws(n, count, variant, capi.u32ptr(a), capi.u32ptr(b))
    return a, b


@pytest.mark.parametrize("n", [3, 7, 150, 1000, 65537, 3000000000])
def test_restated_uniform_int_matches_libstdcxx(oracle, n):
    """Variant 1 (Lemire) is what libstdc++ >= 11 does for a 32-bit generator: the restatement must
    reproduce std::uniform_int_distribution<unsigned>(0, n-1) on std::mt19937(42) of THIS compiler
    draw for draw.  Variant 0 (scaling + rejection, libstdc++ <= 10) maps a raw draw r to
    r // (0xFFFFFFFF // n) instead of (r * n) >> 32: the same value except near bucket edges, so the
    two sequences agree almost everywhere for small n and both stay in range."""
    a, b = _draws(oracle, n, 5000, 1)
    assert np.array_equal(a, b)
    c, _ = _draws(oracle, n, 5000, 0)
    assert c.max() < n
    if n <= 1000:
        assert (c == a).mean() > 0.99
    if n >= 1000:
        assert abs(c.astype(np.float64).mean() / (n - 1) - 0.5) < 0.02


def test_classic_variant_known_values(oracle):
    """Scaling + rejection, worked by hand for n = 150: scaling = 0xFFFFFFFF // 150, the first
    mt19937(42) output 1608637542 maps to 1608637542 // 28633115 = 56."""
    c, _ = _draws(oracle, 150, 3, 0)
    assert 0xFFFFFFFF // 150 == 28633115
    assert c[0] == 1608637542 // 28633115 == 56


def test_product_triples_match_oracle_draws(oracle):
    """cslam_ransac_triples (host code of the product) against the oracle's draw loop: distinct
    indices, same sequence, both variants."""
    lib = capi.load_product()
    for variant in (0, 1):
        for n in (3, 4, 150):
            tri = np.zeros(3 * 400, dtype=np.uint32)
            assert lib.ransac_triples(n, 400, variant, capi.u32ptr(tri)) == 0
            tri = tri.reshape(400, 3)
            assert tri.max() < n
            assert np.all(tri[:, 0] != tri[:, 1]) and np.all(tri[:, 0] != tri[:, 2]) and np.all(tri[:, 1] != tri[:, 2])
            # replay with the oracle's single-draw function: duplicates are re-drawn in the same order
            draws, _ = _draws(oracle, n, 6000, variant)
            it = iter(draws)
            for a, b, c in tri:
                x = next(it)
                y = next(it)
                while y == x:
                    y = next(it)
                z = next(it)
                while z == x or z == y:
                    z = next(it)
                assert (a, b, c) == (x, y, z)


def _svd_rotation(p0, p1):
    pb, qb = p0.mean(axis=0), p1.mean(axis=0)
    W = (p1 - qb).T @ (p0 - pb) / p0.shape[0]
    U, _, Vt = np.linalg.svd(W)
    M = np.diag([1.0, 1.0, np.linalg.det(U) * np.linalg.det(Vt)])
    C = U @ M @ Vt
    return np.concatenate([qb - C @ pb, C.reshape(9)])


@pytest.mark.parametrize("n", [3, 4, 50])
def test_kabsch_against_lapack_svd(oracle, n):
    """compute_transformation (point_cloud_aligner.cpp:12-62) against numpy/LAPACK: the formula
    U diag(1, 1, det U det V) V^T written out with an independent SVD."""
    rng = np.random.default_rng(n)
    for trial in range(20):
        p0 = rng.normal(0, 5, (n, 3)) + np.array([0, 0, 15.0])
        R = syn.so3_exp(rng.normal(0, 0.3, (1, 3)))[0]
        p1 = p0 @ R.T + rng.normal(0, 0.5, 3) + rng.normal(0, 0.05, (n, 3))
        T = np.zeros(12)
        oracle.kabsch(n, capi.dptr(np.ascontiguousarray(p0)), capi.dptr(np.ascontiguousarray(p1)), capi.dptr(T))
        ref = _svd_rotation(p0, p1)
        assert np.abs(T - ref).max() < 1e-11
        C = T[3:].reshape(3, 3)
        assert abs(np.linalg.det(C) - 1) < 1e-12 and np.abs(C @ C.T - np.eye(3)).max() < 1e-12


def _track_with_outliers(n_poses=30, seed=13, frac=0.15):
    # 0.25 px noise: the reference's threshold (squared distance 4 over u, v, d of TWO noisy
    # observations) then keeps nearly every true match
    tr = syn.make_track(n_poses, 15, 10, seed=seed, pix_sigma=0.25)
    rng = np.random.default_rng(seed + 1)
    bad = rng.random(tr["uvd"].shape[0]) < frac
    tr["uvd"][bad] += rng.normal(0, 25.0, (int(bad.sum()), 3))
    tr["uvd"][:, 2] = np.maximum(tr["uvd"][:, 2], 1.0)
    return tr


def test_initial_guess_oracle_follows_ground_truth():
    """compute_initial_guess end to end on the CPU restatement: chained RANSAC poses stay close to the
    ground truth despite 15 % gross outliers, and only inliers initialise points."""
    tr = _track_with_outliers()
    # the synthetic track ramps up (its first and last frames share only a handful of points):
    # run the front end on the steady part, anchored at the ground-truth pose k1
    k1, k2 = 9, tr["n_poses"] - 9
    poses = np.tile(tr["poses_gt"][k1], (tr["n_poses"], 1))
    points = np.zeros((tr["n_points"], 3))
    init = np.zeros(tr["n_points"], dtype=bool)
    st = orc.compute_initial_guess(tr, poses, points, init, k1=k1, k2=k2)
    assert st["n_matches"].min() >= 40 and np.all(st["n_inliers"] >= 0.5 * st["n_matches"])
    assert np.abs(poses[k1:k2, :3] - tr["poses_gt"][k1:k2, :3]).max() < 0.3
    assert np.abs(poses[k1:k2, 3:] - tr["poses_gt"][k1:k2, 3:]).max() < 0.03
    assert init.sum() > 100
    err = np.linalg.norm(points[init] - tr["points_gt"][init], axis=1)
    assert np.median(err) < 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_ransac_gpu_matches_oracle(product, variant):
    """The batched kernel against the oracle, pair by pair: same best hypothesis (hence the same
    transformation to rounding), same inlier set, same count — including pairs below 3 matches."""
    tr = _track_with_outliers(40, seed=5)
    rng = ig.state_ranges(tr["obs_cam"], tr["n_poses"])
    pt = tr["obs_pt"].astype(np.int64)
    pairs0, pairs1 = [], []
    for k in range(1, tr["n_poses"]):
        kp, kc = ig.match_pair(pt[rng[k - 1]:rng[k]], pt[rng[k]:rng[k + 1]])
        pairs0.append(ig.triangulate(tr["cam"], tr["uvd"][rng[k - 1]:rng[k]][kp]))
        pairs1.append(ig.triangulate(tr["cam"], tr["uvd"][rng[k]:rng[k + 1]][kc]))
    R = syn.so3_exp(np.array([[0.02, -0.05, 0.01]]))[0]
    tri0 = np.array([[1.0, 0.5, 9.0], [-2.0, 0.3, 14.0], [0.5, -1.0, 20.0]])
    pairs0 += [pairs0[0][:2], pairs0[1][:0], tri0]               # 2, 0 and exactly 3 (rigid) correspondences
    pairs1 += [pairs1[0][:2], pairs1[1][:0], tri0 @ R.T + np.array([0.1, 0.0, -0.3])]
    Tg, ig_in, cg = ig.ransac_align(pairs0, pairs1, tr["cam"], rng_variant=variant)
    To, io_in, co = orc.ransac_align(pairs0, pairs1, tr["cam"], rng_variant=variant)
    assert np.array_equal(cg, co)
    assert cg[-3] == 0 and cg[-2] == 0 and cg[-1] == 3
    for a, b in zip(ig_in, io_in):
        assert np.array_equal(a, b)
    assert np.abs(Tg - To).max() < 1e-9
    assert (cg[8:-11] > 40).all()      # steady part of the track (the ends share few points)


@pytest.mark.gpu
def test_initial_guess_gpu_matches_oracle(product):
    """compute_initial_guess through the GPU kernel equals the oracle's: poses, points, flags."""
    tr = _track_with_outliers()
    out = {}
    for backend in ("b200", "oracle"):
        poses = np.tile(tr["poses_gt"][0], (tr["n_poses"], 1))
        points = np.zeros((tr["n_points"], 3))
        init = np.zeros(tr["n_points"], dtype=bool)
        (ig if backend == "b200" else orc).compute_initial_guess(tr, poses, points, init)
        out[backend] = (poses, points, init)
    assert np.array_equal(out["b200"][2], out["oracle"][2])
    assert np.abs(out["b200"][0] - out["oracle"][0]).max() < 1e-8
    assert np.abs(out["b200"][1] - out["oracle"][1]).max() < 1e-8
