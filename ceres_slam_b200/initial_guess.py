"""Host mirror of `DatasetProblem::compute_initial_guess` (src/ceres_slam/dataset_problem.cpp:179-270)
over the batched RANSAC entry point of the C ABI (`cslam_ransac_align`, KR in csrc/ransac.cu).

The reference walks the pose pairs (k-1, k) one at a time: match the point ids seen by both poses,
triangulate both clouds, run the 3-point RANSAC (400 hypotheses, threshold 4), chain
`poses[k] = T_k_km1 * poses[k-1]` and initialise every inlier point that has no guess yet from the
first cloud.  The RANSAC of a pair only reads observations, so all pairs of a call go to the GPU as
ONE launch; the chaining and the point initialisation are then a cheap sequential pass on the host
in the reference's order.
"""
import ctypes as C

import numpy as np

from . import capi


def triangulate(cam, uvd):
    """StereoCamera::triangulate (stereo_camera.hpp:112-120), same operation order."""
    b_over_d = cam["b"] / uvd[:, 2]
    fu_over_fv = cam["fu"] / cam["fv"]
    return np.stack([(uvd[:, 0] - cam["cu"]) * b_over_d, (uvd[:, 1] - cam["cv"]) * b_over_d * fu_over_fv,
                     cam["fu"] * b_over_d], axis=1)


def state_ranges(obs_cam, n_states):
    """obs_indices_at_state for a pose-major file: [start, end) of every state's observations."""
    k = np.asarray(obs_cam, dtype=np.int64)
    assert np.all(np.diff(k) >= 0), "observations must be grouped by state, as in the reference's files"
    return np.searchsorted(k, np.arange(n_states + 1), side="left")


def match_pair(ids_prev, ids_cur):
    """Reciprocal matches (dataset_problem.cpp:207-221): both lists keep their own order and are then
    paired BY POSITION (pts_km1[i] <-> pts_k[i], :224-229) — the reference's behaviour, kept."""
    keep_prev = np.flatnonzero(np.isin(ids_prev, ids_cur))
    keep_cur = np.flatnonzero(np.isin(ids_cur, ids_prev))
    return keep_prev, keep_cur


def ransac_align(pairs0, pairs1, cam, num_iters=400, thresh=4.0, rng_variant=0, device=0, entry=None):
    """Batched compute_transformation_and_inliers.  pairs0/pairs1: lists of (n_i, 3) arrays.
    `entry` is the C entry point to call (default: the library's `cslam_ransac_align`)."""
    entry = entry or capi.load_product().ransac_align
    n_pairs = len(pairs0)
    sizes = np.array([p.shape[0] for p in pairs0], dtype=np.int64)
    offsets = np.zeros(n_pairs + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum(sizes)
    total = int(offsets[-1])
    p0 = np.ascontiguousarray(np.concatenate(pairs0, axis=0) if total else np.zeros((1, 3)), dtype=np.float64)
    p1 = np.ascontiguousarray(np.concatenate(pairs1, axis=0) if total else np.zeros((1, 3)), dtype=np.float64)
    intr = np.array([cam["fu"], cam["fv"], cam["cu"], cam["cv"], cam["b"]], dtype=np.float64)
    T = np.zeros((max(n_pairs, 1), 12))
    inl = np.zeros(max(total, 1), dtype=np.uint8)
    cnt = np.zeros(max(n_pairs, 1), dtype=np.uint32)
    st = entry(device, n_pairs, capi.u32ptr(offsets), capi.dptr(p0), capi.dptr(p1), capi.dptr(intr),
                          num_iters, float(thresh), rng_variant, capi.dptr(T), capi.u8ptr(inl), capi.u32ptr(cnt))
    if st != 0:
        raise RuntimeError(f"cslam_ransac_align status {st}")
    return T[:n_pairs], [inl[offsets[i]:offsets[i + 1]].astype(bool) for i in range(n_pairs)], cnt[:n_pairs]


def compute_initial_guess(track, poses, points, initialized, k1=0, k2=None, num_iters=400,
                          thresh=4.0, rng_variant=0, device=0, entry=None):
    """dataset_problem.cpp:179-270.  `poses` (n, 12) [t | R], `points` (m, 3) and the boolean
    `initialized` (m,) are updated in place; poses[k1] is the anchor.  Returns per-pair statistics."""
    n_states = track["n_poses"]
    if k2 is None or k1 >= k2:
        k1, k2 = 0, n_states
    cam = track["cam"]
    rng = state_ranges(track["obs_cam"], n_states)
    pt = np.asarray(track["obs_pt"], dtype=np.int64)
    uvd = np.asarray(track["uvd"], dtype=np.float64)
    pairs0, pairs1, ids = [], [], []
    for k in range(k1 + 1, k2):
        a0, a1, b0, b1 = rng[k - 1], rng[k], rng[k], rng[k + 1]
        kp, kc = match_pair(pt[a0:a1], pt[b0:b1])
        n = min(kp.size, kc.size)
        kp, kc = kp[:n], kc[:n]
        pairs0.append(triangulate(cam, uvd[a0:a1][kp]))
        pairs1.append(triangulate(cam, uvd[b0:b1][kc]))
        ids.append(pt[a0:a1][kp])
    T, inl, cnt = ransac_align(pairs0, pairs1, cam, num_iters, thresh, rng_variant, device, entry)
    for i, k in enumerate(range(k1 + 1, k2)):
        Rr, tr = T[i, 3:].reshape(3, 3), T[i, :3]
        Rp, tp = poses[k - 1, 3:].reshape(3, 3), poses[k - 1, :3]
        poses[k, :3] = Rr @ tp + tr                      # T_k_km1 * poses[k-1]  (se3group.hpp:176-183)
        poses[k, 3:] = (Rr @ Rp).reshape(9)
        sel = np.flatnonzero(inl[i])
        j = ids[i][sel]
        fresh = ~initialized[j]
        if fresh.any():
            Rt = Rp.T
            t_inv = -(Rt @ tp)                          # poses[k-1].inverse()  (se3group.hpp:152-157)
            # a point id can appear once per pair, so the vectorised write equals the reference's loop
            points[j[fresh]] = pairs0[i][sel[fresh]] @ Rt.T + t_inv
            initialized[j[fresh]] = True
    return {"T_rel": T, "n_matches": np.array([p.shape[0] for p in pairs0]), "n_inliers": cnt}
