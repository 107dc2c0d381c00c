"""Shared by tests/golden/make_ref_golden.py (generator) and tests/test_ref_pin.py.

`build_inputs()` — a deterministic set of inputs for every functor and plus operation on the path:
the inputs of the mpmath golden file, blocks of a synthetic track (stereo with shared and with
per-observation stiffness, sun, pose prior, intensity with a point and a directional light, normal)
and the branch cases the reference's code has (sun thresholds and wrap, prior at its own reference
(theta = 0 branch of SO3::log), small rotations around the `angle <= eps` switch of SO3::exp,
shadowed / back-facing / clamped Phong terms).

`evaluate(lib_kind, lib, cases)` — runs every case through `oracle/_ref` (the reference's own
headers) or through the oracle restatement, returning the same nested dict of lists.
"""
import json
import os

import numpy as np

from ceres_slam_b200 import capi, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_MP60 = os.path.join(ROOT, "tests", "golden", "functors_mp60.json")
REF_BLOCKS = os.path.join(ROOT, "tests", "golden", "ref_blocks.json")
d = capi.dptr


def A(x, dtype=np.float64):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def _so3_exp(phi):
    return syn.so3_exp(np.asarray(phi, dtype=np.float64).reshape(1, 3))[0]


def _pose(t, phi):
    return np.concatenate([np.asarray(t, dtype=np.float64), _so3_exp(phi).reshape(9)])


def build_inputs():
    g = json.load(open(GOLDEN_MP60))
    rng = np.random.default_rng(20261018)
    cases = {"camera": g["camera"]}

    # ---- stereo: golden inputs + blocks of a track (shared W) + per-observation W ----------------
    tr = syn.make_track(30, 12, 6, seed=5)
    sel = rng.choice(tr["uvd"].shape[0], 40, replace=False)
    stereo = [dict(pose=c["pose"], point=c["point"], uvd=c["uvd"], W=c["W"]) for c in g["stereo"]]
    for i in sel:
        stereo.append(dict(pose=tr["poses"][tr["obs_cam"][i]].tolist(), point=tr["points"][tr["obs_pt"][i]].tolist(),
                           uvd=tr["uvd"][i].tolist(), W=tr["W"].reshape(9).tolist()))
    trw = syn.make_track(20, 8, 5, seed=6, per_obs_W=True)
    for i in rng.choice(trw["uvd"].shape[0], 24, replace=False):
        stereo.append(dict(pose=trw["poses"][trw["obs_cam"][i]].tolist(), point=trw["points"][trw["obs_pt"][i]].tolist(),
                           uvd=trw["uvd"][i].tolist(), W=trw["W"][i].reshape(9).tolist()))
    cases["stereo"] = stereo

    # ---- sun: golden + track + thresholds / wrap -------------------------------------------------
    sun = [dict(pose=c["pose"], obs_c=c["obs_c"], ref_g=c["ref_g"], W=c["W"], az_thresh=c["az_thresh"],
                zen_thresh=c["zen_thresh"]) for c in g["sun"]]
    trs = syn.add_sun(syn.make_track(24, 6, 5, seed=7))
    for k in range(0, 24, 2):
        sun.append(dict(pose=trs["poses"][k].tolist(), obs_c=trs["sun_obs_c"][k].tolist(),
                        ref_g=trs["sun_ref_g"][k].tolist(), W=trs["sun_W"][k].tolist(), az_thresh=0.5, zen_thresh=0.5))
    for k in range(6):   # tight thresholds: one or both components replaced by the constant 0
        sun.append(dict(pose=trs["poses"][k].tolist(), obs_c=trs["sun_obs_c"][k].tolist(),
                        ref_g=trs["sun_ref_g"][k].tolist(), W=[30.0, 1.5, -2.0, 25.0],
                        az_thresh=[0.01, 1.0, 0.02][k % 3], zen_thresh=[1.0, 0.01, 0.02][k % 3]))
    for s in (1.0, -1.0):  # azimuths either side of +-pi: the wrap branches (sun_sensor_error.hpp:80-84)
        pose = _pose([0.1, -0.2, 0.3], [0.0, 0.02 * s, 0.0])
        e = np.array([0.02 * s, -0.4, -1.0]); e /= np.linalg.norm(e)
        o = np.array([-0.03 * s, -0.41, -1.0]); o /= np.linalg.norm(o)
        sun.append(dict(pose=pose.tolist(), obs_c=o.tolist(), ref_g=e.tolist(), W=[20.0, 0.0, 0.0, 20.0],
                        az_thresh=7.0, zen_thresh=7.0))
    cases["sun"] = sun

    # ---- pose prior: golden + perturbed references + at its own reference + tiny rotations -------
    prior = [dict(pose=c["pose"], Tref=c["Tref"], W=c["W"]) for c in g["prior"]]
    for k in range(10):
        T = tr["poses_gt"][k + 2]
        dphi = rng.normal(0, [1e-3, 0.05, 0.5, 2.0][k % 4], 3)
        Tref = _pose(T[:3] + rng.normal(0, 0.1, 3), dphi)
        Tref[3:] = (_so3_exp(dphi) @ T[3:].reshape(3, 3)).reshape(9)
        Wm = np.eye(6) * rng.uniform(1, 50) + 0.3 * rng.normal(size=(6, 6))
        prior.append(dict(pose=T.tolist(), Tref=Tref.tolist(), W=((Wm + Wm.T) / 2).reshape(36).tolist()))
    prior.append(dict(pose=tr["poses_gt"][3].tolist(), Tref=tr["poses_gt"][3].tolist(), W=np.eye(6).reshape(36).tolist()))
    cases["prior"] = prior

    # ---- lighting: golden + blocks of a synthetic scene ------------------------------------------
    normal = [dict(pose=c["pose"], normal=c["normal"], obs=c["obs"], W=c["W"]) for c in g["normal"]]
    intensity = [dict(kind=c["kind"], directional=c["directional"], pose=c["pose"], point=c["point"], normal=c["normal"],
                      phong=c["phong"], texture=c["texture"], light=c["light"], colour=c["colour"],
                      stiffness=c["stiffness"]) for c in g["intensity"]]
    for directional in (False, True):
        trp = syn.add_phong(syn.make_track(14, 10, 5, seed=9 + int(directional)), directional=directional)
        for i in rng.choice(trp["uvd"].shape[0], 30, replace=False):
            k, j = int(trp["obs_cam"][i]), int(trp["obs_pt"][i])
            m = int(trp["material_id"][j])
            intensity.append(dict(kind="track", directional=int(directional), pose=trp["poses"][k].tolist(),
                                  point=trp["points"][j].tolist(), normal=trp["normals"][j].tolist(),
                                  phong=trp["phong"][m].tolist(), texture=[float(trp["textures"][j])],
                                  light=trp["light"].tolist(), colour=float(trp["intensity"][i]),
                                  stiffness=float(trp["int_stiffness"])))
            normal.append(dict(pose=trp["poses"][k].tolist(), normal=trp["normals"][j].tolist(),
                               obs=trp["normal_obs"][i].tolist(), W=trp["W_normal"].tolist()))
        # a normal turned away from the light (diffuse and specular off) and one over-bright (clamp at 1)
        i = 0
        k, j = int(trp["obs_cam"][i]), int(trp["obs_pt"][i])
        base = dict(directional=int(directional), pose=trp["poses"][k].tolist(), point=trp["points"][j].tolist(),
                    light=trp["light"].tolist(), colour=0.4, stiffness=100.0)
        intensity.append(dict(base, kind="backfacing", normal=(-trp["normals"][j]).tolist(), phong=[0.0, 0.4, 10.0],
                              texture=[0.5]))
        intensity.append(dict(base, kind="clamped", normal=trp["normals"][j].tolist(), phong=[0.0, 3.0, 1.5], texture=[4.0]))
    cases["normal"] = normal
    cases["intensity"] = intensity

    # ---- plus operations --------------------------------------------------------------------------
    se3_plus = [dict(pose=c["pose"], delta=c["delta"]) for c in g["se3_plus"]]
    for k, scale in enumerate([0.0, 1e-17, 1e-16, 3e-16, 1e-12, 1e-8, 1e-3, 0.3, 2.5, 3.5]):
        dl = rng.normal(size=6)
        dl[3:] *= scale / max(np.linalg.norm(dl[3:]), 1e-300)
        se3_plus.append(dict(pose=tr["poses"][k].tolist(), delta=dl.tolist()))
    cases["se3_plus"] = se3_plus
    unit_plus = [dict(x=c["x"], delta=c["delta"]) for c in g["unit_plus"]]
    for k in range(8):
        x = rng.normal(size=3)
        x /= np.linalg.norm(x)
        unit_plus.append(dict(x=x.tolist(), delta=(rng.normal(size=3) * [1e-9, 1e-3, 0.1, 1.0][k % 4]).tolist()))
    unit_plus.append(dict(x=[0.6, 0.0, 0.8], delta=[0.0, 0.0, 0.0]))
    unit_plus.append(dict(x=[1.2, -0.4, 0.3], delta=[0.01, 0.02, -0.03]))      # non-unit x: the formula divides by |x|^2
    cases["unit_plus"] = unit_plus
    so3 = []
    for scale in [0.0, 1e-17, 2e-16, 2.3e-16, 1e-10, 1e-5, 0.1, 1.0, 3.0, 3.14, 3.7]:
        v = rng.normal(size=3)
        so3.append(dict(phi=(v / np.linalg.norm(v) * scale).tolist()))
    cases["so3"] = so3
    return cases


def _out(n):
    return np.zeros(n)


def evaluate(kind, lib, cases):
    """kind: "ref" (oracle/_ref) or "oracle" (the restatement).  Returns the outputs of every case."""
    from oracle import pybinding as orc
    intr = A(cases["camera"])
    res = {}
    # stereo
    out = []
    for c in cases["stereo"]:
        pose, point, uvd, W = A(c["pose"]), A(c["point"]), A(c["uvd"]), A(c["W"])
        if kind == "ref":
            r, Jc, Jp = _out(3), _out(18), _out(9)
            z = np.zeros(1, dtype=np.uint32)
            assert lib.stereo_blocks(1, d(intr), capi.u32ptr(z), capi.u32ptr(z), d(uvd), d(W), 0, d(pose), d(point),
                                     d(r), d(Jc), d(Jp), None) == 0
        else:
            p = orc.OracleProblem()
            p.set_camera(*cases["camera"])
            p.set_poses(pose.reshape(1, 12).copy(), np.zeros(1, dtype=np.uint8))
            p.set_points(point.reshape(1, 3).copy())
            p.add_stereo(np.zeros(1, np.uint32), np.zeros(1, np.uint32), uvd, W)
            e = p.evaluate()
            r, Jc, Jp = e["r_stereo"].ravel(), e["Jpose_stereo"].ravel(), e["Jpoint_stereo"].ravel()
            p.close()
        out.append(dict(r=r.tolist(), J_pose=Jc.tolist(), J_point=Jp.tolist()))
    res["stereo"] = out
    # sun
    out = []
    for c in cases["sun"]:
        pose, o, e_g, W = A(c["pose"]), A(c["obs_c"]), A(c["ref_g"]), A(c["W"])
        if kind == "ref":
            r, J = _out(2), _out(12)
            z = np.zeros(1, dtype=np.uint32)
            assert lib.sun_blocks(1, capi.u32ptr(z), d(o), d(e_g), d(W), c["az_thresh"], c["zen_thresh"], d(pose), d(r), d(J)) == 0
        else:
            p = orc.OracleProblem()
            p.set_camera(*cases["camera"])
            p.set_poses(pose.reshape(1, 12).copy(), np.zeros(1, dtype=np.uint8))
            p.set_points(np.array([[0.0, 0.0, 5.0]]))
            p.add_sun(np.zeros(1, np.uint32), o, e_g, W, c["az_thresh"], c["zen_thresh"])
            e = p.evaluate()
            r, J = e["r_sun"].ravel(), e["J_sun"].ravel()
            p.close()
        out.append(dict(r=r.tolist(), J_pose=J.tolist()))
    res["sun"] = out
    # prior
    out = []
    for c in cases["prior"]:
        pose, Tref, W = A(c["pose"]), A(c["Tref"]), A(c["W"])
        if kind == "ref":
            r, J = _out(6), _out(36)
            assert lib.prior_block(d(pose), d(Tref), d(W), d(r), d(J)) == 0
        else:
            p = orc.OracleProblem()
            p.set_camera(*cases["camera"])
            p.set_poses(pose.reshape(1, 12).copy(), np.zeros(1, dtype=np.uint8))
            p.set_points(np.array([[0.0, 0.0, 5.0]]))
            p.add_pose_prior(0, Tref, W)
            e = p.evaluate()
            r, J = e["r_prior"].ravel(), e["J_prior"].ravel()
            p.close()
        out.append(dict(r=r.tolist(), J_pose=J.tolist()))
    res["prior"] = out
    # normal / intensity: same signature in both libraries
    out = []
    for c in cases["normal"]:
        r, Jc, Jn = _out(3), _out(18), _out(9)
        assert lib.normal_block(d(A(c["pose"])), d(A(c["normal"])), d(A(c["obs"])), d(A(c["W"])), d(r), d(Jc), d(Jn)) == 0
        out.append(dict(r=r.tolist(), J_pose=Jc.tolist(), J_normal=Jn.tolist()))
    res["normal"] = out
    out = []
    for c in cases["intensity"]:
        r, Jc, Jp, Jn, Jk, Jt, Jl = (_out(n) for n in (1, 6, 3, 3, 3, 1, 3))
        assert lib.intensity_block(d(A(c["pose"])), d(A(c["point"])), d(A(c["normal"])), d(A(c["phong"])),
                                   d(A(c["texture"])), d(A(c["light"])), c["colour"], c["stiffness"], c["directional"],
                                   d(r), d(Jc), d(Jp), d(Jn), d(Jk), d(Jt), d(Jl)) == 0
        out.append(dict(r=r.tolist(), J_pose=Jc.tolist(), J_point=Jp.tolist(), J_normal=Jn.tolist(), J_phong=Jk.tolist(),
                        J_tex=Jt.tolist(), J_light=Jl.tolist()))
    res["intensity"] = out
    # plus operations
    out = []
    for c in cases["se3_plus"]:
        o, J = _out(12), _out(72)
        lib.se3_plus(d(A(c["pose"])), d(A(c["delta"])), d(o))
        lib.se3_plus_jacobian(d(A(c["pose"])), d(J))
        out.append(dict(out=o.tolist(), J_plus=J.tolist()))
    res["se3_plus"] = out
    out = []
    for c in cases["unit_plus"]:
        o, J = _out(3), _out(9)
        lib.unit_plus(d(A(c["x"])), d(A(c["delta"])), d(o))
        lib.unit_plus_jacobian(d(A(c["x"])), d(J))
        out.append(dict(out=o.tolist(), J_plus=J.tolist()))
    res["unit_plus"] = out
    out = []
    for c in cases["so3"]:
        R, back = _out(9), _out(3)
        lib.so3_exp(d(A(c["phi"])), d(R))
        lib.so3_log(d(R), d(back))
        out.append(dict(R=R.tolist(), log_of_exp=back.tolist()))
    res["so3"] = out
    return res


# ---- front end: pose pairs for PointCloudAligner (src/ceres_slam/point_cloud_aligner.cpp) --------------------
REF_RANSAC = os.path.join(ROOT, "tests", "golden", "ref_ransac.json")
RANSAC_SETTINGS = [(400, 4.0), (400, 25.0), (60, 9.0)]   # (num_iters, thresh); 400 / 25 are the header's defaults


def build_ransac_pairs():
    """Matched, triangulated point clouds of consecutive frames of a noisy synthetic track with 15 % gross
    outliers (what compute_initial_guess hands to the aligner, dataset_problem.cpp:196-234), every pair with at
    least 3 matches (below that the reference's draw loop does not terminate), plus one exactly rigid triple."""
    from ceres_slam_b200 import initial_guess as ig, synthetic as syn
    tr = syn.make_track(40, 15, 10, seed=5, pix_sigma=0.25)
    rng = np.random.default_rng(6)
    bad = rng.random(tr["uvd"].shape[0]) < 0.15
    tr["uvd"][bad] += rng.normal(0, 25.0, (int(bad.sum()), 3))
    tr["uvd"][:, 2] = np.maximum(tr["uvd"][:, 2], 1.0)
    rg = ig.state_ranges(tr["obs_cam"], tr["n_poses"])
    pt = tr["obs_pt"].astype(np.int64)
    pairs0, pairs1 = [], []
    for k in range(1, tr["n_poses"]):
        kp, kc = ig.match_pair(pt[rg[k - 1]:rg[k]], pt[rg[k]:rg[k + 1]])
        if min(kp.size, kc.size) < 3:
            continue
        pairs0.append(ig.triangulate(tr["cam"], tr["uvd"][rg[k - 1]:rg[k]][kp]))
        pairs1.append(ig.triangulate(tr["cam"], tr["uvd"][rg[k]:rg[k + 1]][kc]))
    R = _so3_exp(np.array([0.02, -0.05, 0.01]))
    tri0 = np.array([[1.0, 0.5, 9.0], [-2.0, 0.3, 14.0], [0.5, -1.0, 20.0]])
    pairs0.append(tri0)
    pairs1.append(tri0 @ R.T + np.array([0.1, 0.0, -0.3]))
    return tr["cam"], pairs0, pairs1


def ransac_inputs_digest(cam, pairs0, pairs1):
    import hashlib
    h = hashlib.sha256()
    h.update(np.array([cam[k] for k in ("fu", "fv", "cu", "cv", "b")], dtype=np.float64).tobytes())
    for a, b in zip(pairs0, pairs1):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        h.update(np.ascontiguousarray(b, dtype=np.float64).tobytes())
    return h.hexdigest()


def evaluate_ransac(entry, cam, pairs0, pairs1, rng_variant=1):
    """[{num_iters, thresh, T (n_pairs x 12), inliers (index lists), counts}] through any library's
    ransac_align entry (the reference's source, the oracle, the CUDA kernel: one signature)."""
    from ceres_slam_b200 import initial_guess as ig
    out = []
    for num_iters, thresh in RANSAC_SETTINGS:
        T, inl, cnt = ig.ransac_align(pairs0, pairs1, cam, num_iters=num_iters, thresh=thresh, rng_variant=rng_variant,
                                      entry=entry)
        out.append({"num_iters": num_iters, "thresh": thresh, "T": [[float(x) for x in row] for row in T],
                    "inliers": [[int(i) for i in np.flatnonzero(m)] for m in inl], "counts": [int(c) for c in cnt]})
    return out
