"""Deterministic synthetic stereo tracks in the shapes BASELINE.json names.

Intrinsics are the ones the reference's camera_test.cpp:11-15 uses; the base seed 42 is the seed
the reference hard-codes for RANSAC (point_cloud_aligner.cpp:72).  Poses are stored the way the
reference stores them: 12 doubles [t | R row-major] of T_c_g (se3group.hpp:115-118), so that
p_c = R p_g + t.  Observations come out grouped by pose index like the reference's CSV rows
(dataset_problem.cpp:71-98).
"""
import numpy as np

KITTI = dict(fu=707.0912, fv=707.0912, cu=601.8873, cv=183.1104, b=0.535105804)
IMG_W, IMG_H = 1242.0, 375.0


def so3_exp(phi):
    """Rodrigues, vectorised over rows of phi (n,3) -> (n,3,3)."""
    phi = np.atleast_2d(phi)
    ang = np.linalg.norm(phi, axis=1)
    small = ang < 1e-12
    axis = phi / np.where(small, 1.0, ang)[:, None]
    K = np.zeros((phi.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -axis[:, 2], axis[:, 1]
    K[:, 1, 0], K[:, 1, 2] = axis[:, 2], -axis[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -axis[:, 1], axis[:, 0]
    c, s = np.cos(ang)[:, None, None], np.sin(ang)[:, None, None]
    R = c * np.eye(3) + (1 - c) * axis[:, :, None] * axis[:, None, :] + s * K
    R[small] = np.eye(3)
    return R


def pose_pack(R, t):
    return np.concatenate([t, R.reshape(-1, 9)], axis=1)


def pose_R(p):
    return p[:, 3:].reshape(-1, 3, 3)


def pose_t(p):
    return p[:, :3]


def perturb_poses(poses, eps):
    """Left-multiplicative decoupled update T <- exp(eps) T (perturbations.hpp:62)."""
    E = so3_exp(eps[:, 3:])
    R = E @ pose_R(poses)
    t = np.einsum("nij,nj->ni", E, pose_t(poses)) + eps[:, :3]
    return pose_pack(R, t)


def loop_poses(n_poses, spacing=0.3, min_radius=10.0):
    """Closed planar loop, camera looking along the tangent (z forward, x outward, y down)."""
    radius = max(min_radius, n_poses * spacing / (2 * np.pi))
    th = 2 * np.pi * np.arange(n_poses) / n_poses
    c = np.stack([radius * np.cos(th), radius * np.sin(th), np.zeros_like(th)], axis=1)
    x = np.stack([np.cos(th), np.sin(th), np.zeros_like(th)], axis=1)
    y = np.tile(np.array([0.0, 0.0, -1.0]), (n_poses, 1))
    z = np.stack([-np.sin(th), np.cos(th), np.zeros_like(th)], axis=1)
    R = np.stack([x, y, z], axis=1)
    t = -np.einsum("nij,nj->ni", R, c)
    return pose_pack(R, t)


def project(cam, pc):
    iz = 1.0 / pc[:, 2]
    return np.stack([cam["fu"] * pc[:, 0] * iz + cam["cu"], cam["fv"] * pc[:, 1] * iz + cam["cv"],
                     cam["fu"] * cam["b"] * iz], axis=1)


def make_track(n_poses, new_per_frame, track_len, seed=42, pix_sigma=1.0, pose_sigma=(0.02, 0.01),
               point_sigma=0.05, spacing=0.3, per_obs_W=False, cam=None, closed=False, ragged=None):
    """A sims-style stereo track.

    Every frame sees `new_per_frame` new landmarks, each tracked over `track_len` consecutive
    poses, i.e. about new_per_frame * track_len observations per frame in steady state
    (C1/C2: 15 x 10; C5: 100 x 10 -> 20 k poses, 2 M landmarks, 20 M observations).

    closed=True closes the loop the way the reference's simulated trajectories do
    (scripts/ba_all_sims.sh:8-13: triangle / square / penta / circle tracks): the tracks of the last
    frames run on into the first ones, so the last poses share landmarks with the first and the
    reduced camera system gets blocks far from its diagonal.

    ragged=dict(mean=..., max=..., drop=...) makes the tracks what a real stereo front end produces
    instead of `track_len` consecutive frames each: lengths 2 + Geometric(1 / (mean - 2)) clipped to
    `max` (default 30), and every observation after the first dropped with probability `drop`
    (default 0.1; at least two are kept).  `track_len` is then only the mean used when `mean` is absent.
    """
    cam = dict(cam or KITTI)
    rng = np.random.default_rng(seed)
    poses_gt = loop_poses(n_poses, spacing)
    n_starts = n_poses if closed else n_poses - track_len + 1
    n_pts = n_starts * new_per_frame
    first = np.repeat(np.arange(n_starts, dtype=np.int64), new_per_frame)
    # place each landmark in the frustum of the middle pose of its track
    mid = (first + track_len // 2) % n_poses
    u = rng.uniform(150.0, IMG_W - 150.0, n_pts)
    v = rng.uniform(60.0, IMG_H - 60.0, n_pts)
    z = rng.uniform(4.0, 30.0, n_pts)
    pc = np.stack([(u - cam["cu"]) * z / cam["fu"], (v - cam["cv"]) * z / cam["fv"], z], axis=1)
    Rm, tm = pose_R(poses_gt)[mid], pose_t(poses_gt)[mid]
    points_gt = np.einsum("nji,nj->ni", Rm, pc - tm)  # R^T (p_c - t)
    # observations, pose-major
    if ragged is None:
        k = ((first[:, None] + np.arange(track_len)[None, :]) % n_poses).reshape(-1)
        j = np.repeat(np.arange(n_pts, dtype=np.int64), track_len)
    else:
        mean = float(ragged.get("mean", track_len))
        lmax = int(ragged.get("max", 30))
        drop = float(ragged.get("drop", 0.1))
        n_starts = n_poses if closed else n_poses - 1
        n_pts = n_starts * new_per_frame
        first = np.repeat(np.arange(n_starts, dtype=np.int64), new_per_frame)
        length = np.minimum(2 + rng.geometric(1.0 / max(mean - 2.0, 1e-9), n_pts) - 1, lmax)
        if not closed:
            length = np.minimum(length, n_poses - first)
        # the landmark sits in the frustum of the middle frame of ITS track
        mid = (first + length // 2) % n_poses
        u = rng.uniform(150.0, IMG_W - 150.0, n_pts)
        v = rng.uniform(60.0, IMG_H - 60.0, n_pts)
        z = rng.uniform(4.0, 30.0, n_pts)
        pc = np.stack([(u - cam["cu"]) * z / cam["fu"], (v - cam["cv"]) * z / cam["fv"], z], axis=1)
        Rm, tm = pose_R(poses_gt)[mid], pose_t(poses_gt)[mid]
        points_gt = np.einsum("nji,nj->ni", Rm, pc - tm)
        off = np.arange(lmax)[None, :]
        keep = off < length[:, None]
        dropped = (rng.random((n_pts, lmax)) < drop) & (off >= 2)   # the first two observations always stay
        keep &= ~dropped
        jj, oo = np.nonzero(keep)
        k = (first[jj] + oo) % n_poses
        j = jj.astype(np.int64)
    order = np.argsort(k, kind="stable")
    k, j = k[order], j[order]
    pck = np.einsum("nij,nj->ni", pose_R(poses_gt)[k], points_gt[j]) + pose_t(poses_gt)[k]
    uvd = project(cam, pck)
    ok = (pck[:, 2] > 0.5) & (uvd[:, 0] > 0) & (uvd[:, 0] < IMG_W) & (uvd[:, 1] > 0) & \
         (uvd[:, 1] < IMG_H) & (uvd[:, 2] > 1.0)
    k, j, uvd = k[ok], j[ok], uvd[ok]
    uvd = uvd + rng.normal(0.0, pix_sigma, uvd.shape)
    # initial guess: ground truth perturbed, first pose held at ground truth (gauge)
    eps = np.concatenate([rng.normal(0, pose_sigma[0], (n_poses, 3)),
                          rng.normal(0, pose_sigma[1], (n_poses, 3))], axis=1)
    eps[0] = 0.0
    poses_init = perturb_poses(poses_gt, eps)
    points_init = points_gt + rng.normal(0, point_sigma, points_gt.shape)
    if per_obs_W:
        # SPD covariance near diag(sigma^2) per observation -> symmetric inverse square root
        A = rng.normal(0, 0.15, (uvd.shape[0], 3, 3))
        cov = pix_sigma ** 2 * np.eye(3) + 0.5 * (A + A.transpose(0, 2, 1)) * pix_sigma ** 2 * 0.3
        w, V = np.linalg.eigh(cov)
        W = np.einsum("nij,nj,nkj->nik", V, 1.0 / np.sqrt(w), V).reshape(-1, 9)
    else:
        W = (np.eye(3) / pix_sigma).reshape(9)
    constant = np.zeros(n_poses, dtype=np.uint8)
    constant[0] = 1
    return dict(cam=cam, poses_gt=poses_gt, poses=poses_init, points_gt=points_gt,
                points=points_init, obs_cam=k.astype(np.uint32), obs_pt=j.astype(np.uint32),
                uvd=np.ascontiguousarray(uvd), W=np.ascontiguousarray(W), constant=constant,
                n_poses=n_poses, n_points=n_pts)


def add_sun(track, seed=43, sigma_deg=2.0):
    """Ephemeris direction fixed in the global frame, observed direction = R e_g + noise
    (dataset_problem_sun.cpp:139-175 file shapes)."""
    rng = np.random.default_rng(seed)
    n = track["n_poses"]
    e_g = np.array([0.3, -0.5, 0.81])
    e_g = e_g / np.linalg.norm(e_g)
    R = pose_R(track["poses_gt"])
    obs = np.einsum("nij,j->ni", R, e_g)
    noise = so3_exp(rng.normal(0, np.deg2rad(sigma_deg), (n, 3)))
    obs = np.einsum("nij,nj->ni", noise, obs)
    W2 = np.tile((np.eye(2) / np.deg2rad(sigma_deg)).reshape(4), (n, 1))
    track["sun_cam"] = np.arange(n, dtype=np.uint32)
    track["sun_obs_c"] = np.ascontiguousarray(obs)
    track["sun_ref_g"] = np.tile(e_g, (n, 1))
    track["sun_W"] = W2
    return track


def add_phong(track, seed=44, n_materials=8, directional=False, int_var=1e-4, normal_var=1e-4,
              shared_textures=False):
    """Lighting data in the shape of dataset_ba_phong's input (dataset_problem_phong.cpp:29-117):
    per vertex a unit normal facing the cameras that see it, a diffuse texture kd and a material
    id; per material [ka, ks, alpha]; one point light at (-2, -2, 2) (light_test.cpp:49) or a
    directional light; per observation the rendered Phong intensity (+ noise, variance int_var)
    and the normal in the camera frame (+ noise, variance normal_var).  `shared_textures`: one kd per
    material, shared by its vertices, as the reference's reader builds them
    (dataset_problem_phong.cpp:262-279, :342-343)."""
    rng = np.random.default_rng(seed)
    n_pts = track["n_points"]
    k, j = track["obs_cam"].astype(np.int64), track["obs_pt"].astype(np.int64)
    R, t = pose_R(track["poses_gt"]), pose_t(track["poses_gt"])
    # normal: towards the mean camera centre of the observing poses, perturbed
    centres = -np.einsum("nji,nj->ni", R, t)
    acc = np.zeros((n_pts, 3))
    np.add.at(acc, j, centres[k] - track["points_gt"][j])
    nrm = acc + rng.normal(0, 0.3, acc.shape) * np.linalg.norm(acc, axis=1, keepdims=True).clip(1e-9)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).clip(1e-12)
    nrm[~np.isfinite(nrm).all(axis=1)] = np.array([0.0, 0.0, 1.0])
    mat_id = rng.integers(0, n_materials, n_pts).astype(np.uint32)
    phong = np.stack([rng.uniform(0.0, 0.2, n_materials), rng.uniform(0.1, 0.5, n_materials),
                      rng.uniform(5.0, 30.0, n_materials)], axis=1)
    tex = rng.uniform(0.3, 0.9, n_pts)
    tex_shared = rng.uniform(0.3, 0.9, n_materials)
    if shared_textures:
        tex = tex_shared[mat_id]
    if directional:
        light = np.array([0.2, -0.4, 0.9])
        light /= np.linalg.norm(light)
    else:
        light = np.array([-2.0, -2.0, 2.0])
    # render with the model (float64 numpy restatement of phong.hpp:25-104 for data generation only)
    pc = np.einsum("nij,nj->ni", R[k], track["points_gt"][j]) + t[k]
    nc = np.einsum("nij,nj->ni", R[k], nrm[j])
    if directional:
        lv = np.einsum("nij,j->ni", R[k], light)
    else:
        lv = np.einsum("nij,nj->ni", R[k], light[None, :] - track["points_gt"][j])
    lhat = lv / np.linalg.norm(lv, axis=1, keepdims=True)
    chat = -pc / np.linalg.norm(pc, axis=1, keepdims=True)
    a = np.einsum("ni,ni->n", lhat, nc)
    diffuse = np.where(a > 0, tex[j] * a, 0.0)
    m = 2 * a[:, None] * nc - lhat
    mn = np.linalg.norm(m, axis=1, keepdims=True).clip(1e-300)
    s = np.einsum("ni,ni->n", m / mn, chat)
    spec = np.where(s > 0, phong[mat_id[j], 1] * np.power(np.clip(s, 1e-300, None), phong[mat_id[j], 2]), 0.0)
    inten = np.clip(diffuse + spec, 0.0, 1.0) + rng.normal(0, np.sqrt(int_var), k.size)
    nobs = nc + rng.normal(0, np.sqrt(normal_var), nc.shape)
    # initial guesses: perturbed normals (re-normalised), textures and material parameters
    n0 = nrm + rng.normal(0, 0.05, nrm.shape)
    n0 /= np.linalg.norm(n0, axis=1, keepdims=True)
    track.update(normals_gt=nrm, normals=n0, textures_gt=tex, textures=np.clip(tex + rng.normal(0, 0.05, n_pts), 0, 1),
                 material_id=mat_id, phong_gt=phong, phong=phong * rng.uniform(0.9, 1.1, phong.shape),
                 light_gt=light, light=light + (0 if directional else rng.normal(0, 0.05, 3)),
                 directional=bool(directional), intensity=inten, normal_obs=np.ascontiguousarray(nobs),
                 int_stiffness=1.0 / np.sqrt(int_var), W_normal=(np.eye(3) / np.sqrt(normal_var)).reshape(9))
    if shared_textures:
        track.update(tex_shared_gt=tex_shared, tex_shared=np.clip(tex_shared + rng.normal(0, 0.05, n_materials), 0, 1),
                     texture_id=mat_id.copy())
        track["textures"] = track["tex_shared"][mat_id]
    return track


def build_phong_problem(track, bounds=False, problem_cls=None, **options):
    """dataset_ba_phong's problem (stereo + intensity + normal blocks, first pose constant).  With a
    track made with `shared_textures` the texture blocks are shared per material
    (cslam_set_textures) — the shape the joint solve takes; `bounds` adds the box constraints of
    dataset_ba_phong.cpp:143-181."""
    p, poses, points = build_problem(track, problem_cls=problem_cls, **options)
    normals, textures = p.set_vertices(track["normals"].copy(), track["textures"].copy(), track["material_id"])
    if "tex_shared" in track:
        textures = p.set_textures(track["tex_shared"].copy(), track["texture_id"])
    phong = p.set_materials(track["phong"].copy())
    light = p.set_light(track["light"].copy(), track["directional"])
    p.add_phong(track["obs_cam"], track["obs_pt"], track["intensity"], track["int_stiffness"],
                track["normal_obs"], track["W_normal"])
    if bounds:
        p.set_bounds("material", [0.0, 0.0, 1.0], [1.0, 1.0, np.inf])
        p.set_bounds("texture", [0.0], [1.0])
    return p, dict(poses=poses, points=points, normals=normals, textures=textures, phong=phong, light=light)


def window_of(track, k1, k2):
    """The sub-problem `solveWindow(dataset, k1, k2)` builds (dataset_vo.cpp:40-62): observations
    of poses k1..k2-1, restricted to points seen by at least two poses of the window (the
    reference only optimises points its two-frame RANSAC initialised), re-indexed compactly."""
    sel = (track["obs_cam"] >= k1) & (track["obs_cam"] < k2)
    cam = track["obs_cam"][sel].astype(np.int64) - k1
    pt = track["obs_pt"][sel].astype(np.int64)
    uvd = track["uvd"][sel]
    W = track["W"] if track["W"].size == 9 else track["W"][sel]
    cnt = np.bincount(pt, minlength=track["n_points"])
    keep = cnt[pt] >= 2
    cam, pt, uvd = cam[keep], pt[keep], uvd[keep]
    if W.size != 9:
        W = W[keep]
    ids, pt_local = np.unique(pt, return_inverse=True)
    constant = np.zeros(k2 - k1, dtype=np.uint8)
    constant[0] = 1
    out = dict(cam=track["cam"], poses=track["poses"][k1:k2].copy(), poses_gt=track["poses_gt"][k1:k2],
               points=track["points"][ids].copy(), points_gt=track["points_gt"][ids],
               obs_cam=cam.astype(np.uint32), obs_pt=pt_local.astype(np.uint32),
               uvd=np.ascontiguousarray(uvd), W=np.ascontiguousarray(W), constant=constant,
               n_poses=k2 - k1, n_points=ids.size, k1=k1)
    if "sun_cam" in track:
        out["sun_cam"] = np.arange(k2 - k1, dtype=np.uint32)
        out["sun_obs_c"] = track["sun_obs_c"][k1:k2].copy()
        out["sun_ref_g"] = track["sun_ref_g"][k1:k2].copy()
        out["sun_W"] = track["sun_W"][k1:k2].copy()
    return out


def build_problem(track, sun=False, prior=None, huber=0.0, hold_first=True, problem_cls=None, **options):
    """Assemble the problem the way the reference drivers do (dataset_vo.cpp:40-62 /
    dataset_vo_sun.cpp:49-129).  Returns (BAProblem, poses array, points array)."""
    from .problem import BAProblem
    p = (problem_cls or BAProblem)(**options)
    c = track["cam"]
    p.set_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["b"])
    const = track["constant"] if hold_first else np.zeros(track["n_poses"], dtype=np.uint8)
    poses = p.set_poses(track["poses"].copy(), const)
    points = p.set_points(track["points"].copy())
    p.add_stereo(track["obs_cam"], track["obs_pt"], track["uvd"], track["W"])
    if sun:
        p.add_sun(track["sun_cam"], track["sun_obs_c"], track["sun_ref_g"], track["sun_W"], huber=huber)
    if prior is not None:
        cam, Tref, W6 = prior
        p.add_pose_prior(cam, Tref, W6)
    return p, poses, points


def write_track_csv(track, path):
    """The reference's plain track CSV (src/ceres_slam/dataset_problem.cpp:27-83): counts,
    intrinsics, variances, first pose as a 4x4 row-major matrix, then rows k,j,u,v,d grouped by k."""
    c = track["cam"]
    W = np.asarray(track["W"]).reshape(-1)[:9].reshape(3, 3)
    var = 1.0 / np.diag(W) ** 2
    T0 = np.eye(4)
    T0[:3, :3] = track["poses_gt"][0, 3:].reshape(3, 3)
    T0[:3, 3] = track["poses_gt"][0, :3]
    with open(path, "w") as f:
        f.write(f"{track['n_poses']},{track['n_points']}\n")
        f.write(",".join(repr(float(c[k])) for k in ("fu", "fv", "cu", "cv", "b")) + "\n")
        f.write(",".join(repr(float(v)) for v in var) + "\n")
        f.write(",".join(repr(float(v)) for v in T0.reshape(-1)) + "\n")
        for k, j, z in zip(track["obs_cam"], track["obs_pt"], track["uvd"]):
            f.write(f"{int(k)},{int(j)},{z[0]!r},{z[1]!r},{z[2]!r}\n".replace("np.float64(", "").replace(")", ""))


def _fmt(values):
    return ",".join(repr(float(v)) for v in values)


def _first_pose_row(track):
    T0 = np.eye(4)
    T0[:3, :3] = track["poses_gt"][0, 3:].reshape(3, 3)
    T0[:3, 3] = track["poses_gt"][0, :3]
    return _fmt(T0.reshape(-1))


def write_sun_csvs(track, track_path, ref_sun_path, obs_sun_path):
    """The three files of dataset_vo_sun (dataset_problem_sun.cpp:33-175): track rows
    `k,j,u,v,d,c00..c22` (no variance line), ephemeris `k,e,n,u`, observed sun
    `k,x,y,z,c00,c01,c10,c11`.  Covariances are the inverse squares of the stiffness matrices."""
    c = track["cam"]
    W = np.asarray(track["W"])
    W = np.tile(W.reshape(1, 9), (track["obs_cam"].size, 1)) if W.size == 9 else W.reshape(-1, 9)
    with open(track_path, "w") as f:
        f.write(f"{track['n_poses']},{track['n_points']}\n")
        f.write(_fmt(c[k] for k in ("fu", "fv", "cu", "cv", "b")) + "\n")
        f.write(_first_pose_row(track) + "\n")
        for k, j, z, w in zip(track["obs_cam"], track["obs_pt"], track["uvd"], W):
            Wm = w.reshape(3, 3)
            cov = np.linalg.inv(Wm @ Wm)
            cov = 0.5 * (cov + cov.T)
            f.write(f"{int(k)},{int(j)},{_fmt(z)},{_fmt(cov.reshape(-1))}\n")
    with open(ref_sun_path, "w") as f:
        for k, e in zip(track["sun_cam"], track["sun_ref_g"]):
            f.write(f"{int(k)},{_fmt(e)}\n")
    with open(obs_sun_path, "w") as f:
        for k, o, w in zip(track["sun_cam"], track["sun_obs_c"], track["sun_W"]):
            Wm = w.reshape(2, 2)
            cov = np.linalg.inv(Wm @ Wm)
            f.write(f"{int(k)},{_fmt(o)},{_fmt(cov.reshape(-1))}\n")


def write_phong_csv(track, path, int_var=1e-4, normal_var=1e-4):
    """dataset_ba_phong's input (dataset_problem_phong.cpp:29-117): counts, intrinsics, the seven
    variances, the initial light position / direction, the first pose, then rows
    `t,j,mat_id,u,v,d,I,nx,ny,nz` grouped by timestamp."""
    c = track["cam"]
    W = np.asarray(track["W"]).reshape(-1)[:9].reshape(3, 3)
    var = 1.0 / np.diag(W) ** 2
    n_mat = track["phong"].shape[0]
    with open(path, "w") as f:
        f.write(f"{track['n_poses']},{track['n_points']},{n_mat}\n")
        f.write(_fmt(c[k] for k in ("fu", "fv", "cu", "cv", "b")) + "\n")
        f.write(_fmt(list(var) + [normal_var] * 3 + [int_var]) + "\n")
        f.write(_fmt(track["light"]) + "\n")
        f.write(_first_pose_row(track) + "\n")
        for k, j, z, inten, nobs in zip(track["obs_cam"], track["obs_pt"], track["uvd"], track["intensity"],
                                        track["normal_obs"]):
            f.write(f"{float(k)!r},{int(j)},{int(track['material_id'][j])},{_fmt(z)},{float(inten)!r},{_fmt(nobs)}\n")
