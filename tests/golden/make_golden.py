#!/usr/bin/env python
"""Generate tests/golden/functors_mp60.json: known-answer vectors for every residual functor and
plus operation on the hot path, computed INDEPENDENTLY of oracle/ and of the CUDA code in 60-digit
arithmetic (mpmath).

The reference's own tests hold no numeric expectations for this path and
neither Ceres nor Eigen is in the image, so the reference cannot be executed to produce vectors.
What "Ceres autodiff" returns is, mathematically, d r(Plus(x, delta)) / d delta at delta = 0 of the
functor as written in the reference headers.  This script restates each functor from those headers
in mpmath (citations below) and differentiates r(Plus(x, delta)) by central differences with
h = 1e-25 at 60 digits: truncation error O(h^2) = 1e-50, so the stored doubles are correctly
rounded.  No Jet type and no closed form is involved, which makes the file a third, independent
leg next to oracle/ (Jets) and csrc/closed_form.h (closed forms).

    python tests/golden/make_golden.py          # rewrites tests/golden/functors_mp60.json

Inputs are drawn as float64 (seeded) and converted exactly to mpf; outputs are rounded to float64.
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 60
H = mp.mpf(10) ** -25
EPS = mp.mpf(np.finfo(np.float64).eps)  # std::numeric_limits<double>::epsilon() branch tests
PI_D = mp.mpf(float(np.arctan(1.0) * 4.0))  # utils.hpp:14  `pi = std::atan(1.) * 4.`


def M(x):
    return [mp.mpf(float(v)) for v in np.asarray(x, dtype=np.float64).ravel()]


def F(x):
    return [float(v) for v in x]


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def norm(a):
    return mp.sqrt(dot(a, a))


def matvec(A, v, n):  # A row-major n x n
    return [sum(A[n * i + k] * v[k] for k in range(n)) for i in range(n)]


def matmul3(A, B):
    return [sum(A[3 * i + k] * B[3 * k + j] for k in range(3)) for i in range(3) for j in range(3)]


def wedge(p):  # so3group.hpp:248-254
    return [0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0]


def so3_exp(phi):  # so3group.hpp:272-292
    angle = norm(phi)
    I = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    if angle <= EPS:
        w = wedge(phi)
        return [I[i] + w[i] for i in range(9)]
    ax = [p / angle for p in phi]
    cp, sp = mp.cos(angle), mp.sin(angle)
    w = wedge(ax)
    return [cp * I[i] + (1 - cp) * ax[i // 3] * ax[i % 3] + sp * w[i] for i in range(9)]


def so3_log(C):  # so3group.hpp:299-349
    axis = [C[7] - C[5], C[2] - C[6], C[3] - C[1]]
    sin_angle = mp.mpf("0.5") * norm(axis)
    cos_angle = mp.mpf("0.5") * (C[0] + C[4] + C[8] - 1)
    angle = mp.atan2(sin_angle, cos_angle)
    if abs(angle) <= EPS:
        return [mp.mpf("0.5") * a for a in axis]  # vee(C - I), so3group.hpp:258-263
    return [mp.mpf("0.5") * angle * a / sin_angle for a in axis]


def se3_plus(x, d):  # perturbations.hpp:56-64: T' = exp(delta) * T, se3group.hpp:176-183,323-325
    t, R = x[:3], x[3:]
    E = so3_exp(d[3:])
    Rn = matmul3(E, R)
    tn = [a + b for a, b in zip(matvec(E, t, 3), d[:3])]
    return tn + Rn


def unit_plus(x, d):  # perturbations.hpp:97-104
    s = dot(d, x) / dot(x, x)
    y = [x[i] + d[i] - s * x[i] for i in range(3)]
    n = norm(y)
    return [v / n for v in y]


def ident_plus(x, d):
    return [a + b for a, b in zip(x, d)]


def project(cam, p):  # stereo_camera.hpp:77-108
    fu, fv, cu, cv, b = cam
    iz = 1 / p[2]
    return [fu * p[0] * iz + cu, fv * p[1] * iz + cv, fu * b * iz]


def transform_point(x, p):  # se3group.hpp:193
    return [a + b for a, b in zip(matvec(x[3:], p, 3), x[:3])]


def transform_vector(x, v):  # se3group.hpp:244
    return matvec(x[3:], v, 3)


# ---- functors ------------------------------------------------------------------------------
def stereo(cam, z, W):  # stereo_reprojection_error.hpp:27-56
    def f(pose, pt):
        e = [a - b for a, b in zip(project(cam, transform_point(pose, pt)), z)]
        return matvec(W, e, 3)
    return f


def sun(obs_c, ref_g, W, az_t, zen_t):  # sun_sensor_error.hpp:20-96
    no, ng = norm(obs_c), norm(ref_g)
    obs_c = [v / no for v in obs_c]
    ref_g = [v / ng for v in ref_g]

    def f(pose):
        e = transform_vector(pose, ref_g)
        ez, ea = mp.acos(-e[1]), mp.atan2(e[0], e[2])
        oz, oa = mp.acos(-obs_c[1]), mp.atan2(obs_c[0], obs_c[2])
        ra, rz = ea - oa, ez - oz
        if ra > PI_D:
            ra = ra - 2 * PI_D
        elif ra < -PI_D:
            ra = ra + 2 * PI_D
        if abs(ra) > az_t:
            ra = mp.mpf(0)
        if abs(rz) > zen_t:
            rz = mp.mpf(0)
        return [W[0] * ra + W[1] * rz, W[2] * ra + W[3] * rz]
    return f


def prior(Tref, W):  # pose_error.hpp:17-47; se3group.hpp:152 (inverse), :176 (product), :344 (log)
    def f(pose):
        t, R = pose[:3], pose[3:]
        Rt = [R[3 * j + i] for i in range(3) for j in range(3)]
        ti = [-v for v in matvec(Rt, t, 3)]
        Rr = matmul3(Tref[3:], Rt)
        tr = [a + b for a, b in zip(matvec(Tref[3:], ti, 3), Tref[:3])]
        xi = tr + so3_log(Rr)
        return matvec(W, xi, 6)
    return f


def normal_err(obs, W):  # normal_error.hpp:16-41
    def f(pose, n):
        e = [a - b for a, b in zip(transform_vector(pose, n), obs)]
        return matvec(W, e, 3)
    return f


def phong_shade(normal, kd, ks, alpha, light_dir, cam_dir):  # phong.hpp:25-139
    # diffuse (phong.hpp:60-74)
    ld_n = dot(light_dir, normal)
    diffuse = kd * ld_n if ld_n > 0 else mp.mpf(0)
    # specular (phong.hpp:77-103)
    m = [2 * ld_n * normal[i] - light_dir[i] for i in range(3)]
    spec = mp.mpf(0)
    if dot(m, m) > 0:
        nm = norm(m)
        m = [v / nm for v in m]
        s = dot(m, cam_dir)
        if s > 0:
            spec = ks * mp.power(s, alpha)
    col = diffuse + spec  # ambient disabled (phong.hpp:31-33), light colour 1
    col = mp.mpf(0) if 0 >= col else col  # fmax(0, col): constant wins ties (utils.hpp:16-19)
    col = mp.mpf(1) if 1 <= col else col  # fmin(1, col)
    return col


def intensity(colour, w, directional):  # intensity_error_{point,directional}_light.hpp:24-90
    def f(pose, pt, n, phong, tex, light):
        p_c = transform_point(pose, pt)
        n_c = transform_vector(pose, n)
        if directional:
            l_c = transform_vector(pose, light)
            nl = norm(l_c)  # DirectionalLight ctor normalises (directional_light.hpp:51-54)
            light_dir = [v / nl for v in l_c]
        else:
            l_c = transform_point(pose, light)
            lv = [a - b for a, b in zip(l_c, p_c)]  # point_light.hpp:79-81
            nl = norm(lv)
            light_dir = [v / nl for v in lv]
        cv = [-v for v in p_c]  # camera at the origin (intensity_error_point_light.hpp:83)
        nc = norm(cv)
        cam_dir = [v / nc for v in cv]
        col = phong_shade(n_c, tex[0], phong[1], phong[2], light_dir, cam_dir)
        return [w * (col - colour)]
    return f


# ---- differentiation through Plus ------------------------------------------------------------
def jac(f, blocks, pluses, sizes, which):
    """d f(..., Plus_which(x_which, delta), ...) / d delta at 0, central differences."""
    cols = []
    for c in range(sizes[which]):
        outs = []
        for sgn in (1, -1):
            d = [mp.mpf(0)] * sizes[which]
            d[c] = sgn * H
            args = list(blocks)
            args[which] = pluses[which](blocks[which], d)
            outs.append(f(*args))
        cols.append([(a - b) / (2 * H) for a, b in zip(*outs)])
    nres = len(cols[0])
    return [cols[c][r] for r in range(nres) for c in range(sizes[which])]  # row-major nres x size


def random_pose(rng, t_scale=5.0, ang=1.0):
    phi = rng.normal(size=3) * ang
    R = np.array(F(so3_exp(M(phi)))).reshape(3, 3)  # rounded to double: not exactly orthonormal, fine
    t = rng.normal(size=3) * t_scale
    return np.concatenate([t, R.ravel()])


def main():
    rng = np.random.default_rng(20261018)
    cam = [707.0912, 707.0912, 601.8873, 183.1104, 0.535105804]  # camera_test.cpp:11-15
    out = {"about": "60-digit mpmath restatement of the reference functors; Jacobians are "
                    "d r(Plus(x,delta))/d delta at 0 by central differences (h=1e-25). "
                    "Generated by tests/golden/make_golden.py.", "camera": cam,
           "stereo": [], "sun": [], "prior": [], "normal": [], "intensity": [], "se3_plus": [], "unit_plus": []}
    camm = M(cam)

    # stereo ------------------------------------------------------------------------------
    for i in range(12):
        pose = random_pose(rng)
        R, t = pose[3:].reshape(3, 3), pose[:3]
        p_c = np.array([rng.uniform(-4, 4), rng.uniform(-2, 2), rng.uniform(2.5, 40)])
        pt = R.T @ (p_c - t)
        z = np.array(F(project(camm, M(p_c)))) + rng.normal(size=3)
        if i % 2 == 0:
            W = np.diag(1.0 / np.sqrt(rng.uniform(0.5, 4.0, size=3)))
        else:
            A = rng.normal(size=(3, 3))
            cov = np.eye(3) + 0.1 * A @ A.T
            w, V = np.linalg.eigh(cov)
            W = V @ np.diag(w ** -0.5) @ V.T
        f = stereo(camm, M(z), M(W))
        blocks, pl, sz = [M(pose), M(pt)], [se3_plus, ident_plus], [6, 3]
        out["stereo"].append({"pose": F(blocks[0]), "point": F(blocks[1]), "uvd": F(M(z)), "W": F(M(W)),
                              "r": F(f(*blocks)), "J_pose": F(jac(f, blocks, pl, sz, 0)),
                              "J_point": F(jac(f, blocks, pl, sz, 1))})

    # sun ---------------------------------------------------------------------------------
    for i in range(10):
        pose = random_pose(rng)
        R = pose[3:].reshape(3, 3)
        ref_g = rng.normal(size=3) * 3.0  # un-normalised on purpose: the ctor normalises
        e_c = R @ (ref_g / np.linalg.norm(ref_g))
        az_t = zen_t = 1000.0
        if i == 6 or i == 7:
            # azimuth wrap-around: expected az near +-pi, observed on the other side
            sgn = 1.0 if i == 6 else -1.0
            az_e, az_o, zen = sgn * (np.pi - 0.05), -sgn * (np.pi - 0.08), 1.1
            want = np.array([np.sin(zen) * np.sin(az_e), -np.cos(zen), np.sin(zen) * np.cos(az_e)])
            # rotate the pose so that R ref = want: build R from two frames
            a = ref_g / np.linalg.norm(ref_g)
            v = np.cross(a, want)
            c = float(a @ want)
            vx = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
            R = np.eye(3) + vx + vx @ vx / (1 + c)
            pose = np.concatenate([pose[:3], R.ravel()])
            obs = np.array([np.sin(zen + 0.02) * np.sin(az_o), -np.cos(zen + 0.02), np.sin(zen + 0.02) * np.cos(az_o)])
        else:
            noise = rng.normal(size=3) * (0.03 if i < 8 else 0.6)
            obs = e_c + noise
        if i == 8:
            az_t, zen_t = 0.1, 1000.0  # azimuth error rejected -> constant 0
        if i == 9:
            az_t, zen_t = 1000.0, 0.05
        obs = obs * rng.uniform(0.5, 2.0)
        A = rng.normal(size=(2, 2))
        cov = np.eye(2) * (np.deg2rad(2.0) ** 2) + 1e-4 * A @ A.T
        w, V = np.linalg.eigh(cov)
        W = V @ np.diag(w ** -0.5) @ V.T
        f = sun(M(obs), M(ref_g), M(W), mp.mpf(az_t), mp.mpf(zen_t))
        blocks = [M(pose)]
        out["sun"].append({"pose": F(blocks[0]), "obs_c": F(M(obs)), "ref_g": F(M(ref_g)), "W": F(M(W)),
                           "az_thresh": az_t, "zen_thresh": zen_t, "r": F(f(*blocks)),
                           "J_pose": F(jac(f, blocks, [se3_plus], [6], 0))})

    # pose prior --------------------------------------------------------------------------
    for i in range(8):
        pose = random_pose(rng)
        if i < 2:
            Tref = pose.copy()  # residual identically 0: Taylor branch of log (so3group.hpp:329)
        else:
            d = np.concatenate([rng.normal(size=3) * 0.3, rng.normal(size=3) * (0.05 if i < 5 else 1.2)])
            Tref = np.array(F(se3_plus(M(pose), M(d))))
        if i % 2 == 0:
            W = np.diag([1e6, 1e6, 1e6, 1e3, 1e3, 1e3]).astype(float)  # Sigma0 = 1e-12 I style stiffness
        else:
            A = rng.normal(size=(6, 6))
            W = np.eye(6) * 10 + 0.5 * (A + A.T)
        f = prior(M(Tref), M(W))
        blocks = [M(pose)]
        out["prior"].append({"pose": F(blocks[0]), "Tref": F(M(Tref)), "W": F(M(W)), "r": F(f(*blocks)),
                             "J_pose": F(jac(f, blocks, [se3_plus], [6], 0))})

    # normal ------------------------------------------------------------------------------
    for i in range(6):
        pose = random_pose(rng)
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        obs = pose[3:].reshape(3, 3) @ n + rng.normal(size=3) * 0.05
        A = rng.normal(size=(3, 3))
        W = np.eye(3) * 5 + 0.3 * (A + A.T)
        f = normal_err(M(obs), M(W))
        blocks, pl, sz = [M(pose), M(n)], [se3_plus, unit_plus], [6, 3]
        out["normal"].append({"pose": F(blocks[0]), "normal": F(blocks[1]), "obs": F(M(obs)), "W": F(M(W)),
                              "r": F(f(*blocks)), "J_pose": F(jac(f, blocks, pl, sz, 0)),
                              "J_normal": F(jac(f, blocks, pl, sz, 1))})

    # intensity ---------------------------------------------------------------------------
    kinds = ["lit", "lit", "lit", "lit", "light_behind", "no_specular", "saturated", "lit",
             "dir", "dir", "dir", "dir_behind"]
    for i, kind in enumerate(kinds):
        directional = kind.startswith("dir")
        pose = random_pose(rng, t_scale=1.0, ang=0.4)
        R, t = pose[3:].reshape(3, 3), pose[:3]
        p_c = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(2, 6)])
        pt = R.T @ (p_c - t)
        # normal roughly facing the camera
        n_c = -p_c / np.linalg.norm(p_c) + rng.normal(size=3) * 0.25
        n_c /= np.linalg.norm(n_c)
        n = R.T @ n_c
        n /= np.linalg.norm(n)
        if directional:
            l_c = n_c + rng.normal(size=3) * 0.3
            if kind == "dir_behind":
                l_c = -l_c
            light = R.T @ l_c
            light /= np.linalg.norm(light)
        else:
            l_c = p_c + (n_c + rng.normal(size=3) * 0.3) * rng.uniform(1, 4)
            if kind == "light_behind":
                l_c = p_c - n_c * 2.0 + rng.normal(size=3) * 0.1
            light = R.T @ (l_c - t)
        ks, alpha, kd = rng.uniform(0.1, 0.5), rng.uniform(5, 30), rng.uniform(0.3, 0.9)
        if kind == "no_specular":
            # camera far off the mirror direction: tilt the normal away
            n_c2 = np.cross(p_c, [0.3, 1.0, 0.2])
            n_c2 /= np.linalg.norm(n_c2)
            n = R.T @ (0.2 * n_c + 0.98 * n_c2)
            n /= np.linalg.norm(n)
        if kind == "saturated":
            ks, kd, alpha = 0.95, 0.99, 1.0
        phong = np.array([rng.uniform(0, 0.2), ks, alpha])
        colour, w = rng.uniform(0.1, 0.9), 1.0 / np.sqrt(1e-4)
        f = intensity(mp.mpf(colour), mp.mpf(w), directional)
        blocks = [M(pose), M(pt), M(n), M(phong), M([kd]), M(light)]
        pl = [se3_plus, ident_plus, unit_plus, ident_plus, ident_plus, unit_plus if directional else ident_plus]
        sz = [6, 3, 3, 3, 1, 3]
        rec = {"kind": kind, "directional": int(directional), "pose": F(blocks[0]), "point": F(blocks[1]),
               "normal": F(blocks[2]), "phong": F(blocks[3]), "texture": F(blocks[4]), "light": F(blocks[5]),
               "colour": float(colour), "stiffness": float(w), "r": F(f(*blocks))}
        for k, name in enumerate(["J_pose", "J_point", "J_normal", "J_phong", "J_tex", "J_light"]):
            rec[name] = F(jac(f, blocks, pl, sz, k))
        rec["predicted"] = rec["r"][0] / float(w) + float(colour)
        out["intensity"].append(rec)

    # plus operations ---------------------------------------------------------------------
    for i in range(6):
        pose = random_pose(rng)
        d = np.concatenate([rng.normal(size=3) * 0.2, rng.normal(size=3) * [0.3, 1e-3, 1e-9, 1e-17, 2.5, 0.0][i]])
        out["se3_plus"].append({"pose": F(M(pose)), "delta": F(M(d)), "out": F(se3_plus(M(pose), M(d)))})
    for i in range(4):
        x = rng.normal(size=3)
        x /= np.linalg.norm(x)
        d = rng.normal(size=3) * [0.3, 1e-4, 1.5, 0.0][i]
        f = lambda xx: xx  # noqa: E731
        out["unit_plus"].append({"x": F(M(x)), "delta": F(M(d)), "out": F(unit_plus(M(x), M(d))),
                                 "J_plus": F(jac(f, [M(x)], [unit_plus], [3], 0))})

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "functors_mp60.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", path, {k: len(v) for k, v in out.items() if isinstance(v, list) and k != "camera"})


if __name__ == "__main__":
    main()
