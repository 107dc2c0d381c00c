"""The host front end PINNED to the reference's own sources (CPU only).

oracle/_ref compiles, unmodified and where they lie, the reference's
`src/ceres_slam/{point_cloud_aligner, dataset_problem, dataset_problem_sun, dataset_problem_phong}.cpp`
(against the Eigen stand-in; oracle/ref_capi.cpp wraps the three DatasetProblem* classes): the CSV readers,
`compute_initial_guess(k1, k2)` and `reset_points` — everything the reference's drivers do around `solveWindow`.

The product side is the HOST code of the three restated drivers (`ceres_slam_b200/host/dataset_*_b200.cpp`,
`dataset.hpp`, `sun_dataset.hpp`): tests/frontend_host.cpp includes each driver source as it is (`main` renamed) and
redirects its `cslam_ransac_align` calls to an injected alignment entry, so the whole front end runs here without a
GPU.  With the reference's own aligner injected, readers, reciprocal matching, triangulation, pose chaining, point /
vertex initialisation (incl. the Phong variant's normal, texture medians and its material-by-inlier-index quirk) must
reproduce the reference's state; with the oracle's restated aligner injected the same holds to 1e-9.  The CUDA
aligner itself is compared with the same reference source in tests/test_ref_pin.py (GPU).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ceres_slam_b200 import capi, synthetic as syn
from oracle import pybinding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = capi.dptr
u32 = capi.u32ptr
u8 = capi.u8ptr
_vp, _dp, _u32p, _u8p = C.c_void_p, capi._dp, capi._u32p, capi._u8p

needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built and /root/reference absent")

_SIGS = {
    "open": (_vp, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "close": (None, [_vp]),
    "dims": (None, [_vp, _u32p, _u32p, C.POINTER(C.c_uint64), _u32p]),
    "observations": (None, [_vp, _u32p, _u32p, _dp, _dp, _dp]),
    "sun_data": (None, [_vp, _dp, _u8p, _dp, _dp, _dp]),
    "phong_data": (None, [_vp, _u32p, _dp, _dp, _dp, _dp, _dp]),
    "initial_guess": (C.c_int, [_vp, C.c_uint32, C.c_uint32]),
    "reset_points": (None, [_vp]),
    "state": (None, [_vp, _dp, _dp, _u8p, _dp, _dp, _dp]),
}


class FrontEnd:
    """One dataset behind either library: the reference's DatasetProblem* (`cslam_ref_dataset_*`, kind as first
    argument of open) or the product's host code (`fh_*`, one shim per kind)."""

    def __init__(self, dll, prefix, kind, files, dir_light=False):
        self.kind = kind
        self.f = {}
        for name, (res, args) in _SIGS.items():
            if (name == "sun_data" and kind != 1) or (name == "phong_data" and kind != 2):
                continue
            fn = getattr(dll, prefix + name)
            fn.restype = res
            fn.argtypes = ([C.c_int] + args) if (name == "open" and prefix.startswith("cslam_ref")) else args
            self.f[name] = fn
        fs = [os.fsencode(x) for x in files] + [None] * (3 - len(files))
        a = ([kind] if prefix.startswith("cslam_ref") else []) + fs + [int(dir_light)]
        self.h = self.f["open"](*a)
        assert self.h, "reader failed"
        ns, npt, nm, no = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        self.f["dims"](self.h, C.byref(ns), C.byref(npt), C.byref(no), C.byref(nm))
        self.n_states, self.n_points, self.n_obs, self.n_materials = ns.value, npt.value, no.value, nm.value

    def close(self):
        self.f["close"](self.h)

    def observations(self):
        k, j = np.zeros(self.n_obs, np.uint32), np.zeros(self.n_obs, np.uint32)
        uvd, intr, var = np.zeros((self.n_obs, 3)), np.zeros(5), np.zeros(3)
        self.f["observations"](self.h, u32(k), u32(j), d(uvd), d(intr), d(var))
        out = dict(k=k, j=j, uvd=uvd, intr=intr, var=var)
        if self.kind == 1:
            out.update(covars=np.zeros((self.n_obs, 9)), has_sun=np.zeros(self.n_states, np.uint8),
                       sun_obs=np.zeros((self.n_states, 3)), sun_covar=np.zeros((self.n_states, 4)),
                       sun_dir=np.zeros((self.n_states, 3)))
            self.f["sun_data"](self.h, d(out["covars"]), u8(out["has_sun"]), d(out["sun_obs"]), d(out["sun_covar"]),
                               d(out["sun_dir"]))
            # states without an observation hold unspecified values in the reference's vectors
            m = out["has_sun"].astype(bool)
            for key in ("sun_obs", "sun_covar", "sun_dir"):
                out[key][~m] = 0.0
        if self.kind == 2:
            out.update(mat=np.zeros(self.n_obs, np.uint32), inten=np.zeros(self.n_obs), nobs=np.zeros((self.n_obs, 3)),
                       nvar=np.zeros(3), ivar=np.zeros(1), light=np.zeros(3))
            self.f["phong_data"](self.h, u32(out["mat"]), d(out["inten"]), d(out["nobs"]), d(out["nvar"]), d(out["ivar"]),
                                 d(out["light"]))
        return out

    def initial_guess(self, k1, k2):
        return self.f["initial_guess"](self.h, k1, k2)

    def reset_points(self):
        if self.kind != 2:      # DatasetProblemPhong has no reset_points: its vertices stay initialised
            self.f["reset_points"](self.h)

    def state(self):
        out = dict(poses=np.zeros((self.n_states, 12)), points=np.zeros((self.n_points, 3)),
                   init=np.zeros(self.n_points, np.uint8))
        extra = [None, None, None]
        if self.kind == 2:
            out.update(normals=np.zeros((self.n_points, 3)), phong=np.zeros((self.n_points, 3)), tex=np.zeros(self.n_points))
            extra = [d(out["normals"]), d(out["phong"]), d(out["tex"])]
        self.f["state"](self.h, d(out["poses"]), d(out["points"]), u8(out["init"]), *extra)
        return out


@pytest.fixture(scope="module")
def shims():
    """tests/_build/libfrontend_host_{0,1,2}.so: g++ over tests/frontend_host.cpp + the driver sources, linked against
    the product library (which loads without a GPU; no CUDA call is made here)."""
    capi.load_product()
    out = {}
    src = os.path.join(ROOT, "tests", "frontend_host.cpp")
    host = os.path.join(ROOT, "ceres_slam_b200", "host")
    csrc = os.path.join(ROOT, "ceres_slam_b200", "csrc")
    deps = [src, os.path.join(ROOT, "include", "cslam_b200.h")] + \
        [os.path.join(host, f) for f in os.listdir(host) if f.endswith((".cpp", ".hpp"))]
    os.makedirs(os.path.join(ROOT, "tests", "_build"), exist_ok=True)
    for kind in (0, 1, 2):
        so = os.path.join(ROOT, "tests", "_build", f"libfrontend_host_{kind}.so")
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", "-Wno-unused-variable",
                                   "-fPIC", "-shared", f"-DKIND={kind}", "-o", so, src, "-L" + csrc, "-lcslam_b200",
                                   "-Wl,-rpath," + csrc])
        out[kind] = C.CDLL(so)
        out[kind].fh_set_ransac_entry.argtypes = [_vp]
        out[kind].fh_set_ransac_entry.restype = None
    return out


def _entry_address(lib, name):
    return C.cast(getattr(lib.dll, lib.prefix + name), _vp)


def _noisy_track(n_poses=24, seed=31):
    """0.25 px noise and 12 % gross outliers, as in tests/test_ransac.py: RANSAC has something to reject."""
    tr = syn.make_track(n_poses, 12, 8, seed=seed, pix_sigma=0.25)
    rng = np.random.default_rng(seed + 1)
    bad = rng.random(tr["uvd"].shape[0]) < 0.12
    tr["uvd"][bad] += rng.normal(0, 25.0, (int(bad.sum()), 3))
    tr["uvd"][:, 2] = np.maximum(tr["uvd"][:, 2], 1.0)
    return tr


def _files(kind, tmp_path, directional=False):
    tr = _noisy_track(seed=31 + kind)
    if kind == 0:
        f = [str(tmp_path / "track.csv")]
        syn.write_track_csv(tr, f[0])
    elif kind == 1:
        tr = syn.add_sun(syn.make_track(24, 12, 8, seed=32, pix_sigma=0.25, per_obs_W=True))
        f = [str(tmp_path / "track.csv"), str(tmp_path / "ref_sun.csv"), str(tmp_path / "obs_sun.csv")]
        syn.write_sun_csvs(tr, *f)
    else:
        tr = syn.add_phong(tr, directional=directional, shared_textures=True)
        f = [str(tmp_path / "phong.csv")]
        syn.write_phong_csv(tr, f[0])
    return tr, f


def _compare_tables(a, b):
    assert set(a) == set(b)
    for key in a:
        assert np.array_equal(a[key], b[key]), key


def _compare_states(a, b, tol, n_reached):
    """States no window has reached yet are left out: the reference's vector holds default-constructed poses there,
    the product's readers fill every pose with the first one (neither is ever read before a window sets it)."""
    assert np.array_equal(a["init"], b["init"])
    worst = 0.0
    for key in a:
        if key == "init":
            continue
        x, y = (a[key][:n_reached], b[key][:n_reached]) if key == "poses" else (a[key], b[key])
        worst = max(worst, float(np.abs(x - y).max()))
    assert worst <= tol, worst
    return worst


def _usable_states(tr):
    """Number of leading states whose consecutive pairs all share at least 3 points.  Below 3 matches the reference's
    draw loop (point_cloud_aligner.cpp:83-90: three DISTINCT indices) never terminates, so its front end cannot be run
    there; the synthetic track thins out over its last frames."""
    from ceres_slam_b200 import initial_guess as ig
    rg = ig.state_ranges(tr["obs_cam"], tr["n_poses"])
    pt = tr["obs_pt"].astype(np.int64)
    for k in range(1, tr["n_poses"]):
        kp, kc = ig.match_pair(pt[rg[k - 1]:rg[k]], pt[rg[k]:rg[k + 1]])
        if min(kp.size, kc.size) < 3:
            return k
    return tr["n_poses"]


def _window_plan(n):
    """What the drivers do around solveWindow, minus the solve: one batch over the usable states, then sliding windows
    of 2 and of 5 with reset_points after each (dataset_vo.cpp:118-131, dataset_vo_sun.cpp:262-289,
    dataset_ba_phong.cpp:316-330)."""
    plan = [(0, n, False)]
    plan += [(k1, k1 + 2, True) for k1 in range(0, n - 1)]
    plan += [(k1, k1 + 5, True) for k1 in range(0, n - 4, 3)]
    return plan


@needs_ref
@pytest.mark.parametrize("kind,directional", [(0, False), (1, False), (2, False), (2, True)])
def test_host_readers_match_reference_readers(shims, tmp_path, kind, directional):
    """Every field the reference's read_csv stores, against the product's readers, exactly (both parse the same text
    with correctly rounded conversions), and against the arrays the file was written from."""
    tr, files = _files(kind, tmp_path, directional)
    ref = FrontEnd(orc.load_ref().dll, "cslam_ref_dataset_", kind, files, directional)
    host = FrontEnd(shims[kind], "fh_", kind, files, directional)
    assert (ref.n_states, ref.n_points, ref.n_obs, ref.n_materials) == (host.n_states, host.n_points, host.n_obs, host.n_materials)
    assert ref.n_states == tr["n_poses"] and ref.n_obs == tr["obs_cam"].size
    tr_ref, tr_host = ref.observations(), host.observations()
    _compare_tables(tr_ref, tr_host)
    assert np.array_equal(tr_ref["k"], tr["obs_cam"]) and np.array_equal(tr_ref["j"], tr["obs_pt"])
    assert np.array_equal(tr_ref["uvd"], tr["uvd"])
    s_ref, s_host = ref.state(), host.state()
    assert np.array_equal(s_ref["poses"][0], tr["poses_gt"][0]) and np.array_equal(s_host["poses"][0], tr["poses_gt"][0])
    if kind == 2 and directional:
        assert abs(np.linalg.norm(tr_ref["light"]) - 1.0) < 1e-15      # light_dir.normalize()
    ref.close()
    host.close()


@needs_ref
@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind,directional", [(0, False), (1, False), (2, False), (2, True)])
@pytest.mark.parametrize("aligner", ["reference", "oracle"])
def test_host_front_end_matches_reference_front_end(shims, tmp_path, kind, directional, aligner):
    """compute_initial_guess(k1, k2) over the drivers' window sequences: the reference's DatasetProblem* against the
    product's host code.  `reference`: both sides align with the reference's own point_cloud_aligner.cpp, so any
    difference is the host logic's — tolerance 1e-12 (the chaining products are written in a different order).
    `oracle`: the product side aligns with the restated RANSAC — 1e-9."""
    tr, files = _files(kind, tmp_path, directional)
    lib = orc.load_ref() if aligner == "reference" else orc.load_oracle()
    shims[kind].fh_set_ransac_entry(_entry_address(lib, "ransac_align"))
    ref = FrontEnd(orc.load_ref().dll, "cslam_ref_dataset_", kind, files, directional)
    host = FrontEnd(shims[kind], "fh_", kind, files, directional)
    tol = 1e-12 if aligner == "reference" else 1e-9
    worst, n_init, n_use = 0.0, 0, _usable_states(tr)
    for k1, k2, reset in _window_plan(n_use):
        ok_r, ok_h = ref.initial_guess(k1, k2), host.initial_guess(k1, k2)
        assert ok_r == ok_h
        if not ok_r:
            break    # the sun variant gave up; its driver copies the previous pose (dataset_vo_sun.cpp:283-288)
        s_ref, s_host = ref.state(), host.state()
        worst = max(worst, _compare_states(s_ref, s_host, tol, n_use))
        n_init = max(n_init, int(s_ref["init"].sum()))
        if reset:
            ref.reset_points()
            host.reset_points()
    assert n_init > 50            # the comparison saw initialised points, not empty tables
    s = ref.state()
    assert np.abs(s["poses"][1:n_use] - s["poses"][0]).max() > 1e-3   # and poses that moved
    print(f"kind {kind} aligner {aligner}: worst difference {worst:.3g}, up to {n_init} initialised points")
    ref.close()
    host.close()


@needs_ref
@pytest.mark.timeout(300)
def test_sun_front_end_gives_up_below_three_inliers(shims, tmp_path):
    """DatasetProblemSun::compute_initial_guess returns false when a pair has fewer than 3 inliers
    (dataset_problem_sun.cpp:323-326); the host code must stop at the same pair with the same state."""
    tr = syn.add_sun(syn.make_track(24, 12, 8, seed=35, pix_sigma=0.25, per_obs_W=True))
    # wreck every observation of state 12: no hypothesis keeps 3 inliers between states 11 and 12
    rng = np.random.default_rng(1)
    m = tr["obs_cam"] == 12
    tr["uvd"][m] += rng.normal(0, 300.0, (int(m.sum()), 3))
    tr["uvd"][:, 2] = np.maximum(tr["uvd"][:, 2], 1.0)
    f = [str(tmp_path / "track.csv"), str(tmp_path / "ref_sun.csv"), str(tmp_path / "obs_sun.csv")]
    syn.write_sun_csvs(tr, *f)
    shims[1].fh_set_ransac_entry(_entry_address(orc.load_ref(), "ransac_align"))
    ref = FrontEnd(orc.load_ref().dll, "cslam_ref_dataset_", 1, f)
    host = FrontEnd(shims[1], "fh_", 1, f)
    assert _usable_states(tr) >= 16
    assert ref.initial_guess(0, 16) == 0 and host.initial_guess(0, 16) == 0
    s_ref, s_host = ref.state(), host.state()
    assert np.array_equal(s_ref["init"], s_host["init"]) and s_ref["init"].sum() > 20
    assert np.abs(s_ref["poses"][:12] - s_host["poses"][:12]).max() <= 1e-12      # states 1..11 were chained
    assert np.abs(s_ref["points"] - s_host["points"]).max() <= 1e-12
    assert ref.initial_guess(0, 12) == 1 and host.initial_guess(0, 12) == 1
    ref.close()
    host.close()


def _read_table(path):
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    header = [h.strip() for h in lines[0].split(",")]
    rows = np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]]) if len(lines) > 1 else np.zeros((0, len(header)))
    return header, rows


@needs_ref
@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind,directional", [(0, False), (1, False), (2, False), (2, True)])
def test_host_writers_match_reference_writers(shims, tmp_path, kind, directional):
    """write_csv (dataset_problem.cpp:125-166, dataset_problem_sun.cpp:181-204, dataset_problem_phong.cpp:175-232)
    against the drivers' output: the same files (<stem>_poses.csv, _map.csv, _lights.csv), the same header lines, the
    same rows.  The reference prints 4 significant digits (utils.hpp:34 CommaInitFmt), the product 17: values agree
    to the reference's rounding."""
    tr, files = _files(kind, tmp_path, directional)
    lib = orc.load_ref()
    shims[kind].fh_set_ransac_entry(_entry_address(lib, "ransac_align"))
    shims[kind].fh_write.argtypes = [_vp, C.c_char_p]
    shims[kind].fh_write.restype = None
    lib.dll.cslam_ref_dataset_write.argtypes = [_vp, C.c_char_p]
    lib.dll.cslam_ref_dataset_write.restype = C.c_int
    ref = FrontEnd(lib.dll, "cslam_ref_dataset_", kind, files, directional)
    host = FrontEnd(shims[kind], "fh_", kind, files, directional)
    n = _usable_states(tr)
    assert ref.initial_guess(0, n) == 1 and host.initial_guess(0, n) == 1      # a state worth writing
    (tmp_path / "r").mkdir()
    (tmp_path / "h").mkdir()
    assert lib.dll.cslam_ref_dataset_write(ref.h, os.fsencode(str(tmp_path / "r" / "out.csv"))) == 1
    shims[kind].fh_write(host.h, os.fsencode(str(tmp_path / "h" / "out.csv")))
    names = sorted(os.listdir(tmp_path / "r"))
    assert names == sorted(os.listdir(tmp_path / "h"))
    assert names == {0: ["out_map.csv", "out_poses.csv"], 1: ["out_poses.csv"],
                     2: ["out_lights.csv", "out_map.csv", "out_poses.csv"]}[kind]
    for name in names:
        hr, rr = _read_table(tmp_path / "r" / name)
        hh, rh = _read_table(tmp_path / "h" / name)
        assert hr == hh, name
        if name == "out_poses.csv":      # states beyond the batch: default-constructed vs first pose (see above)
            rr, rh = rr[:n], rh[:n]
        assert rr.shape == rh.shape and rr.shape[0] > 0, name
        assert np.all(np.abs(rr - rh) <= 5.01e-4 * np.abs(rh) + 1e-12), name
    ref.close()
    host.close()
