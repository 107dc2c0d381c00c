"""The reference's OWN drivers, unmodified, on this repo's back end (all three configs: the
second part of the file is tests/dataset_vo_sun.cpp — Huber loss, pose prior, SUBSPACE_DOGLEG, ceres::Covariance — the
third compares the C ABI call streams of tests/dataset_ba_phong.cpp / dataset_vo.cpp with the restated drivers').

`oracle/_ref/libref_dataset_vo_{oracle,b200}.so` are /root/reference/tests/dataset_vo.cpp compiled as it is (its
`main` renamed) together with the reference's dataset_problem.cpp / point_cloud_aligner.cpp, against
`oracle/ref_driver/ceres/ceres.h`: a facade with Ceres' API (Problem, AddResidualBlock, SetParameterization,
SetParameterBlockConstant, Solver::Options, Solve) that states the problem through `cslam_b200::Problem` and the C ABI.
Everything around the solve — CSV reader, window loop, compute_initial_guess with the reference's RANSAC,
solveWindow's block list, stiffness from `SelfAdjointEigenSolver::operatorInverseSqrt`, constant first pose, solver
options, reset_points — is the reference's code.

CPU: the `_oracle` build answers the C ABI calls with the CPU oracle.  Its window-by-window results must equal
oracle/driver_mirror.dataset_vo — the Python restatement of this driver that the GPU parity tests of the restated C++
driver (tests/test_driver_sequences.py) are anchored on — so that anchor is itself pinned to the reference's driver.
GPU: the `_b200` build answers them with the CUDA library: the reference's driver source running on the B200 back end
with no change, compared with the same mirror at BASELINE.json's 1e-6.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from ceres_slam_b200 import synthetic as syn
from oracle import driver_mirror as dm
from oracle import pybinding as orc
from test_gpu_parity import _steady_track

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _lib(kind, driver="dataset_vo"):
    so = os.path.join(REF_DIR, f"libref_{driver}_{kind}.so")
    if not os.path.exists(so) and os.path.isdir(orc.REFERENCE_INCLUDE):
        orc.build_ref()
    if not os.path.exists(so):
        pytest.skip(f"{so} not built and /root/reference absent")
    return so


_RUNNER = r"""
import ctypes as C, sys
lib = C.CDLL(sys.argv[1])
main = getattr(lib, "cslam_ref_" + sys.argv[2] + "_main")
main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
args = [sys.argv[2].encode()] + [a.encode() for a in sys.argv[3:]]
sys.exit(main(len(args), (C.c_char_p * len(args))(*args)))
"""


def _run_reference_driver(so, csv, window, trace, driver="dataset_vo", extra=(), timeout=600):
    """In a child process: the driver narrates on stdout / stderr, and a crash must not take the test session down.
    `csv`: the input file, or the list of input files."""
    env = dict(os.environ, CSLAM_FACADE_TRACE=trace)
    if os.path.exists(trace):
        os.remove(trace)
    files = [csv] if isinstance(csv, str) else list(csv)
    r = subprocess.run([sys.executable, "-c", _RUNNER, so, driver] + files + ["--window", str(window)] + list(extra),
                       env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(trace) as f:
        return [json.loads(line) for line in f], r.stdout


def _compare_with_mirror(tr, rows, window, tol, max_iteration_mismatches):
    n = tr["n_poses"]
    w = n if window == 0 else window
    var = 1.0 / np.diag(np.asarray(tr["W"]).reshape(3, 3)) ** 2
    its, costs, per_window = [], [], []

    def on_window(k1, s):
        its.append(s.num_iterations)
        costs.append((s.initial_cost, s.final_cost))

    To = dm.dataset_vo(tr, var, tr["poses_gt"][0], window, 1000, on_window=on_window)   # 1000: dataset_vo.cpp:69
    assert len(rows) == len(its) == n - w + 1
    P = np.zeros((n, 12))
    for k1, r in enumerate(rows):
        assert r["n_poses"] == w
        P[k1:k1 + w] = np.array(r["poses"]).reshape(-1, 12)
    bad = sum(r["iterations"] != i for r, i in zip(rows, its))
    assert bad <= max_iteration_mismatches, ([r["iterations"] for r in rows], its)
    for r, (c0, c1) in zip(rows, costs):
        assert abs(r["initial_cost"] - c0) <= tol * max(c0, 1e-300)
        assert abs(r["final_cost"] - c1) <= max(tol, 1e-9) * max(c1, 1e-300)
    err = float(np.abs(P - To).max())
    assert err <= tol * np.abs(To).max(), err
    return err


@pytest.mark.parametrize("window", [2, 4, 0])
def test_reference_driver_on_oracle_equals_driver_mirror(tmp_path, window):
    """The reference's driver + facade + oracle against the Python mirror of that driver on the same oracle: the same
    windows, blocks, options and iteration counts, poses to 1e-11 (the two differ only in how the stiffness's inverse
    square root and the chained products are rounded)."""
    so = _lib("oracle")
    tr = _steady_track(16, seed=17)
    csv = str(tmp_path / "track.csv")
    syn.write_track_csv(tr, csv)
    rows, out = _run_reference_driver(so, csv, window, str(tmp_path / "trace.jsonl"))
    assert out.count("Report:") == len(rows)                 # summary.BriefReport() of every window (dataset_vo.cpp:82)
    assert os.path.exists(tmp_path / "track_poses.csv") and os.path.exists(tmp_path / "track_map.csv")   # write_csv
    err = _compare_with_mirror(tr, rows, window, 1e-11, 0)
    print(f"window {window}: reference driver on the oracle vs mirror, worst pose difference {err:.3g}")


@pytest.mark.gpu
@pytest.mark.parametrize("window", [2, 0])
def test_reference_driver_on_b200_matches_driver_mirror(product, tmp_path, window):
    """The reference's tests/dataset_vo.cpp, unmodified, with ceres::Solve answered by the CUDA library (measured on a
    B200 through scripts/ref_driver_on_b200.py, the same comparison: 4.8e-14, profiles/r02_reference_drivers_on_b200.log)."""
    so = _lib("b200")
    tr = _steady_track(30, seed=17)
    csv = str(tmp_path / "track.csv")
    syn.write_track_csv(tr, csv)
    rows, _ = _run_reference_driver(so, csv, window, str(tmp_path / "trace.jsonl"), timeout=120)
    _compare_with_mirror(tr, rows, window, 1e-6, max(1, len(rows) // 10))


# ---- config 2: tests/dataset_vo_sun.cpp (both passes, Huber loss, pose prior, SUBSPACE_DOGLEG, covariance chain) ----
def _sun_case(n, tmp_path):
    tr = syn.add_sun(_steady_track(n, seed=23, per_obs_W=True), sigma_deg=1.0)
    paths = [str(tmp_path / f) for f in ("track.csv", "sun_ref.csv", "sun_obs.csv")]
    syn.write_sun_csvs(tr, *paths)
    # the mirror reads what the driver reads: covariances are the CSV's (inverse squares of the stiffness)
    W = np.asarray(tr["W"]).reshape(-1, 3, 3)
    cov = np.stack([0.5 * (c + c.T) for c in (np.linalg.inv(w @ w) for w in W)]).reshape(-1, 9)
    sW = tr["sun_W"].reshape(-1, 2, 2)
    sun = dict(dir_g=tr["sun_ref_g"], obs=tr["sun_obs_c"], covars=np.stack([np.linalg.inv(w @ w) for w in sW]).reshape(-1, 4),
               has=np.ones(n, dtype=bool))
    return tr, paths, cov, sun


def _compare_sun_with_mirror(tr, cov, sun, rows, tol, cov_tol, max_iteration_mismatches):
    n = tr["n_poses"]
    solves = [r for r in rows if "poses" in r]
    covs = [r for r in rows if "covariance" in r]
    assert len(solves) == len(covs) == 2 * (n - 1)           # pass 1 (VO) and pass 2 (with the sun blocks)
    kw = dict(window=2, max_iters=1000, dogleg=True)         # dataset_vo_sun.cpp:140-143
    its = []
    p1, c1 = dm.dataset_vo_sun(tr, cov, sun, tr["poses_gt"][0], use_sun=False, on_window=lambda k1, s: its.append(s.num_iterations), **kw)
    p2, c2 = dm.dataset_vo_sun(tr, cov, sun, tr["poses_gt"][0], use_sun=True, huber=1.0, poses=p1.copy(), pose_covars=c1.copy(),
                               on_window=lambda k1, s: its.append(s.num_iterations), **kw)
    bad = sum(r["iterations"] != i for r, i in zip(solves, its))
    assert bad <= max_iteration_mismatches, ([r["iterations"] for r in solves], its)
    worst = 0.0
    for q, (pm, cm) in enumerate(((p1, c1), (p2, c2))):
        P, Cv = np.zeros((n, 12)), np.zeros((n, 36))
        for k1 in range(n - 1):
            P[k1:k1 + 2] = np.array(solves[q * (n - 1) + k1]["poses"]).reshape(2, 12)
            Cv[k1 + 1] = covs[q * (n - 1) + k1]["covariance"]
        err = float(np.abs(P - pm).max())
        assert err <= tol * np.abs(pm).max(), (q, err)
        cerr = float(np.abs(Cv[1:] - cm[1:]).max() / np.abs(cm[1:]).max())
        assert cerr <= cov_tol, (q, cerr)
        worst = max(worst, err)
    return worst


def test_reference_sun_driver_on_oracle_equals_driver_mirror(tmp_path):
    """tests/dataset_vo_sun.cpp, unmodified: per window the stereo blocks with per-observation stiffness (indexed by
    point id, as the reference does), the sun block under a Huber loss, the prior from the previous window's
    covariance, SUBSPACE_DOGLEG, then ceres::Covariance of the second pose.  In this build the covariance is formed
    inside the facade from the REFERENCE'S functors (autodiff Jacobians, dense Cholesky), the mirror forms it from the
    oracle's Jacobians through a sparse LU: poses to 1e-10, covariances to 1e-9 relative over both passes."""
    so = _lib("oracle", "dataset_vo_sun")
    tr, paths, cov, sun = _sun_case(16, tmp_path)
    rows, _ = _run_reference_driver(so, paths, 2, str(tmp_path / "trace.jsonl"), "dataset_vo_sun", ["--huber-param", "1.0"])
    assert os.path.exists(tmp_path / "track_poses.csv") and os.path.exists(tmp_path / "track_obs_poses.csv")   # :301, :322
    err = _compare_sun_with_mirror(tr, cov, sun, rows, 1e-10, 1e-9, 0)
    print(f"reference sun driver on the oracle vs mirror: worst pose difference {err:.3g}")


@pytest.mark.gpu
def test_reference_sun_driver_on_b200_matches_driver_mirror(product, tmp_path):
    """The reference's tests/dataset_vo_sun.cpp, unmodified, with ceres::Solve and ceres::Covariance answered by the
    CUDA library (window kernel with the DOGLEG loop on the device, cslam_covariance_block); 1.3e-12 on a B200
    (profiles/r02_reference_drivers_on_b200.log)."""
    so = _lib("b200", "dataset_vo_sun")
    tr, paths, cov, sun = _sun_case(25, tmp_path)
    rows, _ = _run_reference_driver(so, paths, 2, str(tmp_path / "trace.jsonl"), "dataset_vo_sun", ["--huber-param", "1.0"],
                                    timeout=120)
    _compare_sun_with_mirror(tr, cov, sun, rows, 1e-6, 1e-5, 5)


# ---- the C ABI call streams of the reference's drivers and of this repo's restated drivers -------------------------------
# Both end in the same boundary: include/cslam_b200.h.  With every call recorded on its way to the CPU oracle
# (oracle/ref_driver/abi_trace.cpp) two drivers that assemble the same problems — blocks in the same order, stiffness,
# constant flags, bounds, options, and (config 3) vertex / material / texture tables, across every window and stage —
# leave the same stream.  Left: the reference's driver source over the Ceres-API facade.  Right: the product's driver
# source (tests/frontend_host.cpp includes it as it is; its RANSAC calls answered by the reference's own aligner).
_RUN_HOST = r"""
import ctypes as C, sys
ref = C.CDLL(sys.argv[2])
lib = C.CDLL(sys.argv[1])
lib.fh_set_ransac_entry.argtypes = [C.c_void_p]
lib.fh_set_ransac_entry(C.cast(ref.cslam_ref_ransac_align, C.c_void_p))
lib.fh_driver_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
args = [b"driver"] + [a.encode() for a in sys.argv[3:]]
sys.exit(lib.fh_driver_main(len(args), (C.c_char_p * len(args))(*args)))
"""


def _host_trace_lib(kind):
    from ceres_slam_b200 import capi
    capi.load_product()
    name = {0: "dataset_vo", 1: "dataset_vo_sun", 2: "dataset_ba_phong"}[kind]
    so = os.path.join(ROOT, "tests", "_build", f"libdriver_{name}_trace.so")
    host = os.path.join(ROOT, "ceres_slam_b200", "host")
    srcs = [os.path.join(ROOT, "tests", "frontend_host.cpp"), os.path.join(ROOT, "oracle", "ref_driver", "abi_trace.cpp")]
    deps = srcs + [os.path.join(ROOT, "oracle", "ref_driver", "abi_remap.h"), os.path.join(ROOT, "include", "cslam_b200.h")] + \
        [os.path.join(host, f) for f in os.listdir(host) if f.endswith((".cpp", ".hpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        orc.build_oracle()
        ob, cs = os.path.join(ROOT, "oracle", "_build"), os.path.join(ROOT, "ceres_slam_b200", "csrc")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", "-Wno-unused-variable", "-fPIC", "-shared",
                               f"-DKIND={kind}", "-DFH_ABI_TRACE", "-o", so] + srcs +
                              ["-L" + ob, "-loracle", "-Wl,-rpath," + ob, "-L" + cs, "-lcslam_b200", "-Wl,-rpath," + cs])
    return so


def _abi_streams(driver, kind, write_input, flags, tmp_path):
    ref_so, host_so = _lib("trace", driver), _host_trace_lib(kind)
    aligner = os.path.join(REF_DIR, "libcslam_ref.so")
    out = []
    for side, cmd in (("ref", [sys.executable, "-c", _RUNNER, ref_so, driver]), ("host", [sys.executable, "-c", _RUN_HOST, host_so, aligner])):
        d = tmp_path / side
        d.mkdir()
        files = write_input(d)
        trace = str(d / "abi.jsonl")
        r = subprocess.run(cmd + files + list(flags), env=dict(os.environ, CSLAM_ABI_TRACE=trace), capture_output=True, text=True,
                           timeout=900)
        assert r.returncode == 0, (side, r.stderr[-2000:])
        with open(trace) as f:
            out.append([json.loads(line) for line in f])
    return out


def _compare_streams(a_calls, b_calls):
    assert [c["call"] for c in a_calls] == [c["call"] for c in b_calls]
    worst, n_solves = 0.0, 0
    for a, b in zip(a_calls, b_calls):
        assert set(a) == set(b), a["call"]
        n_solves += a["call"] == "solve"
        for k in a:
            if k == "call":
                continue
            # inert under LEVENBERG_MARQUARDT, and the one default cslam_options_init does not take from Ceres
            # (it starts at SUBSPACE_DOGLEG, what the reference's drivers set whenever they choose DOGLEG)
            if k == "dogleg_type" and a["trust_region_strategy"] == 0 and b["trust_region_strategy"] == 0:
                continue
            x, y = np.asarray(a[k], dtype=float), np.asarray(b[k], dtype=float)
            assert x.shape == y.shape, (a["call"], k, x.shape, y.shape)
            if x.size:
                worst = max(worst, float(np.abs(x - y).max() / max(1.0, np.abs(x).max())))
    return worst, n_solves


@pytest.mark.parametrize("flags,n_solves", [((), 1), (("--nolight",), 1), (("--dirlight",), 1), (("--multistage",), 3),
                                            (("--window", "4"), 7), (("--window", "3", "--multistage"), 24)])
def test_phong_driver_abi_stream_equals_reference_driver(tmp_path, flags, n_solves):
    """Config 3: tests/dataset_ba_phong.cpp (unmodified) against host/dataset_ba_phong_b200.cpp, call for call: the
    initial guess, every window, every stage (stereo only / lighting only with poses and positions constant / joint),
    point and directional light, bounds, SUBSPACE_DOGLEG options — and what each solve leaves in the parameter blocks.
    The two streams are bit-identical today; the test allows 1e-12."""
    def write_input(d):
        tr = syn.add_phong(_steady_track(10, seed=29), directional="--dirlight" in flags, shared_textures=True)
        path = str(d / "scene.csv")
        syn.write_phong_csv(tr, path)
        return [path]

    a, b = _abi_streams("dataset_ba_phong", 2, write_input, flags, tmp_path)
    worst, solves = _compare_streams(a, b)
    assert solves == n_solves
    assert worst <= 1e-12, worst
    print(f"dataset_ba_phong {' '.join(flags) or '(joint)'}: {len(a)} C ABI calls, {solves} solves, worst difference {worst:.3g}")


@pytest.mark.parametrize("window", [2, 0])
def test_vo_driver_abi_stream_equals_reference_driver(tmp_path, window):
    """Config 1 the same way.  The product driver first solves its first window once on a copy (an untimed warm-up for
    CUDA context creation): that leading group of calls is dropped before comparing."""
    def write_input(d):
        path = str(d / "track.csv")
        syn.write_track_csv(_steady_track(14, seed=17), path)
        return [path]

    a, b = _abi_streams("dataset_vo", 0, write_input, ("--window", str(window)), tmp_path)
    first_solve = [c["call"] for c in b].index("solve")
    b = b[first_solve + 1:]
    worst, solves = _compare_streams(a, b)
    assert solves == (13 if window == 2 else 1)
    assert worst <= 1e-12, worst


def test_sun_driver_abi_stream_equals_reference_driver(tmp_path):
    """Config 2 the same way, both passes: stereo blocks with per-observation stiffness, sun blocks with their Huber
    parameter and thresholds, the prior (reference pose and stiffness from the previous window's covariance), DOGLEG
    options, each solve's result and each covariance block.  The covariance is the oracle's C entry here on both sides
    (cslam_oracle_covariance_block); the facade's own route — the reference's functors, dense — is checked against it
    at the end.  The product driver's untimed warm-up (first window on a copy) is dropped."""
    paths_of = {}

    def write_input(d):
        tr, paths, _, _ = _sun_case(14, d)
        paths_of[d.name] = paths
        return paths

    flags = ("--window", "2", "--huber-param", "1.0")
    a, b = _abi_streams("dataset_vo_sun", 1, write_input, flags, tmp_path)
    b = b[[c["call"] for c in b].index("covariance_block") + 1:]
    worst, solves = _compare_streams(a, b)
    assert solves == 2 * 13 and sum(c["call"] == "covariance_block" for c in a) == 2 * 13
    assert worst <= 1e-12, worst
    # the same driver with the facade's independent covariance (reference functors + autodiff stand-in, dense Cholesky)
    rows, _ = _run_reference_driver(_lib("oracle", "dataset_vo_sun"), paths_of["ref"], 2, str(tmp_path / "trace.jsonl"),
                                    "dataset_vo_sun", ["--huber-param", "1.0"])
    own = [np.array(r["covariance"]) for r in rows if "covariance" in r]
    abi = [np.array(c["covariance"]) for c in a if c["call"] == "covariance_block"]
    assert len(own) == len(abi) == 26
    rel = max(float(np.abs(x - y).max() / np.abs(y).max()) for x, y in zip(own, abi))
    assert rel <= 1e-9, rel
    print(f"dataset_vo_sun: {len(a)} C ABI calls, worst stream difference {worst:.3g}; covariance routes agree to {rel:.3g}")
