"""The reference's OWN driver, unmodified, on this repo's back end.

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


def _lib(kind):
    so = os.path.join(REF_DIR, f"libref_dataset_vo_{kind}.so")
    if not os.path.exists(so) and os.path.isdir(orc.REFERENCE_INCLUDE):
        orc.build_ref()
    if not os.path.exists(so):
        pytest.skip(f"{so} not built and /root/reference absent")
    return so


_RUNNER = r"""
import ctypes as C, sys
lib = C.CDLL(sys.argv[1])
lib.cslam_ref_dataset_vo_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
args = [b"dataset_vo"] + [a.encode() for a in sys.argv[2:]]
sys.exit(lib.cslam_ref_dataset_vo_main(len(args), (C.c_char_p * len(args))(*args)))
"""


def _run_reference_driver(so, csv, window, trace):
    """In a child process: the driver narrates on stdout / stderr, and a crash must not take the test session down."""
    env = dict(os.environ, CSLAM_FACADE_TRACE=trace)
    if os.path.exists(trace):
        os.remove(trace)
    r = subprocess.run([sys.executable, "-c", _RUNNER, so, csv, "--window", str(window)], env=env, capture_output=True,
                       text=True, timeout=600)
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
@pytest.mark.xfail(strict=False, reason="added after this round's GPU budget was spent: the CPU leg (same facade, "
                   "oracle instead of the CUDA library) is verified, this leg has not run on a B200 yet")
@pytest.mark.parametrize("window", [2, 0])
def test_reference_driver_on_b200_matches_driver_mirror(product, tmp_path, window):
    """The reference's tests/dataset_vo.cpp, unmodified, with ceres::Solve answered by the CUDA library."""
    so = _lib("b200")
    tr = _steady_track(30, seed=17)
    csv = str(tmp_path / "track.csv")
    syn.write_track_csv(tr, csv)
    rows, _ = _run_reference_driver(so, csv, window, str(tmp_path / "trace.jsonl"))
    _compare_with_mirror(tr, rows, window, 1e-6, max(1, len(rows) // 10))
