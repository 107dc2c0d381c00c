"""Known-answer tests against tests/golden/functors_mp60.json — residuals and tangent-space
Jacobians of every functor on the path, computed independently in 60-digit arithmetic by
tests/golden/make_golden.py (no Jets, no closed forms).

CPU (`not gpu`): the Jet oracle and the host build of the device closed forms against the golden
vectors.  GPU: the CUDA path through the C ABI (`cslam_evaluate`) against the same vectors.
Tolerance: BASELINE.json's 1e-10 relative, per block (relative to the block's largest entry)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from ceres_slam_b200 import capi
from ceres_slam_b200.problem import BAProblem
from oracle import pybinding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = capi.dptr
RTOL = 1e-10


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "functors_mp60.json")) as f:
        return json.load(f)


def A(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def close(a, b, rtol=RTOL, floor=0.0):
    a, b = np.asarray(a, dtype=float).ravel(), np.asarray(b, dtype=float).ravel()
    scale = max(float(np.abs(b).max()), floor, 1e-300)
    return float(np.abs(a - b).max()) <= rtol * scale


def one_pose_problem(backend, g, pose, point=None):
    p = (orc.OracleProblem if backend == "oracle" else BAProblem)()
    p.set_camera(*g["camera"])
    poses = p.set_poses(A(pose).reshape(1, 12).copy(), np.zeros(1, dtype=np.uint8))
    pts = p.set_points(A(point if point is not None else [0, 0, 5]).reshape(1, 3).copy())
    return p, poses, pts


def eval_stereo(backend, g):
    out = []
    for c in g["stereo"]:
        p, _, _ = one_pose_problem(backend, g, c["pose"], c["point"])
        p.add_stereo(np.zeros(1, np.uint32), np.zeros(1, np.uint32), A(c["uvd"]), A(c["W"]))
        out.append(p.evaluate())
        p.close()
    return out


def eval_sun(backend, g):
    out = []
    for c in g["sun"]:
        p, _, _ = one_pose_problem(backend, g, c["pose"])
        p.add_sun(np.zeros(1, np.uint32), A(c["obs_c"]), A(c["ref_g"]), A(c["W"]), c["az_thresh"], c["zen_thresh"])
        out.append(p.evaluate())
        p.close()
    return out


def eval_prior(backend, g):
    out = []
    for c in g["prior"]:
        p, _, _ = one_pose_problem(backend, g, c["pose"])
        p.add_pose_prior(0, A(c["Tref"]), A(c["W"]))
        out.append(p.evaluate())
        p.close()
    return out


def check_core(g, ev_st, ev_sun, ev_pr):
    for c, e in zip(g["stereo"], ev_st):
        assert close(e["r_stereo"], c["r"]) and close(e["Jpose_stereo"], c["J_pose"])
        assert close(e["Jpoint_stereo"], c["J_point"])
        assert abs(e["cost"] - 0.5 * np.dot(c["r"], c["r"])) <= 1e-12 * max(e["cost"], 1e-300)
    for c, e in zip(g["sun"], ev_sun):
        assert close(e["r_sun"], c["r"]) and close(e["J_sun"], c["J_pose"])
    for c, e in zip(g["prior"], ev_pr):
        # the residual of a prior evaluated at its own reference is rounding noise of R R^T - I
        # scaled by W (up to 1e6): compare against the scale |W| * eps-level pose entries
        floor = float(np.abs(c["W"]).max()) * 1e-5
        assert close(e["r_prior"], c["r"], floor=floor) and close(e["J_prior"], c["J_pose"])


# ---- CPU: oracle and closed forms against the golden vectors ----------------------------------
def test_oracle_core_functors(golden, oracle):
    check_core(golden, eval_stereo("oracle", golden), eval_sun("oracle", golden), eval_prior("oracle", golden))


def test_oracle_normal_and_intensity(golden, oracle):
    for c in golden["normal"]:
        r, Jc, Jn = np.zeros(3), np.zeros(18), np.zeros(9)
        assert oracle.normal_block(d(A(c["pose"])), d(A(c["normal"])), d(A(c["obs"])), d(A(c["W"])), d(r), d(Jc), d(Jn)) == 0
        assert close(r, c["r"]) and close(Jc, c["J_pose"]) and close(Jn, c["J_normal"])
    for c in golden["intensity"]:
        r, Jc, Jp, Jn, Jk, Jt, Jl = (np.zeros(n) for n in (1, 6, 3, 3, 3, 1, 3))
        assert oracle.intensity_block(d(A(c["pose"])), d(A(c["point"])), d(A(c["normal"])), d(A(c["phong"])),
                                      d(A(c["texture"])), d(A(c["light"])), c["colour"], c["stiffness"],
                                      c["directional"], d(r), d(Jc), d(Jp), d(Jn), d(Jk), d(Jt), d(Jl)) == 0
        assert close(r, c["r"]), c["kind"]
        # one common scale for all Jacobian blocks of the residual (it is one row of J)
        full = np.concatenate([c[k] for k in ("J_pose", "J_point", "J_normal", "J_phong", "J_tex", "J_light")])
        got = np.concatenate([Jc, Jp, Jn, Jk, Jt, Jl])
        assert close(got, full), c["kind"]


def test_oracle_plus_operations(golden, oracle):
    for c in golden["se3_plus"]:
        out = np.zeros(12)
        oracle.se3_plus(d(A(c["pose"])), d(A(c["delta"])), d(out))
        assert close(out, c["out"], rtol=1e-14)
    for c in golden["unit_plus"]:
        out, J = np.zeros(3), np.zeros(9)
        oracle.unit_plus(d(A(c["x"])), d(A(c["delta"])), d(out))
        oracle.unit_plus_jacobian(d(A(c["x"])), d(J))
        assert close(out, c["out"], rtol=1e-14) and close(J, c["J_plus"])


def test_closed_forms(golden, cf):
    intr = A(golden["camera"])
    for c in golden["stereo"]:
        r, Jc, Jp = np.zeros(3), np.zeros(18), np.zeros(9)
        cf.cf_stereo_block(d(intr), d(A(c["pose"])), d(A(c["point"])), d(A(c["uvd"])), d(A(c["W"])), d(r), d(Jc), d(Jp))
        assert close(r, c["r"]) and close(Jc, c["J_pose"]) and close(Jp, c["J_point"])
    for c in golden["sun"]:
        obs, ref = A(c["obs_c"]), A(c["ref_g"])
        obs, ref = obs / np.linalg.norm(obs), ref / np.linalg.norm(ref)  # cslam_add_sun normalises
        r, J = np.zeros(2), np.zeros(12)
        cf.cf_sun_block(d(A(c["pose"])), d(obs), d(ref), d(A(c["W"])), c["az_thresh"], c["zen_thresh"], d(r), d(J))
        assert close(r, c["r"]) and close(J, c["J_pose"])
    for c in golden["prior"]:
        r, J = np.zeros(6), np.zeros(36)
        cf.cf_prior_block(d(A(c["pose"])), d(A(c["Tref"])), d(A(c["W"])), d(r), d(J))
        assert close(r, c["r"], floor=float(np.abs(c["W"]).max()) * 1e-5) and close(J, c["J_pose"])
    for c in golden["se3_plus"]:
        out = np.zeros(12)
        cf.cf_se3_plus(d(A(c["pose"])), d(A(c["delta"])), d(out))
        assert close(out, c["out"], rtol=1e-14)
    for c in golden["normal"]:
        r, Jc, Jn = np.zeros(3), np.zeros(18), np.zeros(9)
        cf.cf_normal_block(d(A(c["pose"])), d(A(c["normal"])), d(A(c["obs"])), d(A(c["W"])), d(r), d(Jc), d(Jn))
        assert close(r, c["r"]) and close(Jc, c["J_pose"]) and close(Jn, c["J_normal"])
    for c in golden["intensity"]:
        r, Jc, Jp, Jn, Jk, Jt, Jl = (np.zeros(n) for n in (1, 6, 3, 3, 3, 1, 3))
        cf.cf_intensity_block(d(A(c["pose"])), d(A(c["point"])), d(A(c["normal"])), d(A(c["phong"])),
                              d(A(c["texture"])), d(A(c["light"])), c["colour"], c["stiffness"],
                              c["directional"], d(r), d(Jc), d(Jp), d(Jn), d(Jk), d(Jt), d(Jl))
        full = np.concatenate([c[k] for k in ("J_pose", "J_point", "J_normal", "J_phong", "J_tex", "J_light")])
        assert close(r, c["r"]), c["kind"]
        assert close(np.concatenate([Jc, Jp, Jn, Jk, Jt, Jl]), full), c["kind"]
    for c in golden["unit_plus"]:
        out = np.zeros(3)
        cf.cf_unit_plus(d(A(c["x"])), d(A(c["delta"])), d(out))
        assert close(out, c["out"], rtol=1e-14)


# ---- GPU: the CUDA path through the C ABI against the golden vectors --------------------------
@pytest.mark.gpu
def test_cuda_core_functors(golden, product):
    check_core(golden, eval_stereo("b200", golden), eval_sun("b200", golden), eval_prior("b200", golden))
