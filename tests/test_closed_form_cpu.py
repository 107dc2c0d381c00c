"""The closed-form residual/Jacobian formulas the CUDA kernels use (csrc/closed_form.h, compiled
for the host by tests/closed_form_host.cpp) against the Jet-autodiff oracle, on CPU.
Tolerance: 1e-10 relative to the block's largest entry — BASELINE.json's "match Ceres autodiff
within 1e-10 relative" (autodiff leaves 1e-16-level residue where the closed form has exact
zeros, so the comparison is per block, not per element)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ceres_slam_b200 import capi
from ceres_slam_b200 import synthetic as syn
from oracle import pybinding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = capi.dptr
RTOL = 1e-10


@pytest.fixture(scope="module")
def cf():
    so = os.path.join(ROOT, "tests", "_build", "libclosedform_host.so")
    srcs = [os.path.join(ROOT, "tests", "closed_form_host.cpp"),
            os.path.join(ROOT, "ceres_slam_b200", "csrc", "closed_form.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", so, srcs[0]])
    lib = C.CDLL(so)
    for name in ("cf_stereo_block", "cf_sun_block", "cf_prior_block", "cf_se3_plus", "cf_so3_log"):
        getattr(lib, name).restype = None
    lib.cf_sun_block.argtypes = [capi._dp] * 4 + [C.c_double, C.c_double] + [capi._dp] * 2
    return lib


def blockwise_close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= rtol * scale


def test_stereo_closed_form_vs_autodiff(cf):
    tr = syn.make_track(60, 4, 6, seed=1, per_obs_W=True)
    p, poses, points = orc.build_problem(tr, hold_first=False)
    ev = p.evaluate()
    intr = np.array([tr["cam"][k] for k in ("fu", "fv", "cu", "cv", "b")])
    n = tr["obs_cam"].size
    assert n > 500
    for i in range(n):
        r, Jc, Jp = np.zeros(3), np.zeros(18), np.zeros(9)
        cf.cf_stereo_block(d(intr), d(poses[tr["obs_cam"][i]].copy()), d(points[tr["obs_pt"][i]].copy()),
                           d(tr["uvd"][i].copy()), d(tr["W"][i].copy()), d(r), d(Jc), d(Jp))
        assert blockwise_close(r, ev["r_stereo"][i])
        assert blockwise_close(Jc.reshape(3, 6), ev["Jpose_stereo"][i])
        assert blockwise_close(Jp.reshape(3, 3), ev["Jpoint_stereo"][i])


def _rand_pose(rng, oracle, scale=1.0):
    T = np.zeros(12)
    xi = rng.normal(size=6) * scale
    oracle.se3_exp(d(xi), d(T))
    return T


def test_sun_closed_form_vs_autodiff(cf, oracle):
    rng = np.random.default_rng(5)
    from ceres_slam_b200.problem import BAProblem
    for trial in range(200):
        pose = _rand_pose(rng, oracle)
        obs = rng.normal(size=3)
        ref = rng.normal(size=3)
        A = rng.normal(size=(2, 2))
        W2 = (A @ A.T + np.eye(2)).reshape(4)
        # a third of the trials exercise the hard thresholds (sun_sensor_error.hpp:87-93)
        az_t, zen_t = (1000.0, 1000.0) if trial % 3 else (0.8, 0.5)
        p = orc.OracleProblem()
        p.set_camera(1, 1, 0, 0, 1)
        p.set_poses(pose[None, :].copy())
        p.set_points(np.zeros((1, 3)))
        p.add_sun([0], obs[None], ref[None], W2[None], az_t, zen_t, 0.0)
        ev = p.evaluate()
        r, J = np.zeros(2), np.zeros(12)
        cf.cf_sun_block(d(pose), d(obs / np.linalg.norm(obs)), d(ref / np.linalg.norm(ref)), d(W2),
                        az_t, zen_t, d(r), d(J))
        assert blockwise_close(r, ev["r_sun"][0])
        assert blockwise_close(J.reshape(2, 6), ev["J_sun"][0])


@pytest.mark.parametrize("offset_scale", [0.0, 1e-18, 1e-12, 1e-9, 1e-6, 1e-3, 0.1, 1.0])
def test_prior_closed_form_vs_autodiff(cf, oracle, offset_scale):
    """pose_error.hpp:22-55.  offset 0 is how dataset_vo_sun.cpp:120-124 always starts (T_ref is a
    copy of the pose): first-order log branch, J = -W exactly."""
    rng = np.random.default_rng(7)
    from ceres_slam_b200.problem import BAProblem
    for trial in range(25):
        Tref = _rand_pose(rng, oracle)
        eps = rng.normal(size=6) * offset_scale
        pose = np.zeros(12)
        oracle.se3_plus(d(Tref), d(eps), d(pose))
        A = rng.normal(size=(6, 6))
        W6 = (A @ A.T + 6 * np.eye(6)) * (1e6 if trial % 2 else 1.0)   # Sigma0 = 1e-12 I -> W = 1e6 I
        p = orc.OracleProblem()
        p.set_camera(1, 1, 0, 0, 1)
        p.set_poses(pose[None, :].copy())
        p.set_points(np.zeros((1, 3)))
        p.add_pose_prior(0, Tref, W6)
        ev = p.evaluate()
        r, J = np.zeros(6), np.zeros(36)
        cf.cf_prior_block(d(pose), d(Tref), d(W6.reshape(36).copy()), d(r), d(J))
        # residual: the rotation part is 1e-16-level noise when pose == Tref
        assert np.abs(r - ev["r_prior"][0]).max() <= RTOL * max(np.abs(ev["r_prior"][0]).max(), np.abs(W6).max() * 1e-15)
        assert blockwise_close(J.reshape(6, 6), ev["J_prior"][0])
        if offset_scale == 0.0:
            assert blockwise_close(J.reshape(6, 6), -W6, 1e-14)


def test_se3_plus_matches_oracle(cf, oracle):
    rng = np.random.default_rng(11)
    for scale in (0.0, 1e-17, 1e-8, 1e-2, 1.0, 3.0):
        for _ in range(20):
            T = _rand_pose(rng, oracle)
            eps = rng.normal(size=6) * scale
            a, b = np.zeros(12), np.zeros(12)
            oracle.se3_plus(d(T), d(eps), d(a))
            cf.cf_se3_plus(d(T), d(eps), d(b))
            assert np.allclose(a, b, rtol=0, atol=1e-15 * max(1.0, np.abs(a).max()))
