"""TEST INFRASTRUCTURE — Python access to the checkers.

* `load_oracle()`  — oracle/_build/liboracle.so: the CPU restatement (prefix `cslam_oracle_`), same
  problem-building signatures as the product's C ABI, plus functor-level entry points;
* `load_ref()`     — oracle/_ref/libcslam_ref.so: the reference's own unmodified headers and its
  point_cloud_aligner.cpp compiled against the Eigen / Jet stand-ins (oracle/ref_capi.cpp): functor level and
  the RANSAC front end;
* `OracleProblem`, `build_problem`, `build_phong_problem`, `compute_initial_guess` — the host
  mirror of the product package driven by the oracle library instead, so a test builds the same
  problem twice and compares.
"""
import ctypes as C
import os
import subprocess

from ceres_slam_b200 import capi, initial_guess as _ig, synthetic as _syn
from ceres_slam_b200.capi import _dp, _u32p
from ceres_slam_b200.problem import BAProblem

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcslam_ref.so")
REFERENCE_INCLUDE = "/root/reference/include"

_ORACLE_ONLY = {
    "get_iteration_seconds": (C.c_int, [capi._h, _dp, C.c_int, capi._ip]),
    "poly_root_real_parts": (C.c_int, [_dp, C.c_int, _dp]),
    "dogleg_boundary_minimum": (C.c_int, [_dp, _dp, C.c_double, _dp]),
    "ransac_draws": (None, [C.c_uint32, C.c_uint32, C.c_int, _u32p, _u32p]),
    "covariance_block": (C.c_int, [capi._h, C.c_uint32, _dp]),   # same contract as cslam_covariance_block
    "kabsch": (None, [C.c_uint32, _dp, _dp, _dp]),
    "so3_exp": (None, [_dp, _dp]),
    "so3_log": (None, [_dp, _dp]),
    "se3_exp": (None, [_dp, _dp]),
    "se3_log": (None, [_dp, _dp]),
    "se3_mul": (None, [_dp, _dp, _dp]),
    "se3_inverse": (None, [_dp, _dp]),
    "se3_adjoint": (None, [_dp, _dp]),
    "se3_transform": (None, [_dp, _dp, C.c_int, _dp]),
    "se3_plus": (None, [_dp, _dp, _dp]),
    "se3_plus_jacobian": (None, [_dp, _dp]),
    "unit_plus": (None, [_dp, _dp, _dp]),
    "unit_plus_jacobian": (None, [_dp, _dp]),
    "camera_project": (None, [_dp, _dp, _dp]),
    "camera_triangulate": (None, [_dp, _dp, _dp]),
    "point_light_shade": (C.c_double, [_dp, _dp, _dp, _dp, C.c_double, _dp]),
    "intensity_block": (C.c_int, [_dp] * 6 + [C.c_double, C.c_double, C.c_int] + [_dp] * 7),
    "normal_block": (C.c_int, [_dp] * 7),
}

# oracle/ref_capi.cpp
_REF = {
    "describe": (C.c_char_p, []),
    "so3_exp": (None, [_dp, _dp]),
    "so3_log": (None, [_dp, _dp]),
    "so3_wedge": (None, [_dp, _dp]),
    "se3_exp": (None, [_dp, _dp]),
    "se3_log": (None, [_dp, _dp]),
    "se3_compose": (None, [_dp, _dp, _dp]),
    "se3_inverse": (None, [_dp, _dp]),
    "se3_transform_point": (None, [_dp, _dp, _dp, _dp]),
    "se3_adjoint": (None, [_dp, _dp]),
    "camera_project": (None, [_dp, _dp, _dp, _dp]),
    "camera_triangulate": (None, [_dp, _dp, _dp, _dp]),
    "se3_plus": (None, [_dp, _dp, _dp]),
    "se3_plus_jacobian": (None, [_dp, _dp]),
    "so3_plus": (None, [_dp, _dp, _dp]),
    "so3_plus_jacobian": (None, [_dp, _dp]),
    "unit_plus": (None, [_dp, _dp, _dp]),
    "unit_plus_jacobian": (None, [_dp, _dp]),
    "stereo_blocks": (C.c_int, [C.c_uint64, _dp, _u32p, _u32p, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "sun_blocks": (C.c_int, [C.c_uint32, _u32p, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]),
    "prior_block": (C.c_int, [_dp] * 5),
    "intensity_block": (C.c_int, [_dp] * 6 + [C.c_double, C.c_double, C.c_int] + [_dp] * 7),
    "normal_block": (C.c_int, [_dp] * 7),
    "point_light_shade": (C.c_double, [_dp, _dp, _dp, _dp, C.c_double, _dp]),
    # the reference's own src/ceres_slam/point_cloud_aligner.cpp (same signature as cslam_ransac_align)
    "ransac_align": capi._RANSAC_SIG,
    "kabsch": (None, [C.c_uint32, _dp, _dp, _dp]),
    "svd3": (None, [_dp, _dp, _dp, _dp]),
}

_cache = {}


def build_oracle():
    subprocess.check_call(["make", "-C", HERE])


def build_ref():
    """oracle/_ref from the reference's sources where they lie; a no-op (False) without /root/reference
    (the GPU box: the prebuilt library travels with the snapshot)."""
    if not os.path.isdir(REFERENCE_INCLUDE):
        return os.path.exists(REF_SO)
    subprocess.check_call(["make", "-C", HERE, "ref"])
    return True


def load_oracle():
    if "oracle" not in _cache:
        _cache["oracle"] = capi.Lib(ORACLE_SO, "cslam_oracle_", (capi.PROBLEM_API, _ORACLE_ONLY))
    return _cache["oracle"]


def have_ref():
    return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_INCLUDE)


def load_ref():
    if "ref" not in _cache:
        if not os.path.exists(REF_SO):
            build_ref()
        _cache["ref"] = capi.Lib(REF_SO, "cslam_ref_", (_REF,))
    return _cache["ref"]


class OracleProblem(BAProblem):
    """The product's host handle bound to the CPU restatement instead of the CUDA library."""

    @staticmethod
    def _library():
        return load_oracle()

    def iteration_seconds(self):
        """Steady-clock stamp of every row of iteration_log() (row 0 = the initial evaluation)."""
        import numpy as np
        n = C.c_int(0)
        self.lib.get_iteration_seconds(self._h, None, 0, C.byref(n))
        t = np.zeros(n.value)
        if n.value:
            self.lib.get_iteration_seconds(self._h, capi.dptr(t), n.value, C.byref(n))
        return t


def build_problem(track, **kw):
    return _syn.build_problem(track, problem_cls=OracleProblem, **kw)


def build_phong_problem(track, **kw):
    return _syn.build_phong_problem(track, problem_cls=OracleProblem, **kw)


def ransac_align(pairs0, pairs1, cam, **kw):
    return _ig.ransac_align(pairs0, pairs1, cam, entry=load_oracle().ransac_align, **kw)


def compute_initial_guess(track, poses, points, initialized, **kw):
    return _ig.compute_initial_guess(track, poses, points, initialized, entry=load_oracle().ransac_align, **kw)
