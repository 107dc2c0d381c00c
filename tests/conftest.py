import os
import subprocess
import sys

import pytest

import ctypes as _C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is the checker: make sure its library exists (g++ only, seconds)."""
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in
            ("oracle_capi.cpp", "jet.hpp", "geometry.hpp", "functors.hpp", "problem.hpp", "phong_problem.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle import pybinding
    return pybinding.load_oracle()


@pytest.fixture(scope="session")
def product():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ceres_slam_b200 import capi
    return capi.load_product()


@pytest.fixture(scope="session")
def cf():
    """Host build of the device closed forms (csrc/closed_form.h) for CPU-side checks."""
    from ceres_slam_b200 import capi as _capi
    so = os.path.join(ROOT, "tests", "_build", "libclosedform_host.so")
    srcs = [os.path.join(ROOT, "tests", "closed_form_host.cpp"),
            os.path.join(ROOT, "ceres_slam_b200", "csrc", "closed_form.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", so, srcs[0]])
    lib = _C.CDLL(so)
    for name in ("cf_stereo_block", "cf_sun_block", "cf_prior_block", "cf_se3_plus", "cf_so3_log",
                 "cf_normal_block", "cf_intensity_block", "cf_unit_plus"):
        getattr(lib, name).restype = None
    lib.cf_sun_block.argtypes = [_capi._dp] * 4 + [_C.c_double, _C.c_double] + [_capi._dp] * 2
    lib.cf_intensity_block.argtypes = [_capi._dp] * 6 + [_C.c_double, _C.c_double, _C.c_int] + [_capi._dp] * 7
    return lib
