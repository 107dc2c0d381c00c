import os
import subprocess
import sys

import pytest

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
    from ceres_slam_b200 import capi
    return capi.load_oracle()


@pytest.fixture(scope="session")
def product():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ceres_slam_b200 import capi
    return capi.load_product()
