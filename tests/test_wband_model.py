"""CPU check of the layout the wide-band solver (K3e, ceres_slam_b200/csrc/kernels_wband.cu) works in: the numpy model
`scripts/wband_model.py` uses the kernels' storage formulas (band storage with the previous column's tail, border rows
[left separator | rhs | right separator], inactive right-separator rows before r_start, the gathered separator system)
and loop structure, and must solve random SPD block-banded systems to rounding."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    spec = importlib.util.spec_from_file_location("wband_model", os.path.join(ROOT, "scripts", "wband_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("n_free,w,C", [(40, 3, 1), (60, 3, 3), (75, 5, 4), (90, 13, 2), (64, 2, 5)])
def test_chunked_bordered_band_solve(model, n_free, w, C):
    M, blocks, rhs = model.random_system(n_free, w, 11 + n_free)
    y = model.solve(blocks, rhs, n_free, w, C)
    ref = np.linalg.solve(M, rhs)
    assert np.abs(y - ref).max() <= 1e-10 * np.abs(ref).max()
