"""CPU checks of the DOGLEG restatement (oracle/problem.hpp): the quartic of the
subspace boundary problem against numpy and brute force, and the strategy against Levenberg-Marquardt
on the same problem (same minimum, different paths)."""
import numpy as np
import pytest

from ceres_slam_b200 import capi, synthetic as syn
from oracle import pybinding as orc


def test_polynomial_roots_match_numpy(oracle):
    rng = np.random.default_rng(1)
    for trial in range(50):
        c = rng.normal(0, 1, 5) * 10.0 ** rng.integers(-3, 4, 5)
        c[0] = abs(c[0]) + 1e-3
        out = np.zeros(4)
        assert oracle.poly_root_real_parts(capi.dptr(c), 5, capi.dptr(out)) == 4
        ref = np.sort(np.roots(c).real)
        assert np.allclose(np.sort(out), ref, rtol=1e-8, atol=1e-8 * np.abs(ref).max())


def test_subspace_boundary_minimum_is_the_constrained_minimum(oracle):
    """min 1/2 x^T B x + g^T x on |x| = r, against a dense scan of the circle."""
    rng = np.random.default_rng(2)
    th = np.linspace(0, 2 * np.pi, 400001)
    for trial in range(30):
        M = rng.normal(0, 1, (2, 2))
        B = M @ M.T + (0.0 if trial % 3 else 1e-9) * np.eye(2)
        g = rng.normal(0, 1, 2)
        r = 10.0 ** rng.uniform(-2, 1)
        x = np.zeros(2)
        assert oracle.dogleg_boundary_minimum(capi.dptr(np.ascontiguousarray(B.reshape(4))), capi.dptr(g), r, capi.dptr(x)) == 1
        p = r * x / np.linalg.norm(x)
        f = 0.5 * p @ B @ p + g @ p
        X = r * np.stack([np.cos(th), np.sin(th)])
        F = 0.5 * np.einsum("in,ij,jn->n", X, B, X) + g @ X
        assert f <= F.min() + 1e-9 * max(1.0, abs(F.min()))
        assert abs(np.linalg.norm(x) - r) < 1e-6 * r          # a true boundary root


@pytest.mark.parametrize("dogleg_type", [0, 1])
def test_dogleg_and_lm_reach_the_same_minimum(dogleg_type):
    tr = syn.add_sun(syn.make_track(40, 12, 6, seed=21))
    out = {}
    for strat in (0, 1):
        p, poses, points = orc.build_problem(tr, sun=True, max_num_iterations=50, num_threads=4,
                                             trust_region_strategy=strat, dogleg_type=dogleg_type,
                                             initial_trust_region_radius=2.0)
        s = p.solve()
        assert s.termination_type == 0
        out[strat] = (s.final_cost, poses, points)
    assert abs(out[0][0] - out[1][0]) < 1e-6 * out[0][0]
    assert np.abs(out[0][1] - out[1][1]).max() < 5e-3   # both stop on the function tolerance, not at the exact minimum
