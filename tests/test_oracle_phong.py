"""CPU checks of the lighting-solve oracle (oracle/phong_problem.hpp) — no GPU needed.

The reference's tests assert nothing on this path, so the restatement is
pinned to what can be pinned: its stereo part against the independent stereo oracle
(oracle/problem.hpp), recovery of the generating parameters on noise-free data, and the box."""
import numpy as np
import pytest

from ceres_slam_b200 import synthetic as syn
from oracle import pybinding as orc
from ceres_slam_b200.problem import CslamError

FIXED = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)


def test_phong_oracle_reduces_to_stereo_oracle():
    """With zero intensity / normal stiffness the lighting rows vanish: the joint solve must walk
    the stereo-only LM trajectory of the independent stereo oracle (same cost, radius, poses,
    points) and leave normals, materials, textures and the light where they were."""
    tr = syn.add_phong(syn.make_track(20, 10, 5, seed=3), shared_textures=True)
    tr["int_stiffness"] = 0.0
    tr["W_normal"] = np.zeros(9)
    kw = dict(FIXED, max_num_iterations=5, num_threads=4)
    pj, sj = orc.build_phong_problem(tr, **kw)
    ps, poses_s, points_s = orc.build_problem(tr, **kw)
    before = {k: sj[k].copy() for k in ("normals", "phong", "textures", "light")}
    rj, rs = pj.solve(), ps.solve()
    lj, ls = pj.iteration_log(), ps.iteration_log()
    assert lj.shape == ls.shape and rj.num_successful_steps == rs.num_successful_steps
    assert np.allclose(lj[:, 1], ls[:, 1], rtol=1e-9), "cost trajectory"
    assert np.allclose(lj[:, 6], ls[:, 6], rtol=1e-9), "radius trajectory"
    assert np.allclose(sj["poses"], poses_s, rtol=0, atol=1e-9)
    assert np.allclose(sj["points"], points_s, rtol=0, atol=1e-9)
    for k, v in before.items():
        assert np.allclose(sj[k], v, rtol=0, atol=1e-12), k


@pytest.mark.parametrize("directional", [False, True])
def test_phong_oracle_recovers_noise_free_scene(directional):
    """Noise-free observations rendered by the generator's own numpy Phong model: the joint solve
    drives the cost to ~0 and recovers light, textures and materials."""
    base = syn.make_track(14, 25, 6, seed=9)
    # noise-free stereo observations: re-project the ground truth
    k, j = base["obs_cam"].astype(np.int64), base["obs_pt"].astype(np.int64)
    pc = np.einsum("nij,nj->ni", syn.pose_R(base["poses_gt"])[k], base["points_gt"][j]) + syn.pose_t(base["poses_gt"])[k]
    base["uvd"] = np.ascontiguousarray(syn.project(base["cam"], pc))
    tr = syn.add_phong(base, directional=directional, int_var=1e-4, normal_var=1e-4, shared_textures=True)
    # the same scene rendered with (numerically) no intensity / normal noise, same stiffness as above
    clean = syn.add_phong(dict(base), directional=directional, int_var=1e-30, normal_var=1e-30, shared_textures=True)
    for key in ("intensity", "normal_obs"):
        tr[key] = clean[key]
    p, st = orc.build_phong_problem(tr, bounds=True, max_num_iterations=60, num_threads=8)
    s = p.solve()
    assert s.final_cost < 1e-6 * s.initial_cost
    assert np.abs(st["light"] - tr["light_gt"]).max() < 1e-3
    assert np.abs(st["textures"] - tr["tex_shared_gt"]).max() < 5e-3
    assert np.median(np.abs(st["phong"][:, 1] - tr["phong_gt"][:, 1])) < 1e-2  # ks (weakly observable for some materials)
    assert np.abs(st["normals"] - tr["normals_gt"]).max() < 5e-3
    assert np.allclose(np.linalg.norm(st["normals"], axis=1), 1.0, atol=1e-12)  # UnitVectorPerturbation


def test_phong_oracle_box_and_reference_start():
    """The reference's starting point (materials (0, 0, 1), dataset_problem_phong.cpp:262-279) sits
    on the box of dataset_ba_phong.cpp:143-181; every iterate stays inside it."""
    tr = syn.add_phong(syn.make_track(20, 12, 6, seed=5), shared_textures=True)
    tr["phong"] = np.tile(np.array([0.0, 0.0, 1.0]), (tr["phong"].shape[0], 1))
    p, st = orc.build_phong_problem(tr, bounds=True, max_num_iterations=10, num_threads=8, **FIXED)
    s = p.solve()
    assert s.final_cost < 0.1 * s.initial_cost
    assert np.all(st["phong"][:, :2] >= 0) and np.all(st["phong"][:, :2] <= 1) and np.all(st["phong"][:, 2] >= 1)
    assert np.all(st["textures"] >= 0) and np.all(st["textures"] <= 1)
    assert np.all(st["phong"][:, 0] == 0.0)   # ka has no effect (ambient disabled, phong.hpp:31-33)


def test_phong_oracle_refuses_unpaired_blocks():
    tr = syn.add_phong(syn.make_track(10, 8, 4, seed=2), shared_textures=True)
    p, _ = orc.build_phong_problem(tr)
    n = tr["obs_cam"].size - 3
    p.add_phong(tr["obs_cam"][:n], tr["obs_pt"][:n], tr["intensity"][:n], tr["int_stiffness"], tr["normal_obs"][:n],
                tr["W_normal"])
    with pytest.raises(CslamError):
        p.solve()


def test_phong_oracle_stage2_holds_poses_and_positions():
    """Stage 2 of dataset_ba_phong --multistage (dataset_ba_phong.cpp:207-246): every pose and vertex
    position constant, only the lighting parameters move; the stereo blocks are dropped and their
    cost is carried as Ceres' fixed_cost."""
    tr = syn.add_phong(syn.make_track(14, 20, 6, seed=9), shared_textures=True)
    tr["constant"] = np.ones(tr["n_poses"], dtype=np.uint8)
    p, st = orc.build_phong_problem(tr, bounds=True, max_num_iterations=20, num_threads=4)
    p.set_points_constant(True)
    before = {k: st[k].copy() for k in st}
    pj, _ = orc.build_phong_problem(tr, bounds=True, max_num_iterations=0, num_threads=4)
    s = p.solve()
    assert np.array_equal(st["poses"], before["poses"]) and np.array_equal(st["points"], before["points"])
    for k in ("normals", "phong", "textures", "light"):
        assert np.abs(st[k] - before[k]).max() > 1e-6, k
    assert s.final_cost < s.initial_cost
    # the reported cost includes the dropped stereo blocks: same initial cost as the joint problem
    assert abs(s.initial_cost - pj.solve().initial_cost) < 1e-9 * s.initial_cost


@pytest.mark.parametrize("dogleg_type", [0, 1])
def test_phong_oracle_dogleg(dogleg_type):
    """dataset_ba_phong.cpp:88-89 sets DOGLEG / SUBSPACE_DOGLEG.  (a) With zero lighting stiffness the
    lighting oracle's DOGLEG trajectory equals the stereo oracle's (an independent implementation of the
    same strategy); (b) on a bounded lighting problem LM and DOGLEG reach the same basin."""
    tr = syn.add_phong(syn.make_track(25, 6, 5, seed=3), shared_textures=True)
    tr["int_stiffness"] = 0.0
    tr["W_normal"] = np.zeros(9)
    kw = dict(max_num_iterations=6, num_threads=4, trust_region_strategy=1, dogleg_type=dogleg_type,
              initial_trust_region_radius=2.0, **FIXED)
    pj, sj = orc.build_phong_problem(tr, **kw)
    summ_j = pj.solve()
    ps, poses_s, points_s = orc.build_problem(tr, **kw)
    summ_s = ps.solve()
    lj, ls = pj.iteration_log(), ps.iteration_log()
    assert lj.shape == ls.shape
    assert np.allclose(lj[:, 1], ls[:, 1], rtol=1e-9), "cost trajectory"
    assert np.allclose(lj[:, 6], ls[:, 6], rtol=1e-7), "radius trajectory"
    assert np.array_equal(lj[:, 9], ls[:, 9])
    tr = syn.add_phong(syn.make_track(25, 6, 5, seed=4))
    out = {}
    for strat in (0, 1):
        p, st = orc.build_phong_problem(tr, bounds=True, max_num_iterations=80, num_threads=4, trust_region_strategy=strat,
                                        dogleg_type=dogleg_type)
        s = p.solve()
        assert s.termination_type in (0, 1)
        out[strat] = s.final_cost
    # with a material parameter on its bound the projected Gauss-Newton steps of DOGLEG crawl along the face
    # (every step accepted, tiny): both strategies get within 2 % of each other, far below the initial cost
    assert abs(out[0] - out[1]) < 2e-2 * out[0] and out[1] < 0.05 * s.initial_cost
