"""CPU-only checks of the product library: it loads without a GPU, exports every symbol
include/cslam_b200.h declares, refuses to compute without a device (no CPU fallback), and its
host-side structure analysis (free blocks, grouping, S pattern, landmark sharding) is sound."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from ceres_slam_b200 import capi
from ceres_slam_b200 import synthetic as syn
from ceres_slam_b200.problem import BAProblem, CslamError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ceres_slam_b200 import build
    build.build()
    return capi.load_product()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "cslam_b200.h")).read()
    names = sorted(set(re.findall(r"\b(cslam_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib.dll, n)]
    assert not missing, missing
    bound = set(capi.PRODUCT_SYMBOLS)
    assert bound <= set(names), bound - set(names)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    tr = syn.make_track(40, 5, 4, seed=1)
    p, _, _ = syn.build_problem(tr)
    with pytest.raises(CslamError, match="no CUDA device|cuda"):
        p.solve()
    with pytest.raises(CslamError):
        p.evaluate()


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: nothing in the
    product package or the public header includes, dlopens, imports or names the checker's libraries."""
    import re
    code_ref = re.compile(r"liboracle|libcslam_ref|cslam_oracle_|cslam_ref_|load_oracle|pybinding|ref_standin"
                          r"|^\s*(from|import)\s+oracle\b|[\"'/]oracle[/\"']")
    roots = [os.path.join(ROOT, "ceres_slam_b200"), os.path.join(ROOT, "include")]
    seen = 0
    for root in roots:
        for dirpath, dirs, files in os.walk(root):
            dirs[:] = [x for x in dirs if x not in ("build", "__pycache__", ".pytest_cache")]
            for f in files:
                if not f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".py")):
                    continue
                seen += 1
                for line in open(os.path.join(dirpath, f), errors="replace"):
                    assert not code_ref.search(line), (f, line)
                    if line.lstrip().startswith("#include") or "dlopen" in line or "CDLL" in line:
                        assert "oracle" not in line, (f, line)
    assert seen > 20
    # and at run time: importing the whole package maps neither checker library
    import subprocess, sys
    code = ("import ceres_slam_b200, ceres_slam_b200.problem, ceres_slam_b200.initial_guess, ceres_slam_b200.synthetic;"
            "from ceres_slam_b200 import capi; capi.load_product(); import sys;"
            "maps = open('/proc/self/maps').read();"
            "assert 'libcslam_b200' in maps; assert 'liboracle' not in maps and 'libcslam_ref' not in maps;"
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_structure_analysis_single_rank(lib):
    tr = syn.make_track(100, 15, 10, seed=3)
    p, _, _ = syn.build_problem(tr)
    info = p.analyze()
    n_obs = tr["obs_cam"].size
    seen = np.unique(tr["obs_pt"])
    assert info["n_observations"] == n_obs
    assert info["n_landmarks"] == seen.size
    assert info["landmark_id_sum"] == int(seen.sum())
    assert info["n_free_cams"] == 99          # first pose constant (dataset_vo.cpp:62)
    # co-visibility pattern from numpy: pairs of free cameras sharing a landmark
    cam, pt = tr["obs_cam"].astype(np.int64), tr["obs_pt"].astype(np.int64)
    pairs = set()
    order = np.argsort(pt, kind="stable")
    cam_s, pt_s = cam[order], pt[order]
    start = 0
    for end in list(np.flatnonzero(np.diff(pt_s)) + 1) + [pt_s.size]:
        cs = sorted(set(c for c in cam_s[start:end] if c != 0))
        for x in range(len(cs)):
            for y in range(x, len(cs)):
                pairs.add((cs[x], cs[y]))
        start = end
    for c in range(1, 100):
        pairs.add((c, c))
    assert info["nnz_blocks"] == len(pairs)
    assert info["n_grouped_landmarks"] > 0.5 * info["n_landmarks"]
    # forcing the generic path leaves no groups
    p1, _, _ = syn.build_problem(tr, schur_path=1)
    assert p1.analyze()["n_groups"] == 0


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from ceres_slam_b200 import synthetic as syn
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
tr = syn.make_track(120, 20, 8, seed=5)
p, _, _ = syn.build_problem(tr)
whole = p.analyze(1, 0)
mine = p.analyze(2, rank)
t = torch.tensor([mine["n_landmarks"], mine["n_observations"], mine["landmark_id_sum"], mine["nnz_blocks"],
                  mine["pattern_hash"] %% (2 ** 62), mine["n_free_cams"]], dtype=torch.int64)
both = [torch.zeros_like(t) for _ in range(2)]
dist.all_gather(both, t)
a, b = both
assert int(a[0] + b[0]) == whole["n_landmarks"], "landmark shards must partition the landmarks"
assert int(a[1] + b[1]) == whole["n_observations"], "observation shards must partition the observations"
assert int(a[2] + b[2]) == whole["landmark_id_sum"], "every landmark in exactly one shard"
assert int(a[3]) == int(b[3]) == whole["nnz_blocks"] and int(a[4]) == int(b[4]), "same global S pattern on every rank"
assert int(a[5]) == int(b[5]) == whole["n_free_cams"]
assert abs(int(a[1]) - int(b[1])) < 0.05 * whole["n_observations"], "shards balanced by observation count"
dist.barrier()
print("SHARD_OK", rank)
'''


def test_landmark_sharding_world_size_2_gloo(lib, tmp_path):
    """N > 1 host logic on CPU: two gloo ranks each analyse the whole problem for their rank."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": 29533})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [pr.communicate(timeout=240)[0] for pr in procs]
    for r, (pr, out) in enumerate(zip(procs, outs)):
        assert pr.returncode == 0 and f"SHARD_OK {r}" in out, out


def test_options_struct_layout_matches_the_header():
    """The ctypes mirror of cslam_options must have the header's layout: the LAST field carries a
    distinctive default, so any drift in the fields before it shows up there (both libraries)."""
    import ctypes as C
    from ceres_slam_b200 import capi
    from oracle import pybinding
    for lib in (capi.load_product(), pybinding.load_oracle()):
        opt = capi.Options()
        C.memset(C.byref(opt), 0xAB, C.sizeof(opt))
        lib.options_init(C.byref(opt))
        assert opt.line_search_sufficient_function_decrease == 1e-4
        assert opt.max_num_iterations == 1000 and opt.dogleg_type == 1 and opt.trust_region_strategy == 0
        assert opt.initial_trust_region_radius == 1e4


def test_host_mirror_and_drivers_compile_and_refuse_without_a_gpu(lib, tmp_path):
    """The C++ host mirror (host/cslam_problem.hpp), the restated drivers and the ba_all batch runner build against
    the C ABI with g++ alone; without a CUDA device a driver must fail loudly (no CPU fallback), not produce output."""
    import os
    import subprocess
    from ceres_slam_b200 import build as b
    from ceres_slam_b200 import synthetic as syn
    exes = {n: b.build_host_driver(n) for n in b.HOST_DRIVERS}
    assert set(exes) == {"dataset_vo_b200", "dataset_vo_sun_b200", "dataset_ba_phong_b200", "ba_all_b200"}
    for exe in exes.values():
        assert os.access(exe, os.X_OK)
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert out.returncode != 0 and "usage" in out.stderr      # no arguments: usage, like the reference's drivers
    import torch
    if not torch.cuda.is_available():
        tr = syn.make_track(12, 8, 5, seed=3)
        csv = os.path.join(tmp_path, "track.csv")
        syn.write_track_csv(tr, csv)
        out = subprocess.run([exes["dataset_vo_b200"], csv, "--window", "2", "--init", "constant"], capture_output=True, text=True,
                             timeout=120, cwd=tmp_path)
        assert out.returncode != 0 and not os.path.exists(os.path.join(tmp_path, "track_poses.csv")), out.stdout + out.stderr
