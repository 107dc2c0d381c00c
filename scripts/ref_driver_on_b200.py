"""One-off GPU check without pytest / torch: the reference's own tests/dataset_vo.cpp and dataset_vo_sun.cpp (unmodified,
oracle/_ref/libref_*_b200.so: Ceres-API facade -> C ABI -> CUDA library) against the Python driver mirror on the oracle.
The same comparison as the GPU legs of tests/test_ref_driver.py."""
import os
import pathlib
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_ref_driver as t  # noqa: E402
from ceres_slam_b200 import synthetic as syn  # noqa: E402
from test_gpu_parity import _steady_track  # noqa: E402

tmp = pathlib.Path(tempfile.mkdtemp())
t0 = time.time()
for window in (2, 0):
    d = tmp / f"vo{window}"
    d.mkdir()
    tr = _steady_track(30, seed=17)
    csv = str(d / "track.csv")
    syn.write_track_csv(tr, csv)
    rows, _ = t._run_reference_driver(t._lib("b200"), csv, window, str(d / "trace.jsonl"), timeout=60)
    err = t._compare_with_mirror(tr, rows, window, 1e-6, max(1, len(rows) // 10))
    print(f"dataset_vo.cpp on the B200 back end, window {window}: {len(rows)} solves, worst pose difference vs mirror {err:.3g}"
          f" ({time.time() - t0:.1f} s)", flush=True)
d = tmp / "sun"
d.mkdir()
tr, paths, cov, sun = t._sun_case(25, d)
rows, _ = t._run_reference_driver(t._lib("b200", "dataset_vo_sun"), paths, 2, str(d / "trace.jsonl"), "dataset_vo_sun",
                                  ["--huber-param", "1.0"], timeout=60)
err = t._compare_sun_with_mirror(tr, cov, sun, rows, 1e-6, 1e-5, 5)
print(f"dataset_vo_sun.cpp on the B200 back end: {len(rows)} trace rows, worst pose difference vs mirror {err:.3g} ({time.time() - t0:.1f} s)")
