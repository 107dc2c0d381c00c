"""C1 / C2 driver phases (initial guess vs solve) on the bench tracks."""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ceres_slam_b200 import synthetic as syn, build as b

with tempfile.TemporaryDirectory() as tmp:
    tr = bench._cut_states(syn.make_track(118, 15, 10, seed=42, pix_sigma=0.25), 9, 100)
    csv = os.path.join(tmp, "c1.csv")
    syn.write_track_csv(tr, csv)
    exe = b.build_host_driver("dataset_vo_b200")
    for window in (2, 2, 0):
        r = subprocess.run([exe, csv, "--window", str(window), "--max-iters", "100"], capture_output=True, text=True, cwd=tmp)
        print([l for l in r.stderr.splitlines() if "timing" in l][-1][-160:])
