#!/usr/bin/env python
"""Benchmark of the bundle-adjustment hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA back end
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (the oracle restatement of
                                                             # the reference's Ceres path)

A "step" is one Levenberg-Marquardt iteration of the hot path: fused residual/Jacobian + Schur
elimination, block-Jacobi PCG on the reduced camera system, back-substitution + Plus + candidate
cost.  Workload at every N: BASELINE.json config 5, the large full-batch stereo BA (20 k poses x
2 M landmarks x 20 M observations, synthetic), landmarks sharded over the N GPUs with one NCCL
all-reduce of [S | g] per Schur build ("strong" scaling: the total problem is fixed).
One JSON line on stdout from rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ceres_slam_b200 import synthetic as syn  # noqa: E402
from oracle import pybinding as orc

METRIC = "BA LM iterations/s (full-batch stereo BA, C5: 20k poses x 2M landmarks x 20M obs)"
LM_EXACT = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, linear_solver=0)
LM_ITERATIVE = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0,
                    linear_solver=1, preconditioner=1, eta=0.1, max_linear_solver_iterations=500)
LM_OPTS = dict(LM_EXACT)  # main() switches to LM_ITERATIVE with --linear iterative


def c5_track(scale=1.0, seed=42):
    """Config 5 at `scale` (1.0 = 20 k poses, 100 new landmarks per frame tracked for 10 frames)."""
    n_poses = max(40, int(round(20000 * scale)))
    return syn.make_track(n_poses, 100, 10, seed=seed)


def schur_algorithmic_bytes(n_obs, n_lm, n_cam, nnzU):
    # 32 n_o + 24 n_l + 96 n_c + 288 nnzU + 48 n_c
    return 32 * n_obs + 24 * n_lm + 96 * n_cam + 288 * nnzU + 48 * n_cam


def schur_algorithmic_flops(n_lm, L):
    # n_l (618 L + 108 L (L + 1) + 50)
    return n_lm * (618 * L + 108 * L * (L + 1) + 50)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~4 ms from
    a thread (an `nvidia-smi -lms` child needs longer to start than a 10-step region lasts)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_sm = [], set(), None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].strip().isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                    "hw_thermal_slowdown": 0x40}

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        try:
                            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.004)

            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_oracle_run(steps, warmup, threads, sample_scale):
    """The reference's path on the host cores: the oracle restatement (Jet autodiff functors +
    Ceres-semantics LM + Schur elimination + exact band Cholesky / the same PCG rule) on config 5 at
    `sample_scale` (1.0 = the full problem).  ONE solve of warmup + steps LM iterations; the oracle
    stamps every iteration with the steady clock, so the timed region is exactly the last `steps`."""
    tr = c5_track(sample_scale)
    opts = dict(LM_OPTS, num_threads=threads)
    p, _, _ = orc.build_problem(tr, max_num_iterations=warmup + steps, **opts)
    s = p.solve()
    t = p.iteration_seconds()          # row 0 = the initial evaluation, row r = end of iteration r
    iters_done = len(t) - 1
    w = min(warmup, max(0, iters_done - 1))
    iters = iters_done - w
    t_steps = max(1e-9, float(t[-1] - t[w]))
    p.close()
    n_obs = int(tr["obs_cam"].size)
    return dict(ms_per_iter_sample=1e3 * t_steps / max(1, iters), n_obs_sample=n_obs, iters=iters, warmup=w,
                timed_s=t_steps, final_cost=s.final_cost, n_poses=int(tr["n_poses"]), n_landmarks=int(tr["n_points"]))


REFERENCE_BUDGET_S = 600.0   # the whole CPU arm (generation + W + K iterations) must end within a few minutes


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port; Ceres itself
    is not in the image) with all host threads, on the SAME configuration as the GPU arm — the full
    config 5 for all W + K iterations.  Only if a probe says the host is too slow for the budget is the
    problem scaled down, and then the line says so (`extrapolated: true`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    full_obs = int(c5_obs_count())
    W, K = max(0, args.warmup), max(1, args.steps)
    scale = args.ref_scale
    if scale <= 0:
        probe = cpu_oracle_run(2, 1, threads, 0.02)
        est_full_s = probe["ms_per_iter_sample"] * 1e-3 * full_obs / probe["n_obs_sample"]
        need = est_full_s * (W + K + 1.5) * 1.15 + 25.0      # + initial evaluation / Jacobi scaling, + generation
        scale = 1.0 if need <= REFERENCE_BUDGET_S else max(0.02, (REFERENCE_BUDGET_S - 25.0) / (need - 25.0))
    r = cpu_oracle_run(K, W, threads, scale)
    extrapolated = scale < 1.0
    ms_step = r["ms_per_iter_sample"] * (full_obs / r["n_obs_sample"] if extrapolated else 1.0)
    value = 1e3 / ms_step
    if extrapolated:
        sample = (f"config 5 at {scale:.3g} scale ({r['n_obs_sample']} observations), {r['iters']} timed LM iterations "
                  f"after {r['warmup']}, {threads} threads; per-iteration time SCALED by n_obs to the full problem "
                  "(host too slow for the full problem within the arm's budget)")
    else:
        sample = (f"the full config 5 ({r['n_obs_sample']} observations, {r['n_poses']} poses, {r['n_landmarks']} landmarks), "
                  f"{r['iters']} timed LM iterations after {r['warmup']} warm-up iterations in one solve, {threads} threads; "
                  "nothing extrapolated")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "LM iter/s", "n_gpus": args.gpus,
        "steps": r["iters"], "warmup": r["warmup"], "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus), "extrapolated": extrapolated, "sample_scale": scale,
        "timed_region_s": r["timed_s"],
        "cpu_baseline": {"value": value, "unit": "LM iter/s", "cores": threads, "kind": "port", "sample": sample,
                         "extrapolated": extrapolated, "sample_scale": scale},
        "e2e": {"value": value, "unit": "LM iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "obs_per_s": value * full_obs,
        "note": "CPU restatement of the reference's Ceres SPARSE_SCHUR path (oracle/, functor level pinned to the "
                "reference's own headers; Ceres itself is not in the image)",
    }
    emit(line)


def e2e_transfer_bytes(world, n_obs, n_lm, n_cam):
    """Bytes one cslam_solve call moves between host and devices, summed over the ranks (what
    csrc/engine.cu::upload / download copy): the two index arrays (8 B / observation), the measurements
    (24 B / observation) and the points cross PCIe ONCE — rank r uploads the r-th slice of each and the slices
    are exchanged over NVLink (in-place all-gather) — the poses go to every rank; every rank downloads all
    poses and all points."""
    h2d = n_obs * (8 + 24) + 24 * n_lm + world * 96 * n_cam
    d2h = world * (96 * n_cam + 24 * n_lm)
    return h2d, d2h


def c5_obs_count():
    return 19_991 * 100 * 10  # n_starts * new_per_frame * track_len before visibility filtering


def workload_config(n_gpus):
    return {"workload": "C5 large full-batch stereo BA: 20k poses x 2M landmarks x 20M observations "
                        "(synthetic loop track, 100 new landmarks/frame tracked 10 frames)",
            "lm": ("Ceres-semantics LM, SPARSE_SCHUR-equivalent exact reduced solve (banded block Cholesky), "
                   "first pose constant") if LM_OPTS["linear_solver"] == 0 else
                  "Ceres-semantics LM, ITERATIVE_SCHUR-equivalent (block-Jacobi PCG, eta=0.1), first pose constant",
            "sharding": f"landmarks over {n_gpus} GPU(s), NCCL all-reduce of [S|g]" if n_gpus > 1 else "single GPU",
            "l2": "inputs (640 MB of observations) larger than the 126 MB L2"}


def bench_c4_windows(lib, n_windows=256, iters=6, reps=5, cpu_baseline=True, dogleg=False):
    """BASELINE.json config 4: 256 independent sliding windows (dataset_vo --window 2 shape: 2 poses,
    ~150 landmarks, ~300 observations each) packed into ONE launch; latency-bound, so the figures
    are windows/s and window-LM-iterations/s, not a roofline fraction."""
    import ctypes as C
    from ceres_slam_b200.problem import solve_batch
    tracks = [syn.make_track(100, 15, 10, seed=42 + t) for t in range((n_windows + 89) // 90)]
    wins = []
    for tr in tracks:
        for k1 in range(5, 95):
            if len(wins) < n_windows:
                wins.append(syn.window_of(tr, k1, k1 + 2))
    kw = dict(max_num_iterations=iters, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
    if dogleg:  # the strategy dataset_vo_sun sets (dataset_vo_sun.cpp:142-143): DOGLEG, SUBSPACE_DOGLEG
        kw.update(trust_region_strategy=1, dogleg_type=1)
    best_wall, dev_ms, n_it, n_obs = None, None, 0, sum(int(w["obs_cam"].size) for w in wins)
    for rep in range(reps + 1):
        probs = [syn.build_problem(w, **kw)[0] for w in wins]
        t0 = time.perf_counter()
        sums = solve_batch(probs)
        wall = time.perf_counter() - t0
        if rep > 0 and (best_wall is None or wall < best_wall):
            best_wall, dev_ms = wall, sums[0].device_ms
            n_it = sum(s.num_iterations for s in sums)
        for p in probs:
            p.close()
    cpu = None
    if cpu_baseline:
        # the oracle on the same windows, one window per host thread at a time (each solve single-threaded)
        from concurrent.futures import ThreadPoolExecutor
        threads = os.cpu_count() or 1
        sample = wins[:min(n_windows, 8 * threads)]
        probs = [orc.build_problem(w, num_threads=1, **kw)[0] for w in sample]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            sums_o = list(ex.map(lambda q: q.solve(), probs))
        dt = time.perf_counter() - t0
        cpu = {"value": len(sample) / dt, "unit": "windows/s", "cores": threads, "kind": "port",
               "sample": f"{len(sample)} of the {n_windows} windows, {iters} LM iterations each, one oracle solve per thread",
               "window_lm_iters_per_s": sum(x.num_iterations for x in sums_o) / dt}
        for q in probs:
            q.close()
    return {"windows": n_windows, "lm_iterations_per_window": iters, "observations": n_obs, "cpu_baseline": cpu,
            "kernel_ms": dev_ms, "windows_per_s_kernel": n_windows / (dev_ms * 1e-3),
            "window_lm_iters_per_s_kernel": n_it / (dev_ms * 1e-3),
            "e2e_wall_ms": best_wall * 1e3, "windows_per_s_e2e": n_windows / best_wall,
            "strategy": "SUBSPACE_DOGLEG" if dogleg else "LEVENBERG_MARQUARDT",
            "note": "one CTA per window, trust-region loop on the device, one launch per batch; e2e = host packing + "
                    "H2D + kernel + D2H through cslam_solve_batch"}


def bench_phong_blocks(peak_gbs, reps=10):
    """Lighting blocks of BASELINE.json config 3 (2 k poses x 200 k vertices, ~2 M observations):
    one intensity block + one normal block per observation, residuals and Jacobians materialised
    (the K1-style throughput figure for the Phong residuals)."""
    tr = syn.add_phong(syn.make_track(2000, 100, 10, seed=42))
    p, _ = syn.build_phong_problem(tr)
    n = int(tr["obs_cam"].size)
    ms = p.time_phong(reps)
    p.close()
    bytes_per_obs = 12 + 8 + 24 + 50 * 8  # indices, intensity, observed normal, 50 output doubles
    gbs = bytes_per_obs * n / (ms * 1e-3) / 1e9
    return {"observations": n, "ms": ms, "obs_per_s": n / (ms * 1e-3), "bytes_per_obs": bytes_per_obs,
            "achieved_gbs": gbs, "frac_of_hbm": gbs / peak_gbs,
            "note": "IntensityErrorPointLight + NormalError residuals and tangent-space Jacobians per observation"}


def bench_c3_phong_solve(iters=6, cpu=True):
    """BASELINE.json config 3: dataset_ba_phong's joint solve (2 k poses x 200 k vertices, ~2 M
    observations, each a stereo + an intensity + a normal block; 8 materials and textures, one point
    light, the box of dataset_ba_phong.cpp:143-181).  LM iterations/s with the problem resident in
    HBM, split by kernel class, next to the oracle on a bounded sample (scaled by observations)."""
    fixed = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
    tr = syn.add_phong(syn.make_track(2000, 100, 10, seed=42), shared_textures=True)
    n = int(tr["obs_cam"].size)
    p, _ = syn.build_phong_problem(tr, bounds=True, max_num_iterations=10 ** 6, profile_kernels=1, **fixed)
    t0 = time.perf_counter()
    p.upload()
    up = time.perf_counter() - t0
    p.lm_begin()
    p.lm_iterate(3, ignore_convergence=True)
    p.reset_profile()
    s0 = p.lm_iterate(0, ignore_convergence=True).device_ms
    s = p.lm_iterate(iters, ignore_convergence=True)
    ms = (s.device_ms - s0) / iters
    prof = {k: {"ms": v[0] / max(1, v[1]), "launches": v[1]} for k, v in p.profile().items() if v[1]}
    log = p.iteration_log()
    p.close()
    out = {"poses": int(tr["n_poses"]), "vertices": int(tr["n_points"]), "observations": n, "shared_columns": 35,
           "ms_per_lm_iteration": ms, "lm_iters_per_s": 1e3 / ms, "obs_per_s": n * 1e3 / ms, "step_profile_ms": prof,
           "upload_and_structure_s": up, "cost_first": float(log[0, 1]), "cost_last": float(log[-1, 1]),
           "note": "vertex (position + normal) blocks eliminated on the GPU, arrowhead reduced system: banded "
                   "camera part + 35 dense border columns solved with 36 concurrent banded solves"}
    if cpu:
        trs = syn.add_phong(syn.make_track(60, 100, 10, seed=42), shared_textures=True)
        threads = os.cpu_count() or 1
        po, _ = orc.build_phong_problem(trs, bounds=True, max_num_iterations=3, num_threads=threads, **fixed)
        t0 = time.perf_counter()
        so = po.solve()
        dt = (time.perf_counter() - t0) / max(1, so.num_iterations)
        po.close()
        ns = int(trs["obs_cam"].size)
        out["cpu_baseline"] = {"value": 1.0 / (dt * n / ns), "unit": "LM iter/s", "cores": threads, "kind": "port",
                               "sample": f"60 poses / {ns} observations of the same track, {so.num_iterations} LM iterations, "
                                         "scaled by observations (the oracle factors the reduced system densely, "
                                         "which favours it at this size)"}
    return out


def bench_loop_closure(iters=6, n_poses=500):
    """Closed-loop full batch (scripts/ba_all_sims.sh:8-13 runs closed trajectories; dataset_vo.cpp:118-121
    --window 0): the last poses re-observe the first landmarks, so the reduced camera system has blocks
    far from its diagonal and the exact solve is the dense FP64 Cholesky (DMMA trailing update) instead
    of the banded solver; the PCG fallback run to 1e-15 (dense_solver = -1) is timed beside it."""
    fixed = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
    tr = syn.make_track(n_poses, 100, 10, seed=42, closed=True)
    out = {"poses": n_poses, "landmarks": int(tr["n_points"]), "observations": int(tr["obs_cam"].size),
           "reduced_dimension": 6 * (n_poses - 1)}
    for name, ds in (("dense_cholesky", 0), ("pcg_1e-15", -1)):
        p, _, _ = syn.build_problem(tr, max_num_iterations=10 ** 6, profile_kernels=1, dense_solver=ds, **fixed)
        p.upload()
        p.lm_begin()
        p.lm_iterate(2, ignore_convergence=True)
        p.reset_profile()
        s0 = p.lm_iterate(0, ignore_convergence=True).device_ms
        s = p.lm_iterate(iters, ignore_convergence=True)
        prof = {k: v[0] / max(1, v[1]) for k, v in p.profile().items() if v[1]}
        log = p.iteration_log()
        out[name] = {"ms_per_lm_iteration": (s.device_ms - s0) / iters, "linear_solve_ms": prof.get("linear_solve"),
                     "schur_ms": prof.get("schur"), "linear_iterations_per_solve": float(log[-iters:, 7].mean()),
                     "cost_last": float(log[-1, 1])}
        p.close()
    n = out["reduced_dimension"]
    ms = out["dense_cholesky"]["linear_solve_ms"]
    if ms:
        out["dense_cholesky"]["factor_tflops"] = 2.0 * n ** 3 / 3.0 / (ms * 1e-3) / 1e12
    out["note"] = "factor_tflops = (n^3 / 3 FMA) / whole reduced solve (fill + panels + DMMA updates + both substitutions)"
    return out


def bench_ragged(iters=5, n_poses=5000, mean=8.0, lmax=30, drop=0.1):
    """Tracks as a stereo front end produces them — lengths 2 + Geometric clipped to `lmax`, 10 % drop-outs — next
    to the regular track (every landmark seen by exactly 10 consecutive frames) with the same number of poses and
    landmarks per frame: how many landmarks reach the grouped DMMA Schur kernel (landmarks that share their
    camera list exactly, or whose cameras fit a common window of at most 10), and what an LM iteration costs."""
    fixed = dict(function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
    out = {"poses": n_poses, "landmarks_per_frame": 100,
           "ragged": {"mean_length": mean, "max_length": lmax, "dropout": drop}}
    for name, rg in (("regular", None), ("ragged", dict(mean=mean, max=lmax, drop=drop))):
        tr = syn.make_track(n_poses, 100, 10, seed=42, ragged=rg)
        p, _, _ = syn.build_problem(tr, max_num_iterations=10 ** 6, profile_kernels=1, **dict(LM_EXACT, **fixed))
        info = p.analyze()
        p.upload()
        p.lm_begin()
        p.lm_iterate(2, ignore_convergence=True)
        p.reset_profile()
        s0 = p.lm_iterate(0, ignore_convergence=True).device_ms
        s = p.lm_iterate(iters, ignore_convergence=True)
        prof = {k: v[0] / max(1, v[1]) for k, v in p.profile().items() if v[1]}
        log = p.iteration_log()
        n_obs = int(tr["obs_cam"].size)
        out[name] = {"observations": n_obs, "landmarks": int(info["n_landmarks"]),
                     "grouped_fraction": info["n_grouped_landmarks"] / max(1, info["n_landmarks"]),
                     "groups": int(info["n_groups"]), "ms_per_lm_iteration": (s.device_ms - s0) / iters,
                     "step_profile_ms": prof, "ns_per_observation": (s.device_ms - s0) / iters * 1e6 / n_obs,
                     "cost_first": float(log[0, 1]), "cost_last": float(log[-1, 1])}
        p.close()
    out["ragged_over_regular_per_observation"] = out["ragged"]["ns_per_observation"] / out["regular"]["ns_per_observation"]
    out["ragged_over_regular_per_iteration"] = out["ragged"]["ms_per_lm_iteration"] / out["regular"]["ms_per_lm_iteration"]
    out["note"] = ("ragged: exact + ragged landmark groups and wide-window slices of the long tracks on the FP64 tensor cores "
                   "(K2 / K2w), reduced system of half-bandwidth ~29 blocks factored exactly as two levels of chunked bordered "
                   "bands (K3e); regular: grouped DMMA kernel + narrow-band solver (K3b)")
    return out


def bench_ransac_front_end(n_poses=1000):
    """the RANSAC front end (compute_initial_guess's 400-hypothesis point-cloud
    alignment per consecutive pose pair) for a 1 k-pose track with ~900 matches per pair, all pairs
    in ONE launch; wall time of cslam_ransac_align from host arrays (H2D + kernel + D2H)."""
    from ceres_slam_b200 import initial_guess as ig
    tr = syn.make_track(n_poses, 100, 10, seed=42, pix_sigma=0.25)
    rng = ig.state_ranges(tr["obs_cam"], tr["n_poses"])
    pt = tr["obs_pt"].astype(np.int64)
    p0, p1 = [], []
    for k in range(1, tr["n_poses"]):
        kp, kc = ig.match_pair(pt[rng[k - 1]:rng[k]], pt[rng[k]:rng[k + 1]])
        p0.append(ig.triangulate(tr["cam"], tr["uvd"][rng[k - 1]:rng[k]][kp]))
        p1.append(ig.triangulate(tr["cam"], tr["uvd"][rng[k]:rng[k + 1]][kc]))
    ig.ransac_align(p0[:8], p1[:8], tr["cam"])          # context / module warm-up
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        T, inl, cnt = ig.ransac_align(p0, p1, tr["cam"])
        best = min(best, time.perf_counter() - t0)
    n_pts = int(sum(p.shape[0] for p in p0))
    hyp_pts = 400.0 * n_pts
    return {"pose_pairs": len(p0), "correspondences": n_pts, "hypotheses_per_pair": 400, "wall_ms": best * 1e3,
            "pairs_per_s": len(p0) / best, "hypothesis_point_tests_per_s": hyp_pts / best,
            "median_inlier_fraction": float(np.median(cnt / np.maximum(1, [p.shape[0] for p in p0]))),
            "note": "host lists -> one concatenation + H2D + one kernel launch + D2H (python packing included)"}


def _cut_states(tr, k0, n):
    keep = (tr["obs_cam"] >= k0) & (tr["obs_cam"] < k0 + n)
    out = dict(tr, n_poses=n, obs_cam=(tr["obs_cam"][keep] - k0).astype(np.uint32), obs_pt=tr["obs_pt"][keep].copy(),
               uvd=tr["uvd"][keep].copy(), poses=tr["poses"][k0:k0 + n].copy(), poses_gt=tr["poses_gt"][k0:k0 + n].copy())
    if np.asarray(tr["W"]).size != 9:
        out["W"] = tr["W"][keep].copy()
    for k in ("sun_obs_c", "sun_ref_g", "sun_W"):
        if k in tr:
            out[k] = tr[k][k0:k0 + n].copy()
    if "sun_cam" in tr:
        out["sun_cam"] = np.arange(n, dtype=np.uint32)
    return out


def _driver_timing(stderr):
    out = []
    for line in stderr.splitlines():
        if "cslam_b200 timing:" in line:      # (may follow an unterminated progress message on the same line)
            kv = dict(t.split("=") for t in line.split("cslam_b200 timing:", 1)[1].split())
            out.append((int(kv["windows"]), float(kv["loop_s"]), kv.get("pass", "vo"), kv))
    return out


_REF_DRIVER_RUNNER = r"""
import ctypes as C, sys, time
lib = C.CDLL(sys.argv[1])
main = getattr(lib, "cslam_ref_" + sys.argv[2] + "_main")
main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
args = [sys.argv[2].encode()] + [a.encode() for a in sys.argv[3:]]
t0 = time.perf_counter()
rc = main(len(args), (C.c_char_p * len(args))(*args))
sys.stderr.write("\nreference_driver_wall_s=%.6f\n" % (time.perf_counter() - t0))
sys.exit(rc)
"""


def reference_driver_baseline(driver, files, flags, cwd):
    """CPU figure for configs 1 / 2 from the REFERENCE'S OWN driver source (tests/dataset_vo.cpp, tests/dataset_vo_sun.cpp,
    unmodified) compiled over the Ceres-API facade with the CPU oracle answering ceres::Solve / ceres::Covariance
    (oracle/_ref/libref_<driver>_oracle.so, prebuilt by `make -C oracle ref`): no Python in the window loop.
    Wall time of the driver's whole `main` (its CSV reader and writer included; one solve per Report line).  Never
    raises: a missing library or a failed run is reported as {"unavailable": why}."""
    try:
        so = os.path.join(ROOT, "oracle", "_ref", f"libref_{driver}_oracle.so")
        if not os.path.exists(so):
            return {"unavailable": f"{os.path.relpath(so, ROOT)} not built (make -C oracle ref needs /root/reference)"}
        r = subprocess.run([sys.executable, "-c", _REF_DRIVER_RUNNER, so, driver] + list(files) + list(flags), capture_output=True,
                           text=True, cwd=cwd, timeout=900)
        if r.returncode != 0:
            return {"unavailable": "reference driver exited with %d: %s" % (r.returncode, r.stderr[-300:])}
        wall = float(r.stderr.rsplit("reference_driver_wall_s=", 1)[1].split()[0])
        its = [int(l.split("Iterations:")[1].split(",")[0]) for l in r.stdout.splitlines() if "Iterations:" in l]
        return {"value": len(its) / wall, "unit": "windows/s", "cores": 1, "kind": "reference driver + port solver",
                "windows": len(its), "wall_s": wall, "lm_iters_per_s": sum(its) / wall,
                "sample": f"the reference's own tests/{driver}.cpp, unmodified, over the Ceres-API facade on the CPU oracle: the whole "
                          "run of its main on the same files (CSV reader / writer included)"}
    except Exception as e:  # noqa: BLE001 - a baseline must not take the bench line down
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def bench_c1_c2_drivers(cpu=True):
    """BASELINE.json configs 1 and 2 through the restated C++ drivers (host/dataset_vo_b200,
    host/dataset_vo_sun_b200): sequential sliding windows, each = RANSAC initial guess + window solve
    (+ marginal covariance -> prior of the next window for config 2), windows/s from the driver's own
    clock around its window loop (CSV input / output excluded).  CPU beside it: the same window loop on
    the oracle (oracle/driver_mirror.py) — single-threaded like the reference's per-window Ceres solve
    of a 300-observation problem; config 2's CPU figure is timed on a 100-window sample."""
    import tempfile
    from ceres_slam_b200 import build as b
    from oracle import driver_mirror as dm
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- config 1: 100 poses, ~150 landmarks/frame ------------------------------------------
        tr = _cut_states(syn.make_track(118, 15, 10, seed=42, pix_sigma=0.25), 9, 100)
        csv = os.path.join(tmp, "c1.csv")
        syn.write_track_csv(tr, csv)
        exe = b.build_host_driver("dataset_vo_b200")
        res = {}
        for window in (2, 0):
            best = None
            for _ in range(3):
                r = subprocess.run([exe, csv, "--window", str(window), "--max-iters", "100"], capture_output=True, text=True, cwd=tmp)
                if r.returncode != 0:
                    raise RuntimeError(r.stderr[-1000:])
                nw, sec, _p, kv = _driver_timing(r.stderr)[0]
                if best is None or sec < best:
                    best, phases = sec, {k: float(kv[k]) * 1e3 for k in ("initial_guess_s", "solve_s", "warmup_s") if k in kv}
            its = [int(l.split("Iterations:")[1].split(",")[0]) for l in r.stdout.splitlines() if "Iterations:" in l]
            key = "window2" if window == 2 else "full_batch"
            res[key] = {"windows": nw, "loop_ms": best * 1e3, "windows_per_s": nw / best, "lm_iterations": int(sum(its)),
                        "lm_iters_per_s": sum(its) / best, "phases_ms": phases}
            if cpu:
                var = 1.0 / np.diag(np.asarray(tr["W"]).reshape(3, 3)) ** 2
                its_o = []
                t0 = time.perf_counter()
                dm.dataset_vo(tr, var, tr["poses_gt"][0], window, 100, on_window=lambda k1, s_: its_o.append(s_.num_iterations))
                dt = time.perf_counter() - t0
                res[key]["cpu_baseline"] = {"value": len(its_o) / dt, "unit": "windows/s", "cores": 1, "kind": "port",
                                            "lm_iters_per_s": sum(its_o) / dt,
                                            "sample": "the same track and window loop on the oracle (RANSAC + solve per window)"}
                res[key]["cpu_reference_driver"] = reference_driver_baseline("dataset_vo", [csv], ["--window", str(window)], tmp)
        out["c1_dataset_vo"] = dict(res, track="100 poses, ~150 landmarks/frame, 14.6 k stereo observations",
                                    note="dataset_vo_b200: window 2 = 99 sequential windows (1 free pose, ~300 blocks each), "
                                         "full batch = one RANSAC launch over 99 pairs + one bundle adjustment; the driver "
                                         "runs its first window once, untimed, before the loop (phases_ms.warmup_s: CUDA "
                                         "context creation + kernel module loading, once per process)")
        # ---- config 2: 1 k poses, sun blocks, prior chain ------------------------------------------
        trs = _cut_states(syn.add_sun(syn.make_track(1018, 15, 10, seed=42, per_obs_W=True, pix_sigma=0.25), sigma_deg=1.0), 9, 1000)
        paths = [os.path.join(tmp, f) for f in ("c2.csv", "c2_ref.csv", "c2_obs.csv")]
        syn.write_sun_csvs(trs, *paths)
        exe = b.build_host_driver("dataset_vo_sun_b200")
        res = {}
        for strategy in ("dogleg", "lm"):
            r = subprocess.run([exe, *paths, "--window", "2", "--huber-param", "1.0", "--max-iters", "100", "--strategy", strategy],
                               capture_output=True, text=True, cwd=tmp)
            if r.returncode != 0:
                raise RuntimeError(r.stderr[-1000:])
            tm = _driver_timing(r.stderr)
            nw = sum(t[0] for t in tm)
            sec = sum(t[1] for t in tm)
            its = [int(l.split("Iterations:")[1].split(",")[0]) for l in r.stdout.splitlines() if "Iterations:" in l]
            res[strategy] = {"windows": nw, "loop_ms": sec * 1e3, "windows_per_s": nw / sec, "lm_iterations": int(sum(its)),
                             "lm_iters_per_s": sum(its) / sec, "passes": [{"pass": t[2], "windows": t[0], "loop_ms": t[1] * 1e3} for t in tm]}
        if cpu:
            n = 101
            cut = _cut_states(trs, 0, n)
            W = np.asarray(cut["W"]).reshape(-1, 3, 3)
            cov = np.stack([0.5 * (c + c.T) for c in (np.linalg.inv(w @ w) for w in W)]).reshape(-1, 9)
            sW = cut["sun_W"].reshape(-1, 2, 2)
            sun = dict(dir_g=cut["sun_ref_g"], obs=cut["sun_obs_c"],
                       covars=np.stack([np.linalg.inv(w @ w) for w in sW]).reshape(-1, 4), has=np.ones(n, dtype=bool))
            its_o = []
            t0 = time.perf_counter()
            dm.dataset_vo_sun(cut, cov, sun, cut["poses_gt"][0], 2, 100, use_sun=True, huber=1.0, dogleg=True,
                              on_window=lambda k1, s_: its_o.append(s_.num_iterations))
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": len(its_o) / dt, "unit": "windows/s", "cores": 1, "kind": "port",
                                   "lm_iters_per_s": sum(its_o) / dt,
                                   "sample": "the sun pass over the first 100 windows of the same track on the oracle (RANSAC + "
                                             "SUBSPACE_DOGLEG solve + sparse-LU covariance per window)"}
            res["cpu_reference_driver"] = reference_driver_baseline("dataset_vo_sun", paths, ["--window", "2", "--huber-param", "1.0"], tmp)
        out["c2_dataset_vo_sun"] = dict(res, track="1000 poses, ~150 landmarks/frame, per-observation covariances, sun + prior blocks",
                                        note="dataset_vo_sun_b200 --window 2, both passes (VO, then with sun blocks): 2 x 999 sequential "
                                             "windows, each RANSAC + solve + covariance block; dogleg = the reference's SUBSPACE_DOGLEG")
    return out


_RESULT_FD = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="workload scale (1.0 = config 5)")
    ap.add_argument("--cpu-scale", type=float, default=0.25,
                    help="bounded sample of config 5 timed on the CPU beside the GPU line (cpu_baseline)")
    ap.add_argument("--ref-scale", type=float, default=0.0,
                    help="--impl reference: scale of config 5 (default 0 = the full problem unless a probe says it "
                         "cannot finish within the arm's budget)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--no-phong", action="store_true")
    ap.add_argument("--linear", default="exact", choices=["exact", "iterative"],
                    help="reduced-system solve: exact = SPARSE_SCHUR-equivalent (banded direct solver), "
                         "iterative = ITERATIVE_SCHUR-equivalent (block-Jacobi PCG, eta = 0.1)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: libraries that print there on their own (NCCL
    # with NCCL_DEBUG=VERSION does) are sent to stderr for the whole run
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    LM_OPTS.clear()
    LM_OPTS.update(LM_EXACT if args.linear == "exact" else LM_ITERATIVE)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from ceres_slam_b200 import capi
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cslam_b200 back end has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load_product()
    W = max(3, args.warmup)
    K = args.steps

    t0 = time.perf_counter()
    tr = c5_track(args.scale)
    n_obs, n_lm, n_cam = int(tr["obs_cam"].size), int(tr["n_points"]), int(tr["n_poses"])
    gen_s = time.perf_counter() - t0

    def make_problem(**extra):
        p, poses, points = syn.build_problem(tr, device=local, **dict(LM_OPTS, **extra))
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (C.c_uint8 * 128)()
                assert lib.comm_unique_id(buf) == 0
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            uid = uid.cuda()
            dist.broadcast(uid, 0)
            p.attach_comm(world, rank, uid.cpu().numpy())
        return p, poses, points

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (value) ------------------------------------------------------
    p, _, _ = make_problem(max_num_iterations=10 ** 6)
    stream = torch.cuda.current_stream()
    p.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    p.upload()
    upload_s = time.perf_counter() - t0
    p.lm_begin()
    p.lm_iterate(W, ignore_convergence=True)
    launches0 = C.c_uint64(0)
    lib.get_launch_count(C.byref(launches0))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    s = p.lm_iterate(K, ignore_convergence=True)
    ev1.record(stream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches1 = C.c_uint64(0)
    lib.get_launch_count(C.byref(launches1))
    log = p.iteration_log()
    value = K / (ms * 1e-3)

    # ---- per-kernel-class profile and roofline of the dominant kernel (separate short run) ----
    p.set_options(profile_kernels=1)
    p.reset_profile()
    p.lm_iterate(3, ignore_convergence=True)
    prof = p.profile()
    p.set_options(profile_kernels=0)
    # (the sampler also covers the profile iterations above: the same workload, still under load — a
    # 10-step timed region alone lasts ~55 ms, a handful of NVML polls)
    clocks = sampler.stop() if rank == 0 else None
    nf, nnz = C.c_int(0), C.c_int(0)
    lib.get_reduced_sizes(p._h, C.byref(nf), C.byref(nnz))
    schur_ms = max_over_ranks(prof["schur"][0] / max(1, prof["schur"][1]))
    peak, peak_src = measured_peaks()
    alg_bytes = schur_algorithmic_bytes(n_obs // world, n_lm // world, n_cam, nnz.value)
    achieved = alg_bytes / (schur_ms * 1e-3) / 1e9
    L = n_obs / max(1, n_lm)
    alg_flops = schur_algorithmic_flops(n_lm // world, L)
    fp64 = C.c_double(0)
    lib.measure_fp64_peak(local, C.byref(fp64))
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and args.scale == 1.0:
            traffic = tj["schur_grouped2_kernel"]["bytes"]  # dram read + write of ONE launch, `ncu --set full`
            traffic_src = tj["schur_grouped2_kernel"].get("source", "profiles/traffic.json (committed ncu --set full capture)")
    except (OSError, KeyError, ValueError):
        pass
    tf = alg_flops / (schur_ms * 1e-3) / 1e12
    # The binding roof of this kernel is FP64 arithmetic (DFMA + DMMA share one pipe and one peak): 36 GFLOP vs
    # 0.75 GB per launch.  MEASURED_PEAKS.json carries no FP64 entry, so the peak is measured in this run.
    roofline = {"kernel": "schur_grouped2_kernel (fused residual/Jacobian + Schur elimination, FP64 DMMA)",
                "bound": "tensor", "precision": "fp64 (DMMA.8x8x4 and DFMA: one pipe, one peak)",
                "achieved": tf, "peak": fp64.value, "unit": "TFLOP/s", "frac": tf / max(fp64.value, 1e-9),
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "FP64 peak measured in this run (register-resident DFMA chains on all SMs; "
                               "MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_launch": alg_flops, "algorithmic_bytes_per_launch": alg_bytes,
                "ms_per_launch": schur_ms,
                "flops_per_unit": "n_landmarks x (618 L + 108 L (L + 1) + 50), L = observations per landmark (SURVEY.md 8d)"}
    roofline_hbm = {"kernel": "schur_build", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_source": peak_src,
                    "note": "secondary roof: the kernel moves its algorithmic bytes once (traffic ~ algorithmic) and is "
                            "nowhere near the HBM roof; FP64 arithmetic binds"}
    step_profile = {k: {"ms": v[0] / max(1, v[1]), "launches": v[1]} for k, v in prof.items() if v[1]}
    allreduce = None
    if world > 1 and prof.get("allreduce", (0, 0))[1]:
        # [S | Bdiag | rhs | g | scalars] summed over the ranks once per Schur build
        red_bytes = 8 * (36 * nnz.value + 36 * nf.value + 12 * nf.value + 32)
        ar_ms = max_over_ranks(prof["allreduce"][0] / prof["allreduce"][1])
        allreduce = {"bytes": red_bytes, "ms": ar_ms, "algorithm_gbs": red_bytes / (ar_ms * 1e-3) / 1e9,
                     "bus_gbs": 2 * (world - 1) / world * red_bytes / (ar_ms * 1e-3) / 1e9}

    # ---- materialised residual + Jacobian throughput (K1) -------------------------------------
    resjac = None
    if rank == 0:
        rj_ms = p.time_resjac(5)
        resjac = {"obs_per_s": n_obs / (rj_ms * 1e-3), "ms": rj_ms, "bytes_per_obs": 272,
                  "achieved_gbs": 272 * n_obs / (rj_ms * 1e-3) / 1e9,
                  "frac_of_hbm": 272 * n_obs / (rj_ms * 1e-3) / 1e9 / peak}
    p.close()
    c4 = bench_c4_windows(lib, cpu_baseline=not args.no_cpu and world == 1) if (rank == 0 and not args.no_c4) else None
    c4_dl = bench_c4_windows(lib, cpu_baseline=not args.no_cpu and world == 1, dogleg=True) if (rank == 0 and not args.no_c4) else None
    phong = bench_phong_blocks(peak) if (rank == 0 and not args.no_phong) else None
    c3 = bench_c3_phong_solve(cpu=not args.no_cpu) if (rank == 0 and world == 1 and not args.no_phong) else None
    ransac = bench_ransac_front_end() if (rank == 0 and world == 1 and not args.no_c4) else None
    loop = bench_loop_closure() if (rank == 0 and world == 1 and not args.no_c4) else None
    ragged = bench_ragged() if (rank == 0 and world == 1 and not args.no_c4) else None
    drivers = bench_c1_c2_drivers(cpu=not args.no_cpu) if (rank == 0 and world == 1 and not args.no_c4) else None

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        pe, poses_e, points_e = make_problem(max_num_iterations=K)
        barrier()
        t0 = time.perf_counter()
        se = pe.solve()
        torch.cuda.synchronize()
        wall = max_over_ranks(time.perf_counter() - t0)
        pe.close()
        # the fixed part of that call, measured on its own: upload (structure analysis + H2D) and download
        pf, _, _ = make_problem(max_num_iterations=K)
        barrier()
        t0 = time.perf_counter()
        pf.upload()
        torch.cuda.synchronize()
        up_ms = max_over_ranks(time.perf_counter() - t0) * 1e3
        pf.lm_begin()
        pf.lm_iterate(1, ignore_convergence=True)
        barrier()
        t0 = time.perf_counter()
        pf.download()
        torch.cuda.synchronize()
        down_ms = max_over_ranks(time.perf_counter() - t0) * 1e3
        pf.close()
        h2d, d2h = e2e_transfer_bytes(world, n_obs, n_lm, n_cam)
        e2e = {"value": se.num_iterations / wall, "unit": "LM iter/s", "h2d_bytes_per_step": h2d / max(1, K),
               "d2h_bytes_per_step": d2h / max(1, K), "wall_s": wall, "iterations": se.num_iterations,
               "h2d_bytes_total": h2d, "d2h_bytes_total": d2h,
               "fixed_ms": {"upload_structure_h2d": up_ms, "download_d2h": down_ms, "total": up_ms + down_ms},
               "per_iteration_ms": (wall * 1e3 - up_ms - down_ms) / max(1, se.num_iterations),
               "note": "one cslam_solve call from the caller's (pageable) host arrays: structure analysis + H2D + K LM "
                       "iterations + D2H; the upload is paid once per call, so value moves with --steps: fixed_ms is the "
                       "part that does not amortise; bytes are summed over the ranks"}

    # ---- CPU baseline (rank 0, N = 1 only): a bounded sample of the same workload ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        r = cpu_oracle_run(3, 1, threads, args.cpu_scale)
        ms_full = r["ms_per_iter_sample"] * n_obs / r["n_obs_sample"]
        cpu = {"value": 1e3 / ms_full, "unit": "LM iter/s", "cores": threads, "kind": "port",
               "sample": f"config 5 at {args.cpu_scale:g} scale ({r['n_obs_sample']} observations), "
                         f"{r['iters']} timed LM iterations after {r['warmup']}, per-iteration time scaled by n_obs to the "
                         "full problem; `bench.py --impl reference` runs the full problem",
               "extrapolated": True, "sample_scale": args.cpu_scale, "ms_per_iter_sample": r["ms_per_iter_sample"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "LM iter/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "obs_per_s": value * n_obs, "n_obs": n_obs, "n_landmarks": n_lm, "n_poses": n_cam,
            "e2e": e2e, "gpu_launches": int(launches1.value - launches0.value), "clocks": clocks,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
            "resjac": resjac, "step_profile_ms": step_profile, "allreduce": allreduce, "c4_windows": c4, "c4_windows_dogleg": c4_dl, "phong_blocks": phong,
            "c3_phong_solve": c3, "ransac_front_end": ransac, "loop_closure_dense_solve": loop, "c5_ragged": ragged,
            "c1_dataset_vo": (drivers or {}).get("c1_dataset_vo"), "c2_dataset_vo_sun": (drivers or {}).get("c2_dataset_vo_sun"),
            "lm": {"cost_first": float(log[0, 1]), "cost_last": float(log[-1, 1]),
                   "linear_iterations_timed": int(log[-K:, 7].sum()), "accepted_timed": int(log[-K:, 9].sum())},
            "setup_s": {"generate": gen_s, "upload_and_structure": upload_s},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
