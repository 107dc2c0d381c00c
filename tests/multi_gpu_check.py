"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Landmarks are sharded over the ranks; after every Schur build the partial reduced systems are
summed with one NCCL all-reduce.  The sharded solve must give the same iterates as the same
problem solved on one GPU (1e-9 relative: only the summation order of S differs)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ceres_slam_b200 import capi  # noqa: E402
from ceres_slam_b200 import synthetic as syn  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load_product()
    tr_shared = syn.add_sun(syn.make_track(300, 40, 8, seed=77))
    tr_per_obs = syn.add_sun(syn.make_track(200, 30, 6, seed=78, per_obs_W=True))   # a weight per observation
    # tracks of up to 24 frames with drop-outs: ragged groups and wide-window slices per shard, the wide-band solver
    # (chunks of bordered bands) on every rank
    tr_ragged = syn.add_sun(syn.make_track(400, 12, 6, seed=79, ragged=dict(mean=10, max=24, drop=0.1)))
    kw = dict(max_num_iterations=6, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0,
              device=local, sun=True)
    # (with CSLAM_GPU_STRUCTURE_MIN=0 the ranks analyse the structure on their GPUs; CSLAM_VERIFY_STRUCTURE=1
    # checks that layout against the host analysis)
    for linear_solver, tr in ((0, tr_shared), (1, tr_shared), (0, tr_per_obs), (0, tr_ragged)):
        # reference: the whole problem on this rank's GPU
        p1, poses1, points1 = syn.build_problem(tr, linear_solver=linear_solver, **kw)
        s1 = p1.solve()
        log1 = p1.iteration_log()
        # sharded
        pn, posesn, pointsn = syn.build_problem(tr, linear_solver=linear_solver, **kw)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            assert lib.comm_unique_id(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        pn.attach_comm(world, rank, uid.cpu().numpy())
        sn = pn.solve()
        logn = pn.iteration_log()
        # exact Schur solves agree to rounding; with the inexact (eta = 0.1) PCG the iterate is
        # sensitive to the summation order of S at the 1e-7 level (same as run-to-run on one GPU)
        tol_c, tol_x = (1e-9, 1e-9) if linear_solver == 0 else (1e-6, 1e-5)
        assert sn.num_iterations == s1.num_iterations, (sn.num_iterations, s1.num_iterations)
        assert np.allclose(logn[:, 1], log1[:, 1], rtol=tol_c), (logn[:, 1], log1[:, 1])
        assert np.array_equal(logn[:, 9], log1[:, 9])
        assert np.abs(posesn - poses1).max() <= tol_x * np.abs(poses1).max()
        # every rank's point array receives the COMPLETE solution (the landmarks it does not own are gathered
        # from their owners in the collective download)
        assert np.abs(pointsn - points1).max() <= tol_x * np.abs(points1).max()
        touched_n = np.any(pointsn != tr["points"], axis=1)
        touched_1 = np.any(points1 != tr["points"], axis=1)
        assert np.array_equal(touched_n, touched_1)
        tp = torch.from_numpy(pointsn.copy()).cuda()
        tp0 = tp.clone()
        dist.broadcast(tp0, 0)
        assert torch.equal(tp, tp0)
        # all ranks hold identical poses
        t = torch.from_numpy(posesn.copy()).cuda()
        t0 = t.clone()
        dist.broadcast(t0, 0)
        assert torch.equal(t, t0)
        if rank == 0:
            print(f"linear_solver={linear_solver}: {world}-GPU solve matches 1-GPU "
                  f"(cost {sn.initial_cost:.6e} -> {sn.final_cost:.6e}, {sn.num_iterations} iterations)")
    # the joint lighting solve of dataset_ba_phong (config 3): vertices sharded, the arrowhead reduced system
    # [S_cc S_cg; S_gc S_gg] all-reduced; LM with the box and SUBSPACE_DOGLEG as the reference sets it
    trp = syn.add_phong(syn.make_track(120, 30, 6, seed=79), shared_textures=True)
    for strategy in (0, 1):
        kwp = dict(max_num_iterations=5, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, device=local,
                   trust_region_strategy=strategy, dogleg_type=1)
        p1, st1 = syn.build_phong_problem(trp, bounds=True, **kwp)
        s1 = p1.solve()
        log1 = p1.iteration_log()
        pn, stn = syn.build_phong_problem(trp, bounds=True, **kwp)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            assert lib.comm_unique_id(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        pn.attach_comm(world, rank, uid.cpu().numpy())
        sn = pn.solve()
        logn = pn.iteration_log()
        assert sn.num_iterations == s1.num_iterations, (sn.num_iterations, s1.num_iterations)
        assert np.allclose(logn[:, 1], log1[:, 1], rtol=1e-9), (logn[:, 1], log1[:, 1])
        assert np.array_equal(logn[:, 9], log1[:, 9])
        for k in ("poses", "points", "normals", "phong", "textures", "light"):
            assert np.abs(stn[k] - st1[k]).max() <= 1e-8 * max(1.0, np.abs(st1[k]).max()), k
            t = torch.from_numpy(np.ascontiguousarray(stn[k], dtype=np.float64).copy()).cuda()
            t0 = t.clone()
            dist.broadcast(t0, 0)
            assert torch.equal(t, t0), k          # every rank returns the complete, identical solution
        if rank == 0:
            print(f"lighting solve, strategy {strategy}: {world}-GPU solve matches 1-GPU "
                  f"(cost {sn.initial_cost:.6e} -> {sn.final_cost:.6e}, {sn.num_iterations} iterations)")
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
