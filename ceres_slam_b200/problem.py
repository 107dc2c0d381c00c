"""Host-side handle over the C ABI (include/cslam_b200.h).

`BAProblem` plays the role `ceres::Problem` + `ceres::Solve` play inside the reference's
`solveWindow` (tests/dataset_vo.cpp:22-85): parameter blocks are numpy arrays owned by the
caller and are updated in place by `solve()`.  It binds the CUDA library and nothing else.
"""
import ctypes as C

import numpy as np

from . import capi


class CslamError(RuntimeError):
    pass


def default_options(lib=None, **overrides):
    lib = lib or capi.load_product()
    opt = capi.Options()
    lib.options_init(C.byref(opt))
    for k, v in overrides.items():
        if not hasattr(opt, k):
            raise AttributeError(k)
        setattr(opt, k, v)
    return opt


class BAProblem:
    def __init__(self, **options):
        self.lib = self._library()
        self.options = default_options(self.lib, **options)
        self._h = capi._h()
        self._check(self.lib.problem_create(C.byref(self._h), C.byref(self.options)))
        self._keep = {}
        self.n_stereo = self.n_sun = self.n_prior = self.n_phong = 0

    # -- plumbing ---------------------------------------------------------------------------
    @staticmethod
    def _library():
        return capi.load_product()

    def _check(self, status):
        if status != 0:
            msg = self.lib.last_error(self._h) if self._h else b""
            raise CslamError(f"cslam status {status}: {(msg or b'').decode()}")

    def close(self):
        if self._h:
            self.lib.problem_destroy(self._h)
            self._h = capi._h()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **overrides):
        for k, v in overrides.items():
            if not hasattr(self.options, k):
                raise AttributeError(k)
            setattr(self.options, k, v)
        self._check(self.lib.set_options(self._h, C.byref(self.options)))

    # -- problem construction ---------------------------------------------------------------
    def set_camera(self, fu, fv, cu, cv, b):
        self._check(self.lib.set_camera(self._h, fu, fv, cu, cv, b))

    def set_poses(self, poses12, constant=None):
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        assert poses12.ndim == 2 and poses12.shape[1] == 12
        const = None if constant is None else np.ascontiguousarray(constant, dtype=np.uint8)
        self._keep["poses"], self._keep["const"] = poses12, const
        self._check(self.lib.set_poses(self._h, poses12.shape[0], capi.dptr(poses12), capi.u8ptr(const)))
        return poses12

    def set_points(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        assert xyz.ndim == 2 and xyz.shape[1] == 3
        self._keep["points"] = xyz
        self._check(self.lib.set_points(self._h, xyz.shape[0], capi.dptr(xyz)))
        return xyz

    def add_stereo(self, cam, pt, uvd, W):
        cam = np.ascontiguousarray(cam, dtype=np.uint32)
        pt = np.ascontiguousarray(pt, dtype=np.uint32)
        uvd = np.ascontiguousarray(uvd, dtype=np.float64).reshape(-1, 3)
        W = np.ascontiguousarray(W, dtype=np.float64)
        per_obs = 1 if W.size != 9 else 0
        assert W.size == (9 * cam.size if per_obs else 9)
        self._keep["stereo"] = (cam, pt, uvd, W)
        self.n_stereo = cam.size
        self._check(self.lib.add_stereo(self._h, cam.size, capi.u32ptr(cam), capi.u32ptr(pt),
                                        capi.dptr(uvd), capi.dptr(W), per_obs))

    def add_sun(self, cam, obs_c, ref_g, W2x2, az_thresh=1000.0, zen_thresh=1000.0, huber=0.0):
        cam = np.ascontiguousarray(cam, dtype=np.uint32)
        obs_c = np.ascontiguousarray(obs_c, dtype=np.float64).reshape(-1, 3)
        ref_g = np.ascontiguousarray(ref_g, dtype=np.float64).reshape(-1, 3)
        W2x2 = np.ascontiguousarray(W2x2, dtype=np.float64).reshape(-1, 4)
        assert obs_c.shape[0] == cam.size == ref_g.shape[0] == W2x2.shape[0]
        self.n_sun += cam.size
        self._check(self.lib.add_sun(self._h, cam.size, capi.u32ptr(cam), capi.dptr(obs_c),
                                     capi.dptr(ref_g), capi.dptr(W2x2), az_thresh, zen_thresh, huber))

    def add_pose_prior(self, cam, Tref12, W6x6):
        Tref12 = np.ascontiguousarray(Tref12, dtype=np.float64).reshape(12)
        W6x6 = np.ascontiguousarray(W6x6, dtype=np.float64).reshape(36)
        self.n_prior += 1
        self._check(self.lib.add_pose_prior(self._h, int(cam), capi.dptr(Tref12), capi.dptr(W6x6)))

    # -- lighting blocks (dataset_ba_phong) ---------------------------------------------------
    def set_vertices(self, normals, textures, material_id):
        normals = np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, 3)
        textures = np.ascontiguousarray(textures, dtype=np.float64).reshape(-1)
        material_id = np.ascontiguousarray(material_id, dtype=np.uint32)
        assert normals.shape[0] == textures.size == material_id.size
        self._keep["vertices"] = (normals, textures, material_id)
        self._check(self.lib.set_vertices(self._h, textures.size, capi.dptr(normals), capi.dptr(textures),
                                          capi.u32ptr(material_id)))
        return normals, textures

    def set_textures(self, kd, vertex_texture_id):
        """Texture blocks shared between vertices (one per material in the reference)."""
        kd = np.ascontiguousarray(kd, dtype=np.float64).reshape(-1)
        tid = np.ascontiguousarray(vertex_texture_id, dtype=np.uint32)
        self._keep["textures"] = (kd, tid)
        self._check(self.lib.set_textures(self._h, kd.size, capi.dptr(kd), capi.u32ptr(tid)))
        return kd

    def set_bounds(self, block_kind, lower, upper):
        """SetParameterLowerBound / UpperBound on every material (3) or texture (1) block."""
        kind = {"material": 0, "texture": 1}[block_kind]
        lo = np.ascontiguousarray(lower, dtype=np.float64)
        hi = np.ascontiguousarray(upper, dtype=np.float64)
        assert lo.size == hi.size == (3 if kind == 0 else 1)
        self._check(self.lib.set_bounds(self._h, kind, capi.dptr(lo), capi.dptr(hi)))

    def set_points_constant(self, constant=True):
        """SetParameterBlockConstant / Variable on every vertex position (lighting solves)."""
        self._check(self.lib.set_points_constant(self._h, int(bool(constant))))

    def set_materials(self, phong):
        phong = np.ascontiguousarray(phong, dtype=np.float64).reshape(-1, 3)
        self._keep["materials"] = phong
        self._check(self.lib.set_materials(self._h, phong.shape[0], capi.dptr(phong)))
        return phong

    def set_light(self, light, directional=False):
        light = np.ascontiguousarray(light, dtype=np.float64).reshape(3)
        self._keep["light"] = light
        self._check(self.lib.set_light(self._h, capi.dptr(light), int(directional)))
        return light

    def add_phong(self, cam, vertex, intensity, int_stiffness, normal_obs, W_normal):
        cam = np.ascontiguousarray(cam, dtype=np.uint32)
        vertex = np.ascontiguousarray(vertex, dtype=np.uint32)
        intensity = np.ascontiguousarray(intensity, dtype=np.float64).reshape(-1)
        normal_obs = np.ascontiguousarray(normal_obs, dtype=np.float64).reshape(-1, 3)
        W_normal = np.ascontiguousarray(W_normal, dtype=np.float64).reshape(9)
        assert cam.size == vertex.size == intensity.size == normal_obs.shape[0]
        self._keep["phong"] = (cam, vertex, intensity, normal_obs, W_normal)
        self.n_phong = cam.size
        self._check(self.lib.add_phong(self._h, cam.size, capi.u32ptr(cam), capi.u32ptr(vertex),
                                       capi.dptr(intensity), float(int_stiffness), capi.dptr(normal_obs),
                                       capi.dptr(W_normal)))

    def evaluate_phong(self):
        """Residuals and tangent-space Jacobians of the intensity and normal blocks."""
        n = self.n_phong
        out = {"r_int": np.zeros(n), "J_int": np.zeros((n, 19)), "r_normal": np.zeros((n, 3)),
               "Jpose_normal": np.zeros((n, 3, 6)), "Jn_normal": np.zeros((n, 3, 3))}
        cost = C.c_double(0.0)
        self._check(self.lib.evaluate_phong(self._h, C.byref(cost), *[capi.dptr(out[k]) for k in
                                            ("r_int", "J_int", "r_normal", "Jpose_normal", "Jn_normal")]))
        out["cost"] = cost.value
        return out

    def covariance_block(self, cam):
        """GetCovarianceBlockInTangentSpace(pose cam, pose cam) at the current parameter values."""
        cov = np.zeros((6, 6))
        self._check(self.lib.covariance_block(self._h, int(cam), capi.dptr(cov)))
        return cov

    def time_phong(self, reps):
        ms = C.c_double(0)
        self._check(self.lib.time_phong(self._h, reps, C.byref(ms)))
        return ms.value

    # -- evaluation / solve -----------------------------------------------------------------
    def evaluate(self, apply_loss=True, jacobians=True):
        """ceres::Problem::Evaluate: cost, residuals and tangent-space Jacobians per block."""
        n, m, q = self.n_stereo, self.n_sun, self.n_prior
        out = {
            "r_stereo": np.zeros((n, 3)), "Jpose_stereo": np.zeros((n, 3, 6)),
            "Jpoint_stereo": np.zeros((n, 3, 3)), "r_sun": np.zeros((m, 2)),
            "J_sun": np.zeros((m, 2, 6)), "r_prior": np.zeros((q, 6)), "J_prior": np.zeros((q, 6, 6)),
        }
        cost = C.c_double(0.0)
        ptr = {k: (capi.dptr(v) if v.size else None) for k, v in out.items()}
        if not jacobians:
            for k in ("Jpose_stereo", "Jpoint_stereo", "J_sun", "J_prior"):
                ptr[k] = None
        self._check(self.lib.evaluate(self._h, int(apply_loss), C.byref(cost), ptr["r_stereo"],
                                      ptr["Jpose_stereo"], ptr["Jpoint_stereo"], ptr["r_sun"],
                                      ptr["J_sun"], ptr["r_prior"], ptr["J_prior"]))
        out["cost"] = cost.value
        return out

    def solve(self):
        s = capi.Summary()
        self._check(self.lib.solve(self._h, C.byref(s)))
        return s

    def iteration_log(self):
        n = C.c_int(0)
        self._check(self.lib.get_iteration_log(self._h, None, 0, C.byref(n)))
        rows = np.zeros((n.value, capi.LOG_COLS))
        if n.value:
            self._check(self.lib.get_iteration_log(self._h, capi.dptr(rows), n.value, C.byref(n)))
        return rows

    # -- device-resident API (product only) ---------------------------------------------------
    def upload(self):
        self._check(self.lib.upload(self._h))

    def lm_begin(self):
        self._check(self.lib.lm_begin(self._h))

    def lm_iterate(self, n, ignore_convergence=False):
        s = capi.Summary()
        self._check(self.lib.lm_iterate(self._h, int(n), int(ignore_convergence), C.byref(s)))
        return s

    def download(self):
        self._check(self.lib.download(self._h))

    def reset_state(self):
        self._check(self.lib.reset_state(self._h))

    def set_stream(self, stream_ptr):
        self._check(self.lib.set_stream(self._h, C.c_void_p(stream_ptr)))

    def reduced_system(self):
        nf, nnz = C.c_int(0), C.c_int(0)
        self._check(self.lib.get_reduced_sizes(self._h, C.byref(nf), C.byref(nnz)))
        rowptr = np.zeros(nf.value + 1, dtype=np.int32)
        col = np.zeros(nnz.value, dtype=np.int32)
        val = np.zeros((nnz.value, 6, 6))
        rhs = np.zeros((nf.value, 6))
        ids = np.zeros(nf.value, dtype=np.int32)
        ip = lambda a: a.ctypes.data_as(capi._ip)
        self._check(self.lib.get_reduced_system(self._h, ip(rowptr), ip(col), capi.dptr(val),
                                                capi.dptr(rhs), ip(ids)))
        return rowptr, col, val, rhs, ids

    def profile(self):
        pr = capi.Profile()
        self._check(self.lib.get_profile(self._h, C.byref(pr)))
        return {capi.KERNEL_CLASSES[i]: (pr.ms[i], pr.launches[i]) for i in range(len(capi.KERNEL_CLASSES))}

    def reset_profile(self):
        self._check(self.lib.reset_profile(self._h))

    def time_resjac(self, reps):
        ms = C.c_double(0)
        self._check(self.lib.time_resjac(self._h, reps, C.byref(ms)))
        return ms.value

    def time_schur(self, reps):
        ms = C.c_double(0)
        self._check(self.lib.time_schur(self._h, reps, C.byref(ms)))
        return ms.value

    def analyze(self, n_ranks=1, rank=0):
        """Host-only structure analysis (works without a GPU)."""
        info = capi.StructureInfo()
        self._check(self.lib.analyze(self._h, n_ranks, rank, C.byref(info)))
        return info.as_dict()

    def attach_comm(self, n_ranks, rank, uid):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        assert uid.size == 128
        self._check(self.lib.attach_comm(self._h, n_ranks, rank, capi.u8ptr(uid)))


def solve_batch(problems):
    """cslam_solve_batch: independent small problems, one launch per GPU (config 4)."""
    lib = capi.load_product()
    n = len(problems)
    handles = (capi._h * n)(*[p._h for p in problems])
    sums = (capi.Summary * n)()
    status = lib.solve_batch(handles, n, sums)
    if status != 0:
        raise CslamError(f"cslam_solve_batch status {status}: "
                         f"{(lib.last_error(problems[0]._h) or b'').decode()}")
    return list(sums)
