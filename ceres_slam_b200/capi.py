"""ctypes binding of the C ABI in include/cslam_b200.h.

`load_product()` binds `libcslam_b200.so` — hand-written sm_100a CUDA kernels behind the C ABI — and
fails loudly when it has not been built: there is no CPU fallback and nothing in this package
knows of any other implementation.  (`PROBLEM_API` is the signature table of the problem-building
entry points; the test-side checker under `oracle/` binds its own library with the same table.)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_SO = os.environ.get("CSLAM_B200_LIB") or os.path.join(ROOT, "ceres_slam_b200", "csrc", "libcslam_b200.so")

LOG_COLS = 10
LOG_NAMES = ("iteration", "cost", "cost_change", "gradient_max_norm", "step_norm",
             "relative_decrease", "radius", "linear_iterations", "step_is_valid",
             "step_is_successful")

K_RESJAC, K_COLNORM, K_SCHUR, K_FINALIZE, K_PCG, K_BACKSUB, K_ALLREDUCE, K_WINDOW, K_OTHER = range(9)
KERNEL_CLASSES = ("resjac", "colnorm", "schur", "finalize", "linear_solve", "backsub", "allreduce",
                  "window", "other")


class Options(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int),
        ("use_nonmonotonic_steps", C.c_int),
        ("max_consecutive_nonmonotonic_steps", C.c_int),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("jacobi_scaling", C.c_int),
        ("linear_solver", C.c_int),
        ("preconditioner", C.c_int),
        ("eta", C.c_double),
        ("max_linear_solver_iterations", C.c_int),
        ("min_linear_solver_iterations", C.c_int),
        ("num_threads", C.c_int),
        ("device", C.c_int),
        ("profile_kernels", C.c_int),
        ("schur_path", C.c_int),
        ("band_leaves", C.c_int),
        ("window_path", C.c_int),
        ("band_separator_solver", C.c_int),
        ("trust_region_strategy", C.c_int),
        ("dogleg_type", C.c_int),
        ("line_search_sufficient_function_decrease", C.c_double),
        ("dense_solver", C.c_int),
        ("bandpc_solver", C.c_int),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("num_iterations", C.c_int),
        ("num_successful_steps", C.c_int),
        ("num_unsuccessful_steps", C.c_int),
        ("termination_type", C.c_int),
        ("termination_reason", C.c_int),
        ("final_radius", C.c_double),
        ("total_linear_iterations", C.c_int),
        ("device_ms", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class StructureInfo(C.Structure):
    _fields_ = [("n_free_cams", C.c_int), ("n_landmarks", C.c_int), ("n_observations", C.c_longlong),
                ("nnz_blocks", C.c_int), ("pattern_hash", C.c_ulonglong), ("n_groups", C.c_int),
                ("n_grouped_landmarks", C.c_int), ("n_work_items", C.c_int), ("landmark_id_sum", C.c_ulonglong),
                ("layout_hash", C.c_ulonglong)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * 16), ("launches", C.c_int * 16)]


_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)
_h = C.c_void_p

# name -> (restype, argtypes): the problem-building entry points of include/cslam_b200.h
PROBLEM_API = {
    "options_init": (None, [C.POINTER(Options)]),
    "problem_create": (C.c_int, [C.POINTER(_h), C.POINTER(Options)]),
    "problem_destroy": (None, [_h]),
    "last_error": (C.c_char_p, [_h]),
    "set_options": (C.c_int, [_h, C.POINTER(Options)]),
    "set_camera": (C.c_int, [_h] + [C.c_double] * 5),
    "set_poses": (C.c_int, [_h, C.c_uint32, _dp, _u8p]),
    "set_points": (C.c_int, [_h, C.c_uint32, _dp]),
    "add_stereo": (C.c_int, [_h, C.c_uint64, _u32p, _u32p, _dp, _dp, C.c_int]),
    "add_sun": (C.c_int, [_h, C.c_uint32, _u32p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double]),
    "add_pose_prior": (C.c_int, [_h, C.c_uint32, _dp, _dp]),
    "evaluate": (C.c_int, [_h, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "solve": (C.c_int, [_h, C.POINTER(Summary)]),
    "get_iteration_log": (C.c_int, [_h, _dp, C.c_int, _ip]),
    "set_vertices": (C.c_int, [_h, C.c_uint32, _dp, _dp, _u32p]),
    "set_textures": (C.c_int, [_h, C.c_uint32, _dp, _u32p]),
    "set_materials": (C.c_int, [_h, C.c_uint32, _dp]),
    "set_light": (C.c_int, [_h, _dp, C.c_int]),
    "set_bounds": (C.c_int, [_h, C.c_int, _dp, _dp]),
    "set_points_constant": (C.c_int, [_h, C.c_int]),
    "add_phong": (C.c_int, [_h, C.c_uint64, _u32p, _u32p, _dp, C.c_double, _dp, _dp]),
}
_RANSAC_SIG = (C.c_int, [C.c_int, C.c_uint32, _u32p, _dp, _dp, _dp, C.c_uint32, C.c_double, C.c_int, _dp, _u8p, _u32p])
PROBLEM_API["ransac_align"] = _RANSAC_SIG
_PRODUCT_ONLY = {
    "ransac_triples": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, _u32p]),
    "evaluate_phong": (C.c_int, [_h, _dp, _dp, _dp, _dp, _dp, _dp]),
    "time_phong": (C.c_int, [_h, C.c_int, _dp]),
    "covariance_block": (C.c_int, [_h, C.c_uint32, _dp]),
    "upload": (C.c_int, [_h]),
    "lm_begin": (C.c_int, [_h]),
    "lm_iterate": (C.c_int, [_h, C.c_int, C.c_int, C.POINTER(Summary)]),
    "download": (C.c_int, [_h]),
    "reset_state": (C.c_int, [_h]),
    "solve_batch": (C.c_int, [C.POINTER(_h), C.c_int, C.POINTER(Summary)]),
    "get_reduced_sizes": (C.c_int, [_h, _ip, _ip]),
    "get_reduced_system": (C.c_int, [_h, _ip, _ip, _dp, _dp, _ip]),
    "get_profile": (C.c_int, [_h, C.POINTER(Profile)]),
    "reset_profile": (C.c_int, [_h]),
    "set_stream": (C.c_int, [_h, C.c_void_p]),
    "time_resjac": (C.c_int, [_h, C.c_int, _dp]),
    "time_schur": (C.c_int, [_h, C.c_int, _dp]),
    "measure_fp64_peak": (C.c_int, [C.c_int, _dp]),
    "analyze": (C.c_int, [_h, C.c_int, C.c_int, C.POINTER(StructureInfo)]),
    "get_launch_count": (C.c_int, [C.POINTER(C.c_uint64)]),
    "comm_unique_id": (C.c_int, [_u8p]),
    "attach_comm": (C.c_int, [_h, C.c_int, C.c_int, _u8p]),
}
PRODUCT_SYMBOLS = sorted("cslam_" + n for n in list(PROBLEM_API) + list(_PRODUCT_ONLY))


class Lib:
    """A loaded library plus typed access to its entry points by unprefixed name."""

    def __init__(self, path, prefix, tables):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing — build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'`. There is no CPU fallback for the CUDA back end.")
        self.path = path
        self.prefix = prefix
        self.dll = C.CDLL(path)
        for table in tables:
            for name, (res, args) in table.items():
                fn = getattr(self.dll, prefix + name)
                fn.restype = res
                fn.argtypes = args
                setattr(self, name, fn)


_cache = {}


def load_product():
    if "product" not in _cache:
        _cache["product"] = Lib(PRODUCT_SO, "cslam_", (PROBLEM_API, _PRODUCT_ONLY))
    return _cache["product"]


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def u32ptr(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


def u8ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u8p)
