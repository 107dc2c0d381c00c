"""ceres_slam_b200 — B200-native bundle-adjustment back end for ceres-slam's hot path.

Package layout: `csrc/` holds the sm_100a CUDA kernels and the C ABI (include/cslam_b200.h);
`capi`/`problem` are the ctypes host handle; `synthetic` generates the benchmark tracks;
`host/` holds the C++ mirror of the reference's API surface.
"""
from .problem import BAProblem, CslamError, default_options, solve_batch  # noqa: F401
