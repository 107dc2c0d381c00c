"""TEST INFRASTRUCTURE.  CPU restatement of the reference's hot path (the checker) and, under
`_ref/`, the reference's own headers compiled against stand-ins for Eigen and Ceres' Jet.  Only
tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline leg import this package; the product
(`ceres_slam_b200/`, `include/`) never does (tests/test_abi_and_host.py enforces it)."""
