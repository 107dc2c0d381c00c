"""Build the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libcslam_b200.so")
SOURCES = ["kernels.cu", "kernels_grouped.cu", "kernels_pcg.cu", "kernels_window.cu", "kernels_band.cu", "kernels_dense.cu", "kernels_wband.cu", "kernels_bandpcg.cu", "kernels_phong.cu", "kernels_phong_solve.cu", "kernels_phong_grouped.cu", "kernels_phong_long.cu", "ransac.cu", "structure.cu", "engine.cu", "capi.cu", "comm.cu"]
HEADERS = ["closed_form.h", "common.cuh", "engine.h", "dogleg_host.h", "kernels.cuh", "comm.h", "phong_common.cuh", "chol_chain.cuh",
           os.path.join("..", "..", "include", "cslam_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--fmad=true"]
# diagnostic builds only (e.g. CSLAM_NVCC_EXTRA=-DCSLAM_ODD2_PROF); combine with force=True
NVCC_FLAGS += os.environ.get("CSLAM_NVCC_EXTRA", "").split()


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(bdir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(bdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        # link beside the target and rename: a snapshot of the tree never sees a half-written library
        run([nvcc, "-shared", "-o", OUT + ".tmp"] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
        os.replace(OUT + ".tmp", OUT)
    return OUT


def build_host_driver(name="dataset_vo_b200"):
    """C++ host mirror + a restated driver (dataset_vo_b200, dataset_vo_sun_b200,
    dataset_ba_phong_b200); links the C ABI library."""
    host = os.path.join(HERE, "host")
    out = os.path.join(host, name)
    src = os.path.join(host, name + ".cpp")
    deps = [src, os.path.join(host, "cslam_problem.hpp"), os.path.join(host, "dataset.hpp"),
            os.path.join(host, "sun_dataset.hpp"), OUT,
            os.path.join(os.path.dirname(HERE), "include", "cslam_b200.h")]
    if _stale(out, deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", out, src, "-L" + CSRC, "-lcslam_b200",
                               "-Wl,-rpath," + CSRC, "-Wl,-rpath,$ORIGIN/../csrc"])
    return out


HOST_DRIVERS = ("dataset_vo_b200", "dataset_vo_sun_b200", "dataset_ba_phong_b200", "ba_all_b200")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
