"""Build the CUDA library in-tree with nvcc for sm_100a (B200).

    python -m trajectory_optimization_matrix_lie_groups_b200.build [--force] [-v]

Produces `libtrajopt_b200.so` next to this file.  nvcc cross-compiles without a GPU, so the same
command is the "does it build" check on the CPU box and the real build for the B200 box (the .so
is git-ignored but travels with the repo snapshot).  One translation unit per problem family
(csrc/kind_*.cu) plus the C ABI (csrc/api.cu), compiled in parallel, then linked.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtrajopt_b200.so")
SOURCES = ["api.cu", "kind_so3.cu", "kind_se3.cu", "kind_drone.cu", "kind_rigid.cu", "kind_pend.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "trajopt_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
# A/B builds of kernel experiments: TRAJOPT_NVCC_EXTRA="-DFOO ..." adds flags, TRAJOPT_LIB_SUFFIX=_foo writes
# libtrajopt_b200_foo.so (objects under build_foo/) next to the product library; `TRAJOPT_LIB=<path>` makes _lib.py load it.
EXTRA = os.environ.get("TRAJOPT_NVCC_EXTRA", "").split()
SUFFIX = os.environ.get("TRAJOPT_LIB_SUFFIX", "")
if SUFFIX:
    OBJ = os.path.join(HERE, "build" + SUFFIX)
    LIB = os.path.join(HERE, f"libtrajopt_b200{SUFFIX}.so")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [nvcc] + NVCC_FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return obj


def build(force=False, verbose=False):
    """Compile if the library is missing or older than its sources.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(lambda s: _compile(nvcc, s, verbose), SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
