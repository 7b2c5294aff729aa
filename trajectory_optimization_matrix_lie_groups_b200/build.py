"""Build the CUDA library in-tree with nvcc for sm_100a (B200).

    python -m trajectory_optimization_matrix_lie_groups_b200.build [--force]

Produces `libtrajopt_b200.so` next to this file.  nvcc cross-compiles without a GPU, so the same
command is the "does it build" check on the CPU box and the real build for the B200 box (the .so
is git-ignored but travels with the repo snapshot).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtrajopt_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "lie.cuh", "model.cuh", "kernels.cuh", "backward.cuh", "backward3.cuh", "kernels_fwd.cuh", "debug.cuh",
           os.path.join("..", "..", "include", "trajopt_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile if the library is missing or older than its sources.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
