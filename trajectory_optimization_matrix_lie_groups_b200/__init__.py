"""B200-native batched DDP/iLQR on SO(3)/SE(3): drop-in for traoptlibrary's solver path.

`traoptlibrary` (sub-package) mirrors the reference's Dynamics / Cost / Constraint / Controller
classes; `BatchSolver` is the handle of the native CUDA library they drive.  Importing this package
requires the built CUDA library (see build.py); there is no CPU fallback.

The library is mapped on first use of anything that computes (`BatchSolver`, `PipelinedSolver`, `lie_op`,
`launch_count`, the `traoptlibrary` classes).  The data-only modules — `workloads` (problem definitions),
`layout`, `io`, `distributed` — import without it, so that bench.py's CPU reference arm, which needs the
workload definitions and nothing else, does not load product code.
"""
import os as _os

LIB_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "libtrajopt_b200.so")
if not _os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m trajectory_optimization_matrix_lie_groups_b200.build` "
        "(nvcc, sm_100a).  This package has no CPU fallback.")

_LAZY = {"TrajoptError": "_lib", "BatchSolver": "solver", "lie_op": "solver", "launch_count": "solver",
         "PipelinedSolver": "pipeline"}

__all__ = ["BatchSolver", "PipelinedSolver", "TrajoptError", "lie_op", "launch_count", "LIB_PATH"]


def __getattr__(name):
    if name in _LAZY:
        import importlib
        value = getattr(importlib.import_module("." + _LAZY[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
