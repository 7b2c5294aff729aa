"""B200-native batched DDP/iLQR on SO(3)/SE(3): drop-in for traoptlibrary's solver path.

`traoptlibrary` (sub-package) mirrors the reference's Dynamics / Cost / Constraint / Controller
classes; `BatchSolver` is the handle of the native CUDA library they drive.  Importing this package
requires the built CUDA library (see build.py); there is no CPU fallback.
"""
from ._lib import TrajoptError, LIB_PATH  # noqa: F401
from .solver import BatchSolver, lie_op, launch_count  # noqa: F401
from .pipeline import PipelinedSolver  # noqa: F401

__all__ = ["BatchSolver", "PipelinedSolver", "TrajoptError", "lie_op", "launch_count", "LIB_PATH"]
