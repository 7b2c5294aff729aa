"""ctypes binding of the C ABI declared in include/trajopt_b200.h.

There is no CPU fallback: if the CUDA library is missing this module raises at import, and every
non-zero return code of the library becomes a `TrajoptError`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TRAJOPT_LIB: an A/B build of the same library (see build.py), for kernel experiments only
LIB_PATH = os.environ.get("TRAJOPT_LIB") or os.path.join(HERE, "libtrajopt_b200.so")

SO3, SE3, DRONE, RIGID, PEND = 0, 1, 2, 3, 4
SS, MS, AL_MS = 0, 1, 2
STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_NO_DESCENT, STATUS_RUNNING = 0, 1, 2, 3
FLAG_REG_EXCEEDED, FLAG_NONFINITE = 16, 32

# trajopt_debug_lie op codes (csrc/debug.cuh): name -> (code, input width, output width)
LIE_OPS = {
    "so3_exp": (0, 3, 4), "so3_log": (1, 4, 3), "so3_jr": (2, 3, 9), "so3_jr_inv": (3, 3, 9),
    "so3_jl": (4, 3, 9), "so3_jl_inv": (5, 3, 9), "se3_exp": (6, 6, 7), "se3_log": (7, 7, 6),
    "se3_Q": (8, 6, 9), "se3_jr": (9, 6, 36), "se3_jr_inv": (10, 6, 36), "se3_adj": (11, 7, 36),
    "se3_compose": (12, 14, 7), "se3_rminus": (13, 14, 6), "se3_lminus": (14, 14, 6),
}


class TrajoptError(RuntimeError):
    pass


class Params(C.Structure):
    """struct trajopt_params (include/trajopt_b200.h)."""
    _fields_ = [
        ("dt", C.c_double),
        ("Ib", C.c_double * 9),
        ("mass", C.c_double),
        ("gravity", C.c_double),
        ("Q", C.c_double * 144),
        ("P", C.c_double * 144),
        ("R", C.c_double * 36),
        ("lb", C.c_double * 6),
        ("ub", C.c_double * 6),
        ("has_constraints", C.c_int32),
        ("rollout_linear", C.c_int32),
        ("line_search", C.c_int32),
        ("n_alphas", C.c_int32),
        ("max_iters", C.c_int32),
        ("tol_grad_norm", C.c_double),
        ("tol_d_norm", C.c_double),
        ("max_reg", C.c_double),
        ("defect_kappa", C.c_double),
        ("n_al_iters", C.c_int32),
        ("al_mu0", C.c_double),
        ("al_mu_scale", C.c_double),
        ("al_mu_max", C.c_double),
        ("tol_constr", C.c_double),
        ("length", C.c_double),
        ("xi_lb", C.c_double * 6),
        ("xi_ub", C.c_double * 6),
        ("has_state_bounds", C.c_int32),
    ]


_P = C.c_void_p
_I = C.c_int
# every symbol include/trajopt_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "trajopt_last_error": (C.c_char_p, []),
    "trajopt_version": (_I, []),
    "trajopt_create": (_I, [_I, _I, _I, _I, _I, C.POINTER(_P)]),
    "trajopt_destroy": (_I, [_P]),
    "trajopt_set_params": (_I, [_P, C.POINTER(Params)]),
    "trajopt_set_reference": (_I, [_P, _P, _P]),
    "trajopt_set_reference_batch": (_I, [_P, _P, _P, _P]),
    "trajopt_set_reference_long": (_I, [_P, _P, _P, C.c_int64]),
    "trajopt_set_reference_offset": (_I, [_P, C.c_int64]),
    "trajopt_set_horizons": (_I, [_P, _P, _P]),
    "trajopt_solve_stream": (_I, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trajopt_solve_stream_host": (_I, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trajopt_begin": (_I, [_P, _P, _P, _I, _P]),
    "trajopt_iterate": (_I, [_P, _I, C.POINTER(_I), _P]),
    "trajopt_iterate_inner": (_I, [_P, _I, C.POINTER(_I), _P]),
    "trajopt_export": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trajopt_export_hist": (_I, [_P, _P, _P, _P, _P, _P]),
    "trajopt_export_al": (_I, [_P, _P, _P, _P, _P, _P, _P]),
    "trajopt_export_reg": (_I, [_P, _P, _P, _P]),
    "trajopt_export_al_state": (_I, [_P, _P, _P, _P]),
    "trajopt_solve": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trajopt_solve_host": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trajopt_solve_host_begin": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_I)]),
    "trajopt_solve_host_wait": (_I, [_P, _I]),
    "trajopt_debug_linearize": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trajopt_debug_gains": (_I, [_P, _P, _P, _P]),
    "trajopt_debug_linesearch_rows": (_I, [_P]),
    "trajopt_debug_linesearch": (_I, [_P, _P, _P]),
    "trajopt_debug_stage": (_I, [_P, _I, _I, _I] + [_P] * 11),
    "trajopt_debug_lie": (_I, [_I, _I, _P, _P, _P]),
    "trajopt_debug_fp64_peak": (_I, [C.c_double, C.POINTER(C.c_double), _P]),
    "trajopt_launch_count": (C.c_int64, [_I]),
    "trajopt_phase_times": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), _I]),
    "trajopt_set_profiling": (_I, [_P, _I]),
    "trajopt_set_compaction": (_I, [_P, _I, _I]),
    "trajopt_set_sweep": (_I, [_P, _I, _I]),
    "trajopt_set_line_search_batch": (_I, [_P, _I]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m trajectory_optimization_matrix_lie_groups_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise TrajoptError(f"trajopt error {rc}: {lib.trajopt_last_error().decode()}")
