"""Glue between the mirror classes and the native `BatchSolver`."""
import numpy as np

from .. import layout
from ..solver import BatchSolver
from . import manif_compat

KIND_OF_DYNAMICS = {}   # filled by traopt_dynamics (class -> kind string)


def state_row(kind, x):
    """Reference-style state [q, xi] -> device row."""
    q, xi = x
    if kind in ("so3", "pendulum"):
        return np.concatenate((manif_compat.so3_quat(q), manif_compat.so3_vel(xi)))
    return np.concatenate((layout.pose_rows(False, np.asarray(q, dtype=float)), np.asarray(xi, dtype=float).reshape(6)))


def row_state(kind, row):
    """Device row -> reference-style state [q, xi]."""
    row = np.asarray(row, dtype=float)
    if kind in ("so3", "pendulum"):
        return [manif_compat.SO3(row[:4]), manif_compat.SO3Tangent(row[4:7])]
    return [layout.rows_to_se3(row[:7]), row[7:13].copy()]


def rows_states(kind, rows):
    rows = np.asarray(rows, dtype=float)
    if kind in ("so3", "pendulum"):
        return [[manif_compat.SO3(r[:4]), manif_compat.SO3Tangent(r[4:7])] for r in rows]
    T = layout.rows_to_se3(rows[:, :7])
    return [[T[i], rows[i, 7:13].copy()] for i in range(rows.shape[0])]


def ref_rows(kind, q_ref):
    if kind in ("so3", "pendulum"):
        return np.stack([manif_compat.so3_quat(q) for q in q_ref])
    return layout.pose_rows(False, np.asarray(q_ref, dtype=float))


def dynamics_params(kind, dynamics):
    """The dynamics object's share of trajopt_params."""
    so3_like = kind in ("so3", "pendulum")
    return dict(dt=dynamics.dt, Ib=dynamics.J if so3_like else dynamics.Ib,
                mass=getattr(dynamics, "m", 1.0) if so3_like else dynamics.m, gravity=getattr(dynamics, "g", 9.8),
                length=getattr(dynamics, "l", 0.0))


def make_solver(kind, method, N, B, dynamics, cost, q_ref, xi_ref, device=None, bounds=None, **params):
    s = BatchSolver(kind, method, N, B, device=device)
    if bounds is not None:
        params["lb"], params["ub"] = bounds
    s.set_params(Q=cost.Q, R=cost.R, P=cost.P, **dynamics_params(kind, dynamics), **params)
    xi_ref = np.asarray([manif_compat.so3_vel(w) for w in xi_ref]) if kind in ("so3", "pendulum") else np.asarray(xi_ref, dtype=float)
    s.set_reference(ref_rows(kind, q_ref)[:N + 1], xi_ref[:N + 1])
    return s
