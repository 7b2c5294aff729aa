"""Problem builders for the oracle: load a golden fixture / config and return ready-to-run pieces.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import os

import numpy as np

from . import lie, models, solvers

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_golden(name):
    """`rigid_*` = the SE3 fixture of the same name re-labelled as a rigid body under gravity (the reference
    ships no result file for RigidBodyDynamics; such problems are checked against the oracle only)."""
    if name.startswith("rigid_"):
        g = load_golden("se3_" + name[len("rigid_"):])
        g["kind"] = np.array("rigid")
        return g
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def build(kind, J, dt, Q, R, P, q_ref, xi_ref, m=None, length=None):
    """Return (dynamics, cost, group, q_ref in the group's pose type)."""
    if kind == "se3":
        return (models.SE3Dynamics(J, dt), models.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref),
                solvers.SE3Group, [np.asarray(T, dtype=float) for T in q_ref])
    if kind == "rigid":
        return (models.RigidBodyDynamics(J, dt), models.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref),
                solvers.SE3Group, [np.asarray(T, dtype=float) for T in q_ref])
    if kind == "drone":
        return (models.DroneDynamics(J, dt), models.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref),
                solvers.SE3Group, [np.asarray(T, dtype=float) for T in q_ref])
    if kind == "pendulum":
        return (models.Pendulum3dDynamics(J, m, length, dt), models.SO3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref),
                solvers.SO3Group, [lie.rot_to_quat(np.asarray(Rm, dtype=float)) for Rm in q_ref])
    if kind == "so3":
        return (models.SO3Dynamics(J, dt), models.SO3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref),
                solvers.SO3Group, [lie.rot_to_quat(np.asarray(Rm, dtype=float)) for Rm in q_ref])
    raise ValueError(kind)


def from_golden(g, horizon=None):
    kind = str(g["kind"])
    q_ref, xi_ref = g["prob_q_ref"], g["prob_xi_ref"]
    if horizon is not None:
        q_ref, xi_ref = q_ref[:horizon + 1], xi_ref[:horizon + 1]
    dyn, cost, group, q_ref_g = build(kind, g["prob_J"], float(g["prob_dt"]), g["prob_Q"], g["prob_R"],
                                      g["prob_P"], q_ref, xi_ref, m=g.get("prob_m"), length=g.get("prob_length"))
    if kind in ("so3", "pendulum"):
        x0 = models.so3_state(g["prob_x0_q"], g["prob_x0_xi"])
    else:
        x0 = [np.array(g["prob_x0_q"]), np.array(g["prob_x0_xi"])]
    N = q_ref.shape[0] - 1
    return dyn, cost, group, q_ref_g, xi_ref, x0, N


def poses_to_matrices(kind, xs):
    """xs (oracle states) -> (poses as (N+1,4,4)|(N+1,3,3), velocities (N+1,6|3))."""
    if kind in ("so3", "pendulum"):
        return np.stack([lie.quat_to_rot(x[0]) for x in xs]), np.stack([x[1] for x in xs])
    return np.stack([x[0] for x in xs]), np.stack([x[1] for x in xs])
