"""Helpers shared by the GPU parity tests: build a native BatchSolver from a golden fixture."""
import numpy as np

from oracle import problems


def inertia_parts(kind, J):
    J = np.asarray(J, dtype=float)
    if kind in ("so3", "pendulum"):
        return J, 1.0
    return J[:3, :3], float(J[4, 4])


def make_solver(g, method, B, horizon=None, device="cuda", **params):
    """g: golden dict.  Returns (BatchSolver, x0 row (NS,), N)."""
    from trajectory_optimization_matrix_lie_groups_b200 import BatchSolver, layout
    kind = str(g["kind"])
    q_ref, xi_ref = g["prob_q_ref"], g["prob_xi_ref"]
    if horizon is not None:
        q_ref, xi_ref = q_ref[:horizon + 1], xi_ref[:horizon + 1]
    N = q_ref.shape[0] - 1
    s = BatchSolver(kind, method, N, B, device=device)
    Ib, mass = inertia_parts(kind, g["prob_J"])
    if kind == "pendulum":
        mass, params = float(g["prob_m"]), dict(params, length=float(g["prob_length"]))
    s.set_params(dt=float(g["prob_dt"]), Ib=Ib, mass=mass, Q=g["prob_Q"], R=g["prob_R"], P=g["prob_P"], **params)
    s.set_reference(layout.pose_rows(kind in ("so3", "pendulum"), q_ref), xi_ref)
    x0 = np.concatenate((layout.pose_rows(kind in ("so3", "pendulum"), g["prob_x0_q"]), np.asarray(g["prob_x0_xi"], dtype=float).reshape(-1)))
    return s, x0, N


def perturbed_x0(x0, B, seed=24234156, scale=0.05):
    """Batch of initial states around x0 (row 0 unperturbed): small quaternion + vector noise."""
    rng = np.random.default_rng(seed)
    X = np.tile(x0, (B, 1))
    if B > 1:
        X[1:] += scale * rng.standard_normal((B - 1, x0.size))
        X[:, :4] /= np.linalg.norm(X[:, :4], axis=1, keepdims=True)
    return X


def oracle_state(kind, row):
    """device state row -> oracle state [pose, velocity]."""
    from oracle import lie
    row = np.asarray(row, dtype=float)
    if kind in ("so3", "pendulum"):
        return [row[:4] / np.linalg.norm(row[:4]), row[4:7].copy()]
    q = row[:4] / np.linalg.norm(row[:4])
    return [lie.se3_to_matrix(q, row[4:7]), row[7:13].copy()]


def oracle_rows(kind, xs):
    """oracle states -> (N+1, NS) device rows."""
    from oracle import lie
    out = []
    for x in xs:
        if kind in ("so3", "pendulum"):
            out.append(np.concatenate((x[0], x[1])))
        else:
            q, p = lie.se3_from_matrix(x[0])
            out.append(np.concatenate((q, p, x[1])))
    return np.stack(out)


def quat_rows_close(a, b, NS):
    """max abs difference of state rows with the quaternion sign ambiguity removed."""
    a, b = np.array(a, dtype=float), np.array(b, dtype=float)
    s = np.sign(np.sum(a[..., :4] * b[..., :4], axis=-1, keepdims=True))
    b = b.copy()
    b[..., :4] *= s
    return float(np.max(np.abs(a - b)))
