"""GPU parity, level 2: per-stage quantities against the oracle, through the C ABI.

  * `trajopt_debug_linearize`: f_x, f_u, defect, l, l_x, l_u, l_xx of whole trajectories
    (the reference's `_linearization`, traopt_controller.py:2098-2176 / 2823-2910);
  * `trajopt_debug_stage` via the drop-in Dynamics / Cost classes: the reference's per-stage
    callbacks `f`, `f_x`, `f_u`, `l`, `l_x`, `l_xx`, `_err` (traopt_dynamics.py / traopt_cost.py);
  * `trajopt_debug_gains`: k, K of one backward sweep (`_backward_pass`, :2178-2261 / :2912-3006).

Tolerance: 1e-11 relative to the magnitude of each array (FP64; observed <= 3e-14).
"""
import numpy as np
import pytest

import gpu_common as gc
from oracle import models, problems, solvers

pytestmark = pytest.mark.gpu
RTOL = 1e-11


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, float(np.max(np.abs(b)))))


@pytest.mark.parametrize("name,method,horizon", [
    ("se3_n120", "ms", None), ("se3_n120", "ss", None), ("so3_n249", "ms", 40), ("so3_n249", "ss", 40),
    ("drone_n150", "ms", 40), ("drone_n150", "ss", 40), ("rigid_n120", "ms", 40),
    ("pendulum_n80", "ms", 40), ("pendulum_n80", "ss", 40),
])
def test_linearization_matches_oracle(name, method, horizon):
    g = problems.load_golden(name)
    kind = str(g["kind"])
    B = 5
    s, x0, N = gc.make_solver(g, method, B, horizon=horizon, max_iters=3, tol_grad_norm=1e-12)
    X0 = gc.perturbed_x0(x0, B)
    rng = np.random.default_rng(1)
    us = 0.1 * rng.standard_normal((B, N, s.NU))
    s.begin(X0, us)
    out = {k: v.cpu().numpy() for k, v in s.debug_linearize().items()}
    dyn, cost, group, q_ref, xi_ref, _, _ = problems.from_golden(g, horizon)
    for b in (0, 2, B - 1):
        x0o = gc.oracle_state(kind, X0[b])
        if method == "ms":
            xs = [x0o] + [[q_ref[i], np.array(xi_ref[i], dtype=float)] for i in range(1, N + 1)]
        else:
            xs = [x0o]
            for i in range(N):
                xs.append(dyn.f(xs[i], us[b, i], i))
        d, F_x, F_u, L, L_x, L_u, L_xx, L_ux, L_uu = solvers._linearize(dyn, cost, group, xs, us[b], N, method == "ms")
        ref = {"F_x": F_x, "F_u": F_u, "L": L, "L_x": L_x, "L_u": L_u, "L_xx": L_xx}
        if d is not None:
            ref["d"] = d
        for k, v in ref.items():
            assert _rel(out[k][b], v) < RTOL, (k, b)
        assert np.all(L_ux == 0.0)      # the device never stores l_ux: it is identically zero on this path


@pytest.mark.parametrize("name,method", [("se3_n120", "ms"), ("se3_n120", "ss"), ("so3_n249", "ms"), ("drone_n150", "ms"),
                                         ("pendulum_n80", "ms"), ("pendulum_n80", "ss")])
def test_backward_gains_match_oracle(name, method):
    """One backward sweep from the initial trajectory: k, K against the oracle's LU-based solve."""
    horizon = 40
    g = problems.load_golden(name)
    kind = str(g["kind"])
    s, x0, N = gc.make_solver(g, method, 2, horizon=horizon, max_iters=1, tol_grad_norm=1e-30)
    X0 = gc.perturbed_x0(x0, 2)
    s.begin(X0)
    s.iterate(1)
    k_gpu, K_gpu = (t.cpu().numpy() for t in s.debug_gains())
    mu_gpu, _ = s.export_reg()
    dyn, cost, group, q_ref, xi_ref, _, _ = problems.from_golden(g, horizon)
    for b in (0, 1):
        x0o = gc.oracle_state(kind, X0[b])
        us0 = np.zeros((N, dyn.action_size))
        if method == "ms":
            r = solvers.ilqr_ms(dyn, cost, group, N, q_ref, xi_ref, x0o, us0, n_iterations=1, tol_grad_norm=1e-30,
                                n_alphas=13 if kind in ("so3", "pendulum") else 20)
        else:
            r = solvers.ilqr_ss(dyn, cost, group, N, x0o, us0, n_iterations=1, tol_grad_norm=1e-30)
        assert _rel(k_gpu[b], r.k) < 1e-9
        assert _rel(K_gpu[b], r.K) < 1e-9
        assert float(mu_gpu[b]) == r.mu_hist[0]


@pytest.mark.parametrize("name", ["se3_n120", "drone_n150", "so3_n249", "rigid_n120", "pendulum_n80"])
def test_dropin_class_callbacks_match_oracle(name):
    """The mirror classes' f / f_x / f_u / l / l_x / l_xx / _err are the CUDA library's numbers."""
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import traopt_cost, traopt_dynamics
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary.manif_compat import SO3, SO3Tangent
    g = problems.load_golden(name)
    kind = str(g["kind"])
    dyn_o, cost_o, group, q_ref_o, xi_ref, x0_o, N = problems.from_golden(g)
    J, dt, Q, R, P = g["prob_J"], float(g["prob_dt"]), g["prob_Q"], g["prob_R"], g["prob_P"]
    if kind in ("so3", "pendulum"):
        dyn = (traopt_dynamics.SO3Dynamics(J, dt) if kind == "so3" else
               traopt_dynamics.Pendulum3dDyanmics(J, float(g["prob_m"]), float(g["prob_length"]), dt))
        q_ref = [SO3(q) for q in q_ref_o]
        cost = traopt_cost.SO3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, [SO3Tangent(w) for w in xi_ref])
    elif kind == "drone":
        dyn = traopt_dynamics.DroneDynamics(J, dt)
        cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, g["prob_q_ref"], xi_ref, action_size=4)
    elif kind == "rigid":
        dyn = traopt_dynamics.RigidBodyDynamics(J, dt)
        cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, g["prob_q_ref"], xi_ref)
    else:
        dyn = traopt_dynamics.SE3Dynamics(J, dt)
        cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, g["prob_q_ref"], xi_ref)
    rng = np.random.default_rng(5)
    m = dyn.action_size
    assert dyn.state_size == dyn_o.state_size and m == dyn_o.action_size
    for i in (0, 7, N - 1):
        # a state a little off the reference at stage i
        row = gc.oracle_rows(kind, [[q_ref_o[i], np.asarray(xi_ref[i], dtype=float)]])[0]
        row = gc.perturbed_x0(row, 2, seed=i, scale=0.05)[1]
        xo = gc.oracle_state(kind, row)
        x = [SO3(xo[0]), SO3Tangent(xo[1])] if kind in ("so3", "pendulum") else [xo[0], xo[1]]
        u = rng.standard_normal(m)
        fo = dyn_o.f(xo, u, i)
        fg = dyn.f(x, u, i)
        if kind in ("so3", "pendulum"):
            assert gc.quat_rows_close(np.concatenate((fg[0].coeffs(), fg[1].coeffs())), np.concatenate((fo[0], fo[1])), 7) < 1e-13
        else:
            assert np.max(np.abs(fg[0] - fo[0])) < 1e-13 and np.max(np.abs(fg[1] - fo[1])) < 1e-13
        assert _rel(dyn.f_x(x, u, i), dyn_o.f_x(xo, u, i)) < RTOL
        assert _rel(dyn.f_u(x, u, i), dyn_o.f_u(xo, u, i)) < RTOL
        for terminal in (False, True):
            uu = None if terminal else u
            assert abs(cost.l(x, uu, i, terminal=terminal) - cost_o.l(xo, uu, i, terminal=terminal)) < RTOL * max(1.0, abs(cost_o.l(xo, uu, i, terminal=terminal)))
            assert _rel(cost.l_x(x, uu, i, terminal=terminal), cost_o.l_x(xo, uu, i, terminal=terminal)) < RTOL
            assert _rel(cost.l_xx(x, uu, i, terminal=terminal), cost_o.l_xx(xo, uu, i, terminal=terminal)) < RTOL
        assert _rel(cost.l_u(x, u, i), cost_o.l_u(xo, u, i)) < RTOL
        assert _rel(cost.l_uu(x, u, i), cost_o.l_uu(xo, u, i)) < RTOL
        e_q, e_v = cost._err(x, i)
        eo = cost_o._err(xo, i)
        assert _rel(np.concatenate((e_q, e_v)), np.concatenate([np.ravel(a) for a in eo])) < RTOL


def test_al_cost_terms_match_oracle():
    """ALConstrainedCost on top of the native SE3 cost (traopt_cost.py:1236-1320)."""
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import traopt_constraints, traopt_cost
    g = problems.load_golden("se3_n120")
    dyn_o, cost_o, group, q_ref_o, xi_ref, x0_o, N = problems.from_golden(g)
    cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(g["prob_Q"], g["prob_R"], g["prob_P"], g["prob_q_ref"], xi_ref)
    con = traopt_constraints.InputConstraint(-0.5, 0.5)
    con_o = models.InputConstraint(np.full(6, -0.5), np.full(6, 0.5))
    al = traopt_cost.ALConstrainedCost(cost, con, N)
    al_o = models.ALConstrainedCost(cost_o, con_o, N)
    rng = np.random.default_rng(9)
    lmbd = np.abs(rng.standard_normal((N + 1, 12)))
    Imu = np.stack([np.diag(np.where(rng.random(12) < 0.5, 0.0, 10.0)) for _ in range(N + 1)])
    al.lmbd, al.Imu, al.mu = lmbd, Imu, 10.0
    al_o.lmbd, al_o.Imu, al_o.mu = lmbd, Imu, 10.0
    x = x0_o
    u = rng.standard_normal(6)
    for fn in ("l", "l_x", "l_u", "l_xx", "l_ux", "l_uu"):
        a = getattr(al, fn)(x, u, 3)
        b = getattr(al_o, fn)(x, u, 3)
        assert _rel(a, b) < RTOL, fn
