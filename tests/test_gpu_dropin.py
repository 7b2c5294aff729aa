"""GPU parity, level 5: the drop-in `traoptlibrary` class API (the reference's boundary, SURVEY.md
section 8b) used the way the reference's main_* / benchmark_* scripts use it — construct Dynamics,
Cost, Controller, call `fit` with the scripts' own `on_iteration` callback — against the reference's
shipped results, plus the new `fit_batch`.
"""
import warnings

import numpy as np
import pytest

from oracle import problems

pytestmark = pytest.mark.gpu


def _mk(name, method, **ctrl_kw):
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import (traopt_controller as tc, traopt_cost,
                                                                              traopt_dynamics)
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary.manif_compat import SO3, SO3Tangent
    g = problems.load_golden(name)
    kind = str(g["kind"])
    J, dt, Q, R, P = g["prob_J"], float(g["prob_dt"]), g["prob_Q"], g["prob_R"], g["prob_P"]
    q_ref, xi_ref = g["prob_q_ref"], g["prob_xi_ref"]
    N = q_ref.shape[0] - 1
    if kind == "so3":
        dyn = traopt_dynamics.SO3Dynamics(J, dt)
        q_ref = [SO3.from_matrix(Rm) for Rm in q_ref]
        xi_ref = [SO3Tangent(w) for w in xi_ref]
        cost = traopt_cost.SO3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref)
        x0 = [SO3.from_matrix(g["prob_x0_q"]), SO3Tangent(g["prob_x0_xi"])]
        cls = tc.iLQR_Tracking_SO3_MS if method == "ms" else tc.iLQR_Tracking_SO3
    else:
        if kind == "drone":
            dyn = traopt_dynamics.DroneDynamics(J, dt)
            cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref, action_size=4)
        else:
            dyn = traopt_dynamics.SE3Dynamics(J, dt)
            cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref)
        x0 = [np.array(g["prob_x0_q"]), np.array(g["prob_x0_xi"])]
        cls = tc.iLQR_Tracking_SE3_MS if method == "ms" else tc.iLQR_Tracking_SE3
    if method == "ms":
        ctrl = cls(dyn, cost, N, q_ref, xi_ref, line_search=False, rollout="nonlinear", **ctrl_kw)
    else:
        ctrl = cls(dyn, cost, N, rollout="nonlinear", **ctrl_kw)
    return g, kind, dyn, cost, ctrl, x0, N


def _pose_mats(kind, xs):
    if kind == "so3":
        return np.stack([x[0].rotation() for x in xs]), np.stack([x[1].coeffs() for x in xs])
    return np.stack([x[0] for x in xs]), np.stack([x[1] for x in xs])


# the callbacks of the reference scripts (benchmark_SE3_tracking.py:22-42)
def _cb_ss(iteration, xs, us, J_opt, accepted, converged, grad_wrt_input_norm, alpha, mu, J_hist, xs_hist, us_hist):
    J_hist.append(J_opt)
    xs_hist.append(list(xs))
    us_hist.append(us.copy())


def _cb_ms(iteration, xs, us, J_opt, accepted, converged, defect_norm, grad_wrt_input_norm, alpha, mu, J_hist, xs_hist,
           us_hist, grad_hist, defect_hist):
    J_hist.append(J_opt)
    xs_hist.append(list(xs))
    us_hist.append(us.copy())
    grad_hist.append(grad_wrt_input_norm)
    defect_hist.append(defect_norm)


@pytest.mark.parametrize("name", ["se3_n120", "so3_n249", "drone_n150"])
def test_ms_fit_with_script_callback(name):
    g, kind, dyn, cost, ctrl, x0, N = _mk(name, "ms")
    us_init = np.zeros((N, dyn.action_size))
    xs, us, J_hist, xs_hist, us_hist, grad_hist, defect_hist = ctrl.fit(
        x0, us_init, n_iterations=200, tol_grad_norm=1e-12, on_iteration=_cb_ms)
    Jg = g["ms_J_hist"]
    assert len(J_hist) == len(Jg)
    assert np.max(np.abs(np.array(J_hist) - Jg) / np.abs(Jg)) < 1e-9
    assert len(grad_hist) == len(g["ms_grad_hist"]) and len(defect_hist) == len(g["ms_defect_hist"])
    assert len(xs_hist) == len(Jg) + 1 and len(us_hist) == len(Jg) + 1
    assert abs(defect_hist[0] - g["ms_defect_hist"][0]) < 1e-9 * g["ms_defect_hist"][0]
    Pm, V = _pose_mats(kind, xs)
    assert np.max(np.abs(us - g["ms_us"])) < 1e-7
    assert np.max(np.abs(Pm - g["ms_xs_q"])) < 1e-7 and np.max(np.abs(V - g["ms_xs_xi"])) < 1e-7
    assert np.all(us_init == 0.0)                      # inputs are never mutated
    # without a callback: same solution, histories stay with the caller (empty), like the reference
    xs2, us2, J2, *_ = ctrl.fit(x0, us_init, n_iterations=200, tol_grad_norm=1e-12)
    assert J2 == [] and np.array_equal(us2, us)


@pytest.mark.parametrize("name", ["se3_n120", "drone_n150"])
def test_ss_fit_with_script_callback(name):
    g, kind, dyn, cost, ctrl, x0, N = _mk(name, "ss")
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        xs, us, J_hist, xs_hist, us_hist, grad_hist = ctrl.fit(
            x0, np.zeros((N, dyn.action_size)), n_iterations=200, tol_grad_norm=1e-12, on_iteration=_cb_ss)
    Jg = g["ss_J_hist"]
    assert len(J_hist) == len(Jg)
    assert np.max(np.abs(np.array(J_hist) - Jg) / np.abs(Jg)) < 1e-9
    assert len(grad_hist) == len(g["ss_grad_hist"])
    assert np.max(np.abs(np.array(grad_hist) - g["ss_grad_hist"]) / g["ss_grad_hist"]) < 1e-6
    assert any("descent direction" in str(w.message) for w in wlist)      # both goldens end on the warning
    Pm, V = _pose_mats(kind, xs)
    assert np.max(np.abs(us - g["ss_us"])) < 1e-7
    assert np.max(np.abs(Pm - g["ss_xs_q"])) < 1e-7 and np.max(np.abs(V - g["ss_xs_xi"])) < 1e-7


def test_fit_batch_matches_fit():
    g, kind, dyn, cost, ctrl, x0, N = _mk("se3_n120", "ms")
    rng = np.random.default_rng(0)
    x0s = [x0]
    for _ in range(4):
        T = x0[0].copy()
        T[:3, 3] += 0.05 * rng.standard_normal(3)
        x0s.append([T, x0[1] + 0.02 * rng.standard_normal(6)])
    res = ctrl.fit_batch(x0s, n_iterations=200, tol_grad_norm=1e-12, return_hist=True)
    assert res.J.shape == (5,) and np.all(res.converged)
    for b in (0, 3):
        xs, us, *_ = ctrl.fit(x0s[b], np.zeros((N, 6)), n_iterations=200, tol_grad_norm=1e-12)
        assert np.array_equal(us, res.us[b])
        assert np.array_equal(np.stack([x[0] for x in xs]), np.stack([x[0] for x in res.states(b)]))
    assert res.iters[0] == len(g["ms_J_hist"])


def test_al_controller_runs_and_respects_bounds():
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import (traopt_constraints, traopt_controller as tc,
                                                                              traopt_cost, traopt_dynamics)
    N, dt = 40, 0.01
    q_ref, xi_ref = workloads.helix_reference(N, dt)
    J = np.diag([0.5, 0.7, 0.9, 1.0, 1.0, 1.0])
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    dyn = traopt_dynamics.SE3Dynamics(J, dt)
    cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, np.zeros((6, 6)), 10 * Q, q_ref, xi_ref)
    con = traopt_constraints.InputConstraint(-8.0, 8.0)
    ctrl = tc.AL_iLQR_Tracking_SE3_MS(dyn, cost, con, N, q_ref, xi_ref)
    T0 = np.eye(4)
    T0[:3, 3] = [-0.3, -0.3, -0.1]
    x0 = [T0, np.array([0, 0, 0.1, 2.0, 0, 0.2])]
    seen = []

    def cb(iteration, constr_converged, lmbd, Imu, mu, constr_eval, lmbd_hist, mu_hist, violation_hist, nactive_hist):
        seen.append((iteration, constr_converged, mu, float(np.max(constr_eval))))
        mu_hist.append(mu)
        violation_hist.append(float(np.max(constr_eval)))

    out = ctrl.fit(x0, np.zeros((N, 6)), n_al_iters=15, n_ilqr_iters=60, on_iteration_al=cb)
    xs, us = out[0], out[1]
    assert len(out) == 10
    assert seen[-1][1] and np.max(np.abs(us)) < 8.0 + 1e-2
    assert [s[2] for s in seen[:3]] == [1e-2, 1e-1, 1.0]      # penalty schedule mu0 = 1e-2, x10 per outer iteration
    assert len(seen) == 10                                     # oracle: 10 outer iterations for this problem


def test_fit_batch_with_per_problem_references():
    g, kind, dyn, cost, ctrl, x0, N = _mk("se3_n120", "ms")
    import copy
    B, Nh = 3, 40
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import traopt_controller as tc, traopt_cost
    q_full, xi_full = g["prob_q_ref"], g["prob_xi_ref"]
    qb = np.stack([q_full[7 * b:7 * b + Nh + 1] for b in range(B)])
    xb = np.stack([xi_full[7 * b:7 * b + Nh + 1] for b in range(B)])
    cost_h = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(g["prob_Q"], g["prob_R"], g["prob_P"], qb[0], xb[0])
    ctrl_h = tc.iLQR_Tracking_SE3_MS(dyn, cost_h, Nh, qb[0], xb[0], rollout="nonlinear")
    res = ctrl_h.fit_batch([x0] * B, n_iterations=60, tol_grad_norm=1e-10, q_ref_batch=qb, xi_ref_batch=xb)
    assert np.all(res.converged)
    for b in range(B):      # the same problem through the single-problem API with that reference
        cost_b = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(g["prob_Q"], g["prob_R"], g["prob_P"], qb[b], xb[b])
        ctrl_b = tc.iLQR_Tracking_SE3_MS(dyn, cost_b, Nh, qb[b], xb[b], rollout="nonlinear")
        xs, us, *_ = ctrl_b.fit(x0, np.zeros((Nh, 6)), n_iterations=60, tol_grad_norm=1e-10)
        assert np.max(np.abs(us - res.us[b])) < 1e-9


def test_al_controller_with_velocity_bounds():
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import (traopt_constraints, traopt_controller as tc,
                                                                              traopt_cost, traopt_dynamics)
    N, dt = 40, 0.01
    q_ref, xi_ref = workloads.helix_reference(N, dt)
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    dyn = traopt_dynamics.SE3Dynamics(np.diag([0.5, 0.7, 0.9, 1.0, 1.0, 1.0]), dt)
    cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, 1e-3 * np.eye(6), 10 * Q, q_ref, xi_ref)
    con = traopt_constraints.InputVelocityConstraint(-30.0, 30.0, [-5, -5, -0.5, -5, -5, -5], [5, 5, 0.5, 5, 5, 5])
    ctrl = tc.AL_iLQR_Tracking_SE3_MS(dyn, cost, con, N, q_ref, xi_ref)
    T0 = np.eye(4)
    T0[:3, 3] = [-0.3, -0.3, -0.1]
    seen = []
    out = ctrl.fit([T0, np.array([0, 0, 0.1, 2.0, 0, 0.2])], np.zeros((N, 6)), n_al_iters=12, n_ilqr_iters=60,
                   on_iteration_al=lambda it, conv, lmbd, Imu, mu, ce, *h: seen.append((conv, lmbd.shape, Imu.shape, ce.shape, float(ce.max()))))
    xs = out[0]
    assert seen[-1][0] and len(seen) == 6                       # oracle: 6 outer iterations
    assert seen[0][1:4] == ((N + 1, 24), (N + 1, 24, 24), (N + 1, 24))
    assert max(abs(x[1][2]) for x in xs) < 0.5 + 1e-2          # omega_z respects its bound


def test_al_inner_callback_is_called_per_inner_iteration():
    """`on_iteration_ilqr` (main_SE3ddp_tracking_exact_al_ms.py:14-16, traopt_controller.py:3236-3240): the inner fit of
    every outer iteration calls it once per accepted iteration with the multiple-shooting callback's arguments, and the
    returned histories are those of the LAST inner fit.  Stepping the device one inner iteration at a time changes no
    number."""
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import (traopt_constraints, traopt_controller as tc,
                                                                              traopt_cost, traopt_dynamics)
    N, dt = 40, 0.01
    q_ref, xi_ref = workloads.helix_reference(N, dt)
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    dyn = traopt_dynamics.SE3Dynamics(np.diag([0.5, 0.7, 0.9, 1.0, 1.0, 1.0]), dt)
    cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, np.zeros((6, 6)), 10 * Q, q_ref, xi_ref)
    con = traopt_constraints.InputConstraint(-8.0, 8.0)
    T0 = np.eye(4)
    T0[:3, 3] = [-0.3, -0.3, -0.1]
    x0 = [T0, np.array([0, 0, 0.1, 2.0, 0, 0.2])]
    calls, outer = [], []

    def cb_inner(iteration, xs, us, J_opt, accepted, converged, defect_norm, grad_norm, alpha, mu, J_hist, xs_hist, us_hist,
                 grad_hist, defect_hist):
        calls.append((len(outer), iteration, J_opt, accepted))
        _cb_ms(iteration, xs, us, J_opt, accepted, converged, defect_norm, grad_norm, alpha, mu, J_hist, xs_hist, us_hist,
               grad_hist, defect_hist)

    def cb_outer(iteration, constr_converged, lmbd, Imu, mu, constr_eval, lmbd_hist, mu_hist, violation_hist, nactive_hist):
        outer.append((iteration, constr_converged, float(np.max(constr_eval))))
        violation_hist.append(float(np.max(constr_eval)))

    ctrl = tc.AL_iLQR_Tracking_SE3_MS(dyn, cost, con, N, q_ref, xi_ref)
    a = ctrl.fit(x0, np.zeros((N, 6)), n_al_iters=15, n_ilqr_iters=60, on_iteration_al=cb_outer, on_iteration_ilqr=cb_inner)
    ctrl2 = tc.AL_iLQR_Tracking_SE3_MS(dyn, cost, con, N, q_ref, xi_ref)
    b = ctrl2.fit(x0, np.zeros((N, 6)), n_al_iters=15, n_ilqr_iters=60)
    assert np.array_equal(a[1], b[1]) and len(outer) == 10 and outer[-1][1]
    per_outer = np.bincount([c[0] for c in calls], minlength=len(outer))
    assert np.all(per_outer >= 1)
    assert all(c[3] for c in calls)
    # histories of the last inner fit, filled by the callback; they equal the device's own (returned when no callback is given)
    assert len(a[2]) == per_outer[-1] == len(b[2]) and np.allclose(a[2], b[2], rtol=0, atol=0)
    assert len(a[3]) == per_outer[-1] + 1 and len(a[4]) == per_outer[-1] + 1      # initial guess + one entry per iteration
    assert [c[1] for c in calls if c[0] == len(outer) - 1] == list(range(per_outer[-1]))
