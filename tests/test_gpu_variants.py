"""GPU parity, level 4: solver variants the reference ships no result files for, against the oracle
on the same seeded inputs (short horizons, so the NumPy oracle finishes in seconds):

  * batches of perturbed initial states (every problem of the batch, not just problem 0);
  * single shooting with `rollout='linear'` (the SE3 class default, traopt_controller.py:1837);
  * multiple shooting with `rollout='linear'` (:2720-2726);
  * multiple shooting with the merit line search `line_search=True` (:2549-2590);
  * the regularisation-increase branch of the backward pass (non-PD Q_uu, :2233-2246);
  * the augmented-Lagrangian outer loop with input bounds (:3218-3290);
  * non-zero / per-problem initial controls, horizon edge cases (N = 1, 2), ragged batch sizes.

Bar: identical iteration counts, stopping reasons and accepted step indices; J_hist <= 1e-9 relative;
controls / states <= 1e-7.
"""
import warnings

import numpy as np
import pytest

import gpu_common as gc
from oracle import models, problems, solvers

pytestmark = pytest.mark.gpu


def _oracle(g, kind, method, horizon, x0_row, us0, n_iter, **kw):
    dyn, cost, group, q_ref, xi_ref, _, N = problems.from_golden(g, horizon)
    x0 = gc.oracle_state(kind, x0_row)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if method == "ms":
            return solvers.ilqr_ms(dyn, cost, group, N, q_ref, xi_ref, x0, us0, n_iterations=n_iter,
                                   n_alphas=13 if kind in ("so3", "pendulum") else 20,
                                   defect_kappa=1e-14 if kind in ("so3", "pendulum") else 1e-12, **kw)
        return solvers.ilqr_ss(dyn, cost, group, N, x0, us0, n_iterations=n_iter, **kw)


def _decision_window(J):
    """Number of leading iterations whose accept/reject decision is above rounding noise: once the
    oracle's own cost changes by less than 1e-11 relative, `J_new < J_opt` (or the merit test)
    compares numbers that differ in their last bits and the outcome depends on the rounding of
    every operation before it — in the reference as much as here."""
    J = np.asarray(J, dtype=float)
    for i in range(1, len(J)):
        if abs(J[i] - J[i - 1]) < 1e-11 * abs(J[i]):
            return i
    return len(J)


def _compare(kind, out, hist, b, r, check_alpha=True):
    it = int(out["iters"][b])
    Jo = np.array(r.J_hist)
    n = _decision_window(Jo)
    if n == len(Jo):
        assert it == r.iterations, (b, it, r.iterations)
        assert (int(out["status"][b]) & 15) == r.status, (b, int(out["status"][b]), r.status)
        traj_tol = 1e-7
    else:       # the solve ran into the rounding-noise tail: decisions are compared up to it
        assert it >= n, (b, it, n)
        traj_tol = 1e-6
    if n:
        rel = np.abs(hist["J_hist"][b, :n] - Jo[:n]) / np.maximum(np.abs(Jo[:n]), 1e-300)
        assert rel.max() < 1e-9, (b, rel.max())
        assert abs(hist["J_hist"][b, it - 1] - Jo[-1]) < 1e-9 * abs(Jo[-1])
    if check_alpha:
        assert hist["alpha_hist"][b, :n].tolist() == r.alpha_hist[:n], b
    assert np.max(np.abs(out["us"][b].cpu().numpy() - r.us)) < traj_tol
    assert gc.quat_rows_close(out["xs"][b].cpu().numpy(), gc.oracle_rows(kind, r.xs), 0) < traj_tol


@pytest.mark.parametrize("name,method,horizon,n_iter,kw", [
    ("se3_n120", "ms", 30, 6, {}),
    ("se3_n120", "ss", 30, 6, {}),
    ("se3_n120", "ss", 30, 6, {"rollout": "linear"}),
    ("se3_n120", "ms", 30, 6, {"rollout": "linear"}),
    ("se3_n120", "ms", 30, 8, {"line_search": True}),
    ("se3_n120", "ms", 30, 8, {"line_search": True, "rollout": "linear"}),
    ("so3_n249", "ms", 30, 6, {}),
    ("so3_n249", "ss", 30, 6, {}),
    ("so3_n249", "ss", 30, 6, {"rollout": "linear"}),
    ("so3_n249", "ms", 30, 8, {"line_search": True}),
    ("drone_n150", "ms", 30, 6, {}),
    ("drone_n150", "ss", 30, 6, {}),
    ("drone_n150", "ms", 30, 8, {"line_search": True}),
    ("pendulum_n80", "ms", 30, 6, {}),
    ("pendulum_n80", "ss", 30, 6, {}),
    ("pendulum_n80", "ss", 30, 6, {"rollout": "linear"}),
    ("pendulum_n80", "ms", 30, 6, {"rollout": "linear"}),
    ("pendulum_n80", "ms", 30, 8, {"line_search": True}),
    ("rigid_n120", "ms", 30, 6, {}),
    ("rigid_n120", "ss", 12, 6, {}),      # (free fall over 30 stages of 0.05 s runs into the give-up branch)
    ("rigid_n120", "ms", 30, 6, {"rollout": "linear"}),
], ids=lambda v: str(v).replace(" ", "") if not isinstance(v, dict) else ",".join(f"{k}={x}" for k, x in v.items()) or "default")
def test_batch_against_oracle(name, method, horizon, n_iter, kw):
    g = problems.load_golden(name)
    kind = str(g["kind"])
    B = 6
    s, x0, N = gc.make_solver(g, method, B, horizon=horizon, max_iters=n_iter, tol_grad_norm=1e-12, **kw)
    X0 = gc.perturbed_x0(x0, B, scale=0.05)
    rng = np.random.default_rng(11)
    us0 = 0.05 * rng.standard_normal((B, N, s.NU))
    us0[0] = 0.0
    out = s.solve(X0, us0)
    hist = {k: v.cpu().numpy() for k, v in s.export_hist().items()}
    for b in range(B):
        r = _oracle(g, kind, method, horizon, X0[b], us0[b], n_iter, tol_grad_norm=1e-12, **kw)
        _compare(kind, out, hist, b, r)


@pytest.mark.parametrize("method", ["ms", "ss"])
def test_regularisation_increase_branch(method, monkeypatch):
    """A negative terminal velocity weight makes Q_uu non-PD at mu = 1 in the last stages: the
    Levenberg-Marquardt state machine (delta, mu persisting across stages and iterations, 12 failed
    factorisations per solve here) must take the same path as the reference."""
    g = dict(problems.load_golden("se3_n120"))
    P = g["prob_P"].copy()
    P[6:, 6:] = -0.5 * np.eye(6)
    g["prob_P"] = P
    kind, horizon, n_iter, B = "se3", 25, 3, 4
    s, x0, N = gc.make_solver(g, method, B, horizon=horizon, max_iters=n_iter, tol_grad_norm=1e-12)
    X0 = gc.perturbed_x0(x0, B, scale=0.02)
    out = s.solve(X0)
    hist = {k: v.cpu().numpy() for k, v in s.export_hist().items()}
    mu_gpu, delta_gpu = (t.cpu().numpy() for t in s.export_reg())
    fails = []
    orig = solvers.is_pos_def
    monkeypatch.setattr(solvers, "is_pos_def", lambda A: (fails.append(orig(A)), fails[-1])[1])
    for b in range(B):
        del fails[:]
        r = _oracle(g, kind, method, horizon, X0[b], np.zeros((N, 6)), n_iter, tol_grad_norm=1e-12)
        assert fails.count(False) >= 6 and not r.reg_exceeded, "the test problem no longer exercises the regularisation branch"
        _compare(kind, out, hist, b, r)
        assert mu_gpu[b] == r.mu_hist[-1]


def test_max_regularisation_gives_up():
    """mu >= max_reg: the reference warns "exceeded max regularization term" (:2238-2240); the
    device stops the problem and raises TRAJOPT_FLAG_REG_EXCEEDED."""
    g = dict(problems.load_golden("se3_n120"))
    g["prob_R"] = -1e6 * np.eye(6)
    s, x0, N = gc.make_solver(g, "ms", 2, horizon=10, max_iters=3, tol_grad_norm=1e-12, max_reg=1e3)
    out = s.solve(gc.perturbed_x0(x0, 2))
    st = out["status"].cpu().numpy()
    assert np.all(st & 16) and np.all((st & 15) == 2)


def test_augmented_lagrangian_against_oracle():
    """AL_iLQR_Tracking_SE3_MS (:3139-3293) on a helix reference with tight input bounds."""
    from trajectory_optimization_matrix_lie_groups_b200 import BatchSolver, layout, workloads
    N, dt, B = 40, 0.01, 3
    q_ref, xi_ref = workloads.helix_reference(N, dt)
    J = np.diag([0.5, 0.7, 0.9, 1.0, 1.0, 1.0])
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    R = np.zeros((6, 6))
    P = 10 * Q
    lb, ub = -8.0, 8.0
    rng = np.random.default_rng(24234156)
    X0 = workloads.perturb_se3(np.eye(3), np.array([-0.3, -0.3, -0.1]), np.array([0, 0, 0.1, 2.0, 0, 0.2]), B, 0.02, rng)
    n_al, n_ilqr = 15, 60
    s = BatchSolver("se3", "al_ms", N, B)
    s.set_params(dt=dt, Ib=J[:3, :3], mass=1.0, Q=Q, R=R, P=P, max_iters=n_ilqr, tol_grad_norm=1e-6, tol_d_norm=1e-6,
                 lb=lb, ub=ub, n_al_iters=n_al, tol_constr=1e-2)
    s.set_reference(layout.pose_rows(False, q_ref), xi_ref)
    out = s.solve(X0)
    al = {k: v.cpu().numpy() for k, v in s.export_al().items()}
    us = out["us"].cpu().numpy()
    dyn = models.SE3Dynamics(J, dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref)
    con = models.InputConstraint(np.full(6, lb), np.full(6, ub))
    any_active = False
    for b in range(B):
        alc = models.ALConstrainedCost(cost, con, N)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = solvers.al_ilqr_ms(dyn, alc, con, solvers.SE3Group, N, [T for T in q_ref], xi_ref,
                                   gc.oracle_state("se3", X0[b]), np.zeros((N, 6)), n_al_iters=n_al, n_ilqr_iters=n_ilqr,
                                   tol_constr=1e-2)
        assert int(al["outer_iters"][b]) == r.outer_iterations, (b, al["outer_iters"][b], r.outer_iterations)
        assert int(out["iters"][b]) == r.inner.iterations
        assert abs(al["violation"][b] - r.violation_hist[-1]) < 1e-7
        assert abs(al["mu"][b] - r.mu) <= 1e-12 * r.mu
        assert np.max(np.abs(us[b] - r.inner.us)) < 1e-7
        assert np.max(np.abs(al["lmbd"][b] - r.lmbd)) < 1e-6 * max(1.0, np.max(np.abs(r.lmbd)))
        assert np.array_equal(al["imu"][b][:N], np.stack([np.diag(M) for M in r.Imu])[:N])
        assert abs(float(out["J"][b]) - r.inner.J_hist[-1]) < 1e-9 * abs(r.inner.J_hist[-1])
        any_active |= r.outer_iterations > 1
        assert r.constr_converged and np.max(np.abs(us[b])) < ub + 1e-2
    assert any_active, "bounds never became active: the test does not exercise the multiplier update"


@pytest.mark.parametrize("N", [1, 2, 3])
@pytest.mark.parametrize("method", ["ms", "ss"])
def test_tiny_horizons(N, method):
    g = problems.load_golden("se3_n120")
    s, x0, n = gc.make_solver(g, method, 2, horizon=N, max_iters=5, tol_grad_norm=1e-12)
    assert n == N
    X0 = gc.perturbed_x0(x0, 2)
    out = s.solve(X0)
    hist = {k: v.cpu().numpy() for k, v in s.export_hist().items()}
    for b in range(2):
        r = _oracle(g, "se3", method, N, X0[b], np.zeros((N, 6)), 5, tol_grad_norm=1e-12)
        _compare("se3", out, hist, b, r)


@pytest.mark.parametrize("B", [1, 31, 33, 65])
def test_ragged_batches_are_position_independent(B):
    """A problem's result must not depend on its slot or on its neighbours: solve the same x0 in
    batches of different (non multiple-of-32) sizes and positions -> bit-identical results."""
    g = problems.load_golden("so3_n249")
    s1, x0, N = gc.make_solver(g, "ms", 1, horizon=60, max_iters=20, tol_grad_norm=1e-10)
    one = s1.solve(x0[None, :])
    s, _, _ = gc.make_solver(g, "ms", B, horizon=60, max_iters=20, tol_grad_norm=1e-10)
    X0 = gc.perturbed_x0(x0, B, scale=0.05)
    X0[B - 1] = x0
    out = s.solve(X0)
    for k in ("J", "iters", "status", "grad", "defect", "xs", "us"):
        assert np.array_equal(out[k][B - 1].cpu().numpy(), one[k][0].cpu().numpy()), k


def test_max_iters_zero_and_shared_us_init():
    g = problems.load_golden("se3_n120")
    s, x0, N = gc.make_solver(g, "ss", 2, horizon=20, max_iters=0, tol_grad_norm=1e-12)
    rng = np.random.default_rng(3)
    us0 = 0.1 * rng.standard_normal((N, 6))          # one control path shared by the batch (us_mode 1)
    X0 = gc.perturbed_x0(x0, 2)
    out = s.solve(X0, us0)
    assert np.all(out["iters"].cpu().numpy() == 0)
    assert np.array_equal(out["us"].cpu().numpy(), np.stack([us0, us0]))
    dyn, cost, group, *_ = problems.from_golden(g, 20)
    xs = [gc.oracle_state("se3", X0[1])]
    for i in range(N):
        xs.append(dyn.f(xs[i], us0[i], i))
    assert gc.quat_rows_close(out["xs"][1].cpu().numpy(), gc.oracle_rows("se3", xs), 0) < 1e-12


def test_augmented_lagrangian_with_velocity_bounds():
    """Velocity box bounds next to the input bounds (an addition to the reference, same AL algebra): device vs oracle."""
    from trajectory_optimization_matrix_lie_groups_b200 import BatchSolver, layout, workloads
    N, dt, B = 40, 0.01, 3
    q_ref, xi_ref = workloads.helix_reference(N, dt)
    J = np.diag([0.5, 0.7, 0.9, 1.0, 1.0, 1.0])
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    R = 1e-3 * np.eye(6)
    P = 10 * Q
    lb, ub = -30.0, 30.0
    xlb = np.array([-5.0, -5, -0.5, -5, -5, -5])     # the reference spins at omega_z = 1: this bound becomes active
    xub = -xlb
    rng = np.random.default_rng(24234156)
    X0 = workloads.perturb_se3(np.eye(3), np.array([-0.3, -0.3, -0.1]), np.array([0, 0, 0.1, 2.0, 0, 0.2]), B, 0.02, rng)
    n_al, n_ilqr = 12, 60
    s = BatchSolver("se3", "al_ms", N, B)
    s.set_params(dt=dt, Ib=J[:3, :3], mass=1.0, Q=Q, R=R, P=P, max_iters=n_ilqr, tol_grad_norm=1e-6, tol_d_norm=1e-6,
                 lb=lb, ub=ub, xi_lb=xlb, xi_ub=xub, n_al_iters=n_al, tol_constr=1e-2)
    s.set_reference(layout.pose_rows(False, q_ref), xi_ref)
    out = s.solve(X0)
    al = {k: v.cpu().numpy() for k, v in s.export_al().items()}
    us = out["us"].cpu().numpy()
    xs = out["xs"].cpu().numpy()
    dyn = models.SE3Dynamics(J, dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(Q, R, P, q_ref, xi_ref)
    con = models.InputVelocityConstraint(np.full(6, lb), np.full(6, ub), xlb, xub)
    active = False
    for b in range(B):
        alc = models.ALConstrainedCost(cost, con, N)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = solvers.al_ilqr_ms(dyn, alc, con, solvers.SE3Group, N, [T for T in q_ref], xi_ref,
                                   gc.oracle_state("se3", X0[b]), np.zeros((N, 6)), n_al_iters=n_al, n_ilqr_iters=n_ilqr,
                                   tol_constr=1e-2)
        assert int(al["outer_iters"][b]) == r.outer_iterations, (b, al["outer_iters"][b], r.outer_iterations)
        assert int(out["iters"][b]) == r.inner.iterations
        assert abs(al["violation"][b] - r.violation_hist[-1]) < 1e-6 * max(1.0, abs(r.violation_hist[-1]))
        assert np.max(np.abs(us[b] - r.inner.us)) < 1e-6
        assert gc.quat_rows_close(xs[b], gc.oracle_rows("se3", r.inner.xs), 0) < 1e-6
        lm = np.concatenate((al["lmbd"][b], al["lmbd_state"][b]), axis=1)
        assert np.max(np.abs(lm - r.lmbd)) < 1e-5 * max(1.0, np.max(np.abs(r.lmbd)))
        active |= r.outer_iterations > 1 and np.max(r.lmbd[:, 12:]) > 0
    assert active, "the velocity bounds never became active"


@pytest.mark.parametrize("name,method,kw", [("se3_n120", "ss", {}), ("drone_n150", "ss", {}), ("so3_n249", "ss", {}),
                                             ("se3_n120", "ms", {"line_search": True}),
                                             ("so3_n249", "ms", {"line_search": True})])
def test_one_launch_line_search_is_bit_identical(name, method, kw):
    """Small batches roll out every line-search step size in one launch and keep every candidate (the default up to 256
    problems, so the rest of this suite already runs that way against the oracle); `set_line_search_batch(0)` restores the
    three-pass scheme of the large batches (step size 0, then the rest, then the accepted one again).  Same decisions,
    every export bit-identical — with per-problem horizons too."""
    g = problems.load_golden(name)
    B = 37
    s, x0, N = gc.make_solver(g, method, B, horizon=60, max_iters=25, tol_grad_norm=1e-12, **kw)
    X0 = gc.perturbed_x0(x0, B, scale=0.3)      # large perturbations: several problems reject the full step
    for horizons in (None, [N - (b % 5) * 7 for b in range(B)]):
        s.set_horizons(horizons)
        ref = None
        for max_batch in (0, 256):
            s.set_line_search_batch(max_batch)
            out = {k: v.cpu().numpy() for k, v in s.solve(X0).items()}
            out.update({k: v.cpu().numpy() for k, v in s.export_hist().items()})
            if ref is None:
                ref = out
                assert np.any(out["alpha_hist"] > 0), "every problem accepted the full step: the test exercises nothing"
            for k in ref:
                assert np.array_equal(ref[k], out[k]), (horizons is not None, max_batch, k)
    s.set_line_search_batch(256)
