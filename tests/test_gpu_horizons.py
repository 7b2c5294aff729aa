"""Per-problem horizons (`trajopt_set_horizons`): problem b is the N_b-stage problem on the first N_b + 1 rows of the
reference; each one is checked against the oracle solving exactly that problem."""
import warnings

import numpy as np
import pytest

import gpu_common as gc
from oracle import problems, solvers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,method,kw", [
    ("se3_n120", "ms", {}), ("se3_n120", "ss", {}), ("so3_n249", "ms", {}), ("drone_n150", "ms", {"line_search": True}),
    ("se3_n120", "ms", {"rollout": "linear"}),
])
def test_each_problem_has_its_own_horizon(name, method, kw):
    g = problems.load_golden(name)
    kind = str(g["kind"])
    N, n_iter = 40, 5
    horizons = np.array([40, 7, 23, 1, 39, 12], dtype=np.int32)
    B = horizons.size
    s, x0, _ = gc.make_solver(g, method, B, horizon=N, max_iters=n_iter, tol_grad_norm=1e-12, **kw)
    s.set_horizons(horizons)
    X0 = gc.perturbed_x0(x0, B, scale=0.02)
    out = s.solve(X0)
    hist = {k: v.cpu().numpy() for k, v in s.export_hist().items()}
    for b in range(B):
        Nb = int(horizons[b])
        dyn, cost, group, q_ref, xi_ref, _, _ = problems.from_golden(g, Nb)
        xo = gc.oracle_state(kind, X0[b])
        us0 = np.zeros((Nb, dyn.action_size))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if method == "ms":
                r = solvers.ilqr_ms(dyn, cost, group, Nb, q_ref, xi_ref, xo, us0, n_iterations=n_iter, tol_grad_norm=1e-12,
                                    n_alphas=13 if kind == "so3" else 20, defect_kappa=1e-14 if kind == "so3" else 1e-12, **kw)
            else:
                r = solvers.ilqr_ss(dyn, cost, group, Nb, xo, us0, n_iterations=n_iter, tol_grad_norm=1e-12, **kw)
        Jo = np.array(r.J_hist)
        n = min(len(Jo), int(out["iters"][b]))
        assert n >= 1
        rel = np.abs(hist["J_hist"][b, :n] - Jo[:n]) / np.abs(Jo[:n])
        assert rel.max() < 1e-9, (b, Nb, rel.max())
        if len(Jo) == int(out["iters"][b]) and hist["alpha_hist"][b, :n].tolist() == r.alpha_hist:
            assert np.max(np.abs(out["us"][b, :Nb].cpu().numpy() - r.us)) < 1e-6
            assert gc.quat_rows_close(out["xs"][b, :Nb + 1].cpu().numpy(), gc.oracle_rows(kind, r.xs), 0) < 1e-6
        assert abs(hist["grad_hist"][b, 0] - r.grad_hist[0]) < 1e-6 * r.grad_hist[0] + 1e-12


def test_default_horizons_unchanged_and_reset():
    g = problems.load_golden("se3_n120")
    s, x0, N = gc.make_solver(g, "ms", 4, horizon=30, max_iters=6, tol_grad_norm=1e-12)
    X0 = gc.perturbed_x0(x0, 4)
    a = s.solve(X0)["us"].cpu().numpy()
    s.set_horizons([30, 30, 10, 30])
    b = s.solve(X0)["us"].cpu().numpy()
    assert np.array_equal(a[[0, 1, 3]], b[[0, 1, 3]]) and not np.array_equal(a[2, :10], b[2, :10])
    s.set_horizons(None)
    assert np.array_equal(s.solve(X0)["us"].cpu().numpy(), a)


def test_fit_batch_horizons_match_single_problem_fits():
    """Mirror class API: fit_batch(horizons=...) against `fit` of a controller built with that horizon."""
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import traopt_controller as tc, traopt_cost, traopt_dynamics
    g = problems.load_golden("se3_n120")
    N = 40
    q, xi = g["prob_q_ref"][:N + 1], g["prob_xi_ref"][:N + 1]
    dyn = traopt_dynamics.SE3Dynamics(g["prob_J"], float(g["prob_dt"]))
    cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(g["prob_Q"], g["prob_R"], g["prob_P"], q, xi)
    ctrl = tc.iLQR_Tracking_SE3_MS(dyn, cost, N, q, xi, rollout="nonlinear")
    x0 = [g["prob_x0_q"], np.asarray(g["prob_x0_xi"], dtype=float).reshape(-1)]
    horizons = [40, 9, 25]
    res = ctrl.fit_batch([x0] * 3, n_iterations=60, tol_grad_norm=1e-10, horizons=horizons)
    assert np.all(res.converged)
    for b, Nb in enumerate(horizons):
        cost_b = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(g["prob_Q"], g["prob_R"], g["prob_P"], q[:Nb + 1], xi[:Nb + 1])
        ctrl_b = tc.iLQR_Tracking_SE3_MS(dyn, cost_b, Nb, q[:Nb + 1], xi[:Nb + 1], rollout="nonlinear")
        xs, us, *_ = ctrl_b.fit(x0, np.zeros((Nb, 6)), n_iterations=60, tol_grad_norm=1e-10)
        assert np.max(np.abs(us - res.us[b][:Nb])) < 1e-9
