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
