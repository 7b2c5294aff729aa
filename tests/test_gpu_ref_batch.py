"""Per-problem references (`trajopt_set_reference_batch`): a batch whose problems differ in their reference as
well as in their initial state.  Each problem is checked against the oracle solving it alone with its own
reference; and with B copies of ONE reference the results are bit-identical to the shared-reference path."""
import warnings

import numpy as np
import pytest

import gpu_common as gc
from oracle import problems, solvers

pytestmark = pytest.mark.gpu


def _windows(g, B, N, stride):
    q, xi = g["prob_q_ref"], g["prob_xi_ref"]
    return [(q[b * stride:b * stride + N + 1], xi[b * stride:b * stride + N + 1]) for b in range(B)]


@pytest.mark.parametrize("name,method", [("se3_n120", "ms"), ("se3_n120", "ss"), ("so3_n249", "ms"), ("drone_n150", "ms")])
def test_each_problem_tracks_its_own_reference(name, method):
    from trajectory_optimization_matrix_lie_groups_b200 import layout
    g = problems.load_golden(name)
    kind = str(g["kind"])
    B, N, n_iter = 4, 30, 5
    s, x0, _ = gc.make_solver(g, method, B, horizon=N, max_iters=n_iter, tol_grad_norm=1e-12)
    wins = _windows(g, B, N, 9)                      # time-shifted windows of the fixture's reference
    q_rows = np.stack([layout.pose_rows(kind == "so3", q) for q, _ in wins])
    xi_rows = np.stack([xi for _, xi in wins])
    s.set_reference_batch(q_rows, xi_rows)
    X0 = gc.perturbed_x0(x0, B, scale=0.02)
    out = s.solve(X0)
    hist = {k: v.cpu().numpy() for k, v in s.export_hist().items()}
    for b in range(B):
        gb = dict(g)
        gb["prob_q_ref"], gb["prob_xi_ref"] = wins[b]
        dyn, cost, group, q_ref, xi_ref, _, _ = problems.from_golden(gb, N)
        xo = gc.oracle_state(kind, X0[b])
        us0 = np.zeros((N, dyn.action_size))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if method == "ms":
                r = solvers.ilqr_ms(dyn, cost, group, N, q_ref, xi_ref, xo, us0, n_iterations=n_iter, tol_grad_norm=1e-12,
                                    n_alphas=13 if kind == "so3" else 20)
            else:
                r = solvers.ilqr_ss(dyn, cost, group, N, xo, us0, n_iterations=n_iter, tol_grad_norm=1e-12)
        it = int(out["iters"][b])
        assert it == r.iterations
        Jo = np.array(r.J_hist)
        assert np.max(np.abs(hist["J_hist"][b, :it] - Jo) / np.abs(Jo)) < 1e-9
        assert np.max(np.abs(out["us"][b].cpu().numpy() - r.us)) < 1e-7
        assert gc.quat_rows_close(out["xs"][b].cpu().numpy(), gc.oracle_rows(kind, r.xs), 0) < 1e-7


def test_identical_references_equal_the_shared_path():
    """(Not bit-identical: the per-problem-reference linearisation is a separate kernel instantiation and the compiler
    is free to contract its multiply-adds differently; the packed reference rows themselves are bit-identical.)"""
    from trajectory_optimization_matrix_lie_groups_b200 import layout
    g = problems.load_golden("se3_n120")
    B, N = 70, 40
    a, x0, _ = gc.make_solver(g, "ms", B, horizon=N, max_iters=30, tol_grad_norm=1e-10)
    b, _, _ = gc.make_solver(g, "ms", B, horizon=N, max_iters=30, tol_grad_norm=1e-10)
    q_rows = layout.pose_rows(False, g["prob_q_ref"][:N + 1])
    b.set_reference_batch(np.tile(q_rows, (B, 1, 1)), np.tile(g["prob_xi_ref"][:N + 1], (B, 1, 1)))
    b.set_compaction(0, 1)                       # the per-problem reference array moves with its problem
    X0 = gc.perturbed_x0(x0, B, scale=0.05)
    ra, rb = a.solve(X0), b.solve(X0)
    for k in ("iters", "status"):
        assert np.array_equal(ra[k].cpu().numpy(), rb[k].cpu().numpy()), k
    Ja, Jb = ra["J"].cpu().numpy(), rb["J"].cpu().numpy()
    assert np.max(np.abs(Ja - Jb) / np.abs(Ja)) < 1e-12
    assert np.max(np.abs(ra["us"].cpu().numpy() - rb["us"].cpu().numpy())) < 1e-9
    assert np.max(np.abs(ra["xs"].cpu().numpy() - rb["xs"].cpu().numpy())) < 1e-10
    # switching back to a shared reference
    b.set_reference(q_rows, g["prob_xi_ref"][:N + 1])
    rc = b.solve(X0)
    assert np.array_equal(ra["us"].cpu().numpy(), rc["us"].cpu().numpy())
