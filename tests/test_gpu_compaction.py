"""Compaction of the running problems into the leading slots (host_impl.cuh `maybe_compact`) must change no
number: every export — trajectories, summaries, histories, AL multipliers — is bit-identical with and without it."""
import numpy as np
import pytest

import gpu_common as gc
from oracle import problems

pytestmark = pytest.mark.gpu


def _all_outputs(s, X0, us0=None):
    out = {k: v.cpu().numpy() for k, v in s.solve(X0, us0).items()}
    out.update({k: v.cpu().numpy() for k, v in s.export_hist().items()})
    mu, delta = s.export_reg()
    out["mu"], out["delta"] = mu.cpu().numpy(), delta.cpu().numpy()
    return out


@pytest.mark.parametrize("name,method,kw", [
    ("drone_n150", "ms", {}),
    ("se3_n120", "ss", {}),
    ("so3_n249", "ms", {"line_search": True}),
    ("pendulum_n80", "ms", {}),
])
def test_compaction_changes_nothing(name, method, kw):
    g = problems.load_golden(name)
    B = 200                      # ragged (not a multiple of 32); iteration counts differ across the batch
    horizon = 60
    a, x0, N = gc.make_solver(g, method, B, horizon=horizon, max_iters=40, tol_grad_norm=1e-10, **kw)
    b, _, _ = gc.make_solver(g, method, B, horizon=horizon, max_iters=40, tol_grad_norm=1e-10, **kw)
    a.set_compaction(-1, 4)      # never
    b.set_compaction(0, 1)       # as soon as a single slot in use is idle
    X0 = gc.perturbed_x0(x0, B, scale=0.05)
    rng = np.random.default_rng(2)
    us0 = 0.05 * rng.standard_normal((B, N, a.NU))     # per-problem initial controls: exercised through `orig`
    ra, rb = _all_outputs(a, X0, us0), _all_outputs(b, X0, us0)
    assert len(np.unique(ra["iters"])) > 1, "the batch should finish at different iterations"
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), k
    # and again on the same handles (state of a previous compaction must not leak into the next solve)
    rb2 = _all_outputs(b, X0, us0)
    for k in ra:
        assert np.array_equal(ra[k], rb2[k]), k


def test_compaction_augmented_lagrangian():
    from trajectory_optimization_matrix_lie_groups_b200 import BatchSolver, layout, workloads
    N, dt, B = 40, 0.01, 70
    q_ref, xi_ref = workloads.helix_reference(N, dt)
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    rng = np.random.default_rng(24234156)
    X0 = workloads.perturb_se3(np.eye(3), np.array([-0.3, -0.3, -0.1]), np.array([0, 0, 0.1, 2.0, 0, 0.2]), B, 0.03, rng)
    res = []
    for min_batch, ratio in ((-1, 4), (0, 1)):
        s = BatchSolver("se3", "al_ms", N, B)
        s.set_params(dt=dt, Ib=np.diag([0.5, 0.7, 0.9]), mass=1.0, Q=Q, R=np.zeros((6, 6)), P=10 * Q, max_iters=60,
                     tol_grad_norm=1e-6, tol_d_norm=1e-6, lb=-8.0, ub=8.0, n_al_iters=15, tol_constr=1e-2)
        s.set_reference(layout.pose_rows(False, q_ref), xi_ref)
        s.set_compaction(min_batch, ratio)
        out = {k: v.cpu().numpy() for k, v in s.solve(X0).items()}
        out.update({k: v.cpu().numpy() for k, v in s.export_al().items()})
        res.append(out)
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k


def test_second_solve_on_one_handle_after_compaction_keeps_references_and_horizons():
    """Per-problem references and horizons are persistent settings ("set before trajopt_begin"); a compaction moves them
    with their problems, and the next solve on the same handle must find them back in the caller's order.  DISTINCT
    references and horizons at a batch size where compaction is on by default: a mix-up changes the results."""
    from trajectory_optimization_matrix_lie_groups_b200 import layout
    g = problems.load_golden("se3_n120")
    B, N = 1056, 30
    rng = np.random.default_rng(5)
    shift = rng.integers(0, 12, size=B)
    q_rows_all = layout.pose_rows(False, g["prob_q_ref"])
    q_rows = np.stack([q_rows_all[s:s + N + 1] for s in shift])
    xi_rows = np.stack([g["prob_xi_ref"][s:s + N + 1] for s in shift])
    horizons = rng.integers(5, N + 1, size=B).astype(np.int32)
    X0 = None
    res = []
    for min_batch, ratio in ((-1, 4), (0, 1), (1024, 4)):
        s, x0, _ = gc.make_solver(g, "ms", B, horizon=N, max_iters=40, tol_grad_norm=1e-10)
        s.set_compaction(min_batch, ratio)
        s.set_reference_batch(q_rows, xi_rows)
        s.set_horizons(horizons)
        if X0 is None:
            X0 = gc.perturbed_x0(x0, B, scale=0.02)
        first = _all_outputs(s, X0)
        second = _all_outputs(s, X0)              # same handle, nothing set again
        res.append((first, second))
        if min_batch >= 0:                        # settings given anew after a compacted solve are taken in the caller's order
            s.set_horizons(horizons)
            third = _all_outputs(s, X0)
            for k in first:
                assert np.array_equal(first[k], third[k]), k
        s.close()
    ref = res[0][0]
    assert len(np.unique(ref["iters"])) > 2 and np.all(np.isfinite(ref["J"]))
    for first, second in res:
        for k in ref:
            assert np.array_equal(ref[k], first[k]), k
            assert np.array_equal(ref[k], second[k]), ("second solve", k)
