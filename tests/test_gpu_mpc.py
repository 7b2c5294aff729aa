"""Receding-horizon driver: closed loop over a batch, warm starts, plant = library dynamics."""
import numpy as np
import pytest

import gpu_common as gc
from oracle import problems, solvers

pytestmark = pytest.mark.gpu


def _setup(name="se3_n120"):
    g = problems.load_golden(name)
    kind = str(g["kind"])
    Ib, mass = gc.inertia_parts(kind, g["prob_J"])
    from trajectory_optimization_matrix_lie_groups_b200 import layout
    x0 = np.concatenate((layout.pose_rows(kind == "so3", g["prob_x0_q"]), np.asarray(g["prob_x0_xi"], dtype=float).reshape(-1)))
    return g, kind, Ib, mass, x0


@pytest.mark.parametrize("method", ["ms", "ss"])
def test_first_step_is_the_plain_solve_and_plant_is_the_oracle_dynamics(method):
    from trajectory_optimization_matrix_lie_groups_b200 import mpc
    g, kind, Ib, mass, x0 = _setup()
    B, N, T = 3, 20, 4
    X0 = gc.perturbed_x0(x0, B, scale=0.02)
    res = mpc.receding_horizon(kind, method, q_ref=g["prob_q_ref"], xi_ref=g["prob_xi_ref"], x0_rows=X0, N=N, T=T,
                               dt=float(g["prob_dt"]), Ib=Ib, mass=mass, Q=g["prob_Q"], R=g["prob_R"], P=g["prob_P"],
                               n_iterations=4, tol_grad_norm=1e-12)
    assert res.xs.shape == (B, T + 1, 13) and res.us.shape == (B, T, 6)
    dyn, cost, group, q_ref, xi_ref, _, _ = problems.from_golden(g, N)
    for b in range(B):
        xo = gc.oracle_state(kind, X0[b])
        if method == "ms":
            r = solvers.ilqr_ms(dyn, cost, group, N, q_ref, xi_ref, xo, np.zeros((N, 6)), n_iterations=4, tol_grad_norm=1e-12)
        else:
            r = solvers.ilqr_ss(dyn, cost, group, N, xo, np.zeros((N, 6)), n_iterations=4, tol_grad_norm=1e-12)
        assert np.max(np.abs(res.us[b, 0] - r.us[0])) < 1e-7            # first applied control = first control of the plan
        x1 = dyn.f(xo, res.us[b, 0], 0)                                 # plant step = the reference's f
        assert gc.quat_rows_close(res.xs[b, 1], gc.oracle_rows(kind, [x1])[0], 0) < 1e-12
        assert abs(res.J[0, b] - r.J_hist[-1]) < 1e-9 * abs(r.J_hist[-1])


def test_closed_loop_tracks_and_warm_start_helps():
    from trajectory_optimization_matrix_lie_groups_b200 import mpc, layout
    g, kind, Ib, mass, x0 = _setup()
    B, N, T = 8, 25, 60
    X0 = gc.perturbed_x0(x0, B, scale=0.02)
    common = dict(q_ref=g["prob_q_ref"], xi_ref=g["prob_xi_ref"], x0_rows=X0, N=N, T=T, dt=float(g["prob_dt"]), Ib=Ib,
                  mass=mass, Q=g["prob_Q"], R=g["prob_R"], P=g["prob_P"], n_iterations=2, tol_grad_norm=1e-9)
    warm = mpc.receding_horizon(kind, "ms", warm_start=True, **common)
    cold = mpc.receding_horizon(kind, "ms", warm_start=False, **common)
    ref_p = g["prob_q_ref"][:T + 1, :3, 3]
    err_w = np.linalg.norm(warm.xs[:, :, 4:7] - ref_p[None], axis=2)
    err_c = np.linalg.norm(cold.xs[:, :, 4:7] - ref_p[None], axis=2)
    assert np.all(np.isfinite(warm.xs)) and np.max(np.abs(np.linalg.norm(warm.xs[..., :4], axis=-1) - 1)) < 1e-12
    assert np.all(err_w[:, -1] < 0.25 * err_w[:, 0])          # the loop closes the initial position error
    # with 2 iterations per step the warm-started plans are at least as good on average as cold-started ones
    assert warm.J[5:].mean() <= cold.J[5:].mean() * (1 + 1e-6)
    assert np.mean(err_w[:, -1]) <= np.mean(err_c[:, -1]) * 1.05
    # determinism: the same call twice gives bit-identical closed loops
    again = mpc.receding_horizon(kind, "ms", warm_start=True, **common)
    assert np.array_equal(again.xs, warm.xs) and np.array_equal(again.us, warm.us)


def test_sliding_reference_window_equals_an_explicit_window():
    """trajopt_set_reference_long + _offset(t) is the same problem as trajopt_set_reference on samples [t, t + N]."""
    from trajectory_optimization_matrix_lie_groups_b200 import layout
    g, kind, Ib, mass, x0 = _setup()
    B, N, t = 5, 20, 7
    X0 = gc.perturbed_x0(x0, B, scale=0.02)
    q_rows = layout.pose_rows(False, g["prob_q_ref"])
    a, _, _ = gc.make_solver(g, "ms", B, horizon=N, max_iters=6, tol_grad_norm=1e-12)
    b, _, _ = gc.make_solver(g, "ms", B, horizon=N, max_iters=6, tol_grad_norm=1e-12)
    a.set_reference(q_rows[t:t + N + 1], g["prob_xi_ref"][t:t + N + 1])
    b.set_reference_long(q_rows, g["prob_xi_ref"])
    b.set_reference_offset(t)
    ra, rb = a.solve(X0), b.solve(X0)
    for k in ("J", "iters", "status", "xs", "us"):
        assert np.array_equal(ra[k].cpu().numpy(), rb[k].cpu().numpy()), k
    with pytest.raises(Exception):
        b.set_reference_offset(q_rows.shape[0] - N)        # window would run past the last sample
    b.set_reference_offset(0)
    a.set_reference(q_rows[:N + 1], g["prob_xi_ref"][:N + 1])
    assert np.array_equal(a.solve(X0)["us"].cpu().numpy(), b.solve(X0)["us"].cpu().numpy())


def test_mpc_reports_its_rate():
    from trajectory_optimization_matrix_lie_groups_b200 import mpc
    g, kind, Ib, mass, x0 = _setup()
    B, N, T = 64, 20, 10
    res = mpc.receding_horizon(kind, "ms", q_ref=g["prob_q_ref"], xi_ref=g["prob_xi_ref"], x0_rows=gc.perturbed_x0(x0, B, scale=0.01),
                               N=N, T=T, dt=float(g["prob_dt"]), Ib=Ib, mass=mass, Q=g["prob_Q"], R=g["prob_R"], P=g["prob_P"],
                               n_iterations=2, tol_grad_norm=1e-9)
    assert res.seconds > 0 and res.steps_per_second > 0 and res.us.shape == (B, T, 6)
