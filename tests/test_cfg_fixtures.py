"""The oracle-generated configuration fixtures (tests/golden/cfg*_oracle.npz) stay in sync with their generator, the
workload definitions and the oracle: problem selection, initial states, and a short re-run of the oracle that must
reproduce the stored histories bit for bit (whole solves are re-run by tests/golden/make_cfg_fixtures.py only)."""
import os
import re
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)

import gpu_common as gc  # noqa: E402
from oracle import models, solvers  # noqa: E402


def _fx(name):
    with np.load(os.path.join(GOLD, name + "_oracle.npz")) as z:
        return {k: z[k] for k in z.files}


def test_every_fixture_of_the_generators_is_committed():
    import make_cfg_fixtures
    import make_golden
    for name in make_golden.FILES:
        assert os.path.exists(os.path.join(GOLD, name + ".npz")), name
    for name in ("cfg1", "cfg3", "cfg4", "cfg4_first3"):
        assert os.path.exists(os.path.join(GOLD, name + "_oracle.npz")), name
    fx = _fx("cfg3")
    assert fx["b"].tolist() == make_cfg_fixtures.cfg3_selection()
    assert set(np.asarray(fx["b"]) % 12) == set(range(12))               # every perturbed parameter
    assert int(fx["iters"].min()) == 16 and int(fx["iters"].max()) == 27    # both ends of the batch's histogram
    assert np.all(fx["status"] == solvers.STATUS_CONVERGED)


def test_fixture_initial_states_are_the_workloads():
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    fx = _fx("cfg3")
    assert np.array_equal(workloads.se3_tracking_ms(B=16384).x0_rows[fx["b"]], fx["x0"])
    assert np.array_equal(workloads.se3_tracking_ss(B=1).x0_rows[0], _fx("cfg1")["x0"])
    wl4 = workloads.se3_tracking_al_ms(B=1)
    assert np.array_equal(wl4.x0_rows[0], _fx("cfg4")["x0"]) and wl4.N == 1400


def test_oracle_reproduces_the_cfg3_fixture_bitwise():
    """first 2 iterations of a 27-iteration problem of the headline batch"""
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    fx = _fx("cfg3")
    i = int(np.nonzero(fx["iters"] == 27)[0][0])
    wl = workloads.se3_tracking_ms(B=16384)
    dyn = models.SE3Dynamics(wl.J, wl.dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
    r = solvers.ilqr_ms(dyn, cost, solvers.SE3Group, wl.N, [np.asarray(T, dtype=float) for T in wl.q_ref], wl.xi_ref,
                        gc.oracle_state("se3", fx["x0"][i]), np.zeros((wl.N, 6)), n_iterations=2, tol_grad_norm=1e-12)
    assert np.array_equal(np.array(r.J_hist), fx["J_hist"][i, :2])
    assert np.array_equal(np.array(r.defect_hist), fx["defect_hist"][i, :3])


def test_oracle_reproduces_the_cfg1_fixture_bitwise():
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    fx = _fx("cfg1")
    wl = workloads.se3_tracking_ss(B=1)
    dyn = models.SE3Dynamics(wl.J, wl.dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
    r = solvers.ilqr_ss(dyn, cost, solvers.SE3Group, wl.N, gc.oracle_state("se3", fx["x0"]), np.zeros((wl.N, 6)),
                        n_iterations=2, tol_grad_norm=1e-3, rollout="nonlinear")
    assert np.array_equal(np.array(r.J_hist), fx["J_hist"][:2]) and r.alpha_hist == fx["alpha_hist"][:2].tolist()
    # the survey's replay of this configuration: 0 x 37, 4, 6 ..., J_final = 26963.216612992604
    assert fx["alpha_hist"][:39].tolist() == [0] * 37 + [4, 6]
    assert abs(fx["J_hist"][-1] - 26963.216612992604) < 1e-9 * 26963.2


def test_oracle_reproduces_the_cfg4_fixture_first_outer_iteration():
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    fx = _fx("cfg4_first3")
    full = _fx("cfg4")
    assert np.array_equal(full["violation_hist"][:3], fx["violation_hist"]) and np.array_equal(full["inner_iters_hist"][:3], fx["inner_iters_hist"])
    assert bool(full["converged"]) and full["violation_hist"][-1] < 1e-2
    wl = workloads.se3_tracking_al_ms(B=1)
    dyn = models.SE3Dynamics(wl.J, wl.dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
    con = models.InputConstraint(np.full(6, wl.bounds[0]), np.full(6, wl.bounds[1]))
    alc = models.ALConstrainedCost(cost, con, wl.N)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = solvers.al_ilqr_ms(dyn, alc, con, solvers.SE3Group, wl.N, [T for T in wl.q_ref], wl.xi_ref,
                               gc.oracle_state("se3", fx["x0"]), np.zeros((wl.N, 6)), n_al_iters=1, n_ilqr_iters=200, tol_constr=1e-2)
    assert r.inner_iters_hist == [int(fx["inner_iters_hist"][0])] and r.violation_hist[0] == fx["violation_hist"][0]


def test_workloads_do_not_read_the_test_fixtures():
    src = open(os.path.join(ROOT, "trajectory_optimization_matrix_lie_groups_b200", "workloads.py")).read()
    assert not re.search(r"tests|golden", src.split('"""', 2)[2])
