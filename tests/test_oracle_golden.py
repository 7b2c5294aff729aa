"""Pin the CPU oracle to the reference's own shipped results.

Each golden .npz is a re-packed result pickle of the reference (tests/golden/make_golden.py).
The oracle must reproduce the reference's iteration count, every cost-history entry to 1e-12
relative, the accepted step-size sequence, and the final trajectories.
"""
import numpy as np
import pytest

from oracle import problems, solvers


def _run(name, mode, max_iter=200, horizon=None):
    g = problems.load_golden(name)
    kind = str(g["kind"])
    dyn, cost, group, q_ref, xi_ref, x0, N = problems.from_golden(g, horizon)
    us0 = np.zeros((N, dyn.action_size))
    if mode == "ms":
        so3 = kind in ("so3", "pendulum")
        r = solvers.ilqr_ms(dyn, cost, group, N, q_ref, xi_ref, x0, us0, n_iterations=max_iter,
                            tol_grad_norm=1e-12, n_alphas=13 if so3 else 20,
                            defect_kappa=1e-14 if so3 else 1e-12, append_final_grad=so3)
    else:
        r = solvers.ilqr_ss(dyn, cost, group, N, x0, us0, n_iterations=max_iter, tol_grad_norm=1e-12,
                            rollout="nonlinear")
    return g, kind, r


def _check_hist(r, g, mode, n=None, rtol=1e-12):
    Jg = g[mode + "_J_hist"]
    n = len(Jg) if n is None else n
    J = np.array(r.J_hist[:n])
    assert len(J) == n
    assert np.max(np.abs(J - Jg[:n]) / np.abs(Jg[:n])) < rtol
    if mode == "ms":
        dg = g["ms_defect_hist"][:n + 1]
        assert np.max(np.abs(np.array(r.defect_hist[:n + 1]) - dg)) < 1e-10 * max(1.0, dg[0])


def _check_final(r, g, kind, mode, us_atol, xs_atol):
    P, V = problems.poses_to_matrices(kind, r.xs)
    assert np.max(np.abs(r.us - g[mode + "_us"])) < us_atol
    assert np.max(np.abs(P - g[mode + "_xs_q"])) < xs_atol
    assert np.max(np.abs(V - g[mode + "_xs_xi"])) < xs_atol


def test_se3_ms_n120_full():
    g, kind, r = _run("se3_n120", "ms")
    assert r.iterations == 56 and r.status == solvers.STATUS_CONVERGED
    _check_hist(r, g, "ms")
    _check_final(r, g, kind, "ms", 1e-9, 1e-11)
    assert r.mu_hist[0] == 0.0 and not r.reg_exceeded


def test_se3_ss_n120_full_line_search_decisions():
    g, kind, r = _run("se3_n120", "ss")
    # the reference stops with "Couldn't find descent direction" after 25 iterations
    assert r.iterations == 25 and r.status == solvers.STATUS_NO_DESCENT
    assert r.alpha_hist == [1] + [0] * 23 + [-1]
    _check_hist(r, g, "ss")
    _check_final(r, g, kind, "ss", 1e-9, 1e-11)
    gh = g["ss_grad_hist"]
    assert np.max(np.abs(np.array(r.grad_hist) - gh) / gh) < 1e-8


def test_so3_ms_n249_full():
    g, kind, r = _run("so3_n249", "ms")
    assert r.iterations == 15 and r.status == solvers.STATUS_CONVERGED
    _check_hist(r, g, "ms")
    _check_final(r, g, kind, "ms", 1e-10, 1e-12)
    assert len(r.grad_hist) == len(g["ms_grad_hist"]) == 16      # SO3_MS.fit appends the last one itself


def test_so3_ss_n249_first_20_of_50():
    g, kind, r = _run("so3_n249", "ss", max_iter=20)
    assert r.alpha_hist == [0] * 20
    _check_hist(r, g, "ss", n=20)


def test_drone_ms_n150_full():
    g, kind, r = _run("drone_n150", "ms")
    assert r.iterations == 26 and r.status == solvers.STATUS_CONVERGED
    _check_hist(r, g, "ms")
    _check_final(r, g, kind, "ms", 1e-9, 1e-11)


def test_drone_ss_n150_full():
    g, kind, r = _run("drone_n150", "ss")
    assert r.iterations == 10 and r.status == solvers.STATUS_NO_DESCENT
    assert r.alpha_hist == [0] * 9 + [-1]
    _check_hist(r, g, "ss", rtol=1e-10)
    _check_final(r, g, kind, "ss", 1e-8, 1e-10)


@pytest.mark.parametrize("name", ["se3_n955_r1e-5", "se3_n955_r1e-4"])
def test_se3_ms_n955_first_iterations(name):
    """Headline-config horizon: the first 2 iterations pin the arithmetic at N=955 in seconds.

    (The 1st-draft SE3 pickle, `draft1_se3_n955`, was produced by an older revision of the
    reference library and does not replay with today's code: 1.4 % off at iteration 0.  It is kept
    as a fixture for its problem definition only.  1st-draft SO3 and drone pickles do replay.)
    """
    g, kind, r = _run(name, "ms", max_iter=2)
    _check_hist(r, g, "ms", n=2)


@pytest.mark.slow
@pytest.mark.parametrize("name,iters", [("se3_n955_r1e-5", 20), ("se3_n955_r1e-4", 24)])
def test_se3_ms_n955_full(name, iters):
    g, kind, r = _run(name, "ms")
    assert r.iterations == iters and r.status == solvers.STATUS_CONVERGED
    _check_hist(r, g, "ms")
    _check_final(r, g, kind, "ms", 1e-9, 1e-11)


@pytest.mark.slow
@pytest.mark.parametrize("name,mode,n", [("draft1_so3_n249", "ms", 50), ("draft1_drone_n500", "ms", 20),
                                         ("so3_n249", "ss", 50), ("se3_n955_r1e-5", "ss", 32)])
def test_more_goldens(name, mode, n):
    g, kind, r = _run(name, mode, max_iter=n)
    _check_hist(r, g, mode, n=n, rtol=1e-11)


def test_drone_ss_n500_line_search_walks_down_the_step_sizes():
    """1st-draft quadrotor result file, single shooting: the reference's line search accepts step
    indices 2,0,0,3,5,7,8,9,9,9,10,10,11,11 and then fails — decision parity at every rung."""
    g, kind, r = _run("draft1_drone_n500", "ss")
    assert r.iterations == 15 and r.status == solvers.STATUS_NO_DESCENT
    assert r.alpha_hist == [2, 0, 0, 3, 5, 7, 8, 9, 9, 9, 10, 10, 11, 11, -1]
    _check_hist(r, g, "ss")
    _check_final(r, g, kind, "ss", 1e-8, 1e-10)


def test_so3_draft1_both_methods():
    for mode in ("ms", "ss"):
        g, kind, r = _run("draft1_so3_n249", mode, max_iter=50)
        assert r.iterations == 50
        _check_hist(r, g, mode)
        _check_final(r, g, kind, mode, 1e-10, 1e-12)


@pytest.mark.slow
def test_se3_ss_n955_line_search_decisions():
    g, kind, r = _run("se3_n955_r1e-4", "ss")
    assert r.iterations == 34 and r.alpha_hist == [0] * 32 + [1, -1]
    _check_hist(r, g, "ss")
    g, kind, r = _run("se3_n955_r1e-5", "ss", max_iter=32)
    assert r.alpha_hist == [0] * 28 + [2, 8, 1, 11]
    _check_hist(r, g, "ss")


def test_pendulum_n80_both_methods():
    """Pendulum3dDyanmics (state-dependent f_u, gravity torque): the reference's swing-up result file."""
    g, kind, r = _run("pendulum_n80", "ms", max_iter=100)
    assert r.iterations == 19 and r.status == solvers.STATUS_CONVERGED
    _check_hist(r, g, "ms")
    _check_final(r, g, kind, "ms", 1e-10, 1e-12)
    g, kind, r = _run("pendulum_n80", "ss", max_iter=100)
    assert r.iterations == 100 and r.alpha_hist == [0] * 100
    _check_hist(r, g, "ss")
    _check_final(r, g, kind, "ss", 1e-10, 1e-12)
