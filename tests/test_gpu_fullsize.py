"""GPU parity at BASELINE.json's full sizes, through size-independent properties (the NumPy oracle
needs ~50 s per N=955 solve, so it cannot check 16384 of them):

  * batch / slot / shard independence: a problem's result is bit-identical whether it is solved
    alone, inside the 16384 batch, or inside a shard of it (this is what makes multi-GPU sharding
    an equality);
  * duplicates inside a batch give bit-identical results;
  * every problem ends with the reference's stopping rule satisfied (gradient and defect norms
    under the tolerances, status CONVERGED), J_hist entries are finite, the final defect is at
    rounding level;
  * problem 0 (the script's own unperturbed x0) reproduces the replayed reference result;
  * a sample of problems spread over the batch is checked against the oracle for the first
    iterations (where the perturbation matters most).
"""
import numpy as np
import pytest

import gpu_common as gc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def headline():
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    wl = workloads.se3_tracking_ms(B=16384)
    s, x0 = wl.make_solver()
    x0 = x0.copy()
    x0[7777] = x0[5]              # a duplicate far away in the batch
    out = s.solve(x0)
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res.update({k: v.cpu().numpy() for k, v in s.export_hist().items()})
    s.close()
    return wl, x0, res


def test_headline_batch_converges(headline):
    wl, x0, res = headline
    st = res["status"]
    assert np.all(st == 0), f"{np.count_nonzero(st)} of {st.size} problems did not converge"
    assert np.all(res["grad"] < 1e-12) and np.all(res["defect"] < 1e-6)
    assert res["iters"].min() >= 10 and res["iters"].max() <= 60
    assert int(res["iters"][0]) == 20
    assert abs(res["J"][0] - 19482.538657476107) < 1e-9 * 19482.5
    it = res["iters"]
    Jh = res["J_hist"]
    mask = np.arange(Jh.shape[1])[None, :] < it[:, None]
    assert np.all(np.isfinite(Jh[mask]))
    assert np.array_equal(Jh[np.arange(it.size), it - 1], res["J"])     # J = last history entry
    # unit quaternions all along the returned trajectories (states stay on the group)
    qn = np.linalg.norm(res["xs"][::97, :, :4], axis=-1)
    assert np.max(np.abs(qn - 1.0)) < 1e-12


def test_duplicates_are_bit_identical(headline):
    wl, x0, res = headline
    for k in ("J", "iters", "grad", "defect", "xs", "us", "J_hist"):
        assert np.array_equal(res[k][7777], res[k][5]), k


def test_shard_and_single_solves_are_bit_identical(headline):
    """The multi-GPU decomposition: rank r of G solves problems [r*B/G, (r+1)*B/G) on its own."""
    wl, x0, res = headline
    lo, hi = 3 * 2048, 4 * 2048          # shard 3 of 8
    s, _ = wl.make_solver(B=hi - lo, offset=lo)
    out = s.solve(x0[lo:hi])
    for k in ("J", "iters", "status", "grad", "defect", "xs", "us"):
        assert np.array_equal(out[k].cpu().numpy(), res[k][lo:hi]), k
    s.close()
    s1, _ = wl.make_solver(B=1, offset=12345)
    one = s1.solve(x0[12345:12346])
    for k in ("J", "iters", "status", "xs", "us"):
        assert np.array_equal(one[k].cpu().numpy()[0], res[k][12345]), k


def test_sample_against_oracle_first_iterations(headline):
    """Problems spread over the batch (one per perturbed parameter), first 2 DDP iterations at N=955."""
    from oracle import lie, models, solvers
    wl, x0, res = headline
    dyn = models.SE3Dynamics(wl.J, wl.dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
    q_ref = [np.asarray(T, dtype=float) for T in wl.q_ref]
    for b in (1, 4099, 16383):
        r = solvers.ilqr_ms(dyn, cost, solvers.SE3Group, wl.N, q_ref, wl.xi_ref, gc.oracle_state("se3", x0[b]),
                            np.zeros((wl.N, 6)), n_iterations=2, tol_grad_norm=1e-12)
        Jo = np.array(r.J_hist)
        assert np.max(np.abs(res["J_hist"][b, :2] - Jo) / np.abs(Jo)) < 1e-9, b
        assert abs(res["defect_hist"][b, 0] - r.defect_hist[0]) < 1e-9 * r.defect_hist[0]


def test_drone_sweep_shard():
    """BASELINE configs[4] shape (quadrotor, N=150): one 32768-problem shard of the 2^20 sweep."""
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    B = 32768
    wl = workloads.drone_racing_ms(B=B)
    s, x0 = wl.make_solver()
    out = s.solve(x0, trajectories=False)
    st = out["status"].cpu().numpy()
    assert np.all(st == 0)
    assert int(out["iters"][0]) == 26
    assert abs(float(out["J"][0]) - 125.24481945554696) < 1e-9 * 125.2     # reference result file (2nd draft, ms_se3)
    s.close()


def test_so3_batch_1024():
    """BASELINE configs[1] (benchmark_SO3_tracking.py, batch 1024): problem 0 vs the replayed reference."""
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    from oracle import problems
    wl = workloads.so3_tracking_ms(B=1024)
    s, x0 = wl.make_solver()
    out = s.solve(x0)
    hist = s.export_hist()
    g = problems.load_golden("so3_n249")
    it = int(out["iters"][0])
    Jg = g["ms_J_hist"]
    n = min(it, len(Jg))
    assert n >= 5
    rel = np.abs(hist["J_hist"][0, :n].cpu().numpy() - Jg[:n]) / np.abs(Jg[:n])
    assert rel.max() < 1e-9
    assert np.all((out["status"].cpu().numpy() & 15) <= 1)       # converged or ran out of the script's 50 iterations
    s.close()
