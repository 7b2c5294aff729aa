"""Oracle-generated whole-solve fixtures at the BASELINE.json configurations' OWN sizes.

    python tests/golden/make_cfg_fixtures.py [cfg3] [cfg1] [cfg4] [--jobs 8]

The reference ships result files for problem 0 of configs 2, 3 and 5 only (tests/golden/make_golden.py).  What the
headline metric is quoted on, though, is a batch of PERTURBED problems whose slowest members set the trip count of the
batch, plus two configurations (1 and 4) for which the reference ships nothing.  This script runs the CPU oracle
(oracle/, pinned to the reference's own result files by tests/test_oracle_golden.py) on

  cfg3  33 problems of the 16384-problem headline batch (workloads.se3_tracking_ms): two per perturbed parameter,
        the one with the largest perturbation and the one at the median, plus the extremes of the batch's
        iteration-count histogram (the p_y-perturbed problems that run to 27 iterations and set the batch's trip
        count, and the fastest ones at 16); WHOLE solves to the script's own stopping rule
        (main_SE3ddp_tracking_exact_ms.py:181-190: n_iter 200, tol_grad 1e-12, tol_d 1e-6).
  cfg1  problem 0 of main_SE3ddp_tracking_exact.py (single shooting, N=955, dt=0.01, 13-step line search,
        n_iter 200, tol_grad 1e-3): J_hist, accepted step-size indices, gradient norms, final trajectory.
  cfg4  the nominal problem of main_SE3ddp_tracking_exact_al_ms.py at its own parameters (helix reference, N=1400,
        R=0, u in [-10,10]^6), twice: the first 3 augmented-Lagrangian outer iterations (traopt_controller.py:
        3218-3290: inner iteration counts, max g, multipliers and penalty diagonals after the third update) and the
        whole solve (n_al 100, n_ilqr 200, tol_constr 1e-2: outer count, per-outer inner counts and violations,
        final multipliers / penalties / controls).

and stores, per problem: iteration count, stopping status, J_hist, grad_hist, defect_hist, accepted step-size indices,
the final controls and every 5th state row.  tests/test_gpu_configs.py compares the CUDA path with these on the GPU
box (the oracle needs ~1 min per N=955 multiple-shooting solve on one core, ~10 min for the single-shooting one: far
too slow to run inside the GPU tests).  Output: tests/golden/cfg{3,1,4,4_first3}_oracle.npz.  Run once here, committed with its
output; nothing under /root/reference is needed (the oracle and the package data suffice).
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.dirname(os.path.abspath(__file__))

for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[var] = "1"

XS_STRIDE = 5
# Problems at the two ends of the batch's iteration-count histogram, found by a GPU probe of the whole batch
# (scripts/probe_configs.py: 16 .. 27 iterations, 12642 of 16384 at 20): six of the 35 that run the maximum 27
# iterations and so set the batch's trip count (all p_y-perturbed), two at 26, two at the minimum 16 (p_x-perturbed).
CFG3_EXTREMES = (295, 763, 1171, 1819, 10339, 15871, 11562, 42, 162, 270)


def cfg3_selection(B=16384):
    """Two problems per perturbed parameter j = b mod 12: largest |perturbation| and the median one."""
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    rng = np.random.default_rng(workloads.SEED)
    j = np.arange(B) % 12
    val = rng.uniform(-1.0, 1.0, size=B)          # the same draw as workloads.perturb_se3
    val[0] = 0.0
    sel = []
    for p in range(12):
        idx = np.nonzero(j == p)[0]
        a = np.abs(val[idx])
        order = np.argsort(a)
        sel.append(int(idx[order[-1]]))
        sel.append(int(idx[order[len(order) // 2]]))
    return sorted(set(sel) | set(CFG3_EXTREMES))


def _state_rows(xs):
    import gpu_common as gc
    return gc.oracle_rows("se3", xs)


def job_cfg3(b):
    from oracle import models, solvers
    import gpu_common as gc
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    wl = workloads.se3_tracking_ms(B=16384)
    dyn = models.SE3Dynamics(wl.J, wl.dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
    q_ref = [np.asarray(T, dtype=float) for T in wl.q_ref]
    t0 = time.perf_counter()
    r = solvers.ilqr_ms(dyn, cost, solvers.SE3Group, wl.N, q_ref, wl.xi_ref, gc.oracle_state("se3", wl.x0_rows[b]),
                        np.zeros((wl.N, 6)), n_iterations=wl.solver["max_iters"], tol_grad_norm=wl.solver["tol_grad_norm"],
                        tol_d_norm=wl.solver["tol_d_norm"], rollout="nonlinear", line_search=False)
    return dict(b=b, x0=wl.x0_rows[b], iters=r.iterations, status=r.status, J_hist=np.array(r.J_hist), grad_hist=np.array(r.grad_hist),
                defect_hist=np.array(r.defect_hist), alpha_hist=np.array(r.alpha_hist), final_grad=r.final_grad,
                us=r.us, xs=_state_rows(r.xs)[::XS_STRIDE], seconds=time.perf_counter() - t0)


def job_cfg1(_):
    from oracle import models, solvers
    import gpu_common as gc
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    wl = workloads.se3_tracking_ss(B=1)
    dyn = models.SE3Dynamics(wl.J, wl.dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = solvers.ilqr_ss(dyn, cost, solvers.SE3Group, wl.N, gc.oracle_state("se3", wl.x0_rows[0]), np.zeros((wl.N, 6)),
                            n_iterations=wl.solver["max_iters"], tol_grad_norm=wl.solver["tol_grad_norm"], rollout="nonlinear")
    return dict(b=0, x0=wl.x0_rows[0], iters=r.iterations, status=r.status, J_hist=np.array(r.J_hist), grad_hist=np.array(r.grad_hist),
                alpha_hist=np.array(r.alpha_hist), mu_hist=np.array(r.mu_hist), us=r.us, xs=_state_rows(r.xs)[::XS_STRIDE],
                seconds=time.perf_counter() - t0)


def job_cfg4(n_outer):
    from oracle import models, solvers
    import gpu_common as gc
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    wl = workloads.se3_tracking_al_ms(B=1)
    dyn = models.SE3Dynamics(wl.J, wl.dt)
    cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
    con = models.InputConstraint(np.full(6, wl.bounds[0]), np.full(6, wl.bounds[1]))
    alc = models.ALConstrainedCost(cost, con, wl.N)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = solvers.al_ilqr_ms(dyn, alc, con, solvers.SE3Group, wl.N, [T for T in wl.q_ref], wl.xi_ref,
                               gc.oracle_state("se3", wl.x0_rows[0]), np.zeros((wl.N, 6)), n_al_iters=n_outer,
                               n_ilqr_iters=wl.solver["max_iters"], tol_constr=wl.solver["tol_constr"], verbose=True)
    return dict(b=0, x0=wl.x0_rows[0], n_outer=n_outer, outer_iterations=r.outer_iterations, converged=r.constr_converged,
                violation_hist=np.array(r.violation_hist), inner_iters_hist=np.array(r.inner_iters_hist), mu=r.mu,
                lmbd=r.lmbd, imu_diag=np.stack([np.diag(M) for M in r.Imu]), J_hist=np.array(r.inner.J_hist),
                grad_hist=np.array(r.inner.grad_hist), defect_hist=np.array(r.inner.defect_hist), inner_status=r.inner.status,
                us=r.inner.us, xs=_state_rows(r.inner.xs)[::XS_STRIDE], seconds=time.perf_counter() - t0)


def _run(job):
    name, arg = job
    t0 = time.perf_counter()
    out = {"cfg3": job_cfg3, "cfg1": job_cfg1, "cfg4": job_cfg4}[name](arg)
    print(f"[{name} {arg}] done in {time.perf_counter() - t0:.0f} s: iters {out.get('iters', out.get('inner_iters_hist'))}", flush=True)
    return name, out


def _pad(rows, fill=np.nan):
    n = max(len(r) for r in rows)
    out = np.full((len(rows), n), fill)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def main():
    import multiprocessing as mp
    which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg3", "cfg1", "cfg4"]
    jobs_n = int(sys.argv[sys.argv.index("--jobs") + 1]) if "--jobs" in sys.argv else (os.cpu_count() or 1)
    jobs = []
    if "cfg1" in which:
        jobs.append(("cfg1", 0))         # longest first
    if "cfg4" in which:
        jobs += [("cfg4", 100), ("cfg4", 3)]
    if "cfg3" in which:
        jobs += [("cfg3", b) for b in cfg3_selection()]
    with mp.get_context("fork").Pool(min(jobs_n, len(jobs))) as pool:
        results = pool.map(_run, jobs, chunksize=1)
    c3 = sorted((o for n, o in results if n == "cfg3"), key=lambda o: o["b"])
    if c3:
        np.savez_compressed(
            os.path.join(OUT, "cfg3_oracle.npz"), b=np.array([o["b"] for o in c3]), x0=np.stack([o["x0"] for o in c3]),
            iters=np.array([o["iters"] for o in c3]), status=np.array([o["status"] for o in c3]),
            J_hist=_pad([o["J_hist"] for o in c3]), grad_hist=_pad([o["grad_hist"] for o in c3]),
            defect_hist=_pad([o["defect_hist"] for o in c3]), alpha_hist=_pad([o["alpha_hist"] for o in c3], -9).astype(np.int32),
            final_grad=np.array([o["final_grad"] if o["final_grad"] is not None else np.nan for o in c3]),
            us=np.stack([o["us"] for o in c3]), xs=np.stack([o["xs"] for o in c3]), xs_stride=np.array(XS_STRIDE),
            seconds=np.array([o["seconds"] for o in c3]))
    for n, o in results:
        if n in ("cfg1", "cfg4"):
            name = n if n == "cfg1" else ("cfg4" if o["n_outer"] > 3 else "cfg4_first3")
            np.savez_compressed(os.path.join(OUT, name + "_oracle.npz"), xs_stride=np.array(XS_STRIDE),
                                **{k: np.asarray(v) for k, v in o.items() if v is not None})
    for f in ("cfg3_oracle.npz", "cfg1_oracle.npz", "cfg4_oracle.npz", "cfg4_first3_oracle.npz"):
        p = os.path.join(OUT, f)
        if os.path.exists(p):
            print(f, f"{os.path.getsize(p) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
