"""`fit_batch(shard=True)` under a torch.distributed job: two ranks (gloo, both on cuda:0 — the
N>1 NCCL path itself is what bench.py --gpus N runs) solve the two halves of a batch; every rank ends with the whole
batch's summaries, its own shard's trajectories and histories, and — for the augmented-Lagrangian controller — its
shard's multipliers.  Results equal the unsharded solve bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(al):
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import (traopt_constraints, traopt_controller as tc,
                                                                              traopt_cost, traopt_dynamics)
    N, dt, B = 40, 0.01, 7
    q_ref, xi_ref = workloads.helix_reference(N, dt)
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    dyn = traopt_dynamics.SE3Dynamics(np.diag([0.5, 0.7, 0.9, 1.0, 1.0, 1.0]), dt)
    rng = np.random.default_rng(24234156)
    X0 = workloads.perturb_se3(np.eye(3), np.array([-0.3, -0.3, -0.1]), np.array([0, 0, 0.1, 2.0, 0, 0.2]), B, 0.03, rng)
    if al:
        cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, np.zeros((6, 6)), 10 * Q, q_ref, xi_ref)
        ctrl = tc.AL_iLQR_Tracking_SE3_MS(dyn, cost, traopt_constraints.InputConstraint(-8.0, 8.0), N, q_ref, xi_ref)
        kw = dict(n_al_iters=15, n_ilqr_iters=60)
    else:
        cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(Q, 1e-3 * np.eye(6), 10 * Q, q_ref, xi_ref)
        ctrl = tc.iLQR_Tracking_SE3_MS(dyn, cost, N, q_ref, xi_ref, rollout="nonlinear")
        kw = dict(n_iterations=60, tol_grad_norm=1e-10)
    return ctrl, X0, kw


def _pack(res):
    out = {k: getattr(res, k) for k in ("J", "iters", "status", "grad", "defect", "us", "xs_rows", "J_hist", "alpha_hist")}
    out["shard"] = np.array(res.shard)
    out.update({"al_" + k: v for k, v in res.extra.items()})
    return out


def _worker(rank, world, port, al, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        ctrl, X0, kw = _problem(al)
        res = ctrl.fit_batch(X0, return_hist=True, **kw)          # shard defaults to True inside a job
        q.put((rank, _pack(res)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("al", [False, True], ids=["ms", "al_ms"])
def test_fit_batch_sharded_world2(al):
    import torch.multiprocessing as mp
    ctrl, X0, kw = _problem(al)
    whole = _pack(ctrl.fit_batch(X0, return_hist=True, shard=False, **kw))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, al, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0]["shard"].tolist() == [0, 4] and got[1]["shard"].tolist() == [4, 7]
    for rank in (0, 1):
        o = got[rank]
        lo, hi = o["shard"]
        for k in ("J", "iters", "status", "grad", "defect"):          # whole batch on every rank
            assert np.array_equal(o[k], whole[k]), (rank, k)
        for k in [k for k in whole if k not in ("J", "iters", "status", "grad", "defect", "shard")]:   # this rank's shard
            assert np.array_equal(o[k], whole[k][lo:hi]), (rank, k)
    if al:
        assert whole["al_outer_iters"].max() > 1
