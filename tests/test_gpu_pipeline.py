"""Batches in flight (`PipelinedSolver`): same results as solving the batches one after the other."""
import numpy as np
import pytest

import gpu_common as gc
from oracle import problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_pipelined_results_equal_serial(depth):
    from trajectory_optimization_matrix_lie_groups_b200 import PipelinedSolver
    g = problems.load_golden("drone_n150")
    B = 96
    serial, x0, N = gc.make_solver(g, "ms", B, max_iters=40, tol_grad_norm=1e-12)
    batches = [gc.perturbed_x0(x0, B, seed=s, scale=0.02) for s in range(5)]
    want = [{k: v.cpu().numpy() for k, v in serial.solve(X).items()} for X in batches]
    with PipelinedSolver(lambda: gc.make_solver(g, "ms", B, max_iters=40, tol_grad_norm=1e-12)[0], depth=depth) as pipe:
        got = list(pipe.map(batches))
        for w, o in zip(want, got):
            for k in ("J", "iters", "status", "grad", "defect", "xs", "us"):
                assert np.array_equal(o[k].cpu().numpy(), w[k]), k
        # host-buffer path through the lanes, results in submission order
        futs = [pipe.submit(X, host=True) for X in batches]
        for w, f in zip(want, futs):
            o = f.result()
            assert np.array_equal(o["iters"], w["iters"]) and np.array_equal(o["us"], w["us"])


@pytest.mark.parametrize("pinned", [True, False])
def test_host_solves_overlap_their_copies(pinned):
    """trajopt_solve_host_begin / _wait: the next solve on a handle starts while the previous one's device->host copies are
    still queued; several tickets outstanding, waited out of band.  Every host array equals the device result."""
    import torch
    g = problems.load_golden("drone_n150")
    B = 300                              # ragged convergence: the early copy (three quarters stopped) is taken
    s, x0, N = gc.make_solver(g, "ms", B, horizon=60, max_iters=40, tol_grad_norm=1e-10)
    batches = [gc.perturbed_x0(x0, B, seed=k, scale=0.05) for k in range(6)]
    want = [{k: v.cpu().numpy() for k, v in s.solve(X).items()} for X in batches]
    assert len(np.unique(want[0]["iters"])) > 2

    def buffers():
        o = {"J": torch.empty(B, dtype=torch.float64), "grad": torch.empty(B, dtype=torch.float64), "defect": torch.empty(B, dtype=torch.float64),
             "iters": torch.empty(B, dtype=torch.int32), "status": torch.empty(B, dtype=torch.int32),
             "xs": torch.empty(B, N + 1, s.NS, dtype=torch.float64), "us": torch.empty(B, N, s.NU, dtype=torch.float64)}
        if pinned:
            o = {k: v.pin_memory() for k, v in o.items()}
        return o, {k: v.numpy() for k, v in o.items()}
    bufs = [buffers() for _ in batches]
    tickets = []
    for k, X in enumerate(batches):
        if len(tickets) == 3:                          # at most three tickets in flight here (the handle allows four)
            t, i = tickets.pop(0)
            s.solve_host_wait(t)
            for key in want[i]:
                assert np.array_equal(bufs[i][1][key], want[i][key]), (i, key)
        t, _ = s.solve_host_begin(X, out=bufs[k][1])
        tickets.append((t, k))
    for t, i in tickets:
        s.solve_host_wait(t)
        for key in want[i]:
            assert np.array_equal(bufs[i][1][key], want[i][key]), (i, key)
    s.close()
