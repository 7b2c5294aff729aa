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
