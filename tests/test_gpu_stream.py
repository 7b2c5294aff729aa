"""Continuous batching (`trajopt_solve_stream`): M problems through B slots, a finished slot is refilled before the
next iteration.  Every problem's result must be what `solve` gives for the same x0 — bit for bit, since a problem's
arithmetic does not depend on its slot or its neighbours — whatever the order in which problems finish."""
import numpy as np
import pytest
import torch

import gpu_common as gc
from oracle import problems

pytestmark = pytest.mark.gpu


def _batched(s, X0, B):
    """the same problems, B at a time, through solve()"""
    outs = []
    for lo in range(0, X0.shape[0], B):
        chunk = X0[lo:lo + B]
        n = chunk.shape[0]
        if n < B:
            chunk = np.concatenate((chunk, np.repeat(chunk[-1:], B - n, axis=0)))
        o = s.solve(chunk)
        outs.append({k: (None if v is None else v[:n].clone()) for k, v in o.items()})
    return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}


@pytest.mark.parametrize("name,method,kw,M", [
    ("se3_n120", "ms", {}, 200), ("se3_n120", "ss", {}, 150), ("so3_n249", "ms", {}, 200), ("drone_n150", "ms", {"line_search": True}, 150),
    ("se3_n120", "ms", {"rollout": "linear"}, 100), ("pendulum_n80", "ms", {}, 100), ("se3_n120", "ms", {}, 20),
])
def test_stream_matches_batched_solves(name, method, kw, M):
    g = problems.load_golden(name)
    B = 64
    s, x0, N = gc.make_solver(g, method, B, horizon=40, max_iters=30, tol_grad_norm=1e-9, **kw)
    # a spread of difficulties, so that problems finish after different numbers of iterations
    X0 = gc.perturbed_x0(x0, M, scale=0.02)
    X0[::3] = gc.perturbed_x0(x0, M, seed=7, scale=0.3)[::3]
    ref = _batched(s, X0, B)
    out = s.solve_stream(X0)
    it = ref["iters"].cpu().numpy()
    assert it.min() < it.max()                      # the refill path is really exercised
    for k in ("iters", "status", "J", "grad", "defect", "us", "xs"):
        assert torch.equal(out[k], ref[k]), k
    host = s.solve_stream_host(X0)                  # host buffers, rows copied out while the solve goes on
    for k in ("iters", "status", "J", "grad", "defect", "us", "xs"):
        assert np.array_equal(host[k], ref[k].cpu().numpy()), k
    # and again with the same handle: no state leaks from one stream to the next, or into plain solves
    out2 = s.solve_stream(X0[: M // 2], trajectories=False)
    assert torch.equal(out2["J"], ref["J"][: M // 2]) and out2["xs"] is None
    again = s.solve(X0[:B] if M >= B else np.concatenate((X0, np.repeat(X0[-1:], B - M, axis=0))))
    n = min(M, B)
    assert torch.equal(again["J"][:n], ref["J"][:n])


def test_stream_rejects_what_belongs_to_a_batch():
    g = problems.load_golden("se3_n120")
    s, x0, N = gc.make_solver(g, "ms", 8, horizon=20, max_iters=5)
    s.set_horizons([20, 10, 20, 20, 5, 20, 20, 20])
    with pytest.raises(RuntimeError):
        s.solve_stream(gc.perturbed_x0(x0, 16))
    s.set_horizons(None)
    assert s.solve_stream(gc.perturbed_x0(x0, 16))["J"].shape == (16,)
    assert s.solve_stream(np.zeros((0, s.NS)))["J"].shape == (0,)


@pytest.mark.parametrize("pinned", [True, False])
def test_solve_host_sends_finished_rows_early(pinned):
    """trajopt_solve_host copies the rows of the problems that have stopped while the stragglers still iterate, and
    the stragglers' rows at the end (by a kernel writing pinned memory, or run by run for pageable memory): same bytes
    as the device path either way."""
    g = problems.load_golden("se3_n120")
    B = 96
    s, x0, N = gc.make_solver(g, "ms", B, horizon=40, max_iters=60, tol_grad_norm=1e-6)
    X0 = gc.perturbed_x0(x0, B, scale=0.01)
    X0[5::8] = gc.perturbed_x0(x0, B, seed=7, scale=0.5)[5::8]      # one straggler in eight
    ref = s.solve(X0)
    it = ref["iters"].cpu().numpy()
    still = [int((it > k).sum()) for k in range(int(it.max()))]
    assert any(0 < n <= B // 4 for n in still), still                   # some iteration ends with 0 < running <= B/4: early path
    out = None
    if pinned:
        mk = lambda *shape, dtype=torch.float64: torch.full(shape, -7, dtype=dtype).pin_memory().numpy()   # noqa: E731
        out = {"J": mk(B), "iters": mk(B, dtype=torch.int32), "status": mk(B, dtype=torch.int32), "grad": mk(B), "defect": mk(B),
               "xs": mk(B, N + 1, s.NS), "us": mk(B, N, s.NU)}
    host = s.solve_host(X0, out=out)
    for k in ("iters", "status", "J", "grad", "defect", "us", "xs"):
        assert np.array_equal(host[k], ref[k].cpu().numpy()), k
