"""Several batches in flight on one GPU.

A batch's solve is a host-driven loop of latency-bound kernels whose trip count is set by its slowest
problem (16..27 DDP iterations on the headline workload, 20 on average): in the last third of the
launches most problems have already converged and most CTAs exit at once, leaving the SMs idle.
`PipelinedSolver` keeps `depth` independent `BatchSolver`s, each with its own CUDA stream and host
thread, and hands successive batches to them round-robin: the tail of one batch then shares the GPU
with the head of the next, and (for host buffers) the D2H copy of one batch overlaps the compute of
the next.  Results are bit-identical to solving the batches one after the other: the lanes share
nothing but the device.

The reference's counterpart is its joblib pool (visualization/perturb_all_compute.py:240-250), which
also keeps every core busy with whichever job is ready.
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import torch


class PipelinedSolver:
    def __init__(self, make_solver, depth=2, device=None):
        """make_solver() -> a configured BatchSolver (called `depth` times, on `device`)."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        self.solvers = [make_solver() for _ in range(depth)]
        for sv in self.solvers:
            sv.set_sweep(0, depth)        # the lanes share the SMs: each one's sweep picks its CTA shape accordingly
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
        self.pools = [ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"trajopt-lane{i}") for i in range(depth)]
        # host-buffer solves: a second thread per lane waits for the device->host copies of a finished solve while the
        # lane's own thread already runs the next one (trajopt_solve_host_begin / _wait)
        self.drains = [ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"trajopt-drain{i}") for i in range(depth)]
        self.active_depth = depth     # lanes in use (<= depth): submit() goes round-robin over the first `active_depth` lanes
        self._next = 0
        self._lock = threading.Lock()

    def _run(self, lane, fn, args, kwargs):
        with torch.cuda.device(self.device), torch.cuda.stream(self.streams[lane]):
            out = fn(self.solvers[lane], *args, **kwargs)
            self.streams[lane].synchronize()
        return out

    def submit(self, x0, us_init=None, trajectories=True, host=False, out=None):
        """Queue one batch; returns a Future of the result dict of BatchSolver.solve / solve_host.

        The result tensors of a lane are owned by that lane's solver call; they stay valid until the
        caller drops them (device results are fresh tensors; host results are `out` or fresh arrays).
        """
        with self._lock:
            lane = self._next % max(1, min(self.active_depth, self.depth))
            self._next = (lane + 1) % max(1, min(self.active_depth, self.depth))
        if host:
            fn = lambda s, *a, **k: s.solve_host_begin(*a, **k)      # noqa: E731
            begun = self.pools[lane].submit(self._run, lane, fn, (x0, us_init), dict(trajectories=trajectories, out=out))

            def drain():
                ticket, res = begun.result()
                with torch.cuda.device(self.device):
                    self.solvers[lane].solve_host_wait(ticket)
                return res
            return self.drains[lane].submit(drain)
        fn = lambda s, *a, **k: s.solve(*a, **k)               # noqa: E731
        return self.pools[lane].submit(self._run, lane, fn, (x0, us_init), dict(trajectories=trajectories))

    def map(self, batches, **kw):
        """Solve an iterable of x0 batches, `depth` in flight; yields results in submission order."""
        futs = [self.submit(x0, **kw) for x0 in batches]
        for f in futs:
            yield f.result()

    def close(self):
        for p in self.pools + self.drains:
            p.shutdown(wait=True)
        for s in self.solvers:
            s.close()
        self.solvers = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
