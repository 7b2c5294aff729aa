"""Small solves of every kind / method in one process: a quick everything-launches check (and the workload to put
under compute-sanitizer where that tool is available; it is closed on the development pool).
Covers: linearise, both Riccati sweeps (2-warp TMA-staged and one-warp), the TMA-staged rollout, the generic
rollouts, SS and merit line searches, the AL outer loop, compaction, per-problem references and horizons, exports."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gpu_common as gc  # noqa: E402
from oracle import problems  # noqa: E402

N, B = 12, 40     # 40 problems: one full group of 32 and a ragged one


def run(name, method, **kw):
    g = problems.load_golden(name)
    horizons = kw.pop("horizons", None)
    s, x0, n = gc.make_solver(g, method, B, horizon=N, max_iters=4, tol_grad_norm=1e-12, **kw)
    s.set_compaction(8, 1.5)
    if horizons is not None:
        s.set_horizons(horizons)
    X0 = gc.perturbed_x0(x0, B, scale=0.02)
    out = s.solve(X0)
    s.export_hist()
    s.debug_linearize()
    print(name, method, kw, "iters", int(out["iters"].max()), "J", float(out["J"].max()), flush=True)


for name in ("se3_n120", "drone_n150", "rigid_n120", "so3_n249", "pendulum_n80"):
    run(name, "ms")
    run(name, "ss")
run("se3_n120", "ms", line_search=True)
run("se3_n120", "ms", rollout="linear")
run("se3_n120", "ms", horizons=np.arange(B) % N + 1)
run("se3_n120", "ss", horizons=np.arange(B) % N + 1)
run("so3_n249", "ms", horizons=np.arange(B) % N + 1)
run("se3_n120", "al_ms", lb=np.full(6, -0.5), ub=np.full(6, 0.5), n_al_iters=3)
print("done")
