"""Development probe (GPU): the backward-sweep mappings side by side.

    python scripts/sweep_ab.py [config] [batch]

Solves one batch alone with the sweep forced to two-, four-, six-warp CTAs and chosen automatically;
prints the solve time, the per-phase device time of a profiled solve, and checks that every result is bit-identical."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from trajectory_optimization_matrix_lie_groups_b200 import workloads

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
wl = workloads.CONFIGS[cfg](B=B)
s, x0 = wl.make_solver(B=B, device=torch.device("cuda", 0))
x0d = torch.as_tensor(x0, device="cuda")
ref = None
for variant in (2, 4, 6, 0):
    s.set_sweep(variant, 1)
    s.solve(x0d, trajectories=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = s.solve(x0d)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res.update({k: v.cpu().numpy() for k, v in s.export_hist().items()})
    if ref is None:
        ref = res
    same = all(np.array_equal(ref[k], res[k]) for k in ref)
    s.set_profiling(True)
    s.phase_times(reset=True)
    s.solve(x0d, trajectories=False)
    torch.cuda.synchronize()
    ph = s.phase_times(reset=True)
    s.set_profiling(False)
    print(f"cfg {cfg} B {B} sweep variant {variant}: solve {min(ts):.1f} ms (iters max {int(res['iters'].max())}), bit-identical to variant 2: {same}; "
          + ", ".join(f"{k} {v[0]:.1f} ms / {v[1]}" for k, v in ph.items()))
