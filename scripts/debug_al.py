"""Development diagnostic: augmented-Lagrangian config (cfg 4) on a small batch, per outer iteration."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from trajectory_optimization_matrix_lie_groups_b200 import workloads

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1400
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
wl = workloads.se3_tracking_al_ms(B=B, N=N, frac=frac)
s, x0 = wl.make_solver()
s.begin(x0)
t0 = time.time()
for outer in range(100):
    t1 = time.time()
    act = s.iterate(1)
    torch.cuda.synchronize()
    out = s.export(trajectories=False)
    al = s.export_al()
    it = out["iters"].cpu().numpy()
    st = out["status"].cpu().numpy()
    viol = al["violation"].cpu().numpy()
    oi = al["outer_iters"].cpu().numpy()
    print(f"outer {outer}: active {act} dt {time.time() - t1:.2f}s inner iters min/mean/max {it.min()}/{it.mean():.1f}/{it.max()} "
          f"status {np.bincount(st & 15, minlength=4).tolist()} viol max {viol.max():.3e} min {viol.min():.3e} mu {float(al['mu'][0]):.1e} "
          f"outer_iters max {oi.max()}")
    if act == 0:
        break
print("total", time.time() - t0)
