"""Development diagnostic: histogram of DDP iteration counts over the headline batch."""
import sys
import numpy as np
sys.path.insert(0, ".")
from trajectory_optimization_matrix_lie_groups_b200 import workloads
wl = workloads.se3_tracking_ms(B=16384)
s, x0 = wl.make_solver()
out = s.solve(x0, trajectories=False)
it = out["iters"].cpu().numpy()
h = np.bincount(it)
act = [(it >= k).sum() for k in range(it.max() + 2)]
print("hist", {k: int(v) for k, v in enumerate(h) if v})
print("active at launch k:", act)
print("sum active / (launches * B):", sum(act[:it.max() + 1]) / ((it.max() + 1) * it.size))
j = np.arange(it.size) % 12
for p in range(12):
    print(p, "mean iters", it[j == p].mean(), "max", it[j == p].max())
# warps with all lanes done per launch
w = it.reshape(-1, 32).max(axis=1)
print("warp-level active at launch k:", [(w >= k).sum() for k in range(it.max() + 2)])
