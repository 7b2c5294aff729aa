"""Development probe: wall time of each DDP iteration of one headline batch (begin + iterate(1) x n)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from trajectory_optimization_matrix_lie_groups_b200 import workloads

wl = workloads.CONFIGS[3](B=16384)
s, x0 = wl.make_solver(B=16384, device=torch.device("cuda", 0))
for rep in range(2):
    s.begin(x0)
    torch.cuda.synchronize()
    ts = []
    t_all = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        act = s.iterate(1)
        torch.cuda.synchronize()
        ts.append((1e3 * (time.perf_counter() - t0), act))
        if act == 0:
            break
    print("total %.1f ms" % (1e3 * (time.perf_counter() - t_all)))
    print(" ".join("%.1f(%d)" % t for t in ts))

# per-phase device time of each iteration (profiling mode: serial order, events around every launch)
s.set_profiling(True)
s.begin(x0)
s.phase_times(reset=True)
while True:
    act = s.iterate(1)
    ph = s.phase_times(reset=True)
    print(act, " ".join("%s %.2f" % (k, v[0]) for k, v in ph.items()))
    if act == 0:
        break
