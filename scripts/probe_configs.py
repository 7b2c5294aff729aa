"""Development probe (GPU): per-problem iteration counts / costs of the BASELINE configs, written to
gpurun_out/probe_configs.npz so that the oracle fixtures (tests/golden/make_cfg_fixtures.py) can be chosen to
cover the problems that set each batch's trip count."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from trajectory_optimization_matrix_lie_groups_b200 import workloads

os.makedirs("gpurun_out", exist_ok=True)
res = {}


def run(tag, wl, hist=False):
    s, x0 = wl.make_solver()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = s.solve(x0, trajectories=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it = out["iters"].cpu().numpy()
    st = out["status"].cpu().numpy()
    res[tag + "_iters"], res[tag + "_status"], res[tag + "_J"] = it, st, out["J"].cpu().numpy()
    res[tag + "_grad"] = out["grad"].cpu().numpy()
    print(tag, f"{dt:.3f}s", "iters hist", {int(k): int(v) for k, v in enumerate(np.bincount(it)) if v}, "status", np.unique(st & 15, return_counts=True))
    if hist:
        h = s.export_hist()
        for k, v in h.items():
            res[tag + "_" + k] = v.cpu().numpy()[:8]
    if wl.method == "al_ms":
        al = s.export_al()
        res[tag + "_outer"] = al["outer_iters"].cpu().numpy()
        res[tag + "_viol"] = al["violation"].cpu().numpy()
        res[tag + "_almu"] = al["mu"].cpu().numpy()
        print(tag, "outer hist", {int(k): int(v) for k, v in enumerate(np.bincount(res[tag + "_outer"])) if v})
    s.close()
    return dt


run("cfg3", workloads.se3_tracking_ms(B=16384))
run("cfg1", workloads.se3_tracking_ss(B=12), hist=True)
run("cfg1b1", workloads.se3_tracking_ss(B=1), hist=True)
run("cfg4", workloads.se3_tracking_al_ms(B=2048), hist=True)
run("cfg4_nom3", workloads.se3_tracking_al_ms(B=12), hist=True)
run("cfg5", workloads.drone_racing_ms(B=131072))
run("cfg2", workloads.so3_tracking_ms(B=1024))
np.savez_compressed("gpurun_out/probe_configs.npz", **res)
