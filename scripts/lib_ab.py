"""Development probe (GPU): one batch alone under the library / environment given by the caller, with a digest of every
export so that runs under different builds (TRAJOPT_LIB=...) or switches (TRAJOPT_B3_GROUPS, TRAJOPT_SWEEP ...) can be
compared for bit-identity.

    TRAJOPT_LIB=.../libtrajopt_b200_x.so python scripts/lib_ab.py [config] [batch] [sweep variant]
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from trajectory_optimization_matrix_lie_groups_b200 import workloads

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = workloads.CONFIGS[cfg](B=B)
s, x0 = wl.make_solver(B=B, device=torch.device("cuda", 0))
x0d = torch.as_tensor(x0, device="cuda")
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
h = hashlib.sha1()
for k in sorted(res):
    h.update(np.ascontiguousarray(res[k]).tobytes())
s.set_profiling(True)
s.phase_times(reset=True)
s.solve(x0d, trajectories=False)
torch.cuda.synchronize()
ph = s.phase_times(reset=True)
env = {k: v for k, v in os.environ.items() if k.startswith("TRAJOPT_")}
print(f"cfg {cfg} B {B} sweep {variant} env {env}: solve {min(ts):.1f} ms (iters max {int(res['iters'].max())}) digest {h.hexdigest()[:12]}; "
      + ", ".join(f"{k} {v[0]:.1f} ms / {v[1]}" for k, v in ph.items()))
