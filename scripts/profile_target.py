"""Target of the `ncu --set full` captures: two DDP iterations of one batch of the headline workload, nothing else.

    python scripts/profile_target.py 16384     # two-warp sweep; run with TRAJOPT_OVERLAP=0 for whole-horizon launches
    python scripts/profile_target.py 2048      # a 2048-problem shard: six-warp sweep
    python scripts/profile_target.py 1024 2    # BASELINE config 2 (SO3 x 1024): the one-warp sweep
"""
import sys

import torch

sys.path.insert(0, ".")
from trajectory_optimization_matrix_lie_groups_b200 import workloads

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = workloads.CONFIGS[cfg](B=B)
s, x0 = wl.make_solver(B=B, device=torch.device("cuda", 0))
s.begin(torch.as_tensor(x0, device="cuda"))
act = s.iterate(2)
torch.cuda.synchronize()
print(B, "running after 2 iterations:", act)
s.close()
