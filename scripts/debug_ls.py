"""Development diagnostic: drone multiple-shooting merit line search, device vs oracle, per iteration."""
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gpu_common as gc  # noqa: E402
from oracle import problems, solvers  # noqa: E402

warnings.simplefilter("ignore")
name, horizon, n_iter, B = "drone_n150", 30, 8, 6
g = problems.load_golden(name)
kind = str(g["kind"])
rng = np.random.default_rng(11)
for it_cap in (3, 8):
    s, x0, N = gc.make_solver(g, "ms", B, horizon=horizon, max_iters=it_cap, tol_grad_norm=1e-12, line_search=True)
    X0 = gc.perturbed_x0(x0, B, scale=0.05)
    rng = np.random.default_rng(11)
    us0 = 0.05 * rng.standard_normal((B, N, s.NU))
    us0[0] = 0.0
    out = s.solve(X0, us0)
    hist = {k: v.cpu().numpy() for k, v in s.export_hist().items()}
    tab = s.debug_linesearch().cpu().numpy()
    for b in range(B):
        it = int(out["iters"][b])
        print(f"cap {it_cap} b={b} gpu iters {it} status {int(out['status'][b])} alpha {hist['alpha_hist'][b, :it].tolist()}")
        print("    J", ["%.17g" % j for j in hist["J_hist"][b, :it]])
        print("    table J_new", ["%.15g" % v for v in tab[:20, b]])
        print("    table d_new", ["%.3g" % v for v in tab[20:40, b]])
        print("    c1 c2 dw merit", ["%.17g" % v for v in tab[40:44, b]])
