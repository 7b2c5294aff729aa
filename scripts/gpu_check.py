"""Development diagnostic (not a test): run the native solver against the oracle / goldens and
print the differences.  Usage on the GPU box: python scripts/gpu_check.py [quick]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import lie, problems, solvers  # noqa: E402
import gpu_common as gc  # noqa: E402
from trajectory_optimization_matrix_lie_groups_b200 import lie_op  # noqa: E402


def check_lie():
    rng = np.random.default_rng(0)
    n = 64
    w = rng.standard_normal((n, 3)) * np.concatenate((np.full(16, 1e-7), np.full(16, 1e-3), np.full(16, 0.3), np.full(16, 0.9)))[:, None]
    tau = np.concatenate((w, rng.standard_normal((n, 3))), axis=1)
    for name, fn, x in (
        ("so3_exp", lie.so3_exp, w), ("so3_jr", lambda a: lie.so3_jr(a).ravel(), w),
        ("so3_jr_inv", lambda a: lie.so3_jr_inv(a).ravel(), w), ("so3_jl", lambda a: lie.so3_jl(a).ravel(), w),
        ("so3_jl_inv", lambda a: lie.so3_jl_inv(a).ravel(), w),
        ("se3_exp", lambda a: np.concatenate(lie.se3_exp(a)), tau), ("se3_Q", lambda a: lie.se3_Q(a[:3], a[3:]).ravel(), tau),
        ("se3_jr", lambda a: lie.se3_jr(a).ravel(), tau), ("se3_jr_inv", lambda a: lie.se3_jr_inv(a).ravel(), tau),
    ):
        got = lie_op(name, x).cpu().numpy()
        ref = np.stack([fn(r) for r in x])
        print(f"lie {name:12s} max abs err {np.max(np.abs(got - ref)):.3e}")
    q = np.stack([lie.so3_exp(r) for r in w])
    got = lie_op("so3_log", q).cpu().numpy()
    print(f"lie so3_log      max abs err {np.max(np.abs(got - w)):.3e}")
    qp = np.stack([np.concatenate(lie.se3_exp(r)) for r in tau])
    got = lie_op("se3_log", qp).cpu().numpy()
    ref = np.stack([lie.se3_log(r[:4], r[4:]) for r in qp])
    print(f"lie se3_log      max abs err {np.max(np.abs(got - ref)):.3e}")
    got = lie_op("se3_adj", qp).cpu().numpy()
    ref = np.stack([lie.se3_adj(r[:4], r[4:]).ravel() for r in qp])
    print(f"lie se3_adj      max abs err {np.max(np.abs(got - ref)):.3e}")
    ab = np.concatenate((qp, np.roll(qp, 1, axis=0)), axis=1)
    for name, fn in (("se3_compose", lambda a: np.concatenate(lie.se3_compose(a[:4], a[4:7], a[7:11], a[11:]))),
                     ("se3_rminus", lambda a: lie.se3_log(*lie.se3_compose(*lie.se3_inverse(a[7:11], a[11:]), a[:4], a[4:7]))),
                     ("se3_lminus", lambda a: lie.se3_log(*lie.se3_compose(a[:4], a[4:7], *lie.se3_inverse(a[7:11], a[11:]))))):
        got = lie_op(name, ab).cpu().numpy()
        ref = np.stack([fn(r) for r in ab])
        print(f"lie {name:12s} max abs err {np.max(np.abs(got - ref)):.3e}")


def check_linearize(name, method, horizon=None):
    g = problems.load_golden(name)
    kind = str(g["kind"])
    B = 5
    s, x0, N = gc.make_solver(g, method, B, horizon=horizon, max_iters=3, tol_grad_norm=1e-12)
    X0 = gc.perturbed_x0(x0, B)
    rng = np.random.default_rng(1)
    us = 0.1 * rng.standard_normal((B, N, s.NU))
    s.begin(X0, us)
    out = {k: v.cpu().numpy() for k, v in s.debug_linearize().items()}
    dyn, cost, group, q_ref, xi_ref, _, _ = problems.from_golden(g, horizon)
    for b in (0, B - 1):
        x0o = gc.oracle_state(kind, X0[b])
        if method == "ms":
            xs = [x0o] + [[q_ref[i], np.array(xi_ref[i], dtype=float)] for i in range(1, N + 1)]
        else:
            xs = [x0o]
            for i in range(N):
                xs.append(dyn.f(xs[i], us[b, i], i))
        d, F_x, F_u, L, L_x, L_u, L_xx, L_ux, L_uu = solvers._linearize(dyn, cost, group, xs, us[b], N, method == "ms")
        ref = {"F_x": F_x, "F_u": F_u, "L": L, "L_x": L_x, "L_u": L_u, "L_xx": L_xx}
        if d is not None:
            ref["d"] = d
        for k, v in ref.items():
            err = np.max(np.abs(out[k][b] - v))
            scale = max(1.0, np.max(np.abs(v)))
            print(f"lin {name} {method} b={b} {k:5s} max abs err {err:.3e} (scale {scale:.3e})")


def check_solve(name, method, iters_expected=None, horizon=None, max_iters=200, **kw):
    g = problems.load_golden(name)
    kind = str(g["kind"])
    B = 3
    s, x0, N = gc.make_solver(g, method, B, horizon=horizon, max_iters=max_iters, tol_grad_norm=1e-12, **kw)
    X0 = gc.perturbed_x0(x0, B, scale=0.01)
    t0 = time.time()
    out = s.solve(X0)
    torch.cuda.synchronize()
    dt = time.time() - t0
    hist = {k: v.cpu().numpy() for k, v in s.export_hist().items()}
    it = out["iters"].cpu().numpy()
    st = out["status"].cpu().numpy()
    print(f"solve {name} {method}: iters {it} status {st} J {out['J'].cpu().numpy()} wall {dt:.3f}s")
    tag = method
    if tag + "_J_hist" in g and horizon is None:
        Jg = g[tag + "_J_hist"]
        n = min(len(Jg), int(it[0]))
        Jh = hist["J_hist"][0, :n]
        rel = np.abs(Jh - Jg[:n]) / np.abs(Jg[:n])
        print(f"   golden iters {len(Jg)}; J_hist max rel err over first {n}: {rel.max():.3e}; per-iter {np.array2string(rel[:8], precision=2)}")
        print(f"   alpha_hist {hist['alpha_hist'][0, :n].tolist()}")
        xs = out["xs"].cpu().numpy()[0]
        us = out["us"].cpu().numpy()[0]
        print(f"   us max abs err {np.max(np.abs(us - g[tag + '_us'])):.3e} (scale {np.max(np.abs(g[tag + '_us'])):.3e})")
        from trajectory_optimization_matrix_lie_groups_b200 import layout
        if kind == "so3":
            P = layout.quat_to_rot(xs[:, :4]); V = xs[:, 4:]
        else:
            P = layout.rows_to_se3(xs[:, :7]); V = xs[:, 7:]
        print(f"   xs pose err {np.max(np.abs(P - g[tag + '_xs_q'])):.3e} vel err {np.max(np.abs(V - g[tag + '_xs_xi'])):.3e}")
        if tag == "ms":
            print(f"   grad_hist tail gpu {hist['grad_hist'][0, max(0, n - 3):n + 1]} golden {g['ms_grad_hist'][-3:]}")
            print(f"   defect_hist gpu {hist['defect_hist'][0, :3]} golden {g['ms_defect_hist'][:3]}")
    return s, out


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print(torch.cuda.get_device_name(0))
    check_lie()
    check_linearize("se3_n120", "ms")
    check_linearize("se3_n120", "ss")
    check_linearize("so3_n249", "ms", horizon=30)
    check_linearize("drone_n150", "ms", horizon=30)
    check_solve("se3_n120", "ms")
    check_solve("se3_n120", "ss", rollout="nonlinear")
    check_solve("so3_n249", "ms")
    check_solve("so3_n249", "ss", max_iters=50)
    check_solve("drone_n150", "ms")
    check_solve("drone_n150", "ss")
    if not quick:
        check_solve("se3_n955_r1e-5", "ms")
        check_solve("se3_n955_r1e-4", "ms")
