"""Oracle restatement of traoptlibrary's controllers (traopt_controller.py), loop by loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).

  * `ilqr_ss`    : iLQR_Tracking_SE3.fit (:1880-2013) and iLQR_Tracking_SO3.fit (:575-695)
  * `ilqr_ms`    : iLQR_Tracking_SE3_MS.fit (:2443-2639) and iLQR_Tracking_SO3_MS.fit (:1131-1325)
  * `al_ilqr_ms` : AL_iLQR_Tracking_SE3_MS.fit (:3218-3290), with the reference's 6-vs-7 tuple
                   unpack bug at :3236 fixed (SURVEY.md finding 6) — the only way it can run.

The reference appends to its history lists from the *callbacks* the scripts pass in
(benchmark_SE3_tracking.py:22-42); the solvers here append the same values at the same points, so
`J_hist`, `grad_hist`, `defect_hist` line up entry by entry with the shipped pickles.

Summation orders are part of the contract (SURVEY.md "Decision parity"): `J_opt = L.sum()` is
NumPy pairwise summation (:1935, :2507), `J_new` is a Python left-to-right sum over stages plus
the terminal term (:2094-2096), the gradient-norm sum runs from t = N-1 down to 0 (:2343-2349).
"""
import warnings
from dataclasses import dataclass, field

import numpy as np

from . import lie

STATUS_CONVERGED = 0      # gradient (and defect) tolerance met
STATUS_MAX_ITER = 1       # n_iterations exhausted
STATUS_NO_DESCENT = 2     # line search exhausted ("Couldn't find descent direction")


# ============================================================================================
# Group adapters: the few manif operations the controllers call directly
# ============================================================================================

class SE3Group:
    """Pose = 4x4 matrix, re-entering (quat, p) at every operation like the reference does."""
    pose_dim = 6

    @staticmethod
    def rminus(Ta, Tb):
        """manif `a - b` = Log(b^-1 a), [omega, v] (traopt_controller.py:2056-2058, 2680-2683)."""
        qa, pa = lie.se3_from_matrix(Ta)
        qb, pb = lie.se3_from_matrix(Tb)
        qi, pi = lie.se3_inverse(qb, pb)
        return lie.se3_log(*lie.se3_compose(qi, pi, qa, pa))

    @staticmethod
    def rplus(T, tau):
        """manif `X + tau` = X Exp(tau) (:2069, :2723)."""
        q, p = lie.se3_from_matrix(T)
        return lie.se3_to_matrix(*lie.se3_compose(q, p, *lie.se3_exp(tau)))

    @staticmethod
    def ms_step(T_next, d_q, alpha, T_f, T_f_new):
        """q_next Exp(alpha d_q) f(x,u).q^-1 f(x_new,u_new).q, left to right (:2713-2715)."""
        q, p = lie.se3_from_matrix(T_next)
        q, p = lie.se3_compose(q, p, *lie.se3_exp(alpha * d_q))
        qf, pf = lie.se3_from_matrix(T_f)
        q, p = lie.se3_compose(q, p, *lie.se3_inverse(qf, pf))
        q, p = lie.se3_compose(q, p, *lie.se3_from_matrix(T_f_new))
        return lie.se3_to_matrix(q, p)


class SO3Group:
    """Pose = unit quaternion (the reference keeps manif SO3 objects on this path)."""
    pose_dim = 3

    @staticmethod
    def rminus(qa, qb):
        return lie.so3_log(lie.quat_normalize(lie.quat_mul(lie.quat_conj(qb), qa)))  # :738, :1372

    @staticmethod
    def rplus(q, w):
        return lie.quat_normalize(lie.quat_mul(q, lie.so3_exp(w)))                   # :747

    @staticmethod
    def ms_step(q_next, d_q, alpha, q_f, q_f_new):
        q = lie.quat_normalize(lie.quat_mul(q_next, lie.so3_exp(alpha * d_q)))       # :1393
        q = lie.quat_normalize(lie.quat_mul(q, lie.quat_conj(q_f)))
        return lie.quat_normalize(lie.quat_mul(q, q_f_new))


@dataclass
class SolveResult:
    xs: list
    us: np.ndarray
    J_hist: list = field(default_factory=list)
    grad_hist: list = field(default_factory=list)
    defect_hist: list = field(default_factory=list)
    alpha_hist: list = field(default_factory=list)      # accepted step-size index per iteration, -1 = none
    mu_hist: list = field(default_factory=list)         # regulariser after each backward pass
    status: int = STATUS_MAX_ITER
    reg_exceeded: bool = False
    final_grad: float = None                            # MS: gradient norm that met the tolerance
    k: np.ndarray = None
    K: np.ndarray = None

    @property
    def iterations(self):
        return len(self.J_hist)


def is_pos_def(A):
    """traopt_utilis.py:320-329."""
    if np.array_equal(A, A.T):
        try:
            np.linalg.cholesky(A)
            return True
        except np.linalg.LinAlgError:
            return False
    return False


# ============================================================================================
# Shared pieces
# ============================================================================================

class _Regulariser:
    """Levenberg-Marquardt state, reset per fit (:1899-1900), persisting across stages/iterations."""

    def __init__(self, max_reg):
        self.mu = 1.0
        self.mu_min = 1e-6
        self.mu_max = max_reg
        self.delta_0 = 2.0
        self.delta = 2.0
        self.exceeded = False


def _trajectory_cost(cost, xs, us, N):
    """:2084-2096 / :2742-2754: Python `sum` (left to right from 0) + terminal."""
    J = 0
    for i in range(N):
        J = J + cost.l(xs[i], us[i], i)
    return J + cost.l(xs[N], None, N, terminal=True)


def _linearize(dynamics, cost, group, xs, us, N, with_defect):
    """:2098-2176 (SS) / :2823-2910 (MS)."""
    n, m = dynamics.state_size, dynamics.action_size
    F_x = np.empty((N, n, n))
    F_u = np.empty((N, n, m))
    L = np.empty(N + 1)
    L_x = np.empty((N + 1, n))
    L_u = np.empty((N, m))
    L_xx = np.empty((N + 1, n, n))
    L_ux = np.empty((N, m, n))
    L_uu = np.empty((N, m, m))
    d = np.empty((N, n)) if with_defect else None
    for i in range(N):
        x, u = xs[i], us[i]
        if with_defect:
            fx = dynamics.f(x, u, i)
            d[i] = np.concatenate((group.rminus(fx[0], xs[i + 1][0]), fx[1] - xs[i + 1][1]))  # :2882-2888
        F_x[i] = dynamics.f_x(x, u, i)
        F_u[i] = dynamics.f_u(x, u, i)
        L[i] = cost.l(x, u, i, terminal=False)
        L_x[i] = cost.l_x(x, u, i, terminal=False)
        L_u[i] = cost.l_u(x, u, i, terminal=False)
        L_xx[i] = cost.l_xx(x, u, i, terminal=False)
        L_ux[i] = cost.l_ux(x, u, i, terminal=False)
        L_uu[i] = cost.l_uu(x, u, i, terminal=False)
    x = xs[N]
    L[N] = cost.l(x, None, N, terminal=True)
    L_x[N] = cost.l_x(x, None, N, terminal=True)
    L_xx[N] = cost.l_xx(x, None, N, terminal=True)
    return d, F_x, F_u, L, L_x, L_u, L_xx, L_ux, L_uu


def _backward_pass(reg, n, N, d, F_x, F_u, L_x, L_u, L_xx, L_ux, L_uu):
    """:2178-2321 (SS, d=None) / :2912-3068 (MS).  Returns k, K, V_x[N+1], V_xx[N+1]."""
    m = F_u.shape[2]
    V_x = np.empty((N + 1, n))
    V_xx = np.empty((N + 1, n, n))
    V_x[N] = L_x[N]
    V_xx[N] = L_xx[N]
    k = np.empty((N, m))
    K = np.empty((N, m, n))
    for i in range(N - 1, -1, -1):
        f_x, f_u = F_x[i], F_u[i]
        Vx, Vxx = V_x[i + 1], V_xx[i + 1]
        while True:
            v = Vx if d is None else Vx + Vxx.dot(d[i])             # :3053-3054
            Q_x = L_x[i] + f_x.T.dot(v)
            Q_u = L_u[i] + f_u.T.dot(v)
            Q_xx = L_xx[i] + f_x.T.dot(Vxx).dot(f_x)
            r = reg.mu * np.eye(n)                                   # :2311 / :3058
            Q_ux = L_ux[i] + f_u.T.dot(Vxx + r).dot(f_x)
            Q_uu = L_uu[i] + f_u.T.dot(Vxx + r).dot(f_u)
            if not is_pos_def(Q_uu + Q_uu.T):                        # :2232 / :2977
                reg.delta = max(1.0, reg.delta) * reg.delta_0
                reg.mu = max(reg.mu_min, reg.mu * reg.delta)
                if reg.mu_max and reg.mu >= reg.mu_max:
                    warnings.warn("exceeded max regularization term")
                    reg.exceeded = True
                    break
            else:
                reg.delta = min(1.0, reg.delta) / reg.delta_0
                reg.mu *= reg.delta
                if reg.mu <= reg.mu_min:
                    reg.mu = 0.0
                break
        k[i] = -np.linalg.solve(Q_uu, Q_u)                           # :2249-2250
        K[i] = -np.linalg.solve(Q_uu, Q_ux)
        V_x[i] = Q_x + K[i].T.dot(Q_uu).dot(k[i])                    # :2253-2259
        V_x[i] += K[i].T.dot(Q_u) + Q_ux.T.dot(k[i])
        V_xx[i] = Q_xx + K[i].T.dot(Q_uu).dot(K[i])
        V_xx[i] += K[i].T.dot(Q_ux) + Q_ux.T.dot(K[i])
        V_xx[i] = 0.5 * (V_xx[i] + V_xx[i].T)
    return k, K, V_x, V_xx


def _alphas(n):
    return 1.1 ** (-np.arange(n) ** 2)                               # :1908 / :2472


# ============================================================================================
# Single shooting
# ============================================================================================

def ilqr_ss(dynamics, cost, group, N, x0, us_init, n_iterations=100, tol_grad_norm=1e-3,
            max_reg=1e10, rollout="nonlinear", n_alphas=13, verbose=False):
    """iLQR_Tracking_SE3.fit :1880-2013 (SO3 twin :575-695).

    Class defaults differ (SE3: rollout='linear', tol 1e-3; SO3: 'nonlinear', 1e-6) — pass them.
    """
    n, m = dynamics.state_size, dynamics.action_size
    pd = group.pose_dim
    reg = _Regulariser(max_reg)
    alphas = _alphas(n_alphas)
    us = np.array(us_init, dtype=float)
    xs = [None] * (N + 1)                                             # _init_rollout :2015-2028
    xs[0] = x0
    for i in range(N):
        xs[i + 1] = dynamics.f(xs[i], us[i], i)
    res = SolveResult(xs=xs, us=us)
    k = K = None

    for iteration in range(n_iterations):
        accepted = False
        _, F_x, F_u, L, L_x, L_u, L_xx, L_ux, L_uu = _linearize(dynamics, cost, group, xs, us, N, False)
        J_opt = L.sum()                                               # pairwise (:1935)

        p = L_x[N]                                                    # _gradient_wrt_control :2323-2349
        g_norm_sum = 0
        for t in range(N - 1, -1, -1):
            g = L_u[t] + np.matmul(F_u[t].T, p)
            p = L_x[t] + np.matmul(F_x[t].T, p)
            g_norm_sum = g_norm_sum + np.linalg.norm(g)
        grad = g_norm_sum / N
        res.grad_hist.append(grad)
        if grad < tol_grad_norm:                                      # :1939-1942, before the callback
            res.status = STATUS_CONVERGED
            break

        k, K, _, _ = _backward_pass(reg, n, N, None, F_x, F_u, L_x, L_u, L_xx, L_ux, L_uu)
        res.mu_hist.append(reg.mu)

        a_idx = -1
        for j, alpha in enumerate(alphas):                            # :1972-1990
            xs_new = [None] * (N + 1)                                 # _rollout :2030-2082
            us_new = np.zeros_like(us)
            xs_new[0] = [np.array(xs[0][0]), np.array(xs[0][1])]
            for i in range(N):
                q_new, xi_new = xs_new[i]
                q, xi = xs[i]
                xs_err = np.concatenate((group.rminus(q_new, q), xi_new - xi))
                us_err = alpha * k[i] + K[i].dot(xs_err)
                us_new[i] = us[i] + us_err
                if rollout == "linear":                               # :2066-2071
                    q_next, xi_next = xs[i + 1]
                    xs_new[i + 1] = [group.rplus(q_next, F_x[i, :pd, :] @ xs_err + F_u[i, :pd, :] @ us_err),
                                     xi_next + F_x[i, pd:, :] @ xs_err + F_u[i, pd:, :] @ us_err]
                else:                                                 # :2073-2080
                    xs_new[i + 1] = dynamics.f([q_new, xi_new], us_new[i], i)
            J_new = _trajectory_cost(cost, xs_new, us_new, N)
            if verbose:
                print(f"  it {iteration} alpha[{j}]={alpha:.3e} J_new={J_new!r} J_opt={J_opt!r}")
            if J_new < J_opt:
                J_opt, xs, us = J_new, xs_new, us_new
                accepted = True
                a_idx = j
                break

        res.J_hist.append(J_opt)                                      # the scripts' callback
        res.alpha_hist.append(a_idx)
        if not accepted:                                              # :2005-2007
            res.status = STATUS_NO_DESCENT
            break

    res.xs, res.us, res.k, res.K = xs, us, k, K
    res.reg_exceeded = reg.exceeded
    return res


# ============================================================================================
# Multiple shooting
# ============================================================================================

def _ms_rollout(dynamics, group, N, xs, us, k, K, d, F_x, F_u, alpha, mode):
    """iLQR_Tracking_SE3_MS._rollout :2641-2740 (SO3 twin :1327-1420)."""
    n = dynamics.state_size
    pd = group.pose_dim
    xs_new = [None] * (N + 1)
    us_new = np.zeros_like(us)
    xs_new[0] = [np.array(xs[0][0]), np.array(xs[0][1])]
    xs_errs = np.empty((N + 1, n))
    us_errs = np.empty((N, dynamics.action_size))
    for i in range(N):
        q_new, xi_new = xs_new[i]
        q, xi = xs[i]
        q_next, xi_next = xs[i + 1]
        xs_err = np.concatenate((group.rminus(q_new, q), xi_new - xi))
        us_err = alpha * k[i] + K[i].dot(xs_err)
        us_new[i] = us[i] + us_err
        xs_errs[i] = xs_err
        us_errs[i] = us_err
        d_q, d_xi = d[i, :pd], d[i, pd:]
        if mode == "nonlinear":                                       # :2697-2718
            fq_new, fxi_new = dynamics.f([q_new, xi_new], us_new[i], i)
            fq, fxi = dynamics.f([q, xi], us[i], i)
            xs_new[i + 1] = [group.ms_step(q_next, d_q, alpha, fq, fq_new),
                             xi_next + fxi_new - fxi + alpha * d_xi]
        else:                                                         # :2720-2726
            xs_new[i + 1] = [group.rplus(q_next, (F_x[i, :pd, :] @ xs_err + F_u[i, :pd, :] @ us_err) + alpha * d_q),
                             xi_next + F_x[i, pd:, :] @ xs_err + F_u[i, pd:, :] @ us_err + alpha * d_xi]
    q_new, xi_new = xs_new[N]
    q, xi = xs[N]
    xs_errs[N] = np.concatenate((group.rminus(q_new, q), xi_new - xi))
    return xs_new, us_new, xs_errs, us_errs


def _compute_defect(dynamics, group, N, xs, us):
    """:2790-2810."""
    d = np.empty((N, dynamics.state_size))
    for i in range(N):
        fx = dynamics.f(xs[i], us[i], i)
        d[i] = np.concatenate((group.rminus(fx[0], xs[i + 1][0]), fx[1] - xs[i + 1][1]))
    return d


def ilqr_ms(dynamics, cost, group, N, q_ref, xi_ref, x0, us_init, n_iterations=100,
            tol_grad_norm=1e-6, tol_d_norm=1e-6, max_reg=1e10, line_search=False,
            rollout="nonlinear", n_alphas=20, defect_kappa=1e-12, append_final_grad=False,
            verbose=False):
    """iLQR_Tracking_SE3_MS.fit :2443-2639 (SO3 twin :1131-1325: n_alphas=13, kappa=1e-14).

    `q_ref` must already be in the group adapter's pose type (4x4 for SE3, unit quat for SO3).
    `append_final_grad`: the SO3 twin appends the converged gradient norm to grad_hist itself (:1219).
    """
    n, m = dynamics.state_size, dynamics.action_size
    reg = _Regulariser(max_reg)
    alphas = _alphas(n_alphas)
    us = np.array(us_init, dtype=float)
    xs = [x0] + [[q_ref[i], np.array(xi_ref[i], dtype=float)] for i in range(1, N + 1)]  # :3123-3136
    res = SolveResult(xs=xs, us=us)
    k = K = None
    defect_mu0, defect_rho, defect_gamma = 10.0, 0.5, 0.05            # :2406-2410
    d_weight_prev = defect_mu0

    for iteration in range(n_iterations):
        accepted = False
        d, F_x, F_u, L, L_x, L_u, L_xx, L_ux, L_uu = _linearize(dynamics, cost, group, xs, us, N, True)
        d_norm = np.linalg.norm(d.reshape(-1), 2)                     # :2812-2821
        if iteration == 0:
            res.defect_hist.append(d_norm)
        J_opt = L.sum()

        k, K, V_x, V_xx = _backward_pass(reg, n, N, d, F_x, F_u, L_x, L_u, L_xx, L_ux, L_uu)
        res.mu_hist.append(reg.mu)

        g_norm_sum = 0                                                # :3070-3093
        for t in range(N - 1, -1, -1):
            g = L_u[t] + np.matmul(F_u[t].T, V_x[t + 1] + np.matmul(V_xx[t + 1].T, d[t]))
            g_norm_sum = g_norm_sum + np.linalg.norm(g)
        grad = g_norm_sum / N
        if grad < tol_grad_norm and d_norm < tol_d_norm:              # :2528-2532, before the callback
            res.status = STATUS_CONVERGED
            res.final_grad = grad
            if append_final_grad:
                res.grad_hist.append(grad)
            break

        a_idx = -1
        if line_search:                                               # :2549-2590
            _, _, xs_errs, us_errs = _ms_rollout(dynamics, group, N, xs, us, k, K, d, F_x, F_u, 1.0, "linear")
            c1 = 0.0                                                  # _expected_cost_change :2756-2769
            c2 = 0.0
            for i in range(N):
                c1 += L_x[i].T @ xs_errs[i] + L_u[i].T @ us_errs[i]
                c2 += xs_errs[i].T @ L_xx[i] @ xs_errs[i] + us_errs[i].T @ L_uu[i] @ us_errs[i] \
                    + 2 * us_errs[i].T @ L_ux[i] @ xs_errs[i]
            c1 += L_x[N].T @ xs_errs[N]
            c2 += xs_errs[N].T @ L_xx[N] @ xs_errs[N]
            if d_norm < defect_kappa:                                 # _update_defect_weight :2774-2788
                d_weight = d_weight_prev
            else:
                d_weight = max(defect_mu0, defect_mu0 + np.abs(c1 + 0.5 * c2) / ((1 - defect_rho) * d_norm))
            d_weight_prev = d_weight
            merit = J_opt + d_weight * d_norm
            for j, alpha in enumerate(alphas):
                xs_new, us_new, _, _ = _ms_rollout(dynamics, group, N, xs, us, k, K, d, F_x, F_u, alpha, rollout)
                J_new = _trajectory_cost(cost, xs_new, us_new, N)
                d_norm_new = np.linalg.norm(_compute_defect(dynamics, group, N, xs_new, us_new).reshape(-1), 2)
                J_exp = alpha * c1 + 0.5 * (alpha ** 2) * c2          # _scale_cost_change :2771-2772
                merit_new = J_new + d_weight * d_norm_new
                if merit_new - merit < defect_gamma * (J_exp - alpha * d_weight * d_norm):  # :2576
                    accepted = True
                    a_idx = j
                    break
        else:                                                         # :2592-2600
            xs_new, us_new, _, _ = _ms_rollout(dynamics, group, N, xs, us, k, K, d, F_x, F_u, 1, rollout)
            J_new = _trajectory_cost(cost, xs_new, us_new, N)
            d_norm_new = np.linalg.norm(_compute_defect(dynamics, group, N, xs_new, us_new).reshape(-1), 2)
            accepted = True
            a_idx = 0

        if accepted:                                                  # :2609-2612
            J_opt, xs, us = J_new, xs_new, us_new
        if verbose:
            print(f"  it {iteration} J={J_opt!r} d={d_norm_new:.3e} grad={grad:.3e} a={a_idx} mu={reg.mu}")
        res.J_hist.append(J_opt)                                      # the scripts' callback
        res.grad_hist.append(grad)
        res.defect_hist.append(d_norm_new)
        res.alpha_hist.append(a_idx)
        if not accepted:                                              # :2631-2633
            res.status = STATUS_NO_DESCENT
            break

    res.xs, res.us, res.k, res.K = xs, us, k, K
    res.reg_exceeded = reg.exceeded
    return res


# ============================================================================================
# Augmented Lagrangian outer loop
# ============================================================================================

@dataclass
class ALResult:
    inner: SolveResult
    lmbd: np.ndarray
    Imu: np.ndarray
    mu: float
    outer_iterations: int
    constr_converged: bool
    violation_hist: list
    inner_iters_hist: list


def al_ilqr_ms(dynamics, al_cost, constraints, group, N, q_ref, xi_ref, x0, us_init,
               n_al_iters=100, n_ilqr_iters=200, tol_constr=1e-2, mu_scale=10.0, max_reg=1e10,
               line_search=False, verbose=False):
    """AL_iLQR_Tracking_SE3_MS.fit :3218-3267 and _al_update_param :3270-3290."""
    c = constraints.constr_size
    mu0, mu_max = 1e-2, 1e8                                           # :3182-3184
    al_cost.lmbd = np.zeros((N + 1, c))                               # :3223, _al_inital_param
    al_cost.Imu = np.tile(mu0 * np.eye(c), (N + 1, 1, 1))
    al_cost.mu = mu0
    violation_hist, inner_hist = [], []
    converged = False
    inner = None
    it = 0
    for it in range(n_al_iters):
        inner = ilqr_ms(dynamics, al_cost, group, N, q_ref, xi_ref, x0, us_init,    # cold start :3237
                        n_iterations=n_ilqr_iters, tol_grad_norm=1e-6, tol_d_norm=1e-6,
                        max_reg=max_reg, line_search=line_search, rollout="nonlinear")
        g = np.array([constraints.g(inner.xs[i], inner.us[i], i) for i in range(N)]
                     + [constraints.g(inner.xs[N], None, N, terminal=True)])        # :3242-3248
        violation_hist.append(float(np.max(g)))
        inner_hist.append(inner.iterations)
        if verbose:
            print(f"AL it {it}: inner {inner.iterations} max g {np.max(g):.4e} mu {al_cost.mu}")
        if np.max(g) < tol_constr:                                    # :3250
            converged = True
            break
        lmbd, Imu = al_cost.lmbd, al_cost.Imu                         # _al_update_param
        mu_new = min(al_cost.mu * mu_scale, mu_max)
        lmbd_new = lmbd.copy()
        Imu_new = Imu.copy()
        for i in range(N + 1):
            lmbd_new[i] = np.clip(lmbd[i] + Imu[i] @ g[i], a_min=0.0, a_max=None)
            Imu_new[i] = np.diag(np.where((g[i] < 0.0) & (lmbd_new[i] == 0.0), 0.0, mu_new))
        al_cost.lmbd, al_cost.Imu, al_cost.mu = lmbd_new, Imu_new, mu_new
    return ALResult(inner, al_cost.lmbd, al_cost.Imu, al_cost.mu, it + 1, converged,
                    violation_hist, inner_hist)
