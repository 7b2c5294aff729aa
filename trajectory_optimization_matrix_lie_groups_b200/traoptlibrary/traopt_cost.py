"""Cost classes with the reference's interface (traoptlibrary/traopt_cost.py).

`SE3TrackingQuadraticGaussNewtonCost` (:570-867), `SO3TrackingQuadraticGaussNewtonCost` (:280-564)
and `ALConstrainedCost` (:1173-1320) carry the weights / reference / multipliers that the
controllers pack into the native solver.  `l`, `l_x`, `l_u`, `l_xx`, `l_ux`, `l_uu` and `_err`
evaluate through the CUDA library (trajopt_debug_stage); the augmented-Lagrangian terms of the
box constraint are added on the host from the closed forms of traopt_cost.py:1236-1320.
"""
import numpy as np

from . import _native


class BaseCost:
    """Instantaneous cost interface (traopt_cost.py:14-110)."""

    def l(self, x, u, i, terminal=False):
        raise NotImplementedError

    def l_x(self, x, u, i, terminal=False):
        raise NotImplementedError

    def l_u(self, x, u, i, terminal=False):
        raise NotImplementedError

    def l_xx(self, x, u, i, terminal=False):
        raise NotImplementedError

    def l_ux(self, x, u, i, terminal=False):
        raise NotImplementedError

    def l_uu(self, x, u, i, terminal=False):
        raise NotImplementedError


class _PlaceholderCost:
    def __init__(self, nx, nu):
        self.Q = np.eye(nx)
        self.P = np.eye(nx)
        self.R = np.eye(nu)


def _placeholder_cost(nx, nu):
    return _PlaceholderCost(nx, nu)


class _PlaceholderDynamics:
    def __init__(self, kind):
        self.J = np.eye(3) if kind in ("so3", "pendulum") else np.eye(6)
        self.Ib = np.eye(3)
        self.m = 1.0
        self.dt = 1.0


class _TrackingCost(BaseCost):
    KIND = None

    def _init_common(self, Q, R, P, q_ref, xi_ref, state_size, action_size):
        self._state_size = state_size[0] + state_size[1]
        self._error_state_size = state_size[0]
        self._vel_state_size = state_size[1]
        self._action_size = action_size
        self._Q, self._R, self._P = (np.asarray(a, dtype=float) for a in (Q, R, P))
        self._q_ref = q_ref
        self._xi_ref = xi_ref
        self._evaluator = None

    state_size = property(lambda self: self._state_size)
    error_state_size = property(lambda self: self._error_state_size)
    vel_state_size = property(lambda self: self._vel_state_size)
    action_size = property(lambda self: self._action_size)
    Q = property(lambda self: self._Q)
    R = property(lambda self: self._R)
    P = property(lambda self: self._P)
    q_ref = property(lambda self: self._q_ref)
    xi_ref = property(lambda self: self._xi_ref)

    def _kind(self):
        return "drone" if (self.KIND == "se3" and self._action_size == 4) else self.KIND

    def _eval(self, x, u, i, terminal, want):
        if self._evaluator is None:
            kind = self._kind()
            N = len(self._q_ref) - 1
            self._evaluator = _native.make_solver(kind, "ss", N, 1, _PlaceholderDynamics(kind), self, self._q_ref,
                                                  self._xi_ref, max_iters=1)
        kind = self._kind()
        u_row = None if terminal else np.asarray(u, dtype=float)
        out = self._evaluator.stage_eval(i, _native.state_row(kind, x), u_row, terminal=terminal, want=want)
        return {k: v.cpu().numpy()[0] for k, v in out.items()}

    def _err(self, x, i):
        """(Log(q q_ref_i^-1), xi - xi_ref_i) (traopt_cost.py:659-673 / :366-379)."""
        e = self._eval(x, np.zeros(self._action_size), i, False, ("err",))["err"]
        n = self._error_state_size
        return e[:n], e[n:]

    def l(self, x, u, i, terminal=False):
        return float(self._eval(x, u, i, terminal, ("l",))["l"])

    def l_x(self, x, u, i, terminal=False):
        return self._eval(x, u, i, terminal, ("l_x",))["l_x"]

    def l_u(self, x, u, i, terminal=False):
        if terminal:
            return np.zeros(self._action_size)
        return self._eval(x, u, i, terminal, ("l_u",))["l_u"]

    def l_xx(self, x, u, i, terminal=False):
        return self._eval(x, u, i, terminal, ("l_xx",))["l_xx"]

    def l_ux(self, x, u, i, terminal=False):
        return np.zeros((self._action_size, self._state_size))

    def l_uu(self, x, u, i, terminal=False):
        if terminal:
            return np.zeros((self._action_size, self._action_size))
        return 2.0 * self._R


class SE3TrackingQuadraticGaussNewtonCost(_TrackingCost):
    """e^T Q1 e + dxi^T Q2 dxi + u^T R u with e = Log(q q_ref^-1), Gauss-Newton Hessian (traopt_cost.py:570-867)."""
    KIND = "se3"

    def __init__(self, Q, R, P, q_ref, xi_ref, state_size=(6, 6), action_size=6, **kwargs):
        self._init_common(Q, R, P, q_ref, xi_ref, state_size, action_size)


class SO3TrackingQuadraticGaussNewtonCost(_TrackingCost):
    """SO(3) twin (traopt_cost.py:280-564); the terminal value/gradient use Q, the terminal Hessian P."""
    KIND = "so3"

    def __init__(self, Q, R, P, q_ref, xi_ref, state_size=(3, 3), action_size=3, **kwargs):
        self._init_common(Q, R, P, q_ref, xi_ref, state_size, action_size)


class ALConstrainedCost(BaseCost):
    """l + lambda^T g + 1/2 g^T Imu g (traopt_cost.py:1173-1320)."""

    def __init__(self, cost, constraints, N, state_size=(6, 6), action_size=6, **kwargs):
        self._state_size = state_size[0] + state_size[1]
        self._error_state_size = state_size[0]
        self._vel_state_size = state_size[1]
        self._action_size = action_size
        self._constr_size = constraints.constr_size
        self.constr = constraints
        self.cost = cost
        self.N = N
        self.lmbd = np.zeros((N + 1, self._constr_size))                     # :1205
        self.mu = 0.0                                                         # :1206
        self.Imu = np.zeros((N + 1, self._constr_size, self._constr_size))    # :1207

    state_size = property(lambda self: self._state_size)
    error_state_size = property(lambda self: self._error_state_size)
    vel_state_size = property(lambda self: self._vel_state_size)
    action_size = property(lambda self: self._action_size)
    constr_size = property(lambda self: self._constr_size)
    Q = property(lambda self: self.cost.Q)
    R = property(lambda self: self.cost.R)
    P = property(lambda self: self.cost.P)

    def _err(self, x, i):
        return self.cost._err(x, i)

    def l(self, x, u, i, terminal=False):
        g = self.constr.g(x, u, i, terminal=terminal)
        return self.cost.l(x, u, i, terminal=terminal) + self.lmbd[i] @ g + 0.5 * (g @ self.Imu[i] @ g)

    def l_x(self, x, u, i, terminal=False):
        g = self.constr.g(x, u, i, terminal=terminal)
        gx = self.constr.g_x(x, u, i, terminal=terminal)
        return self.cost.l_x(x, u, i, terminal=terminal) + gx.T @ (self.lmbd[i] + self.Imu[i] @ g)

    def l_u(self, x, u, i, terminal=False):
        g = self.constr.g(x, u, i, terminal=terminal)
        gu = self.constr.g_u(x, u, i, terminal=terminal)
        return self.cost.l_u(x, u, i, terminal=terminal) + gu.T @ (self.lmbd[i] + self.Imu[i] @ g)

    def l_xx(self, x, u, i, terminal=False):
        gx = self.constr.g_x(x, u, i, terminal=terminal)
        return self.cost.l_xx(x, u, i, terminal=terminal) + gx.T @ self.Imu[i] @ gx

    def l_ux(self, x, u, i, terminal=False):
        gx = self.constr.g_x(x, u, i, terminal=terminal)
        gu = self.constr.g_u(x, u, i, terminal=terminal)
        return self.cost.l_ux(x, u, i, terminal=terminal) + gu.T @ self.Imu[i] @ gx

    def l_uu(self, x, u, i, terminal=False):
        gu = self.constr.g_u(x, u, i, terminal=terminal)
        return self.cost.l_uu(x, u, i, terminal=terminal) + gu.T @ self.Imu[i] @ gu
