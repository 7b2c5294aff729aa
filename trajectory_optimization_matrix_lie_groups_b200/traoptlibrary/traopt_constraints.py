"""Constraints with the reference's interface (traoptlibrary/traopt_constraints.py).

`InputConstraint` is a parameter carrier for the augmented-Lagrangian controller: its bounds are
packed into the native solver.  `g`, `g_x`, `g_u` are the closed forms of the reference
(:116-169) for scripts that evaluate violations on a returned trajectory.
"""
import numpy as np


class BaseConstraint:
    """Inequality constraint g(x, u, i) <= 0 (traopt_constraints.py:5-63)."""

    def g(self, x, u, i, terminal=False):
        raise NotImplementedError

    def g_x(self, x, u, i, terminal=False):
        raise NotImplementedError

    def g_u(self, x, u, i, terminal=False):
        raise NotImplementedError


class InputConstraint(BaseConstraint):
    """Box bounds lb <= u <= ub as g = [lb - u; u - ub] (traopt_constraints.py:66-169)."""

    def __init__(self, input_lb, input_ub, state_size=(6, 6), action_size=6):
        self._state_size = state_size[0] + state_size[1]
        self._action_size = action_size
        self._lb = np.broadcast_to(np.asarray(input_lb, dtype=float), (action_size,)).copy()
        self._ub = np.broadcast_to(np.asarray(input_ub, dtype=float), (action_size,)).copy()
        self._constr_size = 2 * action_size

    @property
    def state_size(self):
        return self._state_size

    @property
    def action_size(self):
        return self._action_size

    @property
    def constr_size(self):
        return self._constr_size

    @property
    def input_lb(self):
        return self._lb

    @property
    def input_ub(self):
        return self._ub

    def g(self, x, u, i, terminal=False):
        if terminal:
            return np.zeros(self._constr_size)
        u = np.asarray(u, dtype=float).reshape(self._action_size)
        return np.concatenate((self._lb - u, u - self._ub))

    def g_x(self, x, u, i, terminal=False):
        return np.zeros((self._constr_size, self._state_size))

    def g_u(self, x, u, i, terminal=False):
        if terminal:
            return np.zeros((self._constr_size, self._action_size))
        return np.vstack((-np.identity(self._action_size), np.identity(self._action_size)))


class InputVelocityConstraint(InputConstraint):
    """Box bounds on the input AND on the body velocity xi = [omega, v] — an addition to the reference (which only has
    `InputConstraint`), with the same BaseConstraint interface: g = [lb_u - u; u - ub_u; lb_xi - xi; xi - ub_xi].
    The velocity rows stay active at the terminal stage."""

    def __init__(self, input_lb, input_ub, xi_lb, xi_ub, state_size=(6, 6), action_size=6):
        super().__init__(input_lb, input_ub, state_size, action_size)
        self._nv = state_size[1]
        self._xlb = np.broadcast_to(np.asarray(xi_lb, dtype=float), (self._nv,)).copy()
        self._xub = np.broadcast_to(np.asarray(xi_ub, dtype=float), (self._nv,)).copy()
        self._constr_size = 2 * action_size + 2 * self._nv

    xi_lb = property(lambda self: self._xlb)
    xi_ub = property(lambda self: self._xub)

    def g(self, x, u, i, terminal=False):
        xi = np.asarray(x[1], dtype=float).reshape(self._nv)
        gu = np.zeros(2 * self._action_size) if terminal else super().g(x, u, i)
        return np.concatenate((gu, self._xlb - xi, xi - self._xub))

    def g_x(self, x, u, i, terminal=False):
        G = np.zeros((self._constr_size, self._state_size))
        m2, nv, npose = 2 * self._action_size, self._nv, self._state_size - self._nv
        G[m2:m2 + nv, npose:] = -np.identity(nv)
        G[m2 + nv:, npose:] = np.identity(nv)
        return G

    def g_u(self, x, u, i, terminal=False):
        G = np.zeros((self._constr_size, self._action_size))
        if not terminal:
            G[:2 * self._action_size] = super().g_u(x, u, i)
        return G
