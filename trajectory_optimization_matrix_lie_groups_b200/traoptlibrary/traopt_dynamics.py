"""Dynamics classes with the reference's interface (traoptlibrary/traopt_dynamics.py).

`SO3Dynamics` (:275-418), `SE3Dynamics` (:629-898), `RigidBodyDynamics` (:901-1206) and `DroneDynamics`
(:1209-1530) carry the
inertia / time step the controllers pack into the native solver.  Their per-stage callbacks
`f`, `f_x`, `f_u` are evaluated by the CUDA library (trajopt_debug_stage) — there is no NumPy
implementation here.
"""
import numpy as np

from . import _native


class BaseDynamics:
    """Dynamics model interface (traopt_dynamics.py:14-130)."""

    @property
    def state_size(self):
        raise NotImplementedError

    @property
    def action_size(self):
        raise NotImplementedError

    @property
    def has_hessians(self):
        raise NotImplementedError

    def f(self, x, u, i):
        raise NotImplementedError

    def f_x(self, x, u, i):
        raise NotImplementedError

    def f_u(self, x, u, i):
        raise NotImplementedError

    def f_xx(self, x, u, i):
        raise NotImplementedError

    def f_ux(self, x, u, i):
        raise NotImplementedError

    def f_uu(self, x, u, i):
        raise NotImplementedError


class _NativeDynamics(BaseDynamics):
    KIND = None

    def _init_common(self, J, dt, integration_method, state_size, action_size, hessians, debug):
        if integration_method != "euler":
            # the reference raises for "rk4" too (traopt_dynamics.py:676-678)
            raise ValueError("Invalid integration method. Choose 'euler' or 'rk4'." if integration_method != "rk4"
                             else "RK4 not implemented yet.")
        self._state_size = state_size[0] + state_size[1]
        self._error_state_size = state_size[0]
        self._vel_state_size = state_size[1]
        self._action_size = action_size
        self._J = np.asarray(J, dtype=float)
        self._Jinv = np.linalg.inv(self._J)
        self._dt = float(dt)
        self._integration_method = integration_method
        self._has_hessians = hessians
        self._debug = debug
        self._evaluator = None

    state_size = property(lambda self: self._state_size)
    error_state_size = property(lambda self: self._error_state_size)
    vel_state_size = property(lambda self: self._vel_state_size)
    action_size = property(lambda self: self._action_size)
    has_hessians = property(lambda self: self._has_hessians)
    J = property(lambda self: self._J)
    Jinv = property(lambda self: self._Jinv)
    dt = property(lambda self: self._dt)

    # ---- per-stage callbacks through the native library ------------------------------------
    def _eval(self, x, u, want):
        if self._evaluator is None:
            from .traopt_cost import _placeholder_cost
            nx = self._state_size
            cost = _placeholder_cost(nx, self._action_size)
            if self.KIND in ("so3", "pendulum"):
                q_ref, xi_ref = np.tile(np.array([0.0, 0, 0, 1]), (2, 1)), np.zeros((2, 3))
            else:
                q_ref, xi_ref = np.tile(np.eye(4), (2, 1, 1)), np.zeros((2, 6))
            self._evaluator = _native.make_solver(self.KIND, "ss", 1, 1, self, cost, q_ref, xi_ref, max_iters=1)
        out = self._evaluator.stage_eval(0, _native.state_row(self.KIND, x), np.asarray(u, dtype=float), want=want)
        return {k: v.cpu().numpy()[0] for k, v in out.items()}

    def f(self, x, u, i):
        """x+ = f(x, u): exact discrete rigid-body step on the group."""
        return _native.row_state(self.KIND, self._eval(x, u, ("f",))["f"])

    def f_x(self, x, u, i):
        return self._eval(x, u, ("F_x",))["F_x"]

    def f_u(self, x, u, i):
        return self._eval(x, u, ("F_u",))["F_u"]

    def _second_order(self, name):
        """The reference's error behaviour (traopt_dynamics.py:852-898): NotImplementedError without `hessians=True`;
        with it, the attribute error of a `_f_xx` that the exact models never define (:684-686 is commented out)."""
        if not self._has_hessians:
            raise NotImplementedError
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '_{name}'")

    def f_xx(self, x, u, i):
        return self._second_order("f_xx")

    def f_ux(self, x, u, i):
        return self._second_order("f_ux")

    def f_uu(self, x, u, i):
        return self._second_order("f_uu")


class SO3Dynamics(_NativeDynamics):
    """Rigid-body attitude dynamics on SO(3) (traopt_dynamics.py:275-418)."""
    KIND = "so3"

    def __init__(self, J, dt, integration_method="euler", state_size=(3, 3), action_size=3, hessians=False,
                 debug=None, **kwargs):
        self._init_common(J, dt, integration_method, state_size, action_size, hessians, debug)


class Pendulum3dDyanmics(_NativeDynamics):
    """3-D pendulum on SO(3) actuated by a force at the pivot (traopt_dynamics.py:421-626; the class name,
    typo included, is the reference's)."""
    KIND = "pendulum"

    def __init__(self, J, m, length, dt, integration_method="euler", state_size=(3, 3), action_size=3, hessians=False,
                 debug=None, **kwargs):
        self._init_common(J, dt, integration_method, state_size, action_size, hessians, debug)
        self._m = float(m)
        self._l = float(length)
        self._g = 9.8                                         # :460

    m = property(lambda self: self._m)
    l = property(lambda self: self._l)
    g = property(lambda self: self._g)


class SE3Dynamics(_NativeDynamics):
    """Rigid-body dynamics on SE(3), twist [w, v], J = diag(I_b, m I) (traopt_dynamics.py:629-898)."""
    KIND = "se3"

    def __init__(self, J, dt, integration_method="euler", state_size=(6, 6), action_size=6, hessians=False,
                 debug=None, **kwargs):
        self._init_common(J, dt, integration_method, state_size, action_size, hessians, debug)
        self._Ib = self._J[0:3, 0:3]          # :662
        self._m = float(self._J[4, 4])        # :663

    Ib = property(lambda self: self._Ib)
    m = property(lambda self: self._m)


class RigidBodyDynamics(SE3Dynamics):
    """Rigid body on SE(3) under gravity, 6 inputs (traopt_dynamics.py:901-1206)."""
    KIND = "rigid"

    def __init__(self, J, dt, integration_method="euler", state_size=(6, 6), action_size=6, hessians=False,
                 debug=None, **kwargs):
        super().__init__(J, dt, integration_method, state_size, action_size, hessians, debug)
        self._g = 9.8                                         # :936

    g = property(lambda self: self._g)


class DroneDynamics(SE3Dynamics):
    """Quadrotor on SE(3): gravity + (tau_x, tau_y, tau_z, f_z) input map (traopt_dynamics.py:1209-1530)."""
    KIND = "drone"

    def __init__(self, J, dt, integration_method="euler", state_size=(6, 6), action_size=4, hessians=False,
                 debug=None, **kwargs):
        super().__init__(J, dt, integration_method, state_size, action_size, hessians, debug)
        self._g = 9.8                                         # :1245
        self._Pu = np.zeros((6, 4))                           # :1250-1254
        self._Pu[0, 0] = self._Pu[1, 1] = self._Pu[2, 2] = self._Pu[5, 3] = 1.0

    g = property(lambda self: self._g)
    Pu = property(lambda self: self._Pu)
