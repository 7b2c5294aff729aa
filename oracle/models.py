"""Oracle restatement of traoptlibrary's dynamics / cost / constraint callbacks.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every method cites the reference lines it follows;
the three load-bearing reference quirks (SURVEY.md finding 4) are reproduced on purpose and marked
QUIRK.  States keep the reference's shapes: SE3/drone x = [T (4x4 ndarray), xi (6,)],
SO3 x = [quat (4,) unit, omega (3,)] (the reference holds a manif SO3 object there, i.e. a unit
quaternion; `so3_state` / `so3_state_to_matrix` convert from/to the 3x3 arrays stored in pickles).
"""
import numpy as np

from . import lie


def _adjoint(xi):
    """traopt_utilis.py:75-88: ad_xi = [[w^, 0], [v^, w^]] with xi = [first3, last3]."""
    w, v = xi[:3], xi[3:]
    A = np.zeros((6, 6))
    A[:3, :3] = lie.skew(w)
    A[3:, 3:] = A[:3, :3]
    A[3:, :3] = lie.skew(v)
    return A


def _coadjoint(xi):
    """traopt_utilis.py:90-92."""
    return _adjoint(xi).T


def so3_state(R, w):
    return [lie.rot_to_quat(np.asarray(R, dtype=float)), np.array(w, dtype=float)]


def so3_state_to_matrix(x):
    return [lie.quat_to_rot(x[0]), np.array(x[1])]


# ============================================================================================
# Dynamics
# ============================================================================================

class SO3Dynamics:
    """traopt_dynamics.py:275-418."""
    state_size = 6
    action_size = 3

    def __init__(self, J, dt):
        self.J = np.asarray(J, dtype=float)
        self.Jinv = np.linalg.inv(self.J)                                   # :308
        self.dt = float(dt)
        self._Bt = np.vstack((np.zeros((3, 3)), self.Jinv))                 # :311-313

    def f(self, x, u, i):
        """fd_euler, :369-380.  q+ = q Exp(w dt); w+ = w + J^-1 (w^T^ J w + u) dt."""
        q, w = x
        q_next = lie.quat_normalize(lie.quat_mul(q, lie.so3_exp(w * self.dt)))
        w_next = w + self.Jinv @ (lie.skew(w).T @ self.J @ w + u) * self.dt  # smallAdj = w^
        return [q_next, w_next]

    def f_x(self, x, u, i):
        """:385-400."""
        q, w = x
        th = w * self.dt
        J_q_q = lie.quat_to_rot(lie.so3_exp(th)).T       # rplus: Ad(Exp(th))^-1 = Exp(th)^T
        J_q_xi = lie.so3_jr(th) * self.dt
        G = lie.skew(self.J @ w)
        H = self.Jinv @ (lie.skew(w).T @ self.J + G)
        A = np.zeros((6, 6))
        A[:3, :3] = J_q_q
        A[:3, 3:] = J_q_xi
        A[3:, 3:] = np.eye(3) + H * self.dt
        return A

    def f_u(self, x, u, i):
        return self._Bt * self.dt                                           # :402-403


class Pendulum3dDynamics(SO3Dynamics):
    """traopt_dynamics.py:421-626: 3-D pendulum, gravity torque + force input at the pivot."""

    def __init__(self, J, m, length, dt):
        super().__init__(J, dt)
        self.m, self.l, self.g = float(m), float(length), 9.8               # :458-460

    def _rho(self):
        return self.l / 2 * np.array([0.0, 0.0, -1.0])                      # :524

    def f(self, x, u, i):
        """fd_euler :520-541."""
        q, w = x
        down = np.array([0.0, 0.0, -1.0])
        Rt = lie.quat_to_rot(q).T
        g_term = lie.skew(self.m * self.g * self._rho()) @ (Rt @ down)       # smallAdj of an SO3 tangent = skew
        M = lie.skew(self.m * self._rho()) @ (Rt @ np.asarray(u, dtype=float))
        qn = lie.quat_normalize(lie.quat_mul(q, lie.so3_exp(w * self.dt)))
        wn = w + self.Jinv @ (lie.skew(w).T @ self.J @ w + g_term + M) * self.dt
        return [qn, wn]

    def f_x(self, x, u, i):
        """:543-583.  d(R^T v)/dR = skew(R^T v) for a right perturbation of R (act and inverse Jacobians of manif)."""
        q, w = x
        A = super().f_x(x, u, i)
        down = np.array([0.0, 0.0, -1.0])
        Rt = lie.quat_to_rot(q).T
        L1 = lie.skew(self.m * self.g * self._rho()) @ lie.skew(Rt @ down)
        L2 = lie.skew(self.m * self._rho()) @ lie.skew(Rt @ np.asarray(u, dtype=float))
        A[3:, :3] = self.Jinv @ (L1 + L2) * self.dt
        return A

    def f_u(self, x, u, i):
        """:585-603: J^-1 skew(m rho) R^T dt in the velocity rows."""
        q, w = x
        bt = self.Jinv @ lie.skew(self.m * self._rho()) @ lie.quat_to_rot(q).T
        return np.vstack((np.zeros((3, 3)), bt)) * self.dt


class SE3Dynamics:
    """traopt_dynamics.py:629-898."""
    state_size = 12
    action_size = 6

    def __init__(self, J, dt):
        self.J = np.asarray(J, dtype=float)
        self.Ib = self.J[0:3, 0:3]                                          # :662
        self.m = self.J[4, 4]                                               # :663
        self.Jinv = np.linalg.inv(self.J)                                   # :665
        self.dt = float(dt)
        self._Bt = np.vstack((np.zeros((6, self.action_size)), self._input_map()))  # :668-670

    def _input_map(self):
        return self.Jinv

    def _force(self, q_quat, u):
        """Generalised force on the twist equation besides the Coriolis term."""
        return u

    def f(self, x, u, i):
        """fd_euler, :763-787."""
        T, xi = x
        q, p = lie.se3_from_matrix(T)                                       # :777
        eq, ep = lie.se3_exp(xi * self.dt)
        qn, pn = lie.se3_compose(q, p, eq, ep)                              # :783 rplus
        xi_next = xi + self.Jinv @ (_coadjoint(xi) @ self.J @ xi + self._force(q, u)) * self.dt  # :785
        return [lie.se3_to_matrix(qn, pn), xi_next]

    def _f_x_lower_left(self, q_quat):
        return np.zeros((6, 6))

    def f_x(self, x, u, i):
        """:802-837."""
        T, xi = x
        omega, v = xi[:3], xi[3:]
        tau = xi * self.dt
        eq, ep = lie.se3_exp(-tau)
        J_q_q = lie.se3_adj(eq, ep)                       # rplus d/dX = Ad(Exp(tau))^-1 = Ad(Exp(-tau))
        J_q_xi = lie.se3_jr(tau) * self.dt                # :826
        G = np.zeros((6, 6))                              # :828-831
        G[:3, :3] = lie.skew(self.Ib @ omega)
        G[:3, 3:] = self.m * lie.skew(v)
        G[3:, :3] = self.m * lie.skew(v)
        # QUIRK 1 (:819 then :832): `xi` was rebound to a manif tangent, whose coeffs() are
        # [v, omega]; coadjoint() then treats v as the angular part.
        xi_swapped = np.concatenate((v, omega))
        H = self.Jinv @ (_coadjoint(xi_swapped) @ self.J + G)
        A = np.zeros((12, 12))
        A[:6, :6] = J_q_q
        A[:6, 6:] = J_q_xi
        A[6:, :6] = self._f_x_lower_left(lie.se3_from_matrix(T)[0])
        A[6:, 6:] = np.eye(6) + H * self.dt
        return A

    def f_u(self, x, u, i):
        return self._Bt * self.dt                                           # :839-850


class RigidBodyDynamics(SE3Dynamics):
    """traopt_dynamics.py:901-1206: SE3 + gravity, inputs act on all six twist components."""
    action_size = 6

    def __init__(self, J, dt):
        self.g = 9.8                                                        # :936
        super().__init__(J, dt)

    def _force(self, q_quat, u):
        """:1066-1072: [0; m g R^T (-e3)] + u."""
        down = np.array([0.0, 0.0, -1.0])
        g_acc = self.m * self.g * (lie.quat_to_rot(q_quat).T @ down)
        return np.concatenate((np.zeros(3), g_acc)) + u

    def _f_x_lower_left(self, q_quat):
        """:1110-1134: the same gravity block as the quadrotor, also without the m*g factor."""
        down = np.array([0.0, 0.0, -1.0])
        J_xi_q = np.zeros((6, 6))
        J_xi_q[3:, :3] = lie.skew(lie.quat_to_rot(q_quat).T @ down)
        return (self.Jinv @ J_xi_q) * self.dt


class DroneDynamics(SE3Dynamics):
    """traopt_dynamics.py:1209-1530: SE3 + gravity + 4->6 input map."""
    action_size = 4

    def __init__(self, J, dt):
        self.g = 9.8                                                        # :1245
        self.Pu = np.zeros((6, 4))                                          # :1250-1254
        self.Pu[0, 0] = self.Pu[1, 1] = self.Pu[2, 2] = self.Pu[5, 3] = 1.0
        super().__init__(J, dt)

    def _input_map(self):
        return self.Jinv @ self.Pu                                          # :1256-1258

    def _force(self, q_quat, u):
        """:1393-1399: [0; m g R^T (-e3)] + Pu u."""
        down = np.array([0.0, 0.0, -1.0])
        g_acc = self.m * self.g * (lie.quat_to_rot(q_quat).T @ down)
        return np.concatenate((np.zeros(3), g_acc)) + self.Pu @ u

    def _f_x_lower_left(self, q_quat):
        """:1445-1458.  J_v_R = d(R^T e)/dR = skew(R^T e), e = -e3.

        QUIRK 2: the m*g factor present in f (:1394) is missing here.
        """
        down = np.array([0.0, 0.0, -1.0])
        J_v_R = lie.skew(lie.quat_to_rot(q_quat).T @ down)
        J_xi_q = np.zeros((6, 6))
        J_xi_q[3:, :3] = J_v_R
        return (self.Jinv @ J_xi_q) * self.dt


# ============================================================================================
# Costs
# ============================================================================================

class SE3TrackingQuadraticGaussNewtonCost:
    """traopt_cost.py:570-867."""
    state_size = 12

    def __init__(self, Q, R, P, q_ref, xi_ref):
        self.Q, self.R, self.P = (np.asarray(a, dtype=float) for a in (Q, R, P))
        self.action_size = self.R.shape[0]
        self._q_ref = [lie.se3_from_matrix(T) for T in q_ref]               # :614
        self._xi_ref = np.asarray(xi_ref, dtype=float)

    def _err_jac(self, x, i, want_jac):
        """lminus(q, q_ref) = Log(q q_ref^-1) and d/dq = Jr^-1(e) Ad(q_ref) (:776-779)."""
        T, xi = x
        q, p = lie.se3_from_matrix(T)
        qr, pr = self._q_ref[i]
        qi, pi = lie.se3_inverse(qr, pr)
        qe, pe = lie.se3_compose(q, p, qi, pi)
        e = lie.se3_log(qe, pe)
        if not want_jac:
            return e, None
        return e, lie.se3_jr_inv(e) @ lie.se3_adj(qr, pr)

    def _err(self, x, i):
        return self._err_jac(x, i, False)[0], x[1] - self._xi_ref[i]        # :659-673

    def l(self, x, u, i, terminal=False):
        """:675-756."""
        W = self.P if terminal else self.Q
        e, dxi = self._err(x, i)
        c = e @ W[:6, :6] @ e + dxi @ W[6:, 6:] @ dxi
        if not terminal:
            c = c + u @ self.R @ u
        return float(c)

    def l_x(self, x, u, i, terminal=False):
        """:758-790."""
        W = self.P if terminal else self.Q
        e, Je = self._err_jac(x, i, True)
        return np.concatenate(((Je.T * 2) @ W[:6, :6] @ e,
                               2 * W[6:, 6:] @ (x[1] - self._xi_ref[i])))

    def l_u(self, x, u, i, terminal=False):
        return 2 * self.R @ u                                               # :792-804

    def l_xx(self, x, u, i, terminal=False):
        """:806-839 (Gauss-Newton)."""
        W = self.P if terminal else self.Q
        _, Je = self._err_jac(x, i, True)
        H = np.zeros((12, 12))
        H[:6, :6] = (Je.T * 2) @ W[:6, :6] @ Je
        H[6:, 6:] = 2 * W[6:, 6:]
        return H

    def l_ux(self, x, u, i, terminal=False):
        return np.zeros((self.action_size, 12))                             # :841-853

    def l_uu(self, x, u, i, terminal=False):
        return 2 * self.R                                                   # :855-867


class SO3TrackingQuadraticGaussNewtonCost:
    """traopt_cost.py:280-564."""
    state_size = 6
    action_size = 3

    def __init__(self, Q, R, P, q_ref, xi_ref):
        self.Q, self.R, self.P = (np.asarray(a, dtype=float) for a in (Q, R, P))
        self._q_ref = [lie.rot_to_quat(np.asarray(Rm, dtype=float)) for Rm in q_ref]   # :322
        self._xi_ref = np.asarray(xi_ref, dtype=float)

    def _err_jac(self, x, i):
        """lminus(q, q_ref) = Log(q q_ref^-1); d/dq = Jr^-1(e) Ad(q_ref) = Jr^-1(e) R_ref."""
        q, _ = x
        qr = self._q_ref[i]
        e = lie.so3_log(lie.quat_normalize(lie.quat_mul(q, lie.quat_conj(qr))))
        return e, lie.so3_jr_inv(e) @ lie.quat_to_rot(qr)

    def _err(self, x, i):
        return self._err_jac(x, i)[0], x[1] - self._xi_ref[i]               # :366-379

    def l(self, x, u, i, terminal=False):
        """:381-458.  QUIRK 3: the terminal *value* uses Q, not P (:434, :438)."""
        e, dw = self._err(x, i)
        c = e @ self.Q[:3, :3] @ e + dw @ self.Q[3:, 3:] @ dw
        if not terminal:
            c = c + u @ self.R @ u
        return float(c)

    def l_x(self, x, u, i, terminal=False):
        """:460-487.  QUIRK 3: no terminal branch, always Q."""
        e, Je = self._err_jac(x, i)
        return np.concatenate(((Je.T * 2) @ self.Q[:3, :3] @ e,
                               2 * self.Q[3:, 3:] @ (x[1] - self._xi_ref[i])))

    def l_u(self, x, u, i, terminal=False):
        return 2 * self.R @ u                                               # :489-501

    def l_xx(self, x, u, i, terminal=False):
        """:503-536.  The terminal *Hessian* does use P (:529-531)."""
        W = self.P if terminal else self.Q
        _, Je = self._err_jac(x, i)
        H = np.zeros((6, 6))
        H[:3, :3] = (Je.T * 2) @ W[:3, :3] @ Je
        H[3:, 3:] = 2 * W[3:, 3:]
        return H

    def l_ux(self, x, u, i, terminal=False):
        return np.zeros((3, 6))                                             # :538-550

    def l_uu(self, x, u, i, terminal=False):
        return 2 * self.R                                                   # :552-564


# ============================================================================================
# Constraints + augmented Lagrangian
# ============================================================================================

class InputConstraint:
    """traopt_constraints.py:66-169: g = [lb - u; u - ub]."""

    def __init__(self, input_lb, input_ub, state_size=12, action_size=6):
        self.lb = np.asarray(input_lb, dtype=float)
        self.ub = np.asarray(input_ub, dtype=float)
        self.state_size = state_size
        self.action_size = action_size
        self.constr_size = 2 * action_size                                  # :81

    def g(self, x, u, i, terminal=False):
        if terminal:
            return np.zeros(self.constr_size)                               # :127-128
        return np.concatenate([self.lb - u, u - self.ub])                   # :130-133

    def g_x(self, x, u, i, terminal=False):
        return np.zeros((self.constr_size, self.state_size))                # :150

    def g_u(self, x, u, i, terminal=False):
        if terminal:
            return np.zeros((self.constr_size, self.action_size))           # :164-165
        return np.vstack([-np.identity(self.action_size), np.identity(self.action_size)])  # :167-169


class InputVelocityConstraint(InputConstraint):
    """Box bounds on u AND on the body velocity xi (NOT in the reference, which has InputConstraint only): the same
    BaseConstraint interface (traopt_constraints.py:5-63), g = [lb_u - u; u - ub_u; lb_xi - xi; xi - ub_xi],
    g_x = [0 0; 0 0; 0 -I; 0 I], g_u = [-I; I; 0; 0].  The velocity rows stay active at the terminal stage."""

    def __init__(self, input_lb, input_ub, xi_lb, xi_ub, state_size=12, action_size=6):
        super().__init__(input_lb, input_ub, state_size, action_size)
        self.nv = state_size // 2
        self.xlb = np.broadcast_to(np.asarray(xi_lb, dtype=float), (self.nv,)).copy()
        self.xub = np.broadcast_to(np.asarray(xi_ub, dtype=float), (self.nv,)).copy()
        self.constr_size = 2 * action_size + 2 * self.nv

    def g(self, x, u, i, terminal=False):
        xi = np.asarray(x[1], dtype=float)
        gu = np.zeros(2 * self.action_size) if terminal else np.concatenate([self.lb - u, u - self.ub])
        return np.concatenate([gu, self.xlb - xi, xi - self.xub])

    def g_x(self, x, u, i, terminal=False):
        G = np.zeros((self.constr_size, self.state_size))
        m2, nv, npose = 2 * self.action_size, self.nv, self.state_size - self.nv
        G[m2:m2 + nv, npose:] = -np.identity(nv)
        G[m2 + nv:, npose:] = np.identity(nv)
        return G

    def g_u(self, x, u, i, terminal=False):
        G = np.zeros((self.constr_size, self.action_size))
        if not terminal:
            G[:self.action_size] = -np.identity(self.action_size)
            G[self.action_size:2 * self.action_size] = np.identity(self.action_size)
        return G


class ALConstrainedCost:
    """traopt_cost.py:1173-1320: l + lambda^T g + 1/2 g^T Imu g and its derivatives."""

    def __init__(self, cost, constraints, N):
        self.cost = cost
        self.constr = constraints
        self.N = N
        self.state_size = cost.state_size
        self.action_size = cost.action_size
        c = constraints.constr_size
        self.lmbd = np.zeros((N + 1, c))                                    # :1205
        self.mu = 0.0                                                       # :1206
        self.Imu = np.zeros((N + 1, c, c))                                  # :1207

    def l(self, x, u, i, terminal=False):
        g = self.constr.g(x, u, i, terminal=terminal)                       # :1236-1249
        return self.cost.l(x, u, i, terminal=terminal) + self.lmbd[i] @ g + 0.5 * (g @ self.Imu[i] @ g)

    def l_x(self, x, u, i, terminal=False):
        g = self.constr.g(x, u, i, terminal=terminal)                       # :1251-1264
        gx = self.constr.g_x(x, u, i, terminal=terminal)
        return self.cost.l_x(x, u, i, terminal=terminal) + gx.T @ (self.lmbd[i] + self.Imu[i] @ g)

    def l_u(self, x, u, i, terminal=False):
        g = self.constr.g(x, u, i, terminal=terminal)                       # :1266-1279
        gu = self.constr.g_u(x, u, i, terminal=terminal)
        return self.cost.l_u(x, u, i, terminal=terminal) + gu.T @ (self.lmbd[i] + self.Imu[i] @ g)

    def l_uu(self, x, u, i, terminal=False):
        gu = self.constr.g_u(x, u, i, terminal=terminal)                    # :1281-1292
        return self.cost.l_uu(x, u, i, terminal=terminal) + gu.T @ self.Imu[i] @ gu

    def l_xx(self, x, u, i, terminal=False):
        gx = self.constr.g_x(x, u, i, terminal=terminal)                    # :1294-1305
        return self.cost.l_xx(x, u, i, terminal=terminal) + gx.T @ self.Imu[i] @ gx

    def l_ux(self, x, u, i, terminal=False):
        gx = self.constr.g_x(x, u, i, terminal=terminal)                    # :1307-1320
        gu = self.constr.g_u(x, u, i, terminal=terminal)
        return self.cost.l_ux(x, u, i, terminal=terminal) + gu.T @ self.Imu[i] @ gx
