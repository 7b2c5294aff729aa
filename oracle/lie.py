"""SO(3)/SE(3) closed forms used by the reference through manifpy (artivis/manif, un-pinned).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference never does Lie arithmetic itself:
it converts 4x4 matrices to manif objects (traoptlibrary/traopt_utilis.py:331-342, through
scipy `Rotation.from_matrix`, i.e. a *unit quaternion*), calls manif's rplus / rminus / lminus /
exp / log and their Jacobians, and converts back with `.transform()` (traopt_utilis.py:344-354).
This file restates those manif operations from their published closed forms (Sola et al.,
"A micro Lie theory", and Barfoot's Q matrix for the SE(3) Jacobian), with the conventions of
SURVEY.md Appendix A:

  * tangents are [omega, v] (angular first), the order traoptlibrary uses everywhere
    (traopt_utilis.py:43-55, 75-88); manif's own [v, omega] order and the re-ordering helpers
    `se32manifse3` / `manifse32se3` / `Jmnf2J` (traopt_utilis.py:356-399) therefore vanish here;
  * perturbations are right/local:  X (+) tau = X Exp(tau),  A (-) B = Log(B^-1 A);
    the tracking error alone is left/global:  lminus(A, B) = Log(A B^-1);
  * a pose is (quat, p) with quat = [x, y, z, w] (Eigen / manif / scipy coefficient order), kept
    unit-norm, which is what the reference's matrix -> quaternion -> manif round trip enforces.

Everything is single-element NumPy FP64 on purpose: this is the readable checker, not a fast path.
"""
import math

import numpy as np

# manif: Constants<double>::eps — threshold on theta^2 below which the small-angle forms are used.
MANIF_EPS = 1e-10

_I3 = np.eye(3)


def skew(w):
    """traopt_utilis.py:13-24."""
    return np.array([[0.0, -w[2], w[1]],
                     [w[2], 0.0, -w[0]],
                     [-w[1], w[0], 0.0]])


# --------------------------------------------------------------------------------------------
# quaternions ([x, y, z, w])
# --------------------------------------------------------------------------------------------

def quat_normalize(q):
    return q / math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])


def quat_mul(a, b):
    """Hamilton product a*b, coefficient order [x, y, z, w] (Eigen::Quaternion::operator*)."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def quat_conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_to_rot(q):
    """Eigen::Quaternion::toRotationMatrix (what manif `.rotation()` / `.transform()` return)."""
    x, y, z, w = q
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1.0 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1.0 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1.0 - (txx + tyy)],
    ])


def rot_to_quat(R):
    """scipy `Rotation.from_matrix(R).as_quat()` (Markley's method, then normalised).

    This is where the reference silently re-projects every pose onto SO(3):
    traopt_utilis.py:180 via :312 and :337 (SURVEY.md finding 5).
    """
    m00, m11, m22 = R[0, 0], R[1, 1], R[2, 2]
    tr = m00 + m11 + m22
    dec = (m00, m11, m22, tr)
    c = max(range(4), key=lambda i: dec[i])
    q = np.empty(4)
    if c != 3:
        i = c
        j = (i + 1) % 3
        k = (j + 1) % 3
        q[i] = 1.0 - tr + 2.0 * R[i, i]
        q[j] = R[j, i] + R[i, j]
        q[k] = R[k, i] + R[i, k]
        q[3] = R[k, j] - R[j, k]
    else:
        q[0] = R[2, 1] - R[1, 2]
        q[1] = R[0, 2] - R[2, 0]
        q[2] = R[1, 0] - R[0, 1]
        q[3] = 1.0 + tr
    return quat_normalize(q)


# --------------------------------------------------------------------------------------------
# SO(3)
# --------------------------------------------------------------------------------------------

def so3_exp(w):
    """manif SO3Tangent::exp -> unit quaternion."""
    th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2]
    if th2 > MANIF_EPS:
        th = math.sqrt(th2)
        s = math.sin(0.5 * th) / th
        return np.array([s * w[0], s * w[1], s * w[2], math.cos(0.5 * th)])
    return quat_normalize(np.array([0.5 * w[0], 0.5 * w[1], 0.5 * w[2], 1.0]))


def so3_log(q):
    """manif SO3::log (atan2 form on the unit quaternion), angle in (-pi, pi]."""
    s2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2]
    if s2 > MANIF_EPS:
        s = math.sqrt(s2)
        c = q[3]
        two_angle = 2.0 * (math.atan2(-s, -c) if c < 0.0 else math.atan2(s, c))
        k = two_angle / s
    else:
        k = 2.0
    return np.array([k * q[0], k * q[1], k * q[2]])


def _so3_jac_coeffs(th2):
    """(1-cos)/th^2 and (th-sin)/th^3 with manif's small-angle switch."""
    if th2 <= MANIF_EPS:
        return 0.5, 1.0 / 6.0
    th = math.sqrt(th2)
    return (1.0 - math.cos(th)) / th2, (th - math.sin(th)) / (th2 * th)


def so3_jl(w):
    """Left Jacobian of SO(3): I + (1-cos)/th^2 W + (th-sin)/th^3 W^2 (manif SO3Tangent::ljac)."""
    th2 = float(w @ w)
    W = skew(w)
    if th2 <= MANIF_EPS:
        return _I3 + 0.5 * W
    a, b = _so3_jac_coeffs(th2)
    return _I3 + a * W + b * (W @ W)


def so3_jr(w):
    """Right Jacobian = Jl(-w) (manif SO3Tangent::rjac)."""
    th2 = float(w @ w)
    W = skew(w)
    if th2 <= MANIF_EPS:
        return _I3 - 0.5 * W
    a, b = _so3_jac_coeffs(th2)
    return _I3 - a * W + b * (W @ W)


def _so3_jinv_coeff(th2):
    th = math.sqrt(th2)
    return 1.0 / th2 - (1.0 + math.cos(th)) / (2.0 * th * math.sin(th))


def so3_jr_inv(w):
    """I + W/2 + (1/th^2 - (1+cos)/(2 th sin)) W^2 (manif SO3Tangent::rjacinv)."""
    th2 = float(w @ w)
    W = skew(w)
    if th2 <= MANIF_EPS:
        return _I3 + 0.5 * W
    return _I3 + 0.5 * W + _so3_jinv_coeff(th2) * (W @ W)


def so3_jl_inv(w):
    """I - W/2 + (1/th^2 - (1+cos)/(2 th sin)) W^2 (manif SO3Tangent::ljacinv)."""
    th2 = float(w @ w)
    W = skew(w)
    if th2 <= MANIF_EPS:
        return _I3 - 0.5 * W
    return _I3 - 0.5 * W + _so3_jinv_coeff(th2) * (W @ W)


# --------------------------------------------------------------------------------------------
# SE(3): pose = (quat, p); tangent = [omega, v]
# --------------------------------------------------------------------------------------------

def se3_exp(tau):
    """manif SE3Tangent::exp: (Exp(omega), Jl(omega) v)."""
    w, v = tau[:3], tau[3:]
    return so3_exp(w), so3_jl(w) @ v


def se3_log(q, p):
    """manif SE3::log: [Log(R), Jl(omega)^-1 p]."""
    w = so3_log(q)
    return np.concatenate((w, so3_jl_inv(w) @ p))


def se3_compose(qa, pa, qb, pb):
    """manif SE3::compose: (qa qb, pa + Ra pb); the product quaternion is renormalised."""
    return quat_normalize(quat_mul(qa, qb)), pa + quat_to_rot(qa) @ pb


def se3_inverse(q, p):
    qi = quat_conj(q)
    return qi, -(quat_to_rot(qi) @ p)


def se3_adj(q, p):
    """Ad(T) = [[R, 0], [p^ R, R]] in [omega, v] order (manif SE3::adj after Jmnf2J)."""
    R = quat_to_rot(q)
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[3:, 3:] = R
    A[3:, :3] = skew(p) @ R
    return A


def se3_Q(w, v):
    """Barfoot's Q(omega, v), the off-diagonal block of the SE(3) left Jacobian (manif fillQ)."""
    th2 = float(w @ w)
    W = skew(w)
    V = skew(v)
    if th2 <= MANIF_EPS:
        B = 1.0 / 6.0 + th2 / 120.0
        C = -1.0 / 24.0 + th2 / 720.0
        D = -1.0 / 60.0
    else:
        th = math.sqrt(th2)
        s, c = math.sin(th), math.cos(th)
        B = (th - s) / (th2 * th)
        C = (1.0 - 0.5 * th2 - c) / (th2 * th2)
        D = C - 3.0 * (th - s - th2 * th / 6.0) / (th2 * th2 * th)
    WV = W @ V
    VW = V @ W
    WVW = WV @ W
    WW = W @ W
    return (0.5 * V + B * (WV + VW + WVW)
            - C * (WW @ V + V @ WW - 3.0 * WVW)
            - 0.5 * D * (WVW @ W + W @ WVW))


def se3_jl(tau):
    w, v = tau[:3], tau[3:]
    Jl = so3_jl(w)
    J = np.zeros((6, 6))
    J[:3, :3] = Jl
    J[3:, 3:] = Jl
    J[3:, :3] = se3_Q(w, v)
    return J


def se3_jr(tau):
    """Jr(tau) = Jl(-tau) (manif SE3Tangent::rjac, re-ordered to [omega, v])."""
    return se3_jl(-tau)


def se3_jr_inv(tau):
    """[[Jr^-1, 0], [-Jr^-1 Q(-w,-v) Jr^-1, Jr^-1]] (manif SE3Tangent::rjacinv)."""
    w, v = tau[:3], tau[3:]
    Ji = so3_jr_inv(w)
    J = np.zeros((6, 6))
    J[:3, :3] = Ji
    J[3:, 3:] = Ji
    J[3:, :3] = -Ji @ se3_Q(-w, -v) @ Ji
    return J


# --------------------------------------------------------------------------------------------
# 4x4 / 3x3 matrix <-> (quat, p): the converters of traopt_utilis.py:291-354
# --------------------------------------------------------------------------------------------

def se3_from_matrix(T):
    """SE32manifSE3 (traopt_utilis.py:331-342)."""
    return rot_to_quat(T[:3, :3]), np.array(T[:3, 3], dtype=float)


def se3_to_matrix(q, p):
    """manifSE32SE3 = manif `.transform()` (traopt_utilis.py:344-354)."""
    T = np.zeros((4, 4))
    T[:3, :3] = quat_to_rot(q)
    T[:3, 3] = p
    T[3, 3] = 1.0
    return T
