"""Host utilities with the reference's names (traoptlibrary/traopt_utilis.py).

Only the helpers the tracking scripts call around the solver are provided: hat/vee/ad maps
(:13-92), quaternion / Euler / matrix conversions through scipy (:159-316), `is_pos_def` (:320-329).
The manif converters (:331-399) keep their names and meaning on the light stand-in types of
`manif_compat` (the device state row is already the (unit quaternion, position) pair they produce).
"""
import numpy as np
from scipy.spatial.transform import Rotation

from .. import layout
from .manif_compat import SE3, SE3Tangent


def skew(w):
    """3-vector -> so(3) matrix (traopt_utilis.py:13-24)."""
    w = np.asarray(w, dtype=float).reshape(3)
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def unskew(W):
    """so(3) matrix -> 3-vector (:26-41)."""
    W = np.asarray(W)
    return np.array([W[2, 1], W[0, 2], W[1, 0]])


def se3_hat(xi):
    """[w, v] -> 4x4 se(3) matrix (:43-55)."""
    xi = np.asarray(xi, dtype=float).reshape(6)
    M = np.zeros((4, 4))
    M[:3, :3] = skew(xi[:3])
    M[:3, 3] = xi[3:]
    return M


def se3_vee(M):
    """4x4 se(3) matrix -> [w, v] (:57-73)."""
    M = np.asarray(M)
    return np.concatenate((unskew(M[:3, :3]), M[:3, 3]))


def adjoint(xi):
    """ad_xi = [[w^, 0], [v^, w^]] for xi = [w, v] (:75-88)."""
    xi = np.asarray(xi, dtype=float).reshape(6)
    A = np.zeros((6, 6))
    A[:3, :3] = skew(xi[:3])
    A[3:, 3:] = A[:3, :3]
    A[3:, :3] = skew(xi[3:])
    return A


def coadjoint(xi):
    """ad_xi^T (:90-92)."""
    return adjoint(xi).T


def SE32absangle(m):
    """Geodesic rotation angle of a 4x4 pose, degrees (:94-112)."""
    m = np.asarray(m)
    if m.shape != (4, 4):
        raise ValueError("The input must be a 4x4 SE3 matrix")
    return np.rad2deg(np.arccos((np.trace(m[:3, :3]) - 1) / 2))


def rotm2absangle(m):
    """Geodesic rotation angle of a 3x3 rotation, degrees (:121-138)."""
    m = np.asarray(m)
    if m.shape != (3, 3):
        raise ValueError("The input must be a 3x3 rotation matrix")
    return np.rad2deg(np.arccos((np.trace(m) - 1) / 2))


def parallel_SE32absangle(m_list):
    return np.array([SE32absangle(m) for m in m_list])


def parallel_rotm2absangle(m_list):
    return np.array([rotm2absangle(m) for m in m_list])


def quat2rotm(quat):
    """Scalar-first quaternion(s) -> rotation matrix (:159-161)."""
    return Rotation.from_quat(quat, scalar_first=True).as_matrix()


def quat2euler(quat):
    """Scalar-first quaternion(s) -> 'zxy' Euler angles in degrees (:163-165)."""
    return Rotation.from_quat(quat, scalar_first=True).as_euler('zxy', degrees=True)


def rotm2quat(m):
    """Rotation matrix -> scalar-first quaternion (:167-181)."""
    q1, q2, q3, q0 = layout.rot_to_quat(np.asarray(m, dtype=float))
    return np.array([q0, q1, q2, q3])


def rotm2euler(m, order='zxy'):
    """Rotation matrix -> Euler angles in degrees (:183-199)."""
    return np.asarray(Rotation.from_matrix(m).as_euler(order or 'zxy', degrees=True))


def parallel_rotm2euler(m_list, order):
    return np.array([rotm2euler(m, order) for m in m_list])


def euler2quat(eulerAngles):
    """[roll, pitch, yaw] (rad, ZYX intrinsic) -> scalar-first quaternion (:209-250)."""
    e = np.asarray(eulerAngles, dtype=float).reshape(-1)
    if e.size != 3:
        raise TypeError("The eulerAngles must be given as [3x1] np.ndarray vector or a python list of 3 elements")
    cr, cp, cy = np.cos(e / 2)
    sr, sp, sy = np.sin(e / 2)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])


def quatpos2SE3(x7):
    """[quat (scalar first), position] -> 4x4 pose (:252-271)."""
    x7 = np.asarray(x7, dtype=float)
    if x7.shape not in ((7,), (7, 1)):
        raise ValueError("Input must be a 7-d np or jnp vector")
    x7 = x7.reshape(7)
    T = np.eye(4)
    T[:3, :3] = quat2rotm(x7[:4])
    T[:3, 3] = x7[4:]
    return T


def rotmpos2SE3(m, x):
    """(rotation matrix, position) -> 4x4 pose (:273-289)."""
    x = np.asarray(x, dtype=float)
    if x.shape not in ((3,), (3, 1)):
        raise ValueError("Input dimension incorrect")
    T = np.eye(4)
    T[:3, :3] = np.asarray(m, dtype=float)
    T[:3, 3] = x.reshape(3)
    return T


def SE32quatpos(m):
    """4x4 pose -> (7,1) [quat (scalar first), position] (:299-316)."""
    m = np.asarray(m, dtype=float)
    if m.shape != (4, 4):
        raise ValueError("Input must be a 4*4 np or jnp vector")
    return np.concatenate((rotm2quat(m[:3, :3]), m[:3, 3])).reshape((7, 1))


def is_pos_def(A):
    """Symmetric and Cholesky succeeds (:320-329)."""
    A = np.asarray(A)
    if np.array_equal(A, A.T):
        try:
            np.linalg.cholesky(A)
            return True
        except np.linalg.LinAlgError:
            return False
    return False


def SE32manifSE3(x):
    """4x4 pose -> SE3(position, quaternion [x,y,z,w]); the matrix -> quaternion step re-projects onto SO(3)
    exactly as the reference's does (:331-342)."""
    quatpos = SE32quatpos(x).reshape(7)
    q0, q1, q2, q3 = quatpos[:4]
    return SE3(position=quatpos[4:], quaternion=np.array([q1, q2, q3, q0]))


def manifSE32SE3(x):
    """SE3 object -> 4x4 pose (:344-354)."""
    return x.transform()


def se32manifse3(x):
    """Twist [omega, v] -> SE3Tangent in manif's [v, omega] order (:356-367)."""
    x = np.asarray(x, dtype=float).reshape(6)
    return SE3Tangent(np.concatenate((x[3:], x[:3])))


def manifse32se3(x):
    """SE3Tangent (or its [v, omega] coefficients) -> twist [omega, v] (:369-383)."""
    if isinstance(x, SE3Tangent):
        x = x.coeffs()
    x = np.asarray(x, dtype=float).reshape(6)
    return np.concatenate((x[3:], x[:3]))


def Jmnf2J(J):
    """Reorder a 6x6 manif Jacobian ([v, omega] blocks) to the library's [omega, v] order (:387-399)."""
    J = np.asarray(J)
    return np.block([[J[3:, 3:], J[3:, :3]], [J[:3, 3:], J[:3, :3]]])


def parallel_SE32manifSE3(q_ref):
    """(:401-406); the reference maps over a thread pool, the result is the same list."""
    return [SE32manifSE3(q) for q in q_ref]
