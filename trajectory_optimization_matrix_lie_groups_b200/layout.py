"""Host-side layout conversions between the reference's API types and the device state rows.

At the class API a pose is a 4x4 (SE3) or 3x3 (SO3) float64 matrix and a state is the Python list
`[q, xi]`; on the device a state row is `quat[x,y,z,w] (+ p) + velocity` (include/trajopt_b200.h).
The matrix -> quaternion step is the same one the reference performs at every manif call
(traoptlibrary/traopt_utilis.py:331-342 via scipy `Rotation.from_matrix`), so poses that enter as
matrices land on exactly the unit quaternions the reference would have used.  NumPy, vectorised.
"""
import numpy as np


def rot_to_quat(R):
    """(...,3,3) -> (...,4) [x,y,z,w], Markley's method as in scipy Rotation.from_matrix."""
    R = np.asarray(R, dtype=np.float64)
    lead = R.shape[:-2]
    M = R.reshape(-1, 3, 3)
    n = M.shape[0]
    dec = np.empty((n, 4))
    dec[:, :3] = M[:, [0, 1, 2], [0, 1, 2]]
    dec[:, 3] = dec[:, :3].sum(axis=1)
    choice = dec.argmax(axis=1)
    q = np.empty((n, 4))
    idx = np.nonzero(choice != 3)[0]
    i = choice[idx]
    j = (i + 1) % 3
    k = (j + 1) % 3
    q[idx, i] = 1.0 - dec[idx, 3] + 2.0 * M[idx, i, i]
    q[idx, j] = M[idx, j, i] + M[idx, i, j]
    q[idx, k] = M[idx, k, i] + M[idx, i, k]
    q[idx, 3] = M[idx, k, j] - M[idx, j, k]
    idx = np.nonzero(choice == 3)[0]
    q[idx, 0] = M[idx, 2, 1] - M[idx, 1, 2]
    q[idx, 1] = M[idx, 0, 2] - M[idx, 2, 0]
    q[idx, 2] = M[idx, 1, 0] - M[idx, 0, 1]
    q[idx, 3] = 1.0 + dec[idx, 3]
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.reshape(lead + (4,))


def quat_to_rot(q):
    """(...,4) [x,y,z,w] -> (...,3,3), Eigen's toRotationMatrix (what manif `.rotation()` returns)."""
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1.0 - (tyy + tzz); R[..., 0, 1] = txy - twz; R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz; R[..., 1, 1] = 1.0 - (txx + tzz); R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy; R[..., 2, 1] = tyz + twx; R[..., 2, 2] = 1.0 - (txx + tyy)
    return R


def se3_to_rows(T):
    """(...,4,4) -> (...,7) rows quat + position."""
    T = np.asarray(T, dtype=np.float64)
    return np.concatenate((rot_to_quat(T[..., :3, :3]), T[..., :3, 3]), axis=-1)


def rows_to_se3(rows):
    """(...,7) -> (...,4,4)."""
    rows = np.asarray(rows, dtype=np.float64)
    T = np.zeros(rows.shape[:-1] + (4, 4))
    T[..., :3, :3] = quat_to_rot(rows[..., :4])
    T[..., :3, 3] = rows[..., 4:7]
    T[..., 3, 3] = 1.0
    return T


def pose_rows(kind_is_so3, q):
    """Stack of pose matrices (or already-packed rows) -> device pose rows."""
    q = np.asarray(q, dtype=np.float64)
    if kind_is_so3:
        if q.shape[-2:] == (3, 3):
            return rot_to_quat(q)
        if q.shape[-1] == 4:
            return q / np.linalg.norm(q, axis=-1, keepdims=True)
        raise ValueError("SO3 pose must be a 3x3 matrix or a quaternion [x,y,z,w]")
    if q.shape[-2:] == (4, 4):
        return se3_to_rows(q)
    if q.shape[-1] == 7:
        out = q.copy()
        out[..., :4] /= np.linalg.norm(out[..., :4], axis=-1, keepdims=True)
        return out
    raise ValueError("SE3 pose must be a 4x4 matrix or a row quat[x,y,z,w]+position")
