"""Minimal stand-ins for the manifpy types that cross the reference's API.

The reference hands `manifpy.SO3` / `manifpy.SO3Tangent` objects to and from the SO3 controllers
(benchmark_SO3_tracking.py:67-79, 194-201: `x[0].rotation()`, `x[1].coeffs()`).  manifpy is a
third-party C++ extension that is not part of this package; these light objects expose the
accessors the scripts use and nothing else.  Arithmetic stays in the CUDA library.
"""
import numpy as np

from .. import layout


class SO3:
    """Unit quaternion [x, y, z, w] (manif coefficient order)."""

    def __init__(self, *args):
        if len(args) == 1:
            q = np.asarray(args[0], dtype=float).reshape(-1)
        elif len(args) == 4:
            q = np.array(args, dtype=float)
        else:
            raise TypeError("SO3(quat[x,y,z,w]) or SO3(x, y, z, w)")
        if q.shape != (4,):
            raise ValueError("SO3 needs a quaternion [x, y, z, w]")
        self._q = q / np.linalg.norm(q)

    @classmethod
    def from_matrix(cls, R):
        return cls(layout.rot_to_quat(np.asarray(R, dtype=float)))

    @classmethod
    def Identity(cls):
        return cls(np.array([0.0, 0.0, 0.0, 1.0]))

    def coeffs(self):
        return self._q.copy()

    def quat(self):
        return self._q.copy()

    def rotation(self):
        return layout.quat_to_rot(self._q)

    def __repr__(self):
        return f"SO3(quat={self._q})"


class SO3Tangent:
    def __init__(self, w):
        self._w = np.asarray(w, dtype=float).reshape(3).copy()

    def coeffs(self):
        return self._w.copy()

    def copy(self):
        return SO3Tangent(self._w)

    def __array__(self, dtype=None, copy=None):
        return self._w.astype(dtype) if dtype is not None else self._w.copy()

    def __repr__(self):
        return f"SO3Tangent({self._w})"


def so3_quat(q):
    """Anything the SO3 API accepts as a rotation -> unit quaternion [x,y,z,w]."""
    if hasattr(q, "coeffs"):
        c = np.asarray(q.coeffs(), dtype=float).reshape(-1)
        return c / np.linalg.norm(c)
    q = np.asarray(q, dtype=float)
    if q.shape == (3, 3):
        return layout.rot_to_quat(q)
    if q.shape == (4,):
        return q / np.linalg.norm(q)
    raise ValueError("SO3 state must be a manif-like SO3, a 3x3 matrix or a quaternion [x,y,z,w]")


def so3_vel(w):
    if hasattr(w, "coeffs"):
        return np.asarray(w.coeffs(), dtype=float).reshape(3)
    return np.asarray(w, dtype=float).reshape(3)


class SE3:
    """Pose as (position, unit quaternion [x, y, z, w]) — what `SE32manifSE3` builds (traopt_utilis.py:331-342)."""

    def __init__(self, position=None, quaternion=None):
        p = np.zeros(3) if position is None else np.asarray(position, dtype=float).reshape(3)
        q = np.array([0.0, 0.0, 0.0, 1.0]) if quaternion is None else np.asarray(quaternion, dtype=float).reshape(4)
        self._p = p.copy()
        self._q = q / np.linalg.norm(q)

    def translation(self):
        return self._p.copy()

    def quat(self):
        return self._q.copy()

    def coeffs(self):
        return np.concatenate((self._p, self._q))            # manif order: translation, then quaternion

    def rotation(self):
        return layout.quat_to_rot(self._q)

    def transform(self):
        return layout.rows_to_se3(np.concatenate((self._q, self._p)))

    def __repr__(self):
        return f"SE3(position={self._p}, quaternion={self._q})"


class SE3Tangent:
    """Twist in manif's order [v, omega] (the library's own order is [omega, v]: see se32manifse3)."""

    def __init__(self, coeffs):
        self._c = np.asarray(coeffs, dtype=float).reshape(6).copy()

    def coeffs(self):
        return self._c.copy()

    def __mul__(self, s):
        return SE3Tangent(self._c * float(s))

    __rmul__ = __mul__

    def __repr__(self):
        return f"SE3Tangent({self._c})"
