"""GPU parity, level 1: the Lie-group closed forms of csrc/lie.cuh against the oracle (oracle/lie.py),
called through the C ABI (`trajopt_debug_lie`).

Tolerances (absolute, FP64).  manif (and therefore the oracle) switches from the closed forms to a
two-term series only below theta^2 = 1e-10; just above that threshold its closed forms cancel:
`(1 - cos th)/th^2` carries a relative error of ~1e-16/th^2 and `(1 - th^2/2 - cos th)/th^4` one of
~1e-16/th^4.  The device keeps longer series up to a larger angle (it is the more accurate of the
two there), so the two agree to ~1e-11 on the SO(3) Jacobians / SE(3) exp-log and to ~1e-10 on the
SE(3) Jacobian block Q(omega, v) (and on Jr, Jr^-1 which contain it) — far below the 1e-9 / 1e-7
bars of the solver-level parity tests, which are the contract.  Asserted here: 1e-10 and 1e-9.
"""
import numpy as np
import pytest

from oracle import lie

pytestmark = pytest.mark.gpu


def _angles(n_per=16, seed=0):
    """rotation vectors covering the series branch, the cancellation zone, mid angles and near-pi."""
    rng = np.random.default_rng(seed)
    scales = np.concatenate([np.full(n_per, s) for s in (1e-9, 1e-7, 1e-5, 1e-3, 0.3, 0.9, 1.7)])
    w = rng.standard_normal((scales.size, 3)) * scales[:, None]
    near_pi = rng.standard_normal((n_per, 3))
    near_pi *= ((np.pi - 10.0 ** -rng.uniform(1, 6, n_per)) / np.linalg.norm(near_pi, axis=1))[:, None]
    return np.concatenate((w, near_pi, np.zeros((1, 3))))


def _op(name, x):
    from trajectory_optimization_matrix_lie_groups_b200 import lie_op
    return lie_op(name, x).cpu().numpy()


@pytest.mark.parametrize("name,fn,tol", [
    ("so3_exp", lie.so3_exp, 1e-14),
    ("so3_jr", lambda a: lie.so3_jr(a).ravel(), 1e-10),
    ("so3_jr_inv", lambda a: lie.so3_jr_inv(a).ravel(), 1e-10),
    ("so3_jl", lambda a: lie.so3_jl(a).ravel(), 1e-10),
    ("so3_jl_inv", lambda a: lie.so3_jl_inv(a).ravel(), 1e-10),
])
def test_so3_maps(name, fn, tol):
    w = _angles()
    if "inv" in name:
        w = w[np.linalg.norm(w, axis=1) < 3.0]     # Jr^-1 has a pole factor (1 + cos)/sin near pi: compare away from it
    got = _op(name, w)
    ref = np.stack([fn(r) for r in w])
    assert np.max(np.abs(got - ref)) < tol * max(1.0, np.max(np.abs(ref)))


def test_so3_log_roundtrip_and_oracle():
    w = _angles()
    w = w[np.linalg.norm(w, axis=1) < np.pi]       # Log(Exp(w)) = w only inside the injectivity radius
    q = np.stack([lie.so3_exp(r) for r in w])
    got = _op("so3_log", q)
    ref = np.stack([lie.so3_log(r) for r in q])
    assert np.max(np.abs(got - ref)) < 1e-13
    assert np.max(np.abs(got - w)) < 1e-9          # near pi the round trip itself is conditioned like 1/(pi - th)
    # antipodal quaternion: same rotation, same log (w >= 0 canonicalisation of the reference's scipy/manif path)
    got2 = _op("so3_log", -q)
    mid = np.linalg.norm(w, axis=1) < 3.0
    assert np.max(np.abs(got2[mid] - got[mid])) < 1e-13


@pytest.mark.parametrize("name,fn,tol", [
    ("se3_exp", lambda a: np.concatenate(lie.se3_exp(a)), 1e-10),
    ("se3_Q", lambda a: lie.se3_Q(a[:3], a[3:]).ravel(), 1e-9),
    ("se3_jr", lambda a: lie.se3_jr(a).ravel(), 1e-9),
    ("se3_jr_inv", lambda a: lie.se3_jr_inv(a).ravel(), 1e-9),
])
def test_se3_maps(name, fn, tol):
    w = _angles(seed=1)
    if "inv" in name:
        w = w[np.linalg.norm(w, axis=1) < 3.0]
    rng = np.random.default_rng(2)
    tau = np.concatenate((w, rng.standard_normal((w.shape[0], 3))), axis=1)
    got = _op(name, tau)
    ref = np.stack([fn(r) for r in tau])
    assert np.max(np.abs(got - ref)) < tol * max(1.0, np.max(np.abs(ref)))


def test_se3_group_operations():
    w = _angles(seed=3)
    w = w[np.linalg.norm(w, axis=1) < 3.0]
    rng = np.random.default_rng(4)
    tau = np.concatenate((w, rng.standard_normal((w.shape[0], 3))), axis=1)
    qp = np.stack([np.concatenate(lie.se3_exp(r)) for r in tau])
    got = _op("se3_log", qp)
    ref = np.stack([lie.se3_log(r[:4], r[4:]) for r in qp])
    assert np.max(np.abs(got - ref)) < 1e-10
    got = _op("se3_adj", qp)
    ref = np.stack([lie.se3_adj(r[:4], r[4:]).ravel() for r in qp])
    assert np.max(np.abs(got - ref)) < 1e-13
    ab = np.concatenate((qp, np.roll(qp, 1, axis=0)), axis=1)
    ops = {
        "se3_compose": lambda a: np.concatenate(lie.se3_compose(a[:4], a[4:7], a[7:11], a[11:])),
        "se3_rminus": lambda a: lie.se3_log(*lie.se3_compose(*lie.se3_inverse(a[7:11], a[11:]), a[:4], a[4:7])),
        "se3_lminus": lambda a: lie.se3_log(*lie.se3_compose(a[:4], a[4:7], *lie.se3_inverse(a[7:11], a[11:]))),
    }
    for name, fn in ops.items():
        got = _op(name, ab)
        ref = np.stack([fn(r) for r in ab])
        err = np.abs(got - ref)
        if name == "se3_compose":      # quaternion sign is free
            err = np.minimum(err, np.abs(np.concatenate((-got[:, :4], got[:, 4:]), axis=1) - ref))
        assert np.max(err) < 1e-10, name


def test_empty_and_bad_arguments():
    import torch
    from trajectory_optimization_matrix_lie_groups_b200 import lie_op, _lib
    out = lie_op("so3_exp", np.zeros((0, 3)))
    assert out.shape == (0, 4)
    t = torch.zeros(1, 3, dtype=torch.float64, device="cuda")
    rc = _lib.lib.trajopt_debug_lie(999, 1, t.data_ptr(), t.data_ptr(), None)
    assert rc == -1 and b"unknown op" in _lib.lib.trajopt_last_error()
