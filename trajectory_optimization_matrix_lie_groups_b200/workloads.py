"""The BASELINE.json configurations as concrete, seeded, synthetic batches.

Each builder returns a `Workload`: the problem definition the reference script sets up plus a batch
of initial states in the "perturb_all_compute shape" (visualization/perturb_all_compute.py:181-194
in the reference): problem b perturbs ONE of the 12 initial-state parameters, j = b mod 12, of the
script's nominal x0 — {th_z, th_y, th_x (right-multiplied Euler 'zyx', degrees), w_x, w_y, w_z,
p_x, p_y, p_z, v_x, v_y, v_z} — by a value drawn uniformly from +-10 % of the reference sweep's
half-width.  Problem 0 is the script's own unperturbed problem.  RNG: numpy default_rng(24234156),
the seed the scripts declare (e.g. main_SE3ddp_tracking_exact.py:22).

Reference trajectories are package data (`data/*.npy`, in the reference's own on-disk format and under the file
names its scripts load from visualization/optimized_trajectories/; written by scripts/make_workload_data.py and
read with `io.load_reference_trajectory`); neither the reference tree nor the test fixtures are needed at run time.
"""
import os
from dataclasses import dataclass, field

import numpy as np
from scipy.spatial.transform import Rotation

from . import layout
from .io import load_reference_trajectory

SEED = 24234156
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# half-widths of the reference sweep (perturb_all_compute.py:181-194): th 30 deg, w_x 3, w_y 2, w_z 1, p 50, v 10
SWEEP_HALF_WIDTH = np.array([30.0, 30.0, 30.0, 3.0, 2.0, 1.0, 50.0, 50.0, 50.0, 10.0, 10.0, 10.0])


@dataclass
class Workload:
    name: str
    kind: str                 # 'so3' | 'se3' | 'drone'
    method: str               # 'ss' | 'ms' | 'al_ms'
    N: int
    dt: float
    J: np.ndarray             # generalised inertia as the scripts build it (6x6, or 3x3 for SO3)
    Q: np.ndarray
    R: np.ndarray
    P: np.ndarray
    q_ref: np.ndarray         # (N+1, 4, 4) or (N+1, 3, 3)
    xi_ref: np.ndarray        # (N+1, 6) or (N+1, 3)
    x0_rows: np.ndarray       # (B, NS) device state rows
    solver: dict = field(default_factory=dict)     # fit() keyword arguments / controller options
    bounds: tuple = None      # (lb, ub) for the AL config

    @property
    def B(self):
        return self.x0_rows.shape[0]

    @property
    def Ib(self):
        return self.J[:3, :3]

    @property
    def mass(self):
        return 1.0 if self.kind in ("so3", "pendulum") else float(self.J[4, 4])

    def make_solver(self, B=None, device=None, offset=0):
        """BatchSolver configured for problems [offset, offset+B) of this workload, and their x0 rows."""
        from .solver import BatchSolver
        B = self.B - offset if B is None else B
        s = BatchSolver(self.kind, self.method, self.N, B, device=device)
        kw = dict(self.solver)
        if self.bounds is not None:
            kw["lb"], kw["ub"] = self.bounds
        s.set_params(dt=self.dt, Ib=self.Ib, mass=self.mass, Q=self.Q, R=self.R, P=self.P, **kw)
        s.set_reference(layout.pose_rows(self.kind in ("so3", "pendulum"), self.q_ref), self.xi_ref)
        return s, self.x0_rows[offset:offset + B]


def _reference(name):
    """(q_ref, xi_ref, dt) of data/<name>.npy"""
    return load_reference_trajectory(os.path.join(DATA_DIR, name + ".npy"))


def _rigid_J(m=1.0):
    J = np.zeros((6, 6))
    J[:3, :3] = np.diag([0.5, 0.7, 0.9])
    J[3:, 3:] = m * np.eye(3)
    return J


def perturb_se3(R_nom, p_nom, xi_nom, B, frac, rng):
    """(B, 13) state rows: one-parameter-at-a-time perturbations of (R_nom, p_nom, xi_nom)."""
    j = np.arange(B) % 12
    val = rng.uniform(-1.0, 1.0, size=B) * (np.asarray(frac) * SWEEP_HALF_WIDTH)[j]
    val[0] = 0.0
    eul = np.zeros((B, 3))
    for a in range(3):
        eul[j == a, a] = val[j == a]
    dR = Rotation.from_euler("zyx", eul, degrees=True).as_matrix()
    R = np.einsum("ij,bjk->bik", R_nom, dR)
    xi = np.tile(np.asarray(xi_nom, dtype=float), (B, 1))
    p = np.tile(np.asarray(p_nom, dtype=float), (B, 1))
    for a in range(3):
        xi[j == 3 + a, a] += val[j == 3 + a]
        p[j == 6 + a, a] += val[j == 6 + a]
        xi[j == 9 + a, 3 + a] += val[j == 9 + a]
    return np.concatenate((layout.rot_to_quat(R), p, xi), axis=1)


def se3_tracking_ss(B=1):
    """cfg 1: main_SE3ddp_tracking_exact.py — SE3 single shooting, N=955, dt=0.01 (script value)."""
    q_ref, xi_ref, _ = _reference("path_dense_random_columns_4obj")
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    rng = np.random.default_rng(SEED)
    x0 = perturb_se3(q_ref[0][:3, :3], q_ref[0][:3, 3], xi_ref[0], B, 0.1, rng)
    return Workload("main_SE3ddp_tracking_exact", "se3", "ss", q_ref.shape[0] - 1, 0.01, _rigid_J(), Q,
                    1e-5 * np.eye(6), 10 * Q, q_ref, xi_ref, x0,
                    solver=dict(max_iters=200, tol_grad_norm=1e-3, rollout="nonlinear"))


def so3_tracking_ms(B=1024, method="ms"):
    """cfg 2: benchmark_SO3_tracking.py — SO3, N=249, dt=0.04, batch of perturbed initial attitudes."""
    q_ref, xi_ref, dt = _reference("path_3dpendulum_8shape_tryout")
    Q = np.diag([10.0, 10, 10, 1, 1, 1])
    rng = np.random.default_rng(SEED)
    R_nom = Rotation.from_euler("zxy", [90.0, 10.0, 45.0], degrees=True).as_matrix()
    delta = 0.3 * rng.standard_normal((B, 3))
    dw = 0.05 * rng.standard_normal((B, 3))
    delta[0] = 0.0
    dw[0] = 0.0
    R = np.einsum("ij,bjk->bik", R_nom, Rotation.from_rotvec(delta).as_matrix())
    w0 = 0.15 * np.ones((B, 3)) + dw
    x0 = np.concatenate((layout.rot_to_quat(R), w0), axis=1)
    return Workload("benchmark_SO3_tracking", "so3", method, q_ref.shape[0] - 1, float(dt),
                    np.diag([0.5, 0.7, 0.9]), Q, 1e-3 * np.eye(3), 1.5 * Q, q_ref, xi_ref, x0,
                    solver=dict(max_iters=50, tol_grad_norm=1e-8, rollout="nonlinear"))


def se3_tracking_ms(B=16384, frac=0.1):
    """cfg 3 (headline): main_SE3ddp_tracking_exact_ms.py — SE3 multiple shooting, N=955, dt=0.004."""
    q_ref, xi_ref, _ = _reference("path_dense_random_columns_4obj")
    Q = np.diag([25.0, 25, 25, 10, 10, 10, 1, 1, 1, 1, 1, 1])
    rng = np.random.default_rng(SEED)
    R_nom = Rotation.from_euler("zxy", [90.0, 10.0, 45.0], degrees=True).as_matrix()
    x0 = perturb_se3(R_nom, q_ref[0][:3, 3] - 1.0, 0.1 * np.ones(6), B, frac, rng)
    return Workload("main_SE3ddp_tracking_exact_ms", "se3", "ms", q_ref.shape[0] - 1, 0.004, _rigid_J(), Q,
                    1e-5 * np.eye(6), 10 * Q, q_ref, xi_ref, x0,
                    solver=dict(max_iters=200, tol_grad_norm=1e-12, tol_d_norm=1e-6, rollout="nonlinear",
                                line_search=False))


def helix_reference(N=1400, dt=0.01):
    """Constant-twist helix of main_SE3ddp_tracking_exact_al_ms.py:59-80: q_{k+1} = q_k expm(xi^ dt)."""
    xi = np.array([0.0, 0.0, 1.0, 2.0, 0.0, 0.2])
    w, v = xi[:3] * dt, xi[3:] * dt
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    Rs = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th**2 * (W @ W)
    Vs = np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * (W @ W)
    E = np.eye(4)
    E[:3, :3] = Rs
    E[:3, 3] = Vs @ v
    q_ref = np.zeros((N + 1, 4, 4))
    q_ref[0] = np.eye(4)
    for i in range(N):
        q_ref[i + 1] = q_ref[i] @ E
    return q_ref, np.tile(xi, (N + 1, 1))


def se3_tracking_al_ms(B=16384, frac=0.02, N=1400):
    """cfg 4: main_SE3ddp_tracking_exact_al_ms.py — AL multiple shooting with input bounds, helix reference.

    Perturbations are +-2 % of the sweep half-width (p +-1, th +-0.6 deg, v +-0.2): with the +-10 % of the other
    configs a handful of problems per thousand drive the reference's penalty scheme to its cap (mu = 1e8) without
    ever meeting tol_constr, so they run all 100 outer x 200 inner iterations (measured: 7 of 2048 problems, 260 s
    for the batch instead of 2 s) and the benchmark would time those stragglers.  `frac=0.1` gives that stress set."""
    q_ref, xi_ref = helix_reference(N, 0.01)
    Q = np.diag([10.0, 10, 10, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    rng = np.random.default_rng(SEED)
    x0 = perturb_se3(np.eye(3), np.array([-1.0, -1.0, -0.2]), np.array([0, 0, 0.1, 2.0, 0, 0.2]), B, frac, rng)
    return Workload("main_SE3ddp_tracking_exact_al_ms", "se3", "al_ms", N, 0.01, _rigid_J(), Q, np.zeros((6, 6)),
                    10 * Q, q_ref, xi_ref, x0, bounds=(-10.0, 10.0),
                    solver=dict(max_iters=200, tol_grad_norm=1e-6, tol_d_norm=1e-6, rollout="nonlinear",
                                n_al_iters=100, tol_constr=1e-2))


def drone_racing_ms(B=1 << 20, method="ms"):
    """cfg 5: benchmark_drone_racing_tracking.py — quadrotor on SE3, N=150, small perturbation ranges."""
    q_ref, xi_ref, _ = _reference("path_dense_random_columns_4obj")
    q_ref, xi_ref = q_ref[:151], xi_ref[:151]          # Nsim = 150 (benchmark_drone_racing_tracking.py:56-58)
    Q = np.diag([25.0, 25, 25, 10, 10, 10, 1, 1, 1, 1, 1, 1])
    rng = np.random.default_rng(SEED)
    R_nom = Rotation.from_euler("zxy", [1e-4, 0.0, 0.0], degrees=True).as_matrix()
    # ranges of SURVEY.md section 8d cfg 5: th +-1 deg, w +-0.1, p +-0.2, v +-0.2
    frac = np.array([1.0, 1, 1, 0.1, 0.1, 0.1, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2]) / SWEEP_HALF_WIDTH
    x0 = perturb_se3(R_nom, q_ref[0][:3, 3] - 0.1, 1e-3 * np.ones(6), B, frac, rng)
    return Workload("benchmark_drone_racing_tracking", "drone", method, q_ref.shape[0] - 1, 0.004,
                    _rigid_J(), Q, 1e-5 * np.eye(4), 1.5 * Q, q_ref, xi_ref, x0,
                    solver=dict(max_iters=200, tol_grad_norm=1e-12, tol_d_norm=1e-6, rollout="nonlinear"))


CONFIGS = {1: se3_tracking_ss, 2: so3_tracking_ms, 3: se3_tracking_ms, 4: se3_tracking_al_ms, 5: drone_racing_ms}

# algorithmic work per stage-iteration: dense FLOPs and SoA FP64 HBM bytes
#   W_stage = W_backward + W_linearise + n_alpha W_forward;  bytes = backward + n_alpha forward
WORK_MODEL = {
    "se3": dict(flop_bwd=19140.0, flop_lin=3200.0, flop_fwd=2300.0, bytes_bwd=776.0, bytes_fwd=928.0),
    "drone": dict(flop_bwd=14872.0, flop_lin=3400.0, flop_fwd=2400.0, bytes_bwd=552.0, bytes_fwd=688.0),
    "so3": dict(flop_bwd=2730.0, flop_lin=600.0, flop_fwd=500.0, bytes_bwd=248.0, bytes_fwd=328.0),
    "rigid": dict(flop_bwd=19140.0, flop_lin=3400.0, flop_fwd=2400.0, bytes_bwd=776.0, bytes_fwd=928.0),
    "pendulum": dict(flop_bwd=2730.0, flop_lin=800.0, flop_fwd=600.0, bytes_bwd=248.0, bytes_fwd=328.0),
}
# doubles the linearisation kernel writes per stage-iteration as designed (record + G_i + stage cost + defect norm)
LIN_RECORD_DOUBLES = {"se3": 114 + 13 + 2, "rigid": 117 + 13 + 2, "drone": 113 + 13 + 2, "so3": 54 + 7 + 2, "pendulum": 72 + 7 + 2}
