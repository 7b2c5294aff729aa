"""The reference's data formats on either side of the solver.

* Reference trajectories: `visualization/optimized_trajectories/*.npy` are written with consecutive
  `np.save` calls into ONE file — `q_ref` (N+1,4,4) or (N+1,3,3), `xi_ref` (N+1,6|3) and, in most
  files, a scalar `dt` — and read back with consecutive `np.load` calls on the open file
  (main_SE3ddp_tracking_exact_ms.py:108-110, benchmark_SO3_tracking.py:48-57).
* Result files: the `pickle` written by `save_results_pickle` (benchmark_SE3_tracking.py:272-327):
  `{'prob': {J, dt, q_ref, xi_ref, x0, Q, P, R}, '<tag>': {xs, us, J_hist, grad_hist[, defect_hist]}, ...}`
  with `xs` a list of `[q, xi]` states.  These are the files the golden fixtures were re-packed from.
"""
import pickle

import numpy as np


def load_reference_trajectory(path):
    """-> (q_ref, xi_ref, dt or None) from a reference-format .npy file."""
    with open(path, "rb") as f:
        q_ref = np.load(f, allow_pickle=False)
        xi_ref = np.load(f, allow_pickle=False)
        try:
            dt = float(np.load(f, allow_pickle=False))
        except (EOFError, ValueError, OSError):
            dt = None                      # path_dense_random_columns.npy carries no dt
    q_ref = np.asarray(q_ref, dtype=np.float64)
    xi_ref = np.asarray(xi_ref, dtype=np.float64)
    if q_ref.ndim != 3 or q_ref.shape[1:] not in ((4, 4), (3, 3)):
        raise ValueError(f"{path}: q_ref must be (N+1,4,4) or (N+1,3,3), got {q_ref.shape}")
    if xi_ref.shape != (q_ref.shape[0], 6 if q_ref.shape[1] == 4 else 3):
        raise ValueError(f"{path}: xi_ref shape {xi_ref.shape} does not match q_ref {q_ref.shape}")
    return q_ref, xi_ref, dt


def save_reference_trajectory(path, q_ref, xi_ref, dt=None):
    """Write (q_ref, xi_ref[, dt]) the way the reference's conversion notebook does."""
    with open(path, "wb") as f:
        np.save(f, np.asarray(q_ref, dtype=np.float64))
        np.save(f, np.asarray(xi_ref, dtype=np.float64))
        if dt is not None:
            np.save(f, np.float64(dt))


def _plain_state(x):
    q, xi = x
    q = q.rotation() if hasattr(q, "rotation") else np.asarray(q, dtype=np.float64)
    xi = xi.coeffs() if hasattr(xi, "coeffs") else np.asarray(xi, dtype=np.float64)
    return [np.array(q, dtype=np.float64), np.array(xi, dtype=np.float64).reshape(-1)]


def save_results_pickle(filename, prob, **runs):
    """prob: dict with J, dt, q_ref, xi_ref, x0, Q, P, R;  runs: tag -> dict(xs, us, J_hist, grad_hist[, defect_hist]).

    States are stored as plain `[ndarray pose, ndarray velocity]` pairs, so the file loads without manifpy."""
    data = {"prob": dict(prob)}
    data["prob"]["x0"] = _plain_state(prob["x0"])
    for tag, r in runs.items():
        entry = {"xs": [_plain_state(x) for x in r["xs"]], "us": np.asarray(r["us"], dtype=np.float64),
                 "J_hist": list(map(float, r["J_hist"])), "grad_hist": list(map(float, r["grad_hist"]))}
        if "defect_hist" in r:
            entry["defect_hist"] = list(map(float, r["defect_hist"]))
        data[tag] = entry
    with open(filename, "wb") as f:
        pickle.dump(data, f)
    return data


def load_results_pickle(filename):
    """Read a result file of the reference (or of `save_results_pickle`)."""
    with open(filename, "rb") as f:
        return pickle.load(f)


def batch_results_to_runs(result, J_hist=None):
    """BatchResult of `fit_batch` -> list of per-problem run dicts in the result-file layout."""
    out = []
    for b in range(result.J.shape[0]):
        lo, hi = result.shard
        local = lo <= b < hi and result.xs_rows is not None
        n = int(result.iters[b])
        out.append({
            "xs": result.states(b - lo) if local else [],
            "us": result.us[b - lo] if local else np.zeros((0, 0)),
            "J_hist": (result.J_hist[b, :n] if result.J_hist is not None else [result.J[b]]),
            "grad_hist": (result.grad_hist[b, :n] if result.grad_hist is not None else [result.grad[b]]),
            "defect_hist": (result.defect_hist[b, :n + 1] if result.defect_hist is not None else [result.defect[b]]),
        })
    return out


# ---------------------------------------------------------------------------------------------
# Reference-trajectory ingestion: what visualization/convert_path_to_reference.ipynb does before the
# solver ever runs.  Host-side NumPy/SciPy: this is data preparation, not the path.
# ---------------------------------------------------------------------------------------------
CSV_POSITION = ("p_x", "p_y", "p_z")
CSV_QUATERNION = ("q_w", "q_x", "q_y", "q_z")          # scalar first in the planner's CSV
CSV_LINEAR_VELOCITY = ("v_x", "v_y", "v_z")            # world frame in the file
CSV_ANGULAR_VELOCITY = ("w_x", "w_y", "w_z")           # body frame in the file


def reference_from_csv(path, dt=None):
    """Planner CSV (`t,p_*,q_*,v_*,w_*,...`, e.g. path_dense_random_columns_4obj.csv) -> (q_ref, xi_ref, dt).

    The notebook's conversion (cells "file_name = ...path_dense_random_columns_4obj" onward): pose from the
    scalar-first quaternion and the position, twist [w, R^T v_world] (angular first, linear in the body frame).
    `dt` is not derived from the `t` column by the notebook (it hard-codes 0.004); pass it, or get the median step.
    """
    import csv

    from scipy.spatial.transform import Rotation
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    head = [h.strip() for h in rows[0]]
    data = np.array([[float(v) for v in r[:len(head)]] for r in rows[1:] if r], dtype=np.float64)
    col = {h: i for i, h in enumerate(head)}
    missing = [c for c in CSV_POSITION + CSV_QUATERNION + CSV_LINEAR_VELOCITY + CSV_ANGULAR_VELOCITY if c not in col]
    if missing:
        raise ValueError(f"{path}: missing columns {missing}")
    take = lambda names: data[:, [col[n] for n in names]]
    pos, quat, v_world, w = take(CSV_POSITION), take(CSV_QUATERNION), take(CSV_LINEAR_VELOCITY), take(CSV_ANGULAR_VELOCITY)
    rot = Rotation.from_quat(quat, scalar_first=True).as_matrix()
    n = data.shape[0]
    q_ref = np.zeros((n, 4, 4))
    q_ref[:, :3, :3] = rot
    q_ref[:, :3, 3] = pos
    q_ref[:, 3, 3] = 1.0
    xi_ref = np.empty((n, 6))
    xi_ref[:, :3] = w
    # the notebook's per-row `romt[i].T @ v_world[i]` (a (3,3)x(3,1) matmul): same left-to-right dot products
    xi_ref[:, 3:] = np.stack([rot[i].T @ v_world[i] for i in range(n)])
    if dt is None and "t" in col and n > 1:
        dt = float(np.median(np.diff(data[:, col["t"]])))
    return q_ref, xi_ref, dt


def twist_integrated_reference(q0, twists, dt):
    """Reference generated by integrating body twists: q_{k+1} = q_k expm(hat(xi_k) dt)   (the notebook's
    `path_se3_generate_sine*` / `path_se3_spiral_static_velocity` cells; also the helix of BASELINE config 4).

    twists: (N, 6) twist applied over step k, [w, v].  Returns q_ref (N+1, 4, 4).
    """
    from scipy.linalg import expm
    twists = np.asarray(twists, dtype=np.float64)
    q = np.asarray(q0, dtype=np.float64).copy()
    out = [q.copy()]
    for xi in twists:
        H = np.zeros((4, 4))
        H[:3, :3] = [[0.0, -xi[2], xi[1]], [xi[2], 0.0, -xi[0]], [-xi[1], xi[0], 0.0]]
        H[:3, 3] = xi[3:]
        q = q @ expm(H * dt)
        out.append(q.copy())
    return np.stack(out)


def so3_reference_from_rotations(rotations, dt):
    """SO(3) reference from sampled attitudes: the velocity is the finite difference on the group,
    w_i = Log(R_i^T R_{i+1}) / dt (manif's `(R_{i+1} - R_i) / dt`, right-minus), last one repeated
    (the notebook's `path_3dpendulum_8shape*` cells).  -> (q_ref (N,3,3), omega_ref (N,3))."""
    from scipy.spatial.transform import Rotation
    R = np.asarray(rotations, dtype=np.float64)
    rel = Rotation.from_matrix(np.einsum("nji,njk->nik", R[:-1], R[1:]))
    w = rel.as_rotvec() / dt
    return R, np.concatenate((w, w[-1:]), axis=0)


def eight_shape_rotations(period=10.0, dt=0.04, amp_x=np.pi / 3, amp_y=np.pi / 2, periods=1):
    """The pendulum's 8-shape attitude samples: extrinsic 'xy' Euler angles A sin(wt), B sin(2wt)."""
    from scipy.spatial.transform import Rotation
    n = int(periods * period / dt)
    t = np.linspace(0, periods * period, n)
    om = 2 * np.pi / period
    ang = np.stack((amp_x * np.sin(om * t), amp_y * np.sin(2 * om * t)), axis=1)
    return Rotation.from_euler("xy", ang, degrees=False).as_matrix()
