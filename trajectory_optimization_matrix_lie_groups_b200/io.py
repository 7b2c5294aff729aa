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
