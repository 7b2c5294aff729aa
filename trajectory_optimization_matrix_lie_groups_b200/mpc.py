"""Receding-horizon (MPC) driver over the batched solver.

The reference solves single horizons only (its thesis title is MPC; README.md:5) and its
augmented-Lagrangian loop cold-starts every outer iteration (traopt_controller.py:3237).  The natural
consumer of a batched solver is a receding-horizon loop over B closed-loop systems at once:

    for t in 0 .. T-1:
        reference window  <- (q_ref, xi_ref)[t : t+N+1]
        warm start        <- previous controls shifted by one stage (last one repeated)
        solve             <- a few DDP iterations (the solver's own fit(), truncated by n_iterations)
        apply             <- u_t = us[0];  x_{t+1} = f(x_t, u_t)   (the same discrete dynamics, on the GPU)

Everything numeric (solves and the plant step) runs in the CUDA library; this module only moves
windows and shifts controls.  Multiple shooting restarts its shooting nodes from the reference
window, exactly like `_initial_guess` (traopt_controller.py:3123-3136); single shooting rolls out the
warm-started controls.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import layout
from .solver import BatchSolver


@dataclass
class MPCResult:
    xs: np.ndarray        # (B, T+1, NS) closed-loop state rows
    us: np.ndarray        # (B, T, NU) applied controls
    J: np.ndarray         # (T, B) cost of the plan at each step
    iters: np.ndarray     # (T, B) DDP iterations spent at each step
    status: np.ndarray    # (T, B)


def receding_horizon(kind, method, *, q_ref, xi_ref, x0_rows, N, T, dt, Ib, mass, Q, R, P, n_iterations=3,
                     tol_grad_norm=1e-9, tol_d_norm=1e-6, rollout="nonlinear", warm_start=True, device=None,
                     plant_disturbance=None, **params):
    """Closed-loop tracking of (q_ref, xi_ref) over T steps with horizon N for B systems at once.

    q_ref: (>= T+N+1, 4, 4) or (.., 3, 3) poses; x0_rows: (B, NS) device state rows.
    plant_disturbance(t, x_rows ndarray) -> x_rows, applied to the plant state after each step (optional).
    """
    q_ref = np.asarray(q_ref, dtype=np.float64)
    xi_ref = np.asarray(xi_ref, dtype=np.float64)
    if q_ref.shape[0] < T + N + 1:
        raise ValueError(f"reference has {q_ref.shape[0]} samples, needs T + N + 1 = {T + N + 1}")
    x = np.ascontiguousarray(x0_rows, dtype=np.float64)
    B = x.shape[0]
    s = BatchSolver(kind, method, N, B, device=device)
    s.set_params(dt=dt, Ib=Ib, mass=mass, Q=Q, R=R, P=P, max_iters=n_iterations, tol_grad_norm=tol_grad_norm,
                 tol_d_norm=tol_d_norm, rollout=rollout, **params)
    ref_rows = layout.pose_rows(kind in ("so3", "pendulum"), q_ref)
    xs_cl = np.empty((B, T + 1, s.NS))
    us_cl = np.empty((B, T, s.NU))
    J = np.empty((T, B))
    iters = np.empty((T, B), dtype=np.int32)
    status = np.empty((T, B), dtype=np.int32)
    xs_cl[:, 0] = x
    us_warm = None
    for t in range(T):
        s.set_reference(ref_rows[t:t + N + 1], xi_ref[t:t + N + 1])
        out = s.solve(x, us_warm, trajectories=True)
        us_plan = out["us"]                                   # (B, N, NU) device
        u0 = us_plan[:, 0, :].contiguous()
        # plant: the same exact discrete dynamics, evaluated by the library on the B current states
        nxt = s.stage_eval(0, x, u0.cpu().numpy(), want=("f",))["f"].cpu().numpy()
        if plant_disturbance is not None:
            nxt = np.ascontiguousarray(plant_disturbance(t, nxt), dtype=np.float64)
            nxt[:, :4] /= np.linalg.norm(nxt[:, :4], axis=1, keepdims=True)
        us_cl[:, t] = u0.cpu().numpy()
        J[t] = out["J"].cpu().numpy()
        iters[t] = out["iters"].cpu().numpy()
        status[t] = out["status"].cpu().numpy()
        if warm_start:
            us_warm = torch.cat((us_plan[:, 1:, :], us_plan[:, -1:, :]), dim=1).contiguous()
        x = nxt
        xs_cl[:, t + 1] = x
    s.close()
    return MPCResult(xs_cl, us_cl, J, iters, status)
