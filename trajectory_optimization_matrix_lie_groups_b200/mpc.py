"""Receding-horizon (MPC) driver over the batched solver.

The reference solves single horizons only (its thesis title is MPC; README.md:5) and its
augmented-Lagrangian loop cold-starts every outer iteration (traopt_controller.py:3237).  The natural
consumer of a batched solver is a receding-horizon loop over B closed-loop systems at once:

    for t in 0 .. T-1:
        reference window  <- (q_ref, xi_ref)[t : t+N+1]
        warm start        <- previous controls shifted by one stage (last one repeated)
        solve             <- a few DDP iterations (the solver's own fit(), truncated by n_iterations)
        apply             <- u_t = us[0];  x_{t+1} = f(x_t, u_t)   (the same discrete dynamics, on the GPU)

Everything numeric (solves and the plant step) runs in the CUDA library; this module only slides the
reference window and shifts controls, all of it on the device.  Multiple shooting restarts its shooting nodes from the reference
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
    seconds: float = 0.0  # wall time of the closed loop (T steps of B systems), results still on the device

    @property
    def steps_per_second(self):
        """closed-loop control steps per second, summed over the batch (one step = one truncated solve + one plant step)"""
        return self.us.shape[0] * self.us.shape[1] / self.seconds if self.seconds > 0 else float("nan")


def receding_horizon(kind, method, *, q_ref, xi_ref, x0_rows, N, T, dt, Ib, mass, Q, R, P, n_iterations=3,
                     tol_grad_norm=1e-9, tol_d_norm=1e-6, rollout="nonlinear", warm_start=True, device=None,
                     plant_disturbance=None, **params):
    """Closed-loop tracking of (q_ref, xi_ref) over T steps with horizon N for B systems at once.

    q_ref: (>= T+N+1, 4, 4) or (.., 3, 3) poses; x0_rows: (B, NS) device state rows.
    plant_disturbance(t, x_rows ndarray) -> x_rows, applied to the plant state after each step (optional; the only thing
    that brings the state to the host).

    The loop is device-resident: the whole reference is uploaded once and a window slides over it
    (`trajopt_set_reference_long` / `_offset`), the plant state, the shifted warm start and the logged closed loop stay in
    HBM, and the host sees nothing but the solver's own per-iteration counter until the results are copied at the end.
    """
    import time
    q_ref = np.asarray(q_ref, dtype=np.float64)
    xi_ref = np.asarray(xi_ref, dtype=np.float64)
    if q_ref.shape[0] < T + N + 1:
        raise ValueError(f"reference has {q_ref.shape[0]} samples, needs T + N + 1 = {T + N + 1}")
    s = BatchSolver(kind, method, N, np.shape(x0_rows)[0], device=device)
    dev, B = s.device, s.B
    s.set_params(dt=dt, Ib=Ib, mass=mass, Q=Q, R=R, P=P, max_iters=n_iterations, tol_grad_norm=tol_grad_norm,
                 tol_d_norm=tol_d_norm, rollout=rollout, **params)
    s.set_reference_long(layout.pose_rows(kind in ("so3", "pendulum"), q_ref[:T + N + 1]), xi_ref[:T + N + 1])
    x = torch.as_tensor(np.ascontiguousarray(x0_rows, dtype=np.float64), device=dev)
    xs_cl = torch.empty((B, T + 1, s.NS), dtype=torch.float64, device=dev)
    us_cl = torch.empty((B, T, s.NU), dtype=torch.float64, device=dev)
    J = torch.empty((T, B), dtype=torch.float64, device=dev)
    iters = torch.empty((T, B), dtype=torch.int32, device=dev)
    status = torch.empty((T, B), dtype=torch.int32, device=dev)
    xs_cl[:, 0] = x
    us_warm = None
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for t in range(T):
        s.set_reference_offset(t)
        out = s.solve(x, us_warm, trajectories=True)
        us_plan = out["us"]                                   # (B, N, NU) device
        u0 = us_plan[:, 0, :].contiguous()
        # plant: the same exact discrete dynamics, evaluated by the library on the B current states
        nxt = s.stage_eval(0, x, u0, want=("f",))["f"]
        if plant_disturbance is not None:
            h = np.ascontiguousarray(plant_disturbance(t, nxt.cpu().numpy()), dtype=np.float64)
            h[:, :4] /= np.linalg.norm(h[:, :4], axis=1, keepdims=True)
            nxt = torch.as_tensor(h, device=dev)
        us_cl[:, t] = u0
        J[t], iters[t], status[t] = out["J"], out["iters"], out["status"]
        if warm_start:
            us_warm = torch.cat((us_plan[:, 1:, :], us_plan[:, -1:, :]), dim=1).contiguous()
        x = nxt
        xs_cl[:, t + 1] = x
    torch.cuda.synchronize(dev)
    seconds = time.perf_counter() - t0
    res = MPCResult(xs_cl.cpu().numpy(), us_cl.cpu().numpy(), J.cpu().numpy(), iters.cpu().numpy(), status.cpu().numpy(), seconds)
    s.close()
    return res
