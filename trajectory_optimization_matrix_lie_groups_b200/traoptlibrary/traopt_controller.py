"""Controllers with the reference's interface (traoptlibrary/traopt_controller.py), driving the
native CUDA solver.

    iLQR_Tracking_SO3        :526-1026     single shooting + 13-step line search
    iLQR_Tracking_SO3_MS     :1029-1826    multiple shooting
    iLQR_Tracking_SE3        :1831-2349    single shooting (SE3 / drone dynamics)
    iLQR_Tracking_SE3_MS     :2352-3136    multiple shooting
    AL_iLQR_Tracking_SE3_MS  :3139-3293    augmented-Lagrangian outer loop around SE3_MS

`fit` keeps the reference's signature, return tuple and callback protocol (the history lists are
filled by the caller's `on_iteration`, exactly like the reference, except `grad_hist` which the
reference's single-shooting `fit` appends itself).  `fit_batch` is new: B initial states solved
together on the GPU, optionally sharded over the ranks of a torch.distributed job.
"""
import warnings
from dataclasses import dataclass, field

import numpy as np
import torch

from .. import _lib
from .. import distributed as _dist
from . import _native


class BaseController:
    """Controller interface (traopt_controller.py:14-32)."""

    def fit(self, x0, us_init, *args, **kwargs):
        raise NotImplementedError


class PDViolationError(Exception):
    """Defined by the reference (traopt_controller.py:35) and never raised there either."""


@dataclass
class BatchResult:
    """Per-problem outcome of `fit_batch` (NumPy, problem-major)."""
    kind: str
    J: np.ndarray
    iters: np.ndarray
    status: np.ndarray                 # low 4 bits: 0 converged, 1 max-iter, 2 no-descent; flags 16 / 32
    grad: np.ndarray
    defect: np.ndarray
    xs_rows: np.ndarray = None         # (b, N+1, NS) device rows of this rank's shard (None if not requested)
    us: np.ndarray = None              # (b, N, NU)
    shard: tuple = (0, 0)              # [lo, hi) of xs_rows / us inside the global batch
    J_hist: np.ndarray = None
    grad_hist: np.ndarray = None
    defect_hist: np.ndarray = None
    alpha_hist: np.ndarray = None
    extra: dict = field(default_factory=dict)

    def states(self, b):
        """Trajectory of local problem b as the reference's list of [q, xi]."""
        return _native.rows_states(self.kind, self.xs_rows[b])

    @property
    def converged(self):
        return (self.status & 15) == _lib.STATUS_CONVERGED


class _NativeController(BaseController):
    METHOD = None
    N_ALPHAS_SS = 13

    def _setup(self, dynamics, cost, N, max_reg, hessians, rollout, debug, line_search=False, q_ref=None, xi_ref=None):
        self.dynamics = dynamics
        self.cost = cost
        self.N = N
        self._use_hessians = hessians and dynamics.has_hessians
        if hessians and not dynamics.has_hessians:
            warnings.warn("hessians requested but are unavailable in dynamics")
        self._mu = 1.0
        self._mu_min = 1e-6
        self._mu_max = max_reg
        self._delta_0 = 2.0
        self._delta = self._delta_0
        self._action_size = dynamics.action_size
        self._state_size = dynamics.state_size
        self._error_state_size = dynamics.error_state_size
        self._rollout_mode = rollout
        self._line_search = line_search
        self._debug = debug
        self._q_ref = q_ref if q_ref is not None else cost.q_ref
        self._xi_ref = xi_ref if xi_ref is not None else cost.xi_ref
        self._k = np.zeros((N, self._action_size))
        self._K = np.zeros((N, self._action_size, self._state_size))
        self._kind = dynamics.KIND
        self._solvers = {}
        self.last_result = None

    state_size = property(lambda self: self._state_size)
    action_size = property(lambda self: self._action_size)
    error_state_size = property(lambda self: self._error_state_size)
    q_ref = property(lambda self: self._q_ref)
    xi_ref = property(lambda self: self._xi_ref)

    def get_q_ref(self, i):
        return self._q_ref[i]

    def get_xi_ref(self, i):
        return self._xi_ref[i]

    # ------------------------------------------------------------------------------------ native
    def _base_cost(self):
        return self.cost.cost if hasattr(self.cost, "constr") else self.cost

    def _solver(self, B, device=None, **params):
        if self._use_hessians:
            # the reference's `_linearization` calls dynamics.f_xx here (traopt_controller.py:2150-2156, 2890-2896), which
            # fails for every exact-dynamics class: same error, before anything is solved
            self.dynamics.f_xx(None, None, 0)
        key = (B, str(device))
        s = self._solvers.get(key)
        cost = self._base_cost()
        bounds = None
        if self.METHOD == "al_ms":
            bounds = (self.constr.input_lb, self.constr.input_ub)
            if hasattr(self.constr, "xi_lb"):
                params = dict(params, xi_lb=self.constr.xi_lb, xi_ub=self.constr.xi_ub)
        if s is None:
            s = _native.make_solver(self._kind, self.METHOD, self.N, B, self.dynamics, cost, self._q_ref, self._xi_ref,
                                    device=device, bounds=bounds, **params)
            self._solvers[key] = s
        else:
            if bounds is not None:
                params["lb"], params["ub"] = bounds
            s.set_params(Q=cost.Q, R=cost.R, P=cost.P, **_native.dynamics_params(self._kind, self.dynamics), **params)
        return s

    def _xi_ref_array(self):
        from . import manif_compat
        if self._kind in ("so3", "pendulum"):
            return np.asarray([manif_compat.so3_vel(w) for w in self._xi_ref])
        return np.asarray(self._xi_ref, dtype=float)

    def _params(self, n_iterations, tol_grad_norm, tol_d_norm=1e-6):
        return dict(max_iters=n_iterations, tol_grad_norm=tol_grad_norm, tol_d_norm=tol_d_norm,
                    max_reg=self._mu_max if self._mu_max else 0.0, rollout=self._rollout_mode,
                    line_search=self._line_search)

    def _alphas(self):
        n = self.N_ALPHAS_SS if self.METHOD == "ss" else (13 if self._kind in ("so3", "pendulum") else 20)
        return 1.1 ** (-np.arange(n) ** 2)

    def _x0_rows(self, x0_batch):
        if isinstance(x0_batch, np.ndarray) and x0_batch.ndim == 2:
            return np.ascontiguousarray(x0_batch, dtype=float)
        return np.stack([_native.state_row(self._kind, x) for x in x0_batch])

    def _store_gains(self, s):
        k, K = s.debug_gains()
        self._k, self._K = k[0].cpu().numpy(), K[0].cpu().numpy()

    def _ms_callback_loop(self, s, step, n_iterations, on_iteration, J_hist, xs_hist, us_hist, grad_hist, defect_hist):
        """The `for iteration in range(n_iterations)` loop of the multiple-shooting `fit` (traopt_controller.py:2493-2633)
        with the scripts' callback protocol; `step()` advances the device by one iteration.  Returns (out, hist, status).

        Iteration j's cost and new defect are evaluated by the device in pass j+1 (they are the stage costs / defects
        of the next linearisation), so its callback fires one pass later."""
        kind = self._kind
        alphas = self._alphas()
        pending = None
        out = h = None
        status = _lib.STATUS_RUNNING
        for j in range(n_iterations + 1):
            step()
            out = s.export()
            h = s.export_hist()
            status = int(out["status"][0]) & 15
            if j == 0:
                defect_hist.append(float(h["defect_hist"][0, 0]))              # appended by fit at iteration 0 (:2505-2506)
            if pending is not None:
                pj, pxs, pus, pmu, pa = pending
                accepted = pa >= 0
                on_iteration(pj, pxs, pus, float(h["J_hist"][0, pj]), accepted, False, float(h["defect_hist"][0, pj + 1]),
                             np.float64(h["grad_hist"][0, pj].item()), float(alphas[pa]) if self._line_search and accepted else 1,
                             pmu, J_hist, xs_hist, us_hist, grad_hist, defect_hist)
                pending = None
                if not accepted:
                    break
            if status != _lib.STATUS_RUNNING:
                break
            self._mu = float(s.export_reg()[0][0])
            pending = (j, _native.rows_states(kind, out["xs"][0].cpu().numpy()), out["us"][0].cpu().numpy(), self._mu,
                       int(h["alpha_hist"][0, j]))
        return out, h, status

    # -------------------------------------------------------------------------------- fit_batch
    def fit_batch(self, x0_batch, us_init=None, n_iterations=100, tol_grad_norm=None, tol_d_norm=1e-6,
                  return_trajectories=True, return_hist=False, device=None, shard=None, q_ref_batch=None,
                  xi_ref_batch=None, horizons=None, **extra):
        """Solve B problems that differ in their initial state (and optionally initial controls).

        x0_batch: list of reference-style states [q, xi], or an array of device rows (B, NS).
        us_init:  None (zeros), (N, m) shared, or (B, N, m).
        q_ref_batch, xi_ref_batch: one reference per problem — (B, N+1, 4, 4) poses (or (B, N+1, 3, 3) / device rows)
                  and (B, N+1, 6|3) twists; default: the controller's own reference for every problem.
        horizons: (B,) ints, 1 <= N_b <= N: problem b is solved over its first N_b stages (terminal cost at N_b).
        shard:    split the batch over the ranks of the initialised torch.distributed job
                  (default: yes when a job is initialised); summaries are all-gathered, trajectories
                  stay on the rank that solved them.
        Returns a BatchResult whose J/iters/status/grad/defect cover the WHOLE batch.
        """
        if tol_grad_norm is None:
            tol_grad_norm = self.DEFAULT_TOL_GRAD
        rows = self._x0_rows(x0_batch)
        B = rows.shape[0]
        rank, ws = _dist.world()
        if shard is None:
            shard = ws > 1
        params = self._params(n_iterations, tol_grad_norm, tol_d_norm)
        params.update(extra)
        if (q_ref_batch is None) != (xi_ref_batch is None):
            raise ValueError("q_ref_batch and xi_ref_batch go together")

        def ready(b, lo_):
            """solver for b problems starting at global index lo_, with the per-problem references of that range"""
            sv = self._solver(b, device, **params)
            if q_ref_batch is not None:
                q = np.asarray(q_ref_batch)[lo_:lo_ + b, :self.N + 1]
                so3 = self._kind in ("so3", "pendulum")
                q_rows = q if q.shape[-1] in (4, 7) and q.ndim == 3 else _native.layout.pose_rows(so3, q)
                sv.set_reference_batch(q_rows, np.asarray(xi_ref_batch, dtype=float)[lo_:lo_ + b, :self.N + 1])
            else:
                sv.set_reference(_native.ref_rows(self._kind, self._q_ref)[:self.N + 1], self._xi_ref_array()[:self.N + 1])
            sv.set_horizons(None if horizons is None else np.asarray(horizons)[lo_:lo_ + b])
            return sv
        if shard and ws > 1:
            lo, hi = _dist.shard_bounds(B, rank, ws)
            made = []          # the solver that solved the shard: asking _solver() again would reset its state (set_params)

            def make(b):
                made.append(ready(b, lo))
                return made[-1]
            out, summ, (lo, hi) = _dist.solve_sharded(make, rows, us_init, trajectories=return_trajectories)
            s = made[-1]
        else:
            lo, hi = 0, B
            s = ready(B, 0)
            out = s.solve(rows, us_init, trajectories=return_trajectories)
            summ = out
        res = BatchResult(self._kind, *(summ[k].cpu().numpy() for k in ("J", "iters", "status", "grad", "defect")),
                          xs_rows=out["xs"].cpu().numpy() if return_trajectories else None,
                          us=out["us"].cpu().numpy() if return_trajectories else None, shard=(lo, hi))
        if return_hist:
            h = s.export_hist()
            res.J_hist, res.grad_hist, res.defect_hist, res.alpha_hist = (h[k].cpu().numpy() for k in
                                                                          ("J_hist", "grad_hist", "defect_hist", "alpha_hist"))
        if self.METHOD == "al_ms":
            res.extra = {k: v.cpu().numpy() for k, v in s.export_al().items()}
        self.last_result = res
        return res


# ================================================================================================
# single shooting
# ================================================================================================
class _SingleShooting(_NativeController):
    METHOD = "ss"

    def fit(self, x0, us_init, n_iterations=100, tol_J=1e-6, tol_grad_norm=None, on_iteration=None):
        """Reference signature and return tuple: (xs, us, J_hist, xs_hist, us_hist, grad_hist)
        (traopt_controller.py:1880-2013; SO3 twin :575-695)."""
        if tol_grad_norm is None:
            tol_grad_norm = self.DEFAULT_TOL_GRAD
        kind = self._kind
        us0 = np.array(us_init, dtype=float)
        s = self._solver(1, None, **self._params(n_iterations, tol_grad_norm))
        s.begin(_native.state_row(kind, x0)[None, :], us0)
        J_hist, xs_hist, us_hist, grad_hist = [], [], [], []
        out = s.export()
        xs = _native.rows_states(kind, out["xs"][0].cpu().numpy())
        us = out["us"][0].cpu().numpy()
        xs_hist.append(list(xs))
        us_hist.append(us.copy())
        alphas = self._alphas()
        if on_iteration is None:
            s.iterate(n_iterations)
            out = s.export()
            h = s.export_hist()
            it = int(out["iters"][0])
            status = int(out["status"][0]) & 15
            n_grad = it + 1 if status == _lib.STATUS_CONVERGED else it
            grad_hist.extend(h["grad_hist"][0, :n_grad].cpu().numpy().tolist())
            if status == _lib.STATUS_NO_DESCENT:
                warnings.warn("Couldn't find descent direction, regularization and line search step exhausted")
        else:
            for it in range(n_iterations):
                s.iterate(1)
                out = s.export()
                h = s.export_hist()
                status = int(out["status"][0]) & 15
                grad = float(h["grad_hist"][0, it])
                grad_hist.append(grad)
                if status == _lib.STATUS_CONVERGED:       # break before the callback (:1939-1942)
                    break
                a_idx = int(h["alpha_hist"][0, it])
                accepted = a_idx >= 0
                self._mu = float(s.export_reg()[0][0])
                xs = _native.rows_states(kind, out["xs"][0].cpu().numpy())
                us = out["us"][0].cpu().numpy()
                on_iteration(it, xs, us, float(h["J_hist"][0, it]), accepted, False, grad,
                             float(alphas[a_idx] if accepted else alphas[-1]), self._mu, J_hist, xs_hist, us_hist)
                if not accepted:
                    warnings.warn("Couldn't find descent direction, regularization and line search step exhausted")
                    break
        xs = _native.rows_states(kind, out["xs"][0].cpu().numpy())
        us = out["us"][0].cpu().numpy()
        self._store_gains(s)
        self.last_result = {k: v.cpu().numpy() for k, v in {**out, **s.export_hist()}.items()}
        return xs, us, J_hist, xs_hist, us_hist, grad_hist


class iLQR_Tracking_SE3(_SingleShooting):
    """Single-shooting iLQR on SE(3) (traopt_controller.py:1831-2349)."""
    DEFAULT_TOL_GRAD = 1e-3

    def __init__(self, dynamics, cost, N, max_reg=1e10, hessians=False, rollout='linear', debug=None):
        self._setup(dynamics, cost, N, max_reg, hessians, rollout, debug)


class iLQR_Tracking_SO3(_SingleShooting):
    """Single-shooting iLQR on SO(3) (traopt_controller.py:526-1026)."""
    DEFAULT_TOL_GRAD = 1e-6

    def __init__(self, dynamics, cost, N, max_reg=1e10, hessians=False, rollout='nonlinear', debug=None):
        self._setup(dynamics, cost, N, max_reg, hessians, rollout, debug)


# ================================================================================================
# multiple shooting
# ================================================================================================
class _MultipleShooting(_NativeController):
    METHOD = "ms"
    DEFAULT_TOL_GRAD = 1e-6
    APPEND_FINAL_GRAD = False      # the SO3 twin appends the converged gradient norm itself (:1219)

    def fit(self, x0, us_init, n_iterations=100, tol_J=1e-6, tol_grad_norm=1e-6, tol_d_norm=1e-6, on_iteration=None):
        """Reference signature and return tuple:
        (xs, us, J_hist, xs_hist, us_hist, grad_hist, defect_hist) (traopt_controller.py:2443-2639; SO3 :1131-1325)."""
        kind = self._kind
        us0 = np.array(us_init, dtype=float)
        s = self._solver(1, None, **self._params(n_iterations, tol_grad_norm, tol_d_norm))
        s.begin(_native.state_row(kind, x0)[None, :], us0)
        J_hist, xs_hist, us_hist, grad_hist, defect_hist = [], [], [], [], []
        out = s.export()
        xs_hist.append(_native.rows_states(kind, out["xs"][0].cpu().numpy()))
        us_hist.append(out["us"][0].cpu().numpy())
        alphas = self._alphas()
        status = _lib.STATUS_RUNNING
        if on_iteration is None:
            s.iterate(n_iterations + 1)
            out = s.export()
            h = s.export_hist()
            status = int(out["status"][0]) & 15
            defect_hist.append(float(h["defect_hist"][0, 0]))              # appended by fit at iteration 0 (:2505-2506)
        else:
            out, h, status = self._ms_callback_loop(s, lambda: s.iterate(1), n_iterations, on_iteration, J_hist, xs_hist,
                                                    us_hist, grad_hist, defect_hist)
        if status == _lib.STATUS_CONVERGED and self.APPEND_FINAL_GRAD:
            grad_hist.append(float(out["grad"][0]))
        if status == _lib.STATUS_NO_DESCENT:
            warnings.warn("Couldn't find descent direction, regularization and line search step exhausted")
        xs = _native.rows_states(kind, out["xs"][0].cpu().numpy())
        us = out["us"][0].cpu().numpy()
        self._store_gains(s)
        self.last_result = {k: v.cpu().numpy() for k, v in {**out, **s.export_hist()}.items()}
        return xs, us, J_hist, xs_hist, us_hist, grad_hist, defect_hist


class iLQR_Tracking_SE3_MS(_MultipleShooting):
    """Multiple-shooting iLQR on SE(3) (traopt_controller.py:2352-3136)."""

    def __init__(self, dynamics, cost, N, q_ref, xi_ref, max_reg=1e10, hessians=False, line_search=False,
                 rollout='linear', debug=None):
        self._setup(dynamics, cost, N, max_reg, hessians, rollout, debug, line_search, q_ref, xi_ref)


class iLQR_Tracking_SO3_MS(_MultipleShooting):
    """Multiple-shooting iLQR on SO(3) (traopt_controller.py:1029-1826)."""
    APPEND_FINAL_GRAD = True

    def __init__(self, dynamics, cost, N, q_ref, xi_ref, max_reg=1e10, hessians=False, line_search=False,
                 rollout='linear', debug=None):
        self._setup(dynamics, cost, N, max_reg, hessians, rollout, debug, line_search, q_ref, xi_ref)


# ================================================================================================
# augmented Lagrangian
# ================================================================================================
class AL_iLQR_Tracking_SE3_MS(_NativeController):
    """Input-constrained multiple shooting by augmented Lagrangian (traopt_controller.py:3139-3293).

    The reference's `fit` cannot run as committed (it unpacks 6 values from an inner `fit` that
    returns 7, :3236); the semantics implemented are the library's with that unpack fixed.
    """
    METHOD = "al_ms"
    DEFAULT_TOL_GRAD = 1e-6

    def __init__(self, dynamics, cost, constraints, N, q_ref, xi_ref, mu_scale=10., max_reg=1e10, hessians=False,
                 line_search=False, rollout='nonlinear', debug=None):
        from .traopt_cost import ALConstrainedCost
        self.constr = constraints
        self._setup(dynamics, cost, N, max_reg, hessians, 'nonlinear', debug, line_search, q_ref, xi_ref)
        self._constr_size = constraints.constr_size
        self._mu0 = 1e-2
        self._mu_scale = mu_scale
        self._mu_max_al = 1e8
        self.al = ALConstrainedCost(cost, constraints, N)

    def _al_params(self, n_al_iters, n_ilqr_iters, tol_constr):
        p = self._params(n_ilqr_iters, 1e-6, 1e-6)       # the inner tolerances are hard-coded (:3237-3240)
        p.update(n_al_iters=n_al_iters, al_mu0=self._mu0, al_mu_scale=self._mu_scale, al_mu_max=self._mu_max_al,
                 tol_constr=tol_constr)
        return p

    def fit(self, x0, us_init, n_al_iters=100, n_ilqr_iters=200, tol_J=1e-6, tol_grad_norm=1e-6, tol_constr=1e-2,
            on_iteration_al=None, on_iteration_ilqr=None):
        """Returns (xs, us, J_hist, xs_hist, us_hist, grad_hist, lmbd_hist, mu_hist, violation_hist, nactive_hist)."""
        kind = self._kind
        s = self._solver(1, None, **self._al_params(n_al_iters, n_ilqr_iters, tol_constr))
        s.begin(_native.state_row(kind, x0)[None, :], np.array(us_init, dtype=float))
        lmbd_hist, mu_hist, violation_hist, nactive_hist = [], [], [], []
        J_hist, xs_hist, us_hist, grad_hist, defect_hist = [], [], [], [], []
        c = self._constr_size
        for it in range(n_al_iters):
            al = s.export_al()                              # multipliers this outer iteration is solved with
            lmbd = al["lmbd"][0].cpu().numpy()
            Imu = np.stack([np.diag(r) for r in al["imu"][0].cpu().numpy()])
            mu = float(al["mu"][0])
            if on_iteration_ilqr is not None:
                # the inner fit() of this outer iteration, one device iteration per callback (:3236-3240); its history
                # lists start afresh like those of every iLQR_Tracking_SE3_MS.fit call (:2478-2491)
                J_hist, xs_hist, us_hist, grad_hist, defect_hist = [], [], [], [], []
                s.iterate_inner(0)                          # opens the inner solve (cold start) without iterating
                o0 = s.export()
                xs_hist.append(_native.rows_states(kind, o0["xs"][0].cpu().numpy()))
                us_hist.append(o0["us"][0].cpu().numpy())
                self._ms_callback_loop(s, lambda: s.iterate_inner(1), n_ilqr_iters, on_iteration_ilqr, J_hist, xs_hist,
                                       us_hist, grad_hist, defect_hist)
            active = s.iterate(1)                           # (rest of) the inner solve + constraint evaluation + update
            out = s.export()
            us = out["us"][0].cpu().numpy()
            if hasattr(self.constr, "xi_lb"):
                xs_it = _native.rows_states(kind, out["xs"][0].cpu().numpy())
                constr_eval = np.vstack([self.constr.g(xs_it[i], us[i], i) for i in range(self.N)]
                                        + [self.constr.g(xs_it[self.N], None, self.N, terminal=True)])
                lmbd = np.concatenate((lmbd, al["lmbd_state"][0].cpu().numpy()), axis=1)
                Imu = np.stack([np.diag(np.concatenate((np.diag(Imu[i]), r))) for i, r in
                                enumerate(al["imu_state"][0].cpu().numpy())])
            else:
                constr_eval = np.vstack([np.concatenate((self.constr.input_lb - u, u - self.constr.input_ub)) for u in us]
                                        + [np.zeros(c)])
            converged = bool(np.max(constr_eval) < tol_constr)
            self.al.lmbd, self.al.Imu, self.al.mu = lmbd, Imu, mu
            if on_iteration_al:
                on_iteration_al(it, converged, lmbd, Imu, mu, constr_eval, lmbd_hist, mu_hist, violation_hist, nactive_hist)
            if converged or active == 0:
                break
        out = s.export()
        h = s.export_hist()
        n = int(out["iters"][0])
        xs = _native.rows_states(kind, out["xs"][0].cpu().numpy())
        us = out["us"][0].cpu().numpy()
        self.last_result = {k: v.cpu().numpy() for k, v in {**out, **h, **s.export_al()}.items()}
        if on_iteration_ilqr is None:      # no callback to fill the lists: the device's own histories of the last inner solve
            J_hist, grad_hist = h["J_hist"][0, :n].cpu().numpy().tolist(), h["grad_hist"][0, :n].cpu().numpy().tolist()
        return xs, us, J_hist, xs_hist, us_hist, grad_hist, lmbd_hist, mu_hist, violation_hist, nactive_hist

    def fit_batch(self, x0_batch, us_init=None, n_al_iters=100, n_ilqr_iters=200, tol_constr=1e-2, **kw):
        return super().fit_batch(x0_batch, us_init, n_iterations=n_ilqr_iters, tol_grad_norm=1e-6, tol_d_norm=1e-6,
                                 n_al_iters=n_al_iters, al_mu0=self._mu0, al_mu_scale=self._mu_scale,
                                 al_mu_max=self._mu_max_al, tol_constr=tol_constr, **kw)
