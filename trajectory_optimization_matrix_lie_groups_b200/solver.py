"""`BatchSolver`: the one Python object that owns a native solver handle.

Everything numeric happens in `libtrajopt_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/trajopt_b200.h); PyTorch is used here only for device memory and streams.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

KINDS = {"so3": _lib.SO3, "se3": _lib.SE3, "drone": _lib.DRONE, "rigid": _lib.RIGID, "pendulum": _lib.PEND}
METHODS = {"ss": _lib.SS, "ms": _lib.MS, "al_ms": _lib.AL_MS}
DIMS = {  # kind -> (NX, NP, NU, NS)
    "so3": (6, 3, 3, 7), "se3": (12, 6, 6, 13), "drone": (12, 6, 4, 13), "rigid": (12, 6, 6, 13), "pendulum": (6, 3, 3, 7),
}


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class BatchSolver:
    """B independent tracking problems of one family, solved together on one GPU.

    kind: 'so3' | 'se3' | 'drone' | 'rigid' | 'pendulum';  method: 'ss' | 'ms' | 'al_ms'.
    """

    def __init__(self, kind, method, N, B, device=None):
        if not torch.cuda.is_available():
            raise _lib.TrajoptError("no CUDA device: this solver has no CPU path")
        self.kind, self.method, self.N, self.B = kind, method, int(N), int(B)
        self.NX, self.NP, self.NU, self.NS = DIMS[kind]
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        torch.cuda.init()
        torch.zeros(1, device=dev)          # make sure torch's primary context exists on the device
        h = C.c_void_p()
        check(lib.trajopt_create(KINDS[kind], METHODS[method], self.N, self.B, dev.index, C.byref(h)))
        self._h = h
        self.max_iters = None
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            lib.trajopt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------ configuration
    def set_params(self, *, dt, Ib, mass, Q, R, P, gravity=9.8, max_iters=100, tol_grad_norm=1e-6,
                   tol_d_norm=1e-6, max_reg=1e10, rollout="nonlinear", line_search=False, n_alphas=0,
                   defect_kappa=0.0, lb=None, ub=None, n_al_iters=100, al_mu0=1e-2, al_mu_scale=10.0,
                   al_mu_max=1e8, tol_constr=1e-2, length=0.0, xi_lb=None, xi_ub=None):
        NX, NU = self.NX, self.NU
        p = _lib.Params()
        p.dt = float(dt)
        p.Ib[:] = np.asarray(Ib, dtype=np.float64).reshape(9).tolist()
        p.mass = float(mass)
        p.gravity = float(gravity)
        p.length = float(length)
        p.has_state_bounds = int(xi_lb is not None)
        if xi_lb is not None:
            nv = self.NX - self.NP
            xl = np.broadcast_to(np.asarray(xi_lb, dtype=np.float64), (nv,))
            xu = np.broadcast_to(np.asarray(xi_ub, dtype=np.float64), (nv,))
            for i in range(nv):
                p.xi_lb[i], p.xi_ub[i] = xl[i], xu[i]
        self._has_state_bounds = xi_lb is not None
        for name, M, n in (("Q", Q, NX), ("P", P, NX), ("R", R, NU)):
            M = np.asarray(M, dtype=np.float64)
            if M.shape != (n, n):
                raise ValueError(f"{name} must be {n}x{n}")
            buf = getattr(p, name)
            flat = M.reshape(-1)
            for i in range(flat.size):
                buf[i] = flat[i]
        p.has_constraints = int(lb is not None)
        if lb is not None:
            lb = np.broadcast_to(np.asarray(lb, dtype=np.float64), (NU,))
            ub = np.broadcast_to(np.asarray(ub, dtype=np.float64), (NU,))
            for i in range(NU):
                p.lb[i], p.ub[i] = lb[i], ub[i]
        if rollout not in ("linear", "nonlinear"):
            raise ValueError("rollout must be 'linear' or 'nonlinear'")
        p.rollout_linear = int(rollout == "linear")
        p.line_search = int(bool(line_search))
        p.n_alphas = int(n_alphas)
        p.max_iters = int(max_iters)
        p.tol_grad_norm = float(tol_grad_norm)
        p.tol_d_norm = float(tol_d_norm)
        p.max_reg = float(max_reg) if max_reg else 0.0
        p.defect_kappa = float(defect_kappa)
        p.n_al_iters = int(n_al_iters)
        p.al_mu0, p.al_mu_scale, p.al_mu_max, p.tol_constr = float(al_mu0), float(al_mu_scale), float(al_mu_max), float(tol_constr)
        check(lib.trajopt_set_params(self._h, C.byref(p)))
        self.max_iters = int(max_iters)
        self.n_al_iters = int(n_al_iters)

    def set_reference(self, q_ref_rows, xi_ref):
        """q_ref_rows: (N+1, 7|4) quat[+pos] rows; xi_ref: (N+1, 6|3).  Shared by the whole batch."""
        q = np.ascontiguousarray(q_ref_rows, dtype=np.float64)
        xi = np.ascontiguousarray(xi_ref, dtype=np.float64)
        if q.shape != (self.N + 1, self.NS - (self.NX - self.NP)) or xi.shape != (self.N + 1, self.NX - self.NP):
            raise ValueError(f"reference shapes {q.shape}, {xi.shape} do not match N={self.N}, kind={self.kind}")
        check(lib.trajopt_set_reference(self._h, q.ctypes.data_as(C.c_void_p), xi.ctypes.data_as(C.c_void_p)))

    def set_reference_long(self, q_ref_rows, xi_ref):
        """A shared reference longer than the horizon: (n, 7|4) and (n, 6|3), n >= N + 1; `set_reference_offset` slides the window."""
        q = np.ascontiguousarray(q_ref_rows, dtype=np.float64)
        xi = np.ascontiguousarray(xi_ref, dtype=np.float64)
        if q.ndim != 2 or q.shape[1] != self.NS - (self.NX - self.NP) or xi.shape != (q.shape[0], self.NX - self.NP):
            raise ValueError(f"reference shapes {q.shape}, {xi.shape} do not match kind={self.kind}")
        check(lib.trajopt_set_reference_long(self._h, q.ctypes.data_as(C.c_void_p), xi.ctypes.data_as(C.c_void_p), q.shape[0]))

    def set_reference_offset(self, first_row):
        check(lib.trajopt_set_reference_offset(self._h, int(first_row)))

    def set_reference_batch(self, q_ref_rows, xi_ref):
        """One reference per problem: q_ref_rows (B, N+1, 7|4) quat[+pos] rows, xi_ref (B, N+1, 6|3)."""
        npose, nv = self.NS - (self.NX - self.NP), self.NX - self.NP
        q = self._dev(q_ref_rows, (self.B, self.N + 1, npose))
        xi = self._dev(xi_ref, (self.B, self.N + 1, nv))
        check(lib.trajopt_set_reference_batch(self._h, _ptr(q), _ptr(xi), _stream(self.device)))
        torch.cuda.current_stream(self.device).synchronize()      # q, xi may be temporaries

    def set_horizons(self, horizons=None):
        """One horizon per problem, (B,) ints in [1, N]; None restores N for all."""
        if horizons is None:
            check(lib.trajopt_set_horizons(self._h, C.c_void_p(0), _stream(self.device)))
            return
        t = self._dev(np.asarray(horizons, dtype=np.int32), (self.B,), dtype=torch.int32)
        check(lib.trajopt_set_horizons(self._h, _ptr(t), _stream(self.device)))
        torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------------------------ solving
    def _dev(self, a, shape=None, dtype=torch.float64):
        t = torch.as_tensor(a, dtype=dtype)
        if t.device != self.device:
            t = t.to(self.device)
        t = t.contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _us_arg(self, us_init):
        if us_init is None:
            return None, 0
        us = torch.as_tensor(us_init, dtype=torch.float64)
        if us.dim() == 2:
            return self._dev(us, (self.N, self.NU)), 1
        return self._dev(us, (self.B, self.N, self.NU)), 2

    def begin(self, x0, us_init=None):
        x0 = self._dev(x0, (self.B, self.NS))
        us, mode = self._us_arg(us_init)
        self._keep = [x0, us]               # the library reads us_init lazily; keep it alive
        check(lib.trajopt_begin(self._h, _ptr(x0), _ptr(us), mode, _stream(self.device)))

    def iterate(self, n_iters):
        act = C.c_int(0)
        check(lib.trajopt_iterate(self._h, int(n_iters), C.byref(act), _stream(self.device)))
        return act.value

    def iterate_inner(self, n_iters=1):
        """AL handles: inner iterations of the current outer iteration; returns the problems whose inner solve still runs."""
        act = C.c_int(0)
        check(lib.trajopt_iterate_inner(self._h, int(n_iters), C.byref(act), _stream(self.device)))
        return act.value

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def export(self, trajectories=True):
        B, N = self.B, self.N
        out = {
            "J": self._new(B), "iters": self._new(B, dtype=torch.int32), "status": self._new(B, dtype=torch.int32),
            "grad": self._new(B), "defect": self._new(B),
            "xs": self._new(B, N + 1, self.NS) if trajectories else None,
            "us": self._new(B, N, self.NU) if trajectories else None,
        }
        check(lib.trajopt_export(self._h, _ptr(out["xs"]), _ptr(out["us"]), _ptr(out["J"]), _ptr(out["iters"]),
                                 _ptr(out["status"]), _ptr(out["grad"]), _ptr(out["defect"]), _stream(self.device)))
        return out

    def export_hist(self):
        B, M = self.B, self.max_iters
        out = {"J_hist": self._new(B, M), "grad_hist": self._new(B, M + 1), "defect_hist": self._new(B, M + 1),
               "alpha_hist": self._new(B, M, dtype=torch.int32)}
        check(lib.trajopt_export_hist(self._h, _ptr(out["J_hist"]), _ptr(out["grad_hist"]), _ptr(out["defect_hist"]),
                                      _ptr(out["alpha_hist"]), _stream(self.device)))
        return out

    def export_reg(self):
        mu, delta = self._new(self.B), self._new(self.B)
        check(lib.trajopt_export_reg(self._h, _ptr(mu), _ptr(delta), _stream(self.device)))
        return mu, delta

    def export_al(self):
        B, N, c = self.B, self.N, 2 * self.NU
        out = {"lmbd": self._new(B, N + 1, c), "imu": self._new(B, N + 1, c), "mu": self._new(B),
               "outer_iters": self._new(B, dtype=torch.int32), "violation": self._new(B)}
        check(lib.trajopt_export_al(self._h, _ptr(out["lmbd"]), _ptr(out["imu"]), _ptr(out["mu"]),
                                    _ptr(out["outer_iters"]), _ptr(out["violation"]), _stream(self.device)))
        if getattr(self, "_has_state_bounds", False):
            cs = 2 * (self.NX - self.NP)
            out["lmbd_state"], out["imu_state"] = self._new(B, N + 1, cs), self._new(B, N + 1, cs)
            check(lib.trajopt_export_al_state(self._h, _ptr(out["lmbd_state"]), _ptr(out["imu_state"]), _stream(self.device)))
        return out

    def solve(self, x0, us_init=None, trajectories=True):
        """One whole fit() for every problem; inputs/outputs are device tensors."""
        self.begin(x0, us_init)
        units = self.n_al_iters if self.method == "al_ms" else self.max_iters + 1
        self.iterate(units)
        return self.export(trajectories)

    def solve_stream(self, x0, us_init=None, trajectories=True, out=None):
        """Continuous batching: any number M of problems through this solver's B slots (`trajopt_solve_stream`).

        x0: (M, NS) device tensor / array.  us_init: None or ONE (N, NU) sequence shared by all problems.
        Returns the same dict as `solve`, with M rows; problem p's row is what `solve` returns for x0[p].
        `out`: optional dict of preallocated device tensors to fill (bench: no allocation in the timed region).
        """
        x0 = torch.as_tensor(x0, dtype=torch.float64)
        if x0.device != self.device:
            x0 = x0.to(self.device)
        x0 = x0.contiguous()
        if x0.ndim != 2 or x0.shape[1] != self.NS:
            raise ValueError(f"x0 must be (M, {self.NS})")
        M, N = x0.shape[0], self.N
        us_t = None if us_init is None else self._dev(us_init, (N, self.NU))
        if out is None:
            out = {"J": self._new(M), "iters": self._new(M, dtype=torch.int32), "status": self._new(M, dtype=torch.int32),
                   "grad": self._new(M), "defect": self._new(M),
                   "xs": self._new(M, N + 1, self.NS) if trajectories else None,
                   "us": self._new(M, N, self.NU) if trajectories else None}
        if M == 0:
            return out
        check(lib.trajopt_solve_stream(self._h, _ptr(x0), M, _ptr(us_t), _ptr(out["xs"]), _ptr(out["us"]), _ptr(out["J"]),
                                       _ptr(out["iters"]), _ptr(out["status"]), _ptr(out["grad"]), _ptr(out["defect"]),
                                       _stream(self.device)))
        return out

    def solve_stream_host(self, x0, us_init=None, trajectories=True, out=None):
        """`solve_stream` through HOST buffers (NumPy, ideally pinned): (M, NS) in, M rows out, copies overlapped."""
        N = self.N
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        if x0.ndim != 2 or x0.shape[1] != self.NS:
            raise ValueError(f"x0 must be (M, {self.NS})")
        M = x0.shape[0]
        us_p = C.c_void_p(0)
        if us_init is not None:
            us_init = np.ascontiguousarray(us_init, dtype=np.float64)
            if us_init.shape != (N, self.NU):
                raise ValueError(f"us_init must be {(N, self.NU)}")
            us_p = us_init.ctypes.data_as(C.c_void_p)
        if out is None:
            out = {"J": np.empty(M), "iters": np.empty(M, dtype=np.int32), "status": np.empty(M, dtype=np.int32),
                   "grad": np.empty(M), "defect": np.empty(M),
                   "xs": np.empty((M, N + 1, self.NS)) if trajectories else None,
                   "us": np.empty((M, N, self.NU)) if trajectories else None}
        if M == 0:
            return out

        def hp(a):
            return C.c_void_p(0) if a is None else a.ctypes.data_as(C.c_void_p)
        check(lib.trajopt_solve_stream_host(self._h, x0.ctypes.data_as(C.c_void_p), M, us_p, hp(out["xs"]), hp(out["us"]),
                                            hp(out["J"]), hp(out["iters"]), hp(out["status"]), hp(out["grad"]),
                                            hp(out["defect"]), _stream(self.device)))
        return out

    def solve_host_begin(self, x0, us_init=None, trajectories=True, out=None):
        """First half of `solve_host`: returns (ticket, out) once the solve is done and the copies into `out` are queued;
        `solve_host_wait(ticket)` completes them.  The next solve on this handle may begin in between."""
        B, N = self.B, self.N
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        if x0.shape != (B, self.NS):
            raise ValueError(f"x0 must be {(B, self.NS)}")
        mode, us_p = 0, C.c_void_p(0)
        if us_init is not None:
            us_init = np.ascontiguousarray(us_init, dtype=np.float64)
            mode = 1 if us_init.ndim == 2 else 2
            us_p = us_init.ctypes.data_as(C.c_void_p)
        if out is None:
            out = {"J": np.empty(B), "iters": np.empty(B, dtype=np.int32), "status": np.empty(B, dtype=np.int32),
                   "grad": np.empty(B), "defect": np.empty(B),
                   "xs": np.empty((B, N + 1, self.NS)) if trajectories else None,
                   "us": np.empty((B, N, self.NU)) if trajectories else None}

        def hp(a):
            return C.c_void_p(0) if a is None else a.ctypes.data_as(C.c_void_p)
        ticket = C.c_int(-1)
        self._host_keep = (x0, us_init, out)      # the library reads / writes these until the ticket's wait returns
        check(lib.trajopt_solve_host_begin(self._h, x0.ctypes.data_as(C.c_void_p), us_p, mode, hp(out["xs"]), hp(out["us"]),
                                           hp(out["J"]), hp(out["iters"]), hp(out["status"]), hp(out["grad"]),
                                           hp(out["defect"]), _stream(self.device), C.byref(ticket)))
        return ticket.value, out

    def solve_host_wait(self, ticket):
        check(lib.trajopt_solve_host_wait(self._h, int(ticket)))

    def solve_host(self, x0, us_init=None, trajectories=True, out=None):
        """Same through HOST buffers (NumPy, ideally pinned): H2D, solve, D2H inside one call."""
        ticket, out = self.solve_host_begin(x0, us_init, trajectories, out)
        self.solve_host_wait(ticket)
        return out

    # ---------------------------------------------------------------------------- parity exports
    def debug_linearize(self):
        B, N, NX, NU = self.B, self.N, self.NX, self.NU
        out = {"F_x": self._new(B, N, NX, NX), "F_u": self._new(B, N, NX, NU), "d": self._new(B, N, NX),
               "L": self._new(B, N + 1), "L_x": self._new(B, N + 1, NX), "L_xx": self._new(B, N + 1, NX, NX),
               "L_u": self._new(B, N, NU)}
        out["d"].zero_()
        check(lib.trajopt_debug_linearize(self._h, _ptr(out["F_x"]), _ptr(out["F_u"]), _ptr(out["d"]), _ptr(out["L"]),
                                          _ptr(out["L_x"]), _ptr(out["L_xx"]), _ptr(out["L_u"]), _stream(self.device)))
        return out

    def debug_gains(self):
        k = self._new(self.B, self.N, self.NU)
        K = self._new(self.B, self.N, self.NU, self.NX)
        check(lib.trajopt_debug_gains(self._h, _ptr(k), _ptr(K), _stream(self.device)))
        return k, K

    def debug_linesearch(self):
        """Line-search table of the last iteration, (rows, B); see trajopt_debug_linesearch in the header."""
        rows = lib.trajopt_debug_linesearch_rows(self._h)
        if rows < 0:
            check(rows)
        t = self._new(rows, self.B)
        check(lib.trajopt_debug_linesearch(self._h, _ptr(t), _stream(self.device)))
        return t

    def stage_eval(self, i, x_rows, u_rows=None, terminal=False, want=("f", "F_x", "F_u", "l", "l_x", "l_xx", "l_u", "err")):
        """The reference's per-stage callbacks on rows of states/controls against reference row i."""
        def rows(a, width):      # device tensors pass through; anything else comes up from the host
            t = a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a, dtype=np.float64))
            return t.to(self.device, torch.float64).reshape(-1, width).contiguous()
        x = rows(x_rows, self.NS)
        n = x.shape[0]
        u = None
        if not terminal:
            u = rows(u_rows, self.NU)
            if u.shape[0] != n:
                raise ValueError("x and u row counts differ")
        NX, NU, NS = self.NX, self.NU, self.NS
        shapes = {"f": (n, NS), "F_x": (n, NX, NX), "F_u": (n, NX, NU), "l": (n,), "l_x": (n, NX), "l_xx": (n, NX, NX),
                  "l_u": (n, NU), "err": (n, NX)}
        out = {k: (self._new(*shapes[k]) if k in want else None) for k in shapes}
        check(lib.trajopt_debug_stage(self._h, int(i), int(bool(terminal)), n, _ptr(x), _ptr(u), _ptr(out["f"]),
                                      _ptr(out["F_x"]), _ptr(out["F_u"]), _ptr(out["l"]), _ptr(out["l_x"]),
                                      _ptr(out["l_xx"]), _ptr(out["l_u"]), _ptr(out["err"]), _stream(self.device)))
        return {k: v for k, v in out.items() if v is not None}

    def set_compaction(self, min_batch=1024, ratio=4):
        """Pack the running problems into the leading slots once running * ratio <= slots in use (min_batch < 0: never)."""
        check(lib.trajopt_set_compaction(self._h, int(min_batch), int(ratio)))

    def set_line_search_batch(self, max_batch=256):
        """One launch rolls out every line-search step size while the padded batch is <= max_batch (0: never)."""
        check(lib.trajopt_set_line_search_batch(self._h, int(max_batch)))

    def set_sweep(self, variant=0, lanes=1):
        """Backward-sweep mapping: 0 automatic, 2 / 4 / 6 = always two- / four- / six-warp CTAs; lanes = solvers sharing this GPU."""
        check(lib.trajopt_set_sweep(self._h, int(variant), int(lanes)))

    def set_profiling(self, on):
        check(lib.trajopt_set_profiling(self._h, int(bool(on))))

    def phase_times(self, reset=True):
        ms = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        check(lib.trajopt_phase_times(self._h, ms, cnt, int(reset)))
        names = ("linearize", "backward", "forward", "other")
        return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}


def lie_op(name, x, device="cuda"):
    """Evaluate one Lie-group primitive of the device library on rows of `x` (parity tests)."""
    code, win, wout = _lib.LIE_OPS[name]
    t = torch.as_tensor(np.asarray(x, dtype=np.float64)).reshape(-1, win).contiguous().to(device)
    out = torch.empty(t.shape[0], wout, dtype=torch.float64, device=t.device)
    check(lib.trajopt_debug_lie(code, t.shape[0], _ptr(t), _ptr(out), _stream(t.device)))
    return out


def launch_count(reset=False):
    return int(lib.trajopt_launch_count(int(reset)))


def fp64_peak_tflops(ms_target=50.0, device=None):
    """Measured FP64 FMA throughput of the current device (roofline denominator), TFLOP/s."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = C.c_double(0.0)
    with torch.cuda.device(dev):
        check(lib.trajopt_debug_fp64_peak(float(ms_target), C.byref(out), _stream(dev)))
    return out.value
