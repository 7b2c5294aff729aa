#!/usr/bin/env python
"""Headline benchmark: batched SE3 multiple-shooting DDP tracking solves per second (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 3] [--batch B]

A "step" is one pass of the hot path over one synthetic batch: a complete `fit` (to the script's
own stopping rule) of every problem of the batch.  Default workload = BASELINE.json configs[2],
`main_SE3ddp_tracking_exact_ms.py` (SE3, multiple shooting, N=955, dt=0.004, tol_grad 1e-12), batch
16384 perturbed initial poses PER GPU (weak scaling: rank r solves problems [r*B, (r+1)*B) of the
global seeded batch; no collective on the data path, one NCCL all-gather of the per-problem
summaries (J, grad, defect, iters, status) at the end of every step).

`value`   device-timed (CUDA events, max over ranks) with the initial states resident in HBM; the timed solves return
          the per-problem summaries only (trajectories stay on the device; `e2e` exports and copies them).
`strong`  BASELINE.json words the metric as "batch 16k (1-8 GPU)": ONE 16384-problem sweep split over the N ranks
          (the reference's batch is one sweep split over its workers, visualization/perturb_all_compute.py:240-250).
          The `strong` block reports exactly that — 16384 / N problems per GPU — device-timed, end to end, and as the
          latency of one sweep alone; `value` stays the weak-scaling figure (16384 problems per GPU).
`e2e`     the same solves through the host-buffer C-ABI call (`trajopt_solve_host`): pinned host x0
          copied in, trajectories and summaries copied out, all inside the timed region.
`roofline` the backward Riccati sweep (dominant kernel): algorithmic FP64 FLOPs / measured duration
          against the FP64 FMA peak measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry).
`cpu_baseline` / `--impl reference`: the reference's algorithm restated in NumPy (oracle/, validated
          against the reference's shipped results) on the host cores, one process per core like
          the reference's joblib sweep, on a bounded sample (the reference itself needs jax, manifpy
          and casadi, which this image does not have).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SE3 tracking DDP solves/sec"
# the batch BASELINE.json quotes each config on, as ONE sweep (strong scaling splits it over the ranks)
STRONG_GLOBAL_BATCH = {1: 1, 2: 1024, 3: 16384, 4: 16384, 5: 1 << 20}
NOMINAL_ITERS = {1: 40, 2: 16, 3: 20, 4: 60, 5: 26}   # iterations of the unperturbed problem (goldens / SURVEY.md)


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    config, b, n_iters, batch = args
    import numpy as np
    from oracle import lie, models, solvers
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    wl = workloads.CONFIGS[config](B=batch)
    row = wl.x0_rows[b]
    if wl.kind == "so3":
        dyn = models.SO3Dynamics(wl.J, wl.dt)
        cost = models.SO3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
        group = solvers.SO3Group
        q_ref = [lie.rot_to_quat(R) for R in wl.q_ref]
        x0 = [row[:4].copy(), row[4:].copy()]
    else:
        dyn = (models.DroneDynamics if wl.kind == "drone" else models.SE3Dynamics)(wl.J, wl.dt)
        cost = models.SE3TrackingQuadraticGaussNewtonCost(wl.Q, wl.R, wl.P, wl.q_ref, wl.xi_ref)
        group = solvers.SE3Group
        q_ref = [np.asarray(T, dtype=float) for T in wl.q_ref]
        x0 = [lie.se3_to_matrix(row[:4], row[4:7]), row[7:].copy()]
    us0 = np.zeros((wl.N, dyn.action_size))
    if n_iters <= 0:                       # whole solve: the script's own iteration cap
        n_iters = int(wl.solver["max_iters"])
    t0 = time.perf_counter()
    if wl.method == "ss":
        r = solvers.ilqr_ss(dyn, cost, group, wl.N, x0, us0, n_iterations=n_iters, tol_grad_norm=wl.solver["tol_grad_norm"])
    else:
        c = cost
        if wl.method == "al_ms":
            constr = models.InputConstraint(np.full(6, wl.bounds[0]), np.full(6, wl.bounds[1]))
            c = models.ALConstrainedCost(cost, constr, wl.N)
            c.Imu = np.tile(1e-2 * np.eye(12), (wl.N + 1, 1, 1))
        r = solvers.ilqr_ms(dyn, c, group, wl.N, q_ref, wl.xi_ref, x0, us0, n_iterations=n_iters,
                            tol_grad_norm=wl.solver["tol_grad_norm"], n_alphas=13 if wl.kind == "so3" else 20)
    return time.perf_counter() - t0, r.iterations


def cpu_sample(config, n_iters, iters_per_solve, cores=None):
    """One bounded CPU sample: one problem per core, `n_iters` DDP iterations each — or, with n_iters <= 0, WHOLE solves
    to the script's own stopping rule (no extrapolation: solves per second = problems / pool time).

    Returns (solves_per_second, cores, description)."""
    import multiprocessing as mp
    # one single-threaded process per core (the reference's joblib sweep); BLAS pools on top of that only oversubscribe
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    cores = cores or int(os.environ.get("TRAJOPT_BENCH_CORES", "0")) or len(os.sched_getaffinity(0))
    batch = max(cores, 12)
    whole = n_iters <= 0
    jobs = [(config, b, n_iters, batch) for b in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    done_iters = sum(r[1] for r in res)
    busy = sum(r[0] for r in res)
    if whole:
        slowest = max(r[0] for r in res)
        desc = (f"oracle (NumPy restatement of traoptlibrary) on {cores} processes, 1 problem each (problems 0..{cores - 1} of the "
                f"workload), WHOLE solves to the script's stopping rule ({done_iters / cores:.1f} iterations on average, "
                f"{busy / max(done_iters, 1):.2f} s per iteration per core); solves/s = {cores} problems / {slowest:.1f} s (the slowest "
                f"worker: the pool is done when it is); pool wall {wall:.1f} s")
        return cores / slowest, cores, desc
    # throughput of the pool in DDP iterations/s, then whole solves = iterations / iters_per_solve
    it_per_s = done_iters / max(r[0] for r in res)
    value = it_per_s / iters_per_solve
    desc = (f"oracle (NumPy restatement of traoptlibrary) on {cores} processes, 1 problem each, first {n_iters} "
            f"DDP iteration(s) of the workload's solve ({busy / max(done_iters, 1):.2f} s per iteration per core), "
            f"extrapolated to whole solves at {iters_per_solve:.1f} iterations per solve; pool wall {wall:.1f} s")
    return value, cores, desc


def cpu_sample_subprocess(config, n_iters, iters_per_solve):
    """Run `cpu_sample` in a fresh interpreter (no CUDA context to fork)."""
    proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-sample", str(config), str(n_iters),
                           repr(float(iters_per_solve))], capture_output=True, text=True, timeout=600)
    if proc.returncode != 0:
        raise RuntimeError(proc.stderr[-400:])
    d = json.loads(proc.stdout.strip().splitlines()[-1])
    return d["value"], d["cores"], d["desc"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    its = float(NOMINAL_ITERS[args.config])
    # Each step is a bounded sample of the workload: WHOLE solves of one problem per core when the whole run then fits
    # ~4 minutes (judged from a one-iteration probe), else the first DDP iterations of each, extrapolated.
    t_probe = time.perf_counter()
    cpu_sample(args.config, 1, its)
    per_iter = time.perf_counter() - t_probe                      # ~ one iteration per core + pool start-up
    n_it = int(os.environ.get("TRAJOPT_BENCH_CPU_ITERS", "-1"))
    if n_it < 0:
        budget = 240.0 / max(args.steps + max(args.warmup - 1, 0), 1)
        n_it = 0 if per_iter * (its + 4.0) <= budget else max(1, int(budget / per_iter))
    for _ in range(max(args.warmup - 1, 0)):
        cpu_sample(args.config, 1, its)
    vals, desc, cores = [], "", 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, cores, desc = cpu_sample(args.config, n_it, its)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = statistics.mean(vals)
    from trajectory_optimization_matrix_lie_groups_b200 import workloads
    cfg = workload_config(args, workloads.CONFIGS[args.config](B=min(args.batch, 64)))   # same keys / values as the GPU arm's line
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def workload_config(args, wl):
    return {
        "workload": {1: "main_SE3ddp_tracking_exact.py", 2: "benchmark_SO3_tracking.py",
                     3: "main_SE3ddp_tracking_exact_ms.py", 4: "main_SE3ddp_tracking_exact_al_ms.py",
                     5: "benchmark_drone_racing_tracking.py"}[args.config],
        "baseline_config_index": args.config - 1,
        "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
        "horizon": None if wl is None else wl.N, "method": None if wl is None else wl.method,
        "perturbation": "one of 12 x0 parameters per problem, +-10% of the reference sweep half-width, seed 24234156",
        "l2": "working set (trajectories, linearisation, gains: tens of GB) far exceeds the 126 MB L2; no flush needed",
        "parallelism": f"dp{args.gpus} (independent problems, contiguous shards)",
        "schedule": args.schedule,
        "schedule_note": ("continuous batching (trajopt_solve_stream): the problems of the timed steps are one queue through "
                          "batch_per_gpu solver slots; a slot whose problem has finished takes the next problem before the following "
                          "DDP iteration; every problem's result is bit-identical to a plain batched solve (asserted here)")
        if args.schedule == "stream" else
        ("whole batches, `inflight` of them per GPU (own stream + host thread each); the tail of one batch's solve, where most "
         "problems have converged and most CTAs exit at once, overlaps the head of the next; steps complete in order"),
        "inflight": 1 if args.schedule == "stream" else max(1, args.inflight),
    }


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_lo, t_hi):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not (t_lo <= t <= t_hi):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def sweep_kernel_name(kind, B):
    """the Riccati sweep the profiled solve (one batch alone) launches at this batch size: run_backward, host_impl.cuh"""
    if kind in ("so3", "pendulum"):
        return "k_backward (Riccati sweep, one warp per 32 problems)"
    slots = (B + 31) // 32 * 32
    if slots <= 148 * 32:
        return "k_backward6 (Riccati sweep, 6-warp CTA per 32 problems: serial part on its own warps)"
    if slots <= 148 * 2 * 32:
        return "k_backward4 (Riccati sweep, 4-warp CTA per 32 problems)"
    return "k_backward3 (Riccati sweep, 2-warp CTA per 32 problems; the tail launches after compaction run k_backward4 / k_backward6)"


def load_kernel_counters(kind, B=1 << 30):
    """ncu-derived static counters of the dominant kernel (profiles/kernel_counters.json, written from a committed
    `ncu --set full` capture by scripts/ncu_summary.py): executed FP64 thread-instructions and DRAM bytes per launch and
    the stage-iterations that launch processed."""
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    slots = (B + 31) // 32 * 32
    se3 = "k_backward6_se3_2048" if slots <= 148 * 32 else ("k_backward4_se3_2048" if slots <= 148 * 64 else "k_backward3_se3")
    return table.get({"se3": se3, "rigid": se3, "drone": "k_backward3_drone"}.get(kind, "k_backward_so3"))


def measure_host_d2h(torch, dist, dev, world, gib=1.0, reps=3):
    """Concurrent device->host bandwidth of the box: every rank copies `gib` GiB into pinned memory at the same time
    (plain cudaMemcpyAsync, one large buffer).  Returns (this rank's GB/s, sum over ranks)."""
    n = int(gib * (1 << 30)) // 8
    src = torch.empty(n, dtype=torch.float64, device=dev)
    dst = torch.empty(n, dtype=torch.float64).pin_memory()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    gbs = reps * n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    tot = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    del src, dst
    return gbs, float(tot.item())


class Leg:
    """`depth` solver lanes for batches of B problems (rows [offset, offset + B) of the workload): the device-timed and the
    host-buffer (e2e) loops over K steps, with every buffer allocated up front."""

    def __init__(self, torch, dist, wl, B, offset, depth, dev, world, want_traj, compact):
        from trajectory_optimization_matrix_lie_groups_b200 import PipelinedSolver
        self.torch, self.dist, self.wl, self.B, self.dev, self.world, self.want_traj = torch, dist, wl, B, dev, world, want_traj
        self.depth = depth
        self.x0_rows = wl.x0_rows[offset:offset + B]
        self.pipe = PipelinedSolver(lambda: wl.make_solver(B=B, device=dev, offset=offset)[0], depth=depth, device=dev)
        self.solver = self.pipe.solvers[0]
        if compact:
            mb, ratio = (int(v) for v in compact.split(","))
            for sv in self.pipe.solvers:
                sv.set_compaction(mb, ratio)
        self.x0_dev = torch.as_tensor(self.x0_rows, device=dev)
        self.x0_pin = torch.as_tensor(self.x0_rows).pin_memory()
        self.outs = [self._pinned_out() for _ in range(depth)]      # one set of pinned result buffers per lane
        self.gathered = [torch.empty(B, 5, dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def _pinned_out(self):
        torch, B, wl, sv = self.torch, self.B, self.wl, self.solver
        o = {"J": torch.empty(B, dtype=torch.float64).pin_memory(), "grad": torch.empty(B, dtype=torch.float64).pin_memory(),
             "defect": torch.empty(B, dtype=torch.float64).pin_memory(),
             "iters": torch.empty(B, dtype=torch.int32).pin_memory(), "status": torch.empty(B, dtype=torch.int32).pin_memory(),
             "xs": torch.empty(B, wl.N + 1, sv.NS, dtype=torch.float64).pin_memory() if self.want_traj else None,
             "us": torch.empty(B, wl.N, sv.NU, dtype=torch.float64).pin_memory() if self.want_traj else None}
        return o, {k: (None if v is None else v.numpy()) for k, v in o.items()}

    def gather(self, out):
        if self.world > 1:   # the one collective of the path: per-problem summaries to every rank
            torch = self.torch
            summ = torch.stack((out["J"], out["grad"], out["defect"], out["iters"].double(), out["status"].double()), dim=1)
            self.dist.all_gather(self.gathered, summ)

    def run_steps(self, n, host):
        """n steps, `depth` batches in flight; steps complete (and are gathered) in submission order."""
        futs = []
        for k in range(n):
            if host:
                futs.append(self.pipe.submit(self.x0_pin.numpy(), trajectories=self.want_traj, host=True, out=self.outs[k % self.depth][1]))
            else:
                futs.append(self.pipe.submit(self.x0_dev, trajectories=False))
        last = None
        for f in futs:
            last = f.result()
            if not host:
                self.gather(last)
        return last

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed_device(self, steps):
        """K steps between barriers, CUDA events, max over ranks -> (ms_total, last result)"""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        out = self.run_steps(steps, False)      # every lane synchronises its stream before its result is handed back
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()), out

    def timed_host(self, steps, e2e_depth=0):
        """the same through trajopt_solve_host with pinned buffers: wall clock between barriers, max over ranks -> seconds"""
        torch = self.torch
        if e2e_depth > 0:
            self.pipe.active_depth = min(e2e_depth, self.depth)
        self.run_steps(min(self.depth, 2), True)
        self.barrier()
        t0 = time.perf_counter()
        self.run_steps(steps, True)
        self.barrier()
        sec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(sec, op=self.dist.ReduceOp.MAX)
        self.pipe.active_depth = self.depth
        out_np = self.outs[(steps - 1) % self.depth][1]
        return float(sec.item()), out_np

    def timed_single(self):
        """one batch alone on the GPU -> (ms max over ranks, result)"""
        torch = self.torch
        self.barrier()
        self.solver.set_sweep(0, 1)             # nothing else shares the GPU now: let the sweep pick its CTA shape for that
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        out1 = self.solver.solve(self.x0_dev, trajectories=False)
        s1.record()
        torch.cuda.synchronize(self.dev)
        self.solver.set_sweep(0, self.depth)
        ms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()), out1

    def close(self):
        self.pipe.close()
        self.outs = []
        self.torch.cuda.empty_cache()


def auto_inflight(batch):
    return 3 if batch > 2048 else 8


def decision_window(J):
    """Leading iterations of a cost history whose accept/stop decisions are above rounding noise (relative change >= 1e-12):
    the span over which the parity tests assert identical decisions (tests/test_gpu_golden.py, test_gpu_configs.py)."""
    for i in range(1, len(J)):
        if abs(J[i] - J[i - 1]) < 1e-12 * abs(J[i]):
            return i
    return len(J)


def run_stream_arm(args):
    """--schedule stream: continuous batching (trajopt_solve_stream), kept for A/B runs; prints the base contract's keys."""
    import numpy as np
    import torch
    from trajectory_optimization_matrix_lie_groups_b200 import workloads, launch_count
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--schedule stream is a single-GPU A/B mode")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = args.batch
    wl = workloads.CONFIGS[args.config](B=B)
    solver, x0_rows = wl.make_solver(B=B, device=dev)
    n_max = max(args.steps, args.warmup, 1)
    x0_all_dev = torch.as_tensor(x0_rows, device=dev).repeat(n_max, 1)
    x0_all_pin = torch.as_tensor(np.tile(x0_rows, (n_max, 1))).pin_memory()
    M, ns, want_traj = n_max * B, args.steps, not args.no_traj
    dev_out = {"J": torch.empty(M, dtype=torch.float64, device=dev), "grad": torch.empty(M, dtype=torch.float64, device=dev),
               "defect": torch.empty(M, dtype=torch.float64, device=dev), "iters": torch.empty(M, dtype=torch.int32, device=dev),
               "status": torch.empty(M, dtype=torch.int32, device=dev), "xs": None, "us": None}
    host_t = {"J": torch.empty(ns * B, dtype=torch.float64).pin_memory(), "grad": torch.empty(ns * B, dtype=torch.float64).pin_memory(),
              "defect": torch.empty(ns * B, dtype=torch.float64).pin_memory(),
              "iters": torch.empty(ns * B, dtype=torch.int32).pin_memory(), "status": torch.empty(ns * B, dtype=torch.int32).pin_memory(),
              "xs": torch.empty(ns * B, wl.N + 1, solver.NS, dtype=torch.float64).pin_memory() if want_traj else None,
              "us": torch.empty(ns * B, wl.N, solver.NU, dtype=torch.float64).pin_memory() if want_traj else None}
    host_out = {k: (None if v is None else v.numpy()) for k, v in host_t.items()}

    def dev_steps(n):
        o = solver.solve_stream(x0_all_dev[: n * B], trajectories=False, out={k: (None if v is None else v[: n * B]) for k, v in dev_out.items()})
        torch.cuda.synchronize(dev)
        return {k: (None if v is None else v[(n - 1) * B: n * B]) for k, v in o.items()}
    if args.warmup:
        dev_steps(args.warmup)
    launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    out = dev_steps(args.steps)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_total = e0.elapsed_time(e1)
    launches = launch_count(reset=True)
    t0 = time.perf_counter()
    solver.solve_stream_host(x0_all_pin.numpy()[: ns * B], trajectories=want_traj, out=host_out)
    e2e_s = time.perf_counter() - t0
    one = solver.solve(torch.as_tensor(x0_rows, device=dev), trajectories=False)
    assert np.array_equal(one["iters"].cpu().numpy(), out["iters"].cpu().numpy()) and np.array_equal(one["J"].cpu().numpy(), out["J"].cpu().numpy()), \
        "continuous batching and the batched solve disagree"
    step_s = ms_total * 1e-3 / args.steps
    print(json.dumps({"metric": METRIC, "value": B / step_s, "unit": "solves/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": workload_config(args, wl), "gpu_launches": int(launches),
                      "e2e": {"value": B * ns / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": int(x0_rows.nbytes),
                              "d2h_bytes_per_step": int(sum(v.nbytes for v in host_out.values() if v is not None) // ns)}}))
    return 0


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from trajectory_optimization_matrix_lie_groups_b200 import workloads, launch_count
    from trajectory_optimization_matrix_lie_groups_b200.solver import fp64_peak_tflops

    if args.schedule == "stream":
        return run_stream_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    wl = workloads.CONFIGS[args.config](B=B * world)
    depth = max(1, args.inflight)
    want_traj = not args.no_traj
    leg = Leg(torch, dist, wl, B, rank * B, depth, dev, world, want_traj, args.compact)
    solver = leg.solver

    if args.warmup > 0:
        leg.run_steps(args.warmup, False)
    leg.barrier()

    smi_index = torch.cuda.current_device() if os.environ.get("CUDA_VISIBLE_DEVICES") is None else local
    sampler = ClockSampler(smi_index) if rank == 0 else None
    launch_count(reset=True)
    t_lo = time.perf_counter()
    ms_total, out = leg.timed_device(args.steps)
    t_hi = time.perf_counter()
    launches = launch_count(reset=True)
    clocks = sampler.stop(t_lo, t_hi) if sampler else None

    iters = out["iters"].cpu().numpy()
    status = out["status"].cpu().numpy() & 15
    Jfin = out["J"].cpu().numpy()

    # ---- end to end through the host-buffer C-ABI call --------------------------------------
    e2e_s, out_np = leg.timed_host(args.steps, args.e2e_inflight)
    h2d = leg.x0_rows.nbytes
    d2h = sum(v.nbytes for v in out_np.values() if v is not None)
    assert np.array_equal(out_np["iters"], iters), "host-buffer path and device path disagree"
    host_gbs, host_gbs_all = measure_host_d2h(torch, dist, dev, world)

    # ---- one batch alone (no overlap): latency of a step, and the serial throughput for comparison ----
    serial_ms, out1 = leg.timed_single()
    assert np.array_equal(out1["iters"].cpu().numpy(), iters) and np.array_equal(out1["J"].cpu().numpy(), Jfin), \
        "pipelined and serial solves disagree"

    # ---- per-phase device time of one profiled step (event pairs around every launch) -------
    solver.set_profiling(True)
    solver.phase_times(reset=True)
    solver.solve(leg.x0_dev, trajectories=False)
    torch.cuda.synchronize(dev)
    phases = solver.phase_times(reset=True)
    solver.set_profiling(False)
    hist0 = solver.export_hist()["J_hist"][0].cpu().numpy()[: int(iters[0])]

    # ---- the metric as worded: ONE global batch split over the ranks (strong scaling) --------
    strong = None
    G = STRONG_GLOBAL_BATCH[args.config]
    if world > 1 and G % world == 0 and G // world <= 2 * args.batch and not args.no_strong:
        leg.close()
        Bs = G // world
        wl_s = workloads.CONFIGS[args.config](B=G)
        sdepth = auto_inflight(Bs)
        sleg = Leg(torch, dist, wl_s, Bs, rank * Bs, sdepth, dev, world, want_traj, args.compact)
        sleg.run_steps(max(args.warmup, 1), False)
        s_ms, s_out = sleg.timed_device(args.steps)
        s_e2e, s_np = sleg.timed_host(args.steps)
        s_single_ms, _ = sleg.timed_single()
        assert np.array_equal(s_np["iters"], s_out["iters"].cpu().numpy())
        strong = {"global_batch": G, "batch_per_gpu": Bs, "inflight": sdepth, "scaling": "strong",
                  "value": G * args.steps / (s_ms * 1e-3), "unit": "solves/s", "ms_per_step": s_ms / args.steps,
                  "e2e": {"value": G * args.steps / s_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(sleg.x0_rows.nbytes),
                          "d2h_bytes_per_step": int(sum(v.nbytes for v in s_np.values() if v is not None))},
                  "one_sweep_alone": {"ms": s_single_ms, "solves_per_s": G / (s_single_ms * 1e-3),
                                      "note": "ONE 16k-problem sweep split over the ranks, nothing else in flight: its latency is the "
                                              "slowest rank's solve of global_batch / n_gpus problems"},
                  "note": "K sweeps of global_batch problems, each split over the ranks in contiguous shards; `inflight` sweeps in "
                          "flight per GPU like the weak-scaling figure"}
        sleg.close()

    # ---- single-solve latency (the second half of BASELINE.json's metric): config 1, B = 1 ---
    latency = None
    if world == 1 and args.config == 3 and not args.no_latency:
        wl1 = workloads.CONFIGS[1](B=1)
        s1, x1 = wl1.make_solver(B=1, device=dev)
        x1d = torch.as_tensor(x1, device=dev)
        s1.solve(x1d, trajectories=False)
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            o1 = s1.solve_host(x1)
            ts.append(time.perf_counter() - t0)
        J1 = s1.export_hist()["J_hist"][0].cpu().numpy()[: int(o1["iters"][0])]
        latency = {"workload": "main_SE3ddp_tracking_exact.py (BASELINE configs[0]: SE3 single shooting, N=955, 13-step line search), B=1",
                   "ms": 1e3 * statistics.median(ts), "iters": int(o1["iters"][0]), "decision_window_iters": decision_window(J1),
                   "ms_per_iteration": 1e3 * statistics.median(ts) / max(int(o1["iters"][0]), 1), "J": float(o1["J"][0]),
                   "timed": "trajopt_solve_host, host buffers in and out, wall clock, median of 3",
                   "note": "iterations beyond the decision window are the rounding-noise tail of the 13-step line search (cost changes "
                           "< 1e-12 relative); the oracle stops the same problem after 48"}
        s1.close()

    # ---- receding-horizon loop over a batch: closed-loop control steps per second ----
    mpc_block = None
    if world == 1 and args.config == 3 and not args.no_latency:
        from trajectory_optimization_matrix_lie_groups_b200 import mpc
        Bm, Nm, Tm, itm = 4096, 50, 40, 2
        kwm = dict(kind=wl.kind, method="ms", q_ref=wl.q_ref, xi_ref=wl.xi_ref, x0_rows=wl.x0_rows[:Bm], N=Nm, dt=wl.dt, Ib=wl.Ib,
                   mass=wl.mass, Q=wl.Q, R=wl.R, P=wl.P, n_iterations=itm, tol_grad_norm=1e-9, device=dev)
        mpc.receding_horizon(T=4, **kwm)
        rm = mpc.receding_horizon(T=Tm, **kwm)
        mpc_block = {"systems": Bm, "horizon": Nm, "steps": Tm, "ddp_iterations_per_step": itm, "seconds": rm.seconds,
                     "control_steps_per_s": rm.steps_per_second, "ms_per_step_of_the_batch": 1e3 * rm.seconds / Tm,
                     "note": "mpc.receding_horizon: warm-started truncated solves + plant step, reference window slid on the device "
                             "(trajopt_set_reference_long / _offset); state, warm start and the logged loop stay in HBM"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    model = workloads.WORK_MODEL[wl.kind]
    sweeps = float(np.sum(iters + (status == 0)))           # backward sweeps executed (a converged problem runs one more)
    rollouts = float(np.sum(iters))
    bwd_ms, bwd_n = phases["backward"]
    lin_ms, lin_n = phases["linearize"]
    fwd_ms, fwd_n = phases["forward"]
    oth_ms, _ = phases["other"]
    phase_total = bwd_ms + lin_ms + fwd_ms + oth_ms
    psampler = ClockSampler(smi_index)
    tp0 = time.perf_counter()
    peak = fp64_peak_tflops(400.0, dev)
    tp1 = time.perf_counter()
    peak_clocks = psampler.stop(tp0, tp1)
    if wl.method == "al_ms":
        # every outer iteration restarts the inner solve: the per-problem sweep count is not in the exported summaries,
        # so the launches are counted as full-batch ones (an upper bound on the work, hence on the fraction)
        sweeps = float(bwd_n) * B
        rollouts = float(fwd_n) * B
    flop_bwd = model["flop_bwd"] * wl.N * sweeps
    achieved = flop_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0
    counters = load_kernel_counters(wl.kind, B)
    traffic = executed = None
    if counters:
        traffic = counters.get("dram_bytes_per_launch")
        fpu = (2.0 * counters["dfma"] + counters["dmul"] + counters["dadd"]) / counters["units"]
        executed = {"flop_executed_per_unit": fpu, "flop_model_per_unit": model["flop_bwd"],
                    "executed_tflops": fpu * wl.N * sweeps / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0,
                    "fp64_pipe_active_frac_ncu": counters.get("fp64_pipe_active_frac"), "source": counters.get("source")}
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        with open(ppath) as f:
            peaks = json.load(f)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    step_s = ms_total * 1e-3 / args.steps
    flop_step = (model["flop_bwd"] + model["flop_lin"]) * wl.N * sweeps + model["flop_fwd"] * wl.N * rollouts
    bytes_step = model["bytes_bwd"] * wl.N * sweeps + model["bytes_fwd"] * wl.N * rollouts

    value = B * world / step_s
    e2e_value = B * world * args.steps / e2e_s
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, wl),
        "value_returns": "per-problem summaries (J, grad, defect, iters, status); trajectories stay in HBM — `e2e` exports and copies them",
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "returns": "xs, us, J, grad, defect, iters, status" if want_traj else "J, grad, defect, iters, status",
                "host_d2h_gbs": host_gbs_all, "host_d2h_gbs_this_rank": host_gbs,
                "host_d2h_used_frac": (d2h * world * args.steps / e2e_s / 1e9) / host_gbs_all if host_gbs_all else None,
                "host_d2h_note": "concurrent pinned cudaMemcpyAsync of 1 GiB per rank, all ranks at once, summed: the ceiling the e2e "
                                 "leg's trajectory copies share"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "kernel": sweep_kernel_name(wl.kind, B), "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "frac_note": "model fraction: SURVEY 8d's DENSE flop count per stage-iteration over the measured time; the kernel "
                                  "skips the structural zeros of f_x / f_u, see executed_frac for the FLOPs it actually issues",
                     "executed_frac": executed["executed_tflops"] / peak if executed and peak else None,
                     "executed": executed,
                     "peak_source": "DFMA microbenchmark run live by this bench (no FP64 entry in MEASURED_PEAKS.json)",
                     "peak_clocks": peak_clocks,
                     "flop_per_launch": flop_bwd / max(bwd_n, 1), "launch_ms": bwd_ms / max(bwd_n, 1), "launches": int(bwd_n),
                     "share_of_step": bwd_ms / phase_total if phase_total else None},
        "solve_roofline": {"fp64_frac": flop_step / step_s / 1e12 / peak if peak else None,
                           "hbm_frac": bytes_step / step_s / 1e9 / hbm_peak, "hbm_peak_gbs": hbm_peak,
                           "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                           "algorithmic_gflop_per_step": flop_step / 1e9, "algorithmic_gb_per_step": bytes_step / 1e9},
        "kernels": [
            {"kernel": "k_linearize", "bound": "hbm", "ms_per_launch": lin_ms / max(lin_n, 1), "share_of_step": lin_ms / phase_total,
             "achieved": workloads.LIN_RECORD_DOUBLES[wl.kind] * 8.0 * (wl.N + 1) * sweeps / (lin_ms * 1e-3) / 1e9, "peak": hbm_peak,
             "unit": "GB/s", "frac": workloads.LIN_RECORD_DOUBLES[wl.kind] * 8.0 * (wl.N + 1) * sweeps / (lin_ms * 1e-3) / 1e9 / hbm_peak,
             "note": "bytes = the records the kernel writes by design (the SURVEY byte model assumes none: linearisation recomputed in the sweep)"},
            {"kernel": "k_forward_ms_full" if wl.method != "ss" else "k_forward", "bound": "latency (sequential Exp/Log chain); hbm secondary",
             "ms_per_launch": fwd_ms / max(fwd_n, 1), "share_of_step": fwd_ms / phase_total,
             "achieved": model["bytes_fwd"] * wl.N * rollouts / (fwd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
             "frac": model["bytes_fwd"] * wl.N * rollouts / (fwd_ms * 1e-3) / 1e9 / hbm_peak},
        ] if lin_ms > 0 and fwd_ms > 0 else [],
        "single_batch": {"ms": serial_ms, "solves_per_s": B / (serial_ms * 1e-3),
                         "note": "one batch alone on the GPU (nothing else in flight): step latency / serial throughput"},
        "phases_ms": {"linearize": lin_ms, "backward": bwd_ms, "forward": fwd_ms, "other": oth_ms,
                      "launches": {"linearize": int(lin_n), "backward": int(bwd_n), "forward": int(fwd_n)}},
        "solve_stats": {"iters_mean": float(iters.mean()), "iters_min": int(iters.min()), "iters_max": int(iters.max()),
                        "converged_frac": float(np.mean(status == 0)), "J_problem0": float(Jfin[0]),
                        "decision_window_iters": decision_window(hist0), "iters_problem0": int(iters[0]),
                        "decision_window_note": "iterations of problem 0 whose cost change is above rounding noise (>= 1e-12 relative): "
                                                "the span over which iteration counts / accepted step sizes are asserted equal to the "
                                                "reference's (tests/test_gpu_golden.py, test_gpu_configs.py: 33 whole solves of this batch, "
                                                "16..27 iterations, all inside their windows)"},
    }
    if strong is not None:
        line["strong"] = strong
    elif world == 1 and G == B:
        line["strong"] = {"global_batch": B, "batch_per_gpu": B, "scaling": "strong", "value": value, "unit": "solves/s",
                          "e2e": {"value": e2e_value, "unit": "solves/s"}, "one_sweep_alone": {"ms": serial_ms, "solves_per_s": B / (serial_ms * 1e-3)},
                          "note": "at one GPU the strong- and weak-scaling workloads coincide"}
    if latency is not None:
        line["single_solve_latency"] = latency
    if mpc_block is not None:
        line["mpc"] = mpc_block
    if world == 1 and not args.no_cpu:
        try:
            # whole solves of the first `cores` problems (about 10-30 s of CPU work on the box's cores); configs whose solves
            # are long (AL: thousands of iterations) keep the bounded extrapolated sample
            v, cores, desc = cpu_sample_subprocess(args.config, 0 if args.config in (2, 3, 5) else 3, float(iters.mean()) + 1.0)
            line["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": desc}
        except Exception as e:   # the baseline is context, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "solves/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    if len(sys.argv) >= 5 and sys.argv[1] == "--cpu-sample":
        v, cores, desc = cpu_sample(int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]))
        print(json.dumps({"value": v, "cores": cores, "desc": desc}))
        return 0
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU")
    ap.add_argument("--no-traj", action="store_true", help="e2e leg returns only the per-problem summaries")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (N > 1)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-solve latency leg (N = 1, config 3)")
    ap.add_argument("--inflight", type=int, default=0,
                    help="batches in flight per GPU (1 = strictly one after the other; 0 = auto: 3, or 8 for batches <= 2048)")
    ap.add_argument("--e2e-inflight", type=int, default=0,
                    help="batches in flight in the host-buffer (e2e) leg; 0 = same as --inflight")
    ap.add_argument("--schedule", default="pipelined", choices=["stream", "pipelined"],
                    help="pipelined (default): whole batches, --inflight of them at a time; stream: continuous batching, the K "
                         "steps' problems are one queue through batch_per_gpu slots (trajopt_solve_stream) — measured slower on "
                         "the headline workload (32.6 k against 37.8 k solves/s, see DESIGN.md section 5)")
    ap.add_argument("--compact", default="", help="'min_batch,ratio' for trajopt_set_compaction (default: the library's 1024,4; '-1,4' = off)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {1: 1, 2: 1024, 3: 16384, 4: 2048, 5: 131072}[args.config]
    if args.inflight <= 0:
        args.inflight = auto_inflight(args.batch)
    if args.schedule == "stream" and args.config == 4:
        raise SystemExit("--schedule stream: the augmented-Lagrangian method is solved per batch (use pipelined)")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
