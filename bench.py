#!/usr/bin/env python
"""Headline benchmark: batched SE3 multiple-shooting DDP tracking solves per second (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 3] [--batch B]

A "step" is one pass of the hot path over one synthetic batch: a complete `fit` (to the script's
own stopping rule) of every problem of the batch.  Default workload = BASELINE.json configs[2],
`main_SE3ddp_tracking_exact_ms.py` (SE3, multiple shooting, N=955, dt=0.004, tol_grad 1e-12), batch
16384 perturbed initial poses PER GPU (weak scaling: rank r solves problems [r*B, (r+1)*B) of the
global seeded batch; no collective on the data path, one NCCL all-gather of the per-problem
summaries (J, grad, defect, iters, status) at the end of every step).

`value`   device-timed (CUDA events, max over ranks) with the initial states resident in HBM.
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
    """One bounded CPU sample: one problem per core, `n_iters` DDP iterations each.

    Returns (solves_per_second extrapolated to whole solves, cores, description)."""
    import multiprocessing as mp
    # one single-threaded process per core (the reference's joblib sweep); BLAS pools on top of that only oversubscribe
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    cores = cores or int(os.environ.get("TRAJOPT_BENCH_CORES", "0")) or len(os.sched_getaffinity(0))
    batch = max(cores, 12)
    jobs = [(config, b, n_iters, batch) for b in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    done_iters = sum(r[1] for r in res)
    busy = sum(r[0] for r in res)
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
                           repr(float(iters_per_solve))], capture_output=True, text=True, timeout=900)
    if proc.returncode != 0:
        raise RuntimeError(proc.stderr[-400:])
    d = json.loads(proc.stdout.strip().splitlines()[-1])
    return d["value"], d["cores"], d["desc"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    its = float(NOMINAL_ITERS[args.config])
    n_it = int(os.environ.get("TRAJOPT_BENCH_CPU_ITERS", "2"))
    if args.warmup > 0:
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


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from trajectory_optimization_matrix_lie_groups_b200 import workloads, launch_count, PipelinedSolver
    from trajectory_optimization_matrix_lie_groups_b200.solver import fp64_peak_tflops

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
    stream_mode = args.schedule == "stream"
    depth = 1 if stream_mode else max(1, args.inflight)
    x0_rows = wl.x0_rows[rank * B:rank * B + B]
    pipe = PipelinedSolver(lambda: wl.make_solver(B=B, device=dev, offset=rank * B)[0], depth=depth, device=dev)
    solver = pipe.solvers[0]
    if args.compact:
        mb, ratio = (int(v) for v in args.compact.split(","))
        for sv in pipe.solvers:
            sv.set_compaction(mb, ratio)
    x0_dev = torch.as_tensor(x0_rows, device=dev)
    x0_pin = torch.as_tensor(x0_rows).pin_memory()
    want_traj = not args.no_traj

    def pinned_out():
        o = {"J": torch.empty(B, dtype=torch.float64).pin_memory(), "grad": torch.empty(B, dtype=torch.float64).pin_memory(),
             "defect": torch.empty(B, dtype=torch.float64).pin_memory(),
             "iters": torch.empty(B, dtype=torch.int32).pin_memory(), "status": torch.empty(B, dtype=torch.int32).pin_memory(),
             "xs": torch.empty(B, wl.N + 1, solver.NS, dtype=torch.float64).pin_memory() if want_traj else None,
             "us": torch.empty(B, wl.N, solver.NU, dtype=torch.float64).pin_memory() if want_traj else None}
        return o, {k: (None if v is None else v.numpy()) for k, v in o.items()}
    outs = [pinned_out() for _ in range(depth)] if not stream_mode else []   # one set of pinned result buffers per lane
    if stream_mode:
        # continuous batching: the problems of n steps are ONE queue of n x B problems through the solver's B slots
        n_max = max(args.steps, args.warmup, 1)
        x0_all_dev = x0_dev.repeat(n_max, 1)
        x0_all_pin = torch.as_tensor(np.tile(x0_rows, (n_max, 1))).pin_memory()
        M = n_max * B
        dev_out = {"J": torch.empty(M, dtype=torch.float64, device=dev), "grad": torch.empty(M, dtype=torch.float64, device=dev),
                   "defect": torch.empty(M, dtype=torch.float64, device=dev), "iters": torch.empty(M, dtype=torch.int32, device=dev),
                   "status": torch.empty(M, dtype=torch.int32, device=dev), "xs": None, "us": None}
        ns = args.steps
        host_out_t = {"J": torch.empty(ns * B, dtype=torch.float64).pin_memory(), "grad": torch.empty(ns * B, dtype=torch.float64).pin_memory(),
                      "defect": torch.empty(ns * B, dtype=torch.float64).pin_memory(),
                      "iters": torch.empty(ns * B, dtype=torch.int32).pin_memory(), "status": torch.empty(ns * B, dtype=torch.int32).pin_memory(),
                      "xs": torch.empty(ns * B, wl.N + 1, solver.NS, dtype=torch.float64).pin_memory() if want_traj else None,
                      "us": torch.empty(ns * B, wl.N, solver.NU, dtype=torch.float64).pin_memory() if want_traj else None}
        host_out = {k: (None if v is None else v.numpy()) for k, v in host_out_t.items()}
    gathered = [torch.empty(B, 5, dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def gather(out):
        if world > 1:   # the one collective of the path: per-problem summaries to every rank
            summ = torch.stack((out["J"], out["grad"], out["defect"], out["iters"].double(), out["status"].double()), dim=1)
            dist.all_gather(gathered, summ)

    def run_steps(n, host):
        """n steps, `depth` batches in flight; steps complete (and are gathered) in submission order."""
        if stream_mode:
            if host:
                n = min(n, args.steps)
                o = solver.solve_stream_host(x0_all_pin.numpy()[: n * B], trajectories=want_traj,
                                             out={k: (None if v is None else v[: n * B]) for k, v in host_out.items()})
                return {k: (None if v is None else v[(n - 1) * B: n * B]) for k, v in o.items()}
            o = solver.solve_stream(x0_all_dev[: n * B], trajectories=False,
                                    out={k: (None if v is None else v[: n * B]) for k, v in dev_out.items()})
            torch.cuda.current_stream(dev).synchronize()
            last = None
            for k in range(n):
                last = {key: (None if v is None else v[k * B:(k + 1) * B]) for key, v in o.items()}
                gather(last)
            return last
        futs = []
        for k in range(n):
            if host:
                futs.append(pipe.submit(x0_pin.numpy(), trajectories=want_traj, host=True, out=outs[k % depth][1]))
            else:
                futs.append(pipe.submit(x0_dev, trajectories=False))
        last = None
        for f in futs:
            last = f.result()
            if not host:
                gather(last)
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if args.warmup > 0:
        out = run_steps(args.warmup, False)
    barrier()

    sampler = ClockSampler(torch.cuda.current_device() if os.environ.get("CUDA_VISIBLE_DEVICES") is None else local) if rank == 0 else None
    launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_lo = time.perf_counter()
    barrier()
    e0.record()
    out = run_steps(args.steps, False)      # every lane synchronises its stream before its result is handed back
    e1.record()
    barrier()
    t_hi = time.perf_counter()
    launches = launch_count(reset=True)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_lo, t_hi) if sampler else None

    iters = out["iters"].cpu().numpy()
    status = out["status"].cpu().numpy() & 15
    Jfin = out["J"].cpu().numpy()

    # ---- end to end through the host-buffer C-ABI call --------------------------------------
    if args.e2e_inflight > 0 and not stream_mode:
        pipe.active_depth = min(args.e2e_inflight, depth)
    run_steps(min(depth, 2), True)
    barrier()
    t0 = time.perf_counter()
    run_steps(args.steps, True)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = x0_rows.nbytes
    if stream_mode:
        out_np = {k: (None if v is None else v[(args.steps - 1) * B: args.steps * B]) for k, v in host_out.items()}
    else:
        out_np = outs[(args.steps - 1) % depth][1]
    d2h = sum(v.nbytes for v in out_np.values() if v is not None)
    assert np.array_equal(out_np["iters"], iters), "host-buffer path and device path disagree"

    pipe.active_depth = depth
    # ---- one batch alone (no overlap): latency of a step, and the serial throughput for comparison ----
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    out1 = solver.solve(x0_dev, trajectories=False)
    s1.record()
    torch.cuda.synchronize(dev)
    serial_ms = s0.elapsed_time(s1)
    assert np.array_equal(out1["iters"].cpu().numpy(), iters) and np.array_equal(out1["J"].cpu().numpy(), Jfin), \
        "pipelined and serial solves disagree"

    # ---- per-phase device time of one profiled step (event pairs around every launch) -------
    solver.set_profiling(True)
    solver.phase_times(reset=True)
    solver.solve(x0_dev, trajectories=False)
    torch.cuda.synchronize(dev)
    phases = solver.phase_times(reset=True)
    solver.set_profiling(False)

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
    peak = fp64_peak_tflops(60.0, dev)
    flop_bwd = model["flop_bwd"] * wl.N * sweeps
    achieved = flop_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("k_backward_dram_bytes_per_launch")
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
    e2e_value = B * world * args.steps / float(e2e_s.item())
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, wl),
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "returns": "xs, us, J, grad, defect, iters, status" if want_traj else "J, grad, defect, iters, status"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "kernel": "k_backward3 (Riccati sweep, 2-warp CTA per 32 problems)" if wl.kind != "so3" else "k_backward (Riccati sweep)", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "peak_source": "DFMA microbenchmark run live by this bench (no FP64 entry in MEASURED_PEAKS.json)",
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
                        "converged_frac": float(np.mean(status == 0)), "J_problem0": float(Jfin[0])},
    }
    if world == 1 and not args.no_cpu:
        try:
            v, cores, desc = cpu_sample_subprocess(args.config, 3, float(iters.mean()) + 1.0)
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
        args.inflight = 3 if args.batch > 2048 else 8
    if args.schedule == "stream" and args.config == 4:
        raise SystemExit("--schedule stream: the augmented-Lagrangian method is solved per batch (use pipelined)")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
