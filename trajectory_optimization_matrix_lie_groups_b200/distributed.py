"""Multi-GPU sharding of a batch of independent problems.

The reference parallelises a sweep as unrelated joblib jobs (visualization/perturb_all_compute.py:
243-250); here the batch is cut into contiguous shards, one per rank (one process per GPU), each
rank solves its shard with no communication, and the per-problem summaries
(J, grad, defect, iters, status) = 32 B/problem are all-gathered once at the end (NCCL over NVLink on
GPUs; the same code runs over gloo on CPU tensors for the host-logic tests).
"""
import torch
import torch.distributed as dist

SUMMARY_FIELDS = ("J", "grad", "defect", "iters", "status")


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(B, rank, world_size):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the larger blocks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_summary(out):
    """dict of per-problem tensors -> (b, 5) float64 tensor (iters/status are exact in float64)."""
    return torch.stack([out[k].to(torch.float64) for k in SUMMARY_FIELDS], dim=1).contiguous()


def unpack_summary(t):
    return {"J": t[:, 0].contiguous(), "grad": t[:, 1].contiguous(), "defect": t[:, 2].contiguous(),
            "iters": t[:, 3].to(torch.int32), "status": t[:, 4].to(torch.int32)}


def all_gather_summaries(local, B):
    """All-gather the ranks' (b_r, 5) summaries into the (B, 5) table of the whole batch, in problem order.

    Shards may differ by one row, so every rank pads to the largest shard before the collective.
    """
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_bounds(B, r, ws) for r in range(ws)]
    bmax = max(hi - lo for lo, hi in sizes)
    # NCCL gathers device tensors in place; a gloo job (the CPU tests, or several ranks sharing one GPU) goes through the host
    via_host = local.is_cuda and dist.get_backend() != "nccl"
    src = local.cpu() if via_host else local
    pad = torch.zeros(bmax, src.shape[1], dtype=src.dtype, device=src.device)
    pad[: src.shape[0]] = src
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    table = torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
    return table.to(local.device) if via_host else table


def solve_sharded(make_solver, x0_rows, us_init=None, trajectories=False):
    """Solve this rank's shard of `x0_rows` (B, NS) and return (local result dict, global summary dict, (lo, hi)).

    make_solver(b) -> BatchSolver for b problems on this rank's device.
    """
    rank, ws = world()
    B = x0_rows.shape[0]
    lo, hi = shard_bounds(B, rank, ws)
    solver = make_solver(hi - lo)
    us = us_init
    if us is not None and getattr(us, "ndim", 2) == 3:
        us = us[lo:hi]
    out = solver.solve(x0_rows[lo:hi], us, trajectories=trajectories)
    table = all_gather_summaries(pack_summary(out), B)
    return out, unpack_summary(table), (lo, hi)
