"""Frame-index sharding across GPUs.

Frames are independent, so rank g of N owns the contiguous slice
[g*ceil(F/N), min(F, (g+1)*ceil(F/N))) and no data-path collective exists.  The only
cross-rank traffic is the harness reducing N timing scalars (max) and N unit counts (sum).
"""


def frame_slice(total_frames, rank, world_size):
    """Contiguous frame range of `rank` when `total_frames` are split over `world_size` GPUs."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    per = -(-total_frames // world_size)
    lo = min(total_frames, rank * per)
    hi = min(total_frames, lo + per)
    return lo, hi


def gather_job_stats(local_seconds, local_units, group=None):
    """(max seconds over ranks, sum of units) via torch.distributed; identity when not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_seconds), float(local_units)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), float(u.item())
