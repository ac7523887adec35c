"""world_size-2 gloo test of the multi-GPU harness logic (no GPU): frame sharding + the two
scalar reductions bench.py performs (max of times, sum of units).  No data-path collective exists."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from metalbt709decoder_b200.sharding import frame_slice, gather_job_stats
    lo, hi = frame_slice(601, rank, world)
    # every rank "processes" its slice; rank 1 is slower
    secs, units = gather_job_stats(0.5 + rank, float(hi - lo) * 1920 * 1080)
    dist.barrier()
    out.put((rank, lo, hi, secs, units))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, u0), (r1, lo1, hi1, s1, u1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 301, 301, 601)     # contiguous, disjoint, complete
    assert s0 == s1 == 1.5                                  # max over ranks
    assert u0 == u1 == 601.0 * 1920 * 1080                  # sum over ranks = whole job


def test_gather_is_identity_without_process_group():
    from metalbt709decoder_b200.sharding import gather_job_stats
    assert gather_job_stats(0.25, 7.0) == (0.25, 7.0)
