"""Multi-GPU host logic: independent stereo sequences partitioned over ranks, no data-path collective.

A single camera stream cannot shard (frame t+1 needs frame t's pyramids and tracks: reference
src/feature_tracker/feature_tracker.rs:94-96,173-174), so scaling comes from replicas only: every
rank owns a block of streams on its own GPU.  The only communication is the reduction of the final
counters (sum) and of the device time (max over ranks) for reporting.
"""
import os


def rank_info():
    """(rank, local_rank, world_size) from the torchrun environment (single process: 0, 0, 1)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def partition_streams(total_streams, world_size, rank):
    """Contiguous block partition of stream ids [0, total) over ranks; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size) or total_streams < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(total_streams, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def weak_streams(streams_per_gpu, rank):
    """Weak scaling: every rank runs `streams_per_gpu` streams with globally distinct ids (seeds)."""
    return list(range(rank * streams_per_gpu, (rank + 1) * streams_per_gpu))


def aggregate(dist, device, tracks, frames, iterations, ms):
    """Whole-job counters: sums over ranks, and the slowest rank's device time.
    `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(tracks), float(frames), float(iterations), float(ms)
    import torch
    s = torch.tensor([float(tracks), float(frames), float(iterations)], dtype=torch.float64, device=device)
    m = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(s[0]), float(s[1]), float(s[2]), float(m[0])


def throughput(tracks, frames, ms):
    """(tracks/s, stereo frames/s) of the whole job."""
    sec = ms * 1e-3
    return tracks / sec, frames / sec
