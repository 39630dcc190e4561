"""Stream fan-out across the GPUs of one box (replicas only, no data-path collective).

Stream i runs on rank i mod world; a rank's streams are independent encoder sessions.  The only
communication is the measurement plumbing: a barrier and a MAX over the ranks' elapsed times,
done with torch.distributed on whatever backend the process group uses (NCCL on the GPU box,
gloo in the CPU tests)."""
from typing import List


def streams_of_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Indices of the streams rank `rank` encodes (round-robin: stream i -> rank i mod world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [i for i in range(n_streams) if i % world == rank]


def stream_seed(stream: int) -> int:
    """Synthetic-content seed of a stream (distinct content per stream)."""
    return stream


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity without a process group)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.cpu()]


def aggregate_throughput(frames_this_rank: int, elapsed_ms_this_rank: float, device=None) -> float:
    """Whole-job frames/s: all ranks' frames over the slowest rank's time."""
    total = sum_over_ranks([frames_this_rank], device)[0]
    worst = max_over_ranks([elapsed_ms_this_rank], device)[0]
    return total / (worst * 1e-3)
