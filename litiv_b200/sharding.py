"""Stream sharding across GPUs: independent video streams are dealt round-robin to ranks (one process per GPU);
there is no data-path collective because there is no cross-stream state. Only timings are reduced."""


def shard_streams(n_streams, rank, world):
    """stream ids owned by `rank` (stream i -> GPU i mod world)"""
    return [s for s in range(n_streams) if s % world == rank]


def aggregate_max_ms(local_ms, dist=None, device=None):
    """max over ranks of a locally measured duration (ms); identity for a single process"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms)
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
