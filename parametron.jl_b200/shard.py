"""Instance-shard data parallelism: a batch of independent problem instances is cut into one contiguous
shard per GPU; there is NO data-path collective (every instance's (Q, q, c0, C_i, d_i) stays on the GPU that produced it).
`torch.distributed` is used only for the timing barrier and the max-over-ranks reduction of the measured time."""


def instance_shard(total, rank, world):
    """(first_instance, count) of `rank`'s contiguous shard: ceil(total / world) instances per rank, the tail ranks get the rest."""
    if total < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    per = -(-total // world)
    first = min(rank * per, total)
    return first, max(0, min(per, total - first))


def aggregate_throughput(units_per_rank, seconds_per_rank):
    """whole-job throughput = all units processed / the slowest rank's time (never a sum of per-rank rates)"""
    return sum(units_per_rank) / max(seconds_per_rank)


def max_over_ranks(value, device=None):
    """max of a python float over the process group (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count, device=None):
    """every rank's unit count, on every rank"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(count)]
    t = torch.tensor([int(count)], dtype=torch.int64, device=device if device is not None else "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]
