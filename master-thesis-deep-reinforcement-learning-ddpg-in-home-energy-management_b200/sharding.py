"""Multi-GPU host logic: instances shard across ranks by contiguous global env-id ranges, no data-path collective.

Philox streams are keyed by the GLOBAL env id (`env_id_base` of the C ABI), so a sharded job computes exactly what one
big handle would — the number of ranks never changes a result.  torch.distributed is plumbing only (barrier, max-reduce of
the device time, gathering per-instance returns for reporting)."""
import torch


def shard_range(n_total, rank, world):
    """(env_id_base, count) of `rank`: balanced contiguous split of n_total instances (strong scaling)."""
    q, r = divmod(int(n_total), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def weak_range(n_per_rank, rank):
    """(env_id_base, count) with per-rank work fixed (weak scaling; what bench.py uses)."""
    return rank * int(n_per_rank), int(n_per_rank)


def max_over_ranks(value, dist=None, device="cpu"):
    """Device times are reported as the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_concat(x, dist=None):
    """Concatenate equally-shaped per-rank tensors in rank order (per-instance episode returns, for reporting)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    out = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(out, x)
    return torch.cat(out)
