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


POPULATION_CHARGERS = (1, 2, 3, 4, 5, 6, 7, 8, 9, 98)   # capacities of shems_LU1.jl:47-59 without the spare id 97


def population_shard(rank, world, chargers=POPULATION_CHARGERS, seeds_per_charger=64):
    """BASELINE configs[4] — `len(chargers)` chargers x `seeds_per_charger` seeds = one independent learner each (the reference
    starts one Julia process per (JOB_ID, seed), RL-SHEMS_bs_scheduler_*.sh:73-81) — split over `world` ranks in contiguous,
    balanced blocks of global learner ids g = charger_index * seeds_per_charger + seed_index.  Returns the rank's
    (global_ids, charger_ids, seeds): learners of one charger stay adjacent, which is what PopulationDriver's instance groups
    need; seeds follow input.jl:136 (`rng_run = parse(Int, "123" * "$seed")`)."""
    total = len(chargers) * int(seeds_per_charger)
    base, count = shard_range(total, rank, world)
    gids = list(range(base, base + count))
    cids = [int(chargers[g // seeds_per_charger]) for g in gids]
    seeds = [int("123" + str(g % seeds_per_charger + 1)) for g in gids]
    return gids, cids, seeds
