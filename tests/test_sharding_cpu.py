"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard ranges, rank-invariant RNG keying, max-reduce timing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import shems_b200 as sb
    from shems_b200 import sharding
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, count = sharding.shard_range(n_total, rank, world)
    ser = sb.series.synth_charger98(600, seed=3)
    env = O.OracleEnv(O.params_for_charger(98), ser, T, count)          # the oracle stands in for the CUDA shard on CPU
    env.reset(mode=2, seed=21, env_id_base=base)
    ret = env.rollout(1, T, seed=21, env_id_base=base)["ep_return"]
    # pad to equal length for all_gather, then trim
    pad = torch.zeros(n_total // world + 1, dtype=torch.float64)
    pad[:count] = torch.from_numpy(ret)
    allr = sharding.gather_concat(pad, dist).reshape(world, -1)
    t = sharding.max_over_ranks(10.0 + rank, dist)
    if rank == 0:
        counts = [sharding.shard_range(n_total, r, world)[1] for r in range(world)]
        q.put((np.concatenate([allr[r, :counts[r]].numpy() for r in range(world)]), t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_rollout_equals_single(O, sb):
    from shems_b200 import sharding
    assert [sharding.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 3), (7, 3)]
    assert sharding.weak_range(1 << 20, 3) == (3 << 20, 1 << 20)
    n_total, T, world = 101, 48, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, T, q)) for r in range(world)]
    [p.start() for p in procs]
    got, tmax = q.get()
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ser = sb.series.synth_charger98(600, seed=3)
    env = O.OracleEnv(O.params_for_charger(98), ser, T, n_total)
    env.reset(mode=2, seed=21)
    want = env.rollout(1, T, seed=21)["ep_return"]
    np.testing.assert_array_equal(got, want)   # results do not depend on the number of ranks
    assert tmax == 11.0                        # device time = max over ranks


def test_population_shard_covers_every_learner_once():
    """configs[4]: 10 chargers x 64 seeds over 1/2/4/8 ranks — disjoint, complete, balanced; a rank's learners of one charger are
    adjacent (instance groups) and the seeds follow input.jl:136."""
    from shems_b200 import sharding
    for world in (1, 2, 4, 8, 3):
        seen = []
        for rank in range(world):
            gids, cids, seeds = sharding.population_shard(rank, world)
            assert len(gids) == len(cids) == len(seeds) and abs(len(gids) - 640 / world) < 1
            assert gids == list(range(gids[0], gids[0] + len(gids)))
            runs = [c for i, c in enumerate(cids) if i == 0 or cids[i - 1] != c]
            assert len(runs) == len(set(runs))                     # equal chargers adjacent
            for g, c, s in zip(gids, cids, seeds):
                assert c == sharding.POPULATION_CHARGERS[g // 64] and s == int("123%d" % (g % 64 + 1))
            seen += gids
        assert sorted(seen) == list(range(640))
    gids, cids, seeds = sharding.population_shard(0, 8)
    assert len(gids) == 80 and cids[:64] == [1] * 64 and cids[64:] == [2] * 16 and seeds[0] == 1231 and seeds[63] == 12364 and seeds[64] == 1231
