"""Worker of the data-parallel learner tests, one process per rank (run under torch.distributed.run).

Every rank takes its slice of each minibatch through ddpg_update_dp (gradient exchange fused into the optimiser kernels over
CUDA-IPC peer memory).  Rank 0 checks that the replicas are bit-identical and equal one learner on the full minibatch, and
prints DP_OK.  With fewer GPUs than ranks the ranks share a device (separate processes time-slice it; the IPC path is the same)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shems_b200 as sb  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(device)
    dist.init_process_group("gloo")
    n, T, K = 64, 72, 3
    B, tc = int(os.environ.get("DP_BATCH", "32")), int(os.environ.get("DP_TC", "0"))   # per-rank minibatch; TF32 tensor cores
    ser = sb.series.synth_charger98(4320, seed=98)
    env = sb.Shems(T, ser, n_envs=n, device=device)
    mem = sb.Replay(n * T, device=device)
    env.reset(rng=2)
    env.rollout(sb.POLICY_RANDOM, T, seed=2, replay=mem, want_return=False)       # identical transitions on every rank
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)
    idx = np.random.default_rng(0).integers(0, len(mem), (K, world * B)).astype(np.int32)
    le = sb.Learner(params=sb.default_ddpg_params(batch=B, use_tensor_cores=tc), device=device)
    le.init(3)
    le.set_norm(mn, mx)
    le.dp_connect_dist(dist)
    dist.barrier()
    le.replay_fused_dp(mem, n_updates=K, idx=idx[:, rank * B:(rank + 1) * B])
    status = le.dp_status()
    mine = [le.get_layer(net, k) for net in range(4) for k in range(3)]
    everyone = [None] * world if rank == 0 else None
    dist.gather_object((status, mine), everyone, dst=0)
    if rank == 0:
        assert all(st == 0 for st, _ in everyone), [st for st, _ in everyone]
        full = sb.Learner(params=sb.default_ddpg_params(batch=world * B, use_tensor_cores=tc), device=device)
        full.init(3)
        full.set_norm(mn, mx)
        full.replay(mem, n_updates=K, idx=idx)
        p = full.p
        lrs = [p.lr_actor, p.lr_critic, p.lr_actor * p.tau, p.lr_critic * p.tau]
        for j, (net, k) in enumerate((net, k) for net in range(4) for k in range(3)):
            wf, bf = full.get_layer(net, k)
            for r in range(world):
                w, b = everyone[r][1][j]
                np.testing.assert_array_equal(w, everyone[0][1][j][0])           # bit-identical replicas
                np.testing.assert_array_equal(b, everyone[0][1][j][1])
                if not tc:
                    np.testing.assert_allclose(w, wf, rtol=1e-5, atol=0.02 * lrs[net] * K + 1e-7)
                    np.testing.assert_allclose(b, bf, rtol=1e-5, atol=0.02 * lrs[net] * K + 1e-7)
                else:  # TF32 products on both sides, different tiling of the batch: ADAM sign noise on near-zero gradients
                    d = np.abs(w - wf)
                    assert d.max() <= 2.0 * lrs[net] * K + 1e-6 and np.quantile(d, 0.99) <= 0.1 * lrs[net] * K + 1e-7, (net, k, d.max())
    # a native training episode on a connected learner (ddpg_episode): every rank steps ITS OWN instances (different env ids ->
    # different noise, different transitions), every replay() exchanges gradients -> the replicas must stay bit-identical
    env2 = sb.Shems(T, ser, n_envs=4, device=device, env_id_base=rank * 4)
    env2.reset(rng=5)
    le.dp_prepare(0)
    ret = le.episode(env2, mem, 6, train=True, sigma=0.1, rng_ep=5, updates_per_step=1)
    assert le.dp_status() == 0
    mine = ([le.get_layer(net, k) for net in range(4) for k in range(3)], ret.cpu().numpy())
    everyone = [None] * world if rank == 0 else None
    dist.gather_object(mine, everyone, dst=0)
    if rank == 0:
        for r in range(1, world):
            for (w0, b0), (w, b) in zip(everyone[0][0], everyone[r][0]):
                np.testing.assert_array_equal(w, w0)
                np.testing.assert_array_equal(b, b0)
            if world > 1 and torch.cuda.device_count() >= 1:
                assert not np.array_equal(everyone[0][1], everyone[r][1])          # the ranks really saw different episodes
        print("DP_OK world=%d devices=%d batch=%d tc=%d" % (world, torch.cuda.device_count(), B, tc), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
