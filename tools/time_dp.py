"""Times the data-parallel learner (ddpg_update_dp: gradient exchange fused into the optimiser kernels over NVLink peer memory) on N ranks.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/time_dp.py [n_updates]
-> one JSON line from rank 0 (us per update, max over ranks, CUDA events)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import shems_b200 as sb  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_updates = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
TRACE = "dptrace" in os.environ.get("SHEMS_B200_LIB", "")   # a -DDP_TRACE build: globaltimer stamps inside the exchange kernel
ser = sb.series.synth_charger98(4320, seed=98)
env = sb.Shems(72, ser, n_envs=1000, device=local, env_id_base=rank * 1000)
mem = sb.Replay(24_000, device=local)
env.reset(rng=1)
env.rollout(sb.POLICY_RANDOM, 24, seed=1, replay=mem, want_return=False)
mn, mx = mem.min_max_buffer(24_000, rng_mm=1)
out = {}
for fused in (True, False):
    le = sb.Learner(device=local)
    le.set_fused(fused)
    le.init(1)
    le.set_norm(mn, mx)
    le.dp_connect_dist(dist)
    dist.barrier()
    le.replay_fused_dp(mem, rng_rpl=100 + rank, n_updates=50)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    le.replay_fused_dp(mem, rng_rpl=1000 + rank, n_updates=n_updates)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w0 = torch.from_numpy(le.get_layer(0, 1)[0]).cuda()
    ref = w0.clone()
    dist.broadcast(ref, 0)
    ok = torch.tensor([float(torch.equal(w0, ref))], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if TRACE:
        import ctypes
        buf = (ctypes.c_ulonglong * 16)()
        sb._lib.lib().ddpg_dp_trace_read.argtypes = [ctypes.c_void_p]
        sb._lib.lib().ddpg_dp_trace_read(buf)
        for seg, name in ((0, "critic"), (1, "actor")):
            st = list(buf[seg * 8: seg * 8 + 8])
            print("rank %d %s segment, (ns from kernel start): block 0's sums pushed %d, all blocks in %d, flags released %d, "
                  "all peers' flags seen %d, gather+ADAM done %d, grid done %d" % (rank, name, st[1] - st[0], st[2] - st[0], st[3] - st[0], st[4] - st[0], st[5] - st[0], st[6] - st[0]),
                  file=sys.stderr, flush=True)
    out["cluster_fused" if fused else "tiled_gemm"] = dict(us_per_update=1e3 * float(t.item()) / n_updates, status=le.dp_status(),
                                                          replicas_bit_identical=bool(ok.item() == 1.0))
    le.close()
if rank == 0:
    print(json.dumps(dict(world=world, batch_per_rank=120, **out)))
dist.barrier()
dist.destroy_process_group()
