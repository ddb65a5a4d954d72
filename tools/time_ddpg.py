"""Times ddpg_update (CUDA-graph path) at the reference's tuned config.  usage: python tools/time_ddpg.py [batch] [n_updates] [use_tensor_cores]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 120
n_updates = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
tc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ser = sb.series.synth_charger98(4320, seed=98)
env = sb.Shems(72, ser, n_envs=1000)
mem = sb.Replay(max(24_000, B))
env.reset(rng=1)
env.rollout(sb.POLICY_RANDOM, 24, seed=1, replay=mem, want_return=False)
le = sb.Learner(params=sb.default_ddpg_params(batch=B, use_tensor_cores=tc))
le.init(1)
mn, mx = mem.min_max_buffer(24_000, rng_mm=1)
le.set_norm(mn, mx)
le.replay(mem, rng_rpl=1, n_updates=50)
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    le.replay(mem, rng_rpl=2 + rep, n_updates=n_updates)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1))
ms = sorted(res)[1]
lc, la = le.losses()
flops = 10 * 256_500 * B
print(json.dumps(dict(batch=B, tensor_cores=tc, n_updates=n_updates, us_per_update=1e3 * ms / n_updates, updates_per_s=n_updates / ms * 1e3,
                      tflops=flops * n_updates / ms / 1e9, loss_crit=lc, loss_act=la)))
