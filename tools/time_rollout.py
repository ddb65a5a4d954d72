"""Times the fused rollout kernel of one library build (SHEMS_B200_LIB) on the config-3 workload shape.
usage: python tools/time_rollout.py [n_envs] [T] [policy] [nrows]   -> one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
policy = int(sys.argv[3]) if len(sys.argv) > 3 else sb.POLICY_RANDOM
nrows = int(sys.argv[4]) if len(sys.argv) > 4 else T + 1
sink = sys.argv[5] if len(sys.argv) > 5 else "replay"
ser = sb.series.synth_charger98(nrows, seed=98)
env = sb.Shems(T, ser, n_envs=n)
mem = sb.Replay(n * 16) if sink == "replay" else None
ms = []
for it in range(6):
    env.reset(rng=it + 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.rollout(policy, T, seed=it + 1, replay=mem, want_return=True)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
best = sorted(ms[2:])[len(ms[2:]) // 2]
print(json.dumps(dict(lib=os.path.basename(os.environ.get("SHEMS_B200_LIB", "default")), n=n, T=T, policy=policy, sink=sink, ms=best,
                      steps_per_s=n * T / best * 1e3, gbs_88=88 * n * T / best / 1e6)))
