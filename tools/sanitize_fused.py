"""Smallest program that drives the cluster-fused kernels (critic pass, actor pass, partial-sum ADAM, act) on three shapes and
through the graph path — written for `compute-sanitizer --tool memcheck python tools/sanitize_fused.py` (the tool is closed on
this pool in round 1, so it only served as a plain smoke run)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

rng = np.random.default_rng(0)
for B, l1, l2 in ((120, 250, 500), (16, 3, 5), (40, 255, 497)):
    le = sb.Learner(params=sb.default_ddpg_params(batch=B, l1=l1, l2=l2))
    assert le.set_fused(True)
    le.init(3)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    for _ in range(2):
        le.update_batch(dev(rng.uniform(-1, 3, (9, B)).astype(np.float32)), dev(rng.uniform(-1, 1, (2, B)).astype(np.float32)),
                        dev(rng.uniform(-5, 1, B).astype(np.float32)), dev(rng.uniform(-1, 3, (9, B)).astype(np.float32)))
    a, _ = le.act(dev(rng.uniform(0, 1, (9, 5)).astype(np.float32)), train=True, sigma=0.1, rng_act=1, step=1)
    torch.cuda.synchronize()
    print(B, l1, l2, le.losses(), a.cpu().numpy().ravel()[:2], flush=True)
    le.close()
# the graph path with on-device sampling
ser = sb.series.synth_charger98(4320, seed=98)
env = sb.Shems(72, ser, n_envs=64)
mem = sb.Replay(64 * 72)
env.reset(rng=1)
env.rollout(sb.POLICY_RANDOM, 72, seed=1, replay=mem, want_return=False)
le = sb.Learner()
le.init(1)
le.replay(mem, rng_rpl=1, n_updates=3)
torch.cuda.synchronize()
print("graph path", le.losses(), flush=True)
