"""Times the step-API kernel (`step!` for N instances, one launch per vector step): 64 algorithmic bytes per env-step.
usage: python tools/time_step.py [n_envs] [mode]   mode: lockstep (all instances on the same row) | random (random start rows)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
mode = sys.argv[2] if len(sys.argv) > 2 else "random"
steps = 60
ser = sb.series.synth_charger98(4320, seed=98)
env = sb.Shems(72, ser, n_envs=n)
env.reset(rng=-1 if mode == "lockstep" else 3)
act = torch.rand((2, n), device="cuda")
rew = torch.empty(n, device="cuda")
for _ in range(5):
    env.step(act, reward_out=rew)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    env.step(act, reward_out=rew)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps(dict(n=n, mode=mode, us_per_step=ms * 1e3, env_steps_per_s=n / ms * 1e3, gbs_64=64 * n / ms / 1e6,
                      working_set_mb=64 * n / 1e6)))
