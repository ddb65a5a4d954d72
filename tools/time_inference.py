"""Times evaluation / DRL inference (episode!(train = false), inference(track = 1)): ddpg_rollout — actor + step! for all steps in ONE
persistent cluster kernel — against the step-by-step loop (one fused act launch + one step! launch per step) through the same ABI.
usage: python tools/time_inference.py   -> JSON lines"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

le = sb.Learner()
le.init(7)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, nrows, T, n, trace in (("inference(track=1), test set, 1 instance", 3000, 2999, 1, True),
                                 ("inference(track=1), eval set, 1 instance", 1440, 1439, 1, True),
                                 ("run_episodes evaluation: 100 episodes x 72 steps as 100 instances", 1440, 72, 100, False),
                                 ("one evaluation episode, 72 steps, 1 instance", 1440, 72, 1, False)):
    ser = sb.series.synth_charger98(nrows, seed=7)
    env = sb.Shems(nrows - 1, ser, n_envs=n)

    def fused():
        env.reset(rng=-1)
        return le.rollout(env, T, want_trace=trace)

    def loop():
        env.reset(rng=-1)
        tr = []
        for t in range(T):
            a, scaled = le.act(env.state_tensor(), train=False)
            out = env.step(scaled, track=1 if trace else 0)
            if trace:
                tr.append(out[2])
        return tr

    ms_f, ms_l = timed(fused), timed(loop, reps=2)
    print(json.dumps(dict(case=name, steps=T, instances=n, one_kernel_ms=ms_f, us_per_step=1e3 * ms_f / T, step_loop_ms=ms_l,
                          step_loop_us_per_step=1e3 * ms_l / T, speedup=ms_l / ms_f)))
