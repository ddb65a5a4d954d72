"""End-to-end population training (BASELINE configs[4] shape): P learners = chargers x seeds on one GPU.
usage: python tools/train_population.py [--learners 80] [--envs 64] [--episodes 20]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--learners", type=int, default=80)
ap.add_argument("--envs", type=int, default=64)
ap.add_argument("--episodes", type=int, default=20)
ap.add_argument("--tc", type=int, default=1)
ap.add_argument("--eval-every", type=int, default=0)
args = ap.parse_args()
CH = (1, 2, 3, 4, 5, 6, 7, 8, 9, 98)
P = args.learners
ser = sb.series.synth_charger98(4320, seed=98)
drv = sb.PopulationDriver(ser, chargers=[CH[(g // 64) % 10] for g in range(P)], seeds=[1231 + g for g in range(P)], n_envs=args.envs,
                          use_tensor_cores=args.tc)
t0 = time.time()
drv.populate_memory()
drv.min_max_buffer()
torch.cuda.synchronize()
t_pop = time.time() - t0
rule = drv.evaluate_rule_based()
first = drv.episode(train=False, rng_ep=999).cpu().numpy()
t1 = time.time()
t_eval = 0.0
for ep in range(1, args.episodes + 1):
    r = drv.episode(train=True, rng_ep=ep)
    if args.eval_every and ep % args.eval_every == 0:
        torch.cuda.synchronize()
        te = time.time()
        ev = drv.episode(train=False, rng_ep=999).cpu().numpy()
        t_eval += time.time() - te
        lc = [drv.learner.select(l).losses()[0] for l in (0, P - 1)]
        print(json.dumps(dict(episode=ep, train_return=float(r.mean()), eval_return_mean=float(ev.mean()), eval_return_best=float(ev.max()),
                              eval_return_worst=float(ev.min()), beats_rule_based=int((ev > rule).sum()), loss_crit_first_last=lc)), flush=True)
torch.cuda.synchronize()
dt = time.time() - t1 - t_eval
last = drv.episode(train=False, rng_ep=999).cpu().numpy()
print(json.dumps(dict(learners=P, envs_per_learner=args.envs, episodes=args.episodes, populate_seconds=t_pop, train_seconds=dt,
                      vector_steps_per_s=args.episodes * 72 / dt, learner_updates_per_s=P * args.episodes * 72 / dt,
                      env_steps_per_s=P * args.envs * args.episodes * 72 / dt, eval_return_before=float(first.mean()),
                      eval_return_after=float(last.mean()), rule_based=float(rule.mean()))))
