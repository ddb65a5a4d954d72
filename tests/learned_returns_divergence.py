"""How fast do two correct implementations of the reference's training loop drift apart?  (not collected by pytest; run on a GPU box:
    python tests/learned_returns_divergence.py [episodes]  ->  the table in profiles/r2_learned_returns.md)

Three learners train side by side at the reference's shape (ONE instance, EP_LENGTH = 72, B = 120, 250/500, one replay() per step,
DDPG.jl:186-242) from identical initial weights, with identical start rows, injected Gaussian noise and minibatch indices:
    fused   CUDA, cluster-fused replay() (the default at this shape)
    tiled   CUDA, the tiled-GEMM replay() (ddpg_set_fused(0)): the same arithmetic in another fp32 summation order
    oracle  the CPU restatement (Float64 accumulation)
After every episode each policy is evaluated without noise on the first 72 rows of the evaluation series (DDPG.jl:266-279).  The point:
fused-vs-tiled (two orderings of the same CUDA arithmetic) drifts at the same rate as fused-vs-oracle, i.e. the drift past the first few
hundred updates is the loop's sensitivity to rounding (ADAM's normalised steps in a closed data-collection loop), not a discrepancy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402
from oracle import oracle as O  # noqa: E402

EPISODES = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T, B = 72, 120
O.build()
O.set_threads(max(1, min(16, (os.cpu_count() or 2) // 2)))
train = sb.series.synth_charger98(4320, seed=98)
evals = np.load(os.path.join(ROOT, "tests", "golden", "charger98_test_series.npz"))["series"]
P = O.params_for_charger(98)


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


rng = np.random.default_rng(2024)
wenv = sb.Shems(T, train, n_envs=64)
warm = sb.Replay(64 * T)
wenv.reset(rng=12)
wenv.rollout(sb.POLICY_RANDOM, T, seed=12, replay=warm, want_return=False)
S0, A0, R0, S20, D0 = warm.get()
mn, mx = warm.min_max_buffer(len(warm), rng_mm=1)
cap = S0.shape[1] + EPISODES * T


class Cuda:
    def __init__(self, fused):
        self.le = sb.Learner()
        assert self.le.set_fused(fused) is fused
        self.le.init(1231)
        self.le.set_norm(mn, mx)
        self.mem = sb.Replay(cap)
        self.mem.push(dev(S0), dev(A0), dev(R0), dev(S20), dev(D0))
        self.env, self.ev = sb.Shems(T, train, n_envs=1), sb.Shems(T, evals, n_envs=1)
        self.r64 = torch.empty(1, dtype=torch.float64, device="cuda")

    def reset(self, idx0, socb0):
        self.env.reset(idx0=idx0, socb0=socb0)

    def step(self, noise, idx):
        s = self.env.state_tensor().clone()
        a, scaled = self.le.act(s, noise=dev(noise))
        r, s2 = self.env.step(scaled, reward64_out=self.r64)
        self.mem.push(s, a, r, s2)
        self.le.replay(self.mem, n_updates=1, idx=idx)
        return float(self.r64[0])

    def score(self):
        self.ev.reset(rng=-1)
        return float(self.le.rollout(self.ev, T)["ep_return"][0])


class Oracle:
    def __init__(self, like):
        self.orc = O.OracleDdpg(O.default_ddpg_params())
        for net in range(4):
            for k in range(3):
                self.orc.set_layer(net, k, *like.le.get_layer(net, k))
        self.orc.set_norm(mn, mx)
        self.S, self.A, self.R, self.S2, self.D = S0, A0, R0, S20, D0
        self.env, self.ev = O.OracleEnv(P, train, T, 1), O.OracleEnv(P, evals, T, 1)

    def reset(self, idx0, socb0):
        self.env.reset(mode=1, idx0=idx0, socb0=socb0)

    def step(self, noise, idx):
        s = self.env.obs.copy()
        a, scaled = self.orc.act(s, noise=noise)
        r, s2, _ = self.env.step(scaled)
        self.S = np.concatenate([self.S, s], 1); self.A = np.concatenate([self.A, a], 1); self.R = np.concatenate([self.R, r.astype(np.float32)])
        self.S2 = np.concatenate([self.S2, s2], 1); self.D = np.concatenate([self.D, np.zeros(1, np.float32)])
        self.orc.update_batch(self.S[:, idx], self.A[:, idx], self.R[idx], self.S2[:, idx], self.D[idx])
        return float(r[0])

    def score(self):
        self.ev.reset(mode=0)
        tot = 0.0
        for _ in range(T):
            _, scaled = self.orc.act(self.ev.obs.copy())
            r, _, _ = self.ev.step(scaled)
            tot += float(r[0])
        return tot


fused, tiled = Cuda(True), Cuda(False)
runs = {"fused": fused, "tiled": tiled, "oracle": Oracle(fused)}
rel = lambda x, y: abs(x - y) / (abs(y) + 1.0)
print("| episode | eval score fused / tiled / oracle | fused vs oracle | tiled vs oracle | fused vs tiled | train return fused vs oracle |")
print("|---|---|---|---|---|---|")
n_mem = S0.shape[1]
for ep in range(EPISODES):
    idx0 = rng.integers(1, train.shape[1] - T, 1).astype(np.int32)
    socb0 = rng.uniform(0, 6.75, 1).astype(np.float32)
    for r in runs.values():
        r.reset(idx0, socb0)
    ret = dict.fromkeys(runs, 0.0)
    for step in range(T):
        noise = rng.normal(0, 0.1, (2, 1)).astype(np.float32)
        n_mem += 1
        idx = rng.integers(0, n_mem, B).astype(np.int32)
        for k, r in runs.items():
            ret[k] += r.step(noise, idx)
    sc = {k: r.score() for k, r in runs.items()}
    print("| %d | %.5f / %.5f / %.5f | %.1e | %.1e | %.1e | %.1e |" % (ep + 1, sc["fused"], sc["tiled"], sc["oracle"], rel(sc["fused"], sc["oracle"]),
          rel(sc["tiled"], sc["oracle"]), rel(sc["fused"], sc["tiled"]), rel(ret["fused"], ret["oracle"])), flush=True)
