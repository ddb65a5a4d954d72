"""Generates tests/golden/oracle_ddpg_small.npz (run in the build container only): a small, fixed DDPG problem — the library's
Philox initialisation, normalisation constants, four minibatches — and the ORACLE's results after each update (weights of all four
nets, losses, actions of `act`).  These are oracle outputs, not reference outputs (Julia/Flux cannot run here): the file pins the
oracle against accidental change and gives the CUDA path committed vectors to be compared with (tests/test_oracle_ddpg.py,
tests/test_replay_ddpg_gpu.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
from oracle import oracle as O  # noqa: E402

B, L1, L2, K, SEED = 24, 20, 28, 4, 2024


def run(O):
    rng = np.random.default_rng(SEED)
    orc = O.OracleDdpg(O.default_ddpg_params(batch=B, l1=L1, l2=L2))
    orc.init(SEED)
    s_min = rng.uniform(-1, 0, 9).astype(np.float32)
    s_max = (s_min + rng.uniform(0.5, 3, 9)).astype(np.float32)
    s_max[5] = s_min[5]                                    # the constant p_buy column
    orc.set_norm(s_min, s_max)
    out = dict(s_min=s_min, s_max=s_max)
    for u in range(K):
        s = rng.uniform(-1, 3, (9, B)).astype(np.float32)
        s[5] = s_min[5]
        s2 = rng.uniform(-1, 3, (9, B)).astype(np.float32)
        s2[5] = s_min[5]
        a = rng.uniform(-1, 1, (2, B)).astype(np.float32)
        r = rng.uniform(-5, 1, B).astype(np.float32)
        orc.update_batch(s, a, r, s2)
        out.update({f"s{u}": s, f"a{u}": a, f"r{u}": r, f"s2_{u}": s2, f"loss{u}": np.array(orc.losses(), np.float32)})
    for net in range(4):
        for k in range(3):
            w, b = orc.get_layer(net, k)
            out[f"w{net}{k}"], out[f"b{net}{k}"] = w, b
    obs = rng.uniform(-1, 3, (9, 16)).astype(np.float32)
    noise = rng.normal(0, 0.1, (2, 16)).astype(np.float32)
    act, scaled = orc.act(obs, noise=noise)
    out.update(obs=obs, noise=noise, act=act, scaled=scaled)
    return out


if __name__ == "__main__":
    O.build()
    np.savez_compressed(os.path.join(HERE, "oracle_ddpg_small.npz"), **run(O))
    print("written")
