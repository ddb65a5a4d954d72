"""The replay() problem of julia/make_golden.jl, generated identically on both sides (test infrastructure).

Weights, biases and the replay memory come from a counter-based generator (splitmix64 of seed + index -> a Float64 uniform in
[0, 1) with 53 random bits) that is three lines in Julia and vectorises in numpy, so no multi-megabyte weight file has to
travel: the Julia script rebuilds the same problem, runs the UNMODIFIED algorithms/DDPG.jl `replay()` on it and writes digests
of the four nets after each update (tests/golden/reference_julia/ddpg_replay.csv), which tests/test_julia_golden.py compares
with the oracle (CPU) and the CUDA update (GPU) whenever that file is present.
"""
import numpy as np

S, A, L1, L2, B, N_MEM, N_UPDATES = 9, 2, 250, 500, 120, 512, 3
MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def uniform(seed, n):
    """u[i] = (splitmix64(seed * 2^32 + i) >> 11) * 2^-53, i = 0..n-1"""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) << np.uint64(32)) + np.arange(n, dtype=np.uint64)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def layer_dims(net):
    """(in, out) of the three Dense layers; net: 0 actor, 1 critic, 2 actor_target, 3 critic_target"""
    nin, nout = (S + A, 1) if net in (1, 3) else (S, A)
    return [(nin, L1), (L1, L2), (L2, nout)]


def tensor_seed(net, layer, is_bias):
    return 1000 + 10 * net + 2 * layer + (1 if is_bias else 0)


def weights():
    """{(net, layer): (W flat in Flux order (out x in column-major), b)}: hidden layers glorot-uniform-like, last layers
    U(-3e-3, 3e-3) (DDPG.jl:21-22), small non-zero biases; the targets are independent draws so that a mix-up shows"""
    out = {}
    for net in range(4):
        for k, (i, o) in enumerate(layer_dims(net)):
            u = uniform(tensor_seed(net, k, False), i * o)
            w = (u - 0.5) * np.sqrt(24.0 / (i + o)) if k < 2 else 6e-3 * u - 3e-3
            b = (uniform(tensor_seed(net, k, True), o) - 0.5) * 0.02
            out[(net, k)] = (w.astype(np.float32), b.astype(np.float32))
    return out


def memory():
    """N_MEM transitions: s, s2 [9][n] float32, a [2][n] float32, r [n] float64 (the reference's memory holds Float64 rewards)"""
    def states(seed0):
        u = [uniform(seed0 + k, N_MEM) for k in range(9)]
        s = np.stack([u[0] * 6.75, u[1], np.floor(u[2] * 42.0) - 1.0, 0.2 + u[3] * 5.8, u[4] * 20.0, np.full(N_MEM, 0.4),
                      2.0 * u[6] - 1.0, 2.0 * u[7] - 1.0, 1.0 + np.floor(u[8] * 4.0)])
        return s.astype(np.float32)
    s, s2 = states(2000), states(2100)
    a = np.stack([2.0 * uniform(2200 + k, N_MEM) - 1.0 for k in range(2)]).astype(np.float32)
    r = -(uniform(2300, N_MEM) * 5.0)
    return s, a, r, s2


def digest_positions(n):
    return np.arange(n) if n <= 64 else (np.arange(64) * (n // 64))


def digest(x):
    """[sum, sum|x|, x[digest_positions]] in Float64 — the columns of ddpg_replay.csv"""
    x = np.asarray(x, np.float32).ravel().astype(np.float64)
    return np.concatenate([[x.sum(), np.abs(x).sum()], x[digest_positions(len(x))]])
