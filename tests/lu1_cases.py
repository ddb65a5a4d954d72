"""Seeded (state, idx, action, track) cases for the shems_LU1 transition, built to reach EVERY leaf of the flow dispatch
(shems_LU1.jl:362-449).  Shared by tests/test_kat_leaves.py (numpy restatement vs oracle vs CUDA) and by
julia/make_golden.jl's input files (tests/golden/make_julia_inputs.py), so a Julia run of the UNMODIFIED reference pins the
same cases.  Test infrastructure; nothing here is imported by the product."""
import numpy as np

import lu1_numpy as J

F32 = np.float32


def make_cases(series, n, seed=2024, charger_id=98):
    """-> dict of arrays: state [n][9] float32, idx [n] int32 (1-based, idx+1 <= nrows), a [n][2] float32, track [n] float64"""
    K = J.Consts(charger_id)
    rng = np.random.default_rng(seed)
    nrows = series.shape[1]
    cd = series[1]
    arrivals = np.nonzero((cd[:-1] == -1) & (cd[1:] >= 0))[0] + 1      # 1-based idx whose next row is a newly connected EV
    assert len(arrivals) > 0
    state = np.zeros((n, 9), F32)
    idx = rng.integers(1, nrows, size=n).astype(np.int32)
    take = rng.random(n) < 0.15
    idx[take] = rng.choice(arrivals, take.sum())
    smax = float(K.b_soc_max)
    state[:, 0] = rng.uniform(0, smax, n)
    corners_b = np.array([0.0, 5e-4, 1e-3, 1.0000001e-3, smax, smax * 0.95, 0.011, 3.3, 3.4737, 0.0105, smax - 1e-3], F32)
    m = rng.random(n) < 0.2
    state[m, 0] = rng.choice(corners_b, m.sum())
    # exogenous fields: half from the series row (what a real run sees), half free (the transition only reads env.state)
    from_row = rng.random(n) < 0.5
    state[:, 1:] = series[:, idx - 1].T
    free = ~from_row
    k = free.sum()
    state[free, 2] = rng.choice(np.array([-1, -1, 0, 0, 1, 2, 5, 17, 40], F32), k)
    state[free, 3] = rng.choice([rng.uniform(0.05, 6.0), 0.25, 1.0], k) * rng.uniform(0.2, 1.5, k)
    pv_on = rng.random(k) < 0.6
    state[free, 4] = np.where(pv_on, rng.uniform(0, 22, k), 0.0)
    state[free, 5] = np.where(rng.random(k) < 0.7, 0.4, rng.uniform(0.05, 0.6, k))
    connected = state[:, 2] >= 0
    state[:, 1] = np.where(connected, rng.uniform(0, 1, n), 1.0)
    m = connected & (rng.random(n) < 0.2)
    state[m, 1] = rng.choice(np.array([0.0, 0.5, 0.98999995, 0.99, 0.99999994, 1.0], F32), m.sum())
    # small residuals so that the battery can cover demand AND EV (leaves A2a / B1a need BD > need)
    small = rng.random(n) < 0.25
    state[small, 3] = rng.uniform(0.01, 0.8, small.sum())
    state[small & connected, 1] = rng.uniform(0.97, 1.0, (small & connected).sum())
    a = rng.uniform(0, 1, (n, 2)).astype(F32)
    m = rng.random(n) < 0.15
    a[m] = rng.choice(np.array([0.0, 0.99, 0.98999995, 0.9900001, 1.0, 0.5], F32), (m.sum(), 2))
    track = rng.choice([0.0, 1.0, -0.5], n, p=[0.5, 0.15, 0.35])
    for i in np.nonzero(track < 0)[0]:
        if rng.random() < 0.7:   # what episode! passes: a = action(env, track)  (DDPG.jl:209-212)
            a[i] = J.action_rule(K, [F32(v) for v in state[i]])
        else:                    # any feasible-looking (B, EV) pair: step! takes them as given (:350-353)
            a[i] = (F32(rng.uniform(-3.5, 3.5)), F32(rng.choice([0.0, rng.uniform(0, 11), 0.3])))
    return dict(state=state, idx=idx, a=a, track=track)


# every (flow leaf, charging leaf) pair the source can reach, and why the others cannot be reached:
#   A2a, B1a, B1b need a discharge budget BD > 0 (:362) => B < -0.01 => no charging (:412 needs B > 0.01);
#   A2b and B2 set pv_ = 0 (:376, :390) so `pv_ > BC/eta` (:414) is false => only the c2 leaf.
REACHABLE_FLOW_CHARGE = {
    ("A1", "none"), ("A1", "c1"), ("A1", "c2"), ("A2a", "none"), ("A2b", "none"), ("A2b", "c2"),
    ("B1a", "none"), ("B1b", "none"), ("B2", "none"), ("B2", "c2"),
}
TAILS = {"none", "departure", "penalty"}
