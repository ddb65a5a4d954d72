"""CPU tests of the shems_LU1 oracle: hand-derived KATs, invariants from the source, golden regression."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_states

NAMES = "index c_ev EV_target EV Soc_ev rewards profit discomfort penalty PV_DE B_DE GR_DE PV_B PV_GR PV_EV B_EV GR_EV EX_EV GR_B B_GR B B_tar Soc_b".split()
T = {n: i for i, n in enumerate(NAMES)}


def test_params_charger98(O, P98):
    assert P98.b_soc_max == np.float32(7.5) * np.float32(0.9) == np.float32(6.75)
    assert P98.ev_soc_max == np.float32(35.816) and P98.b_rate_max == 3.3
    assert P98.sell_discount == float(np.float32(0.2)) and P98.discomfort_weight_ev == float(np.float32(0.01))
    with pytest.raises(KeyError):
        O.params_for_charger(42)
    assert O.params_for_charger(4).b_rate_max == 4.6


def test_kats_appendix_b(O, P98):
    """SURVEY.md Appendix B (hand-derived by the surveyor) — all agree to 1e-6 rel except K3's Soc_b' (see fixture note)."""
    kats = json.load(open(os.path.join(GOLDEN, "kat_appendix_b.json")))
    for k in kats:
        st = np.array(k["state"] + [1, 0, 1], np.float32)
        ser3 = np.array(k["series3"], np.float32)
        r, s2, i2, tr = O.step_single(P98, ser3, st, 1, np.array(k["action_used"], np.float32), k["track"])
        sv = k["survey"]
        got = dict(B=tr[T["B"]], EV=tr[T["EV"]], Soc_b=s2[0], Soc_ev=s2[1], reward=r, discomfort=tr[T["discomfort"]], penalty=tr[T["penalty"]])
        for n in ("PV_DE", "B_DE", "GR_DE", "PV_B", "PV_GR", "PV_EV", "B_EV", "GR_EV", "EX_EV"):
            got[n] = tr[T[n]]
        for key, want in sv.items():
            tol = 2e-7 if (k["name"], key) == ("K3", "Soc_b") else 1e-7
            assert got[key] == pytest.approx(want, rel=3e-7, abs=tol), (k["name"], key, got[key], want)
        # regression: bit-exact against the oracle values stored with the fixture
        assert r == k["oracle"]["reward"] and list(map(float, s2)) == k["oracle"]["state2"]
    k3 = [k for k in kats if k["name"] == "K3"][0]
    assert k3["oracle"]["Soc_b"] == float(np.float32(0.42103994))


def _trace_batch(O, P, series, obs, idx, act, track):
    env = O.OracleEnv(P, series, 1, obs.shape[1])
    env.obs[:] = obs
    env.idx[:] = idx
    r, s2, tr = env.step(act, track=track, want_trace=True)
    return r, s2.copy(), tr, env.idx.copy()


@pytest.mark.parametrize("track", [0.0, -0.5])
def test_energy_balance_invariants(O, P98, train_series, track):
    """Invariants that follow from the source: every kWh of demand / EV charge is served from exactly one of PV, battery, grid."""
    rng = np.random.default_rng(7)
    n = 20000
    obs, idx = random_states(rng, n, train_series, P98)
    if track < 0:
        env = O.OracleEnv(P98, train_series, 1, n)
        env.obs[:] = obs
        act = env.action()
    else:
        act = rng.uniform(0, 1, (2, n)).astype(np.float32)
        act[:, : n // 10] = rng.choice([0.0, 0.99, 1.0, 0.98999995], (2, n // 10))
    r, s2, tr, idx2 = _trace_batch(O, P98, train_series, obs, idx, act, track)
    d_e, g_e, EV = obs[3].astype(np.float64), obs[4].astype(np.float64), tr[T["EV"]]
    np.testing.assert_allclose(tr[T["PV_DE"]] + tr[T["B_DE"]] + tr[T["GR_DE"]], d_e, rtol=0, atol=2e-6)
    np.testing.assert_allclose(tr[T["PV_EV"]] + tr[T["B_EV"]] + tr[T["GR_EV"]], EV, rtol=0, atol=2e-6)
    assert np.all(tr[T["GR_B"]] == 0) and np.all(tr[T["B_GR"]] == 0)
    # PV is split into demand, EV, battery (PV_B stored; PV_B or PV_B/eta consumed) and export
    pv_used = tr[T["PV_DE"]] + tr[T["PV_EV"]] + tr[T["PV_GR"]]
    assert np.all(pv_used <= g_e + 1e-5) and np.all(pv_used + tr[T["PV_B"]] / 0.95 + 1e-5 >= g_e - 1e-5)
    assert np.all(s2[0] >= 0) and np.all(s2[0] <= P98.b_soc_max * (1 + 1e-6))
    assert np.all(s2[1] >= 0) and np.all(s2[1] <= 1.0 + 1e-6)
    assert np.all((r <= 1e-12) | (tr[T["PV_GR"]] > 0))  # reward > 0 only when exporting
    assert np.all(idx2 == idx + 1) and np.all(tr[T["index"]] == idx + 1)
    np.testing.assert_array_equal(s2[2:], train_series[1:, idx2 - 1][[0, 1, 2, 3, 4, 5, 6]])
    if track < 0:
        assert np.all(tr[T["penalty"]] == 0) and np.all(tr[T["EV_target"]] == 0) and np.all(tr[T["B_tar"]] == 0)
    else:  # penalty only when absent and EV_target < 0.99 (Float64 literal)
        pen_expected = (obs[2] < 0) & (act[1].astype(np.float64) < 0.99)
        assert np.array_equal(tr[T["penalty"]] > 0, pen_expected & (act[1] < 1))


def test_float64_threshold_equivalences():
    """The CUDA kernels evaluate `B < -0.01`, `B > 0.01`, `EV_target < 0.99` (Float64 literals in Julia,
    shems_LU1.jl:362, :412, :447) as Float32 compares against -0.01f0 / 0.01f0 / 0.99f0: check they agree
    on the Float32 neighbours of the thresholds."""
    def neigh(f):
        lo, hi = f, f
        out = [f]
        for _ in range(3):
            lo = np.nextafter(lo, np.float32(-9), dtype=np.float32)
            hi = np.nextafter(hi, np.float32(9), dtype=np.float32)
            out += [lo, hi]
        return out
    for x in neigh(np.float32(-0.01)):
        assert (float(x) < -0.01) == bool(x < np.float32(-0.01))
    for x in neigh(np.float32(0.01)):
        assert (float(x) > 0.01) == bool(x > np.float32(0.01))
    for x in neigh(np.float32(0.99)):
        assert (float(x) < 0.99) == bool(x < np.float32(0.99))


def test_arrival_and_departure_rules(O, P98):
    ser = np.zeros((8, 4), np.float32)
    ser[0] = [1, 0.3, 0.9, 1]      # soc_ev data
    ser[1] = [-1, 2, 1, 0]         # absent, then a session arrives at row 2
    ser[2], ser[3], ser[4], ser[5], ser[7] = 1.0, 0.0, 0.4, 1, 1
    st = np.array([2.0, 1.0, -1, 1.0, 0.0, 0.4, 1, 0, 1], np.float32)
    r, s2, i2, tr = O.step_single(P98, ser, st, 1, [0.5, 1.0], 0)
    assert s2[1] == np.float32(0.3) and s2[2] == 2  # newly connected: SOC loaded from the data (:270-272)
    r, s3, i3, tr = O.step_single(P98, ser, s2, i2, [0.5, 0.0], 0)
    assert s3[1] == np.float32(0.3) and s3[2] == 1  # already connected: data column ignored
    r, s4, i4, tr = O.step_single(P98, ser, s3, i3, [0.5, 0.0], 0)
    assert s4[2] == 0 and tr[T["discomfort"]] == 0  # departure is checked at c_ev == 0 only
    ser2 = np.concatenate([ser, ser[:, -1:]], axis=1)
    ser2[1, -1] = -1
    r, s5, i5, tr = O.step_single(P98, ser2, s4, i4, [0.5, 0.0], 0)
    assert tr[T["discomfort"]] == pytest.approx(70.0, rel=1e-6) and s5[1] == 1.0  # (1-0.3)*100, SOC forced to 1 (:442-446)
    assert r == pytest.approx(tr[T["profit"]] - 0.01 * 70.0**2, rel=1e-6)


def test_step_bounds_error(O, P98, train_series):
    st = np.zeros(9, np.float32)
    with pytest.raises(IndexError):
        O.step_single(P98, train_series, st, train_series.shape[1], [0.5, 0.5], 0)


def test_reset_modes_and_window_shift(O, P98, train_series):
    nrows, T_ = train_series.shape[1], 72
    env = O.OracleEnv(P98, train_series, T_, 1)
    env.reset(mode=0)
    assert env.idx[0] == 1 and env.obs[0, 0] == np.float32(3.375)
    np.testing.assert_array_equal(env.obs[1:, 0], train_series[:, 0])
    cd = train_series[1]
    hi = nrows - T_
    n = hi
    envs = O.OracleEnv(P98, train_series, T_, n)
    idx0 = np.arange(1, hi + 1, dtype=np.int32)
    envs.reset(mode=1, idx0=idx0, socb0=np.full(n, 1.5, np.float32))
    # python restatement of the shift loop (:227-246)
    for i0, got in zip(idx0, envs.idx):
        idx, c, cnt = int(i0), cd[i0 + T_ - 1], 0
        while c > -1 and idx < hi:
            idx += int(c + 1)
            if idx > hi:
                idx = int(i0)
            c = cd[idx + T_ - 1]
            cnt += 1
            if cnt > 100:
                break
        assert got == idx
    ends_free = cd[envs.idx + T_ - 1] == -1
    assert ends_free.mean() > 0.9  # almost every window ends outside a charging session
    assert np.all(envs.obs[0] == 1.5)
    with pytest.raises(ValueError):
        O.OracleEnv(P98, train_series[:, :72], 72, 1).reset(mode=0)  # rand(1:0) throws


def test_philox_reset_is_seeded_and_rank_invariant(O, P98, train_series):
    a = O.OracleEnv(P98, train_series, 72, 64)
    a.reset(mode=2, seed=5)
    b = O.OracleEnv(P98, train_series, 72, 32)
    b.reset(mode=2, seed=5, env_id_base=32)
    np.testing.assert_array_equal(a.obs[:, 32:], b.obs)
    np.testing.assert_array_equal(a.idx[32:], b.idx)
    c = O.OracleEnv(P98, train_series, 72, 64)
    c.reset(mode=2, seed=6)
    assert not np.array_equal(a.idx, c.idx)
    assert a.obs[0].min() >= 0 and a.obs[0].max() <= P98.b_soc_max and a.idx.min() >= 1 and a.idx.max() <= 4320 - 72


def test_rule_based_golden_charger98(O, P98, charger98_test_series):
    """Regression pin on the real Charger98 test series (reconstructed from the MPC benchmark CSV) + sanity vs the MPC optimum."""
    g = np.load(os.path.join(GOLDEN, "oracle_rule_based_charger98.npz"))
    env = O.OracleEnv(P98, charger98_test_series, 2998, 1)
    env.reset(mode=0)
    out = env.rollout(0, 2998, want_trace=True)
    tr = out["trace"][:, :, 0]
    np.testing.assert_array_equal(tr[:5], g["first"])
    np.testing.assert_array_equal(tr[-5:], g["last"])
    np.testing.assert_allclose(tr.sum(0), g["colsum"], rtol=1e-12)
    assert out["ep_return"][0] == g["ep_return"][0]
    # the perfect-foresight LP of the reference's Python benchmark reaches -369.537 EUR on this series: an upper bound
    assert tr[:, T["profit"]].sum() < -369.537
    assert np.all(tr[:, T["penalty"]] == 0)


def test_rollout_equals_step_loop(O, P98, train_series):
    n, T_ = 257, 24
    a = O.OracleEnv(P98, train_series, 72, n)
    a.reset(mode=2, seed=11)
    b = O.OracleEnv(P98, train_series, 72, n)
    b.reset(mode=2, seed=11)
    out = a.rollout(1, T_, seed=3, want_transitions=True)
    ret = np.zeros(n)
    for t in range(T_):
        raw = np.zeros((2, n), np.float32)
        for e in range(n):
            tmp = np.zeros(2, np.float32)
            O.lib().oracle_random_action(3, e, t, O._fp(tmp))
            raw[:, e] = tmp
        np.testing.assert_array_equal(raw, out["a"][t])
        assert raw.min() >= -1 and raw.max() <= 1
        scaled = ((raw.astype(np.float64) + 1.0) * 0.5).astype(np.float32)
        s_before = b.obs.copy()
        r, s2, _ = b.step(scaled, track=0)
        ret += r
        np.testing.assert_array_equal(out["s"][t], s_before)
        np.testing.assert_array_equal(out["s2"][t], s2)
        np.testing.assert_array_equal(out["r"][t], r.astype(np.float32))
    np.testing.assert_array_equal(out["ep_return"], ret)


def test_sibling_env_constants_and_reward_terms(O, train_series):
    """f4: module constants of shems_LU7.jl / shems_LU1_input0607.jl and their two type-level differences to shems_LU1
    (hand-checked against the source: shems_LU7.jl:25, :35, :91, :94, :465-468; shems_LU1_input0607.jl:38-52, :481-484)."""
    lu7 = O.params_for_env(1, 98)
    assert lu7.b_soc_max == 10.0 and lu7.b_rate_max == float(np.float32(4.6)) and lu7.sell_discount == float(np.float32(0.3))
    assert lu7.discomfort_weight_ev == 1.0 and lu7.disc_pot == 1.0 and lu7.penalty_in_f64 == 1 and lu7.penalty_weight_f64 == 0.1
    i67 = O.params_for_env(2, 98)
    assert i67.b_soc_max == np.float32(7.5) * np.float32(0.9) and i67.reward_form == 1 and i67.disc_pot == 1.0
    assert i67.discomfort_weight_ev == float(np.float32(0.1)) and i67.penalty_weight_f64 == 0.1
    lu1 = O.params_for_env(0, 98)
    assert lu1.penalty_in_f64 == 0 and lu1.reward_form == 0 and lu1.disc_pot == 2.0
    with pytest.raises(KeyError):
        O.params_for_env(1, 97)
    # one absent-EV step with EV_target = 0.5: penalty = (1 - 0.5f0) * w.  LU1: Float32 product 0.05f0; siblings: Float64 0.5 * 0.1
    ser = train_series
    row = int(np.flatnonzero(ser[1] < 0)[5]) + 1
    st = np.concatenate([[2.0], ser[:, row - 1]]).astype(np.float32)
    a = np.array([0.3, 0.5], np.float32)
    tr1 = O.step_single(lu1, ser, st, row, a, track=1)[3]
    tr7 = O.step_single(lu7, ser, st, row, a, track=1)[3]
    assert tr1[8] == float(np.float32(0.5) * np.float32(0.1)) and tr7[8] == 0.5 * 0.1
    # departure with an unfinished charge: discomfort d = (1 - Soc_ev') * 100; LU1 subtracts w*d^2, LU7 d*1, input0607 (d*w)^pot
    row0 = int(np.flatnonzero(ser[1] == 0)[0]) + 1
    st0 = np.concatenate([[0.0], ser[:, row0 - 1]]).astype(np.float32)
    st0[1] = 0.5
    a0 = np.array([0.0, 0.0], np.float32)
    out = {}
    for name, Pv in (("lu1", lu1), ("lu7", lu7), ("i67", i67)):
        tr = O.step_single(Pv, ser, st0, row0, a0, track=1)[3]
        out[name] = tr
        assert tr[7] == 50.0                                     # (1 - 0.5f0) * 100
    assert out["lu1"][5] == out["lu1"][6] - float(np.float32(0.01)) * 50.0 ** 2
    assert out["lu7"][5] == out["lu7"][6] - 50.0 * 1.0
    assert out["i67"][5] == out["i67"][6] - 50.0 * float(np.float32(0.1))
    i67.disc_pot = 2.0
    tr = O.step_single(i67, ser, st0, row0, a0, track=1)[3]
    assert tr[5] == tr[6] - (50.0 * float(np.float32(0.1))) ** 2.0
