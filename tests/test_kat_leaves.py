"""Leaf-complete known-answer tests of the shems_LU1 transition.

Two restatements of RL-SHEMS/RL_environments/envs/shems_LU1.jl that share no code — tests/lu1_numpy.py (a statement-for-statement
numpy transliteration whose typing is numpy's own scalar promotion) and oracle/shems_oracle.c (C, tagged Int/Float32/Float64
values) — must agree BIT FOR BIT on seeded cases that reach every leaf of the flow dispatch (:362-449): the six flow leaves
A1 / A2a / A2b / B1a / B1b / B2 x the charging leaves none / c1 / c2 (where reachable) x departure / penalty / neither, with and
without a newly connected EV in the next row (:270-272).  The coverage itself is asserted.  The `-m gpu` half runs the CUDA
step kernel (through the C ABI) on the same cases against the numpy restatement — i.e. NOT against the oracle.
"""
import collections

import numpy as np
import pytest

import lu1_cases as Cs
import lu1_numpy as J

N_CASES = 20000


@pytest.fixture(scope="module")
def cases(charger98_test_series):
    return Cs.make_cases(charger98_test_series, N_CASES)


@pytest.fixture(scope="module")
def numpy_results(cases, charger98_test_series):
    K = J.Consts(98)
    out = []
    for i in range(N_CASES):
        out.append(J.step(K, charger98_test_series, cases["state"][i], int(cases["idx"][i]), cases["a"][i], cases["track"][i]))
    return out


def test_every_leaf_is_reached(numpy_results, cases):
    hit = collections.Counter()
    for (_, _, _, _, tags), track in zip(numpy_results, cases["track"]):
        hit[("fct", tags["flow"], tags["charge"], tags["tail"])] += 1
        hit[("flow_arrival", tags["flow"], tags["arrival"])] += 1
        hit[("tail_arrival", tags["tail"], tags["arrival"])] += 1
        hit[("flow_discharge", tags["flow"], tags["discharge"])] += 1
        hit[("action", tags["act_ev"], tags["act_b"])] += 1
        hit[("track", float(track))] += 1
    for fc in Cs.REACHABLE_FLOW_CHARGE:
        for tail in Cs.TAILS:
            assert hit[("fct",) + fc + (tail,)] >= 3, ("leaf not covered", fc, tail)
    seen_fc = {k[1:3] for k in hit if k[0] == "fct"}
    assert seen_fc == Cs.REACHABLE_FLOW_CHARGE, seen_fc ^ Cs.REACHABLE_FLOW_CHARGE   # and nothing the analysis calls unreachable
    for flow in ("A1", "A2a", "A2b", "B1a", "B1b", "B2"):
        assert hit[("flow_arrival", flow, True)] and hit[("flow_arrival", flow, False)]
    for tail in Cs.TAILS:
        assert hit[("tail_arrival", tail, True)] and hit[("tail_arrival", tail, False)]
    # leaves that need a discharge budget never run without one; the slack leaves run with and without
    for flow in ("A2a", "B1a", "B1b"):
        assert hit[("flow_discharge", flow, True)] and not hit[("flow_discharge", flow, False)]
    for flow in ("A1", "A2b", "B2"):
        assert hit[("flow_discharge", flow, True)] and hit[("flow_discharge", flow, False)]
    # action(env, a) :292-313: EV on/off x battery charge / discharge / idle; action(env, track) through `given`/rule cases
    for ev in ("ev_on", "ev_off"):
        for b in ("b_charge", "b_discharge", "b_idle"):
            assert hit[("action", ev, b)], (ev, b)
    assert hit[("action", "given", "given")] and hit[("track", 0.0)] and hit[("track", 1.0)] and hit[("track", -0.5)]


def test_oracle_equals_numpy_restatement_bit_for_bit(O, P98, cases, numpy_results, charger98_test_series):
    ser = charger98_test_series
    for i, (r, s2, i2, tr, tags) in enumerate(numpy_results):
        ro, so, io, tro = O.step_single(P98, ser, cases["state"][i], int(cases["idx"][i]), cases["a"][i], cases["track"][i])
        assert r == ro and io == i2, (i, tags, r, ro)
        assert s2.tobytes() == so.tobytes(), (i, tags, s2, so)
        assert tr.tobytes() == tro.tobytes() or np.array_equal(tr, tro), (i, tags, tr - tro)


def test_actions_bit_for_bit(O, P98, cases):
    K = J.Consts(98)
    st = cases["state"][:4000]
    n = len(st)
    ref = O.OracleEnv(P98, np.zeros((8, 2), np.float32), 1, n)
    ref.obs[:] = st.T
    rule = ref.action()
    drl = ref.action(cases["a"][:n].T.copy())
    for i in range(n):
        s = [np.float32(v) for v in st[i]]
        assert tuple(rule[:, i]) == J.action_rule(K, s), i
        assert tuple(drl[:, i]) == J.action_drl(K, s, (cases["a"][i, 0], cases["a"][i, 1])), i


@pytest.mark.parametrize("cid", [1, 4, 6, 97])
def test_other_chargers_bit_for_bit(O, charger98_test_series, cid):
    """other rows of `capacities` (:47-59): rate_max 4.6, other battery / EV sizes"""
    ser = charger98_test_series
    K, P = J.Consts(cid), O.params_for_charger(cid)
    assert float(K.b_soc_max) == P.b_soc_max and float(K.ev_soc_max) == P.ev_soc_max and float(K.b_rate_max) == P.b_rate_max
    cs = Cs.make_cases(ser, 3000, seed=cid, charger_id=cid)
    for i in range(3000):
        r, s2, i2, tr, _ = J.step(K, ser, cs["state"][i], int(cs["idx"][i]), cs["a"][i], cs["track"][i])
        ro, so, io, tro = O.step_single(P, ser, cs["state"][i], int(cs["idx"][i]), cs["a"][i], cs["track"][i])
        assert r == ro and s2.tobytes() == so.tobytes() and np.array_equal(tr, tro), i


@pytest.mark.parametrize("variant,vid,cid", [("LU7", 1, 98), ("LU7", 1, 4), ("INPUT0607", 2, 98), ("INPUT0607", 2, 6)])
def test_sibling_envs_bit_for_bit(O, charger98_test_series, variant, vid, cid):
    """shems_LU7.jl / shems_LU1_input0607.jl: other constants, a Float64 penalty weight and another reward line"""
    ser = charger98_test_series
    K, P = J.Consts(cid, variant), O.params_for_env(vid, cid)
    assert float(K.b_soc_max) == P.b_soc_max and float(K.ev_soc_max) == P.ev_soc_max and float(K.b_rate_max) == P.b_rate_max
    assert float(K.sell_discount) == P.sell_discount
    cs = Cs.make_cases(ser, 3000, seed=100 + cid, charger_id=98)
    cs["state"][:, 0] = np.minimum(cs["state"][:, 0], np.float32(K.b_soc_max))
    seen = set()
    for i in range(3000):
        r, s2, i2, tr, tags = J.step(K, ser, cs["state"][i], int(cs["idx"][i]), cs["a"][i], cs["track"][i])
        ro, so, io, tro = O.step_single(P, ser, cs["state"][i], int(cs["idx"][i]), cs["a"][i], cs["track"][i])
        assert r == ro and s2.tobytes() == so.tobytes() and np.array_equal(tr, tro), (i, tags, r, ro)
        seen.add(tags["tail"])
    assert seen == Cs.TAILS


def test_unknown_charger_is_a_key_error(O):
    with pytest.raises(KeyError):
        J.Consts(42)
    with pytest.raises(KeyError):
        O.params_for_charger(42)
    with pytest.raises(KeyError):
        J.Consts(97, "INPUT0607")
    with pytest.raises(KeyError):
        O.params_for_env(2, 97)


def test_reset_bit_for_bit(O, P98, charger98_test_series):
    """reset_state! :216-262 for every admissible start row (all paths of the window-shift loop) and rng == -1"""
    ser, T = charger98_test_series, 72
    K = J.Consts(98)
    n = ser.shape[1] - T
    idx0 = np.arange(1, n + 1, dtype=np.int32)
    socb0 = np.random.default_rng(3).uniform(0, 6.75, n).astype(np.float32)
    ref = O.OracleEnv(P98, ser, T, n)
    ref.reset(mode=1, idx0=idx0, socb0=socb0)
    for i in range(n):
        s, idx = J.reset_state(K, ser, T, False, idx0[i], socb0[i])
        assert idx == ref.idx[i] and s.tobytes() == ref.obs[:, i].tobytes(), i
    ref1 = O.OracleEnv(P98, ser, T, 1)
    ref1.reset(mode=0)
    s, idx = J.reset_state(K, ser, T, True)
    assert idx == 1 and s.tobytes() == ref1.obs[:, 0].tobytes() and s[0] == np.float32(3.375)


def test_closed_loop_episode_bit_for_bit(O, P98, charger98_test_series):
    """72 steps of both restatements feeding on their own outputs (rule-based and random targets)"""
    ser, T = charger98_test_series, 72
    K = J.Consts(98)
    rng = np.random.default_rng(11)
    for track in (-0.5, 0.0):
        s, idx = J.reset_state(K, ser, T, False, 1500, np.float32(2.2))
        so, io = s.copy(), idx
        ret = reto = 0.0
        for t in range(T):
            a = J.action_rule(K, [np.float32(v) for v in s]) if track < 0 else tuple(rng.uniform(0, 1, 2).astype(np.float32))
            r, s, idx, _, _ = J.step(K, ser, s, idx, a, track)
            ro, so, io, _ = O.step_single(P98, ser, so, io, a, track)
            ret, reto = ret + r, reto + ro
            assert s.tobytes() == so.tobytes() and r == ro and idx == io, (track, t)
        assert ret == reto


# ------------------------------------------------------------------------------------------------ CUDA (through the C ABI)
@pytest.mark.gpu
def test_cuda_step_equals_numpy_restatement(sb, cases, numpy_results, charger98_test_series):
    """The CUDA step kernel against the numpy transliteration (the oracle is not involved): Float32 states bit-exact, Float64
    reward and trace columns within 1e-12 relative (north-star tolerance: 1e-5 relative — asserted too)."""
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available()
    ser = charger98_test_series
    for track in (0.0, 1.0, -0.5):
        sel = np.nonzero(cases["track"] == track)[0]
        n = len(sel)
        env = sb.Shems(72, ser, n_envs=n)
        env.set_state(np.ascontiguousarray(cases["state"][sel].T), cases["idx"][sel])
        act = torch.as_tensor(np.ascontiguousarray(cases["a"][sel].T), device="cuda")
        if track == 0.0:
            r, s2 = env.step(act, track=0)
            tr = None
        else:
            r, s2, tr = env.step(act, track=track)
        s2 = s2.cpu().numpy()
        r = r.cpu().numpy()
        want_s = np.stack([numpy_results[i][1] for i in sel], axis=1)
        want_r = np.array([numpy_results[i][0] for i in sel])
        assert s2.tobytes() == want_s.tobytes(), np.nonzero((s2 != want_s).any(axis=0))[0][:10]
        np.testing.assert_array_equal(r, want_r.astype(np.float32))
        np.testing.assert_allclose(r, want_r, rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(env.idx, np.array([numpy_results[i][2] for i in sel]))
        if tr is not None:
            want_tr = np.stack([numpy_results[i][3] for i in sel], axis=1)
            np.testing.assert_allclose(tr.cpu().numpy(), want_tr, rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_cuda_actions_equal_numpy_restatement(sb, cases, charger98_test_series):
    torch = pytest.importorskip("torch")
    K = J.Consts(98)
    n = 4000
    env = sb.Shems(72, charger98_test_series, n_envs=n)
    env.set_state(np.ascontiguousarray(cases["state"][:n].T), cases["idx"][:n])
    rule = env.action(-0.5).cpu().numpy()
    drl = env.action(torch.as_tensor(np.ascontiguousarray(cases["a"][:n].T), device="cuda")).cpu().numpy()
    for i in range(n):
        s = [np.float32(v) for v in cases["state"][i]]
        assert tuple(rule[:, i]) == J.action_rule(K, s), i
        assert tuple(drl[:, i]) == J.action_drl(K, s, (cases["a"][i, 0], cases["a"][i, 1])), i
