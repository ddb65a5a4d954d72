"""GPU parity tests: the CUDA env (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (north_star): per-step states and rewards within 1e-5 relative (fp32).  Because the
kernels replicate Julia's Float32/Float64 promotions the states are expected BIT-EXACT and are
asserted so; rewards/traces are Float64 and asserted to 1e-12 relative.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_states

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


@pytest.fixture(scope="module")
def cuda_ok():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"


def test_kats_through_c_abi(sb, O, cuda_ok):
    kats = json.load(open(os.path.join(GOLDEN, "kat_appendix_b.json")))
    for k in kats:
        ser3 = np.array(k["series3"], np.float32)
        env = sb.Shems(1, ser3, n_envs=1)
        st = np.array(k["state"] + [1, 0, 1], np.float32).reshape(9, 1)
        env.set_state(st, np.array([1], np.int32))
        act = dev(np.array(k["action_used"], np.float32).reshape(2, 1))
        if k["track"] < 0:
            np.testing.assert_array_equal(env.action(-0.5).cpu().numpy().ravel(), np.array(k["action_used"], np.float32))
        r, s2, tr = env.step(act, track=(-0.5 if k["track"] < 0 else 1))
        assert s2.cpu().numpy().ravel().tolist() == k["oracle"]["state2"], k["name"]
        assert float(tr[5, 0]) == k["oracle"]["reward"], k["name"]
        names = list(k["oracle"]["trace"].keys())
        np.testing.assert_array_equal(tr.cpu().numpy().ravel(), np.array([k["oracle"]["trace"][n] for n in names]))
        assert float(r[0]) == np.float32(k["oracle"]["reward"])


@pytest.mark.parametrize("track", [0, -0.5])
@pytest.mark.parametrize("n", [1, 257, 100_000])
def test_step_parity_random_states(sb, O, P98, train_series, cuda_ok, track, n):
    rng = np.random.default_rng(100 + n)
    obs, idx = random_states(rng, n, train_series, P98)
    ref = O.OracleEnv(P98, train_series, 72, n)
    ref.obs[:] = obs
    ref.idx[:] = idx
    env = sb.Shems(72, train_series, n_envs=n)
    env.set_state(obs, idx)
    if track < 0:
        act = ref.action()
        np.testing.assert_array_equal(env.action(track).cpu().numpy(), act)
        if n > 1:  # arbitrary (B, EV) pairs too: the ABI accepts any feasible or infeasible request
            act[:, ::3] = rng.uniform(-4, 12, (2, len(act[0, ::3]))).astype(np.float32)
    else:
        act = rng.uniform(0, 1, (2, n)).astype(np.float32)
        act[:, : n // 10] = rng.choice(np.array([0.0, 0.99, 1.0, 0.98999995], np.float32), (2, n // 10))
        tgt = dev(act)
        np.testing.assert_array_equal(env.action(tgt).cpu().numpy(), ref.action(act))
    r_ref, s_ref, tr_ref = ref.step(act, track=track, want_trace=True)
    r, s2, tr = env.step(dev(act), track=(track if track < 0 else 1))
    np.testing.assert_array_equal(s2.cpu().numpy(), s_ref)           # bit-exact fp32 state
    np.testing.assert_array_equal(env.idx, ref.idx)
    np.testing.assert_allclose(tr.cpu().numpy(), tr_ref, rtol=1e-12, atol=0)
    np.testing.assert_array_equal(r.cpu().numpy(), r_ref.astype(np.float32))
    # north_star tolerance, stated: 1e-5 relative
    np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=1e-5, atol=1e-6)
    # second step runs the FROM_SERIES fast path (state produced by the library): still identical
    act2 = rng.uniform(0, 1, (2, n)).astype(np.float32)
    ok = ref.idx + 1 <= train_series.shape[1]
    if ok.all():
        r_ref, s_ref, _ = ref.step(act2, track=0)
        r, s2 = env.step(dev(act2), track=0)
        np.testing.assert_array_equal(s2.cpu().numpy(), s_ref)
        np.testing.assert_array_equal(r.cpu().numpy(), r_ref.astype(np.float32))


def test_other_chargers_and_rates(sb, O, train_series, cuda_ok):
    rng = np.random.default_rng(5)
    for cid in (4, 6, 9):
        P = O.params_for_charger(cid)
        n = 4096
        obs, idx = random_states(rng, n, train_series, P)
        ref = O.OracleEnv(P, train_series, 72, n)
        ref.obs[:] = obs
        ref.idx[:] = idx
        env = sb.Shems(72, train_series, n_envs=n, charger_id=cid)
        env.set_state(obs, idx)
        act = rng.uniform(0, 1, (2, n)).astype(np.float32)
        r_ref, s_ref, _ = ref.step(act)
        r, s2 = env.step(dev(act))
        np.testing.assert_array_equal(s2.cpu().numpy(), s_ref)
        np.testing.assert_array_equal(r.cpu().numpy(), r_ref.astype(np.float32))


def test_reset_modes(sb, O, P98, train_series, cuda_ok):
    n, T = 4320 - 72, 72
    ref = O.OracleEnv(P98, train_series, T, n)
    env = sb.Shems(T, train_series, n_envs=n)
    ref.reset(mode=0)
    env.reset(rng=-1)
    np.testing.assert_array_equal(env.state, ref.obs)
    np.testing.assert_array_equal(env.idx, ref.idx)
    idx0 = np.arange(1, n + 1, dtype=np.int32)  # every admissible start index -> every path of the shift loop
    socb0 = np.random.default_rng(0).uniform(0, 6.75, n).astype(np.float32)
    ref.reset(mode=1, idx0=idx0, socb0=socb0)
    env.reset(idx0=idx0, socb0=socb0)
    np.testing.assert_array_equal(env.state, ref.obs)
    np.testing.assert_array_equal(env.idx, ref.idx)
    ref.reset(mode=2, seed=77, env_id_base=0)
    env.reset(rng=77)
    np.testing.assert_array_equal(env.state, ref.obs)
    np.testing.assert_array_equal(env.idx, ref.idx)
    assert env.step_count == 0 and env.finished() is False
    with pytest.raises(sb.ShemsError):
        env.reset(idx0=np.zeros(n, np.int32), socb0=socb0)  # idx0 outside 1..nrows-maxsteps


@pytest.mark.parametrize("policy", ["rule", "random", "tape"])
def test_rollout_parity_72_steps(sb, O, P98, train_series, cuda_ok, policy):
    n, T = 4096, 72
    pol = dict(rule=sb.POLICY_RULE, random=sb.POLICY_RANDOM, tape=sb.POLICY_TAPE)[policy]
    ref = O.OracleEnv(P98, train_series, T, n)
    env = sb.Shems(T, train_series, n_envs=n)
    ref.reset(mode=2, seed=3)
    env.reset(rng=3)
    # a tape that feeds a training memory holds the UNSCALED actions remember() stores (DDPG.jl:229); the oracle is given scale_action(a)
    raw = np.random.default_rng(1).uniform(-1, 1, (T, 2, n)).astype(np.float32) if policy == "tape" else None
    tape = ((raw.astype(np.float64) + 1.0) * 0.5).astype(np.float32) if policy == "tape" else None
    want = ref.rollout(pol, T, seed=9, tape=tape, want_transitions=True, want_trace=True)
    if policy == "tape":
        want["a"] = raw
    mem = sb.Replay(T * n)
    got = env.rollout(pol, T, seed=9, tape=dev(raw) if raw is not None else None, replay=mem, want_trace=True, want_obs=True,
                      want_reward=True, tape_unscaled=policy == "tape")
    np.testing.assert_array_equal(got["obs"].cpu().numpy(), want["s2"])          # closed-loop, 72 steps, bit-exact
    np.testing.assert_array_equal(got["reward"].cpu().numpy(), want["r"])
    np.testing.assert_allclose(got["trace"].cpu().numpy(), want["trace"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(got["ep_return"].cpu().numpy(), want["ep_return"], rtol=1e-13, atol=0)
    np.testing.assert_array_equal(env.state, ref.obs)
    np.testing.assert_array_equal(env.idx, ref.idx)
    assert env.step_count == T and len(mem) == T * n
    s, a, r, s2, d = mem.get()  # transitions in push order: step-major, env-minor
    np.testing.assert_array_equal(s, want["s"].transpose(1, 0, 2).reshape(9, -1))
    np.testing.assert_array_equal(a, want["a"].transpose(1, 0, 2).reshape(2, -1))
    np.testing.assert_array_equal(r, want["r"].reshape(-1))
    np.testing.assert_array_equal(s2, want["s2"].transpose(1, 0, 2).reshape(9, -1))
    assert np.all(d == 0)


def test_rollout_equals_step_api(sb, P98, train_series, cuda_ok):
    n, T = 1000, 30
    a = sb.Shems(72, train_series, n_envs=n)
    b = sb.Shems(72, train_series, n_envs=n)
    a.reset(rng=4)
    b.reset(rng=4)
    tape = torch.rand((T, 2, n), device="cuda")
    out = a.rollout(sb.POLICY_TAPE, T, tape=tape, want_obs=True, want_reward=True)
    for t in range(T):
        r, s2 = b.step(tape[t].contiguous())
        assert torch.equal(s2, out["obs"][t]) and torch.equal(r, out["reward"][t])


def test_inference_rule_based_golden_charger98(sb, charger98_test_series, cuda_ok):
    """f1: full-dataset deterministic rule-based inference with the 23-column trace, against the committed golden."""
    g = np.load(os.path.join(GOLDEN, "oracle_rule_based_charger98.npz"))
    env = sb.Shems(2998, charger98_test_series, n_envs=2)
    env.reset(rng=-1)
    out = env.rollout(sb.POLICY_RULE, 2998, want_trace=True)
    tr = out["trace"].cpu().numpy()
    np.testing.assert_array_equal(tr[:, :, 0], tr[:, :, 1])
    np.testing.assert_allclose(tr[:5, :, 0], g["first"], rtol=1e-12)
    np.testing.assert_allclose(tr[-5:, :, 0], g["last"], rtol=1e-12)
    np.testing.assert_allclose(tr[:, :, 0].sum(0), g["colsum"], rtol=1e-11)
    assert float(out["ep_return"][0]) == pytest.approx(float(g["ep_return"][0]), rel=1e-13)
    np.testing.assert_array_equal(env.state[:, 0], g["final_state"])


def test_bounds_error_like_julia(sb, train_series, cuda_ok):
    env = sb.Shems(72, train_series[:, :100], n_envs=8)
    env.reset(rng=-1)
    act = torch.full((2, 8), 0.5, device="cuda")
    for _ in range(99):
        env.step(act)
    before = env.state
    with pytest.raises(IndexError):  # row 101 of a 100-row series (shems_LU1.jl:266-268)
        env.step(act)
    np.testing.assert_array_equal(env.state, before)
    env.reset(rng=-1)
    with pytest.raises(IndexError):
        env.rollout(sb.POLICY_RULE, 100)
    env2 = sb.Shems(72, train_series, n_envs=2)
    with pytest.raises(sb.ShemsError):
        env2.step(torch.zeros((2, 2), device="cuda"))  # step before reset


def test_full_size_properties_config3(sb, O, P98, cuda_ok):
    """BASELINE config 3 shape (2^20 instances, 8761-row year series), checked through size-independent properties:
    battery/EV bounds, the energy-balance identity of the trace on a sampled step, and shard invariance."""
    ser = sb.series.synth_charger98(8761, seed=98)
    n = 1 << 20
    env = sb.Shems(8760, ser, n_envs=n)
    env.reset(rng=1)
    assert np.all(env.idx == 1)  # nrows - maxsteps == 1: the only admissible window
    out = env.rollout(sb.POLICY_RANDOM, 200, seed=5, want_return=True)
    st = env.state_tensor()
    assert float(st[0].min()) >= 0 and float(st[0].max()) <= 6.75 * (1 + 1e-6)
    assert float(st[1].min()) >= 0 and float(st[1].max()) <= 1 + 1e-6
    assert torch.all(st[2] == float(ser[1, 200])) and torch.isfinite(out["ep_return"]).all()
    # the first 4096 instances replayed on the oracle
    ref = O.OracleEnv(P98, ser, 8760, 4096)
    ref.reset(mode=2, seed=1)
    want = ref.rollout(1, 200, seed=5)
    np.testing.assert_allclose(out["ep_return"][:4096].cpu().numpy(), want["ep_return"], rtol=1e-13)
    np.testing.assert_array_equal(st[:, :4096].cpu().numpy(), ref.obs)
    # shard invariance: instances [2^19, 2^19+1024) computed by a handle with env_id_base = 2^19
    sh = sb.Shems(8760, ser, n_envs=1024, env_id_base=1 << 19)
    sh.reset(rng=1)
    o2 = sh.rollout(sb.POLICY_RANDOM, 200, seed=5)
    assert torch.equal(o2["ep_return"], out["ep_return"][(1 << 19):(1 << 19) + 1024])


def test_tracker_csv_and_checkpoint_roundtrip(sb, charger98_test_series, cuda_ok, tmp_path):
    """f1: the 23-column tracker CSV of write_to_results_file (memory_plotting_saving.jl:167-190); f3: checkpoint round trip."""
    env = sb.Shems(2998, charger98_test_series, n_envs=1)
    env.reset(rng=-1)
    out = env.rollout(sb.POLICY_RULE, 2998, want_trace=True)
    sums = sb.tracker.write_results_csv(tmp_path / "rule.csv", out["trace"])
    lines = open(tmp_path / "rule.csv").read().splitlines()
    assert lines[0].split(",") == sb.tracker.TRACE_HEADER and len(lines) == 2999
    back = np.loadtxt(tmp_path / "rule.csv", delimiter=",", skiprows=1)
    np.testing.assert_array_equal(back, out["trace"][:, :, 0].cpu().numpy())          # repr() round-trips Float64 exactly
    assert sums["rewards"] == pytest.approx(float(out["ep_return"][0]), rel=1e-12)
    le = sb.Learner()
    le.init(3)
    sb.tracker.save_checkpoint(tmp_path / "ck.npz", le, s_min=np.zeros(9), s_max=np.ones(9), best_run=7)
    le2 = sb.Learner()
    z = sb.tracker.load_checkpoint(tmp_path / "ck.npz", le2)
    assert int(z["best_run"]) == 7 and z["actor_W2"].shape == (500, 250)
    for net in range(4):
        for k in range(3):
            np.testing.assert_array_equal(le.get_layer(net, k)[0], le2.get_layer(net, k)[0])


def test_instance_groups_equal_separate_handles(sb, O, train_series):
    """shems_create_groups: three chargers in one handle (instance groups with their own constants, two of them with their own
    series) must produce bit-identical states and rewards to three single-charger handles — reset, step API, fused rollout."""
    import torch
    ser2 = sb.series.synth_charger98(4320, seed=7)
    sers = np.stack([train_series, ser2, train_series])
    groups = [(98, 40), (4, 24), (6, 32)]
    multi = sb.Shems(72, sers, groups=groups)
    singles, base = [], 0
    for g, (cid, k) in enumerate(groups):
        singles.append(sb.Shems(72, sers[g], n_envs=k, charger_id=cid, env_id_base=base))
        base += k
    multi.reset(rng=5)
    for e in singles:
        e.reset(rng=5)
    cat = lambda xs: torch.cat(xs, dim=-1)
    assert torch.equal(multi.state_tensor(), cat([e.state_tensor() for e in singles]))
    rng = np.random.default_rng(0)
    for step in range(20):
        a = torch.as_tensor(rng.uniform(0, 1, (2, multi.n_envs)).astype(np.float32), device="cuda")
        r, s2 = multi.step(a)
        rs, ss, off = [], [], 0
        for e in singles:
            ri, si = e.step(a[:, off:off + e.n_envs].contiguous())
            rs.append(ri.clone()); ss.append(si.clone()); off += e.n_envs
        assert torch.equal(r, cat(rs)) and torch.equal(s2, cat(ss))
    multi.reset(rng=9)
    out = multi.rollout(sb.POLICY_RANDOM, 40, seed=3, want_obs=True, want_reward=True)
    off = 0
    for e in singles:
        e.reset(rng=9)
        o = e.rollout(sb.POLICY_RANDOM, 40, seed=3, want_obs=True, want_reward=True)
        sl = slice(off, off + e.n_envs)
        assert torch.equal(out["obs"][:, :, sl], o["obs"]) and torch.equal(out["reward"][:, sl], o["reward"])
        assert torch.equal(out["ep_return"][sl], o["ep_return"])
        off += e.n_envs
    # rule-based controller sees each group's battery size
    b = multi.action(-0.5)
    off = 0
    for e in singles:
        assert torch.equal(b[:, off:off + e.n_envs], e.action(-0.5))
        off += e.n_envs


@pytest.mark.parametrize("variant,cid,dw,pot", [(1, 98, None, None), (1, 4, None, None), (2, 98, None, None), (2, 6, 0.04, 2.0), (0, 98, 0.01, 1.5)])
def test_sibling_env_variants_parity(sb, O, train_series, cuda_ok, variant, cid, dw, pot):
    """f4: shems_LU7.jl (fixed 10 kWh / 4.6f0 kW battery, sell_discount 0.3f0, linear discomfort with weight 1, Float64
    penalty_weight) and shems_LU1_input0607.jl ((discomfort*w)^pot, Float64 penalty_weight) are shems_LU1 with other module
    constants and two type differences; CUDA vs oracle on random states + a 72-step closed rollout, states bit-exact."""
    Pg = sb.params_for_env(variant, cid)
    Po = O.params_for_env(variant, cid)
    for f, _ in Pg._fields_:
        assert getattr(Pg, f) == getattr(Po, f), f                      # library and oracle agree on the module constants
    if dw is not None:
        Pg.discomfort_weight_ev = Po.discomfort_weight_ev = float(np.float32(dw))
        Pg.disc_pot = Po.disc_pot = pot
    n = 50_000
    rng = np.random.default_rng(7 + variant)
    obs, idx = random_states(rng, n, train_series, Po)
    obs[1, : n // 4] = np.where(obs[2, : n // 4] >= 0, rng.uniform(0, 0.9, n // 4), 1.0)   # many unfinished sessions
    obs[2, : n // 8] = 0.0                                                                   # ... departing now: discomfort term
    ref = O.OracleEnv(Po, train_series, 72, n)
    ref.obs[:] = obs
    ref.idx[:] = idx
    env = sb.Shems(72, train_series, n_envs=n, params=Pg)
    env.set_state(obs, idx)
    act = rng.uniform(0, 1, (2, n)).astype(np.float32)
    r_ref, s_ref, tr_ref = ref.step(act, track=1, want_trace=True)
    r, s2, tr = env.step(dev(act), track=1)
    np.testing.assert_array_equal(s2.cpu().numpy(), s_ref)
    np.testing.assert_allclose(tr.cpu().numpy(), tr_ref, rtol=1e-12, atol=0)
    assert (tr_ref[7] > 0).sum() > 100 and (tr_ref[8] > 0).sum() > 100     # both the discomfort and the penalty term were exercised
    if variant != 0:  # Float64 penalty weight: (1 - EV_target)::Float32 * 0.1::Float64, not the Float32 product of shems_LU1
        k = np.flatnonzero(tr_ref[8] > 0)[:1000]
        np.testing.assert_array_equal(tr_ref[8][k], (np.float32(1) - act[1, k]).astype(np.float64) * 0.1)
    env.reset(rng=3)
    ref.reset(mode=2, seed=3)
    out = env.rollout(sb.POLICY_RANDOM, 72, seed=3, want_obs=True, want_reward=True)
    want = ref.rollout(1, 72, seed=3, want_transitions=True)
    np.testing.assert_array_equal(out["obs"].cpu().numpy(), want["s2"])
    np.testing.assert_array_equal(out["reward"].cpu().numpy(), want["r"])
    np.testing.assert_allclose(out["ep_return"].cpu().numpy(), want["ep_return"], rtol=1e-12, atol=0)


def test_sibling_env_unknown_keys(sb):
    with pytest.raises(sb.ShemsKeyError):
        sb.params_for_env(sb.ENV_LU7, 97)           # shems_LU7.jl's ev_capacities has no charger 97
    with pytest.raises(sb.ShemsKeyError):
        sb.params_for_env(sb.ENV_LU1_INPUT0607, 97)
    assert sb.params_for_env(sb.ENV_LU7, 99).ev_soc_max == np.float32(35.816)
    with pytest.raises(sb.ShemsError):
        sb.params_for_env(7, 98)
