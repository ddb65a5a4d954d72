"""GPU tests of the round-2 additions (all through the C ABI):
  * full learner snapshot: a resumed run is BIT-identical to an uninterrupted one (ddpg_get/set_state, checkpoints of tracker.py)
  * Flux.ADAM exactness: injected gradients, CUDA step vs a numpy Float64 restatement of Flux 0.12.1's apply!/update!, bit for bit
  * ddpg_rollout (actor policy + step! as one persistent cluster kernel) vs the step-by-step loop and vs the oracle
  * Float64 reward of step!, noise_eps of episode!, unscaled action tapes
  * learned-policy returns at the reference's own shape (250/500, B = 120): CUDA and the C oracle trained side by side with
    identical initial weights, injected noise and minibatch indices; evaluation returns within a stated tolerance (north star).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def _memory(sb, series, n=64, T=72, seed=2):
    env = sb.Shems(T, series, n_envs=n)
    mem = sb.Replay(n * T)
    env.reset(rng=seed)
    env.rollout(sb.POLICY_RANDOM, T, seed=seed, replay=mem, want_return=False)
    return mem


# ------------------------------------------------------------------------------------------------ resume
@pytest.mark.parametrize("fused", [True, False])
def test_resumed_run_is_bit_identical(sb, train_series, fused):
    mem = _memory(sb, train_series)
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)

    def fresh():
        le = sb.Learner()
        le.set_fused(fused)
        le.init(5)
        le.set_norm(mn, mx)
        return le
    a = fresh()
    for k in range(7):
        a.replay(mem, rng_rpl=100 + k, n_updates=1)
    state, opt = a.get_state()
    assert opt[4] == 7 and opt[0] == pytest.approx(0.9 ** 8) and opt[3] == pytest.approx(0.999 ** 8)
    for k in range(7, 15):
        a.replay(mem, rng_rpl=100 + k, n_updates=1)
    end_a, opt_a = a.get_state()
    b = sb.Learner()           # a different handle, nothing but the snapshot
    b.set_fused(fused)
    b.set_state(state, opt)
    for k in range(7, 15):
        b.replay(mem, rng_rpl=100 + k, n_updates=1)
    end_b, opt_b = b.get_state()
    assert end_a.tobytes() == end_b.tobytes() and opt_a.tobytes() == opt_b.tobytes()
    # a learner that is NOT given the optimiser state does not reproduce the run (the reference's situation: actor-only BSON)
    c = fresh()
    na = int(sb._lib.lib().ddpg_num_params(c._h, 0))
    c.set_state(state, np.array([0.9, 0.999, 0.9, 0.999, 0, 0, 0, 0.0]))   # β powers of a fresh optimiser: a different step size
    for k in range(7, 15):
        c.replay(mem, rng_rpl=100 + k, n_updates=1)
    assert c.get_state()[0][:na].tobytes() != end_b[:na].tobytes()
    # ddpg_init on a used handle starts fresh optimisers
    a.init(5)
    st0, opt0 = a.get_state()
    nc = int(sb._lib.lib().ddpg_num_params(a._h, 1))
    assert opt0[4] == 0 and opt0[0] == 0.9 and opt0[1] == 0.999 and not st0[2 * na + 2 * nc: 4 * na + 4 * nc].any()


def test_checkpoint_file_round_trip_with_memory_and_population(sb, train_series, tmp_path):
    P = 3
    mems = [_memory(sb, train_series, n=16, seed=3 + l) for l in range(P)]
    pop = sb.Learner(params=sb.default_ddpg_params(population=P, batch=64))
    pop.init(11)
    for l in range(P):
        mn, mx = mems[l].min_max_buffer(len(mems[l]), rng_mm=l)
        pop.select(l).set_norm(mn, mx)
    pop.replay(mems, rng_rpl=50, n_updates=3)
    path = str(tmp_path / "ckpt.npz")
    sb.tracker.save_checkpoint(path, pop, memories=mems, total_reward=np.arange(3.0))
    pop.replay(mems, rng_rpl=60, n_updates=4)
    want = [pop.select(l).get_state() for l in range(P)]
    pop2 = sb.Learner(params=sb.default_ddpg_params(population=P, batch=64))
    mems2 = [sb.Replay(m.capacity) for m in mems]
    z = sb.tracker.load_checkpoint(path, pop2, memories=mems2)
    assert np.array_equal(z["total_reward"], np.arange(3.0)) and z["actor_W1"].shape == (250, 9)
    for m, m2 in zip(mems, mems2):
        for x, y in zip(m.get(), m2.get()):
            assert x.tobytes() == y.tobytes()
    pop2.replay(mems2, rng_rpl=60, n_updates=4)
    for l in range(P):
        st, opt = pop2.select(l).get_state()
        assert st.tobytes() == want[l][0].tobytes() and opt.tobytes() == want[l][1].tobytes(), l


# ------------------------------------------------------------------------------------------------ ADAM
def flux_adam_step(x, g, m, v, bp, eta, b1=0.9, b2=0.999, eps=1e-8):
    """Flux 0.12.1 Optimise.ADAM apply! + update! on Float32 arrays with Float64 β, βp, ϵ, η (struct fields are Float64):
         @. mt = β[1] * mt + (1 - β[1]) * Δ;  @. vt = β[2] * vt + (1 - β[2]) * Δ^2
         @. Δ  = mt / (1 - βp[1]) / (√(vt / (1 - βp[2])) + ϵ) * η;  βp .= βp .* β;  x .-= Δ"""
    f64 = np.float64
    m[:] = (b1 * m.astype(f64) + (1.0 - b1) * g.astype(f64)).astype(np.float32)
    v[:] = (b2 * v.astype(f64) + (1.0 - b2) * (g * g).astype(f64)).astype(np.float32)     # Δ^2 is a Float32 product
    d = (m.astype(f64) / (1.0 - bp[0]) / (np.sqrt(v.astype(f64) / (1.0 - bp[1])) + eps) * float(np.float32(eta))).astype(np.float32)
    x[:] = x - d
    bp[0] *= b1
    bp[1] *= b2


def test_adam_bit_exact_against_flux_restatement(sb):
    """The optimiser kernel pays for Flux's Float64 element math (Markstein divisions, DSQRT): prove that it buys exactness."""
    L = sb._lib
    le = sb.Learner()
    le.set_fused(False)            # ddpg_update_phase drives the tiled path's optimiser kernel; the fused path shares adam_element
    le.init(3)
    rng = np.random.default_rng(0)
    na, nc = int(L.lib().ddpg_num_params(le._h, 0)), int(L.lib().ddpg_num_params(le._h, 1))
    g = le.grad_tensor()
    off_a = g.numel() - na        # [critic | pad | actor]
    st, opt = le.get_state()
    x_a, x_c = st[:na].copy(), st[na:na + nc].copy()
    t_a, t_c = st[na + nc:2 * na + nc].copy(), st[2 * na + nc:2 * na + 2 * nc].copy()
    m_a, v_a, m_c, v_c = np.zeros(na, np.float32), np.zeros(na, np.float32), np.zeros(nc, np.float32), np.zeros(nc, np.float32)
    bp_a, bp_c = [0.9, 0.999], [0.9, 0.999]
    tau, omt = np.float32(1e-3), np.float32(1) - np.float32(1e-3)
    for rnd in range(4):
        scale = [1.0, 1e-6, 1e3, 1e-20][rnd]
        gc = (rng.normal(0, 1, nc) * scale).astype(np.float32)
        ga = (rng.normal(0, 1, na) * scale).astype(np.float32)
        gc[::7] = 0.0
        ga[::5] = 0.0
        gc[1::7] *= np.float32(1e-12)
        g[:nc] = dev(gc)
        L.check(L.lib().ddpg_update_phase(le._h, None, 1, None, 0, 1.0))   # ADAM(critic) with the injected gradient (then the actor pass)
        g[off_a:] = dev(ga)
        L.check(L.lib().ddpg_update_phase(le._h, None, 2, None, 0, 1.0))   # ADAM(actor) + soft_update! of both targets
        flux_adam_step(x_c, gc, m_c, v_c, bp_c, 1e-3)
        flux_adam_step(x_a, ga, m_a, v_a, bp_a, 1e-4)
        t_a = omt * t_a + tau * x_a                                        # p_t .= (1f0 - τ) * p_t .+ τ * p_m  (DDPG.jl:99-103)
        t_c = omt * t_c + tau * x_c
        st, opt = le.get_state()
        got = dict(actor=st[:na], critic=st[na:na + nc], actor_t=st[na + nc:2 * na + nc], critic_t=st[2 * na + nc:2 * na + 2 * nc],
                   m_a=st[2 * na + 2 * nc:3 * na + 2 * nc], v_a=st[3 * na + 2 * nc:4 * na + 2 * nc],
                   m_c=st[4 * na + 2 * nc:4 * na + 3 * nc], v_c=st[4 * na + 3 * nc:4 * na + 4 * nc])
        want = dict(actor=x_a, critic=x_c, actor_t=t_a, critic_t=t_c, m_a=m_a, v_a=v_a, m_c=m_c, v_c=v_c)
        for k in want:
            assert got[k].tobytes() == want[k].astype(np.float32).tobytes(), (rnd, k, np.abs(got[k] - want[k]).max())
        assert opt[0] == bp_c[0] and opt[1] == bp_c[1] and opt[2] == bp_a[0] and opt[3] == bp_a[1] and opt[4] == rnd + 1


# ------------------------------------------------------------------------------------------------ actor rollout kernel
@pytest.mark.parametrize("n", [1, 8, 13, 100])
def test_actor_rollout_kernel_equals_step_loop_and_oracle(sb, O, charger98_test_series, n):
    """ddpg_rollout (one persistent cluster kernel) against (a) act + step! launched step by step through the same ABI — bit-exact —
    and (b) the CPU oracle's actor and environment (closed loop: states must stay bit-exact while the actions agree to 1e-6)."""
    ser = charger98_test_series
    T = 96
    le = sb.Learner()
    le.init(21)
    rng = np.random.default_rng(n)
    mn = np.array([0, 0, -1, 0, 0, 0.4, -1, -1, 1], np.float32)
    mx = np.array([6.75, 1, 60, 8, 20, 0.4, 1, 1, 4], np.float32)
    le.set_norm(mn, mx)
    # larger last-layer weights so that the actions move around inside (-1, 1)
    w, b = le.get_layer(0, 2)
    le.set_layer(0, 2, (w * 40).astype(np.float32), rng.normal(0, 0.3, 2).astype(np.float32))
    idx0 = rng.integers(1, ser.shape[1] - T, n).astype(np.int32)
    socb0 = rng.uniform(0, 6.75, n).astype(np.float32)
    env = sb.Shems(T, ser, n_envs=n)
    env.reset(idx0=idx0, socb0=socb0)
    start_state, start_idx = env.state, env.idx
    out = le.rollout(env, T, want_trace=True, want_actions=True)
    end_state, end_idx = env.state, env.idx
    assert env.step_count == T and np.array_equal(end_idx, start_idx + T)
    # (a) the step-by-step loop
    env2 = sb.Shems(T, ser, n_envs=n)
    env2.reset(idx0=idx0, socb0=socb0)
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    r64 = torch.empty(n, dtype=torch.float64, device="cuda")
    for t in range(T):
        st = env2.state_tensor()
        parts = [le.act(st[:, c:c + 50].contiguous(), train=False) for c in range(0, n, 50)]   # <= 64 states: the fused act kernel
        a, scaled = torch.cat([p[0] for p in parts], 1).contiguous(), torch.cat([p[1] for p in parts], 1).contiguous()
        assert torch.equal(a, out["actions"][t]), t
        r, s2, tr = env2.step(scaled, track=1, reward64_out=r64)
        assert torch.equal(tr, out["trace"][t]), t
        assert torch.equal(r64, tr[5]) and torch.equal(r, r64.float())   # Float64 env.reward and its Float32 value
        ret += r64
    assert np.array_equal(env2.state, end_state) and torch.equal(ret, out["ep_return"])
    # (b) the oracle
    orc = O.OracleDdpg(O.default_ddpg_params())
    for net in range(4):
        for k in range(3):
            orc.set_layer(net, k, *le.get_layer(net, k))
    orc.set_norm(mn, mx)
    ref = O.OracleEnv(O.params_for_charger(98), ser, T, n)
    ref.obs[:] = start_state
    ref.idx[:] = start_idx
    acts = out["actions"].cpu().numpy()
    trace = out["trace"].cpu().numpy()
    for t in range(T):
        oa, osc = orc.act(ref.obs.copy())
        assert np.abs(oa - acts[t]).max() < 1e-6, (t, np.abs(oa - acts[t]).max())
        # feed the CUDA action so that the environments stay comparable bit for bit
        sc = ((acts[t].astype(np.float64) + 1.0) * 0.5).astype(np.float32)
        r_ref, s_ref, tr_ref = ref.step(sc, track=1, want_trace=True)
        np.testing.assert_allclose(trace[t], tr_ref, rtol=1e-12, atol=0)
    assert np.array_equal(ref.obs, end_state)


def test_actor_rollout_bounds_and_groups(sb, train_series):
    le = sb.Learner()
    le.init(1)
    env = sb.Shems(72, train_series, n_envs=4)
    env.reset(rng=-1)
    with pytest.raises(IndexError):
        le.rollout(env, train_series.shape[1])        # would read past the last row: BoundsError, nothing launched
    assert env.step_count == 0
    # two chargers in one handle: each group runs with its own constants
    g = sb.Shems(72, train_series, groups=[(98, 5), (4, 6)])
    g.reset(rng=-1)
    out = le.rollout(g, 72)
    single = []
    for cid, k in ((98, 5), (4, 6)):
        e = sb.Shems(72, train_series, n_envs=k, charger_id=cid)
        e.reset(rng=-1)
        single.append(le.rollout(e, 72)["ep_return"])
    assert torch.equal(out["ep_return"], torch.cat(single))


def test_rollout_fallback_for_wide_nets(sb, train_series):
    """nets wider than the cluster kernel's shared-memory plan (l2 > 512) take one act + step! launch pair per step: same contract"""
    le = sb.Learner(params=sb.default_ddpg_params(l1=64, l2=640, batch=32))
    le.init(2)
    env = sb.Shems(72, train_series, n_envs=9)
    env.reset(rng=5)
    s0, i0 = env.state, env.idx
    out = le.rollout(env, 20, want_trace=True, want_actions=True)
    env2 = sb.Shems(72, train_series, n_envs=9)
    env2.set_state(s0, i0)
    ret = np.zeros(9)
    for t in range(20):
        a, scaled = le.act(env2.state_tensor(), train=False)
        r, s2, tr = env2.step(scaled, track=1)
        assert torch.equal(tr, out["trace"][t]) and torch.equal(a, out["actions"][t])
        ret += tr[5].cpu().numpy()
    np.testing.assert_array_equal(ret, out["ep_return"].cpu().numpy())


def test_episode_noise_eps_and_float64_return(sb, O, train_series):
    """episode!(train = true) returns reward_eps as the Float64 sum of the Float64 step rewards (DDPG.jl:223) and noise_eps as the
    sum over the steps of mean(noise) (:224).  The noise is re-derived here from the documented stream — Philox(rng_step, env id,
    step, STREAM_NOISE) + Box-Muller — and the actions from ddpg_act with that noise injected."""
    n, T, sigma = 6, 30, 0.1
    mem = sb.Replay(4096)
    le, le2 = sb.Learner(), sb.Learner()
    for x in (le, le2):
        x.init(8)
    env, env2 = sb.Shems(72, train_series, n_envs=n), sb.Shems(72, train_series, n_envs=n)
    env.reset(rng=3)
    env2.reset(rng=3)
    ret, nz = le.episode(env, mem, T, train=True, sigma=sigma, rng_ep=77, updates_per_step=0, want_noise=True)
    lib = O.lib()
    lib.oracle_u53.restype = C.c_double
    lib.oracle_u53.argtypes = [C.c_uint32, C.c_uint32]
    ret2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    nz2 = np.zeros(n, np.float32)
    r64 = torch.empty(n, dtype=torch.float64, device="cuda")
    sig = float(np.float32(sigma))
    for step in range(1, T + 1):
        rng_step = (77 * 1000003 + step) & (2**63 - 1)
        noise = np.zeros((2, n), np.float32)
        for j in range(n):
            w = (C.c_uint32 * 4)()
            lib.oracle_philox(rng_step, j, step, 0x4e4f, w)
            u1, u2 = 1.0 - lib.oracle_u53(w[0], w[1]), lib.oracle_u53(w[2], w[3])
            rad = np.sqrt(-2.0 * np.log(u1))
            noise[0, j], noise[1, j] = sig * (rad * np.cos(2 * np.pi * u2)), sig * (rad * np.sin(2 * np.pi * u2))
        s = env2.state_tensor().clone()
        a, scaled = le2.act(s, noise=dev(noise))
        env2.step(scaled, reward64_out=r64)
        ret2 += r64
        nz2 = nz2 + (noise[0] + noise[1]) * np.float32(0.5)
    np.testing.assert_array_equal(env.state, env2.state)        # same noise -> same actions -> same trajectory
    assert torch.equal(ret, ret2)
    np.testing.assert_allclose(nz.cpu().numpy(), nz2, rtol=0, atol=1e-6)   # cos/sin(2 pi u) in numpy vs sincospi on the device
    assert len(mem) == n * T


def test_unscaled_tape_feeds_a_training_memory(sb, O, train_series):
    """POLICY_TAPE with a replay sink must carry the UNSCALED action (remember stores a in [-1,1], DDPG.jl:229): the scaled-tape
    form is refused, the unscaled form stores a and steps with scale_action(a)."""
    n, T = 32, 10
    rng = np.random.default_rng(4)
    a = rng.uniform(-1, 1, (T, 2, n)).astype(np.float32)
    env = sb.Shems(72, train_series, n_envs=n)
    mem = sb.Replay(n * T)
    env.reset(rng=9)
    with pytest.raises(sb.ShemsError):
        env.rollout(sb.POLICY_TAPE, T, tape=dev((a + 1) / 2), replay=mem)
    out = env.rollout(sb.POLICY_TAPE, T, tape=dev(a), replay=mem, tape_unscaled=True, want_obs=True)
    S, A, R, S2, D = mem.get()
    np.testing.assert_array_equal(A.reshape(2, T, n).transpose(1, 0, 2), a)
    ref = O.OracleEnv(O.params_for_charger(98), train_series, 72, n)
    ref.reset(mode=2, seed=9)
    for t in range(T):
        sc = ((a[t].astype(np.float64) + 1.0) * 0.5).astype(np.float32)
        ref.step(sc)
        np.testing.assert_array_equal(out["obs"][t].cpu().numpy(), ref.obs)


def test_replay_sample_is_the_minibatch_replay_trains_on(sb, O, train_series):
    """getData(rng) is one sample in both of its uses (memory_plotting_saving.jl:31-42): Replay.sample(B, rng_dt = r) returns the
    transitions Learner.replay(rng_rpl = r) trains on, whatever the learner's update count."""
    mem = _memory(sb, train_series)
    le, le2 = sb.Learner(), sb.Learner()
    for x in (le, le2):
        x.set_fused(False)
        x.init(3)
    le.replay(mem, rng_rpl=5, n_updates=3)      # advance the update counter
    st, opt = le.get_state()
    le2.set_state(st, opt)
    le.replay(mem, rng_rpl=99, n_updates=1)
    s, a, r, s2, d = mem.sample(120, rng_dt=99)
    le2.update_batch(s, a, r, s2, d)
    assert le.get_state()[0].tobytes() == le2.get_state()[0].tobytes()


# ------------------------------------------------------------------------------------------------ learned returns (north star)
def test_learned_returns_match_oracle_at_reference_shape(sb, O, train_series, charger98_test_series):
    """BASELINE configs[0] in small: ONE instance, EP_LENGTH = 72, B = 120, 250/500, one replay() per step (DDPG.jl:186-242), on CUDA
    (cluster-fused update, fused act) and on the CPU oracle with identical initial weights, start rows, injected Gaussian noise and
    host-drawn minibatch indices, for 5 training episodes.  After every episode both policies are evaluated without noise on the
    first 72 rows of an evaluation series (what run_episodes does, :266-279).

    STATED TOLERANCE: evaluation and training returns within 1e-3 relative (+1e-3 absolute).  The fp32 summation order of the CUDA
    kernels differs from the oracle's Float64 accumulation; ADAM's normalised step turns that rounding noise into weight differences
    of a few per cent of lr per update, which 360 updates compound — the measured deviation is printed.  Longer horizons (SHEMS_G1_EPISODES,
    tests/learned_returns_divergence.py): past ~400 updates the closed loop amplifies rounding tenfold per episode, for CUDA against the oracle
    exactly as for the two CUDA update paths against each other (profiles/r2_learned_returns.md) — hence 5 episodes here."""
    O.set_threads(max(1, min(16, (os.cpu_count() or 2) // 2)))
    T, B, EPISODES = 72, 120, int(os.environ.get("SHEMS_G1_EPISODES", "5"))   # longer horizons: profiles/r2_learned_returns.md
    rng = np.random.default_rng(2024)
    warm = _memory(sb, train_series, n=64, T=T, seed=12)             # a warm-up memory both sides share (random policy)
    S, A, R, S2, D = warm.get()
    cap = S.shape[1] + EPISODES * T
    mem = sb.Replay(cap)
    mem.push(dev(S), dev(A), dev(R), dev(S2), dev(D))
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)
    le = sb.Learner()
    assert le.set_fused(True)
    le.init(1231)
    orc = O.OracleDdpg(O.default_ddpg_params())
    for net in range(4):
        for k in range(3):
            orc.set_layer(net, k, *le.get_layer(net, k))
    le.set_norm(mn, mx)
    orc.set_norm(mn, mx)
    P = O.params_for_charger(98)
    env = sb.Shems(T, train_series, n_envs=1)
    ref = O.OracleEnv(P, train_series, T, 1)
    ev = sb.Shems(T, charger98_test_series, n_envs=1)
    ev_ref = O.OracleEnv(P, charger98_test_series, T, 1)
    worst_eval = worst_train = worst_action = 0.0
    for ep in range(EPISODES):
        idx0 = rng.integers(1, train_series.shape[1] - T, 1).astype(np.int32)
        socb0 = rng.uniform(0, 6.75, 1).astype(np.float32)
        env.reset(idx0=idx0, socb0=socb0)
        ref.reset(mode=1, idx0=idx0, socb0=socb0)
        ret_gpu = ret_ref = 0.0
        r64 = torch.empty(1, dtype=torch.float64, device="cuda")
        for step in range(T):
            noise = rng.normal(0, 0.1, (2, 1)).astype(np.float32)
            s_gpu = env.state_tensor().clone()
            a, scaled = le.act(s_gpu, noise=dev(noise))
            r, s2 = env.step(scaled, reward64_out=r64)
            ret_gpu += float(r64[0])
            mem.push(s_gpu, a, r, s2)
            s_ref = ref.obs.copy()
            oa, osc = orc.act(s_ref, noise=noise)
            r_ref, s2_ref, _ = ref.step(osc)
            ret_ref += float(r_ref[0])
            S = np.concatenate([S, s_ref], 1); A = np.concatenate([A, oa], 1); R = np.concatenate([R, r_ref.astype(np.float32)])
            S2 = np.concatenate([S2, s2_ref], 1); D = np.concatenate([D, np.zeros(1, np.float32)])
            idx = rng.integers(0, S.shape[1], B).astype(np.int32)
            le.replay(mem, n_updates=1, idx=idx)
            orc.update_batch(S[:, idx], A[:, idx], R[idx], S2[:, idx], D[idx])
            worst_action = max(worst_action, float(np.abs(a.cpu().numpy() - oa).max()))
        worst_train = max(worst_train, abs(ret_gpu - ret_ref) / (abs(ret_ref) + 1.0))
        # evaluation episode: no noise, first 72 rows, Soc_b = 50 % (reset!(rng = -1))
        ev.reset(rng=-1)
        ev_ref.reset(mode=0)
        score = float(le.rollout(ev, T)["ep_return"][0])
        score_ref = 0.0
        for step in range(T):
            oa, osc = orc.act(ev_ref.obs.copy())
            r_ref, _, _ = ev_ref.step(osc)
            score_ref += float(r_ref[0])
        worst_eval = max(worst_eval, abs(score - score_ref) / (abs(score_ref) + 1.0))
        print("episode %d: train return %.6f / %.6f, eval score %.6f / %.6f (CUDA / oracle)" % (ep + 1, ret_gpu, ret_ref, score, score_ref))
    print("worst relative deviation: eval %.2e, train %.2e; worst action difference %.2e" % (worst_eval, worst_train, worst_action))
    assert worst_eval < 1e-3 and worst_train < 1e-3


# ------------------------------------------------------------------------------------------------ large-batch forward chain
@pytest.mark.parametrize("B,l1,l2", [(1024, 250, 500), (8192, 250, 500), (1000, 250, 500), (384, 64, 128), (2048, 256, 512)])
def test_forward_chain_kernel_equals_layerwise_path(sb, train_series, monkeypatch, B, l1, l2):
    """tc_fwd_chain_kernel (layer 1 by SIMT into the swizzled A operand, tcgen05 layer 2, output layer in the epilogue: one kernel per
    net; with few row tiles a cluster of two CTAs per tile) and tc_bwd_chain_kernel (the dX chain through the critic in the actor pass:
    dz2 in place on the TMA-loaded h2 block, tcgen05 against W2, relu' mask and the action rows of W1 in the epilogue) against the
    layer-by-layer tensor-core path (l1_fwd + tc_gemm + gemm_skinny, outer_mask + tc_gemm + gemm_skinny): same TF32 products in the same k
    order, so activations agree to fp32 rounding of the differently ordered thin dot products; one whole update is compared.  The shapes
    cover one and two CTAs per tile, ragged last tiles (1000 rows), narrow nets and the widest ones the kernels take."""
    mem = _memory(sb, train_series, n=256, seed=4)
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)
    idx = np.random.default_rng(1).integers(0, len(mem), B).astype(np.int32)
    res = []
    for chain in ("1", "0"):
        monkeypatch.setenv("SHEMS_TC_CHAIN", chain)
        le = sb.Learner(params=sb.default_ddpg_params(batch=B, l1=l1, l2=l2, use_tensor_cores=1))
        le.init(6)
        le.set_norm(mn, mx)
        le.replay(mem, n_updates=1, idx=idx)
        grads = [le.get_grad(net, k) for net in (0, 1) for k in range(3)]
        res.append((le.losses(), grads, le.get_state()[0]))
    (lc1, la1), g1, st1 = res[0]
    (lc0, la0), g0, st0 = res[1]
    assert lc1 == pytest.approx(lc0, rel=1e-5) and la1 == pytest.approx(la0, rel=1e-5, abs=1e-7)
    for (w1, b1), (w0, b0) in zip(g1, g0):
        scale = np.abs(w0).max() + 1e-12
        assert np.abs(w1 - w0).max() <= 2e-5 * scale and np.abs(b1 - b0).max() <= 2e-5 * (np.abs(b0).max() + 1e-12)


@pytest.mark.parametrize("P,B,l1,l2", [(3, 96, 48, 64), (5, 120, 250, 500)])
def test_forward_chain_kernel_population(sb, train_series, monkeypatch, P, B, l1, l2):
    """The same kernel with the learner as grid.y (BASELINE configs[4]: one 128-row tile per learner and net, every pointer and TMA
    coordinate moved to the learner's slab) against the population's layer-by-layer path: per learner the same TF32 products."""
    mems = [_memory(sb, train_series, n=64, seed=10 + l) for l in range(P)]
    norms = [m.min_max_buffer(len(m), rng_mm=l) for l, m in enumerate(mems)]
    idx = np.stack([np.random.default_rng(l).integers(0, len(mems[l]), (2, B)) for l in range(P)]).astype(np.int32)
    res = []
    for chain in ("1", "0"):
        monkeypatch.setenv("SHEMS_TC_CHAIN", chain)
        pop = sb.Learner(params=sb.default_ddpg_params(population=P, batch=B, l1=l1, l2=l2, use_tensor_cores=1))
        pop.init(60)
        for l in range(P):
            pop.select(l).set_norm(*norms[l])
        pop.replay(mems, n_updates=2, idx=idx)
        out = []
        for l in range(P):
            pop.select(l)
            out.append((pop.losses(), [pop.get_grad(net, k) for net in (0, 1) for k in range(3)], [pop.get_layer(net, 1)[0] for net in range(4)]))
        res.append(out)
        pop.close()
    for l in range(P):
        (lc1, la1), g1, w1s = res[0][l]
        (lc0, la0), g0, w0s = res[1][l]
        assert lc1 == pytest.approx(lc0, rel=2e-4) and la1 == pytest.approx(la0, rel=2e-4, abs=1e-6), l
        for (w1, b1), (w0, b0) in zip(g1, g0):
            assert np.abs(w1 - w0).max() <= 1e-3 * (np.abs(w0).max() + 1e-12) and np.abs(b1 - b0).max() <= 1e-3 * (np.abs(b0).max() + 1e-12), l
        for w1, w0 in zip(w1s, w0s):
            assert np.abs(w1 - w0).max() <= 2e-5, l
        assert not np.array_equal(res[0][l][2][1], res[0][(l + 1) % P][2][1])   # the learners are different learners
