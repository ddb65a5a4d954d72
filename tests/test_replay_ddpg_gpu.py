"""GPU parity tests: replay ring, on-device sampling and the DDPG minibatch update against the CPU oracle.

Tolerances (stated): the oracle accumulates dot products in double and rounds once; the CUDA kernels accumulate in
fp32 (FMA, tiled order).  Forward outputs / gradients agree to 2e-4 relative (+1e-6·scale absolute); parameters after
K updates agree to 2% of the distance Adam can move them (lr·K) plus 1e-5 relative — Adam's normalised step
g/(|g|+ε) amplifies rounding noise of near-zero gradients, which bounds what any fp32 implementation can match.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def test_ring_push_wrap_and_get(sb):
    cap, n = 1000, 300
    mem = sb.Replay(cap)
    rng = np.random.default_rng(0)
    allr = []
    for k in range(5):  # 1500 transitions through a 1000-slot ring
        s, a, r, s2 = (rng.normal(size=(9, n)).astype(np.float32), rng.normal(size=(2, n)).astype(np.float32),
                       rng.normal(size=n).astype(np.float32), rng.normal(size=(9, n)).astype(np.float32))
        mem.push(dev(s), dev(a), dev(r), dev(s2))
        allr.append((s, a, r, s2))
        assert len(mem) == min(cap, (k + 1) * n)
    S = np.concatenate([x[0] for x in allr], 1)[:, -cap:]
    A = np.concatenate([x[1] for x in allr], 1)[:, -cap:]
    R = np.concatenate([x[2] for x in allr])[-cap:]
    S2 = np.concatenate([x[3] for x in allr], 1)[:, -cap:]
    s, a, r, s2, d = mem.get()  # oldest first, like a CircularBuffer
    np.testing.assert_array_equal(s, S); np.testing.assert_array_equal(a, A)
    np.testing.assert_array_equal(r, R); np.testing.assert_array_equal(s2, S2)
    assert np.all(d == 0)
    # getData with explicit indices
    idx = rng.integers(0, cap, 120).astype(np.int32)
    bs, ba, br, bs2, bd = mem.sample(120, idx=idx)
    np.testing.assert_array_equal(bs.cpu().numpy(), S[:, idx]); np.testing.assert_array_equal(ba.cpu().numpy(), A[:, idx])
    np.testing.assert_array_equal(br.cpu().numpy(), R[idx]); np.testing.assert_array_equal(bs2.cpu().numpy(), S2[:, idx])
    with pytest.raises(sb.ShemsError):
        mem.sample(4, idx=np.array([0, 1, 2, cap], np.int32))
    # one push larger than the ring keeps the newest `cap`
    big = sb.Replay(100)
    s = rng.normal(size=(9, 250)).astype(np.float32)
    z2, z1 = np.zeros((2, 250), np.float32), np.arange(250, dtype=np.float32)
    big.push(dev(s), dev(z2), dev(z1), dev(s))
    assert len(big) == 100
    np.testing.assert_array_equal(big.get()[2], z1[-100:])


def test_device_sampling_matches_oracle_spec(sb, O):
    mem = sb.Replay(24000)
    n = 24000
    r = np.arange(n, dtype=np.float32)
    z9, z2 = np.zeros((9, n), np.float32), np.zeros((2, n), np.float32)
    z9[0] = r
    mem.push(dev(z9), dev(z2), dev(r), dev(z9))
    s, a, rr, s2, d = mem.sample(120, rng_dt=12345)
    want = O.sample_indices(12345, 0, n, 120)  # i.i.d. with replacement (memory_plotting_saving.jl:33)
    np.testing.assert_array_equal(rr.cpu().numpy().astype(np.int64), want)
    mn, mx = mem.min_max_buffer(n, rng_mm=7)
    draws = O.sample_indices(7, 0, n, n)
    assert mn[0] == draws.min() and mx[0] == draws.max()  # min/max over the SAMPLE, not the buffer (quirk 6)
    assert len(np.unique(draws)) / n == pytest.approx(0.632, abs=0.01)
    mn2, mx2 = mem.min_max_buffer(5, idx=np.array([5, 9, 2, 2, 7], np.int32))
    assert (mn2[0], mx2[0]) == (2, 9)
    with pytest.raises(sb.ShemsError):
        sb.Replay(10).sample(4)


def _sync_nets(learner, orc):
    for net in range(4):
        for k in range(3):
            w, b = orc.get_layer(net, k)
            learner.set_layer(net, k, w, b)


def _compare_nets(learner, orc, atol_w, rtol=1e-5, nets=range(4)):
    for net in nets:
        for k in range(3):
            w, b = learner.get_layer(net, k)
            ow, ob = orc.get_layer(net, k)
            np.testing.assert_allclose(w, ow, rtol=rtol, atol=atol_w[net])
            np.testing.assert_allclose(b, ob, rtol=rtol, atol=atol_w[net])


def test_init_matches_oracle_spec(sb, O):
    p = sb.default_ddpg_params()
    le = sb.Learner(params=p)
    le.init(42)
    orc = O.OracleDdpg(O.default_ddpg_params())
    orc.init(42)
    for net in range(4):
        for k in range(3):
            w, b = le.get_layer(net, k)
            ow, ob = orc.get_layer(net, k)
            np.testing.assert_array_equal(w, ow)
            np.testing.assert_array_equal(b, ob)
    assert le.lib.ddpg_num_params(le._h, 0) == 129_002 and le.lib.ddpg_num_params(le._h, 1) == 129_001


@pytest.mark.parametrize("B,l1,l2,tc,K", [(120, 250, 500, 0, 5), (7, 16, 24, 0, 5), (33, 65, 31, 0, 5),
                                          (1100, 64, 96, 0, 3),      # split-K over the batch (fp32 SIMT)
                                          (300, 64, 96, 1, 3),       # TF32 tensor cores, one partial M tile, N < 128
                                          (1024, 250, 500, 1, 2)])   # TF32 tensor cores at the reference's widths (K = 250: ragged k-block)
def test_update_parity_vs_oracle(sb, O, B, l1, l2, tc, K):
    """replay() on a caller-supplied minibatch vs the CPU oracle.  tc = 1 runs the 250x500-type contractions in TF32
    (10-bit significands, fp32 accumulate): stated tolerance 1e-2 of the layer's largest gradient / 2e-3 relative on the
    losses, against 2e-4 / 1e-4 for the fp32 path."""
    rng = np.random.default_rng(B)
    kw = dict(batch=B, l1=l1, l2=l2)
    orc = O.OracleDdpg(O.default_ddpg_params(**kw))
    orc.init(5)
    le = sb.Learner(params=sb.default_ddpg_params(use_tensor_cores=tc, **kw))
    for net in (0, 1):  # non-zero biases, targets different from the models
        for k in range(3):
            w, b = orc.get_layer(net, k)
            b = rng.normal(0, 0.05, b.shape).astype(np.float32)
            orc.set_layer(net, k, w, b)
            orc.set_layer(net + 2, k, w * np.float32(0.9), b * np.float32(1.1))
    _sync_nets(le, orc)
    s_min = rng.uniform(-1, 0, 9).astype(np.float32)
    s_max = (s_min + rng.uniform(0.5, 3, 9)).astype(np.float32)
    s_max[5] = s_min[5]
    orc.set_norm(s_min, s_max)
    le.set_norm(s_min, s_max)
    g_rtol, g_atol, l_rel = (2e-4, 2e-6, 1e-4) if not tc else (0.0, 1e-2, 2e-3)
    worst = 0.0
    for step in range(K):
        s = rng.uniform(-1, 3, (9, B)).astype(np.float32)
        a = rng.uniform(-1, 1, (2, B)).astype(np.float32)
        r = rng.uniform(-5, 1, B).astype(np.float32)
        s2 = rng.uniform(-1, 3, (9, B)).astype(np.float32)
        if tc:  # the constant column (p_buy) sits at its constant: TF32 cannot carry the 1e8-scale inputs the fp32 cases stress
            s[5] = s_min[5]; s2[5] = s_min[5]
        orc.update_batch(s, a, r, s2)
        le.update_batch(dev(s), dev(a), dev(r), dev(s2))
        if step == 0:
            for net in (0, 1):
                for k in range(3):
                    gw, gb = le.get_grad(net, k)
                    ow, ob = orc.get_grad(net, k)
                    sc = max(np.abs(ow).max(), 1e-12)
                    worst = max(worst, np.abs(gw - ow).max() / sc, np.abs(gb - ob).max() / max(np.abs(ob).max(), 1e-12))
                    np.testing.assert_allclose(gw, ow, rtol=g_rtol, atol=g_atol * sc)
                    np.testing.assert_allclose(gb, ob, rtol=g_rtol, atol=g_atol * max(np.abs(ob).max(), 1e-12))
        lc, la = le.losses()
        olc, ola = orc.losses()
        assert lc == pytest.approx(olc, rel=l_rel) and la == pytest.approx(ola, rel=l_rel, abs=1e-6)
    print("B=%d tc=%d: worst gradient error / layer max = %.2e" % (B, tc, worst))
    p = le.p
    travel = {0: p.lr_actor * K, 1: p.lr_critic * K, 2: p.lr_actor * K * p.tau, 3: p.lr_critic * K * p.tau}  # what ADAM can move a weight in K steps
    if not tc:
        _compare_nets(le, orc, {n: 0.02 * t + (1e-7 if n >= 2 else 0.0) for n, t in travel.items()})
    else:
        # TF32 gradients carry ~1e-3 relative noise, and ADAM's normalised step g/(|g|+eps) turns the noise of a near-zero
        # gradient into a full +-lr step: 99 % of the weights within 10 % of the possible travel, every weight within 2x of it
        for net in range(4):
            for k in range(3):
                for x, ox in zip(le.get_layer(net, k), orc.get_layer(net, k)):
                    d = np.abs(x - ox)
                    assert d.max() <= 2.0 * travel[net] + 1e-6 and np.quantile(d, 0.99) <= 0.1 * travel[net] + 1e-7, (net, k, d.max(), np.quantile(d, 0.99))


def test_update_from_replay_graph_path(sb, O, train_series):
    """ddpg_update (CUDA-graph path, on-device sampling) == oracle fed with the same transitions and the spec's indices."""
    n, T, B = 64, 72, 120
    env = sb.Shems(T, train_series, n_envs=n)
    mem = sb.Replay(n * T)
    env.reset(rng=2)
    env.rollout(sb.POLICY_RANDOM, T, seed=2, replay=mem, want_return=False)
    S, A, R, S2, D = mem.get()
    le = sb.Learner()
    le.init(9)
    orc = O.OracleDdpg(O.default_ddpg_params())
    orc.init(9)
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)
    le.set_norm(mn, mx)
    orc.set_norm(mn, mx)
    K = 4
    le.replay(mem, rng_rpl=31, n_updates=K)       # Philox indices, counter = update number
    for u in range(K):
        idx = O.sample_indices(31, u, len(mem), B)
        orc.update_batch(S[:, idx], A[:, idx], R[idx], S2[:, idx], D[idx])
    p = le.p
    atol = {0: 0.02 * p.lr_actor * K, 1: 0.02 * p.lr_critic * K, 2: 0.02 * p.lr_actor * K * p.tau + 1e-7, 3: 0.02 * p.lr_critic * K * p.tau + 1e-7}
    _compare_nets(le, orc, atol)
    lc, la = le.losses()
    olc, ola = orc.losses()
    assert lc == pytest.approx(olc, rel=1e-3) and la == pytest.approx(ola, rel=1e-3, abs=1e-6)
    # explicit host indices drive the same graph
    le2 = sb.Learner()
    le2.init(9)
    le2.set_norm(mn, mx)
    idx_all = np.stack([O.sample_indices(31, u, len(mem), B) for u in range(K)])
    le2.replay(mem, n_updates=K, idx=idx_all)
    for net in range(4):
        for k in range(3):
            np.testing.assert_array_equal(le2.get_layer(net, k)[0], le.get_layer(net, k)[0])  # deterministic: bit-identical


def test_act_parity_and_noise(sb, O):
    rng = np.random.default_rng(3)
    le = sb.Learner()
    le.init(4)
    orc = O.OracleDdpg(O.default_ddpg_params())
    orc.init(4)
    mn, mx = np.zeros(9, np.float32), rng.uniform(1, 5, 9).astype(np.float32)
    le.set_norm(mn, mx)
    orc.set_norm(mn, mx)
    n = 8192
    obs = (rng.uniform(0, 1, (9, n)) * mx[:, None]).astype(np.float32)
    a, sc = le.act(dev(obs), train=False)
    oa, osc = orc.act(obs)
    np.testing.assert_allclose(a.cpu().numpy(), oa, rtol=2e-4, atol=2e-6)
    np.testing.assert_array_equal(sc.cpu().numpy(), ((a.cpu().numpy().astype(np.float64) + 1) * 0.5).astype(np.float32))
    noise = rng.normal(0, 0.1, (2, n)).astype(np.float32)
    a2, _ = le.act(dev(obs), noise=dev(noise))
    oa2, _ = orc.act(obs, noise=noise)
    np.testing.assert_allclose(a2.cpu().numpy(), oa2, rtol=2e-4, atol=2e-6)
    # GNoise(σ=0.1) from Philox: N(0, 0.1²), clamped to [-1, 1], reproducible per (seed, step)
    a3, _ = le.act(dev(obs), train=True, sigma=0.1, rng_act=5, step=1)
    d = (a3 - a).cpu().numpy()
    assert abs(d.mean()) < 3e-3 and d.std() == pytest.approx(0.1, rel=0.03) and np.abs(a3.cpu().numpy()).max() <= 1
    a4, _ = le.act(dev(obs), train=True, sigma=0.1, rng_act=5, step=1)
    assert torch.equal(a3, a4)
    a5, _ = le.act(dev(obs), train=True, sigma=0.1, rng_act=5, step=2)
    assert not torch.equal(a3, a5)


def test_driver_short_training_run(sb, train_series):
    """populate_memory -> min_max_buffer -> a few training episodes -> rule-based and actor inference run end to end."""
    env = sb.Shems(72, train_series, n_envs=32)
    ev = sb.Shems(1439, sb.series.synth_charger98(1440, seed=7), n_envs=4)
    drv = sb.Driver(env, ev, learner=sb.Learner(), mem_size=32 * 72 * 2, ep_length=72, sigma=0.1, rng_run=1231)
    drv.learner.init(1231)
    drv.populate_memory()
    assert len(drv.memory) == 32 * 72 * 2
    mn, mx = drv.min_max_buffer()
    assert np.all(mx >= mn) and mx[5] == mn[5] == np.float32(0.4)  # p_buy is constant -> normalises to 0
    best = []
    st = drv.run_episodes(2, test_every=2, test_runs=2, on_best=lambda i, sc: best.append((i, sc)))
    tot, score = st["total_reward"], st["score_mean"]
    assert tot.shape == (2, 32) and np.isfinite(tot).all() and len(score) == 1 and np.isfinite(score).all()
    # episode 1 is evaluated (1 % test_every == 1), its score beats the initial -100000 -> the "temp" checkpoint hook fires (DDPG.jl:282-289)
    assert st["best_run"] == 1 and best == [(1, score[0])] and st["best_score"] == score[0]
    assert st["noise_mean"].shape == (2, 32) and np.abs(st["noise_mean"]).max() > 0   # sum over 72 steps of mean(N(0, 0.1) x 2)
    assert np.abs(st["noise_mean"]).max() < 72 * 0.1 * 4
    ret, trace = drv.inference(ev, 1439, track=-0.5)
    assert trace.shape == (1439, 23, 4) and torch.isfinite(ret).all()
    ret2, trace2 = drv.inference(ev, 50, track=1)
    assert trace2.shape == (50, 23, 4)
    lc, la = drv.learner.losses()
    assert np.isfinite(lc) and np.isfinite(la)


def test_closed_loop_episode_parity(sb, O, train_series):
    """episode!(train=true) end to end — act(normalize(s)) + noise -> scale_action -> step! -> remember -> replay() every step —
    on CUDA vs the CPU oracle with identical init weights, injected noise and the spec's minibatch indices (DDPG.jl:186-242)."""
    rng = np.random.default_rng(11)
    n, T, B = 8, 40, 32
    kw = dict(batch=B, l1=48, l2=64)
    P = O.params_for_charger(98)
    env = sb.Shems(72, train_series, n_envs=n)
    ref = O.OracleEnv(P, train_series, 72, n)
    mem = sb.Replay(4096)
    # warm-up transitions (random policy) so that replay() has something to sample from
    env.reset(rng=5)
    env.rollout(sb.POLICY_RANDOM, 16, seed=5, replay=mem, want_return=False)
    S, A, R, S2, D = [x.copy() for x in mem.get()]
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=3)
    le = sb.Learner(params=sb.default_ddpg_params(**kw))
    le.init(77)
    orc = O.OracleDdpg(O.default_ddpg_params(**kw))
    orc.init(77)
    le.set_norm(mn, mx)
    orc.set_norm(mn, mx)
    env.reset(rng=6)
    ref.reset(mode=2, seed=6)
    ret_gpu = torch.zeros(n, dtype=torch.float64, device="cuda")
    ret_ref = np.zeros(n)
    worst = 0.0
    for step in range(T):
        noise = rng.normal(0, 0.1, (2, n)).astype(np.float32)
        s_gpu = env.state_tensor().clone()
        a, scaled = le.act(s_gpu, noise=dev(noise))
        r, s2 = env.step(scaled)
        ret_gpu += r.double()
        mem.push(s_gpu, a, r, s2)
        le.replay(mem, rng_rpl=1000 + step, n_updates=1)
        # oracle side
        s_ref = ref.obs.copy()
        oa, osc = orc.act(s_ref, noise=noise)
        r_ref, s2_ref, _ = ref.step(osc)
        ret_ref += r_ref
        S = np.concatenate([S, s_ref], 1); A = np.concatenate([A, oa], 1); R = np.concatenate([R, r_ref.astype(np.float32)])
        S2 = np.concatenate([S2, s2_ref], 1); D = np.concatenate([D, np.zeros(n, np.float32)])
        idx = O.sample_indices(1000 + step, 0, S.shape[1], B)   # Philox counter = index of the update inside the call (one per replay())
        orc.update_batch(S[:, idx], A[:, idx], R[idx], S2[:, idx], D[idx])
        # Stated tolerances.  The CPU oracle itself is insensitive to rounding here (1-ulp perturbations of its initial weights
        # move its actions by < 2e-8 over these 40 steps), so the loop is compared directly: the fp32 summation order of the
        # CUDA GEMMs differs from the oracle's double accumulation, which ADAM's normalised step turns into weight
        # differences of a few percent of lr per update -> actions within 1e-6 + 1e-6 * step.
        da = np.abs(a.cpu().numpy() - oa)
        worst = max(worst, da.max() / (1e-6 + 1e-6 * step))
        assert da.max() < 1e-6 + 1e-6 * step, (step, da.max())
    print("closed loop: worst action difference / tolerance = %.3f" % worst)
    rel = np.abs(ret_gpu.cpu().numpy() - ret_ref) / np.maximum(1e-9, np.abs(ret_ref))
    assert rel.max() < 1e-4, rel
    p = le.p
    for net, lr in ((0, p.lr_actor), (1, p.lr_critic)):
        for k in range(3):
            w, b = le.get_layer(net, k)
            ow, ob = orc.get_layer(net, k)
            assert np.abs(w - ow).max() < 0.05 * lr * T, (net, k, np.abs(w - ow).max() / (lr * T))


def test_data_parallel_phases_equal_full_batch(sb, O, train_series):
    """Two emulated ranks (two learner handles on one GPU) each take half of a minibatch through ddpg_update_phase; summing
    their gradient buffers (what the NCCL all-reduce does) and applying ADAM with grad_scale = 1/2 must equal one learner
    on the full minibatch: mean over 2B samples == mean of the two B-sample means."""
    n, T, B = 64, 72, 64
    env = sb.Shems(T, train_series, n_envs=n)
    mem = sb.Replay(n * T)
    env.reset(rng=2)
    env.rollout(sb.POLICY_RANDOM, T, seed=2, replay=mem, want_return=False)
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)
    idx = np.random.default_rng(0).integers(0, len(mem), 2 * B).astype(np.int32)
    full = sb.Learner(params=sb.default_ddpg_params(batch=2 * B))
    halves = [sb.Learner(params=sb.default_ddpg_params(batch=B)) for _ in range(2)]
    for le in [full] + halves:
        le.init(3)
        le.set_norm(mn, mx)
    K = 3
    for u in range(K):
        idx_u = np.roll(idx, 7 * u)
        full.replay(mem, n_updates=1, idx=idx_u)
        nc = int(full.lib.ddpg_num_params(full._h, sb._lib.NET_CRITIC))
        g = [le.grad_tensor() for le in halves]
        for r, le in enumerate(halves):
            sb._lib.check(le.lib.ddpg_update_phase(le._h, mem._h, 0, idx_u[r * B:(r + 1) * B].ctypes.data_as(sb._lib.PI), 0, 1.0))
        tot = g[0][:nc] + g[1][:nc]            # all_reduce(sum) of the critic part
        g[0][:nc] = tot; g[1][:nc] = tot
        for le in halves:
            sb._lib.check(le.lib.ddpg_update_phase(le._h, mem._h, 1, None, 0, 0.5))
        tot = g[0][nc:] + g[1][nc:]            # all_reduce(sum) of the actor part
        g[0][nc:] = tot; g[1][nc:] = tot
        for le in halves:
            sb._lib.check(le.lib.ddpg_update_phase(le._h, mem._h, 2, None, 0, 0.5))
    p = full.p
    for net, lr in ((0, p.lr_actor), (1, p.lr_critic), (2, p.lr_actor * p.tau), (3, p.lr_critic * p.tau)):
        for k in range(3):
            w0, b0 = halves[0].get_layer(net, k)
            w1, b1 = halves[1].get_layer(net, k)
            np.testing.assert_array_equal(w0, w1)      # the ranks stay bit-identical replicas
            wf, bf = full.get_layer(net, k)
            np.testing.assert_allclose(w0, wf, rtol=1e-5, atol=0.02 * lr * K + 1e-7)
            np.testing.assert_allclose(b0, bf, rtol=1e-5, atol=0.02 * lr * K + 1e-7)


@pytest.mark.parametrize("tc", [0, 1])
def test_population_equals_independent_learners(sb, O, train_series, tc):
    """tc = 1: the population's 48x64 contractions run as one batched TF32 tensor-core launch (grid.z = learner) and are
    compared with fp32 single learners at the TF32 tolerances of test_update_parity_vs_oracle.
    A population handle (P learners advanced by the same launches, BASELINE configs[4]) must do exactly what P separate
    single-learner handles do: same init (seed + l), own replay memory, own minibatch stream, own normalisation constants.
    The batched launches pick other tile shapes than a lone learner, so sums are re-associated: weights within 2 % of
    lr·K, actions within 1e-5."""
    P, n, T, B, K = 3, 32, 40, 96, 4
    kw = dict(batch=B, l1=48, l2=64)
    mems, singles = [], []
    for l in range(P):
        env = sb.Shems(72, train_series, n_envs=n, env_id_base=1000 * l)
        mem = sb.Replay(n * T)
        env.reset(rng=20 + l)
        env.rollout(sb.POLICY_RANDOM, T, seed=20 + l, replay=mem, want_return=False)
        mems.append(mem)
    pop = sb.Learner(params=sb.default_ddpg_params(population=P, use_tensor_cores=tc, **kw))
    assert pop.population == P
    pop.init(50)
    norms = []
    for l in range(P):
        le = sb.Learner(params=sb.default_ddpg_params(**kw))
        le.init(50 + l)
        mn, mx = mems[l].min_max_buffer(n * T, rng_mm=l)
        le.set_norm(mn, mx)
        pop.select(l).set_norm(mn, mx)
        norms.append((mn, mx))
        singles.append(le)
        for net in range(4):   # identical initial weights: Philox(seed + l)
            for k in range(3):
                np.testing.assert_array_equal(pop.get_layer(net, k)[0], le.get_layer(net, k)[0])
    idx = np.stack([np.random.default_rng(l).integers(0, n * T, (K, B)) for l in range(P)]).astype(np.int32)
    pop.replay(mems, n_updates=K, idx=idx)
    for l in range(P):
        singles[l].replay(mems[l], n_updates=K, idx=idx[l])
    p = pop.p
    for l in range(P):
        pop.select(l)
        for net, lr in ((0, p.lr_actor), (1, p.lr_critic), (2, p.lr_actor * p.tau), (3, p.lr_critic * p.tau)):
            for k in range(3):
                for x, y in zip(pop.get_layer(net, k), singles[l].get_layer(net, k)):
                    if not tc:
                        np.testing.assert_allclose(x, y, rtol=1e-5, atol=0.02 * lr * K + 1e-7)
                    else:
                        d = np.abs(x - y)
                        assert d.max() <= 2.0 * lr * K + 1e-6 and np.quantile(d, 0.99) <= 0.1 * lr * K + 1e-7, (l, net, k, d.max())
        lc, la = pop.losses()
        slc, sla = singles[l].losses()
        lrel = 1e-4 if not tc else 5e-3
        assert lc == pytest.approx(slc, rel=lrel) and la == pytest.approx(sla, rel=lrel, abs=1e-6 if not tc else 1e-4)
    # Philox-sampled minibatches: learner l draws from seed + l, exactly like a lone learner with that seed
    pop.replay(mems, rng_rpl=900, n_updates=2)
    for l in range(P):
        singles[l].replay(mems[l], rng_rpl=900 + l, n_updates=2)
        pop.select(l)
        for x, y in zip(pop.get_layer(1, 1), singles[l].get_layer(1, 1)):
            if not tc:
                np.testing.assert_allclose(x, y, rtol=1e-5, atol=0.02 * p.lr_critic * (K + 2))
            else:
                assert np.quantile(np.abs(x - y), 0.99) <= 0.1 * p.lr_critic * (K + 2)
    # act(): obs [P][9][m] -> a [P][2][m]
    m = 40
    obs = torch.stack([torch.as_tensor(mems[l].get()[0][:, :m], device="cuda") for l in range(P)]).contiguous()
    noise = torch.as_tensor(np.random.default_rng(3).normal(0, 0.1, (P, 2, m)).astype(np.float32), device="cuda")
    a, sc = pop.act(obs, noise=noise)
    assert a.shape == (P, 2, m)
    for l in range(P):
        a1, sc1 = singles[l].act(obs[l].contiguous(), noise=noise[l].contiguous())
        np.testing.assert_allclose(a[l].cpu().numpy(), a1.cpu().numpy(), rtol=0, atol=1e-5 if not tc else 2e-3)
        np.testing.assert_allclose(sc[l].cpu().numpy(), sc1.cpu().numpy(), rtol=0, atol=1e-5 if not tc else 2e-3)
    # guards
    with pytest.raises(sb.ShemsError):
        sb._lib.check(pop.lib.ddpg_update(pop._h, mems[0]._h, 1, None, 0))
    with pytest.raises(sb.ShemsError):
        sb.Learner(params=sb.default_ddpg_params(population=2, batch=2048))


def test_act_ou_noise_parity(sb, O):
    """noise_type == "ou" (DDPG.jl:49-55, :157-158): the per-instance OUNoise.X recursion with Julia's Float32/Float64 mix is
    bit-exact against the oracle for injected standard normal draws over 20 steps; Philox draws are N(0,1)-distributed."""
    rng = np.random.default_rng(8)
    le = sb.Learner()
    le.init(4)
    orc = O.OracleDdpg(O.default_ddpg_params())
    orc.init(4)
    mn, mx = np.zeros(9, np.float32), rng.uniform(1, 5, 9).astype(np.float32)
    le.set_norm(mn, mx)
    orc.set_norm(mn, mx)
    n = 512
    theta, mu, sigma, dt = np.float32(0.15), np.float32(0.0), np.float32(0.1), np.float32(1e-2)
    x_gpu = torch.zeros((2, n), dtype=torch.float32, device="cuda")
    x_ref = np.zeros((2, n), np.float32)
    for step in range(20):
        obs = (rng.uniform(0, 1, (9, n)) * mx[:, None]).astype(np.float32)
        z = rng.standard_normal((2, n))
        a, sc = le.act_ou(dev(obs), x_gpu, theta, mu, sigma, dt, z=dev(z))
        noise = O.ou_noise(theta, mu, sigma, dt, x_ref, z)
        np.testing.assert_array_equal(x_gpu.cpu().numpy(), x_ref)          # the OU state itself: bit-exact
        oa, osc = orc.act(obs, noise=noise)
        np.testing.assert_allclose(a.cpu().numpy(), oa, rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(sc.cpu().numpy(), osc, rtol=2e-4, atol=2e-6)
    # device-side draws: X after one step from X = 0 is sigma*sqrt(dt)*z with z ~ N(0, 1)
    x2 = torch.zeros((2, 8192), dtype=torch.float32, device="cuda")
    obs = (rng.uniform(0, 1, (9, 8192)) * mx[:, None]).astype(np.float32)
    le.act_ou(dev(obs), x2, theta, mu, sigma, dt, rng_act=9, step=3)
    zz = x2.cpu().numpy() / (sigma * np.sqrt(dt))
    assert abs(zz.mean()) < 0.03 and abs(zz.std() - 1.0) < 0.03
    x3 = torch.zeros((2, 8192), dtype=torch.float32, device="cuda")
    le.act_ou(dev(obs), x3, theta, mu, sigma, dt, rng_act=9, step=3)
    assert torch.equal(x2, x3)                                                # reproducible per (seed, step)


@pytest.mark.parametrize("cluster_fused", [True, False])
def test_fused_peer_allreduce_world1_equals_plain_update(sb, O, train_series, cluster_fused):
    """ddpg_update_dp with a world of one rank (flags, exchange numbers, captured graph, in-kernel gradient read through the peer
    table) must be bit-identical to the plain update — on the cluster-fused small-batch path and on the tiled-GEMM sequence."""
    n, T, B, K = 64, 72, 64, 4
    env = sb.Shems(T, train_series, n_envs=n)
    mem = sb.Replay(n * T)
    env.reset(rng=2)
    env.rollout(sb.POLICY_RANDOM, T, seed=2, replay=mem, want_return=False)
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)
    plain, fused = (sb.Learner(params=sb.default_ddpg_params(batch=B)) for _ in range(2))
    for le in (plain, fused):
        assert le.set_fused(cluster_fused) is cluster_fused
        le.init(3)
        le.set_norm(mn, mx)
    with pytest.raises(sb.ShemsError):
        fused.replay_fused_dp(mem, n_updates=1)          # not connected yet
    fused.dp_connect(0, 1, [fused.dp_export()])
    fused.dp_prepare()
    plain.replay(mem, rng_rpl=5, n_updates=K)
    fused.replay_fused_dp(mem, rng_rpl=5, n_updates=K)
    assert fused.dp_status() == 0
    for net in range(4):
        for k in range(3):
            for x, y in zip(plain.get_layer(net, k), fused.get_layer(net, k)):
                np.testing.assert_array_equal(x, y)


def test_fused_peer_exchange_gives_up_on_a_missing_peer(sb, train_series, monkeypatch):
    """A rank whose peer never launches its exchange kernel: the bounded spin ends (SHEMS_DP_TIMEOUT_MS, 4 s by default), NO block
    applies the optimiser step (nets, targets untouched), ddpg_dp_status reports 1 and ddpg_sync fails instead of hanging the GPU."""
    monkeypatch.setenv("SHEMS_DP_TIMEOUT_MS", "30")
    n, T, B = 64, 72, 64
    env = sb.Shems(T, train_series, n_envs=n)
    mem = sb.Replay(n * T)
    env.reset(rng=2)
    env.rollout(sb.POLICY_RANDOM, T, seed=2, replay=mem, want_return=False)
    mn, mx = mem.min_max_buffer(len(mem), rng_mm=1)
    here, absent = (sb.Learner(params=sb.default_ddpg_params(batch=B)) for _ in range(2))
    for le in (here, absent):
        le.init(3)
        le.set_norm(mn, mx)
    blobs = [here.dp_export(), absent.dp_export()]
    here.dp_connect(0, 2, blobs)
    here.dp_prepare()
    before = [here.get_layer(net, k) for net in range(4) for k in range(3)]
    here.replay_fused_dp(mem, rng_rpl=5, n_updates=1)
    torch.cuda.synchronize()
    assert here.dp_status() == 1
    with pytest.raises(sb.ShemsError):
        here.sync()
    after = [here.get_layer(net, k) for net in range(4) for k in range(3)]
    for (w0, b0), (w1, b1) in zip(before, after):
        np.testing.assert_array_equal(w0, w1)
        np.testing.assert_array_equal(b0, b1)


@pytest.mark.parametrize("batch,tc", [(32, 0), (512, 1), (2048, 1)])
def test_fused_peer_allreduce_two_processes(batch, tc):
    """(512, 1): the large-batch tensor-core path under the fused exchange; (2048, 1): with the forward / backward chain kernels and their
    programmatic dependent launches behind the exchange kernels.  Two ranks as two processes (torch.distributed.run, gloo for the handle exchange; CUDA IPC for the gradients): replicas
    bit-identical and equal to one learner on the full minibatch — checked inside tests/dp_worker.py.  On a one-GPU box both
    ranks share the device (time-sliced contexts): the same IPC code path, just slower."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "dp_worker.py")]
    env = dict(os.environ, DP_BATCH=str(batch), DP_TC=str(tc))
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0 and "DP_OK world=2" in out.stdout and ("batch=%d tc=%d" % (batch, tc)) in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])


def test_population_driver_trains_end_to_end(sb, train_series):
    """configs[4] in miniature: 4 learners (two chargers x two seeds) act, step their own environments, remember and replay()
    together; returns and losses stay finite and every learner's critic actually moves."""
    drv = sb.PopulationDriver(train_series, chargers=[98, 98, 4, 4], seeds=[11, 12, 13, 14], n_envs=32, mem_size=32 * 72, batch=64, l1=48, l2=64,
                              use_tensor_cores=1)
    drv.populate_memory()
    drv.min_max_buffer()
    w0 = [drv.learner.select(l).get_layer(1, 1)[0].copy() for l in range(4)]
    r1 = drv.episode(train=True, rng_ep=1)
    r2 = drv.episode(train=False, rng_ep=2)
    assert r1.shape == (4,) and torch.isfinite(r1).all() and torch.isfinite(r2).all()
    for l in range(4):
        lc, la = drv.learner.select(l).losses()
        assert np.isfinite(lc) and np.isfinite(la)
        assert np.abs(drv.learner.get_layer(1, 1)[0] - w0[l]).max() > 1e-4
    # different chargers see different physics: learner 0 (Charger98, 6.75 kWh battery) vs learner 2 (charger 4, 9.9 kWh)
    assert drv.env.group_params[0].b_soc_max != drv.env.group_params[1].b_soc_max and drv.env.n_envs == 4 * 32
    rb = drv.evaluate_rule_based()
    assert rb.shape == (4,) and np.isfinite(rb).all()


def test_act_soa_and_push_groups_equal_packed_layout(sb, train_series):
    """ddpg_act_soa / replay_push_groups work on ONE structure-of-arrays over all P*n instances (what an environment handle with
    instance groups holds); they must equal the packed [P][k][n] calls bit for bit."""
    from shems_b200.ddpg import push_groups
    P, n = 3, 50
    N = P * n
    rng = np.random.default_rng(4)
    pop = sb.Learner(params=sb.default_ddpg_params(population=P, batch=32, l1=48, l2=64))
    pop.init(8)
    for l in range(P):
        pop.select(l).set_norm(np.zeros(9, np.float32), rng.uniform(1, 4, 9).astype(np.float32))
    obs_soa = torch.as_tensor(rng.uniform(0, 3, (9, N)).astype(np.float32), device="cuda")
    noise_soa = torch.as_tensor(rng.normal(0, 0.1, (2, N)).astype(np.float32), device="cuda")
    packed = lambda x, k: x.view(k, P, n).permute(1, 0, 2).contiguous()          # [k][N] -> [P][k][n]
    a1, s1 = pop.act(obs_soa, noise=noise_soa, soa=True)
    a2, s2 = pop.act(packed(obs_soa, 9), noise=packed(noise_soa, 2))
    assert a1.shape == (2, N) and torch.equal(packed(a1, 2), a2) and torch.equal(packed(s1, 2), s2)
    a3, _ = pop.act(obs_soa, train=True, sigma=0.1, rng_act=5, step=2, env_id_base=100, soa=True)   # Philox noise keyed by the global env id
    a4, _ = pop.act(packed(obs_soa, 9), train=True, sigma=0.1, rng_act=5, step=2, env_id_base=100)
    assert torch.equal(packed(a3, 2), a4) and not torch.equal(a3, a1)
    # remember(): one launch for all learners vs one push per learner
    mems_a = [sb.Replay(120) for _ in range(P)]
    mems_b = [sb.Replay(120) for _ in range(P)]
    for rep in range(3):                                                          # 3 x 50 transitions into 120 slots: the rings wrap
        r = torch.as_tensor(rng.uniform(-3, 1, N).astype(np.float32), device="cuda")
        s_next = torch.as_tensor(rng.uniform(0, 3, (9, N)).astype(np.float32), device="cuda")
        push_groups(mems_a, obs_soa, a1, r, s_next)
        for l in range(P):
            sl = slice(l * n, (l + 1) * n)
            mems_b[l].push(obs_soa[:, sl].contiguous(), a1[:, sl].contiguous(), r[sl].contiguous(), s_next[:, sl].contiguous())
        obs_soa = s_next
    for l in range(P):
        assert len(mems_a[l]) == len(mems_b[l]) == 120
        for x, y in zip(mems_a[l].get(), mems_b[l].get()):
            np.testing.assert_array_equal(x, y)
    with pytest.raises(sb.ShemsError):                                           # more transitions per learner than a ring holds
        push_groups([sb.Replay(10) for _ in range(P)], obs_soa, a1, r, s_next)


@pytest.mark.parametrize("population,noise", [(1, "gn"), (3, "gn"), (1, "ou")])
def test_native_episode_equals_python_loop(sb, train_series, population, noise):
    """ddpg_episode (episode! enqueued by one native call) must do exactly what the step-by-step Python loop does — same seeds,
    same kernels: returns and every weight bit-identical, for one learner and for a population."""
    kw = dict(batch=32, l1=48, l2=64)
    outs = []
    for native in (False, True):
        if population == 1:
            env = sb.Shems(72, train_series, n_envs=16)
            drv = sb.Driver(env, None, learner=sb.Learner(params=sb.default_ddpg_params(**kw)), mem_size=16 * 72, ep_length=72, sigma=0.1,
                            rng_run=77, native=native, noise_type=noise)
            drv.learner.init(5)
            drv.populate_memory()
            drv.min_max_buffer()
            rets = [drv.episode(env, train=True, rng_ep=3)[0], drv.episode(env, train=False, rng_ep=4)[0]]
            le, P = drv.learner, 1
        else:
            drv = sb.PopulationDriver(train_series, chargers=[98, 98, 4], seeds=[11, 12, 13], n_envs=16, mem_size=16 * 72, use_tensor_cores=0,
                                      native=native, **kw)
            drv.populate_memory()
            drv.min_max_buffer()
            rets = [drv.episode(train=True, rng_ep=3), drv.episode(train=False, rng_ep=4)]
            le, P = drv.learner, 3
        ws = []
        for l in range(P):
            le.select(l)
            ws += [le.get_layer(net, k) for net in range(4) for k in range(3)]
        outs.append((rets, ws))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert torch.equal(a, b)
    for (w0, b0), (w1, b1) in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(w0, w1)
        np.testing.assert_array_equal(b0, b1)
    assert float(outs[0][0][0].abs().sum()) > 0


def test_full_size_properties_ddpg(sb, train_series):
    """BASELINE configs[3] shape (B = 8192, 250/500, TF32 tensor cores) and configs[4] shape (80 learners, B = 120) through
    size-independent properties: zero learning rates leave the models untouched and tau = 1 makes the targets equal to them;
    the update is deterministic (two identical learners stay bit-identical); losses are finite."""
    env = sb.Shems(72, train_series, n_envs=8192)
    mem = sb.Replay(8192 * 8)
    env.reset(rng=1)
    env.rollout(sb.POLICY_RANDOM, 8, seed=1, replay=mem, want_return=False)
    mn, mx = mem.min_max_buffer(20_000, rng_mm=1)

    def make(**kw):
        le = sb.Learner(params=sb.default_ddpg_params(batch=8192, use_tensor_cores=1, **kw))
        le.init(3)
        le.set_norm(mn, mx)
        return le

    frozen = make(lr_actor=0.0, lr_critic=0.0, tau=1.0)
    before = [frozen.get_layer(net, k) for net in (0, 1) for k in range(3)]
    for net in (2, 3):   # make the targets differ from the models first
        for k in range(3):
            w, b = frozen.get_layer(net, k)
            frozen.set_layer(net, k, w * np.float32(0.5), b + np.float32(0.25))
    frozen.replay(mem, rng_rpl=4, n_updates=2)
    after = [frozen.get_layer(net, k) for net in (0, 1) for k in range(3)]
    targets = [frozen.get_layer(net, k) for net in (2, 3) for k in range(3)]
    for (w0, b0), (w1, b1), (wt, bt) in zip(before, after, targets):
        np.testing.assert_array_equal(w0, w1); np.testing.assert_array_equal(b0, b1)      # eta = 0: x - 0
        np.testing.assert_array_equal(w1, wt); np.testing.assert_array_equal(b1, bt)      # tau = 1: 0*p_t + 1*p_m
    lc, la = frozen.losses()
    assert np.isfinite(lc) and np.isfinite(la)
    g = [frozen.get_grad(net, k)[0] for net in (0, 1) for k in range(3)]
    assert all(np.isfinite(x).all() and np.abs(x).max() > 0 for x in g)
    a, b = make(), make()
    a.replay(mem, rng_rpl=9, n_updates=3)
    b.replay(mem, rng_rpl=9, n_updates=3)
    for net in range(4):
        for k in range(3):
            np.testing.assert_array_equal(a.get_layer(net, k)[0], b.get_layer(net, k)[0])
    # 80 learners: deterministic as well, and learners with different seeds end up different
    mems = [mem] * 80
    pops = []
    for _ in range(2):
        pop = sb.Learner(params=sb.default_ddpg_params(population=80, use_tensor_cores=1))
        pop.init(5)
        for l in (0, 41, 79):
            pop.select(l).set_norm(mn, mx)
        pop.replay(mems, rng_rpl=2, n_updates=2)
        pops.append(pop)
    for l in (0, 41, 79):
        x, y = pops[0].select(l).get_layer(1, 1)[0], pops[1].select(l).get_layer(1, 1)[0]
        np.testing.assert_array_equal(x, y)
    assert not np.array_equal(pops[0].select(0).get_layer(1, 1)[0], pops[0].select(79).get_layer(1, 1)[0])


def test_cuda_matches_committed_ddpg_fixture(sb):
    """CUDA replay()/act() against the committed oracle vectors of tests/golden/oracle_ddpg_small.npz (same Philox init, same
    minibatches): losses 1e-4, weights within 2 % of lr*K, actions 2e-4."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_ddpg_small.npz"))
    B, l1, l2, K, seed = 24, 20, 28, 4, 2024
    le = sb.Learner(params=sb.default_ddpg_params(batch=B, l1=l1, l2=l2))
    le.init(seed)
    le.set_norm(z["s_min"], z["s_max"])
    for u in range(K):
        le.update_batch(dev(z[f"s{u}"]), dev(z[f"a{u}"]), dev(z[f"r{u}"]), dev(z[f"s2_{u}"]))
        lc, la = le.losses()
        assert lc == pytest.approx(float(z[f"loss{u}"][0]), rel=1e-4) and la == pytest.approx(float(z[f"loss{u}"][1]), rel=1e-4, abs=1e-6)
    p = le.p
    lrs = [p.lr_actor, p.lr_critic, p.lr_actor * p.tau, p.lr_critic * p.tau]
    for net in range(4):
        for k in range(3):
            w, b = le.get_layer(net, k)
            np.testing.assert_allclose(w, z[f"w{net}{k}"], rtol=1e-5, atol=0.02 * lrs[net] * K + 1e-7)
            np.testing.assert_allclose(b, z[f"b{net}{k}"], rtol=1e-5, atol=0.02 * lrs[net] * K + 1e-7)
    a, sc = le.act(dev(z["obs"]), noise=dev(z["noise"]))
    np.testing.assert_allclose(a.cpu().numpy(), z["act"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(sc.cpu().numpy(), z["scaled"], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("B,l1,l2", [(120, 250, 500), (256, 256, 512), (64, 250, 500), (32, 48, 64), (8, 16, 24), (16, 3, 5), (40, 255, 497)])
def test_cluster_fused_update_equals_tiled_gemm_path(sb, O, B, l1, l2):
    """ddpg_set_fused: the two cluster kernels (csrc/ddpg_fused.cu) against the one-launch-per-product sequence on identical
    weights and minibatches.  Both are fp32; they differ by summation order only: gradients within 2e-5 of the layer's largest
    entry, losses within 1e-5 relative, and after K updates every weight within 2 % of what ADAM can move it (as in the
    oracle tests).  Shapes: the reference's, the largest the fused plan covers, ragged slices (255/497), widths below the
    cluster size (3/5: CTAs with empty slices)."""
    rng = np.random.default_rng(B + l1)
    kw = dict(batch=B, l1=l1, l2=l2)
    fused, tiled = sb.Learner(params=sb.default_ddpg_params(**kw)), sb.Learner(params=sb.default_ddpg_params(**kw))
    assert fused.set_fused(True) is True and tiled.set_fused(False) is False
    orc = O.OracleDdpg(O.default_ddpg_params(**kw))
    orc.init(11)
    for net in (0, 1):
        for k in range(3):
            w, b = orc.get_layer(net, k)
            b = rng.normal(0, 0.05, b.shape).astype(np.float32)
            orc.set_layer(net, k, w, b)
            orc.set_layer(net + 2, k, w * np.float32(0.9), b * np.float32(1.1))
    _sync_nets(fused, orc)
    _sync_nets(tiled, orc)
    s_min = rng.uniform(-1, 0, 9).astype(np.float32)
    s_max = (s_min + rng.uniform(0.5, 3, 9)).astype(np.float32)
    for le in (fused, tiled):
        le.set_norm(s_min, s_max)
    K = 4
    for step in range(K):
        s = rng.uniform(-1, 3, (9, B)).astype(np.float32)
        a = rng.uniform(-1, 1, (2, B)).astype(np.float32)
        r = rng.uniform(-5, 1, B).astype(np.float32)
        s2 = rng.uniform(-1, 3, (9, B)).astype(np.float32)
        d = (rng.uniform(0, 1, B) < 0.1).astype(np.float32)
        for le in (fused, tiled):
            le.update_batch(dev(s), dev(a), dev(r), dev(s2), dev(d))
        if step == 0:
            for net in (0, 1):
                for k in range(3):
                    for gf, gt in zip(fused.get_grad(net, k), tiled.get_grad(net, k)):
                        sc = max(np.abs(gt).max(), 1e-12)
                        assert np.abs(gf - gt).max() <= 2e-5 * sc, (net, k, np.abs(gf - gt).max() / sc)
        (lcf, laf), (lct, lat) = fused.losses(), tiled.losses()
        assert lcf == pytest.approx(lct, rel=1e-5) and laf == pytest.approx(lat, rel=1e-5, abs=1e-7)
    p = fused.p
    travel = {0: p.lr_actor * K, 1: p.lr_critic * K, 2: p.lr_actor * K * p.tau, 3: p.lr_critic * K * p.tau}
    for net in range(4):
        for k in range(3):
            for x, y in zip(fused.get_layer(net, k), tiled.get_layer(net, k)):
                assert np.abs(x - y).max() <= 0.02 * travel[net] + (1e-7 if net >= 2 else 0.0), (net, k, np.abs(x - y).max())
    # the fused path is deterministic: another learner fed the last minibatch twice from the same state ends bit-identical
    twins = [sb.Learner(params=sb.default_ddpg_params(**kw)) for _ in range(2)]
    for le in twins:
        _sync_nets(le, orc)
        le.set_norm(s_min, s_max)
        for _ in range(2):
            le.update_batch(dev(s), dev(a), dev(r), dev(s2), dev(d))
    for net in range(4):
        for k in range(3):
            for x, y in zip(twins[0].get_layer(net, k), twins[1].get_layer(net, k)):
                np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("n", [1, 5, 8, 37, 64])
def test_cluster_fused_act_small_n(sb, O, n):
    """act() on a handful of states (the reference's episode loop acts on one state per step, DDPG.jl:148-176) runs normalize + the
    actor's three layers as one cluster kernel: equal to the oracle within the fp32 tolerance of the large-n test, to the tiled
    path within 1e-6 absolute (tanh outputs; summation order only), the same noise stream on both paths."""
    rng = np.random.default_rng(n)
    fused, tiled = sb.Learner(), sb.Learner()
    assert fused.set_fused(True) is True and tiled.set_fused(False) is False
    orc = O.OracleDdpg(O.default_ddpg_params())
    orc.init(4)
    for net in (0,):
        for k in range(3):
            w, b = orc.get_layer(net, k)
            orc.set_layer(net, k, w, rng.normal(0, 0.05, b.shape).astype(np.float32))
    _sync_nets(fused, orc)
    _sync_nets(tiled, orc)
    mn, mx = rng.uniform(-1, 0, 9).astype(np.float32), rng.uniform(1, 5, 9).astype(np.float32)
    for le in (fused, tiled, orc):
        le.set_norm(mn, mx)
    obs = (rng.uniform(0, 1, (9, n)) * mx[:, None]).astype(np.float32)
    a, sc = fused.act(dev(obs), train=False)
    at, sct = tiled.act(dev(obs), train=False)
    oa, _ = orc.act(obs)
    np.testing.assert_allclose(a.cpu().numpy(), oa, rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(a.cpu().numpy(), at.cpu().numpy(), rtol=0, atol=1e-6)
    np.testing.assert_array_equal(sc.cpu().numpy(), ((a.cpu().numpy().astype(np.float64) + 1) * 0.5).astype(np.float32))
    a3, _ = fused.act(dev(obs), train=True, sigma=0.1, rng_act=5, step=1)
    a4, _ = tiled.act(dev(obs), train=True, sigma=0.1, rng_act=5, step=1)
    np.testing.assert_allclose(a3.cpu().numpy(), a4.cpu().numpy(), rtol=0, atol=1e-6)
    assert not torch.equal(a3, a)
    noise = rng.normal(0, 0.1, (2, n)).astype(np.float32)          # caller-supplied noise
    a5, sc5 = fused.act(dev(obs), noise=dev(noise))
    oa5, _ = orc.act(obs, noise=noise)
    np.testing.assert_allclose(a5.cpu().numpy(), oa5, rtol=2e-4, atol=2e-6)
    np.testing.assert_array_equal(sc5.cpu().numpy(), ((a5.cpu().numpy().astype(np.float64) + 1) * 0.5).astype(np.float32))
