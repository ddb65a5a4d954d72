"""CPU tests of the DDPG oracle: one whole replay() against an independent torch autograd + Adam restatement."""
import numpy as np
import pytest
import torch


def make_batch(rng, B):
    s = rng.uniform(-1, 3, (9, B)).astype(np.float32)
    a = rng.uniform(-1, 1, (2, B)).astype(np.float32)
    r = rng.uniform(-5, 1, B).astype(np.float32)
    s2 = rng.uniform(-1, 3, (9, B)).astype(np.float32)
    return s, a, r, s2


def torch_nets(orc, h, dtype=torch.float64):
    nets = {}
    for net in range(4):
        layers = []
        for k in range(3):
            i, o = h.layer_shape(net, k)
            w, b = h.get_layer(net, k)
            W = torch.tensor(w.reshape(i, o).T.copy(), dtype=dtype, requires_grad=net < 2)  # Flux weight out×in
            bb = torch.tensor(b, dtype=dtype, requires_grad=net < 2)
            layers.append((W, bb))
        nets[net] = layers
    return nets


def fwd(layers, x, last):
    h1 = torch.relu(layers[0][0] @ x + layers[0][1][:, None])
    h2 = torch.relu(layers[1][0] @ h1 + layers[1][1][:, None])
    y = layers[2][0] @ h2 + layers[2][1][:, None]
    return torch.tanh(y) if last == "tanh" else y


@pytest.mark.parametrize("B,l1,l2", [(120, 250, 500), (7, 16, 24)])
def test_update_matches_torch_autograd_adam(O, B, l1, l2):
    rng = np.random.default_rng(3)
    p = O.default_ddpg_params(batch=B, l1=l1, l2=l2)
    h = O.OracleDdpg(p)
    h.init(42)
    s_min = rng.uniform(-1, 0, 9).astype(np.float32)
    s_max = (s_min + rng.uniform(0.5, 3, 9)).astype(np.float32)
    s_max[5] = s_min[5]  # a constant column (p_buy) normalises to 0 (den = 1e-8)
    h.set_norm(s_min, s_max)
    # biases are zero after init: perturb them so the bias path is exercised
    for net in (0, 1):
        for k in range(3):
            w, b = h.get_layer(net, k)
            h.set_layer(net, k, w, rng.normal(0, 0.05, b.shape).astype(np.float32))
            h.set_layer(net + 2, k, w * np.float32(0.9), b)
    n_steps = 3
    T = torch_nets(O, h)
    params = [[t for layer in T[n] for t in layer] for n in (0, 1)]
    opt_a = torch.optim.Adam(params[0], lr=float(np.float32(p.lr_actor)), betas=(0.9, 0.999), eps=1e-8)
    opt_c = torch.optim.Adam(params[1], lr=float(np.float32(p.lr_critic)), betas=(0.9, 0.999), eps=1e-8)
    smin_t, smax_t = torch.tensor(s_min, dtype=torch.float64), torch.tensor(s_max, dtype=torch.float64)
    norm = lambda x: (x - smin_t[:, None]) / ((smax_t - smin_t)[:, None] + float(np.float32(1e-8)))
    gamma, tau = float(np.float32(p.gamma)), float(np.float32(p.tau))
    for step in range(n_steps):
        s, a, r, s2 = make_batch(rng, B)
        h.update_batch(s, a, r, s2)
        ts, ta, tr_, ts2 = (torch.tensor(x, dtype=torch.float64) for x in (s, a, r, s2))
        with torch.no_grad():
            a2 = fwd(T[2], norm(ts2), "tanh")
            q2 = fwd(T[3], torch.cat([norm(ts2), a2]), "id")
            y = tr_[None, :] + gamma * q2
        opt_c.zero_grad()
        loss_c = ((fwd(T[1], torch.cat([norm(ts), ta]), "id") - y) ** 2).mean()
        loss_c.backward()
        if step == 0:
            for k in range(3):
                gw, gb = h.get_grad(1, k)
                i, o = h.layer_shape(1, k)
                np.testing.assert_allclose(gw.reshape(i, o).T, T[1][k][0].grad.numpy(), rtol=2e-4, atol=2e-7)
                np.testing.assert_allclose(gb, T[1][k][1].grad.numpy(), rtol=2e-4, atol=2e-7)
        opt_c.step()
        opt_a.zero_grad()
        opt_c.zero_grad()
        loss_a = -fwd(T[1], torch.cat([norm(ts), fwd(T[0], norm(ts), "tanh")]), "id").mean()
        loss_a.backward()
        if step == 0:
            for k in range(3):
                gw, gb = h.get_grad(0, k)
                i, o = h.layer_shape(0, k)
                np.testing.assert_allclose(gw.reshape(i, o).T, T[0][k][0].grad.numpy(), rtol=5e-4, atol=1e-8)
                np.testing.assert_allclose(gb, T[0][k][1].grad.numpy(), rtol=5e-4, atol=1e-8)
        opt_a.step()
        with torch.no_grad():
            for tn, mn in ((2, 0), (3, 1)):
                for (Wt, bt), (Wm, bm) in zip(T[tn], T[mn]):
                    Wt.mul_(1 - tau).add_(tau * Wm)
                    bt.mul_(1 - tau).add_(tau * bm)
        lc, la = h.losses()
        assert lc == pytest.approx(float(loss_c), rel=1e-4) and la == pytest.approx(float(loss_a), rel=1e-4, abs=1e-6)
    # parameters after 3 updates: Adam's first steps move every weight by ~lr, so compare against lr
    for net, lr in ((0, p.lr_actor), (1, p.lr_critic), (2, p.lr_actor * p.tau), (3, p.lr_critic * p.tau)):
        for k in range(3):
            w, b = h.get_layer(net, k)
            i, o = h.layer_shape(net, k)
            np.testing.assert_allclose(w.reshape(i, o).T, T[net][k][0].detach().numpy(), rtol=1e-5, atol=0.02 * lr * n_steps + 1e-7)
            np.testing.assert_allclose(b, T[net][k][1].detach().numpy(), rtol=1e-5, atol=0.02 * lr * n_steps + 1e-7)


def test_act_and_scale(O):
    rng = np.random.default_rng(5)
    p = O.default_ddpg_params(batch=8, l1=32, l2=48)
    h = O.OracleDdpg(p)
    h.init(1)
    h.set_norm(np.zeros(9, np.float32), np.full(9, 2, np.float32))
    obs = rng.uniform(0, 2, (9, 33)).astype(np.float32)
    a, sc = h.act(obs)
    assert a.shape == (2, 33) and np.all(np.abs(a) <= 1)
    np.testing.assert_array_equal(sc, ((a.astype(np.float64) + 1.0) * 0.5).astype(np.float32))  # scale_action with bounds (0,0)/(1,1)
    noise = np.full((2, 33), 5.0, np.float32)
    a2, sc2 = h.act(obs, noise=noise)
    assert np.all(a2 == 1.0) and np.all(sc2 == 1.0)  # clamp(act + noise, -1, 1)
    T = torch_nets(O, h, torch.float64)
    y = fwd(T[0], torch.tensor(obs, dtype=torch.float64) / (2.0 + float(np.float32(1e-8))), "tanh").detach().numpy()
    np.testing.assert_allclose(a, y, rtol=1e-5, atol=1e-7)


def test_init_distribution(O):
    p = O.default_ddpg_params()
    h = O.OracleDdpg(p)
    h.init(7)
    for net in (0, 1):
        for k in range(3):
            i, o = h.layer_shape(net, k)
            w, b = h.get_layer(net, k)
            assert np.all(b == 0)
            if k < 2:  # glorot_uniform: (rand - 0.5) * sqrt(24/(in+out))  (DDPG.jl:21)
                lim = 0.5 * np.sqrt(24.0 / (i + o))
                assert np.abs(w).max() <= lim and np.abs(w).max() > 0.95 * lim and abs(w.mean()) < 0.05 * lim
            else:      # U(-3e-3, 3e-3)  (DDPG.jl:22)
                assert np.abs(w).max() <= 3e-3 + 1e-9
        wt, bt = h.get_layer(net + 2, 0)
        np.testing.assert_array_equal(wt, h.get_layer(net, 0)[0])  # targets are deepcopies (:38, :46)


def test_sample_indices(O):
    idx = O.sample_indices(9, 3, 24000, 120)
    assert idx.min() >= 0 and idx.max() < 24000 and len(np.unique(idx)) > 100
    np.testing.assert_array_equal(idx, O.sample_indices(9, 3, 24000, 120))
    assert not np.array_equal(idx, O.sample_indices(9, 4, 24000, 120))


def test_oracle_ou_noise_matches_numpy_typed_restatement(O):
    """sample_noise(ou::OUNoise) (DDPG.jl:49-55): Float32 drift term, Float64 diffusion term, Float32 stores — restated with numpy
    scalar types as a second opinion on the C oracle (no reference output exists for it)."""
    rng = np.random.default_rng(1)
    th, mu, sg, dt = np.float32(0.15), np.float32(0.0), np.float32(0.2), np.float32(1e-2)
    x = np.zeros((2, 50), np.float32)
    xn = x.copy()
    for step in range(30):
        z = rng.standard_normal((2, 50))
        out = O.ou_noise(th, mu, sg, dt, x, z)
        dx = (th * (mu - xn)) * dt                                    # Float32 .* Float32
        assert dx.dtype == np.float32
        dx = (dx.astype(np.float64) + np.float64(np.float32(sg * np.sqrt(dt))) * z).astype(np.float32)
        xn = (xn + dx).astype(np.float32)
        np.testing.assert_array_equal(x, xn)
        np.testing.assert_array_equal(out, xn)
    assert np.abs(x).max() < 1.0 and x.std() > 0.01


def test_oracle_matches_committed_ddpg_fixture(O):
    """tests/golden/oracle_ddpg_small.npz (oracle outputs, committed with its generator): the oracle reproduces it bit for bit."""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_ddpg_fixture", os.path.join(here, "make_ddpg_fixture.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = np.load(os.path.join(here, "oracle_ddpg_small.npz"))
    got = mod.run(O)
    assert sorted(want.files) == sorted(got.keys())
    for k in want.files:
        np.testing.assert_array_equal(want[k], got[k], err_msg=k)
