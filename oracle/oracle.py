"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class ShemsParams(C.Structure):
    _fields_ = [
        ("pv_eta", C.c_float), ("b_eta", C.c_float), ("b_soc_min", C.c_float), ("b_soc_max", C.c_float),
        ("b_rate_max", C.c_double), ("b_loss", C.c_float), ("ev_soc_min", C.c_float), ("ev_soc_max", C.c_float),
        ("ev_rate_max", C.c_float), ("penalty_weight", C.c_float), ("sell_discount", C.c_double),
        ("discomfort_weight_ev", C.c_double), ("disc_pot", C.c_double),
        ("penalty_weight_f64", C.c_double), ("penalty_in_f64", C.c_int32), ("reward_form", C.c_int32),
    ]


class DdpgParams(C.Structure):
    _fields_ = [
        ("state_size", C.c_int32), ("action_size", C.c_int32), ("l1", C.c_int32), ("l2", C.c_int32),
        ("batch", C.c_int32), ("gamma", C.c_float), ("tau", C.c_float), ("lr_actor", C.c_float),
        ("lr_critic", C.c_float), ("adam_beta1", C.c_double), ("adam_beta2", C.c_double), ("adam_eps", C.c_double),
        ("act_lo", C.c_float * 2), ("act_hi", C.c_float * 2), ("use_tensor_cores", C.c_int32), ("population", C.c_int32),
    ]


def default_ddpg_params(batch=120, l1=250, l2=500, gamma=0.99, tau=1e-3, lr_actor=1e-4, lr_critic=1e-3):
    p = DdpgParams()
    p.state_size, p.action_size, p.l1, p.l2, p.batch = 9, 2, l1, l2, batch
    p.gamma, p.tau, p.lr_actor, p.lr_critic = gamma, tau, lr_actor, lr_critic
    p.adam_beta1, p.adam_beta2, p.adam_eps = 0.9, 0.999, 1e-8
    p.act_lo[0] = p.act_lo[1] = 0.0
    p.act_hi[0] = p.act_hi[1] = 1.0
    p.use_tensor_cores = 0
    return p


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("shems_oracle.c", "ddpg_oracle.c", "oracle.h", "ddpg_oracle.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "shems_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        PF, PD, PI = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
        PP = C.POINTER(ShemsParams)
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_params_for_charger.argtypes = [C.c_int, PP]
        L.oracle_params_for_env.argtypes = [C.c_int, C.c_int, PP]
        L.oracle_params_for_env.restype = C.c_int
        L.oracle_action_drl.argtypes = [PP, PF, C.c_float, C.c_float, PF]
        L.oracle_action_drl.restype = None
        L.oracle_action_rule.argtypes = [PP, PF, PF]
        L.oracle_action_rule.restype = None
        L.oracle_step.argtypes = [PP, PF, C.c_int, PF, PI, PF, C.c_double, PD, PD]
        L.oracle_reset.argtypes = [PP, PF, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, PF, PI]
        L.oracle_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oracle_philox.restype = None
        L.oracle_reset_draws.argtypes = [PP, C.c_int, C.c_int, C.c_uint64, C.c_uint64, PI, PF]
        L.oracle_reset_draws.restype = None
        L.oracle_random_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, PF]
        L.oracle_random_action.restype = None
        L.oracle_scale_action.argtypes = [PF, PF, PF, PF]
        L.oracle_scale_action.restype = None
        L.oracle_rollout.argtypes = [PP, PF, C.c_int, C.c_longlong, PF, PI, C.c_int, C.c_int, C.c_uint64, C.c_longlong,
                                     PF, PD, PF, PF, PF, PF, PD, C.c_int]
        L.oracle_step_batch.argtypes = [PP, PF, C.c_int, C.c_longlong, PF, PI, PF, C.c_double, PD, PD]
        L.oracle_reset_batch.argtypes = [PP, PF, C.c_int, C.c_int, C.c_longlong, C.c_int, PI, PF, C.c_uint64,
                                         C.c_longlong, PF, PI]
        L.oracle_action_batch.argtypes = [PP, C.c_longlong, PF, PF, PF]
        L.oracle_action_batch.restype = None
        PDP = C.POINTER(DdpgParams)
        L.oracle_ddpg_create.argtypes = [PDP]
        L.oracle_ddpg_create.restype = C.c_void_p
        L.oracle_ddpg_destroy.argtypes = [C.c_void_p]
        L.oracle_ddpg_destroy.restype = None
        for f in (L.oracle_ddpg_set_layer, L.oracle_ddpg_get_layer, L.oracle_ddpg_get_grad):
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, PF, PF]
            f.restype = None
        L.oracle_ddpg_set_norm.argtypes = [C.c_void_p, PF, PF]
        L.oracle_ddpg_set_norm.restype = None
        L.oracle_ddpg_get_losses.argtypes = [C.c_void_p, PF, PF]
        L.oracle_ddpg_get_losses.restype = None
        L.oracle_ddpg_init.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_ddpg_init.restype = None
        L.oracle_ddpg_act.argtypes = [C.c_void_p, PF, C.c_int, PF, PF, PF]
        L.oracle_ddpg_act.restype = None
        L.oracle_ou_noise.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, PF, C.POINTER(C.c_double), C.c_int, C.c_int, PF]
        L.oracle_ou_noise.restype = None
        L.oracle_ddpg_update_batch.argtypes = [C.c_void_p, PF, PF, PF, PF, PF]
        L.oracle_ddpg_update_batch.restype = None
        L.oracle_sample_indices.argtypes = [C.c_uint64, C.c_uint32, C.c_longlong, C.c_int, PI]
        L.oracle_sample_indices.restype = None
        _LIB = L
    return _LIB


def set_threads(n):
    return lib().oracle_set_threads(int(n))


def params_for_charger(cid=98):
    p = ShemsParams()
    st = lib().oracle_params_for_charger(cid, C.byref(p))
    if st != 0:
        raise KeyError(cid)
    return p


def params_for_env(variant, cid=98):
    """module constants of a sibling env file: 0 shems_LU1, 1 shems_LU7, 2 shems_LU1_input0607"""
    p = ShemsParams()
    st = lib().oracle_params_for_env(int(variant), int(cid), C.byref(p))
    if st != 0:
        raise KeyError(cid)
    return p


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleEnv:
    """N instances of the reference `Shems` env, stepped by the oracle (SoA arrays like the ABI)."""

    def __init__(self, params, series, maxsteps, n_envs):
        self.P = params
        self.series = f32(series)
        assert self.series.ndim == 2 and self.series.shape[0] == 8
        self.nrows = self.series.shape[1]
        self.maxsteps = int(maxsteps)
        self.n = int(n_envs)
        self.obs = np.zeros((9, self.n), np.float32)
        self.idx = np.ones(self.n, np.int32)
        self.step_count = 0

    def reset(self, mode=0, idx0=None, socb0=None, seed=0, env_id_base=0):
        i0 = np.ascontiguousarray(idx0, np.int32) if idx0 is not None else None
        s0 = f32(socb0) if socb0 is not None else None
        st = lib().oracle_reset_batch(C.byref(self.P), _fp(self.series), self.nrows, self.maxsteps, self.n, mode,
                                      _ip(i0), _fp(s0), seed, env_id_base, _fp(self.obs), _ip(self.idx))
        if st != 0:
            raise ValueError(f"oracle_reset status {st}")
        self.step_count = 0
        return self.obs

    def step(self, act, track=0.0, want_trace=False):
        act = f32(act)
        reward = np.zeros(self.n, np.float64)
        trace = np.zeros((23, self.n), np.float64) if want_trace else None
        st = lib().oracle_step_batch(C.byref(self.P), _fp(self.series), self.nrows, self.n, _fp(self.obs), _ip(self.idx),
                                     _fp(act), float(track), _dp(reward), _dp(trace))
        if st != 0:
            raise IndexError(f"oracle_step status {st}")
        self.step_count += 1
        return reward, self.obs, trace

    def action(self, target=None):
        out = np.zeros((2, self.n), np.float32)
        t = f32(target) if target is not None else None
        lib().oracle_action_batch(C.byref(self.P), self.n, _fp(self.obs), _fp(t), _fp(out))
        return out

    def rollout(self, policy, T, seed=0, env_id_base=0, tape=None, want_transitions=False, want_trace=False, step0=0):
        n = self.n
        tape = f32(tape) if tape is not None else None
        ret = np.zeros(n, np.float64)
        trs = tra = trr = trs2 = None
        if want_transitions:
            trs = np.zeros((T, 9, n), np.float32)
            tra = np.zeros((T, 2, n), np.float32)
            trr = np.zeros((T, n), np.float32)
            trs2 = np.zeros((T, 9, n), np.float32)
        trace = np.zeros((T, 23, n), np.float64) if want_trace else None
        st = lib().oracle_rollout(C.byref(self.P), _fp(self.series), self.nrows, n, _fp(self.obs), _ip(self.idx), policy, T,
                                  seed, env_id_base, _fp(tape), _dp(ret), _fp(trs), _fp(tra), _fp(trr), _fp(trs2),
                                  _dp(trace), step0)
        if st != 0:
            raise IndexError(f"oracle_rollout status {st}")
        self.step_count += T
        return dict(ep_return=ret, s=trs, a=tra, r=trr, s2=trs2, trace=trace)


def step_single(params, series, state, idx, a, track=0.0):
    """One reference step!; returns (reward, new_state[9], new_idx, trace[23])."""
    series = f32(series)
    s = f32(np.array(state, dtype=np.float32).copy())
    i = C.c_int32(idx)
    r = C.c_double(0)
    tr = np.zeros(23, np.float64)
    a = f32(a)
    st = lib().oracle_step(C.byref(params), _fp(series), series.shape[1], _fp(s), C.byref(i), _fp(a), float(track),
                           C.byref(r), _dp(tr))
    if st != 0:
        raise IndexError(f"oracle_step status {st}")
    return r.value, s, i.value, tr


class OracleDdpg:
    NETS = dict(actor=0, critic=1, actor_target=2, critic_target=3)

    def __init__(self, p):
        self.p = p
        self.h = lib().oracle_ddpg_create(C.byref(p))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:   # module globals may be gone at interpreter exit
            lib().oracle_ddpg_destroy(self.h)
            self.h = None

    def layer_shape(self, net, layer):
        S, A = self.p.state_size, self.p.action_size
        critic = net in (1, 3)
        dims = [(S + A if critic else S), self.p.l1, self.p.l2, (1 if critic else A)]
        return dims[layer], dims[layer + 1]  # (in, out)

    def set_layer(self, net, layer, w, b):
        lib().oracle_ddpg_set_layer(self.h, net, layer, _fp(f32(w)), _fp(f32(b)))

    def _get(self, fn, net, layer):
        i, o = self.layer_shape(net, layer)
        w = np.zeros(i * o, np.float32)
        b = np.zeros(o, np.float32)
        fn(self.h, net, layer, _fp(w), _fp(b))
        return w, b

    def get_layer(self, net, layer):
        return self._get(lib().oracle_ddpg_get_layer, net, layer)

    def get_grad(self, net, layer):
        return self._get(lib().oracle_ddpg_get_grad, net, layer)

    def set_norm(self, s_min, s_max):
        lib().oracle_ddpg_set_norm(self.h, _fp(f32(s_min)), _fp(f32(s_max)))

    def init(self, seed):
        lib().oracle_ddpg_init(self.h, seed)

    def act(self, obs, noise=None):
        obs = f32(obs)
        n = obs.shape[1]
        a = np.zeros((2, n), np.float32)
        sc = np.zeros((2, n), np.float32)
        nz = f32(noise) if noise is not None else None
        lib().oracle_ddpg_act(self.h, _fp(obs), n, _fp(nz), _fp(a), _fp(sc))
        return a, sc

    def update_batch(self, s, a, r, s2, done=None):
        d = f32(done) if done is not None else None
        lib().oracle_ddpg_update_batch(self.h, _fp(f32(s)), _fp(f32(a)), _fp(f32(r)), _fp(f32(s2)), _fp(d))

    def losses(self):
        lc, la = C.c_float(0), C.c_float(0)
        lib().oracle_ddpg_get_losses(self.h, C.byref(lc), C.byref(la))
        return lc.value, la.value


def ou_noise(theta, mu, sigma, dt, ou_x, z):
    """sample_noise(ou::OUNoise) (DDPG.jl:49-55) per instance: advances ou_x [2][n] float32 in place with the standard normal
    draws z [2][n] float64, returns the noise Float32.(X)."""
    assert ou_x.dtype == np.float32 and ou_x.flags.c_contiguous
    z = np.ascontiguousarray(z, np.float64)
    A, n = ou_x.shape
    out = np.zeros((A, n), np.float32)
    lib().oracle_ou_noise(theta, mu, sigma, dt, _fp(ou_x), _dp(z), n, A, _fp(out))
    return out


def sample_indices(seed, update, length, batch):
    out = np.zeros(batch, np.int32)
    lib().oracle_sample_indices(seed, update, length, batch, _ip(out))
    return out
