"""Host-side mirror of the reference environment interface (Reinforce.jl protocol).

`Shems` here is N lock-stepped instances of the reference's `Shems` struct
(RL-SHEMS/RL_environments/envs/shems_LU1.jl:169-203) living on one B200; the methods keep
the reference's names and argument meaning:

    reference (Julia)                         here (Python; `!` -> trailing `_`)
    Shems(maxsteps, path)                     Shems(maxsteps, path_or_series, n_envs=1, ...)
    reset!(env; rng)            :206          reset_(env, rng=...) / env.reset(rng=...)
    step!(env, s, a; track=0)   :343          step_(env, s, a, track=0) / env.step(a, track)
    action(env, a::ShemsAction) :283          action(env, a)
    action(env, track)          :318          action(env, track)  (a negative scalar)
    finished(env, s′)           :487          finished(env, s2)
    env.state / env.a / env.reward / env.step / env.idx / env.maxsteps / env.path

All compute goes through the C ABI (include/shems_b200.h) into CUDA kernels; torch is used
only to own device buffers and the current stream.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import series as _series


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class ShemsAction:
    """ShemsAction (shems_LU1.jl:146-155): default (0.7, 1), bounds (0,0)/(1,1)."""

    def __init__(self, B=0.7, EV=1.0):
        self.B, self.EV = np.float32(B), np.float32(EV)

    def __len__(self):
        return 2

    def __iter__(self):
        return iter((self.B, self.EV))

    @staticmethod
    def minimum():
        return (np.float32(0), np.float32(0))

    @staticmethod
    def maximum():
        return (np.float32(1), np.float32(1))


class Shems:
    def __init__(self, maxsteps, path, n_envs=1, charger_id=98, device=0, params=None, env_id_base=0, groups=None):
        """groups=[(charger_id, n_instances), ...] puts several chargers into one handle (shems_create_groups): group g owns
        n_instances consecutive instances with charger g's capacities (the reference needs one process per charger); `path` may
        then also be a [G][8][nrows] array with one series per group.  n_envs / charger_id / params are ignored in that case."""
        self.lib = L.lib()
        self.maxsteps = int(maxsteps)
        self.path = path if isinstance(path, str) else "<array>"
        ser = _series.load_csv(path) if isinstance(path, str) else np.ascontiguousarray(path, dtype=np.float32)
        per_group = groups is not None and ser.ndim == 3
        if not ((ser.ndim == 2 and ser.shape[0] == 8) or (per_group and ser.shape[0] == len(groups) and ser.shape[1] == 8)):
            raise ValueError("series must be float32 [8][nrows] (or [G][8][nrows] with groups)")
        self.series = ser
        self.nrows = ser.shape[-1]
        self.device = int(device)
        self.env_id_base = int(env_id_base)
        self.a = ShemsAction()
        self.reward = None
        h = C.c_void_p()
        if groups is None:
            self.n_envs = int(n_envs)
            self.params = params if params is not None else L.params_for_charger(charger_id)
            self.groups = None
            L.check(self.lib.shems_create(C.byref(self.params), ser.ctypes.data_as(L.PF), self.nrows, self.maxsteps, self.n_envs,
                                          self.device, C.byref(h)))
        else:
            self.groups = [(int(c), int(k)) for c, k in groups]
            G = len(self.groups)
            self.group_params = [L.params_for_charger(c) for c, _ in self.groups]
            self.params = self.group_params[0]
            self.n_envs = sum(k for _, k in self.groups)
            pa = (L.ShemsParams * G)(*self.group_params)
            sz = (C.c_int64 * G)(*[k for _, k in self.groups])
            L.check(self.lib.shems_create_groups(pa, G, sz, ser.ctypes.data_as(L.PF), 1 if per_group else 0, self.nrows, self.maxsteps,
                                                 self.device, C.byref(h)))
        self._h = h
        self._torch_dev = torch.device("cuda", self.device)
        obs_p, idx_p = C.c_void_p(), C.c_void_p()
        L.check(self.lib.shems_state_ptr(self._h, C.byref(obs_p), C.byref(idx_p)))
        self._obs_ptr, self._idx_ptr = obs_p.value, idx_p.value
        self._reward_buf = torch.empty(self.n_envs, dtype=torch.float32, device=self._torch_dev)
        self._bind_stream()

    # ------------------------------------------------------------------ plumbing
    def _bind_stream(self):
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.shems_set_stream(self._h, C.c_void_p(s)))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.shems_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        L.check(self.lib.shems_sync(self._h))

    # ------------------------------------------------------------------ fields
    @property
    def state(self):
        """env.state of every instance as a float32 [9][N] host array (copy)."""
        obs = np.empty((9, self.n_envs), np.float32)
        L.check(self.lib.shems_get_state(self._h, obs.ctypes.data_as(L.PF), None))
        return obs

    @property
    def idx(self):
        obs = np.empty((9, self.n_envs), np.float32)
        idx = np.empty(self.n_envs, np.int32)
        L.check(self.lib.shems_get_state(self._h, obs.ctypes.data_as(L.PF), idx.ctypes.data_as(L.PI)))
        return idx

    @property
    def step_count(self):
        s = C.c_int32()
        L.check(self.lib.shems_get_step(self._h, C.byref(s)))
        return s.value

    def state_tensor(self):
        """Zero-copy torch view [9][N] of the device-resident state (the handle owns the memory)."""
        return _wrap_device(self._obs_ptr, (9, self.n_envs), torch.float32, self._torch_dev, self)

    def set_state(self, obs, idx):
        obs = np.ascontiguousarray(obs, np.float32)
        idx = np.ascontiguousarray(idx, np.int32)
        assert obs.shape == (9, self.n_envs) and idx.shape == (self.n_envs,)
        L.check(self.lib.shems_set_state(self._h, obs.ctypes.data_as(L.PF), idx.ctypes.data_as(L.PI)))

    # ------------------------------------------------------------------ Reinforce protocol
    def reset(self, rng=0, idx0=None, socb0=None):
        """reset!(env; rng): rng == -1 -> deterministic start; otherwise a seeded random start.
        (idx0, socb0) inject the two draws of shems_LU1.jl:224-225 for exact-parity tests."""
        self._bind_stream()
        if idx0 is not None:
            i0 = np.ascontiguousarray(idx0, np.int32)
            s0 = np.ascontiguousarray(socb0, np.float32)
            L.check(self.lib.shems_reset(self._h, L.RESET_HOST_DRAWS, i0.ctypes.data_as(L.PI), s0.ctypes.data_as(L.PF), 0, 0))
        elif rng == -1:
            L.check(self.lib.shems_reset(self._h, L.RESET_DETERMINISTIC, None, None, 0, 0))
        else:
            L.check(self.lib.shems_reset(self._h, L.RESET_DEVICE_PHILOX, None, None, int(rng) & (2**64 - 1), self.env_id_base))
        self.reward = None
        self.a = ShemsAction()
        return self

    def step(self, a, track=0, reward_out=None, obs_out=None, trace_out=None, reward64_out=None):
        """step!(env, s, a; track): `a` is a device tensor [2][N] (targets for track >= 0, (B, EV) for track < 0).
        Returns (reward tensor [N], state tensor [9][N]) — plus the trace tensor [23][N] (float64) when track != 0.
        reward64_out (float64 [N]): env.reward as the reference holds it (Float64, shems_LU1.jl:171); `reward` is its Float32 value."""
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.ascontiguousarray(a, np.float32).reshape(2, self.n_envs), device=self._torch_dev)
        assert a.dtype == torch.float32 and a.is_contiguous() and a.numel() == 2 * self.n_envs and a.is_cuda
        rw = reward_out if reward_out is not None else self._reward_buf
        if track != 0 and trace_out is None:
            trace_out = torch.empty((23, self.n_envs), dtype=torch.float64, device=self._torch_dev)
        tn = -1 if track < 0 else (1 if track > 0 else 0)
        if reward64_out is not None:
            assert reward64_out.dtype == torch.float64 and reward64_out.numel() == self.n_envs and reward64_out.is_cuda
        L.check(self.lib.shems_step(self._h, _ptr(a), tn, _ptr(rw), _ptr(reward64_out), _ptr(obs_out), _ptr(trace_out)))
        self.reward = rw
        s2 = obs_out if obs_out is not None else self.state_tensor()
        if track == 0:
            return rw, s2
        return rw, s2, trace_out

    def action(self, a_or_track=-1):
        """action(env, track) (rule-based, scalar argument) or action(env, a) (targets tensor [2][N]) -> (B, EV) [2][N]."""
        out = torch.empty((2, self.n_envs), dtype=torch.float32, device=self._torch_dev)
        if torch.is_tensor(a_or_track):
            L.check(self.lib.shems_action_drl(self._h, _ptr(a_or_track), _ptr(out)))
        else:
            L.check(self.lib.shems_action_rule(self._h, _ptr(out)))
        return out

    def finished(self, s2=None):
        f = C.c_int32()
        L.check(self.lib.shems_finished(self._h, C.byref(f)))
        return bool(f.value)

    # ------------------------------------------------------------------ fused rollouts
    def rollout(self, policy, n_steps, seed=0, tape=None, want_return=True, replay=None, want_trace=False, want_obs=False,
                want_reward=False, tape_unscaled=False):
        """T fused steps (episode!/populate_memory/inference loops).  Returns a dict of device tensors.
        tape_unscaled: the tape holds actions in [-1,1] as `remember` stores them (required with a replay sink)."""
        self._bind_stream()
        n, T = self.n_envs, int(n_steps)
        dev = self._torch_dev
        out = {}
        args = L.ShemsRolloutArgs()
        args.policy, args.n_steps, args.seed, args.env_id_base = int(policy), T, int(seed) & (2**64 - 1), self.env_id_base
        if tape is not None:
            assert tape.is_cuda and tape.dtype == torch.float32 and tape.numel() == T * 2 * n
            args.tape_dev = tape.data_ptr()
            args.tape_unscaled = 1 if tape_unscaled else 0
        if want_return:
            out["ep_return"] = torch.empty(n, dtype=torch.float64, device=dev)
            args.ep_return_dev = out["ep_return"].data_ptr()
        if replay is not None:
            args.replay = replay._h
        if want_trace:
            out["trace"] = torch.empty((T, 23, n), dtype=torch.float64, device=dev)
            args.trace_dev = out["trace"].data_ptr()
        if want_obs:
            out["obs"] = torch.empty((T, 9, n), dtype=torch.float32, device=dev)
            args.obs_traj_dev = out["obs"].data_ptr()
        if want_reward:
            out["reward"] = torch.empty((T, n), dtype=torch.float32, device=dev)
            args.reward_traj_dev = out["reward"].data_ptr()
        L.check(self.lib.shems_rollout(self._h, C.byref(args)))
        return out


class _DevView:
    """__cuda_array_interface__ shim so torch can wrap library-owned device memory without copying."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


def _wrap_device(ptr, shape, dtype, device, owner):
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.float64: "<f8"}[dtype]
    with torch.cuda.device(device):
        return torch.as_tensor(_DevView(ptr, shape, typestr, owner), device=device)


# module-level functions named like the Reinforce.jl generics (shems_LU1.jl:62-65)
def reset_(env, rng=0, **kw):
    return env.reset(rng=rng, **kw)


def step_(env, s, a, track=0, **kw):
    return env.step(a, track=track, **kw)


def action(env, a_or_track=-1):
    return env.action(a_or_track)


def finished(env, s2=None):
    return env.finished(s2)


def state(env):
    return env.state


def actions(env, s=None):
    """Reinforce.actions stub (RL_environments/Reinforce.jl:71): the action bounds."""
    return ShemsAction.minimum(), ShemsAction.maximum()
