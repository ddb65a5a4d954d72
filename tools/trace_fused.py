"""Timeline of the cluster-fused critic pass (csrc/ddpg_fused.cu): builds the -DFUSED_TRACE variant of the library, runs a few
updates at the reference's shape and prints the clock64() stamps of CTA 0 / thread 0 as microseconds since kernel start.
usage (GPU box): python tools/trace_fused.py"""
import ctypes as C
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "master-thesis-deep-reinforcement-learning-ddpg-in-home-energy-management_b200")
variant = os.path.join(PKG, "libshems_b200_trace.so")
if not os.path.exists(variant) or "--build" in sys.argv:
    spec = importlib.util.spec_from_file_location("_b", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_variant("trace", ["FUSED_TRACE"])
if "--build" in sys.argv:
    sys.exit(0)
os.environ["SHEMS_B200_LIB"] = variant
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

B = 120
ser = sb.series.synth_charger98(4320, seed=98)
env = sb.Shems(72, ser, n_envs=1000)
mem = sb.Replay(24_000)
env.reset(rng=1)
env.rollout(sb.POLICY_RANDOM, 24, seed=1, replay=mem, want_return=False)
le = sb.Learner(params=sb.default_ddpg_params(batch=B))
le.init(1)
mn, mx = mem.min_max_buffer(24_000, rng_mm=1)
le.set_norm(mn, mx)
le.replay(mem, rng_rpl=1, n_updates=200)
torch.cuda.synchronize()
out = (C.c_longlong * 64)()
lib = C.CDLL(variant)
assert lib.ddpg_fused_trace_read(out) == 0
mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1900
names = ["start", "x + small operands loaded", "f1 actor_t + W2 slot0 landed", "f2 actor_t", "f3 actor_t (+cluster wait)", "f1 critic + slot1 landed",
         "f2 critic", "f3 critic + cluster.sync", "a', q, f1 critic_t + slot0 landed", "f2 critic_t", "f3 critic_t + cluster.sync",
         "TD + b3_dz", "bx2", "cluster.sync (last)", "q/y, b3_grads, bw2", "rs_finish, bw1"]
t0 = out[0]
for i, nme in ((16, "cluster arrive"), (17, "slot 0 copies issued"), (18, "slot 1 copies issued"), (19, "x stored to smem"), (20, "small operands requested")):
    print("   prologue: %-32s at %6.2f us" % (nme, (out[i] - t0) / mhz))
prev = t0
for i, nme in enumerate(names):
    t = out[i]
    print("%2d %-40s +%6.2f us   (at %6.2f us)" % (i, nme, (t - prev) / mhz, (t - t0) / mhz))
    prev = t
