"""Times replay() at a large batch (BASELINE configs[3]) with the layer-2 products on the TF32 tcgen05 kernels, with and without the
fused forward-chain kernel (SHEMS_TC_CHAIN).  usage: python tools/time_ddpg_large.py [batch] [n_updates]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_updates = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ser = sb.series.synth_charger98(4320, seed=98)
env = sb.Shems(72, ser, n_envs=8192)
mem = sb.Replay(1 << 20)
env.reset(rng=1)
env.rollout(sb.POLICY_RANDOM, 72, seed=1, replay=mem, want_return=False)
mn, mx = mem.min_max_buffer(24_000, rng_mm=1)
out = {}
for chain in ("1", "0"):
    os.environ["SHEMS_TC_CHAIN"] = chain
    le = sb.Learner(params=sb.default_ddpg_params(batch=B, use_tensor_cores=1))
    le.init(1)
    le.set_norm(mn, mx)
    le.replay(mem, rng_rpl=1, n_updates=10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    le.replay(mem, rng_rpl=2, n_updates=n_updates)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["chain" if chain == "1" else "layerwise"] = dict(us_per_update=1e3 * ms / n_updates, tflops=10 * 256_500 * B * n_updates / (ms * 1e-3) / 1e12)
    le.close()
print(json.dumps(dict(batch=B, **out)))
if "fctrace" in os.environ.get("SHEMS_B200_LIB", ""):   # a -DFC_TRACE build: timeline of CTA 0 of the last forward-chain launch
    import ctypes
    os.environ["SHEMS_TC_CHAIN"] = "1"
    le = sb.Learner(params=sb.default_ddpg_params(batch=B, use_tensor_cores=1))
    le.init(1)
    le.set_norm(mn, mx)
    le.replay(mem, rng_rpl=1, n_updates=3)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 96)()
    sb._lib.lib().tc_fwd_chain_trace_read.argtypes = [ctypes.c_void_p]
    sb._lib.lib().tc_fwd_chain_trace_read(buf)
    t0 = buf[72]
    rel = [int(x) - int(t0) for x in buf]
    print("TMA slab issued (ns):", rel[0:16])
    print("MMA slab committed:", rel[32:48])
    print("layer-1 k-block written:", rel[64:72])
    print("accumulators seen %d, warp 2's chunks done %d, end %d" % (rel[73], rel[74], rel[76]))
