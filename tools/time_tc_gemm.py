"""TF/s of the TF32 tcgen05 GEMM on the three layer-2 shapes of a B=8192 DDPG update.  usage: python tools/time_tc_gemm.py [B]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import shems_b200 as sb  # noqa: E402

Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
h1 = torch.randn((Bsz, 252), device="cuda"); h2 = torch.randn((Bsz, 500), device="cuda"); W2 = torch.randn((250, 500), device="cuda")
b2 = torch.randn(500, device="cuda"); out2 = torch.empty((Bsz, 500), device="cuda"); out1 = torch.empty((Bsz, 252), device="cuda")
gW = torch.empty((250, 500), device="cuda"); ws = torch.empty(32 * 250 * 500, device="cuda")
cases = [
    ("fwd_L2", lambda: sb._lib.lib().shems_tc_gemm(p(h1), 252, 0, p(W2), 500, 1, p(out2), 500, Bsz, 500, 250, 1, p(b2), None, 0, 1, None, st), 2 * Bsz * 500 * 250),
    ("dX_L2", lambda: sb._lib.lib().shems_tc_gemm(p(h2), 500, 0, p(W2), 500, 0, p(out1), 252, Bsz, 250, 500, 2, None, p(h1), 252, 1, None, st), 2 * Bsz * 500 * 250),
    ("dW_L2_split16", lambda: sb._lib.lib().shems_tc_gemm(p(h1), 252, 1, p(h2), 500, 1, p(gW), 500, 250, 500, Bsz, 0, None, None, 0, 16, p(ws), st), 2 * Bsz * 500 * 250),
    ("dW_L2_split32", lambda: sb._lib.lib().shems_tc_gemm(p(h1), 252, 1, p(h2), 500, 1, p(gW), 500, 250, 500, Bsz, 0, None, None, 0, 32, p(ws), st), 2 * Bsz * 500 * 250),
]
for name, fn, flops in cases:
    for _ in range(5):
        sb._lib.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(json.dumps(dict(case=name, B=Bsz, us=us, tflops=flops / us / 1e6)))
