import ctypes as C, json, os, sys
sys.path.insert(0, "/root/repo")
import torch
import shems_b200 as sb
Bsz = 8192
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
lib = sb._lib.lib()
for (l1, l2) in ((252, 500), (256, 512), (252, 512), (256, 500)):
    h1 = torch.randn((Bsz, l1), device="cuda"); W2 = torch.randn((250, l2), device="cuda"); b2 = torch.randn(500, device="cuda")
    out2 = torch.empty((Bsz, l2), device="cuda")
    fn = lambda: lib.shems_tc_gemm(p(h1), l1, 0, p(W2), l2, 1, p(out2), l2, Bsz, 500, 250, 1, p(b2), None, 0, 1, None, st)
    for _ in range(5): sb._lib.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(json.dumps(dict(lda=l1, ldw=l2, us=us, tflops=2 * Bsz * 500 * 250 / us / 1e6)))
# square-ish big problem for reference: M=8192, N=512, K=2048 aligned
A = torch.randn((8192, 2048), device="cuda"); W = torch.randn((2048, 512), device="cuda"); D = torch.empty((8192, 512), device="cuda")
fn = lambda: lib.shems_tc_gemm(p(A), 2048, 0, p(W), 512, 1, p(D), 512, 8192, 512, 2048, 0, None, None, 0, 1, None, st)
for _ in range(5): sb._lib.check(fn())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(json.dumps(dict(case="8192x512x2048 MN-major B", us=us, tflops=2 * 8192 * 512 * 2048 / us / 1e6)))
Wk = torch.randn((512, 2048), device="cuda")
fn = lambda: lib.shems_tc_gemm(p(A), 2048, 0, p(Wk), 2048, 0, p(D), 512, 8192, 512, 2048, 0, None, None, 0, 1, None, st)
for _ in range(5): sb._lib.check(fn())
torch.cuda.synchronize()
e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(json.dumps(dict(case="8192x512x2048 K-major B", us=us, tflops=2 * 8192 * 512 * 2048 / us / 1e6)))
