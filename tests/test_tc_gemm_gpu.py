"""GPU tests of the TF32 tcgen05 GEMM (csrc/tc_gemm.cu) against an fp32 torch reference.

Tolerance (stated): TF32 keeps 10 mantissa bits of each input, so |D - D_fp32| <= 3e-3 * (|A| . |B|) elementwise
(the bound on the sum of per-product truncation errors; accumulation itself is fp32 in TMEM)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def run(sb, A, a_mn, B, b_mn, M, N, K, ldd=None, epi=0, bias=None, aux=None, splits=1):
    """A/B are the stored 2-D tensors (K-major: [rows][ld>=K]; MN-major: [K][ld>=rows])."""
    ldd = N if ldd is None else ldd
    D = torch.full((M, ldd), float("nan"), device="cuda")
    ws = torch.empty(splits * M * N, device="cuda") if splits > 1 else None
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    st = sb._lib.lib().shems_tc_gemm(p(A), A.stride(0), int(a_mn), p(B), B.stride(0), int(b_mn), p(D), ldd, M, N, K, epi, p(bias), p(aux),
                                     aux.stride(0) if aux is not None else 0, splits, p(ws), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    sb._lib.check(st)
    torch.cuda.synchronize()
    return D


def logical(T, mn, rows, K):
    return (T[:K, :rows].T if mn else T[:rows, :K]).double()


@pytest.mark.parametrize("name,M,N,K,a_mn,b_mn,lda,ldb,ldd,epi,splits", [
    ("fwd_L2", 8192, 500, 250, 0, 1, 252, 500, 500, 1, 1),      # h2 = relu(h1 . W2^T + b2): A K-major, Flux Wt[in][out] MN-major
    ("dX_L2", 8192, 250, 500, 0, 0, 500, 500, 252, 2, 1),       # dz1 = (dz2 . W2) masked: both K-major
    ("dW_L2", 250, 500, 8192, 1, 1, 252, 500, 500, 0, 16),      # dW2 = h1^T . dz2: both MN-major, split-K
    ("odd", 300, 130, 70, 0, 1, 72, 132, 132, 0, 1),
    ("odd_kk", 129, 257, 33, 0, 0, 36, 36, 260, 0, 1),
    ("odd_mnmn_split", 97, 200, 1000, 1, 1, 100, 200, 200, 0, 3),
    ("tiny", 5, 3, 4, 0, 0, 4, 4, 4, 0, 1),
])
def test_tc_gemm_vs_fp32(sb, name, M, N, K, a_mn, b_mn, lda, ldb, ldd, epi, splits):
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 1000)
    A = torch.randn((K, lda) if a_mn else (M, lda), device="cuda", generator=g)
    B = torch.randn((K, ldb) if b_mn else (N, ldb), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g) if epi == 1 else None
    aux = torch.randn((M, N), device="cuda", generator=g) if epi == 2 else None
    D = run(sb, A, a_mn, B, b_mn, M, N, K, ldd=ldd, epi=epi, bias=bias, aux=aux, splits=splits)
    Al, Bl = logical(A, a_mn, M, K), logical(B, b_mn, N, K)
    ref = Al @ Bl.T
    bound = 3e-3 * (Al.abs() @ Bl.abs().T) + 1e-6
    if epi == 1:
        ref = torch.relu(ref + bias.double())
    elif epi == 2:
        ref = torch.where(aux > 0, ref, torch.zeros_like(ref))
    got = D[:, :N].double()
    assert torch.isfinite(got).all(), name
    err = (got - ref).abs()
    assert bool((err <= bound).all()), (name, float(err.max()), float((err / bound).max()))
    if ldd > N:  # padding: the TMA store clips at 16-byte granularity, so the rest of N's last 4-float group may be zeroed; nothing beyond it
        n4 = (N + 3) // 4 * 4
        assert torch.isnan(D[:, n4:]).all()
        pad = D[:, N:n4]
        assert bool((torch.isnan(pad) | (pad == 0)).all())
    # TF32 is not fp32: the error must also be of TF32 size (guards against a silent fp32/SIMT fallback)
    if K >= 64 and epi == 0:
        assert float(err.max()) > 1e-6


def test_tc_gemm_rejects_misaligned(sb):
    A = torch.randn((128, 250), device="cuda")      # row stride 1000 B: not a multiple of 16
    B = torch.randn((128, 250), device="cuda")
    D = torch.empty((128, 128), device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    st = sb._lib.lib().shems_tc_gemm(p(A), 250, 0, p(B), 250, 0, p(D), 128, 128, 128, 250, 0, None, None, 0, 1, None, None)
    assert st == sb._lib.ERR_INVALID
