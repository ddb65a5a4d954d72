import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, shems_b200 as sb
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
def run(M,N,K,a_mn,b_mn):
    g=torch.Generator(device="cuda").manual_seed(1)
    lda = ((M if a_mn else K)+3)//4*4; ldb=((N if b_mn else K)+3)//4*4
    A=torch.randn((K,lda) if a_mn else (M,lda),device="cuda",generator=g); B=torch.randn((K,ldb) if b_mn else (N,ldb),device="cuda",generator=g)
    D=torch.full((M,N),float('nan'),device="cuda")
    st=sb._lib.lib().shems_tc_gemm(p(A),lda,a_mn,p(B),ldb,b_mn,p(D),N,M,N,K,0,None,None,0,1,None,None); sb._lib.check(st); torch.cuda.synchronize()
    Al=(A[:K,:M].T if a_mn else A[:M,:K]).double(); Bl=(B[:K,:N].T if b_mn else B[:N,:K]).double()
    ref=Al@Bl.T; err=(D.double()-ref).abs(); bound=3e-3*(Al.abs()@Bl.abs().T)+1e-6
    bad=(err>bound)
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: max ratio {float((err/bound).max()):.3g} bad frac {float(bad.float().mean()):.3f}", end='')
    if bad.any():
        rows=bad.any(1).nonzero().flatten(); cols=bad.any(0).nonzero().flatten()
        print(f" bad rows {rows[:6].tolist()}..{rows[-3:].tolist()} n={len(rows)} bad cols {cols[:6].tolist()}..{cols[-3:].tolist()} n={len(cols)}")
    else: print()
for a_mn,b_mn in ((0,0),(0,1),(1,0),(1,1)):
    for (M,N,K) in ((128,128,32),(128,128,8),(128,128,64),(128,128,256),(256,256,250),(100,60,40)):
        run(M,N,K,a_mn,b_mn)
