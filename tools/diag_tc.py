import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wap_b200
lib = wap_b200._lib; L = lib.lib()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def rna(t): return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
g = torch.Generator().manual_seed(0)
for (M, N, K, ldd) in [(30, 516, 96, 684), (30, 540, 96, 684), (30, 660, 96, 684), (30, 300, 96, 600), (30, 324, 96, 600), (30, 516, 96, 516), (30, 132, 96, 684), (30, 130, 96, 684)]:
    A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda()
    D0 = torch.randn(M, ldd, generator=g).cuda(); D = D0.clone()
    hi = rna(W); lo = rna(W - hi)
    for x3 in (1, 0):
        D = D0.clone()
        lib.check(L.wap_tc_gemm(st(), p(A), K, p(hi), p(lo) if x3 else None, K, p(D), ldd, M, N, K, 1, None), "tc")
        torch.cuda.synchronize()
        ref = D0.double().clone(); ref[:, :N] += (A.double() @ (W if x3 else hi).double().t())
        err = (D.double() - ref).abs()
        colerr = err.max(0).values
        bad = (colerr > 1e-2).nonzero().flatten().tolist()
        print(f"M{M} N{N} K{K} ldd{ldd} x3={x3}: max err {err.max().item():.3e}  bad cols: {bad[:10]}{'...' if len(bad)>10 else ''} (n={len(bad)})")
