import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, wap_b200
lib = wap_b200._lib; L = lib.lib()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def bench(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
print("tc_gemm (1x1 conv shape) us per launch: M, N, K, ld -> x3 / single / stats")
for (M, N, K, ld) in [(32768, 96, 48, 432), (32768, 96, 240, 432), (32768, 96, 408, 432), (8192, 96, 408, 600), (2048, 96, 660, 684), (2048, 96, 300, 684),
                      (32768, 432, 96, 96), (32768, 120, 96, 96), (8192, 576, 96, 96)]:
    A = torch.randn(M, ld, device="cuda"); W = torch.randn(N, K, device="cuda"); Wl = torch.randn(N, K, device="cuda") * 1e-4
    D = torch.zeros(M, max(N, 96), device="cuda"); stt = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    ldd = D.shape[1]
    t3 = bench(lambda: L.wap_tc_gemm(st(), p(A), ld, p(W), p(Wl), K, p(D), ldd, M, N, K, 0, None))
    t1 = bench(lambda: L.wap_tc_gemm(st(), p(A), ld, p(W), None, K, p(D), ldd, M, N, K, 0, None))
    ts = bench(lambda: L.wap_tc_gemm(st(), p(A), ld, p(W), p(Wl), K, p(D), ldd, M, N, K, 0, p(stt)))
    ta = bench(lambda: L.wap_tc_gemm(st(), p(A), ld, p(W), p(Wl), K, p(D), ldd, M, N, K, 1, None))
    print(f"M{M:6d} N{N:4d} K{K:4d}: x3 {t3:6.1f}  single {t1:6.1f}  x3+stats {ts:6.1f}  x3+accumulate {ta:6.1f}   ({2*M*N*K/t3/1e6:6.1f} TF/s x3)")
print("conv3x3 us: B,H,W")
for (B, H, W) in [(8, 32, 128), (8, 16, 64), (8, 8, 32)]:
    x = torch.randn(B, H, W, 96, device="cuda"); Wt = torch.randn(32, 864, device="cuda"); Wl = Wt * 1e-4
    D = torch.zeros(B * H * W, 24, device="cuda")
    t3 = bench(lambda: L.wap_tc_conv3x3(st(), p(x), 96, B, H, W, 96, p(Wt), p(Wl), 32, 24, p(D), 24, None))
    t1 = bench(lambda: L.wap_tc_conv3x3(st(), p(x), 96, B, H, W, 96, p(Wt), None, 32, 24, p(D), 24, None))
    print(f"B{B} {H}x{W}: x3 {t3:6.1f} single {t1:6.1f}")
