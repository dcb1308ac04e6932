"""Per-k-block clock64 trace of the tcgen05 GEMM pipeline (CTA 0): where does a k-block's time go?"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, wap_b200
L = wap_b200._lib.lib()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(M, N, K, ld, x3, conv=None):
    tr = torch.zeros(4 * 64, dtype=torch.int64, device="cuda")
    if conv is None:
        A = torch.randn(M, ld, device="cuda"); W = torch.randn(N, K, device="cuda"); Wl = W * 1e-4; D = torch.zeros(M, max(N, 96), device="cuda")
        f = lambda: L.wap_tc_gemm(st(), p(A), ld, p(W), p(Wl) if x3 else None, K, p(D), D.shape[1], M, N, K, 0, None)
        nkb = (K + 31) // 32
    else:
        B, H, W_ = conv
        x = torch.randn(B, H, W_, 96, device="cuda"); Wt = torch.randn(32, 864, device="cuda"); Wl = Wt * 1e-4; D = torch.zeros(B * H * W_, 24, device="cuda")
        f = lambda: L.wap_tc_conv3x3(st(), p(x), 96, B, H, W_, 96, p(Wt), p(Wl) if x3 else None, 32, 24, p(D), 24, None)
        nkb = 27
    for _ in range(3): f()
    L.wap_debug_tc_trace(p(tr)); f(); torch.cuda.synchronize(); L.wap_debug_tc_trace(None)
    t = tr.cpu().numpy().reshape(-1, 4)[:nkb]
    t0 = t[0, 0]
    print(f"--- M{M} N{N} K{K} x3={x3} conv={conv}: total {t[-1,3]-t0} cycles for {nkb} k-blocks; per k-block {(t[-1,3]-t[3,3])/max(nkb-4,1):.0f}")
    print(" kb  tma_issue  landed(+lat)  split_done(+dur)  mma_issue(+wait)")
    for kb in range(min(nkb, 14)):
        a, b, c, d = t[kb] - t0
        print(f" {kb:2d}  {a:8d}  {b:8d} (+{b-a:5d})  {c:8d} (+{c-b:5d})  {d:8d} (+{d-(c if x3 else b):5d})")
run(2048, 96, 660, 684, True)
run(2048, 96, 660, 684, False)
run(32768, 96, 408, 432, True)
run(32768, 96, 408, 432, False)
run(0, 0, 0, 0, True, conv=(8, 32, 128))
