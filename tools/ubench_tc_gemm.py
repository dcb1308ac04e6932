"""Latency anatomy of one small tcgen05 GEMM launch (the per-time-step coverage GEMMs of the scans): event-timed back-to-back launches and
the clock64 stamps of CTA (0,0,0).  usage: python tools/ubench_tc_gemm.py [anatomy|kblocks|graph]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wap_b200
from wap_b200.engine import _p, _stream
eng = wap_b200.WAPEngine(seed=0); L = eng.L
CLK = 1.965e3   # cycles per us at the boost clock


def run(M, N, K, tr, splitk, x3=True, reps=200):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda")
    Np = (N + 3) // 4 * 4
    rna = lambda t: ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    hi = torch.zeros(max(N, 128), K, device="cuda"); lo = torch.zeros_like(hi)
    hi[:N] = rna(W.t().contiguous()); lo[:N] = rna(W.t().contiguous() - hi[:N])
    D = torch.zeros((N, M) if tr else (M, Np), device="cuda")
    ldd = M if tr else Np
    trace = torch.zeros(1100, dtype=torch.int64, device="cuda")
    def go(): wap_b200._lib.check(L.wap_tc_gemm_ex(_stream(), _p(A), K, _p(hi), _p(lo) if x3 else None, K, _p(D), ldd, M, N, K, tr, splitk, None), "gemm")
    for _ in range(5): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    L.wap_debug_tc_trace(_p(trace)); go(); torch.cuda.synchronize(); L.wap_debug_tc_trace(None)
    t = trace.cpu().numpy(); nkb = (K + 31) // 32 if splitk <= 1 else -(-((K + 31) // 32) // min(splitk, (K + 31) // 32))
    t0 = t[1024]
    st = [(t[1024 + i] - t0) / CLK for i in range(4)]
    chunks = [round((t[1028 + i] - t0) / CLK, 2) for i in range(4) if t[1028 + i] > 0]
    kb = [[(t[k * 4 + j] - t0) / CLK for j in range(4)] for k in range(min(nkb, 24 if os.environ.get('ALLKB') else 3))]
    D.zero_(); go(); torch.cuda.synchronize()
    ref = (A.double() @ W.double())
    got = (D.t() if tr else D[:, :N]).double()
    err = float((got - ref).norm() / ref.norm())
    print(f"M={M} N={N} K={K} tr={tr} splitk={splitk} x3={x3}: {us:6.2f} us/launch | setup {st[1]:.2f} acc-ready {st[2]:.2f} chunk starts {chunks} all warps done {st[3]:.2f} us | "
          f"k-blocks [tma, landed, split, mma] {[[round(v, 2) for v in r] for r in kb]} | rel err {err:.1e}")


def graph_latency(M, N, K, tr, splitk, interleave, x3=True, reps=50):
    """per-launch time inside a replayed CUDA graph (no host cost): reps x [GEMM (+ a small elementwise kernel when interleave)]"""
    A = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda")
    Np = (N + 3) // 4 * 4
    rna = lambda t: ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    hi = torch.zeros(max(N, 128), K, device="cuda"); lo = torch.zeros_like(hi)
    hi[:N] = rna(W.t().contiguous()); lo[:N] = rna(W.t().contiguous() - hi[:N])
    D = torch.zeros((N, M) if tr else (M, Np), device="cuda"); ldd = M if tr else Np
    small = torch.zeros(8 * 256, device="cuda")
    def go():
        wap_b200._lib.check(L.wap_tc_gemm_ex(_stream(), _p(A), K, _p(hi), _p(lo) if x3 else None, K, _p(D), ldd, M, N, K, tr, splitk, None), "gemm")
        if interleave == 1: small.add_(1.0)
        if interleave == 2: wap_b200._lib.check(L.wap_tc_gemm_ex(_stream(), _p(A), K, _p(hi), _p(lo) if x3 else None, K, _p(D), ldd, 128, N, 32, tr, 1, None), "gemm")
    s = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        go(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            for _ in range(reps): go()
    torch.cuda.synchronize()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    import json, tempfile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay(); torch.cuda.synchronize()
    f = tempfile.mktemp(suffix=".json"); prof.export_chrome_trace(f)
    ev = sorted([e for e in json.load(open(f))["traceEvents"] if e.get("cat") == "kernel" and "tc_gemm" in e["name"]], key=lambda e: e["ts"])
    os.remove(f)
    durs = sorted(e["dur"] for e in ev); gaps = sorted(b["ts"] - a["ts"] for a, b in zip(ev, ev[1:]))
    print(f"graph: M={M} N={N} K={K} tr={tr} splitk={splitk} x3={x3} interleave={interleave}: {e0.elapsed_time(e1) * 1e3 / reps:6.2f} us per iteration; "
          f"CUPTI kernel duration median {durs[len(durs) // 2]:.2f} us, start-to-start median {gaps[len(gaps) // 2]:.2f} us ({len(ev)} kernels)")


mode = sys.argv[1] if len(sys.argv) > 1 else "anatomy"
if mode == "kblocks":        # per-k-block stamps of a block-2-sized 1x1 conv (3xTF32 and single-pass TF32)
    os.environ["ALLKB"] = "1"
    run(8192, 96, 608, 0, 1); run(8192, 96, 608, 0, 1, x3=False)
elif mode == "graph":        # dependent-launch latency inside a replayed graph + CUPTI durations
    graph_latency(128, 121, 32, 1, 1, 0); graph_latency(128, 121, 32, 1, 1, 0, x3=False); graph_latency(128, 121, 32, 0, 1, 0); graph_latency(128, 32, 32, 0, 1, 0)
    L.wap_set_pdl(0)
    graph_latency(128, 121, 32, 1, 1, 0); graph_latency(2048, 121, 512, 1, 8, 0)
    L.wap_set_pdl(1)
    graph_latency(2048, 121, 512, 1, 8, 0); graph_latency(2048, 512, 128, 0, 1, 0)
else:                        # anatomy of one tile: set-up, accumulator ready, each epilogue chunk
    run(128, 121, 32, 1, 1); run(128, 121, 32, 0, 1); run(128, 32, 32, 0, 1); run(128, 121, 32, 0, 1, x3=False)
