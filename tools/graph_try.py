"""Experiment: whole training step captured in a CUDA graph vs stream launches (fixed dropout seed)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import wap_b200
from wap_b200 import synth
eng = wap_b200.WAPEngine(seed=0)
eng.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
eng.L.wap_set_pdl(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
eng.L.wap_set_overlap(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
print("pdl", eng.L.wap_get_pdl(), "overlap", eng.L.wap_get_overlap())
B, H, W, T = 8, 128, 512, 40
imgs, caps = synth.synth_batch(0, B, H, W, T)
x, xm, y, ym = wap_b200.prepare_data(imgs, caps)
xd, xmd = torch.as_tensor(x).cuda(), torch.as_tensor(xm).cuda()
yd, ymd = torch.as_tensor(y).cuda().int(), torch.as_tensor(ym).cuda()
import ctypes as C
trap = (C.c_int * 256)(); eng.L.wap_debug_trap_info(trap)
def step():
    try:
        eng.forward_backward(xd, yd, xmd, ymd, True, 5); eng.apply_update(1.0)
    except Exception as e:
        eng.L.wap_debug_trap_info(trap); print("TRAP NOTES", trap[1], [list(trap[8 + 8 * i: 16 + 8 * i]) for i in range(min(trap[1], 30))]); raise
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); a = torch.cuda.Event(True); b = torch.cuda.Event(True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print("stream ms/step", timeit(step))
c0 = float(eng._cost[0])
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
print("graph  ms/step", timeit(g.replay))
print("cost stream", c0, "graph", float(eng._cost[0]))
