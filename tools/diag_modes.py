"""Diagnostic: per-tensor gradient deviation between conv precision modes on identical inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import O, make_batch, engine_with
B, H, W, cap = 2, 40, 72, 5
x, xm, y, ym = make_batch(52, B, H, W, cap, True)
P = O.lively_params(8)
res = {}
for mode in sys.argv[1:] or ["fp32", "tf32x3"]:
    eng = engine_with(P); eng.set_conv_precision(mode)
    c = float(eng.forward_backward(x, y, xm, ym, dropout=True, seed=777))
    res[mode] = (c, eng.P.unpack(eng.grads))
    print(mode, "cost", c)
a, b = list(res)[:2]
ga, gb = res[a][1], res[b][1]
rows = []
for n in ga:
    e = np.linalg.norm(ga[n] - gb[n]) / max(np.linalg.norm(ga[n]), 1e-30)
    rows.append((n, e, float(np.abs(ga[n]).max())))
def key(n):
    import re
    m = re.search(r"_(\d+)/", n); return int(m.group(1)) if m else (0 if n.startswith(("conv2d/", "batch_normalization/")) else 1000)
for n, e, mx in sorted(rows, key=lambda r: (key(r[0]), r[0])):
    if n.endswith("kernel") or not n.startswith(("conv", "batch")):
        print(f"{n:40s} relL2={e:.2e} max={mx:.2e}")
