"""Per-tensor gradient deviation of the bf16 mode vs the fp64 oracle (and vs the tf32x3 engine) on a small batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import O, engine_with, make_batch, rel_l2, t64, oracle_grads, engine_grads

B, H, W, cap = 2, 48, 80, 5
P = (O.init_params(0) if os.environ.get('WAP_INIT') else O.lively_params(8)); seed = 777
x, xm, y, ym = make_batch(52, B, H, W, cap, False)
ref_cost, G, names = oracle_grads(P, x, xm, y, ym, seed)
for mode in ("tf32x3", "bf16"):
    eng = engine_with(P, precision=mode, use_graphs=False)
    if len(sys.argv) > 1: eng.L.wap_set_overlap(0)
    cost = float(eng.forward_backward(x, y, xm, ym, dropout=True, seed=seed))
    g = engine_grads(eng)
    print(mode, "cost", cost, ref_cost)
    rows = []
    for n in names:
        rows.append((n, rel_l2(g[n], G[n]), float(np.abs(G[n]).max())))
    def key(n):
        import re
        m = re.search(r"_(\d+)/", n)
        return int(m.group(1)) if m else (0 if n.startswith(("conv2d/", "batch_normalization/")) else 1000)
    for n, e, m in sorted(rows, key=lambda r: (key(r[0]), r[0])):
        if (mode == "bf16" and ("kernel" in n or "gamma" in n) and key(n) % 6 == 0) or (mode != "bf16" and e > 1e-3):
            print(f"  {n:45s} l2 {e:9.3e}  |ref|max {m:9.3e}")
