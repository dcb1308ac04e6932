import sys; import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import *

P = O.lively_params(3)
x, xm, y, ym = make_batch(11, 3, 32, 64, 4, True)
eng = engine_with(P)
eng.forward_backward(x, y, xm, ym, True, 1)
names = {n: (o, int(np.prod(sh))) for n, o, sh, f in eng.P.table if not f & 2}
for seed in (5, 6, 7, 5):
    eng.use_graphs = True
    cg = float(eng.forward_backward(x, y, xm, ym, True, seed)); gg = eng.grads.clone()
    eng.use_graphs = False
    ce = float(eng.forward_backward(x, y, xm, ym, True, seed)); ge = eng.grads.clone()
    ce2 = float(eng.forward_backward(x, y, xm, ym, True, seed)); ge2 = eng.grads.clone()
    worst = sorted(((rel_l2(gg[o:o + k].cpu().numpy(), ge[o:o + k].cpu().numpy()), n) for n, (o, k) in names.items() if float(ge[o:o+k].abs().max()) > 0), reverse=True)[:4]
    print(seed, cg, ce, "graph-vs-eager", rel_l2(gg.cpu().numpy(), ge.cpu().numpy()), "eager-vs-eager", rel_l2(ge2.cpu().numpy(), ge.cpu().numpy()), worst)
