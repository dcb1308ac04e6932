import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import O, engine_with, make_batch, rel_l2
P = O.lively_params(3)
a, a2, b = engine_with(P), engine_with(P), engine_with(P)
b.use_graphs = True
x, xm, y, ym = make_batch(11, 3, 32, 64, 4, True)
for step in range(5):
    ca = float(a.train_step(x, y, xm, ym)); ga = a.grads.clone()
    ca2 = float(a2.train_step(x, y, xm, ym)); ga2 = a2.grads.clone()
    cb = float(b.train_step(x, y, xm, ym)); gb = b.grads.clone()
    print(step, ca, ca2, cb, "eager-eager", rel_l2(ga2.cpu().numpy(), ga.cpu().numpy()), "graph-eager", rel_l2(gb.cpu().numpy(), ga.cpu().numpy()),
          "params", rel_l2(a2.P.flat.cpu().numpy(), a.P.flat.cpu().numpy()), rel_l2(b.P.flat.cpu().numpy(), a.P.flat.cpu().numpy()))
