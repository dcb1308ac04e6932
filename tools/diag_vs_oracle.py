import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import O, make_batch, engine_with, t64
B, H, W, cap, ragged, pseed, dseed, drop = 2, 64, 96, 3, True, 9, 60, 500
if len(sys.argv) > 1: B, H, W, cap, ragged = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), bool(int(sys.argv[5]))
mode = sys.argv[6] if len(sys.argv) > 6 else "fp32"
x, xm, y, ym = make_batch(dseed, B, H, W, cap, ragged)
P = O.lively_params(pseed)
eng = engine_with(P); eng.set_conv_precision(mode)
c = float(eng.forward_backward(x, y, xm, ym, dropout=True, seed=drop))
g = eng.P.unpack(eng.grads)
Pt = O.to_torch(P, torch.float64, requires_grad=True)
dfn = O.hash_drop_fn(drop)
ann, am, _ = O.dense_net(Pt, t64(x), t64(xm), True, dfn)
rc = O.get_cost(Pt, ann, torch.as_tensor(y), am, t64(ym), dfn)
names = O.trainable_names(Pt)
gr = torch.autograd.grad(rc, [Pt[n] for n in names])
print("cost", c, float(rc.detach()), "x shape", x.shape)
def key(n):
    m = re.search(r"_(\d+)/", n); return int(m.group(1)) if m else (0 if n.startswith(("conv2d/", "batch_normalization/")) else 1000)
rows = []
for n, r in zip(names, gr):
    r = r.numpy(); e = np.linalg.norm(g[n] - r) / max(np.linalg.norm(r), 1e-30)
    rows.append((key(n), n, e, float(np.abs(r).max())))
rows.sort()
for k, n, e, mx in rows:
    print(f"{n:40s} relL2={e:.2e} max={mx:.2e}")
print("n bad", sum(1 for r in rows if r[2] > 2e-4), "of", len(rows))
