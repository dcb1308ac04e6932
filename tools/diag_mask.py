import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import O, make_batch, engine_with
B, H, W, cap = 2, 40, 72, 5
x, xm, y, ym = make_batch(52, B, H, W, cap, True)
P = O.lively_params(8)
def run(mode, mask):
    eng = engine_with(P); eng.set_conv_precision(mode); eng.L.wap_debug_set_tc_mask(mask)
    eng.forward_backward(x, y, xm, ym, dropout=True, seed=777)
    g = eng.P.unpack(eng.grads); eng.L.wap_debug_set_tc_mask(0x1FF); return g
ref = run("fp32", 0xFF)
for mask in (0, 1, 2, 4):
    g = run("tf32x3", mask | 0x38)
    errs = {k: np.linalg.norm(g[O.conv_name(k)] - ref[O.conv_name(k)]) / np.linalg.norm(ref[O.conv_name(k)]) for k in (98, 90, 86, 85, 84, 80, 67, 66, 40, 1)}
    print("bwd mask", mask, {k: f"{v:.1e}" for k, v in errs.items()})
