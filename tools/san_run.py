import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import O, make_batch, engine_with
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B, H, W, cap = 2, 64, 96, 3
x, xm, y, ym = make_batch(60, B, H, W, cap, True)
eng = engine_with(O.lively_params(9)); eng.set_conv_precision(mode)
c = eng.train_step(x, y, xm, ym, dropout=True, seed=500)
torch.cuda.synchronize(); print("cost", float(c))
ann, _ = eng.dense_net(x[:1], None, training=False)
t, l, _ = eng.greedy_decode(ann, 6, check_every=0); torch.cuda.synchronize(); print(t[0].tolist())
