import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, wap_b200
from wap_b200 import synth
eng = wap_b200.WAPEngine(seed=0); eng.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
imgs, caps = synth.synth_batch(4, 8, 128, 512, 40)
x, xm, y, ym = wap_b200.prepare_data(imgs, caps)
xd = torch.as_tensor(x).cuda(); xmd = torch.as_tensor(xm).cuda(); yd = torch.as_tensor(y).cuda().int(); ymd = torch.as_tensor(ym).cuda()
for _ in range(3): eng.train_step(xd, yd, xmd, ymd)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    eng.train_step(xd, yd, xmd, ymd)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"issue {1e3 * (t1 - t0):.2f} ms   issue+drain {1e3 * (t2 - t0):.2f} ms")
