"""Where a cfg5 image goes: device time of encoder / precompute+decode (CUDA events) and host wall-clock per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wap_b200
from wap_b200 import synth
eng = wap_b200.WAPEngine(seed=0, use_graphs=False)
images, caps = synth.synth_crohme(5, 60)
xs = [torch.as_tensor(np.moveaxis(im, 0, -1).astype(np.float32) / 255.).unsqueeze(0).cuda() for im in images]
for x in xs[:10]:
    a, _ = eng.dense_net(x, None, training=False); eng.greedy_decode(a, 17)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
te = td = 0.0; he = hd = 0.0
for x in xs:
    e0, e1, e2 = ev(), ev(), ev()
    t0 = time.perf_counter(); e0.record(); a, _ = eng.dense_net(x, None, training=False); e1.record(); t1 = time.perf_counter()
    eng.greedy_decode(a, 17); e2.record(); t2 = time.perf_counter()
    torch.cuda.synchronize()
    te += e0.elapsed_time(e1); td += e1.elapsed_time(e2); he += t1 - t0; hd += t2 - t1
n = len(xs)
print(f"per image: encoder device {te / n * 1e3:.0f} us (host call {he / n * 1e6:.0f} us) | precompute+decode(17) device {td / n * 1e3:.0f} us (host call {hd / n * 1e6:.0f} us)")
t0 = time.perf_counter()
for x in xs:
    a, _ = eng.dense_net(x, None, training=False); eng.greedy_decode(a, 17)
torch.cuda.synchronize()
print(f"back to back: {(time.perf_counter() - t0) / n * 1e6:.0f} us per image wall")
