"""A/B of a tensor-core debug mask on the cfg2 training step (graph replay): python tools/ab_mask.py 0x3FF 0x7FF"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, wap_b200
from wap_b200 import synth
imgs, caps = synth.synth_batch(4, 8, 128, 512, 40)
x, xm, y, ym = wap_b200.prepare_data(imgs, caps)
for m in sys.argv[1:]:
    eng = wap_b200.WAPEngine(seed=0)
    eng.L.wap_debug_set_tc_mask(int(m, 0))
    for _ in range(5): eng.train_step(x, y, xm, ym, lr=1.0, dropout=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): eng.train_step(x, y, xm, ym, lr=1.0, dropout=True)
    e1.record(); torch.cuda.synchronize()
    print(m, f"{e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)
    eng.L.wap_debug_set_tc_mask(0x3FF)
