import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, wap_b200
eng = wap_b200.WAPEngine(seed=0)
B, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 8, 32
ann = torch.relu(torch.randn(B, h, w, 684, device="cuda"))
for _ in range(3):
    t, l, s = eng.greedy_decode(ann, 40, ignore_eol=True)
torch.cuda.synchronize()
ws = eng._ws[("greedy_p",)]
# barrier words live after hbuf,pre_h,SP,E,ap,ctx,MO
L = h * w
off = sum(((n + 3) // 4) * 4 for n in (2 * B * 256, B * 256, B * 512, 2 * B * L, 2 * B * L, B * 684, B * 128))
prof = ws[off + 8 * 160: off + 8 * 160 + 16].view(torch.int64).cpu().numpy()
names = ["A gru1", "B pre_h", "C cover", "D ctx", "E gru2", "F logit", "G argmax", "setup"]
tot = prof[:7].sum()
print("steps", int(s[0]), "cycles/step", tot / 40, "~us @1.9GHz", tot / 40 / 1900)
for n, c in zip(names, prof): print(f"  {n:10s} {c / 40:10.0f} cyc/step  {c / 40 / 1900:6.2f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng.greedy_decode(ann, 40, ignore_eol=True)
e1.record(); torch.cuda.synchronize(); print("measured us/step incl precompute", e0.elapsed_time(e1) * 1e3 / 200)
