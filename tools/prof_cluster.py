"""Per-stage SM cycles of the cluster-per-image greedy kernel (cluster 0, rank 0): setup, stages 1..8, cluster barriers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wap_b200
B, h, w, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 8, 32, 40
if len(sys.argv) > 3: h, w = int(sys.argv[2]), int(sys.argv[3])
eng = wap_b200.WAPEngine(seed=0)
L = eng.L
ann = torch.randn(B, h, w, 684, device="cuda") * 0.5
L.wap_debug_set_greedy_impl(2)       # cluster kernel + profiling
for _ in range(2):
    eng.greedy_decode(ann, steps, ignore_eol=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.greedy_decode(ann, steps, ignore_eol=True); e1.record(); torch.cuda.synchronize()
ws = eng._ws[("greedy_p",)]
nfl = int(L.wap_greedy_persistent_workspace_floats(B, h, w))
# prof words: after watch | qf | h0 | mean in the cluster carve-up
o = 0
def take(n):
    global o; p = o; o += (n + 3) & ~3; return p
take(B * h * w * 512); take(121 * 512); take(B * 256); take(B * 684); p = take(32)
prof = ws[p: p + 24].view(torch.int64).cpu().numpy()
names = ["setup", "1 gru1", "2 preh+speller", "3 attention", "4 softmax-local", "5 alpha+ctx", "6 gru2", "7 logits", "8 argmax", "barriers"]
clk = 1.965e3   # cycles per us at 1965 MHz
print(f"B={B} {h}x{w}: {e0.elapsed_time(e1) * 1e3 / steps:.1f} us/step incl. precompute")
for n, v in zip(names, prof[:10]):
    print(f"  {n:18s} {v / clk / (1 if n == 'setup' else steps):8.2f} us" + ("" if n == "setup" else " per step"))
L.wap_debug_set_greedy_impl(0)
