import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import O, make_batch, engine_with
B, H, W, cap = 2, 40, 72, 5
x, xm, y, ym = make_batch(52, B, H, W, cap, True)
P = O.lively_params(8)
eng = engine_with(P); eng.set_conv_precision("tf32x3")
L = eng.L
def rna(t): return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def check(tag):
    ws = eng._ws[("enc", B, H, W, True)]
    bad = []
    for k in range(1, 99):
        w = torch.as_tensor(P[O.conv_name(k)]).cuda()
        kh, kw, ci, co = w.shape
        offb = L.wap_encoder_workspace_offset(B, H, W, 1, 7, k); offbl = L.wap_encoder_workspace_offset(B, H, W, 1, 9, k)
        if kh == 1:
            exp = w.reshape(ci, co)
        else:
            exp = torch.zeros(ci, 9, 32, device="cuda")
            wf = w.reshape(9, ci, co)
            for tap in range(9):
                exp[:, tap, :co] = wf[8 - tap]
            exp = exp.reshape(ci, 288)
        hi = rna(exp); lo = rna(exp - hi)
        ld = 192 if (kh == 1 and co == 96) else exp.shape[1]      # dense-level 1x1 layers live in a [Cin][192] pair buffer (odd level | even level)
        got = torch.as_strided(ws, exp.shape, (ld, 1), offb); gotl = torch.as_strided(ws, exp.shape, (ld, 1), offbl)
        e = (got - hi).abs().max().item(); el = (gotl - lo).abs().max().item()
        if e > 0 or el > 0: bad.append((k, e, el))
    print(tag, "bad bwd weight copies:", bad[:10], len(bad))
c = eng.forward_backward(x, y, xm, ym, dropout=True, seed=777)
torch.cuda.synchronize()
check("after fwd+bwd")
print("total ws floats", L.wap_encoder_workspace_offset(B, H, W, 1, 10, 0), "allocated", eng._ws[("enc", B, H, W, True)].numel())
