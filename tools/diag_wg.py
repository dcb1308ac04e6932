import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, wap_b200
lib = wap_b200._lib; L = lib.lib()
p = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
Mpix, Cc, N = 32, 128, 96
X = torch.zeros(Mpix, Cc, device="cuda"); 
for q in range(32): X[q, q * 4 % 128 + q // 32] = 1.0          # pixel q -> channel 4q (spread over blocks)
dY = (torch.arange(Mpix, device="cuda").float()[:, None] * 100 + torch.arange(N, device="cuda").float()[None, :])   # dY[p,n] = 100p+n
dW = torch.zeros(Cc, N, device="cuda")
lib.check(L.wap_tc_wgrad_1x1(st(), p(X), Cc, p(dY), N, Mpix, Cc, N, p(dW), N, 0), "wg")
torch.cuda.synchronize()
ref = X.t() @ dY
print("max err", (dW - ref).abs().max().item(), "nonzero rows got", (dW.abs().sum(1) > 0).nonzero().flatten().tolist()[:40])
print("nonzero rows ref", (ref.abs().sum(1) > 0).nonzero().flatten().tolist()[:40])
r = (dW.abs().sum(1) > 0).nonzero().flatten().tolist()
for m in r[:6]: print("row", m, dW[m, :12].tolist(), "ref", ref[m, :6].tolist())
# ---- dump stage-0 smem and try K-major flags
dbg = torch.zeros(16384 // 4 + 12288 // 4, device="cuda")
for xor in (0, (1 << 15) | (1 << 16), (1 << 15), (1 << 16)):
    L.wap_debug_wgrad(p(dbg), xor)
    dW.zero_()
    lib.check(L.wap_tc_wgrad_1x1(st(), p(X), Cc, p(dY), N, Mpix, Cc, N, p(dW), N, 0), "wg")
    torch.cuda.synchronize()
    print("idesc xor", hex(xor), "out abs sum", dW.abs().sum().item(), "max", dW.abs().max().item())
L.wap_debug_wgrad(None, 0)
A = dbg[:4096].view(4, 32, 32).cpu()    # [block][row k][32 floats] (swizzled)
Bm = dbg[4096:].view(3, 32, 32).cpu()
print("A tile nonzeros per block", [(A[b] != 0).sum().item() for b in range(4)], "B tile abs-sum per block", [Bm[b].abs().sum().item() for b in range(3)])
print("B block0 row0", Bm[0, 0, :8].tolist(), "row1", Bm[0, 1, :8].tolist())
