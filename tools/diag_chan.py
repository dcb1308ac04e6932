import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import O, make_batch, engine_with, t64
B, H, W, cap, ragged, pseed, dseed, drop = 2, 64, 96, 3, True, 9, 60, 500
x, xm, y, ym = make_batch(dseed, B, H, W, cap, ragged)
P = O.lively_params(pseed)
Pt = O.to_torch(P, torch.float64, requires_grad=True)
dfn = O.hash_drop_fn(drop)
taps = {}
ann, am, _ = O.dense_net(Pt, t64(x), t64(xm), True, dfn, taps=taps)
rc = O.get_cost(Pt, ann, torch.as_tensor(y), am, t64(ym), dfn)
k1 = 93
ref_dy1 = torch.autograd.grad(rc, taps[f"conv{k1}"], retain_graph=True)[0]
M = ref_dy1.shape[0] * ref_dy1.shape[1] * ref_dy1.shape[2]
r1 = ref_dy1.reshape(M, 96)
for it in range(8):
    eng = engine_with(P)
    buf = torch.zeros(M * (96 + 32 + 96), device="cuda")
    eng.L.wap_debug_capture(k1, C.c_void_p(buf.data_ptr()))
    eng.forward_backward(x, y, xm, ym, dropout=True, seed=drop)
    torch.cuda.synchronize(); eng.L.wap_debug_capture(-1, None)
    da1 = buf[: M * 96].view(M, 96).cpu().double(); dy1 = buf[M * 128:].view(M, 96).cpu().double()
    e1 = (dy1 - r1).abs().max(0).values / r1.abs().max()
    ch = int(e1.argmax())
    print(f"run {it}: worst channel {ch} err {float(e1[ch]):.2e}")
    if e1[ch] > 1e-4:
        ws = eng._ws[("enc", B, H, W, True)]
        off = eng.L.wap_encoder_workspace_offset(B, H, W, 1, 2, k1)
        y1 = ws[off: off + M * 96].view(M, 96).cpu().double()
        ry1 = taps[f"conv{k1}"].detach().reshape(M, 96)
        g = torch.as_tensor(P[O.bn_name(k1, "gamma")]).double(); b = torch.as_tensor(P[O.bn_name(k1, "beta")]).double()
        col = y1[:, ch]; mean = col.mean(); var = col.var(unbiased=False); inv = 1 / torch.sqrt(var + 1e-4)
        pre = (col - mean) * inv * g[ch] + b[ch]
        print("  y1 col err vs oracle", float((col - ry1[:, ch]).abs().max()), "mean", float(mean), "std", float(var.sqrt()), "gamma", float(g[ch]), "beta", float(b[ch]))
        print("  pre-activation sorted |pre| smallest:", sorted(pre.abs().tolist())[:6])
        keep = O.drop_hash_keep(drop, k1, M * 96).reshape(M, 96)[:, ch]
        gg = da1[:, ch] * (pre > 0).double() * torch.as_tensor(keep).double() / 0.8
        xh = (col - mean) * inv
        exp = g[ch] * inv * (gg - gg.mean() - xh * (gg * xh).mean())
        print("  dy1 col: engine vs recomputed-from-engine-inputs", float((dy1[:, ch] - exp).abs().max()), " recomputed vs oracle", float((exp - r1[:, ch]).abs().max()), " scale", float(r1.abs().max()))
        print("  da1 col", da1[:, ch].tolist()[:8])
        break
