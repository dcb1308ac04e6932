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
for k1 in (int(a) for a in sys.argv[1:]):
    ref_dy1, ref_dy2 = torch.autograd.grad(rc, [taps[f"conv{k1}"], taps[f"conv{k1 + 1}"]], retain_graph=True)
    M = ref_dy1.shape[0] * ref_dy1.shape[1] * ref_dy1.shape[2]
    eng = engine_with(P)
    buf = torch.zeros(M * (96 + 32 + 96), device="cuda")
    eng.L.wap_debug_capture(k1, C.c_void_p(buf.data_ptr()))
    eng.forward_backward(x, y, xm, ym, dropout=True, seed=drop)
    torch.cuda.synchronize(); eng.L.wap_debug_capture(-1, None)
    da1 = buf[: M * 96].view(M, 96).cpu().double(); dy2 = buf[M * 96: M * 128].view(M, 32)[:, :24].cpu().double(); dy1 = buf[M * 128:].view(M, 96).cpu().double()
    r1 = ref_dy1.reshape(M, 96); r2 = ref_dy2.reshape(M, 24)
    e1 = (dy1 - r1).abs(); e2 = (dy2 - r2).abs()
    print(f"layer {k1}: dy2 relmax {e2.max() / r2.abs().max():.2e}  dy1 relmax {e1.max() / r1.abs().max():.2e}  worst dy1 channel {int(e1.max(0).values.argmax())} err-per-channel top {sorted((e1.max(0).values / r1.abs().max()).tolist(), reverse=True)[:4]}")
    # wgrad check for the 1x1 layer: dW = block_buffer[:, :C]^T @ dy1
    g = eng.P.unpack(eng.grads)[O.conv_name(k1)].reshape(-1, 96)
    Cin = g.shape[0]
    hh, ww = ref_dy1.shape[1], ref_dy1.shape[2]
    ann_e = eng._ws[("ann", B, hh, ww)][: M * 684].view(M, 684).cpu().double()
    blk = taps["block2"].reshape(M, 684)
    print("   block-buffer fwd err", float((ann_e - blk).abs().max() / blk.abs().max()))
    dW_t = ann_e[:, :Cin].t() @ dy1
    ref = torch.autograd.grad(rc, Pt[O.conv_name(k1)], retain_graph=True)[0].reshape(-1, 96)
    print(f"   dW1: engine vs oracle {np.linalg.norm(g - ref.numpy()) / np.linalg.norm(ref.numpy()):.2e}; torch(engine buf, engine dy1) vs oracle {float((dW_t - ref).norm() / ref.norm()):.2e}; rows with worst err: {np.argsort(-np.abs(g - ref.numpy()).max(1))[:6].tolist()} of {Cin}")
