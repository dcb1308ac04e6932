"""Layer-by-layer deviation of the bf16 encoder forward (block buffers) from the fp64 oracle: a kernel bug shows as a jump."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import O, engine_with, make_batch, rel_l2, t64
B, H, W = 2, 64, 128
x, xm, _, _ = make_batch(5, B, H, W, 3, False)
P = O.lively_params(3)
training = len(sys.argv) > 1
if not training:
    P = O.calibrate_moving(P, x, xm)
taps = {}
O.dense_net(O.to_torch(P), t64(x), t64(xm), training, None, taps=taps)
for mode in ("tf32x3", "bf16"):
    eng = engine_with(P, precision=mode, use_graphs=False)
    ann, _ = eng.dense_net(x, xm, training=training, dropout=False)
    ws = eng._ws[("enc",)]
    L = eng.L
    print(mode, "annotation l2", rel_l2(ann.cpu().numpy(), taps["block2"].numpy()))
    geo = [(B * 16 * 32, 432), (B * 8 * 16, 600), (B * 4 * 8, 688 if mode == "bf16" else 684)]
    for b in range(3):
        off = L.wap_encoder_workspace_offset(B, H, W, int(training), 1, b)
        M, ld = geo[b]
        if b == 2 and mode != "bf16":
            buf = ann.reshape(M, 684).cpu().numpy()
        elif mode == "bf16":
            buf = ws[off: off + (M * ld + 1) // 2].view(torch.bfloat16)[: M * ld].view(M, ld).float().cpu().numpy()
        else:
            buf = ws[off: off + M * ld].view(M, ld).cpu().numpy()
        ref = taps[f"block{b}"].numpy().reshape(M, -1)
        C0 = [48, 216, 300][b]
        row = [f"b{b} in[:{C0}] {rel_l2(buf[:, :C0], ref[:, :C0]):.2e}"]
        for j in range(16):
            c = C0 + 24 * j
            row.append(f"{rel_l2(buf[:, c:c + 24], ref[:, c:c + 24]):.1e}")
        print("  " + " ".join(row))
