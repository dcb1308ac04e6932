"""Per-segment device time of one training step (stream launches + PDL): encoder fwd / decoder fwd / decoder bwd / encoder bwd / update."""
import sys, ctypes as C
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wap_b200
from wap_b200 import synth
from wap_b200.engine import _p, _stream, ANNOT_C
eng = wap_b200.WAPEngine(seed=0); eng.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
L = eng.L
B, H, W, T = 8, 128, 512, 40
imgs, caps = synth.synth_batch(0, B, H, W, T)
x, xm, y, ym = wap_b200.prepare_data(imgs, caps)
x, xm = torch.as_tensor(x).cuda(), torch.as_tensor(xm).cuda(); y, ym = torch.as_tensor(y).cuda().int(), torch.as_tensor(ym).cuda()
T = y.shape[0]
h, w = eng.annotation_hw(H, W)
a_mask = eng.mask_subsample(xm)
ews = eng._buf(("enc",), L.wap_encoder_workspace_floats(B, H, W, 1)); dws = eng._buf(("dec",), L.wap_decoder_train_workspace_floats(B, h, w, T))
ann = eng._buf(("ann",), B * h * w * ANNOT_C); d_ann = eng._buf(("dann",), B * h * w * ANNOT_C)
sd = C.c_uint32(5)
segs = [("enc_fwd", lambda st: L.wap_encoder_forward(st, _p(eng.P.flat), _p(eng.P.moving), _p(x), B, H, W, 1, 1, sd, _p(ann), _p(ews))),
        ("dec_fwd", lambda st: L.wap_decoder_train_forward(st, _p(eng.P.flat), _p(ann), _p(a_mask), _p(y), _p(ym), B, h, w, T, 1, sd, _p(eng._cost), _p(dws))),
        ("dec_bwd", lambda st: L.wap_decoder_train_backward(st, _p(eng.P.flat), _p(eng.grads), _p(ann), _p(a_mask), _p(y), _p(ym), B, h, w, T, 1, sd, _p(d_ann), _p(dws))),
        ("enc_bwd", lambda st: L.wap_encoder_backward(st, _p(eng.P.flat), _p(eng.grads), _p(x), B, H, W, 1, sd, _p(ann), _p(d_ann), _p(ews))),
        ("update", lambda st: eng.apply_update(1.0))]
acc = {n: 0.0 for n, _ in segs}; cnt = {n: 0 for n, _ in segs}
N = 12
for it in range(N + 3):
    eng.grads.zero_()
    evs = [torch.cuda.Event(True) for _ in range(len(segs) + 1)]
    evs[0].record()
    for i, (n, f) in enumerate(segs):
        n0 = L.wap_launch_count(); r = f(_stream()); cnt[n] = L.wap_launch_count() - n0
        evs[i + 1].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i, (n, _) in enumerate(segs): acc[n] += evs[i].elapsed_time(evs[i + 1])
print({n: (round(v / N, 3), cnt[n]) for n, v in acc.items()}, "total", round(sum(acc.values()) / N, 3))
