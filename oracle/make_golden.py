"""Generates tests/golden/*.npz from the fp64 oracle (fixed seeds).  The reference ships no golden
vectors and TensorFlow 1.6 cannot run here (parity unpinned, see oracle/wap_oracle.py); these files
pin the ORACLE against accidental change and give the GPU tests a second, frozen comparison point.

    python oracle/make_golden.py            # rewrites tests/golden/

Weights are NOT stored (21 MB): they are regenerated from `lively_params(seed)` /
`init_params(seed)` (numpy RandomState streams are stable across numpy versions)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wap_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cfg1():
    """BASELINE configs[0]: single 64x256 image, 10-token caption, batch 1 (plumbing case)."""
    imgs, caps = O.synth_batch(1, 1, 64, 256, 10)
    x, xm, y, ym = O.prepare_data(imgs, caps)
    P = O.calibrate_moving(O.lively_params(0), x, xm)
    Pt = O.to_torch(P)
    ann, am, _ = O.dense_net(Pt, torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(xm, dtype=torch.float64), False)
    cost = O.get_cost(Pt, ann, torch.as_tensor(y), am, torch.as_tensor(ym, dtype=torch.float64))
    greedy = O.greedy_decode(Pt, ann, 30)
    beam = O.beam_decode(Pt, ann, 10, 30)[0]
    p0, h1, ap1, _ = O.get_word(Pt, torch.tensor([-1]), O.hidden_state_0(Pt, ann), torch.zeros(ann.shape[:3], dtype=torch.float64), ann)
    np.savez_compressed(os.path.join(OUT, "cfg1_forward.npz"), param_seed=0, data_seed=1,
                        annotation=ann.numpy().astype(np.float32), cost=float(cost), greedy=np.array(greedy), beam=np.array(beam),
                        probs0=p0.numpy().astype(np.float32), h1=h1.numpy().astype(np.float32), alpha_past1=ap1.numpy().astype(np.float32))


def small_train():
    """Training-mode forward/backward on a ragged 2-image batch with hashed dropout."""
    imgs, caps = O.synth_batch(7, 2, 32, 48, 3, ragged=True)
    x, xm, y, ym = O.prepare_data(imgs, caps)
    P = O.lively_params(8)
    Pt = O.to_torch(P, torch.float64, requires_grad=True)
    dfn = O.hash_drop_fn(777)
    ann, am, nm = O.dense_net(Pt, torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(xm, dtype=torch.float64), True, dfn)
    cost = O.get_cost(Pt, ann, torch.as_tensor(y), am, torch.as_tensor(ym, dtype=torch.float64), dfn)
    pick = ["conv2d/kernel", "conv2d_1/kernel", "conv2d_98/kernel", "batch_normalization_50/gamma",
            "alpha_past_filter", "embed", "Wo", "ba2h", "V_a_b", "U_rh"]
    grads = torch.autograd.grad(cost, [Pt[n] for n in pick])
    out = {"grad::" + n: g.numpy().astype(np.float32) for n, g in zip(pick, grads)}
    np.savez_compressed(os.path.join(OUT, "small_train.npz"), param_seed=8, data_seed=7, drop_seed=777, cost=float(cost.detach()),
                        annotation=ann.detach().numpy().astype(np.float32),
                        moving_mean_98=nm["batch_normalization_98/moving_mean"].numpy().astype(np.float32),
                        moving_variance_98=nm["batch_normalization_98/moving_variance"].numpy().astype(np.float32), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    cfg1()
    small_train()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
