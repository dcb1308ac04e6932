"""Shared helpers for the parity tests (engine vs CPU oracle)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import wap_oracle as O  # noqa: E402

REL_TOL = 1e-3      # north_star: "within 1e-3 relative fp32 tolerance"


def rel_err(a, b) -> float:
    """max |a-b| / max |b|  (b = oracle)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    return d / max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-30)


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_batch(seed, B, H, W, cap_len, ragged=False):
    imgs, caps = O.synth_batch(seed, B, H, W, cap_len, ragged=ragged)
    return O.prepare_data(imgs, caps)


def engine_with(P_np, precision="tf32x3", use_graphs=True):
    import wap_b200
    eng = wap_b200.WAPEngine(init=False, precision=precision, use_graphs=use_graphs)
    eng.P.load(P_np)
    return eng


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def engine_gates(eng, B, H, W, params=None):
    """{conv index k: uint8 [pixels, Cout_k]} ReLU gates of the training-mode forward the engine has just run (forward_backward /
    train_step on a (B,H,W) batch), exported by wap_debug_bn_gates from the raw conv outputs + statistics in its workspace.
    params: flat parameter tensor that forward used (default: the engine's current one; pass a pre-step clone after train_step,
    whose optimiser update has already changed gamma / beta)."""
    params = eng.P.flat if params is None else params
    import ctypes as C
    from wap_b200 import _lib
    ws = eng._ws[("enc",)]
    out = {}
    geom = []
    hs, ws_ = -(-H // 2), -(-W // 2)
    geom.append(B * hs * ws_)
    h, w = -(-hs // 2), -(-ws_ // 2)
    for b in range(3):
        geom.append(B * h * w); h, w = -(-h // 2), -(-w // 2)
    plan = O.encoder_plan()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for k in range(99):
        cout = int(eng.P.views[O.conv_name(k)].shape[-1])
        block = 0 if k == 0 else (0 if k <= 33 else 1 if k <= 66 else 2)
        npix = geom[0] if k == 0 else geom[1 + block]
        g = torch.empty(npix * cout, dtype=torch.uint8, device=eng.device)
        _lib.check(eng.L.wap_debug_bn_gates(st, C.c_void_p(params.data_ptr()), C.c_void_p(ws.data_ptr()), B, H, W, k, C.c_void_p(g.data_ptr())),
                   "wap_debug_bn_gates")
        out[k] = g.cpu()
    del plan
    return out


def grad_errors(names, got_by_name, ref_by_name, excluded=None, tol=2e-3):
    """Per-tensor max-norm error (tensors whose gradient is tiny relative to the largest one are compared on the global
    scale: they are dominated by fp32 rounding of large cancelling sums).  Nothing is masked (excluded stays empty: the oracle
    is evaluated with the engine's ReLU gates, helpers.engine_gates).
    Returns (failures, worst_err, worst_name, median_rel_l2, n_masked_entries)."""
    excluded = excluded or {}
    gmax = max(float(np.max(np.abs(ref_by_name[n]))) for n in names)
    fails, worst, worst_name, l2s, masked = [], 0.0, None, [], 0
    for n in names:
        got = np.asarray(got_by_name[n], dtype=np.float64); ref = np.asarray(ref_by_name[n], dtype=np.float64)
        d = np.abs(got - ref)
        if n in excluded:
            d[..., excluded[n]] = 0.0
            masked += int(np.prod(d.shape[:-1])) * len(excluded[n])
        scale = max(float(np.max(np.abs(ref))), 1e-4 * gmax)
        err = float(np.max(d)) / scale
        l2s.append(float(np.linalg.norm(d) / max(np.linalg.norm(ref), 1e-30)))
        if err > worst:
            worst, worst_name = err, n
        if err > tol:
            fails.append((n, err))
    return fails, worst, worst_name, float(np.median(l2s)), masked


def oracle_grads(P, x, xm, y, ym, drop_seed=None, dtype=torch.float64, gates=None):
    """(cost, {name: d cost / d name}, names) of the oracle's training-mode forward (no L2 term); gates: see engine_gates."""
    Pt = O.to_torch(P, dtype, requires_grad=True)
    dfn = O.hash_drop_fn(drop_seed) if drop_seed is not None else None
    ann, am, _ = O.dense_net(Pt, torch.as_tensor(x, dtype=dtype), torch.as_tensor(xm, dtype=dtype), True, dfn, gates=gates)
    cost = O.get_cost(Pt, ann, torch.as_tensor(y), am, torch.as_tensor(ym, dtype=dtype), dfn)
    names = O.trainable_names(Pt)
    grads = torch.autograd.grad(cost, [Pt[n] for n in names])
    return float(cost.detach()), {n: g.numpy() for n, g in zip(names, grads)}, names


def engine_grads(eng):
    return eng.P.unpack(eng.grads)
