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


def kink_excluded(P, x, xm, drop_seed=None, tau=5e-5, max_channels=64):
    """Entries of the parameter gradients that a legitimate fp32-vs-fp64 ReLU gating difference may move (see
    oracle.kink_channels): {variable name: channel indices along its LAST axis}.  Everything else must match the oracle on
    THIS batch - there is no retry on another batch.  A cap on the number of excluded channels (of ~6000 BN channels) keeps
    the exclusion from hiding a broad failure."""
    drop = O.hash_drop_fn(drop_seed) if drop_seed is not None else None
    kc = O.kink_channels(P, x, xm, drop, tau)
    n = sum(len(v) for v in kc.values())
    assert n <= max_channels, f"{n} near-kink channels: choose another seed for this test batch"
    ex = {}
    for k, cs in kc.items():
        for name in (O.conv_name(k), O.bn_name(k, "gamma"), O.bn_name(k, "beta")):
            ex[name] = np.asarray(cs, dtype=np.int64)
    return ex


def grad_errors(names, got_by_name, ref_by_name, excluded, tol=2e-3):
    """Per-tensor max-norm error (tensors whose gradient is tiny relative to the largest one are compared on the global
    scale: they are dominated by fp32 rounding of large cancelling sums), with the near-kink channels masked out.
    Returns (failures, worst_err, worst_name, median_rel_l2, n_masked_entries)."""
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


def oracle_grads(P, x, xm, y, ym, drop_seed=None, dtype=torch.float64):
    """(cost, {name: d cost / d name}) of the oracle's training-mode forward (no L2 term)."""
    Pt = O.to_torch(P, dtype, requires_grad=True)
    dfn = O.hash_drop_fn(drop_seed) if drop_seed is not None else None
    ann, am, _ = O.dense_net(Pt, torch.as_tensor(x, dtype=dtype), torch.as_tensor(xm, dtype=dtype), True, dfn)
    cost = O.get_cost(Pt, ann, torch.as_tensor(y), am, torch.as_tensor(ym, dtype=dtype), dfn)
    names = O.trainable_names(Pt)
    grads = torch.autograd.grad(cost, [Pt[n] for n in names])
    return float(cost.detach()), {n: g.numpy() for n, g in zip(names, grads)}, names


def engine_grads(eng):
    return eng.P.unpack(eng.grads)
