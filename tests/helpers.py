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


def engine_with(P_np):
    import wap_b200
    eng = wap_b200.WAPEngine(init=False)
    eng.P.load(P_np)
    return eng


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def well_posed_batch(P, seed, B, H, W, cap_len, ragged, drop_seed=None, margin=1.5e-6, tries=80):
    """First batch (seeds seed, seed+1000, ...) none of whose ~1e6 BN outputs lies within `margin`
    (a few fp32 ulps of an O(1) value) of the ReLU kink (see oracle.relu_kink_margin).  An element
    on the kink is gated differently by fp32 and fp64 arithmetic - legitimately - and with only a few
    dozen pixels per channel in block 2 of these small test images one flipped element moves that
    channel's gradients by ~1e-2, which says nothing about the kernels."""
    for i in range(tries):
        x, xm, y, ym = make_batch(seed + 1000 * i, B, H, W, cap_len, ragged)
        m = O.relu_kink_margin(P, x, xm, O.hash_drop_fn(drop_seed) if drop_seed is not None else None)
        if m > margin:
            return x, xm, y, ym
    raise RuntimeError("no well-posed batch found")
