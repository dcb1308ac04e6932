"""Deterministic synthetic CROHME-shaped inputs (the real offline-{train,test}.pkl are git-LFS
pointers in the reference checkout): uint8 [1,H,W] images with ~10 % ink and uniformly random
captions over token ids 1..110."""
from __future__ import annotations

import numpy as np

VOCAB = 111


def synth_image(rng: np.random.Generator, H: int, W: int) -> np.ndarray:
    return ((rng.random((H, W)) < 0.10) * 255).astype(np.uint8)[None]


def synth_batch(seed: int, B: int, H: int, W: int, caption_len: int, ragged: bool = False):
    """B images (all H x W unless `ragged`) and B captions of `caption_len` tokens."""
    rng = np.random.default_rng(seed)
    images, captions = [], []
    for i in range(B):
        shrink = ragged and i > 0
        h = H - (int(rng.integers(0, H // 3)) if shrink else 0)
        w = W - (int(rng.integers(0, W // 3)) if shrink else 0)
        n = caption_len - (int(rng.integers(0, max(caption_len // 2, 1))) if shrink else 0)
        images.append(synth_image(rng, h, w))
        captions.append([int(t) for t in rng.integers(1, VOCAB, size=n)])
    return images, captions
