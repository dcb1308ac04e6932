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


# Histogram of caption lengths (tokens, without <eol>) of the reference's data/train_caption.txt (8 835 captions: median 13,
# p90 32, max 96) - the empirical distribution SURVEY 8d cfg3 asks the "CROHME-dist" generator to draw from.  Index = length.
CAPTION_LENGTH_COUNTS = [
    0, 498, 288, 449, 365, 473, 391, 436, 307, 348, 268, 324, 263, 301, 179, 195, 262, 209, 195, 250, 216, 225, 189, 300, 148, 210, 155, 45,
    115, 84, 74, 99, 129, 70, 70, 109, 59, 70, 38, 27, 38, 29, 45, 24, 2, 38, 26, 20, 11, 19, 5, 3, 30, 2, 19, 1, 30, 0, 0, 18, 1, 3, 5, 0, 1,
    1, 3, 1, 1, 1, 1, 0, 1, 0, 0, 14, 0, 0, 1, 0, 0, 1, 0, 2, 2, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]


def synth_crohme(seed: int, n_images: int, max_area: int = 400000):
    """SURVEY 8d cfg3 / cfg5 "CROHME-dist" corpus: caption length l ~ empirical lengths of train_caption.txt;
    H = clip(round(96 * lognormal(0, 0.45)), 32, 320); W = clip(round(H * (0.6 + 0.42 l) * lognormal(0, 0.25)), 48, 2000);
    images with H * W > max_area are rejected (model-single-GPU.py:553-554).  Returns (images uint8 [1,H,W], captions)."""
    rng = np.random.default_rng(seed)
    p = np.asarray(CAPTION_LENGTH_COUNTS, dtype=np.float64); p /= p.sum()
    images, captions = [], []
    while len(images) < n_images:
        l = int(rng.choice(len(p), p=p))
        H = int(np.clip(round(96 * rng.lognormal(0.0, 0.45)), 32, 320))
        W = int(np.clip(round(H * (0.6 + 0.42 * l) * rng.lognormal(0.0, 0.25)), 48, 2000))
        if H * W > max_area:
            continue
        images.append(synth_image(rng, H, W))
        captions.append([int(t) for t in rng.integers(1, VOCAB, size=l)])
    return images, captions


def crohme_batches(seed: int, n_batches: int, batch_size: int = 16):
    """cfg3 batching: draw n_batches * batch_size images, sort by area (data.py:33-37), cut into batches of `batch_size`
    (the benchmark batches bypass the reference loader's 400 000-px padded-area cap per batch, data.py:54,61 - SURVEY 8d)."""
    images, captions = synth_crohme(seed, n_batches * batch_size)
    order = sorted(range(len(images)), key=lambda i: images[i].shape[1] * images[i].shape[2])
    return [([images[i] for i in order[b * batch_size:(b + 1) * batch_size]], [captions[i] for i in order[b * batch_size:(b + 1) * batch_size]])
            for b in range(n_batches)]
