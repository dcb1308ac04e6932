"""Loader with the reference's data.py interface (/root/reference/data.py:7-122):

    dataIterator(feature_file, label_file, dictionary, batch_size, batch_Imagesize, maxlen, maxImagesize)
        -> (list[(images, labels)], uidList)
    load_dict(dictFile) -> {token: id}
    prepare_data(images_x, seqs_y) -> (x, x_mask, y, y_mask)

Same inputs (.pkl dict uid -> uint8 [1,H,W]; caption file `uid tok tok ...`; dictionary.txt), same
outputs and the same bucketing decisions, re-implemented (not copied).  `prepare_data` additionally
accepts `pinned=True` to fill page-locked torch buffers for the asynchronous host->device copy.
"""
from __future__ import annotations

import pickle
import sys
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np


def load_dict(dictFile: str) -> Dict[str, int]:
    """dictionary.txt: one `token id` pair per line (data.py:87-97)."""
    lexicon: Dict[str, int] = {}
    with open(dictFile) as fh:
        for line in fh:
            parts = line.split()
            if len(parts) >= 2:
                lexicon[parts[0]] = int(parts[1])
    print('total words/phones', len(lexicon))
    return lexicon


def _read_captions(label_file: str, dictionary: Dict[str, int]) -> Dict[str, List[int]]:
    targets: Dict[str, List[int]] = {}
    with open(label_file) as fh:
        for line in fh:
            parts = line.strip().split()
            if not parts:
                continue
            uid, words = parts[0], parts[1:]
            ids = []
            for wd in words:
                if wd not in dictionary:                      # data.py:26-28: hard exit on OOV
                    print('a word not in the dictionary !! sentence ', uid, 'word ', wd)
                    sys.exit()
                ids.append(dictionary[wd])
            targets[uid] = ids
    return targets


def _load_features(feature_file: str):
    with open(feature_file, 'rb') as fh:
        try:
            return pickle.load(fh)
        except UnicodeDecodeError:                            # python-2 pickles of the original CROHME dumps
            fh.seek(0)
            return pickle.load(fh, encoding='latin1')


def bucket_by_area(features: Dict[str, np.ndarray], targets: Dict[str, List[int]], batch_size: int,
                   batch_Imagesize: int, maxlen: int, maxImagesize: int):
    """Area-sorted greedy bucketing with the reference's exact decisions (data.py:33-84):
    a batch closes when (largest area so far) x (count+1) exceeds `batch_Imagesize` or when it
    already holds `batch_size` images; over-long captions / over-large images are dropped, but
    still advance the running maximum, as in the reference."""
    order = sorted(((uid, im.shape[1] * im.shape[2]) for uid, im in features.items()), key=lambda kv: kv[1])
    batches: List[Tuple[list, list]] = []
    uid_list: List[str] = []
    cur_x: list = []
    cur_y: list = []
    biggest = 0
    for uid, area in order:
        biggest = max(biggest, area)
        projected = biggest * (len(cur_x) + 1)
        label = targets[uid]
        if len(label) > maxlen:
            print('sentence', uid, 'length bigger than', maxlen, 'ignore')
            continue
        if area > maxImagesize:
            print('image', uid, 'size bigger than', maxImagesize, 'ignore')
            continue
        uid_list.append(uid)
        if projected > batch_Imagesize or len(cur_x) == batch_size:
            batches.append((cur_x, cur_y))
            cur_x, cur_y, biggest = [], [], area
        cur_x.append(features[uid])
        cur_y.append(label)
    batches.append((cur_x, cur_y))
    return batches, uid_list


def dataIterator(feature_file, label_file, dictionary, batch_size, batch_Imagesize, maxlen, maxImagesize):
    features = _load_features(feature_file)
    targets = _read_captions(label_file, dictionary)
    batches, uid_list = bucket_by_area(features, targets, batch_size, batch_Imagesize, maxlen, maxImagesize)
    print('total ', len(batches), 'batch data loaded')
    return batches, uid_list


def prepare_data(images_x: Sequence[np.ndarray], seqs_y: Sequence[Sequence[int]], n_words_src=30000, n_words=30000,
                 pinned: bool = False):
    """Pad to the batch maximum, scale to [0,1], build masks (data.py:100-122).

    x [B,H,W,1] f32, x_mask [B,H,W] f32, y [T,B] int64 (trailing 0 = <eol>), y_mask [T,B] f32 with
    T = max caption length + 1 and y_mask covering len+1 positions."""
    B = len(images_x)
    H = max(im.shape[1] for im in images_x)
    W = max(im.shape[2] for im in images_x)
    T = max(len(s) for s in seqs_y) + 1
    if pinned:
        import torch
        tx = torch.zeros((B, H, W, 1), dtype=torch.float32).pin_memory()
        txm = torch.zeros((B, H, W), dtype=torch.float32).pin_memory()
        ty = torch.zeros((T, B), dtype=torch.int64).pin_memory()
        tym = torch.zeros((T, B), dtype=torch.float32).pin_memory()
        x, x_mask, y, y_mask = tx.numpy(), txm.numpy(), ty.numpy(), tym.numpy()
    else:
        x = np.zeros((B, H, W, 1), np.float32)
        x_mask = np.zeros((B, H, W), np.float32)
        y = np.zeros((T, B), np.int64)
        y_mask = np.zeros((T, B), np.float32)
    for b, (im, seq) in enumerate(zip(images_x, seqs_y)):
        hh, ww, n = im.shape[1], im.shape[2], len(seq)
        x[b, :hh, :ww, 0] = im[0] / 255.
        x_mask[b, :hh, :ww] = 1.
        y[:n, b] = seq
        y_mask[:n + 1, b] = 1.
    if pinned:
        return tx, txm, ty, tym
    return x, x_mask, y, y_mask


class PackedBatch:
    """One batch staged for the GPU-side `prepare_data` (wap_prepare_batch): ONE host buffer
    [offs int64 x B | hw int32 x 2B | y int32 x T*B | y_mask f32 x T*B | pad to 16 | raw uint8 pixels of the B images back to back]
    so that a step needs a single H2D copy of B*H*W bytes + a few hundred bytes instead of the padded fp32 x and x_mask
    (8 bytes per padded pixel).  /255, NHWC placement and both masks are produced on the device."""
    __slots__ = ("buf", "B", "H", "W", "T", "off_hw", "off_y", "off_ym", "off_pix", "nbytes")

    def __init__(self, buf, B, H, W, T, off_hw, off_y, off_ym, off_pix):
        self.buf, self.B, self.H, self.W, self.T = buf, B, H, W, T
        self.off_hw, self.off_y, self.off_ym, self.off_pix = off_hw, off_y, off_ym, off_pix
        self.nbytes = int(buf.numel() if hasattr(buf, "numel") else buf.size)


def pack_batch(images_x: Sequence[np.ndarray], seqs_y: Sequence[Sequence[int]], pinned: bool = False) -> PackedBatch:
    """Host half of the device-side prepare_data: same inputs as `prepare_data` (uint8 [1,H,W] images, id lists)."""
    B = len(images_x)
    H = max(im.shape[1] for im in images_x)
    W = max(im.shape[2] for im in images_x)
    T = max(len(s) for s in seqs_y) + 1
    off_hw = 8 * B
    off_y = off_hw + 8 * B
    off_ym = off_y + 4 * T * B
    off_pix = (off_ym + 4 * T * B + 15) & ~15
    npix = sum(int(im.shape[1]) * int(im.shape[2]) for im in images_x)
    total = off_pix + npix
    if pinned:
        import torch
        tbuf = torch.empty(total, dtype=torch.uint8).pin_memory()
        buf = tbuf.numpy()
    else:
        tbuf = None
        buf = np.empty(total, np.uint8)
    offs = buf[:off_hw].view(np.int64); hw = buf[off_hw:off_y].view(np.int32)
    y = buf[off_y:off_ym].view(np.int32).reshape(T, B); ym = buf[off_ym:off_ym + 4 * T * B].view(np.float32).reshape(T, B)
    y[:] = 0; ym[:] = 0.
    buf[off_ym + 4 * T * B: off_pix] = 0                 # alignment padding: deterministic bytes
    o = 0
    for b, (im, seq) in enumerate(zip(images_x, seqs_y)):
        hh, ww, n = int(im.shape[1]), int(im.shape[2]), len(seq)
        if im.dtype != np.uint8:
            raise TypeError("pack_batch expects the loader's uint8 images")
        offs[b] = o; hw[2 * b] = hh; hw[2 * b + 1] = ww
        buf[off_pix + o: off_pix + o + hh * ww] = np.ascontiguousarray(im[0]).reshape(-1)
        o += hh * ww
        if n and (min(seq) < 0 or max(seq) >= 111):
            raise ValueError("token id out of range [0, 111)")
        y[:n, b] = seq
        ym[:n + 1, b] = 1.
    return PackedBatch(tbuf if pinned else buf, B, H, W, T, off_hw, off_y, off_ym, off_pix)


class Prefetcher:
    """Runs `prepare_data` (padding, /255, masks, optional pinned staging) for the next batches on a background thread while
    the GPU works on the current one (the reference prepares every batch synchronously inside the step loop,
    model-single-GPU.py:637-640).  Iterating yields (x, x_mask, y, y_mask) in the order of `batches`; empty batches are skipped;
    an exception in the worker is re-raised in the consumer.  depth = number of prepared batches kept ahead."""

    _END = object()

    def __init__(self, batches: Iterable[Tuple[Sequence[np.ndarray], Sequence[Sequence[int]]]], pinned: bool = False, depth: int = 2,
                 packed: bool = False):
        """packed: yield `PackedBatch` objects (raw bytes for the GPU-side prepare_data) instead of the padded host tensors."""
        import queue
        import threading
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, int(depth)))
        self._stop = threading.Event()
        self._prep = (lambda bx, by, pinned: pack_batch(bx, by, pinned=pinned)) if packed else \
            (lambda bx, by, pinned: prepare_data(bx, by, pinned=pinned))
        self._thread = threading.Thread(target=self._work, args=(iter(batches), pinned), daemon=True)
        self._thread.start()

    def _put(self, item) -> bool:
        import queue
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _work(self, it, pinned):
        try:
            for batch_x, batch_y in it:
                if not batch_x:
                    continue
                if not self._put(self._prep(batch_x, batch_y, pinned)):
                    return
            self._put(self._END)
        except BaseException as e:                         # noqa: BLE001 - handed to the consumer
            self._put(e)

    def __iter__(self):
        while True:
            item = self._q.get()
            if item is self._END:
                return
            if isinstance(item, BaseException):
                raise item
            yield item

    def close(self):
        """Stop early (e.g. on early stopping): the worker exits at its next queue operation."""
        self._stop.set()
