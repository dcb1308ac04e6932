"""WER / ExpRate as the reference computes them (model-distributed.py:30-78; model-single-GPU.py:698
shells out to a `compute-wer.py` that is not shipped): token-level edit distance summed over
expressions / total label length, and the fraction of exactly-matching expressions."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def cmp_result(label: Sequence, rec: Sequence) -> Tuple[int, int]:
    """Levenshtein distance (substitution/insertion/deletion cost 1) and the label length."""
    n, m = len(label), len(rec)
    prev = np.arange(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        cur = np.empty(m + 1, dtype=np.int64)
        cur[0] = i
        for j in range(1, m + 1):
            cur[j] = min(prev[j - 1] + (label[i - 1] != rec[j - 1]), cur[j - 1] + 1, prev[j] + 1)
        prev = cur
    return int(prev[m]), n


def _read(path: str) -> Dict[str, List[str]]:
    out: Dict[str, List[str]] = {}
    with open(path) as fh:
        for line in fh:
            parts = line.split()
            if parts:
                out[parts[0]] = parts[1:]
    return out


def score(rec: Dict[str, List[str]], label: Dict[str, List[str]]) -> Tuple[float, float]:
    dist_total = label_total = lines = exact = 0
    for key, hyp in rec.items():
        d, n = cmp_result(label[key], hyp)
        dist_total += d; label_total += n; lines += 1; exact += (d == 0)
    return dist_total / label_total, exact / lines


def process(recfile: str, labelfile: str, resultfile: str) -> Tuple[float, float]:
    """Same file contract as the reference's scorer: writes 'WER x' and 'ExpRate y' lines."""
    wer, sacc = score(_read(recfile), _read(labelfile))
    with open(resultfile, "w") as fh:
        fh.write('WER {}\n'.format(wer))
        fh.write('ExpRate {}\n'.format(sacc))
    return wer, sacc
