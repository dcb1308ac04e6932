"""End-to-end run of the drop-in entry script on a synthetic corpus (GPU)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("extra", [[], ["--greedy_valid", "--prefetch", "0", "--save", "ckpt.npz"]])
def test_model_single_gpu_entry_point(tmp_path, extra):
    """Default run (beam-search validation on the device, packed prefetch, graph replay for repeated shapes) and a run with the greedy
    two-stream validation pipeline, the reference's synchronous host prepare_data and a checkpoint written on improvement."""
    from wap_b200 import synth
    d = tmp_path; (d / "data").mkdir()
    vocab = [f"tok{i}" for i in range(1, 111)]
    (d / "data" / "dictionary.txt").write_text("".join(f"{t}\t{i + 1}\n" for i, t in enumerate(vocab)) + "<eol>\t0\n")
    for split, n, seed in (("train", 12, 1), ("test", 3, 2)):
        imgs, caps = synth.synth_batch(seed, n, 40, 96, 6, ragged=True)
        feats = {f"{split}{i}": im for i, im in enumerate(imgs)}
        (d / "data" / f"offline-{split}.pkl").write_bytes(pickle.dumps(feats))
        (d / "data" / f"{split}_caption.txt").write_text("".join(f"{split}{i}\t" + " ".join(vocab[t - 1] for t in c) + "\n" for i, c in enumerate(caps)))
    extra = [str(d / a) if a.endswith(".npz") else a for a in extra]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "model-single-GPU.py"), "--path", str(d), "--batch_size", "4", "--max_epoch", "2"] + extra,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    log = (d / "log-bs-6.txt").read_text()
    assert "valid set decode done" in log and "Valid WER:" in log
    dec = (d / "result" / "valid_decode_result-bs-6.txt").read_text().splitlines()
    assert [ln.split()[0] for ln in dec] == sorted(f"test{i}" for i in range(3)) or len(dec) == 3
    wer = (d / "result" / "valid-bs-6.wer").read_text()
    assert wer.startswith("WER ") and "ExpRate " in wer
    if any(a.endswith(".npz") for a in extra):
        z = np.load(d / "ckpt.npz")
        assert "__sched__/lrate" in z.files and "conv2d_98/kernel" in z.files
