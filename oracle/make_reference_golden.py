"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ref_graph_cfg1.npz by running the REFERENCE's own model classes
(/root/reference/model-single-GPU.py: Watcher_train.dense_net, WAP.get_cost, WAP.get_word, WAP.get_sample) on oracle/tf16_shim.py,
the eager stand-in for the TF-1.6 ops they call.  The reference cannot travel to the GPU box, the vectors can:
tests/test_gpu_parity.py::test_engine_matches_reference_generated_golden compares the CUDA path with them there, and
tests/test_reference_graph_cpu.py::test_reference_golden_is_reproducible regenerates them here.

    python oracle/make_reference_golden.py            # rewrites tests/golden/ref_graph_cfg1.npz (needs /root/reference)

Inputs are regenerated from seeds by the tests (helpers.make_batch / oracle.lively_params / oracle.calibrate_moving), so the file
only holds outputs.  Workload = SURVEY 8d cfg1: one 64 x 256 image, a 10-token caption."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tf16_shim as tf          # noqa: E402
from oracle import wap_oracle as O          # noqa: E402

REF = "/root/reference/model-single-GPU.py"
PARAM_SEED, DATA_SEED, CAP_LEN, DROP_SEED, BEAM_K, BEAM_MAXLEN = 3, 12, 10, 4242, 10, 12


def load_reference():
    tf.install()
    sys.path.insert(0, "/root/reference")
    spec = importlib.util.spec_from_file_location("ref_model_single_gpu", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(mod, P, training, drop=None):
    tf.STATE.reset(preset=P, drop=drop)
    watcher = mod.Watcher_train(blocks=3, level=16, growth_rate=24, training=training)
    attender = mod.Attender(684, 256, 512)
    parser = mod.Parser(256, 256, attender, 684)
    wap = mod.WAP(watcher, attender, parser, 256, 256, 684, 111, training)
    mod.wap = wap
    return watcher, wap


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def generate():
    from helpers import make_batch
    mod = load_reference()
    x, xm, y, ym = make_batch(DATA_SEED, 1, 64, 256, CAP_LEN, False)
    P = O.calibrate_moving(O.lively_params(PARAM_SEED), x, xm)
    out = dict(param_seed=PARAM_SEED, data_seed=DATA_SEED, cap_len=CAP_LEN, drop_seed=DROP_SEED, beam_k=BEAM_K, beam_maxlen=BEAM_MAXLEN)
    with torch.no_grad():
        # ---- inference graph (if_trainning = False): annotation, cost, beam search
        watcher, wap = build(mod, P, False)
        ann, am = watcher.dense_net(t64(x), t64(xm))
        out["annotation"] = ann.numpy().astype(np.float32)
        out["annotation_mask"] = am.numpy().astype(np.float32)
        out["cost"] = np.float64(wap.get_cost(ann, torch.as_tensor(y), am, t64(ym)))
        h0 = torch.tanh(torch.tensordot(ann.mean(dim=(1, 2)), t64(P["Wa2h"]), dims=1) + t64(P["ba2h"]))      # S:590

        class Session:
            def run(self, fetches, feed_dict):
                p, wd, hh, ap = wap.get_word(torch.as_tensor(feed_dict[mod.infer_y]), t64(feed_dict[mod.h_pre]),
                                             t64(feed_dict[mod.alpha_past]), t64(feed_dict[mod.anno]))
                return p.to(torch.float32).numpy(), wd.numpy(), hh.numpy(), ap.numpy()
        for name in ("anno", "infer_y", "alpha_past", "h_pre", "if_trainning"):
            setattr(mod, name, name)
        mod.argmax = False
        ann32 = ann.numpy().astype(np.float32)
        samples, scores = wap.get_sample("p", "w", "h", "alpha", ann32, h0.numpy(), BEAM_K, BEAM_MAXLEN, False, Session(), False)
        norm = np.array(scores) / np.array([len(s) for s in samples])                                          # S:671-672
        best = np.array(samples[int(norm.argmin())], np.int64)
        out["beam_best"] = np.concatenate([best, -np.ones(BEAM_MAXLEN + 1 - len(best), np.int64)])
        p0, _, h1, ap1 = wap.get_word(torch.tensor([-1]), h0, torch.zeros(ann.shape[:3], dtype=torch.float64), ann)
        out["probs0"] = p0.numpy().astype(np.float32); out["h1"] = h1.numpy().astype(np.float32); out["alpha_past1"] = ap1.numpy().astype(np.float32)
        # ---- training graph (if_trainning = True) with the hash-keyed dropout masks: annotation, cost, two moving statistics
        watcher, wap = build(mod, P, True, O.hash_drop_fn(DROP_SEED))
        ann_t, am_t = watcher.dense_net(t64(x), t64(xm))
        out["train_annotation"] = ann_t.numpy().astype(np.float32)
        out["train_cost"] = np.float64(wap.get_cost(ann_t, torch.as_tensor(y), am_t, t64(ym)))
        for name in ("batch_normalization/moving_mean", "batch_normalization_98/moving_variance", "batch_normalization_33/moving_variance"):
            out["new::" + name] = tf.STATE.moving_updates[name].numpy().astype(np.float32)
    return out


if __name__ == "__main__":
    out = generate()
    path = os.path.join(ROOT, "tests", "golden", "ref_graph_cfg1.npz")
    np.savez(path, **out)
    print("wrote", path, {k: (np.asarray(v).shape, str(np.asarray(v).dtype)) for k, v in out.items()})
