"""Runs the reference's OWN model classes (model-single-GPU.py, read from /root/reference at test time) on top of
oracle/tf16_shim.py - an eager torch-fp64 stand-in for the TF-1.6 ops they call - and compares them with oracle/wap_oracle.py.

This pins the STRUCTURE of the oracle (what is convolved with what, slices, gate order, where the biases sit, maxout pairing,
embedding shift, scan carry, masks, which variables get which TF names in which order) to the reference source itself.  The
semantics of the individual TF ops stay a documented restatement (the shim), cross-checked against brute-force loops in
tests/test_oracle_cpu.py.  CPU only; /root/reference is not needed on the GPU box."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from helpers import O, make_batch, rel_err

REF = "/root/reference/model-single-GPU.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    from oracle import tf16_shim as tf
    tf.install()
    sys.path.insert(0, "/root/reference")
    try:
        spec = importlib.util.spec_from_file_location("ref_model_single_gpu", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)                      # classes + initialisers only: main() sits behind the __main__ guard
        yield mod, tf
    finally:
        sys.path.remove("/root/reference")
        sys.modules.pop("data", None)                     # the reference's data.py must not shadow anything afterwards
        tf.uninstall()


def _build(mod, tf, P, training, drop=None, requires_grad=False):
    tf.STATE.reset(preset=P, drop=drop, requires_grad=requires_grad)
    watcher = mod.Watcher_train(blocks=3, level=16, growth_rate=24, training=training)
    attender = mod.Attender(684, 256, 512)
    parser = mod.Parser(256, 256, attender, 684)
    wap = mod.WAP(watcher, attender, parser, 256, 256, 684, 111, training)
    return watcher, attender, parser, wap


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def test_variable_names_and_creation_order(ref):
    """The variables the reference code creates (tf.Variable names + tf.layers default names) are exactly the oracle's / the
    library's parameter set, and every one is found in the preset (i.e. nothing was initialised randomly)."""
    mod, tf = ref
    P = O.lively_params(1)
    x, xm, y, ym = make_batch(3, 1, 32, 48, 2, False)
    watcher, attender, parser, wap = _build(mod, tf, P, False)
    with torch.no_grad():
        watcher.dense_net(t64(x), t64(xm))
    assert set(tf.STATE.variables) == set(P), set(tf.STATE.variables) ^ set(P)
    assert tf.STATE.n_conv == 99 and tf.STATE.n_bn == 99 and tf.STATE.n_drop == 98


@pytest.mark.parametrize("training,drop_seed", [(False, None), (True, None), (True, 7)])
def test_dense_net_and_get_cost_match_oracle(ref, training, drop_seed):
    """Watcher_train.dense_net + WAP.get_cost of the reference == oracle.dense_net + oracle.get_cost (fp64): annotation, mask,
    cost and - in training mode - the moving-statistic updates, with and without (hash-keyed) dropout."""
    mod, tf = ref
    P = O.lively_params(5)
    x, xm, y, ym = make_batch(21, 2, 40, 56, 4, True)
    drop = O.hash_drop_fn(drop_seed) if drop_seed is not None else None
    watcher, attender, parser, wap = _build(mod, tf, P, training, drop)
    with torch.no_grad():
        ann, am = watcher.dense_net(t64(x), t64(xm))
        cost = wap.get_cost(ann, torch.as_tensor(y), am, t64(ym))
        Pt = O.to_torch(P)
        o_ann, o_am, o_moving = O.dense_net(Pt, t64(x), t64(xm), training, drop)
        o_cost = O.get_cost(Pt, o_ann, torch.as_tensor(y), o_am, t64(ym), drop if training else None)
    assert tuple(ann.shape) == tuple(o_ann.shape) and torch.equal(am, o_am)
    assert rel_err(ann.numpy(), o_ann.numpy()) < 1e-10
    assert abs(float(cost) - float(o_cost)) < 1e-10 * abs(float(o_cost))
    if training:
        assert set(tf.STATE.moving_updates) == set(o_moving)
        for n, v in o_moving.items():
            assert rel_err(tf.STATE.moving_updates[n].numpy(), v.numpy()) < 1e-10, n


def test_get_context_and_one_time_step_match_oracle(ref):
    """Attender.get_context and Parser.one_time_step / get_ht_ctx of the reference == the oracle, with both masks."""
    mod, tf = ref
    P = O.lively_params(2)
    watcher, attender, parser, wap = _build(mod, tf, P, False)
    Pt = O.to_torch(P)
    rs = np.random.RandomState(4)
    B, h, w, T = 3, 4, 6, 3
    ann = t64(np.maximum(rs.randn(B, h, w, 684), 0) * 0.5)
    am = torch.ones(B, h, w, dtype=torch.float64); am[1, 2:, :] = 0; am[2, :, 4:] = 0
    hp = t64(np.tanh(rs.randn(B, 256))); ap = t64(rs.rand(B, h, w) * 0.2)
    emb = t64(rs.randn(T, B, 256) * 0.3)
    ym = torch.ones(T, B, dtype=torch.float64); ym[2, 1] = 0
    with torch.no_grad():
        for mask in (None, am):
            c, a, apn = attender.get_context(ann, hp, ap.clone(), mask)          # the reference updates alpha_past in place
            oc, oa, oap = O.get_context(Pt, ann, hp, ap, mask)
            assert rel_err(c.numpy(), oc.numpy()) < 1e-12 and rel_err(a.numpy(), oa.numpy()) < 1e-12 and rel_err(apn.numpy(), oap.numpy()) < 1e-12
        hs, cs, als, aps, anns, ams = parser.get_ht_ctx(emb, hp, ann, am, ym)
        assert tuple(anns.shape) == (T, B, h, w, 684)                             # tf.scan stacks every carry, also the annotation
        oh, oap = hp, torch.zeros(B, h, w, dtype=torch.float64)
        for t in range(T):
            oh, oc, oa, oap = O.one_time_step(Pt, oh, oap, ann, am, emb[t], ym[t])
            assert rel_err(hs[t].numpy(), oh.numpy()) < 1e-12 and rel_err(cs[t].numpy(), oc.numpy()) < 1e-12
            assert rel_err(als[t].numpy(), oa.numpy()) < 1e-12 and rel_err(aps[t].numpy(), oap.numpy()) < 1e-12


def test_get_word_and_full_beam_search_of_the_reference(ref):
    """WAP.get_word (S:379-439; it reads the module global `wap`, which main() would have to provide) against oracle.get_word, and
    then the reference's get_sample driven by the reference's get_word - no oracle code in the loop - against oracle.beam_decode."""
    mod, tf = ref
    P = O.lively_params(6)
    watcher, attender, parser, wap = _build(mod, tf, P, False)
    mod.wap = wap
    Pt = O.to_torch(P)
    rs = np.random.RandomState(8)
    h, w = 4, 7
    ann1 = t64(np.maximum(rs.randn(1, h, w, 684), 0) * 0.5)
    with torch.no_grad():
        h0 = O.hidden_state_0(Pt, ann1)
        # first step: y = -1 => zero embedding, one row
        p, _, hn, apn = wap.get_word(torch.tensor([-1]), h0, torch.zeros(1, h, w, dtype=torch.float64), ann1)
        op, oh, oap, _ = O.get_word(Pt, torch.tensor([-1]), h0, torch.zeros(1, h, w, dtype=torch.float64), ann1)
        assert rel_err(p.numpy(), op.numpy()) < 1e-12 and rel_err(hn.numpy(), oh.numpy()) < 1e-12 and rel_err(apn.numpy(), oap.numpy()) < 1e-12
        # later step: k = 3 live hypotheses sharing the annotation
        y3 = torch.tensor([5, 17, 88]); h3 = t64(np.tanh(rs.randn(3, 256))); ap3 = t64(rs.rand(3, h, w) * 0.3)
        p, _, hn, apn = wap.get_word(y3, h3, ap3.clone(), ann1.expand(3, -1, -1, -1))
        op, oh, oap, _ = O.get_word(Pt, y3, h3, ap3, ann1.expand(3, -1, -1, -1))
        assert rel_err(p.numpy(), op.numpy()) < 1e-12 and rel_err(hn.numpy(), oh.numpy()) < 1e-12 and rel_err(apn.numpy(), oap.numpy()) < 1e-12

    class Session:                                             # session.run([p, w, h, alpha], feed_dict) served by the reference's get_word
        def run(self, fetches, feed_dict):
            with torch.no_grad():
                p, wd, hh, ap = wap.get_word(torch.as_tensor(feed_dict[mod.infer_y]), t64(feed_dict[mod.h_pre]),
                                             t64(feed_dict[mod.alpha_past]), t64(feed_dict[mod.anno]))
            return p.to(torch.float32).numpy(), wd.numpy(), hh.numpy(), ap.numpy()

    for name in ("anno", "infer_y", "alpha_past", "h_pre", "if_trainning"):       # the feed_dict keys get_sample looks up as globals
        setattr(mod, name, name)
    mod.argmax = False
    for k, maxlen in ((10, 10), (3, 7)):
        samples, scores = wap.get_sample("p", "w", "h", "alpha", ann1.numpy().astype(np.float32), h0.numpy(), k, maxlen, False, Session(), False)
        best, o_samples, o_scores = O.beam_decode(Pt, t64(ann1.numpy().astype(np.float32)), k, maxlen)
        assert [list(map(int, s)) for s in samples] == [list(map(int, s)) for s in o_samples]
        assert np.allclose(np.array(scores, np.float64), np.array(o_scores, np.float64), rtol=1e-6, atol=1e-6)


def test_reference_golden_is_reproducible_and_matches_oracle(ref):
    """tests/golden/ref_graph_cfg1.npz (what the GPU box compares the CUDA path with) is exactly what oracle/make_reference_golden.py
    produces from the reference's classes today, and the oracle reproduces it."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    try:
        import make_reference_golden as G
    finally:
        sys.path.pop(0)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_graph_cfg1.npz"))
    out = G.generate()
    assert set(out) == set(g.files)
    for k in g.files:
        assert np.array_equal(np.asarray(out[k]), g[k]), k
    # the oracle against the fixture
    x, xm, y, ym = make_batch(int(g["data_seed"]), 1, 64, 256, int(g["cap_len"]), False)
    P = O.calibrate_moving(O.lively_params(int(g["param_seed"])), x, xm)
    Pt = O.to_torch(P)
    with torch.no_grad():
        ann, am, _ = O.dense_net(Pt, t64(x), t64(xm), False)
        assert rel_err(ann.numpy(), g["annotation"]) < 1e-6
        assert abs(float(O.get_cost(Pt, ann, torch.as_tensor(y), am, t64(ym))) - float(g["cost"])) < 1e-9 * abs(float(g["cost"]))
        best = O.beam_decode(Pt, t64(ann.numpy().astype(np.float32)), int(g["beam_k"]), int(g["beam_maxlen"]))[0]
        assert [int(t) for t in best] == [int(t) for t in g["beam_best"] if t >= 0]
        drop = O.hash_drop_fn(int(g["drop_seed"]))
        ann_t, am_t, nm = O.dense_net(Pt, t64(x), t64(xm), True, drop)
        assert rel_err(ann_t.numpy(), g["train_annotation"]) < 1e-6
        assert abs(float(O.get_cost(Pt, ann_t, torch.as_tensor(y), am_t, t64(ym), drop)) - float(g["train_cost"])) < 1e-9 * abs(float(g["train_cost"]))
        for k in g.files:
            if k.startswith("new::"):
                assert rel_err(nm[k[5:]].numpy(), g[k]) < 1e-6, k


def test_gradients_of_the_reference_graph_match_oracle(ref):
    """tf.gradients stand-in: torch autograd THROUGH the reference's dense_net + get_cost (training mode, hash-keyed dropout) gives the
    same gradient for every one of the 329 trainable variables as autograd through the oracle - the gradients the CUDA path is tested
    against (test_full_gradients).  Tensors whose gradient vanishes identically (V_a_b: the softmax is shift-invariant) are compared on
    the global scale."""
    mod, tf = ref
    P = O.lively_params(5)
    x, xm, y, ym = make_batch(21, 2, 40, 56, 4, True)
    drop = O.hash_drop_fn(7)
    watcher, attender, parser, wap = _build(mod, tf, P, True, drop, requires_grad=True)
    ann, am = watcher.dense_net(t64(x), t64(xm))
    cost = wap.get_cost(ann, torch.as_tensor(y), am, t64(ym))
    names = [n for n, v in tf.STATE.variables.items() if v.requires_grad]
    grads = dict(zip(names, torch.autograd.grad(cost, [tf.STATE.variables[n] for n in names])))
    Pt = O.to_torch(P, torch.float64, requires_grad=True)
    o_ann, o_am, _ = O.dense_net(Pt, t64(x), t64(xm), True, drop)
    o_cost = O.get_cost(Pt, o_ann, torch.as_tensor(y), o_am, t64(ym), drop)
    o_names = O.trainable_names(Pt)
    o_grads = dict(zip(o_names, torch.autograd.grad(o_cost, [Pt[n] for n in o_names])))
    assert set(names) == set(o_names) and len(names) == 329
    assert abs(float(cost.detach()) - float(o_cost.detach())) < 1e-10 * abs(float(o_cost.detach()))
    gmax = max(float(g.abs().max()) for g in o_grads.values())
    for n in names:
        scale = max(float(o_grads[n].abs().max()), 1e-6 * gmax)
        assert float((grads[n] - o_grads[n]).abs().max()) / scale < 1e-9, n


@pytest.mark.parametrize("B,H,W", [(2, 37, 75), (1, 33, 65), (3, 18, 50)])
def test_dense_net_on_odd_and_ragged_sizes(ref, B, H, W):
    """Odd heights / widths exercise the asymmetric stem padding, the ceil-mode pools, the valid-count average pool and the
    `[:, 0::2, 0::2]` mask subsampling of the reference (S:80, 87, 130) against the oracle; ragged batches exercise the masks."""
    mod, tf = ref
    P = O.lively_params(4)
    x, xm, y, ym = make_batch(31 + B, B, H, W, 3, True)
    for training in (False, True):
        watcher, attender, parser, wap = _build(mod, tf, P, training)
        with torch.no_grad():
            ann, am = watcher.dense_net(t64(x), t64(xm))
            cost = wap.get_cost(ann, torch.as_tensor(y), am, t64(ym))
            Pt = O.to_torch(P)
            o_ann, o_am, _ = O.dense_net(Pt, t64(x), t64(xm), training)
            o_cost = O.get_cost(Pt, o_ann, torch.as_tensor(y), o_am, t64(ym))
        assert tuple(ann.shape) == tuple(o_ann.shape) == (B, -(-H // 16), -(-W // 16), 684) and torch.equal(am, o_am)
        assert rel_err(ann.numpy(), o_ann.numpy()) < 1e-10 and abs(float(cost) - float(o_cost)) < 1e-10 * abs(float(o_cost))
