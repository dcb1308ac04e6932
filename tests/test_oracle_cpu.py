"""Pins for the CPU oracle itself (no GPU): brute-force TF 'SAME' op semantics, batch-norm
semantics, fp64 finite-difference gradient checks, golden vectors, model-size constants, scorer."""
import os

import numpy as np
import pytest
import torch

from helpers import O, ROOT, t64

GOLDEN = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------ brute-force 'SAME' ops
def _conv_same_loops(x, w, s):
    B, H, W, Ci = x.shape; kh, kw, _, Co = w.shape
    Ho, pt, _ = O.same_pad(H, kh, s); Wo, pl, _ = O.same_pad(W, kw, s)
    y = np.zeros((B, Ho, Wo, Co))
    for b in range(B):
        for i in range(Ho):
            for j in range(Wo):
                for u in range(kh):
                    for v in range(kw):
                        yy, xx = i * s - pt + u, j * s - pl + v
                        if 0 <= yy < H and 0 <= xx < W:
                            y[b, i, j] += x[b, yy, xx] @ w[u, v]
    return y


@pytest.mark.parametrize("H,W,k,s", [(6, 7, 7, 2), (7, 6, 7, 2), (5, 4, 3, 1), (4, 5, 11, 1), (3, 3, 1, 1), (9, 8, 7, 2)])
def test_conv_same_matches_loops(H, W, k, s):
    rs = np.random.RandomState(H * 10 + W)
    x = rs.randn(2, H, W, 2); w = rs.randn(k, k, 2, 3)
    got = O.conv2d_same(t64(x), t64(w), s).numpy()
    assert got.shape == (2, -(-H // s), -(-W // s), 3)
    np.testing.assert_allclose(got, _conv_same_loops(x, w, s), atol=1e-10)


def test_stem_padding_is_asymmetric_on_even_sizes():
    # TF SAME, k=7, s=2, even input: 2 before / 3 after (torch padding=3 would be wrong)
    assert O.same_pad(128, 7, 2) == (64, 2, 3)
    assert O.same_pad(127, 7, 2) == (64, 3, 3)
    assert O.same_pad(9, 3, 1) == (9, 1, 1) and O.same_pad(9, 11, 1) == (9, 5, 5)


@pytest.mark.parametrize("H,W", [(4, 4), (5, 3), (1, 7), (6, 5)])
def test_pools_same_match_loops(H, W):
    rs = np.random.RandomState(H + 31 * W)
    x = rs.randn(2, H, W, 3)
    Ho, Wo = -(-H // 2), -(-W // 2)
    mx = np.full((2, Ho, Wo, 3), -np.inf); sm = np.zeros((2, Ho, Wo, 3)); cnt = np.zeros((Ho, Wo))
    for i in range(H):
        for j in range(W):
            mx[:, i // 2, j // 2] = np.maximum(mx[:, i // 2, j // 2], x[:, i, j])
            sm[:, i // 2, j // 2] += x[:, i, j]; cnt[i // 2, j // 2] += 1
    np.testing.assert_allclose(O.maxpool2_same(t64(x)).numpy(), mx, atol=1e-12)
    np.testing.assert_allclose(O.avgpool2_same(t64(x)).numpy(), sm / cnt[None, :, :, None], atol=1e-12)


def test_batch_norm_semantics():
    rs = np.random.RandomState(0)
    x = rs.randn(3, 4, 5, 6) * 2 + 1
    g, b = rs.randn(6), rs.randn(6); mm, mv = rs.randn(6), rs.rand(6) + 0.5
    y, (nm, nv) = O.batch_norm(t64(x), t64(g), t64(b), t64(mm), t64(mv), True)
    mean = x.reshape(-1, 6).mean(0); var = x.reshape(-1, 6).var(0)            # biased
    np.testing.assert_allclose(y.numpy(), (x - mean) / np.sqrt(var + 1e-4) * g + b, atol=1e-10)
    n = 3 * 4 * 5
    np.testing.assert_allclose(nm.numpy(), 0.9 * mm + 0.1 * mean, atol=1e-12)
    np.testing.assert_allclose(nv.numpy(), 0.9 * mv + 0.1 * var * n / (n - 1), atol=1e-12)   # Bessel-corrected (fused kernel)
    yi, _ = O.batch_norm(t64(x), t64(g), t64(b), t64(mm), t64(mv), False)
    np.testing.assert_allclose(yi.numpy(), (x - mm) / np.sqrt(mv + 1e-4) * g + b, atol=1e-10)


# ------------------------------------------------------------------ model-size constants
def test_model_sizes():
    P = O.init_params(0)
    tr = O.trainable_names(P)
    assert len(P) == 527 and len(tr) == 329
    assert sum(P[n].size for n in tr) == 5428520
    enc = [n for n in tr if n.startswith(("conv2d", "batch_normalization"))]
    assert sum(P[n].size for n in enc) == 2979384
    assert sum(P[n].size for n in O.l2_names(P)) == 5415872
    plan = O.encoder_plan()
    assert len(plan) == 99 and plan[33]["kind"] == "tr" and plan[66]["kind"] == "tr" and plan[33]["cin"] == 432 and plan[66]["cout"] == 300
    imgs, caps = O.synth_batch(1, 1, 64, 256, 10)
    x, xm, y, ym = O.prepare_data(imgs, caps)
    assert y.shape == (11, 1) and ym.sum() == 11
    Pt = O.to_torch(P)
    ann, am, _ = O.dense_net(Pt, t64(x), t64(xm), False)
    assert tuple(ann.shape) == (1, 4, 16, 684) and tuple(am.shape) == (1, 4, 16)


# ------------------------------------------------------------------ finite differences (fp64)
def _fd_check(f, tensors, names, n_dirs=3, eps=1e-6, seed=0):
    cost = f()
    grads = torch.autograd.grad(cost, [tensors[n] for n in names])
    rs = np.random.RandomState(seed)
    for _ in range(n_dirs):
        dirs = [torch.as_tensor(rs.randn(*tensors[n].shape)) for n in names]
        ana = sum(float((g * d).sum()) for g, d in zip(grads, dirs))
        with torch.no_grad():
            for n, d in zip(names, dirs): tensors[n].add_(eps * d)
            cp = float(f())
            for n, d in zip(names, dirs): tensors[n].sub_(2 * eps * d)
            cm = float(f())
            for n, d in zip(names, dirs): tensors[n].add_(eps * d)
        num = (cp - cm) / (2 * eps)
        assert abs(num - ana) <= 1e-5 * max(1.0, abs(ana)), (num, ana)


def test_fd_gradient_decoder():
    P = O.to_torch(O.lively_params(1), torch.float64, requires_grad=True)
    rs = np.random.RandomState(2)
    ann = t64(np.maximum(rs.randn(2, 3, 4, O.ANNOT_C), 0)).requires_grad_(True)
    am = torch.ones(2, 3, 4, dtype=torch.float64); am[1, :, 3] = 0
    y = torch.as_tensor(rs.randint(1, 111, size=(4, 2))); y[3, :] = 0; y[2, 1] = 0
    ym = torch.ones(4, 2, dtype=torch.float64); ym[3, 1] = 0
    tensors = dict(P); tensors["__ann"] = ann
    names = [n for n in O.trainable_names(P) if not n.startswith(("conv2d", "batch_normalization"))] + ["__ann"]
    # eps small enough that no maxout pair (S:357-359) crosses its tie inside the central-difference interval
    _fd_check(lambda: O.get_cost(P, ann, y, am, ym), tensors, names, eps=1e-8)


def test_fd_gradient_encoder_training_mode():
    P = O.to_torch(O.lively_params(2), torch.float64, requires_grad=True)
    imgs, caps = O.synth_batch(3, 2, 20, 36, 2, ragged=True)
    x, xm, _, _ = O.prepare_data(imgs, caps)
    rs = np.random.RandomState(4)
    proj = None

    def f():
        nonlocal proj
        ann, _, _ = O.dense_net(P, t64(x), t64(xm), True, O.hash_drop_fn(5))
        if proj is None:
            proj = t64(rs.randn(*ann.shape))
        return (ann * proj).sum()
    names = [n for n in O.trainable_names(P) if n.startswith(("conv2d", "batch_normalization"))]
    _fd_check(f, P, names, n_dirs=2, eps=1e-8)   # small enough that no BN output crosses the ReLU kink inside the interval


# ------------------------------------------------------------------ golden vectors
def test_golden_cfg1_forward():
    g = np.load(os.path.join(GOLDEN, "cfg1_forward.npz"))
    imgs, caps = O.synth_batch(int(g["data_seed"]), 1, 64, 256, 10)
    x, xm, y, ym = O.prepare_data(imgs, caps)
    P = O.calibrate_moving(O.lively_params(int(g["param_seed"])), x, xm)
    Pt = O.to_torch(P)
    ann, am, _ = O.dense_net(Pt, t64(x), t64(xm), False)
    np.testing.assert_allclose(ann.numpy(), g["annotation"], rtol=0, atol=2e-6 * np.abs(g["annotation"]).max())
    cost = O.get_cost(Pt, ann, torch.as_tensor(y), am, t64(ym))
    assert abs(float(cost) - float(g["cost"])) < 1e-8 * abs(float(g["cost"]))
    assert O.greedy_decode(Pt, ann, 30) == g["greedy"].tolist()
    assert list(O.beam_decode(Pt, ann, 10, 30)[0]) == g["beam"].tolist()


def test_golden_small_train():
    g = np.load(os.path.join(GOLDEN, "small_train.npz"))
    imgs, caps = O.synth_batch(int(g["data_seed"]), 2, 32, 48, 3, ragged=True)
    x, xm, y, ym = O.prepare_data(imgs, caps)
    P = O.to_torch(O.lively_params(int(g["param_seed"])), torch.float64, requires_grad=True)
    dfn = O.hash_drop_fn(int(g["drop_seed"]))
    ann, am, nm = O.dense_net(P, t64(x), t64(xm), True, dfn)
    cost = O.get_cost(P, ann, torch.as_tensor(y), am, t64(ym), dfn)
    assert abs(float(cost) - float(g["cost"])) < 1e-9 * abs(float(g["cost"]))
    names = [k[6:] for k in g.files if k.startswith("grad::")]
    grads = torch.autograd.grad(cost, [P[n] for n in names])
    for n, gr in zip(names, grads):
        ref = g["grad::" + n]
        np.testing.assert_allclose(gr.numpy(), ref, rtol=0, atol=2e-6 * max(np.abs(ref).max(), 1e-12))
    np.testing.assert_allclose(nm["batch_normalization_98/moving_variance"].numpy(), g["moving_variance_98"], rtol=1e-6)


# ------------------------------------------------------------------ misc semantics
def test_dropout_hash_statistics_and_determinism():
    a = O.drop_hash_keep(1234, 5, 200000); b = O.drop_hash_keep(1234, 5, 200000)
    assert np.array_equal(a, b) and abs(a.mean() - 0.8) < 5e-3
    assert not np.array_equal(a, O.drop_hash_keep(1234, 6, 200000))
    assert not np.array_equal(a, O.drop_hash_keep(1235, 5, 200000))


def test_get_word_first_step_uses_zero_embedding_and_gate_order():
    P = O.to_torch(O.lively_params(3))
    rs = np.random.RandomState(1)
    ann = t64(np.maximum(rs.randn(1, 2, 3, O.ANNOT_C), 0))
    h = O.hidden_state_0(P, ann); ap = torch.zeros(1, 2, 3, dtype=torch.float64)
    p, h1, ap1, _ = O.get_word(P, torch.tensor([-1]), h, ap, ann)
    assert abs(float(p.sum()) - 1) < 1e-12 and abs(float(ap1.sum()) - 1) < 1e-12
    # r is the FIRST half of the gate vector (model-single-GPU.py:276-277)
    g = torch.sigmoid(P["b_yz_yr"] + h @ P["U_hz_hr"])
    r1, z1 = g[:, :256], g[:, 256:]
    prop = torch.tanh((h @ P["U_rh"]) * r1 + P["b_yh"])
    pre_h = z1 * h + (1 - z1) * prop
    ctx, _, _ = O.get_context(P, ann, pre_h, ap, None)
    assert ctx.shape == (1, O.ANNOT_C)


def test_wer_exprate_scorer():
    rec = {"a": ["x", "y", "z"], "b": ["1", "2"], "c": []}
    lab = {"a": ["x", "z"], "b": ["1", "2"], "c": ["q"]}
    wer, exp = O.wer_exprate(rec, lab)
    assert abs(wer - (1 + 0 + 1) / 5) < 1e-12 and abs(exp - 1 / 3) < 1e-12
    assert O.edit_distance([1, 2, 3], [1, 3]) == 1 and O.edit_distance([], [1, 2]) == 2


def test_adadelta_update_matches_an_independent_implementation():
    """The oracle's Adadelta (tf.train.AdadeltaOptimizer, S:602: rho = 0.95, epsilon = 1e-8, learning_rate fed = 1.0) against
    torch.optim.Adadelta(rho=0.95, eps=1e-8, lr=1.0) - an independent implementation of the same published update rule - over
    three steps of a full oracle train_step (the gradient includes the L2 term, as in the reference's `cost`)."""
    P = O.lively_params(3)
    imgs, caps = O.synth_batch(9, 2, 24, 40, 3, ragged=True)
    x, xm, y, ym = O.prepare_data(imgs, caps)
    names = O.trainable_names(P)
    acc = {n: np.zeros(P[n].shape, np.float64) for n in names}
    accd = {n: np.zeros(P[n].shape, np.float64) for n in names}
    tp = {n: torch.nn.Parameter(torch.as_tensor(P[n], dtype=torch.float64)) for n in names}
    opt = torch.optim.Adadelta([tp[n] for n in names], lr=1.0, rho=0.95, eps=1e-8)
    Pn = dict(P)
    for step in range(3):
        _, G, Pn, acc, accd = O.train_step(Pn, acc, accd, x, xm, y, ym, 1.0, O.hash_drop_fn(50 + step))
        for n in names:
            tp[n].grad = torch.as_tensor(G[n], dtype=torch.float64)
        opt.step()
        # same gradient stream -> the parameters must follow the same trajectory (the oracle rounds parameters to fp32 each step)
        for n in ("conv2d_5/kernel", "U_a", "embed", "batch_normalization_7/gamma", "bo"):
            got, ref = Pn[n].astype(np.float64), tp[n].detach().numpy()
            assert np.abs(got - ref).max() <= 2e-7 * max(np.abs(ref).max(), 1e-3) * (step + 1), (n, step)


# ------------------------------------------------------------------ the TF-1.6 op shim against the same brute-force loops
@pytest.mark.parametrize("H,W,k,s,pad", [(6, 7, 7, 2, "SAME"), (7, 6, 7, 2, "SAME"), (5, 4, 3, 1, "SAME"), (4, 5, 11, 1, "SAME"), (3, 3, 1, 1, "VALID")])
def test_shim_ops_match_loops(H, W, k, s, pad):
    """oracle/tf16_shim.py (the stand-in the reference's own classes run on) implements the documented TF semantics: conv SAME /
    VALID, 2x2 SAME max / average pooling with valid-count divisor, fused batch norm, dropout keep-scale."""
    from oracle import tf16_shim as tf
    rs = np.random.RandomState(H * 10 + W)
    x = rs.randn(2, H, W, 2); w = rs.randn(k, k, 2, 3)
    tf.STATE.reset(preset={"conv2d/kernel": w})
    got = tf.layers.conv2d(t64(x), filters=3, kernel_size=[k, k], strides=s, padding=pad, use_bias=False).numpy()
    np.testing.assert_allclose(got, _conv_same_loops(x, w, s), atol=1e-10)           # k = 1: SAME == VALID
    np.testing.assert_allclose(tf.nn.conv2d(t64(x), filter=t64(w), strides=[1, 1, 1, 1], padding="SAME").numpy(), _conv_same_loops(x, w, 1), atol=1e-10)
    Ho, Wo = -(-H // 2), -(-W // 2)
    mx = np.full((2, Ho, Wo, 2), -np.inf); sm = np.zeros((2, Ho, Wo, 2)); cnt = np.zeros((Ho, Wo))
    for i in range(H):
        for j in range(W):
            mx[:, i // 2, j // 2] = np.maximum(mx[:, i // 2, j // 2], x[:, i, j])
            sm[:, i // 2, j // 2] += x[:, i, j]; cnt[i // 2, j // 2] += 1
    np.testing.assert_allclose(tf.layers.max_pooling2d(t64(x), pool_size=[2, 2], strides=2, padding="SAME").numpy(), mx, atol=1e-12)
    np.testing.assert_allclose(tf.layers.average_pooling2d(t64(x), pool_size=[2, 2], strides=2, padding="SAME").numpy(), sm / cnt[None, :, :, None], atol=1e-12)
    # batch norm (training): biased batch variance for the normalisation, Bessel-corrected for the moving update
    g, b = rs.randn(2), rs.randn(2); mm, mv = rs.randn(2), rs.rand(2) + 0.5
    tf.STATE.reset(preset={"batch_normalization/gamma": g, "batch_normalization/beta": b, "batch_normalization/moving_mean": mm,
                           "batch_normalization/moving_variance": mv})
    y = tf.layers.batch_normalization(t64(x), training=True, momentum=0.9, scale=True, epsilon=1e-4).numpy()
    mean = x.reshape(-1, 2).mean(0); var = x.reshape(-1, 2).var(0); n = x.size // 2
    np.testing.assert_allclose(y, (x - mean) / np.sqrt(var + 1e-4) * g + b, atol=1e-10)
    np.testing.assert_allclose(tf.STATE.moving_updates["batch_normalization/moving_variance"].numpy(), 0.9 * mv + 0.1 * var * n / (n - 1), atol=1e-12)
    np.testing.assert_allclose(tf.STATE.moving_updates["batch_normalization/moving_mean"].numpy(), 0.9 * mm + 0.1 * mean, atol=1e-12)
    # dropout: keep-mask scaled by 1 / (1 - rate), identity when not training
    tf.STATE.reset(drop=lambda key, shape: torch.as_tensor((np.arange(int(np.prod(shape))) % 5 != 0).astype(np.float64).reshape(shape)))
    d = tf.layers.dropout(t64(x), rate=0.2, training=True).numpy()
    keep = (np.arange(x.size) % 5 != 0).reshape(x.shape)
    np.testing.assert_allclose(d, x * keep / 0.8, atol=1e-12)
    assert np.array_equal(tf.layers.dropout(t64(x), rate=0.2, training=False).numpy(), x)
