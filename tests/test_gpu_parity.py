"""Parity of the sm_100a path (through the C ABI) against the CPU oracle.  `-m gpu` only.

Tolerance: 1e-3 relative (max|engine-oracle| / max|oracle|) — BASELINE.json north_star; greedy
token sequences must be identical.  Oracle side runs in float64.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import (O, REL_TOL, engine_gates, engine_grads, engine_with, grad_errors, make_batch, oracle_grads, rel_err, rel_l2,
                     t64)

pytestmark = pytest.mark.gpu
DEFAULT_TC_MASK = 0x3FF      # csrc/encoder.cu g_tc_debug_mask


def _check(report, test, name, got, ref, tol=REL_TOL, fails=None):
    e = rel_err(got, ref)
    report.add(test, name, e, tol, extra=dict(rel_l2=rel_l2(got, ref), ref_max=float(np.max(np.abs(ref))) if np.size(ref) else 0.0))
    if not (e <= tol) and fails is not None:
        fails.append((name, e))
    return e


def _lib():
    import wap_b200
    return wap_b200._lib


@pytest.mark.parametrize("h,w,k,maxlen,eol_bias", [(4, 16, 10, 12, 0.0), (8, 32, 10, 30, 3.0), (8, 32, 10, 30, 4.5), (3, 5, 1, 10, 2.0), (6, 10, 3, 25, 4.0),
                                                 (5, 40, 16, 20, 5.0), (20, 50, 10, 15, 4.0)])
def test_beam_search_on_device(report, h, w, k, maxlen, eol_bias):
    """wap_beam_decode_persistent (row f1): every sample, its score and the length-normalised pick equal the oracle's beam_decode
    (which is pinned to the reference's own get_sample by tests/test_host_cpu.py and test_reference_graph_cpu.py) and the
    host-bookkeeping path.  eol_bias raises bo[0] so that hypotheses finish at different steps (dead_k > 0, rows shrink)."""
    P = dict(O.lively_params(6))
    P["bo"] = P["bo"].copy(); P["bo"][0] += eol_bias
    eng = engine_with(P)
    Pt = O.to_torch(P)
    ann = _rand_annotation(300 + h * w + k, 1, h, w)
    ann_d = torch.as_tensor(ann).cuda()
    bref, sref, scref = O.beam_decode(Pt, t64(ann), k, maxlen)
    assert eng.beam_supported(h, w, k, maxlen)
    bgot, sgot, scgot = eng.get_sample(ann_d, k, maxlen)
    bhost, shost, schost = eng._get_sample_host(ann_d, k, maxlen)
    ok = list(bgot) == list(bref) and [list(map(int, a)) for a in sgot] == [list(map(int, a)) for a in sref]
    okh = [list(map(int, a)) for a in sgot] == [list(map(int, a)) for a in shost]
    es = rel_err(np.asarray(scgot, np.float64), np.asarray(scref, np.float64)) if ok else 1.0
    ndead = sum(1 for a in sref if a[-1] == 0)
    report.add("beam_device", f"{h}x{w}_k{k}_T{maxlen}_b{eol_bias}", es, 1e-4, extra=dict(identical=ok, identical_to_host_path=okh, finished=ndead, samples=len(sref)))
    assert ok and okh and es <= 1e-4, (bgot, bref, sgot, sref, es)


# --------------------------------------------------------------------------------------
def test_gemm_primitives(report):
    lib = _lib(); L = lib.lib()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    fails = []
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    for (M, N, K) in [(8, 256, 256), (300, 96, 72), (130, 111, 128), (257, 512, 684), (1000, 24, 864), (64, 684, 512)]:
        A = torch.randn(M, K, generator=g).to(dev); B = torch.randn(K, N, generator=g).to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        Cm = torch.zeros(M, N, device=dev)
        lib.check(L.wap_gemm_nn(st, p(A), K, p(B), N, p(Cm), N, M, N, K, p(bias), 0), "gemm_nn")
        ref = (A.double() @ B.double() + bias.double()).cpu().numpy()
        _check(report, "gemm", f"nn_{M}x{N}x{K}", Cm.cpu().numpy(), ref, 1e-5, fails)
        Wt = B.t().contiguous()
        C2 = torch.ones(M, N, device=dev)
        lib.check(L.wap_gemm_nt(st, p(A), K, p(Wt), K, p(C2), N, M, N, K, None, 1), "gemm_nt")
        _check(report, "gemm", f"nt_acc_{M}x{N}x{K}", C2.cpu().numpy(), (A.double() @ B.double() + 1).cpu().numpy(), 1e-5, fails)
    for (M, N, K) in [(72, 96, 5000), (256, 512, 41), (684, 256, 328), (864, 24, 4097)]:
        A = torch.randn(K, M, generator=g).to(dev); B = torch.randn(K, N, generator=g).to(dev)
        Cm = torch.zeros(M, N, device=dev)
        lib.check(L.wap_gemm_tn_atomic(st, p(A), M, p(B), N, p(Cm), N, M, N, K), "gemm_tn")
        _check(report, "gemm", f"tn_{M}x{N}x{K}", Cm.cpu().numpy(), (A.double().t() @ B.double()).cpu().numpy(), 1e-5, fails)
    assert not fails, fails


def test_dropout_hash_matches_numpy(report):
    lib = _lib(); L = lib.lib()
    n = 100003
    for seed, site in [(1234, 5), (0xFFFFFFF0, 99), (7, 0)]:
        m = torch.empty(n, device="cuda")
        lib.check(L.wap_dropout_mask(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_uint32(seed), C.c_uint32(site), n,
                                     C.c_void_p(m.data_ptr())), "dropout_mask")
        ref = O.drop_hash_keep(seed, site, n)
        assert np.array_equal(m.cpu().numpy(), ref)
        assert abs(ref.mean() - 0.8) < 0.01
    report.add("dropout", "hash_bit_exact", 0.0, 0.0)


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,ragged", [(1, 64, 256, False), (2, 37, 75, True), (3, 48, 80, True)])
def test_encoder_inference(report, B, H, W, ragged):
    x, xm, _, _ = make_batch(11 + B, B, H, W, 4, ragged)
    P = O.calibrate_moving(O.lively_params(3), x, xm)
    eng = engine_with(P)
    ann, am = eng.dense_net(x, xm, training=False)
    taps = {}
    Pt = O.to_torch(P)
    ref, ref_m, _ = O.dense_net(Pt, t64(x), t64(xm), False, taps=taps)
    fails = []
    tag = f"enc_infer_B{B}_{H}x{W}"
    _check(report, tag, "annotation", ann.cpu().numpy(), ref.numpy(), REL_TOL, fails)
    _check(report, tag, "mask", am.cpu().numpy(), ref_m.numpy(), 0.0, fails)
    # intermediate block buffers (diagnostics; same tolerance)
    lib = _lib(); L = lib.lib()
    ws = eng._ws[("enc",)]
    hw = [(taps["block0"].shape[1], taps["block0"].shape[2]), (taps["block1"].shape[1], taps["block1"].shape[2])]
    for b in range(2):
        off = L.wap_encoder_workspace_offset(B, H, W, 0, 1, b)
        h, w = hw[b]; Ct = taps[f"block{b}"].shape[3]
        got = ws[off: off + B * h * w * Ct].view(B, h, w, Ct).cpu().numpy()
        _check(report, tag, f"block{b}", got, taps[f"block{b}"].numpy(), REL_TOL, fails)
    assert not fails, fails


@pytest.mark.parametrize("B,H,W,drop", [(2, 40, 72, False), (2, 40, 72, True), (3, 33, 65, True)])
def test_encoder_training_forward(report, B, H, W, drop):
    x, xm, _, _ = make_batch(21 + B, B, H, W, 4, True)
    P = O.lively_params(4)
    eng = engine_with(P)
    seed = 4242
    ann, _ = eng.dense_net(x, xm, training=True, dropout=drop, seed=seed)
    Pt = O.to_torch(P)
    ref, _, new_moving = O.dense_net(Pt, t64(x), t64(xm), True, O.hash_drop_fn(seed) if drop else None)
    fails = []
    tag = f"enc_train_B{B}_{H}x{W}_drop{int(drop)}"
    _check(report, tag, "annotation", ann.cpu().numpy(), ref.numpy(), REL_TOL, fails)
    sd = eng.P.state_dict()
    worst_m = max(rel_err(sd[n], v.numpy()) for n, v in new_moving.items() if n.endswith("moving_mean"))
    worst_v = max(rel_err(sd[n], v.numpy()) for n, v in new_moving.items() if n.endswith("moving_variance"))
    report.add(tag, "moving_mean_worst", worst_m, REL_TOL); report.add(tag, "moving_var_worst", worst_v, REL_TOL)
    assert worst_m <= REL_TOL and worst_v <= REL_TOL, (worst_m, worst_v)
    assert not fails, fails


# --------------------------------------------------------------------------------------
def _rand_annotation(seed, B, h, w):
    rs = np.random.RandomState(seed)
    return np.maximum(rs.randn(B, h, w, O.ANNOT_C), 0).astype(np.float32)      # post-ReLU-like


@pytest.mark.parametrize("k,h,w,shared", [(1, 4, 16, False), (3, 5, 9, False), (10, 8, 32, True), (17, 3, 7, True)])
def test_get_word_step(report, k, h, w, shared):
    P = O.lively_params(5)
    eng = engine_with(P)
    Bn_ann = 1 if shared else k
    ann = _rand_annotation(7, Bn_ann, h, w)
    rs = np.random.RandomState(8)
    hp = np.tanh(rs.randn(k, 256)).astype(np.float32)
    ap = (rs.rand(k, h, w) * 0.3).astype(np.float32)
    y = rs.randint(0, 111, size=(k,)).astype(np.int64)
    ann_d = torch.as_tensor(ann).cuda()
    fails = []
    tag = f"get_word_k{k}_{h}x{w}_shared{int(shared)}"
    for first in (False, True):
        yy = np.full((k,), -1, np.int64) if first else y
        probs, word, hn, apn = eng.get_word(yy, hp, ap, ann_d)
        Pt = O.to_torch(P)
        ann_ref = t64(ann).expand(k, -1, -1, -1) if shared else t64(ann)
        rp, rh, rap, _ = O.get_word(Pt, torch.as_tensor(yy), t64(hp), t64(ap), ann_ref)
        s = "_first" if first else ""
        _check(report, tag, "probs" + s, probs.cpu().numpy(), rp.numpy(), REL_TOL, fails)
        _check(report, tag, "h" + s, hn.cpu().numpy(), rh.numpy(), REL_TOL, fails)
        _check(report, tag, "alpha_past" + s, apn.cpu().numpy(), rap.numpy(), REL_TOL, fails)
        assert np.array_equal(word.cpu().numpy(), rp.numpy().argmax(1)), "argmax word differs"
    assert not fails, fails


def test_hidden_state_0_and_precompute(report):
    P = O.lively_params(5)
    eng = engine_with(P)
    ann = _rand_annotation(9, 3, 5, 11)
    am = np.ones((3, 5, 11), np.float32); am[1, 3:, :] = 0; am[2, :, 7:] = 0
    ann_d = torch.as_tensor(ann).cuda(); am_d = torch.as_tensor(am).cuda()
    Pt = O.to_torch(P)
    fails = []
    _check(report, "precompute", "h0_unmasked", eng.hidden_state_0(ann_d).cpu().numpy(), O.hidden_state_0(Pt, t64(ann)).numpy(), REL_TOL, fails)
    watch, h0, qf = eng.precompute(ann_d, am_d)
    mean = (t64(ann) * t64(am)[..., None]).sum((1, 2)) / t64(am).sum((1, 2))[:, None]
    _check(report, "precompute", "h0_masked", h0.cpu().numpy(), torch.tanh(mean @ Pt["Wa2h"] + Pt["ba2h"]).numpy(), REL_TOL, fails)
    _check(report, "precompute", "watch", watch.cpu().numpy(), (torch.tensordot(t64(ann), Pt["U_a"], dims=1) + Pt["U_a_b"]).numpy(), REL_TOL, fails)
    _check(report, "precompute", "qfused", qf.cpu().numpy(), (Pt["alpha_past_filter"].reshape(121, 512) @ Pt["U_f"]).numpy(), REL_TOL, fails)
    assert not fails, fails


def test_greedy_and_beam_sequences_identical(report):
    P = O.lively_params(6)
    eng = engine_with(P)
    fails = []
    for i, (h, w) in enumerate([(4, 16), (6, 10), (3, 5)]):
        ann = _rand_annotation(30 + i, 1, h, w)
        ann_d = torch.as_tensor(ann).cuda()
        Pt = O.to_torch(P)
        ref = O.greedy_decode(Pt, t64(ann), 25)
        toks, lens, _ = eng.greedy_decode(ann_d, 25, check_every=4)
        got = toks[0, : int(lens[0])].cpu().tolist()
        ok = got == ref
        report.add("greedy", f"img{i}", 0.0 if ok else 1.0, 0.0, extra=dict(got=got, ref=ref))
        if not ok:
            fails.append(("greedy", got, ref))
        bref = O.beam_decode(Pt, t64(ann), 10, 12)[0]
        bgot = eng.get_sample(ann_d, 10, 12)[0]
        okb = list(bgot) == list(bref)
        report.add("beam", f"img{i}", 0.0 if okb else 1.0, 0.0, extra=dict(got=list(map(int, bgot)), ref=list(map(int, bref))))
        if not okb:
            fails.append(("beam", bgot, bref))
    # batched greedy == per-image greedy
    anns = np.concatenate([_rand_annotation(40 + i, 1, 4, 8) for i in range(5)], 0)
    toks, lens, _ = eng.greedy_decode(torch.as_tensor(anns).cuda(), 15, check_every=0)
    for i in range(5):
        ref = O.greedy_decode(O.to_torch(P), t64(anns[i: i + 1]), 15)
        got = toks[i, : int(lens[i])].cpu().tolist()
        report.add("greedy_batched", f"img{i}", 0.0 if got == ref else 1.0, 0.0, extra=dict(got=got, ref=ref))
        if got != ref:
            fails.append(("greedy_batched", i, got, ref))
    assert not fails, fails


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,h,w,T,drop", [(2, 3, 5, 4, False), (3, 4, 7, 6, True)])
def test_get_cost_forward(report, B, h, w, T, drop):
    P = O.lively_params(7)
    eng = engine_with(P)
    rs = np.random.RandomState(12)
    ann = _rand_annotation(13, B, h, w)
    am = np.ones((B, h, w), np.float32); am[-1, :, w - 2:] = 0
    lens = rs.randint(1, T, size=B); lens[0] = T - 1
    y = np.zeros((T, B), np.int64); ym = np.zeros((T, B), np.float32)
    for b in range(B):
        y[: lens[b], b] = rs.randint(1, 111, size=lens[b]); ym[: lens[b] + 1, b] = 1
    seed = 99
    cost = eng.get_cost(torch.as_tensor(ann).cuda(), y, torch.as_tensor(am).cuda(), ym, dropout=drop, seed=seed)
    Pt = O.to_torch(P)
    ref = O.get_cost(Pt, t64(ann), torch.as_tensor(y), t64(am), t64(ym), O.hash_drop_fn(seed) if drop else None)
    e = abs(float(cost) - float(ref)) / abs(float(ref))
    report.add("get_cost", f"B{B}_{h}x{w}_T{T}_drop{int(drop)}", e, REL_TOL, extra=dict(got=float(cost), ref=float(ref)))
    assert e <= REL_TOL, (float(cost), float(ref))


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
@pytest.mark.parametrize("B,H,W,cap,drop,bseed", [(2, 32, 48, 3, False, 52), (2, 40, 72, 5, True, 52), (2, 40, 72, 5, True, 152)])
def test_full_gradients(report, B, H, W, cap, drop, bseed, precision):
    """Every one of the 329 parameter gradients of cost (no L2) vs fp64 autograd of the oracle, in the default 3xTF32
    tensor-core mode and in the fp32 FFMA mode.  EVERY batch must pass and no entry is masked: the oracle evaluates its ReLUs
    with the engine's own gates (helpers.engine_gates), so an element within fp32 rounding of the kink is gated alike on both sides."""
    P = O.lively_params(8)
    seed = 777
    eng = engine_with(P, precision=precision)
    x, xm, y, ym = make_batch(bseed, B, H, W, cap, True)
    cost = float(eng.forward_backward(x, y, xm, ym, dropout=drop, seed=seed))
    ref_cost, G, names = oracle_grads(P, x, xm, y, ym, seed if drop else None, gates=engine_gates(eng, B, x.shape[1], x.shape[2]))
    tag = f"grads_{precision}_B{B}_{H}x{W}_T{cap + 1}_drop{int(drop)}_b{bseed}"
    e = abs(cost - ref_cost) / abs(ref_cost)
    report.add(tag, "cost", e, REL_TOL, extra=dict(got=cost, ref=ref_cost))
    fails, worst, worst_name, med, masked = grad_errors(names, engine_grads(eng), G)
    report.add(tag, "grad_worst_maxnorm", worst, 2e-3, extra=dict(name=worst_name, masked_entries=masked))
    report.add(tag, "grad_median_rel_l2", med, 1e-3)
    assert e <= REL_TOL and not fails and med <= 1e-3, (e, fails[:10], med)


def test_adadelta_l2_step(report):
    lib = _lib(); L = lib.lib()
    n, nd = 4096, 3000 - 3000 % 4
    rs = np.random.RandomState(3)
    p = rs.randn(n).astype(np.float32); g = rs.randn(n).astype(np.float32) * 0.1
    a = (rs.rand(n) * 0.01).astype(np.float32); u = (rs.rand(n) * 0.001).astype(np.float32)
    pt, gt, at, ut = [torch.as_tensor(v).cuda() for v in (p, g, a, u)]
    l2c = torch.zeros(1, dtype=torch.float64, device="cuda")
    f = C.c_float
    lib.check(L.wap_adadelta_l2_step(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(pt.data_ptr()), C.c_void_p(gt.data_ptr()),
                                     C.c_void_p(at.data_ptr()), C.c_void_p(ut.data_ptr()), n, nd, f(0.5), f(0.95), f(1e-8), f(1e-4), f(0.25),
                                     C.c_void_p(l2c.data_ptr())), "adadelta")
    gg = g.astype(np.float64) * 0.25; gg[:nd] += 2e-4 * p[:nd].astype(np.float64)
    a2 = 0.95 * a + 0.05 * gg * gg
    upd = np.sqrt(u + 1e-8) / np.sqrt(a2 + 1e-8) * gg
    u2 = 0.95 * u + 0.05 * upd * upd
    p2 = p - 0.5 * upd
    fails = []
    _check(report, "adadelta", "params", pt.cpu().numpy(), p2, 1e-5, fails)
    _check(report, "adadelta", "accum", at.cpu().numpy(), a2, 1e-5, fails)
    _check(report, "adadelta", "accum_update", ut.cpu().numpy(), u2, 1e-5, fails)
    _check(report, "adadelta", "l2_cost", l2c.cpu().numpy(), np.array([1e-4 * np.sum(p[:nd].astype(np.float64) ** 2)]), 1e-5, fails)
    assert not fails, fails


def test_train_steps_match_oracle(report):
    """Two full optimiser steps from oracle-synchronised state: cost incl. L2, moving statistics and the
    gradient against the fp64 oracle; Adadelta state and parameters against the TF update formula applied
    (in fp64) to the engine's OWN gradient.  (Comparing parameter deltas against the oracle's directly is
    ill-posed: with eps = 1e-8 the first Adadelta updates are ~ g * 1e-4 / sqrt(0.05 g^2 + 1e-8), i.e.
    fp32-vs-fp64 gradient noise of 1e-5 is amplified to O(0.1) of the update for small |g|.)"""
    B, H, W, cap = 2, 48, 80, 3
    P = O.lively_params(9)
    eng = engine_with(P)
    names = O.trainable_names(P)
    acc = {n: np.zeros(P[n].shape, np.float64) for n in names}
    accd = {n: np.zeros(P[n].shape, np.float64) for n in names}
    Pn = dict(P)
    l2set = set(O.l2_names(P))
    fails = []
    for step in range(2):
        seed = 500 + step
        x, xm, y, ym = make_batch(60 + step, B, H, W, cap, True)
        eng.P.load(Pn); eng.P.pack_into(eng.accum, acc); eng.P.pack_into(eng.accum_update, accd)
        flat0 = eng.P.flat.clone()
        cost = float(eng.train_step(x, y, xm, ym, lr=1.0, dropout=True, seed=seed))
        gates = engine_gates(eng, B, x.shape[1], x.shape[2], params=flat0)       # gamma / beta of the forward, not the updated ones
        ref_cost, G, Pn2, acc2, accd2 = O.train_step(Pn, acc, accd, x, xm, y, ym, 1.0, O.hash_drop_fn(seed), gates=gates)
        e = abs(cost - ref_cost) / abs(ref_cost)
        report.add("train_step", f"cost_step{step}", e, REL_TOL, extra=dict(got=cost, ref=ref_cost))
        if e > REL_TOL:
            fails.append(("cost", step, cost, ref_cost))
        sd = eng.P.state_dict()
        g_eng = eng.P.unpack(eng.grads); acc_e = eng.P.unpack(eng.accum); accd_e = eng.P.unpack(eng.accum_update)
        g_l2 = {n: g_eng[n].astype(np.float64) + (2 * O.L2_COEF * Pn[n].astype(np.float64) if n in l2set else 0.0) for n in names}
        gf, worst_g, worst_name, _, masked = grad_errors(names, g_l2, G)
        worst_p = worst_a = 0.0
        for n in names:
            g = g_l2[n]
            a = O.ADADELTA_RHO * acc[n] + (1 - O.ADADELTA_RHO) * g * g
            d = np.sqrt(accd[n] + O.ADADELTA_EPS) / np.sqrt(a + O.ADADELTA_EPS) * g
            u = O.ADADELTA_RHO * accd[n] + (1 - O.ADADELTA_RHO) * d * d
            pnew = Pn[n].astype(np.float64) - d
            worst_a = max(worst_a, rel_err(acc_e[n], a), rel_err(accd_e[n], u))
            worst_p = max(worst_p, float(np.max(np.abs(sd[n] - pnew))) / max(float(np.max(np.abs(d))), 1e-12))
        report.add("train_step", f"gradient_incl_l2_vs_oracle_step{step}", worst_g, 2e-3, extra=dict(name=worst_name, masked_entries=masked))
        report.add("train_step", f"adadelta_state_vs_formula_step{step}", worst_a, 1e-4)
        report.add("train_step", f"param_update_vs_formula_step{step}", worst_p, 1e-3)
        if gf or worst_a > 1e-4 or worst_p > 1e-3:
            fails.append(("update", step, gf[:5], worst_a, worst_p))
        mv = max(rel_err(sd[n], Pn2[n]) for n in P if n.endswith("moving_variance") or n.endswith("moving_mean"))
        report.add("train_step", f"moving_stats_step{step}", mv, 1e-3)
        if mv > 1e-3:
            fails.append(("moving", step, mv))
        Pn, acc, accd = Pn2, acc2, accd2
    assert not fails, fails


def test_fused_bn_backward_reduce_matches_separate_kernels(report):
    """The first pass of the 1x1 output's BN backward inside the 3x3 dgrad epilogue (debug-mask bit 10; off by default - it measured slower,
    DESIGN 3.2) == dgrad3 -> bn_bwd_reduce as separate kernels.  ONE forward, then the backward twice on the same saved activations (a second
    forward could gate an on-kink element differently): same gradients up to the order of the fp32 sums, 48 launches fewer."""
    import ctypes as C
    import wap_b200
    from wap_b200.engine import _p, _stream
    P = O.lively_params(5)
    x, xm, y, ym = make_batch(33, 2, 50, 150, 4, True)
    eng = engine_with(P, use_graphs=False)
    L = wap_b200._lib.lib()
    c0 = float(eng.forward_backward(x, y, xm, ym, True, 9)); g0 = eng.grads.clone()
    xd, yd, xmd, ymd = eng._dev(x, torch.float32), eng._dev(y, torch.int32), eng._dev(xm, torch.float32), eng._dev(ym, torch.float32)
    B, H, W, T = x.shape[0], x.shape[1], x.shape[2], y.shape[0]

    def backward_again():
        eng.grads.zero_()
        n0 = eng.launch_count()
        wap_b200._lib.check(eng.L.wap_train_backward(_stream(), _p(eng.P.flat), _p(eng.grads), _p(xd), B, H, W, _p(eng._ws[("ann",)]), _p(eng._ws[("amask",)]),
                                                     _p(yd), _p(ymd), T, 1, C.c_uint32(9), _p(eng._ws[("dann",)]), _p(eng._ws[("enc",)]), _p(eng._ws[("dec",)])),
                            "wap_train_backward")
        return eng.grads.clone(), eng.launch_count() - n0
    g_plain, n_plain = backward_again()
    try:
        L.wap_debug_set_tc_mask(0x3FF | 1024)
        g_fused, n_fused = backward_again()
    finally:
        L.wap_debug_set_tc_mask(0x3FF)
    e0 = rel_l2(g_plain.cpu().numpy(), g0.cpu().numpy()); e = rel_l2(g_fused.cpu().numpy(), g_plain.cpu().numpy())
    report.add("bn_bwd_fused", "grads_rel_l2_vs_unfused", e, 2e-5, extra=dict(rerun_noise=e0, launches=(n_plain, n_fused)))
    assert e0 <= 2e-5 and e <= 2e-5 and n_fused == n_plain - 48, (c0, e0, e, n_plain, n_fused)


@pytest.mark.parametrize("B,H,W,cap", [(2, 50, 150, 4), (8, 128, 512, 6), (3, 64, 100, 5), (1, 33, 47, 3)])
def test_single_kernel_bn_backward_matches_two_pass(report, B, H, W, cap):
    """BN backward of the small feature maps as ONE kernel (both passes around a grid barrier, elements kept in registers; debug-mask bit 11)
    == reduce kernel + apply kernel.  ONE forward, backward twice on the same saved activations: same gradients up to the order of the sums."""
    import ctypes as C
    import wap_b200
    from wap_b200.engine import _p, _stream
    P = O.lively_params(5)
    x, xm, y, ym = make_batch(35, B, H, W, cap, True)
    eng = engine_with(P, use_graphs=False)
    L = wap_b200._lib.lib()
    mask0 = 0x3FF
    L.wap_debug_set_tc_mask(mask0)
    try:
        c0 = float(eng.forward_backward(x, y, xm, ym, True, 9))
        xd, yd, xmd, ymd = eng._dev(x, torch.float32), eng._dev(y, torch.int32), eng._dev(xm, torch.float32), eng._dev(ym, torch.float32)
        T = y.shape[0]

        def backward_again():
            eng.grads.zero_()
            n0 = eng.launch_count()
            wap_b200._lib.check(eng.L.wap_train_backward(_stream(), _p(eng.P.flat), _p(eng.grads), _p(xd), B, H, W, _p(eng._ws[("ann",)]), _p(eng._ws[("amask",)]),
                                                         _p(yd), _p(ymd), T, 1, C.c_uint32(9), _p(eng._ws[("dann",)]), _p(eng._ws[("enc",)]), _p(eng._ws[("dec",)])),
                                "wap_train_backward")
            return eng.grads.clone(), eng.launch_count() - n0
        g_two, n_two = backward_again()
        L.wap_debug_set_tc_mask(mask0 | 2048)
        g_one, n_one = backward_again()
        g_one2, _ = backward_again()
    finally:
        L.wap_debug_set_tc_mask(DEFAULT_TC_MASK)
    e = rel_l2(g_one.cpu().numpy(), g_two.cpu().numpy()); e2 = rel_l2(g_one2.cpu().numpy(), g_one.cpu().numpy())
    report.add("bn_bwd_single_kernel", f"grads_rel_l2_vs_two_pass_B{B}_{H}x{W}", e, 2e-5, extra=dict(rerun=e2, launches=(n_two, n_one)))
    assert e <= 2e-5 and e2 <= 2e-5 and n_one < n_two, (c0, e, e2, n_two, n_one)


@pytest.mark.parametrize("B,H,W,cap,ragged", [(2, 50, 150, 4, True), (8, 128, 512, 12, False), (3, 64, 100, 7, True)])
def test_fused_bptt_matches_separate_kernels(report, B, H, W, cap, ragged):
    """BPTT with the GRU backward passes evaluated inside the consuming skinny GEMMs and dalpha / softmax backward / dz as one kernel (the
    default, wap_debug_set_fuse_gates bit 1) == the eight separate launches per time step.  ONE forward, backward twice on the same saved
    activations: same gradients up to the order of the fp32 atomics, 5 launches per time step fewer (11 -> 6; dependent chain 8 -> 4), two column sums more after the loop (dV_a / dV_a_b partial rows)."""
    import ctypes as C
    import wap_b200
    from wap_b200.engine import _p, _stream
    P = O.lively_params(5)
    x, xm, y, ym = make_batch(34, B, H, W, cap, ragged)
    eng = engine_with(P, use_graphs=False)
    L = wap_b200._lib.lib()
    c0 = float(eng.forward_backward(x, y, xm, ym, True, 9)); g0 = eng.grads.clone()
    xd, yd, xmd, ymd = eng._dev(x, torch.float32), eng._dev(y, torch.int32), eng._dev(xm, torch.float32), eng._dev(ym, torch.float32)
    T = y.shape[0]

    def backward_again():
        eng.grads.zero_()
        n0 = eng.launch_count()
        wap_b200._lib.check(eng.L.wap_train_backward(_stream(), _p(eng.P.flat), _p(eng.grads), _p(xd), B, H, W, _p(eng._ws[("ann",)]), _p(eng._ws[("amask",)]),
                                                     _p(yd), _p(ymd), T, 1, C.c_uint32(9), _p(eng._ws[("dann",)]), _p(eng._ws[("enc",)]), _p(eng._ws[("dec",)])),
                            "wap_train_backward")
        return eng.grads.clone(), eng.launch_count() - n0
    g_fused, n_fused = backward_again()
    try:
        L.wap_debug_set_fuse_gates(1)
        g_plain, n_plain = backward_again()
    finally:
        L.wap_debug_set_fuse_gates(3)
    e0 = rel_l2(g_fused.cpu().numpy(), g0.cpu().numpy()); e = rel_l2(g_fused.cpu().numpy(), g_plain.cpu().numpy())
    names = {n: (o, int(np.prod(sh))) for n, o, sh, f in eng.P.table if not f & 2}
    gmax = float(g_plain.abs().max())      # V_a_b shifts every energy of a softmax: its gradient is analytically 0 and numerically pure noise
    worst = max((rel_l2(g_fused[o:o + k].cpu().numpy(), g_plain[o:o + k].cpu().numpy()), n) for n, (o, k) in names.items()
                if float(g_plain[o:o + k].abs().max()) > 1e-6 * gmax)
    report.add("bptt_fused", f"grads_rel_l2_vs_unfused_B{B}", e, 2e-5, extra=dict(rerun_noise=e0, worst=worst, launches=(n_plain, n_fused)))
    assert e0 <= 2e-5 and e <= 2e-5 and worst[0] <= 2e-4 and n_fused == n_plain - 5 * T + 2, (c0, e0, e, worst, n_plain, n_fused)


def test_two_stream_decode_pipelines_match_per_image_calls():
    """greedy_decode_images / get_sample_images (encoder of image i+1 overlapping the decode of image i on a second stream) return exactly what the
    per-image calls return."""
    imgs = []
    for i, (H, W) in enumerate([(40, 120), (33, 75), (64, 200), (48, 90), (40, 120)]):
        x, _, _, _ = make_batch(700 + i, 1, H, W, 3)
        imgs.append(x)
    P = O.calibrate_moving(O.lively_params(6), imgs[0], np.ones(imgs[0].shape[:3], np.float32))
    P = dict(P); P["bo"] = P["bo"].copy(); P["bo"][0] += 2.0
    eng = engine_with(P)
    dev = [torch.as_tensor(x).cuda() for x in imgs]
    with eng.decoding():
        pipe = eng.greedy_decode_images(dev, [10, 12, 14, 16, 18])
        beams = eng.get_sample_images(dev, 5, 15)
        for i, x in enumerate(dev):
            ann, _ = eng.dense_net(x, None, training=False)
            t, l, _ = eng.greedy_decode(ann, [10, 12, 14, 16, 18][i])
            assert torch.equal(pipe[i][0], t) and torch.equal(pipe[i][1], l), i
            assert list(eng.get_sample(ann, 5, 15)[0]) == beams[i], i


def test_decoding_context_caches_the_fused_coverage_filter():
    """WAPEngine.decoding(): Q' = Q*U_f computed once and handed to the decode kernels (wap_set_fused_coverage_filter) - same tokens and
    beam samples as the per-call recomputation, for the cluster kernel (B <= 7), the grid-wide kernel and the beam search."""
    P = dict(O.lively_params(6)); P["bo"] = P["bo"].copy(); P["bo"][0] += 2.5
    eng = engine_with(P)
    for B in (3, 9):
        ann = torch.as_tensor(_rand_annotation(900 + B, B, 6, 20)).cuda()
        t0, l0, _ = eng.greedy_decode(ann, 20)
        with eng.decoding():
            t1, l1, _ = eng.greedy_decode(ann, 20)
            b1 = eng.get_sample(ann[:1], 5, 15)
        b0 = eng.get_sample(ann[:1], 5, 15)
        assert torch.equal(t0, t1) and torch.equal(l0, l1)
        assert b0[1] == b1[1] and np.allclose(b0[2], b1[2], rtol=1e-6)
    assert not eng.L.wap_current() is None


@pytest.mark.parametrize("precision", ["tf32x3", "bf16"])
def test_inference_bn_fold_is_bit_identical(precision):
    """Inference encoder with batch norm + ReLU folded into the conv epilogues (default) == the conv -> bn_act kernel sequence, bit for bit
    (same scale / shift expressions), at roughly half the launches."""
    import wap_b200
    x, xm, _, _ = make_batch(21, 2, 50, 150, 3, True)
    P = O.calibrate_moving(O.lively_params(4), x, xm)
    eng = engine_with(P, precision=precision)
    L = wap_b200._lib.lib()
    n0 = eng.launch_count(); a_fused, _ = eng.dense_net(x, xm, training=False); n1 = eng.launch_count()
    try:
        L.wap_debug_set_tc_mask(0x1FF)
        a_plain, _ = eng.dense_net(x, xm, training=False); n2 = eng.launch_count()
    finally:
        L.wap_debug_set_tc_mask(0x3FF)
    if precision == "bf16":      # the unfused bf16 path rounds the raw conv output to bf16 BEFORE the batch norm, the fused one applies it to the fp32 accumulator
        assert rel_l2(a_fused.cpu().numpy(), a_plain.cpu().numpy()) < 2e-1      # two bf16 roundings of every raw conv output apart (inference mode amplifies: 9-12 % vs the oracle)
    else:
        assert torch.equal(a_fused, a_plain)
    assert (n1 - n0) <= (n2 - n1) - 90, (n1 - n0, n2 - n1)


@pytest.mark.parametrize("B,h,w,maxlen", [(7, 8, 32, 30), (8, 8, 32, 30), (16, 5, 9, 20), (1, 4, 16, 25), (3, 20, 50, 12), (2, 1, 1, 6), (5, 7, 13, 40), (1, 16, 61, 10)])
def test_greedy_cluster_kernel(report, B, h, w, maxlen):
    """The cluster-per-image greedy kernel (csrc/decoder_cluster.cu: 16 CTAs per image, DSMEM exchanges, no grid barrier) against the
    oracle's greedy decode, token for token, and against the grid-wide persistent kernel; an <eol> bias makes the images stop at
    different steps (each cluster leaves its loop on its own)."""
    import wap_b200
    P = dict(O.lively_params(6))
    P["bo"] = P["bo"].copy(); P["bo"][0] += 2.5
    eng = engine_with(P)
    L = wap_b200._lib.lib()
    ann = _rand_annotation(500 + B * h + w, B, h, w)
    ann_d = torch.as_tensor(ann).cuda()
    Pt = O.to_torch(P)
    try:
        L.wap_debug_set_greedy_impl(0)
        toks, lens, steps = eng.greedy_decode(ann_d, maxlen, persistent=True)
        L.wap_debug_set_greedy_impl(1)
        toks1, lens1, _ = eng.greedy_decode(ann_d, maxlen, persistent=True)
    finally:
        L.wap_debug_set_greedy_impl(0)
    bad = []
    for b in range(B):
        ref = O.greedy_decode(Pt, t64(ann[b:b + 1]), maxlen)
        got = toks[b, : int(lens[b])].cpu().tolist(); got1 = toks1[b, : int(lens1[b])].cpu().tolist()
        if got != ref or got1 != ref:
            bad.append((b, got, got1, ref))
    report.add("greedy_cluster", f"B{B}_{h}x{w}_T{maxlen}", float(len(bad)), 0.0, extra=dict(mismatch=bad[:2], lens=lens.cpu().tolist()))
    assert not bad and int(steps[0]) == int(lens.max()), (bad[:2], int(steps[0]), lens.cpu().tolist())


@pytest.mark.parametrize("B,h,w", [(8, 8, 32), (16, 5, 9), (1, 4, 16), (3, 20, 50)])
def test_greedy_persistent_matches_stepwise(report, B, h, w):
    """The single-launch persistent decoder emits exactly the token sequences of the step-by-step kernels
    (which test_greedy_and_beam_sequences_identical ties to the oracle), incl. lengths and early stop."""
    P = O.lively_params(6)
    eng = engine_with(P)
    anns = torch.as_tensor(np.concatenate([_rand_annotation(70 + i, 1, h, w) for i in range(B)], 0)).cuda()
    t1, l1, _ = eng.greedy_decode(anns, 30, check_every=0, persistent=False)
    t2, l2, _ = eng.greedy_decode(anns, 30, persistent=True)
    torch.cuda.synchronize()
    ok = bool((l1 == l2).all()) and all(t1[b, : int(l1[b])].tolist() == t2[b, : int(l2[b])].tolist() for b in range(B))
    report.add("greedy_persistent", f"B{B}_{h}x{w}", 0.0 if ok else 1.0, 0.0, extra=dict(l1=l1.tolist(), l2=l2.tolist(), t1=t1[0].tolist(), t2=t2[0].tolist()))
    assert ok, (l1.tolist(), l2.tolist(), t1[0].tolist(), t2[0].tolist())
    if B <= 3:
        ref = O.greedy_decode(O.to_torch(P), t64(anns[:1].cpu().numpy()), 30)
        assert t2[0, : int(l2[0])].tolist() == ref


def test_get_context_and_one_time_step_with_masks(report):
    """Attender.get_context (S:166-206) and Parser.one_time_step / get_ht_ctx (S:246-308) incl. a_mask / y_mask."""
    import wap_b200
    P = O.lively_params(5)
    eng = engine_with(P)
    Pt = O.to_torch(P)
    B, h, w, T = 3, 4, 7, 4
    rs = np.random.RandomState(3)
    ann = _rand_annotation(21, B, h, w)
    am = np.ones((B, h, w), np.float32); am[1, 2:, :] = 0; am[2, :, 5:] = 0
    hp = np.tanh(rs.randn(B, 256)).astype(np.float32); ap = (rs.rand(B, h, w) * 0.2).astype(np.float32)
    ann_d = torch.as_tensor(ann).cuda()
    fails = []
    for mask in (None, am):
        ctx, al, apn = eng.get_context(ann_d, hp, ap, mask)
        rc, ra, rap = O.get_context(Pt, t64(ann), t64(hp), t64(ap), t64(mask) if mask is not None else None)
        tag = "masked" if mask is not None else "unmasked"
        _check(report, "get_context", "context_" + tag, ctx.cpu().numpy(), rc.numpy(), REL_TOL, fails)
        _check(report, "get_context", "alpha_" + tag, al.cpu().numpy(), ra.numpy(), REL_TOL, fails)
        _check(report, "get_context", "alpha_past_" + tag, apn.cpu().numpy(), rap.numpy(), REL_TOL, fails)
    emb = rs.randn(T, B, 256).astype(np.float32) * 0.3
    ym = np.ones((T, B), np.float32); ym[2:, 1] = 0; ym[3, 2] = 0
    att = wap_b200.Attender(684, 256, 512, engine=eng); par = wap_b200.Parser(256, 256, att, 684)
    hs, cs, als, aps, _, _ = par.get_ht_ctx(torch.as_tensor(emb).cuda(), torch.as_tensor(hp).cuda(), ann_d, torch.as_tensor(am).cuda(),
                                            torch.as_tensor(ym).cuda())
    hh, app = t64(hp), torch.zeros(B, h, w, dtype=torch.float64)
    for t in range(T):
        hh, cc, aa, app = O.one_time_step(Pt, hh, app, t64(ann), t64(am), t64(emb[t]), t64(ym[t]))
        _check(report, "one_time_step", f"h_t{t}", hs[t].cpu().numpy(), hh.numpy(), REL_TOL, fails)
        _check(report, "one_time_step", f"ctx_t{t}", cs[t].cpu().numpy(), cc.numpy(), REL_TOL, fails)
        _check(report, "one_time_step", f"alpha_past_t{t}", aps[t].cpu().numpy(), app.numpy(), REL_TOL, fails)
    assert not fails, fails


def test_checkpoint_roundtrip(tmp_path):
    P = O.lively_params(2)
    eng = engine_with(P)
    x, xm, y, ym = make_batch(5, 2, 32, 48, 3, True)
    eng.train_step(x, y, xm, ym, lr=1.0, dropout=True, seed=3)
    path = str(tmp_path / "ckpt.npz")
    eng.save(path)
    c1 = float(eng.train_step(x, y, xm, ym, lr=1.0, dropout=True, seed=4))
    eng2 = engine_with(O.lively_params(7))
    eng2.load(path)
    assert eng2.step_count == 1
    c2 = float(eng2.train_step(x, y, xm, ym, lr=1.0, dropout=True, seed=4))
    assert abs(c1 - c2) <= 2e-5 * abs(c1), (c1, c2)      # two fp32 runs: atomic-order noise only
    z = np.load(path)
    assert "conv2d_98/kernel" in z.files and "batch_normalization_50/moving_variance" in z.files and "__opt__/accum/embed" in z.files


def test_cuda_graph_replay_matches_stream_launches():
    """use_graphs (the default): a shape is captured the SECOND time it is seen; the captured forward+backward, replayed with a
    fresh device-side dropout seed, equals stream launches with the same by-value seed on the same weights (run-to-run noise of
    the atomics only).  Whole training trajectories are not compared: on these tiny batches two *eager* runs already diverge by
    30% in the gradient after one Adadelta step (sign-like first update), which says nothing about the replay."""
    P = O.lively_params(3)
    x, xm, y, ym = make_batch(11, 3, 32, 64, 4, True)
    eng = engine_with(P)
    assert eng.use_graphs and eng.conv_precision() == "tf32x3"
    eng.forward_backward(x, y, xm, ym, True, 1)            # first sighting: plain stream launches, nothing captured
    assert len(eng._graphs) == 0
    got = {}
    for seed in (5, 6, 7, 5):                 # second sighting captures (and runs eagerly), the others replay
        eng.use_graphs = True
        cg = float(eng.forward_backward(x, y, xm, ym, True, seed)); gg = eng.grads.clone()
        eng.use_graphs = False
        ce = float(eng.forward_backward(x, y, xm, ym, True, seed)); ge = eng.grads.clone()
        # two fp32 runs differ by the order of their atomic adds (~1e-6); a wrong seed or mask would differ by O(1)
        assert abs(cg - ce) <= 2e-5 * abs(ce), (seed, cg, ce)
        # ... except that once in a while that noise flips one ReLU gate sitting on its kink (eager vs eager does it too, tools/dbg_graph.py),
        # which moves the gradients of every layer below it by 1e-3..1e-2 (rel_l2 0.007 overall in the runs recorded in NOTES.md).  A wrong
        # seed or dropout mask moves them by O(1), so the bound sits between the two; the usual value is printed for the log.
        e = rel_l2(gg.cpu().numpy(), ge.cpu().numpy())
        print(f"graph vs eager, seed {seed}: cost {cg:.6f} / {ce:.6f}, grads rel_l2 {e:.2e}")
        assert e < 0.05, (seed, e)
        got[seed] = cg
    assert eng.replayed_launches > 0 and len(eng._graphs) == 1
    assert abs(got[5] - got[6]) > 1e-4 * abs(got[5]) and abs(got[6] - got[7]) > 1e-4 * abs(got[6])   # the seed really changed
    # a different shape gets its own graph; a full train_step through the replay path still changes the weights
    x2, xm2, y2, ym2 = make_batch(12, 2, 32, 48, 3, True)
    eng.use_graphs = True
    cs = [float(eng.forward_backward(x2, y2, xm2, ym2, True, 9)) for _ in range(3)]
    assert abs(cs[0] - cs[2]) <= 2e-5 * abs(cs[0]) and len(eng._graphs) == 2
    p0 = eng.P.flat.clone(); eng.train_step(x, y, xm, ym, lr=1.0, dropout=True, seed=5)
    assert float((eng.P.flat - p0).abs().max()) > 0
    # LRU eviction and invalidation when a workspace grows
    eng.max_graphs = 2
    x3, xm3, y3, ym3 = make_batch(13, 2, 24, 40, 2, True)
    for _ in range(2):
        eng.forward_backward(x3, y3, xm3, ym3, True, 3)
    assert len(eng._graphs) == 2 and (tuple(x2.shape), tuple(y2.shape), True, 2) not in eng._graphs
    xb, xmb, yb, ymb = make_batch(14, 4, 64, 128, 6, True)   # larger than anything seen: workspaces are reallocated
    eng.forward_backward(xb, yb, xmb, ymb, True, 3)
    assert len(eng._graphs) == 0


def test_engine_matches_reference_generated_golden(report):
    """The CUDA path against tests/golden/ref_graph_cfg1.npz: vectors produced by the REFERENCE's own classes (dense_net, get_cost,
    get_word, get_sample of /root/reference/model-single-GPU.py) running on oracle/tf16_shim.py - see oracle/make_reference_golden.py.
    The reference cannot travel to the GPU box; these vectors do.  Workload: SURVEY 8d cfg1 (one 64 x 256 image, 10-token caption)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_graph_cfg1.npz"))
    x, xm, y, ym = make_batch(int(g["data_seed"]), 1, 64, 256, int(g["cap_len"]), False)
    P = O.calibrate_moving(O.lively_params(int(g["param_seed"])), x, xm)
    eng = engine_with(P)
    fails = []
    # ---- inference graph
    ann, am = eng.dense_net(x, xm, training=False)
    _check(report, "ref_golden", "annotation", ann.cpu().numpy(), g["annotation"], REL_TOL, fails)
    _check(report, "ref_golden", "annotation_mask", am.cpu().numpy(), g["annotation_mask"], 0.0, fails)
    cost = float(eng.get_cost(ann, y, am, ym))
    e = abs(cost - float(g["cost"])) / abs(float(g["cost"]))
    report.add("ref_golden", "cost", e, REL_TOL, extra=dict(got=cost, ref=float(g["cost"])))
    if e > REL_TOL:
        fails.append(("cost", cost, float(g["cost"])))
    h0 = eng.hidden_state_0(ann)
    probs, word, hn, apn = eng.get_word(np.full((1,), -1, np.int64), h0.cpu().numpy(), np.zeros((1,) + tuple(ann.shape[1:3]), np.float32), ann)
    _check(report, "ref_golden", "probs0", probs.cpu().numpy(), g["probs0"], REL_TOL, fails)
    _check(report, "ref_golden", "h1", hn.cpu().numpy(), g["h1"], REL_TOL, fails)
    _check(report, "ref_golden", "alpha_past1", apn.cpu().numpy(), g["alpha_past1"], REL_TOL, fails)
    best = [int(t) for t in g["beam_best"] if t >= 0]
    got = [int(t) for t in eng.get_sample(ann, int(g["beam_k"]), int(g["beam_maxlen"]))[0]]
    report.add("ref_golden", "beam_identical", 0.0 if got == best else 1.0, 0.0, extra=dict(got=got, ref=best))
    if got != best:
        fails.append(("beam", got, best))
    # ---- training graph (batch statistics, hash-keyed dropout)
    seed = int(g["drop_seed"])
    ann_t, am_t = eng.dense_net(x, xm, training=True, dropout=True, seed=seed)
    _check(report, "ref_golden", "train_annotation", ann_t.cpu().numpy(), g["train_annotation"], REL_TOL, fails)
    cost_t = float(eng.get_cost(ann_t, y, am_t, ym, dropout=True, seed=seed))
    e = abs(cost_t - float(g["train_cost"])) / abs(float(g["train_cost"]))
    report.add("ref_golden", "train_cost", e, REL_TOL, extra=dict(got=cost_t, ref=float(g["train_cost"])))
    if e > REL_TOL:
        fails.append(("train_cost", cost_t, float(g["train_cost"])))
    sd = eng.P.state_dict()
    for key in g.files:
        if key.startswith("new::"):
            _check(report, "ref_golden", key[5:], sd[key[5:]], g[key], REL_TOL, fails)
    assert not fails, fails
