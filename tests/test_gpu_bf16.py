"""bf16 speed mode (wap_set_conv_precision(3), BASELINE configs[2]): the bf16 tcgen05 primitives against torch references computed
from the SAME bf16-rounded inputs in fp64 (so the comparison isolates indexing / descriptors / swizzles from the intrinsic bf16
rounding), and the whole encoder / training step in bf16 mode against the fp64 oracle with the deviation it actually achieves
(it does NOT meet the 1e-3 parity contract of the fp32-grade modes; the tolerances below are what bf16 storage through 99
conv+BN layers gives, stated per quantity)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import O, engine_gates, engine_with, make_batch, oracle_grads, rel_err, rel_l2, t64

pytestmark = pytest.mark.gpu

p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731


def _L():
    import wap_b200
    return wap_b200._lib, wap_b200._lib.lib()


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _bf(t):
    return t.to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("M,N,K,lda_pad,acc,out_bf16,stats", [
    (128, 96, 64, 0, 0, 1, 1), (300, 96, 72, 360, 0, 1, 1), (1000, 24, 96, 96, 1, 0, 0), (257, 128, 660, 28, 1, 0, 0), (5000, 216, 432, 0, 0, 1, 1),
    (131, 300, 600, 88, 0, 1, 1), (64, 32, 8, 0, 0, 0, 0), (2048, 384, 192, 0, 1, 0, 0)])
def test_bf16_gemm(report, M, N, K, lda_pad, acc, out_bf16, stats):
    lib, L = _L()
    g = torch.Generator().manual_seed(M * 7 + N)
    lda = K + lda_pad
    ldb = (K + 7) // 8 * 8                                 # operand row strides must be multiples of 16 bytes
    A = _bf(torch.randn(M, lda, generator=g)).cuda(); Bt = _bf(torch.randn(N, ldb, generator=g)).cuda()
    D0 = torch.randn(M, N, generator=g).cuda()
    D = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda") if out_bf16 else (D0.clone() if acc else torch.full((M, N), float("nan"), device="cuda"))
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda") if stats else None
    lib.check(L.wap_tc_gemm_bf16(_st(), p(A), lda, p(Bt), ldb, p(D), N, M, N, K, acc, out_bf16, p(st)), "wap_tc_gemm_bf16")
    torch.cuda.synchronize()
    ref = A[:, :K].double() @ Bt[:, :K].double().t() + (D0.double() if acc else 0)
    tol = 6e-3 if out_bf16 else 2e-5                       # bf16 result: one rounding of the output (2^-9); fp32 result: accumulation only
    e = rel_err(D.double().cpu().numpy(), ref.cpu().numpy())
    report.add("bf16", f"gemm_{M}x{N}x{K}_acc{acc}_obf{out_bf16}", e, tol)
    assert e <= tol, e
    if stats:
        es = rel_err(st[0].cpu().numpy(), ref.sum(0).cpu().numpy()); eq = rel_err(st[1].cpu().numpy(), (ref ** 2).sum(0).cpu().numpy())
        report.add("bf16", f"gemm_{M}x{N}x{K}_stats", max(es, eq), 2e-5)
        assert es <= 2e-5 and eq <= 2e-5, (es, eq)


def _conv_ref(x, w_hwio):
    """x [B,H,W,C] fp64, w [3,3,C,N] fp64 -> SAME 3x3 conv [B,H,W,N]."""
    return F.conv2d(x.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)


@pytest.mark.parametrize("B,H,W,C,N,out_bf16", [(2, 8, 16, 96, 24, 1), (1, 5, 9, 96, 24, 1), (1, 20, 150, 96, 24, 1), (2, 7, 20, 32, 96, 0), (1, 33, 70, 32, 96, 0)])
def test_bf16_conv3x3_9tap(report, B, H, W, C, N, out_bf16):
    """9-tap implicit GEMM: forward 96 -> 24 (rows wider than a tile) and the dgrad shape 32(24) -> 96 with fp32 output."""
    lib, L = _L()
    g = torch.Generator().manual_seed(B * 100 + H + W)
    ldx = C
    x = _bf(torch.randn(B, H, W, ldx, generator=g)).cuda()
    w = _bf(torch.randn(3, 3, C, N, generator=g) * 0.1).cuda()
    Cpad = -(-C // 64) * 64
    Nrows = 32 if N <= 32 else 96
    Wt = torch.zeros(Nrows, 9, Cpad, dtype=torch.bfloat16, device="cuda")
    Wt[:N, :, :C] = w.reshape(9, C, N).permute(2, 0, 1)
    D = torch.zeros(B * H * W, N, dtype=torch.bfloat16 if out_bf16 else torch.float32, device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    lib.check(L.wap_tc_conv3x3_bf16(_st(), p(x), ldx, B, H, W, C, p(Wt), Nrows, N, p(D), N, out_bf16, p(st)), "wap_tc_conv3x3_bf16")
    torch.cuda.synchronize()
    ref = _conv_ref(x.double(), w.double()).reshape(B * H * W, N)
    tol = 6e-3 if out_bf16 else 2e-5
    e = rel_err(D.double().cpu().numpy(), ref.cpu().numpy())
    es = rel_err(st[0].cpu().numpy(), ref.sum(0).cpu().numpy())
    report.add("bf16", f"conv3x3_B{B}_{H}x{W}_C{C}_N{N}", e, tol); report.add("bf16", f"conv3x3_B{B}_{H}x{W}_C{C}_N{N}_stats", es, 1e-4)
    assert e <= tol and es <= 1e-4, (e, es)


@pytest.mark.parametrize("B,H,W", [(2, 8, 16), (1, 5, 9), (3, 4, 32), (1, 32, 128), (2, 7, 100), (1, 3, 1)])
def test_bf16_conv3x3_rowtile(report, B, H, W):
    lib, L = _L()
    g = torch.Generator().manual_seed(B * 100 + H + W)
    x = _bf(torch.randn(B, H, W, 96, generator=g)).cuda()
    w = _bf(torch.randn(3, 3, 96, 24, generator=g) * 0.1).cuda()
    Wt3 = torch.zeros(96, 3, 128, dtype=torch.bfloat16, device="cuda")            # [24 j + co][dy][ci]
    for j in range(3):
        Wt3[24 * j: 24 * j + 24, :, :96] = w[:, j].permute(2, 0, 1)
    D = torch.zeros(B * H * W, 24, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(2, 24, dtype=torch.float64, device="cuda")
    lib.check(L.wap_tc_conv3x3_rowtile_bf16(_st(), p(x), 96, B, H, W, p(Wt3), p(D), 24, 1, p(st)), "wap_tc_conv3x3_rowtile_bf16")
    torch.cuda.synchronize()
    ref = _conv_ref(x.double(), w.double()).reshape(B * H * W, 24)
    e = rel_err(D.double().cpu().numpy(), ref.cpu().numpy()); es = rel_err(st[0].cpu().numpy(), ref.sum(0).cpu().numpy())
    report.add("bf16", f"rowtile_B{B}_{H}x{W}", e, 6e-3); report.add("bf16", f"rowtile_B{B}_{H}x{W}_stats", es, 1e-4)
    assert e <= 6e-3 and es <= 1e-4, (e, es)


@pytest.mark.parametrize("Mpix,C,N,ldx,ldy", [(64, 48, 96, 432, 192), (1000, 432, 96, 432, 192), (4100, 660, 96, 688, 192), (30, 516, 96, 688, 96), (32768, 120, 96, 432, 192),
                                              (900, 600, 128, 600, 304), (900, 600, 44, 600, 304)])
def test_bf16_wgrad_1x1(report, Mpix, C, N, ldx, ldy):
    """MN-major bf16 operands (plain SWIZZLE_128B atoms of 8 pixel rows x 64 channels)."""
    lib, L = _L()
    g = torch.Generator().manual_seed(Mpix + C)
    X = _bf(torch.randn(Mpix, ldx, generator=g)).cuda(); dY = _bf(torch.randn(Mpix, ldy, generator=g)).cuda()
    dW0 = torch.randn(C, N, generator=g).cuda(); dW = dW0.clone()
    lib.check(L.wap_tc_wgrad_1x1_bf16(_st(), p(X), ldx, p(dY), ldy, Mpix, C, N, p(dW), N), "wap_tc_wgrad_1x1_bf16")
    torch.cuda.synchronize()
    ref = dW0.double() + X[:, :C].double().t() @ dY[:, :N].double()
    e = rel_err(dW.cpu().numpy(), ref.cpu().numpy())
    report.add("bf16", f"wgrad1x1_M{Mpix}_C{C}_N{N}", e, 2e-5)
    assert e <= 2e-5, e


@pytest.mark.parametrize("B,H,W", [(2, 8, 16), (1, 5, 9), (2, 4, 32), (1, 32, 128), (3, 7, 20), (1, 9, 150)])
def test_bf16_wgrad_3x3(report, B, H, W):
    lib, L = _L()
    g = torch.Generator().manual_seed(B * 10 + H * W)
    C, N, ldy = 96, 24, 32
    X = _bf(torch.randn(B, H, W, C, generator=g)).cuda()
    dY = torch.zeros(B, H, W, ldy, dtype=torch.bfloat16); dY[..., :N] = _bf(torch.randn(B, H, W, N, generator=g)); dY = dY.cuda()
    dW0 = torch.randn(9 * C, N, generator=g).cuda(); dW = dW0.clone()
    lib.check(L.wap_tc_wgrad_3x3_bf16(_st(), p(X), C, p(dY), ldy, B, H, W, C, N, p(dW), N), "wap_tc_wgrad_3x3_bf16")
    torch.cuda.synchronize()
    xp = F.pad(X.double(), (0, 0, 1, 1, 1, 1))
    cols = torch.stack([xp[:, dy: dy + H, dx: dx + W, :] for dy in range(3) for dx in range(3)], 3)      # [B,H,W,9,C]
    ref = dW0.double() + cols.reshape(-1, 9 * C).t() @ dY[..., :N].double().reshape(-1, N)
    e = rel_err(dW.cpu().numpy(), ref.cpu().numpy())
    report.add("bf16", f"wgrad3x3_B{B}_{H}x{W}", e, 2e-5)
    assert e <= 2e-5, e


# ---- whole encoder / training step in bf16 mode: deviation from the fp64 oracle, stated honestly
# Measured in round 2 (lively / reference-initialised weights alike): annotation 4-7 % in training mode (batch statistics re-normalise
# the rounding noise at every layer), 9-20 % max-norm / 9-12 % L2 in inference mode.  Gradients: every BN->ReLU site gates ~0.3 % of its
# elements differently once its input carries 2^-9 relative rounding, which alone moves each layer's weight gradient by ~6 % and compounds
# to 50-90 % in the first layers of this 99-layer net (cosine 0.87 over all parameters) - a property of bf16 storage, not of the kernels.
# The kernels are therefore checked with the oracle evaluated under the ENGINE's ReLU gates (helpers.engine_gates), which leaves the pure
# rounding error (measured: median relative L2 12-18 %, cosine 0.985-0.993 - it varies from run to run with the order of the fp32 atomics -
# vs 48-50 % / 0.85-0.87 with the oracle's own gates).
BF16_TOL = dict(annotation=2.5e-1, annotation_l2=1.5e-1, cost=2e-2, grad_median_l2=3e-1, cosine=0.97)     # measured over runs: 0.17-0.20 / 0.10-0.115 / 3e-3 / 0.12-0.18 / 0.985-0.993


@pytest.mark.parametrize("B,H,W,training", [(1, 64, 256, False), (2, 37, 75, False), (2, 40, 72, True), (3, 70, 610, True)])
def test_bf16_encoder_forward(report, B, H, W, training):
    x, xm, _, _ = make_batch(11 + B, B, H, W, 4, True)
    P = O.lively_params(3)
    if not training:
        P = O.calibrate_moving(P, x, xm)
    eng = engine_with(P, precision="bf16")
    assert eng.conv_precision() == "bf16"
    seed = 4242
    ann, _ = eng.dense_net(x, xm, training=training, dropout=training, seed=seed)
    ref, _, nm = O.dense_net(O.to_torch(P), t64(x), t64(xm), training, O.hash_drop_fn(seed) if training else None)
    e, el2 = rel_err(ann.cpu().numpy(), ref.numpy()), rel_l2(ann.cpu().numpy(), ref.numpy())
    report.add("bf16", f"encoder_B{B}_{H}x{W}_train{int(training)}_maxnorm", e, BF16_TOL["annotation"], extra=dict(rel_l2=el2))
    report.add("bf16", f"encoder_B{B}_{H}x{W}_train{int(training)}_rel_l2", el2, BF16_TOL["annotation_l2"])
    assert e <= BF16_TOL["annotation"] and el2 <= BF16_TOL["annotation_l2"], (e, el2)
    if training:                                           # moving statistics come from the fp32 accumulators
        sd = eng.P.state_dict()
        mv = max(rel_err(sd[n], nm[n].numpy()) for n in nm)
        report.add("bf16", f"encoder_B{B}_{H}x{W}_moving_stats", mv, 3e-2)
        assert mv <= 3e-2, mv


def test_bf16_training_step_and_greedy(report):
    """Cost, all 329 gradients (median relative L2 - single tensors of this 99-layer net deviate more in bf16) and the parameter
    update formula; greedy decode on a calibrated model compared with the oracle (identity is reported, not required)."""
    B, H, W, cap = 2, 48, 80, 5
    P = O.lively_params(8)
    seed = 777
    eng = engine_with(P, precision="bf16")
    x, xm, y, ym = make_batch(52, B, H, W, cap, False)
    cost = float(eng.forward_backward(x, y, xm, ym, dropout=True, seed=seed))
    g_eng = eng.P.unpack(eng.grads)
    res = {}
    for tag, gates in (("own_gates", None), ("engine_gates", engine_gates(eng, B, H, W))):
        ref_cost, G, names = oracle_grads(P, x, xm, y, ym, seed, gates=gates)
        e = abs(cost - ref_cost) / abs(ref_cost)
        l2s = {n: rel_l2(g_eng[n], G[n]) for n in names if np.abs(G[n]).max() > 1e-12}
        med = float(np.median(list(l2s.values()))); worst = max(l2s, key=l2s.get)
        dot = sum(float((g_eng[n].astype(np.float64) * G[n]).sum()) for n in names)
        cos = dot / np.sqrt(sum(float((g_eng[n].astype(np.float64) ** 2).sum()) for n in names) * sum(float((G[n] ** 2).sum()) for n in names))
        res[tag] = (e, med, cos, worst, l2s[worst])
        report.add("bf16", f"train_cost_{tag}", e, BF16_TOL["cost"], extra=dict(got=cost, ref=ref_cost))
        report.add("bf16", f"grad_median_rel_l2_{tag}", med, BF16_TOL["grad_median_l2"] if gates is not None else 1.0, extra=dict(worst=worst, worst_l2=l2s[worst], cosine_all=cos))
    e, med, cos, worst, wl2 = res["engine_gates"]
    assert e <= BF16_TOL["cost"] and med <= BF16_TOL["grad_median_l2"] and cos >= BF16_TOL["cosine"], res
    p0 = eng.P.flat.clone()
    eng.apply_update(1.0)
    assert float((eng.P.flat - p0).abs().max()) > 0 and bool(torch.isfinite(eng.P.flat).all())
    # greedy decode on a calibrated model
    x1, xm1, _, _ = make_batch(71, 1, 64, 256, 4)
    P2 = O.calibrate_moving(O.lively_params(6), x1, xm1)
    eng.P.load(P2)
    ann1, _ = eng.dense_net(x1, None, training=False)
    toks, lens, _ = eng.greedy_decode(ann1, 25)
    ref_ann, _, _ = O.dense_net(O.to_torch(P2), t64(x1), t64(xm1), False)
    ref_toks = O.greedy_decode(O.to_torch(P2), ref_ann, 25)
    got = toks[0, : int(lens[0])].cpu().tolist()
    same = sum(1 for a, b in zip(got, ref_toks) if a == b)
    report.add("bf16", "greedy_prefix_agreement", 1.0 - same / max(len(ref_toks), 1), 1.0, extra=dict(got=got, ref=ref_toks, identical=got == ref_toks))
