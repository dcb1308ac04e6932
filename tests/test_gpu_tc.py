"""tcgen05 TF32 tensor-core primitives (through the C ABI) vs torch fp64 references.

Two comparisons per case: against the product of TF32-TRUNCATED inputs (low 13 mantissa bits
cleared: what the tensor core consumes when fed raw fp32; must match to fp32 accumulation accuracy,
proving indexing / swizzle / descriptor logic), and against the exact fp32-input product (the
intrinsic TF32 error, reported)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def _L():
    import wap_b200
    return wap_b200._lib, wap_b200._lib.lib()


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731


@pytest.mark.parametrize("M,N,K,lda_pad,acc,stats", [
    (128, 96, 32, 0, 0, 0), (128, 96, 64, 0, 0, 1), (300, 96, 72, 360, 0, 1), (1000, 24, 864, 0, 0, 1),
    (257, 128, 660, 24, 1, 0), (5000, 216, 432, 0, 0, 1), (131, 300, 600, 84, 0, 1), (64, 32, 8, 0, 0, 0)])
def test_tc_gemm(report, M, N, K, lda_pad, acc, stats):
    lib, L = _L()
    assert L.wap_tc_available() == 1
    g = torch.Generator().manual_seed(M * 7 + N)
    lda = K + lda_pad
    Afull = torch.randn(M, lda, generator=g).cuda()
    Bt = torch.randn(N, K, generator=g).cuda()
    D0 = torch.randn(M, N, generator=g).cuda()
    D = D0.clone() if acc else torch.full((M, N), float("nan"), device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda") if stats else None
    lib.check(L.wap_tc_gemm(_st(), p(Afull), lda, p(Bt), K, p(D), N, M, N, K, acc, p(st)), "wap_tc_gemm")
    torch.cuda.synchronize()
    A = Afull[:, :K]
    ref_t = _trunc_tf32(A).double() @ _trunc_tf32(Bt).double().t() + (D0.double() if acc else 0)
    ref_x = A.double() @ Bt.double().t() + (D0.double() if acc else 0)
    e_t = rel_err(D.cpu().numpy(), ref_t.cpu().numpy()); e_x = rel_err(D.cpu().numpy(), ref_x.cpu().numpy())
    tag = f"tc_gemm_{M}x{N}x{K}_acc{acc}"
    report.add("tc", tag + "_vs_truncated_inputs", e_t, 2e-5); report.add("tc", tag + "_vs_exact(tf32 error)", e_x, 5e-3)
    assert e_t <= 2e-5, (e_t, e_x)
    if stats:
        acc_only = (D.double() - (D0.double() if acc else 0))
        es = rel_err(st[0].cpu().numpy(), acc_only.sum(0).cpu().numpy()); eq = rel_err(st[1].cpu().numpy(), (acc_only ** 2).sum(0).cpu().numpy())
        report.add("tc", tag + "_stats_sum", es, 1e-5); report.add("tc", tag + "_stats_sumsq", eq, 1e-5)
        assert es <= 1e-5 and eq <= 1e-5, (es, eq)


@pytest.mark.parametrize("B,H,W,C,N", [(2, 8, 16, 96, 24), (1, 5, 9, 96, 24), (3, 4, 32, 96, 24), (1, 32, 128, 96, 24), (2, 7, 20, 32, 96)])
def test_tc_conv3x3(report, B, H, W, C, N):
    lib, L = _L()
    g = torch.Generator().manual_seed(B * 100 + H)
    ldx = C + 8
    xfull = torch.randn(B, H, W, ldx, generator=g).cuda()
    w = torch.randn(3, 3, C, N, generator=g).cuda()                       # HWIO
    Nrows = ((N + 31) // 32) * 32
    Wt = torch.zeros(Nrows, 9 * C, device="cuda")
    Wt[:N] = w.reshape(9 * C, N).t()
    D = torch.full((B * H * W, N), float("nan"), device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    lib.check(L.wap_tc_conv3x3(_st(), p(xfull), ldx, B, H, W, C, p(Wt), Nrows, N, p(D), N, p(st)), "wap_tc_conv3x3")
    torch.cuda.synchronize()
    x = xfull[..., :C]
    ref = torch.nn.functional.conv2d(_trunc_tf32(x).double().permute(0, 3, 1, 2), _trunc_tf32(w).double().permute(3, 2, 0, 1), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, N)
    e = rel_err(D.cpu().numpy(), ref.cpu().numpy())
    tag = f"tc_conv3x3_B{B}_{H}x{W}_C{C}_N{N}"
    report.add("tc", tag, e, 2e-5)
    es = rel_err(st[0].cpu().numpy(), D.double().sum(0).cpu().numpy()); eq = rel_err(st[1].cpu().numpy(), (D.double() ** 2).sum(0).cpu().numpy())
    report.add("tc", tag + "_stats", max(es, eq), 1e-5)
    assert e <= 2e-5 and es <= 1e-5 and eq <= 1e-5, (e, es, eq)
