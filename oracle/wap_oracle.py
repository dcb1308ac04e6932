"""CPU oracle for the Watch-Attend-and-Parse hot path.  TEST INFRASTRUCTURE ONLY.

    *** parity: structure pinned to the reference's own code, TF op semantics unpinned ***
    The reference (wwjwhen/Watch-Attend-and-Parse-tensorflow-version) is a TensorFlow-1.6
    graph.  TensorFlow 1.6 cannot be imported in this environment (Python 3.12, no network),
    the reference ships no tests / golden vectors, and its .pkl inputs are git-LFS pointers.
    This file therefore *restates* the graph op-for-op from the reference source and from the
    documented TF-1.6 op semantics (SAME padding, fused batch-norm, tf.scan ...).
    What pins it:
      * tests/test_reference_graph_cpu.py executes the reference's OWN classes (Watcher_train.dense_net,
        Attender.get_context, Parser.one_time_step / get_ht_ctx, WAP.get_cost, WAP.get_word, WAP.get_sample)
        from /root/reference on top of oracle/tf16_shim.py - an eager stand-in for the ~40 TF ops they call -
        and requires agreement with this file to 1e-10: variable names and creation order, every slice,
        gate order, bias placement, maxout pairing, embedding shift, masks, scan carry, moving-statistic
        updates, dropout sites, beam-search bookkeeping.  Only the per-op semantics inside the shim remain a
        restatement (nothing here can be compared with real TensorFlow output);
      * tests/test_oracle_cpu.py: brute-force loops for the SAME ops and batch norm, fp64 finite-difference
        gradient checks, golden .npz files under tests/golden/ produced by oracle/make_golden.py;
      * tests/test_host_cpu.py: the reference's data.py, scorer, initialisers and get_sample run as shipped.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package never imports it.

Everything is torch-CPU, dtype-generic (float64 for truth, float32 for tolerance calibration
and for the timed CPU baseline).  The formulation is deliberately the reference's *as written*
(U_a·a recomputed every step, un-fused coverage conv → U_f, tf.scan stacking) so that the same
code doubles as the timed "reference CPU path" (kind="port").

Citations: S = /root/reference/model-single-GPU.py, D = /root/reference/model-distributed.py,
DATA = /root/reference/data.py.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# model constants (literals of S:574, S:584-588, S:50-55, S:81-82)
# --------------------------------------------------------------------------------------
BLOCKS, LEVELS, GROWTH = 3, 16, 24          # S:574
STEM_C, BOTTLENECK = 48, 4 * GROWTH         # S:53, S:96
DIM_ATT, DIM_DEC, DIM_WORD, VOCAB = 512, 256, 256, 111   # S:584-588
COV_K = 11                                  # S:142
BN_EPS, BN_MOMENTUM = 1e-4, 0.9             # S:81-82
DROP_RATE = 0.2                             # S:50, S:362
L2_COEF = 1e-4                              # S:598
ADADELTA_RHO, ADADELTA_EPS = 0.95, 1e-8     # tf.train.AdadeltaOptimizer defaults (S:602)


def conv_name(k: int) -> str:
    return "conv2d/kernel" if k == 0 else f"conv2d_{k}/kernel"


def bn_name(k: int, what: str) -> str:
    return (f"batch_normalization/{what}" if k == 0 else f"batch_normalization_{k}/{what}")


def encoder_plan() -> List[dict]:
    """The 99 conv+BN layers in TF variable-creation order (S:76-130).

    kind: 'stem' | 'c1' (1x1 bottleneck) | 'c3' (3x3 growth) | 'tr' (transition)."""
    plan = [dict(k=0, kind="stem", cin=1, cout=STEM_C, ks=7)]
    c, k = STEM_C, 1
    for b in range(BLOCKS):
        for _ in range(LEVELS):
            plan.append(dict(k=k, kind="c1", cin=c, cout=BOTTLENECK, ks=1, block=b)); k += 1
            plan.append(dict(k=k, kind="c3", cin=BOTTLENECK, cout=GROWTH, ks=3, block=b)); k += 1
            c += GROWTH
        if b < BLOCKS - 1:
            plan.append(dict(k=k, kind="tr", cin=c, cout=int(c * 0.5), ks=1, block=b)); k += 1
            c = int(c * 0.5)
    return plan


ANNOT_C = 684   # 48→432→216→600→300→684


# --------------------------------------------------------------------------------------
# initialisers  (S:25-42, S:77-82, S:95-99, S:104-108, S:121-125)
# --------------------------------------------------------------------------------------
def _norm_weight(rs, fan_in, fan_out):                       # S:25-27
    b = np.sqrt(6.0 / (fan_in + fan_out))
    return rs.uniform(-b, b, size=(fan_in, fan_out)).astype(np.float32)


def _conv_norm_weight(rs, nin, nout, ks):                    # S:30-36  (HWIO)
    fi, fo = ks[0] * ks[1] * nin, ks[0] * ks[1] * nout
    b = np.sqrt(6.0 / (fi + fo))
    return rs.uniform(-b, b, size=(ks[0], ks[1], nin, nout)).astype(np.float32)


def _ortho_weight(rs, n):                                    # S:39-42
    u, _, _ = np.linalg.svd(rs.randn(n, n))
    return u.astype(np.float32)


def init_params(seed: int = 0) -> Dict[str, np.ndarray]:
    """All variables of the single-GPU graph, keyed by their TF names, float32 numpy.

    Distributions follow the reference initialisers; the random *stream* is ours
    (TF's initialiser seeds are time-based, S:15)."""
    rs = np.random.RandomState(seed)
    P: Dict[str, np.ndarray] = {}
    for L in encoder_plan():
        k, cin, cout, ks = L["k"], L["cin"], L["cout"], L["ks"]
        # Watcher_train.bound (S:67-70); the transition passes the *compressed* count twice (S:120-121)
        bin_ = cout if L["kind"] == "tr" else cin
        lim = np.sqrt(6.0 / (bin_ * ks * ks + cout * ks * ks))
        P[conv_name(k)] = rs.uniform(-lim, lim, size=(ks, ks, cin, cout)).astype(np.float32)
        g = 1.0 / math.sqrt(cout)                                     # S:81,98,107,124
        P[bn_name(k, "gamma")] = rs.uniform(-g, g, size=(cout,)).astype(np.float32)
        P[bn_name(k, "beta")] = np.zeros((cout,), np.float32)
        P[bn_name(k, "moving_mean")] = np.zeros((cout,), np.float32)
        P[bn_name(k, "moving_variance")] = np.ones((cout,), np.float32)
    D, n, m, na, K = ANNOT_C, DIM_DEC, DIM_WORD, DIM_ATT, VOCAB
    z = lambda *s: np.zeros(s, np.float32)
    # Attender (S:149-163)
    P["U_f"] = _norm_weight(rs, na, na); P["U_f_b"] = z(na)
    P["U_a"] = _norm_weight(rs, D, na); P["U_a_b"] = z(na)
    P["W_a"] = _norm_weight(rs, n, na); P["W_a_b"] = z(na)
    P["V_a"] = _norm_weight(rs, na, 1); P["V_a_b"] = z(1)
    P["alpha_past_filter"] = _conv_norm_weight(rs, 1, na, (COV_K, COV_K))
    # Parser (S:218-242)
    P["W_yz_yr"] = np.concatenate([_norm_weight(rs, m, n), _norm_weight(rs, m, n)], 1); P["b_yz_yr"] = z(2 * n)
    P["U_hz_hr"] = np.concatenate([_ortho_weight(rs, n), _ortho_weight(rs, n)], 1)
    P["W_yh"] = _norm_weight(rs, m, n); P["b_yh"] = z(n)
    P["U_rh"] = _ortho_weight(rs, n)
    P["U_hz_hr_nl"] = np.concatenate([_ortho_weight(rs, n), _ortho_weight(rs, n)], 1); P["b_hz_hr_nl"] = z(2 * n)
    P["W_c_z_r"] = _norm_weight(rs, D, 2 * n)
    P["U_rh_nl"] = _ortho_weight(rs, n); P["b_rh_nl"] = z(n)
    P["W_c_h_nl"] = _norm_weight(rs, D, n)
    # WAP (S:319-333)
    P["embed"] = _norm_weight(rs, K, m)
    P["Wa2h"] = _norm_weight(rs, D, n); P["ba2h"] = z(n)
    P["Wc"] = _norm_weight(rs, D, m); P["bc"] = z(m)
    P["Wh"] = _norm_weight(rs, n, m); P["bh"] = z(m)
    P["Wy"] = _norm_weight(rs, m, m); P["by"] = z(m)
    P["Wo"] = _norm_weight(rs, m // 2, K); P["bo"] = z(K)
    return P


def randomize_small_terms(P: Dict[str, np.ndarray], seed: int = 1) -> Dict[str, np.ndarray]:
    """Test helper: the reference initialises every bias / beta to 0 and moving stats to (0,1),
    which hides indexing bugs.  Give them small random values (moving_variance stays > 0)."""
    rs = np.random.RandomState(seed)
    Q = dict(P)
    for name, v in P.items():
        if name.endswith("moving_variance"):
            Q[name] = rs.uniform(0.5, 1.5, size=v.shape).astype(np.float32)
        elif name.endswith("moving_mean") or name.endswith("/beta"):
            Q[name] = rs.uniform(-0.2, 0.2, size=v.shape).astype(np.float32)
        elif v.ndim == 1 and "/" not in name:
            Q[name] = rs.uniform(-0.1, 0.1, size=v.shape).astype(np.float32)
    return Q


def trainable_names(P) -> List[str]:
    return [n for n in P if not (n.endswith("moving_mean") or n.endswith("moving_variance"))]


def l2_names(P) -> List[str]:
    """S:596-598: every trainable variable whose name does not start with 'batch_normalization'."""
    return [n for n in trainable_names(P) if not n.startswith("batch_normalization")]


def to_torch(P, dtype=torch.float64, requires_grad=False) -> Dict[str, torch.Tensor]:
    out = {}
    for n, v in P.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if requires_grad and not (n.endswith("moving_mean") or n.endswith("moving_variance")):
            t.requires_grad_(True)
        out[n] = t
    return out


# --------------------------------------------------------------------------------------
# TF-1.6 op semantics
# --------------------------------------------------------------------------------------
def same_pad(n_in: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF 'SAME': out=ceil(in/s); total=max((out-1)s+k-in,0); before=total//2; after=rest."""
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x: torch.Tensor, w_hwio: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """tf.layers.conv2d / tf.nn.conv2d, NHWC x HWIO, padding='SAME', cross-correlation."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    _, pt, pb = same_pad(x.shape[1], kh, stride)
    _, pl, pr = same_pad(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def maxpool2_same(x: torch.Tensor) -> torch.Tensor:
    """tf.layers.max_pooling2d(2, 2, 'SAME'): padded cells never win (≡ ceil_mode)."""
    y = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2, ceil_mode=True)
    return y.permute(0, 2, 3, 1).contiguous()


def avgpool2_same(x: torch.Tensor) -> torch.Tensor:
    """tf.layers.average_pooling2d(2, 2, 'SAME'): divisor = number of valid cells."""
    y = F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2, ceil_mode=True, count_include_pad=False)
    return y.permute(0, 2, 3, 1).contiguous()


def batch_norm(x, gamma, beta, mov_mean, mov_var, training: bool):
    """tf.layers.batch_normalization(axis=-1, momentum=.9, epsilon=1e-4), fused 4-D kernel.

    training: batch mean / biased variance over (N,H,W) incl. padded pixels; moving update
    uses the Bessel-corrected variance (FusedBatchNorm).  Returns y, (new_mean, new_var)."""
    if training:
        cnt = x.shape[0] * x.shape[1] * x.shape[2]
        mean = x.mean(dim=(0, 1, 2))
        var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
        y = (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta
        unb = var * (cnt / max(cnt - 1, 1))
        new_mean = mov_mean * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM)
        new_var = mov_var * BN_MOMENTUM + unb.detach() * (1 - BN_MOMENTUM)
        return y, (new_mean, new_var)
    y = (x - mov_mean) * torch.rsqrt(mov_var + BN_EPS) * gamma + beta
    return y, (mov_mean, mov_var)


DropFn = Optional[Callable[[str, Tuple[int, ...]], torch.Tensor]]


def _dropout(x, key: str, drop: DropFn):
    """tf.layers.dropout(rate=.2, training=True): keep-mask / 0.8.  `drop(key, shape)` returns the
    0/1 keep-mask (explicit, because TF's RNG is not reproducible); None ⇒ dropout off."""
    if drop is None:
        return x
    m = drop(key, tuple(x.shape)).to(x.dtype)
    return x * m / (1.0 - DROP_RATE)


# --------------------------------------------------------------------------------------
# Watcher  (S:73-132)
# --------------------------------------------------------------------------------------
def dense_net(P, x, mask_x, training: bool, drop: DropFn = None, taps: Optional[dict] = None, gates: Optional[dict] = None):
    """x [B,H,W,1], mask_x [B,H,W] → annotation [B,h,w,684], mask [B,h,w], new moving stats.
    gates (test support): {conv index k: 0/1 tensor of the BN-output shape}; where given, ReLU(y) is evaluated as y * gate, i.e.
    with gating decisions supplied from outside (the fp32 engine's), so that elements within fp32 rounding of the kink do not make
    a cross-precision gradient comparison ill-posed.  The forward value changes by at most the magnitude of such elements."""
    new_moving: Dict[str, torch.Tensor] = {}

    def cbr(x, k, stride=1, do_drop=True):
        y = conv2d_same(x, P[conv_name(k)], stride)
        if taps is not None:
            taps[f"conv{k}"] = y
        y, (nm, nv) = batch_norm(y, P[bn_name(k, "gamma")], P[bn_name(k, "beta")],
                                 P[bn_name(k, "moving_mean")], P[bn_name(k, "moving_variance")], training)
        new_moving[bn_name(k, "moving_mean")] = nm
        new_moving[bn_name(k, "moving_variance")] = nv
        if taps is not None:
            taps[f"minpre{k}"] = float(y.detach().abs().min())     # distance of the closest element to the ReLU kink
        if gates is not None and k in gates:
            y = y * gates[k].to(y.dtype).reshape(y.shape)
        else:
            y = torch.relu(y)
        if do_drop and training:
            y = _dropout(y, f"drop{k}", drop)
        return y

    x = cbr(x, 0, stride=2, do_drop=False)              # S:78-83
    mask_x = mask_x[:, 0::2, 0::2]                      # S:80
    x = maxpool2_same(x)                                # S:84
    mask_x = mask_x[:, 0::2, 0::2]                      # S:86
    if taps is not None:
        taps["stem"] = x
    dense_out = x
    k = 1
    for b in range(BLOCKS):
        for _ in range(LEVELS):
            x = cbr(x, k); k += 1                       # S:96-101
            x = cbr(x, k); k += 1                       # S:105-110
            dense_out = torch.cat([dense_out, x], dim=3)    # S:111
            x = dense_out
        if taps is not None:
            taps[f"block{b}"] = dense_out
        if b < BLOCKS - 1:
            x = cbr(x, k); k += 1                       # S:122-127
            x = avgpool2_same(x)                        # S:128
            dense_out = x
            mask_x = mask_x[:, 0::2, 0::2]              # S:130
            if taps is not None:
                taps[f"trans{b}"] = x
    return dense_out, mask_x, new_moving


# --------------------------------------------------------------------------------------
# Attender  (S:166-206)
# --------------------------------------------------------------------------------------
def get_context(P, annotation, h_t_1, alpha_past, a_mask, taps: Optional[dict] = None):
    Ft = conv2d_same(alpha_past[:, :, :, None], P["alpha_past_filter"])           # S:171
    coverage = torch.tensordot(Ft, P["U_f"], dims=1) + P["U_f_b"]                 # S:173
    watch = torch.tensordot(annotation, P["U_a"], dims=1) + P["U_a_b"]            # S:177 (every step)
    speller = (h_t_1 @ P["W_a"] + P["W_a_b"])[:, None, None, :]                   # S:181-183
    tanh_v = torch.tanh(coverage + watch + speller)                               # S:185
    e = torch.tensordot(tanh_v, P["V_a"], dims=1) + P["V_a_b"]                    # S:188
    alpha = torch.exp(e).squeeze(3)                                               # S:190-192 (no max-subtraction)
    if a_mask is not None:
        alpha = alpha * a_mask                                                    # S:195
    alpha = alpha / alpha.sum(dim=(1, 2), keepdim=True)                           # S:197
    alpha_past = alpha_past + alpha                                               # S:200
    context = (annotation * alpha[:, :, :, None]).sum(dim=(1, 2))                 # S:202
    if taps is not None:
        taps.update(e=e.squeeze(3), coverage=coverage, watch=watch)
    return context, alpha, alpha_past


# --------------------------------------------------------------------------------------
# Parser  (S:260-308)   — also the body of WAP.get_word (S:387-418) with masks = None
# --------------------------------------------------------------------------------------
def one_time_step(P, h, alpha_past, annotation, a_mask, emb_y, y_mask):
    n = DIM_DEC
    g = torch.sigmoid(emb_y @ P["W_yz_yr"] + P["b_yz_yr"] + h @ P["U_hz_hr"])    # S:269-274
    r1, z1 = g[:, :n], g[:, n:]                                                   # S:276-277 (r first!)
    prop = torch.tanh((h @ P["U_rh"]) * r1 + emb_y @ P["W_yh"] + P["b_yh"])       # S:279-284
    pre_h = z1 * h + (1.0 - z1) * prop                                            # S:286
    if y_mask is not None:
        pre_h = y_mask[:, None] * pre_h + (1.0 - y_mask)[:, None] * h             # S:289
    ctx, alpha, alpha_past = get_context(P, annotation, pre_h, alpha_past, a_mask)    # S:291
    g2 = torch.sigmoid(pre_h @ P["U_hz_hr_nl"] + P["b_hz_hr_nl"] + ctx @ P["W_c_z_r"])    # S:292-294
    r2, z2 = g2[:, :n], g2[:, n:]                                                 # S:296-297
    prop2 = torch.tanh((pre_h @ P["U_rh_nl"] + P["b_rh_nl"]) * r2 + ctx @ P["W_c_h_nl"])  # S:299-302
    h_new = z2 * pre_h + (1.0 - z2) * prop2                                       # S:303
    if y_mask is not None:
        h_new = y_mask[:, None] * h_new + (1.0 - y_mask)[:, None] * pre_h         # S:306
    return h_new, ctx, alpha, alpha_past


def _logits(P, h_t, c_t, y_prev_emb, drop: DropFn):
    """S:354-364 / S:424-435."""
    logit = (y_prev_emb @ P["Wy"] + P["by"]) + (c_t @ P["Wc"] + P["bc"]) + (h_t @ P["Wh"] + P["bh"])
    shp = logit.shape
    logit = logit.reshape(*shp[:-1], shp[-1] // 2, 2).max(dim=-1).values          # maxout pairs (2i,2i+1)
    logit = _dropout(logit, "drop_logit", drop)
    return logit @ P["Wo"] + P["bo"]


# --------------------------------------------------------------------------------------
# WAP.get_cost  (S:337-375)
# --------------------------------------------------------------------------------------
def get_cost(P, annotation, y, a_m, y_m, drop: DropFn = None, taps: Optional[dict] = None):
    """annotation [B,h,w,D]; y int64 [T,B]; a_m [B,h,w]; y_m [T,B] → scalar (no L2 term)."""
    T, B = y.shape
    emb = P["embed"][y.reshape(-1)].reshape(T, B, DIM_WORD)                       # S:340-341
    emb_shift = torch.cat([torch.zeros_like(emb[:1]), emb[:-1]], 0)              # S:342-343
    anno_mean = (annotation * a_m[:, :, :, None]).sum(dim=(1, 2)) / a_m.sum(dim=(1, 2))[:, None]  # S:345
    h = torch.tanh(anno_mean @ P["Wa2h"] + P["ba2h"])                             # S:346-347
    alpha_past = torch.zeros(annotation.shape[:3], dtype=annotation.dtype)        # S:251-252
    hs, cs, alphas = [], [], []
    for t in range(T):                                                            # tf.scan S:248
        h, ctx, alpha, alpha_past = one_time_step(P, h, alpha_past, annotation, a_m, emb_shift[t], y_m[t])
        hs.append(h); cs.append(ctx); alphas.append(alpha)
    h_t, c_t = torch.stack(hs), torch.stack(cs)
    logit = _logits(P, h_t, c_t, emb_shift, drop)                                 # [T,B,K]
    logp = torch.log_softmax(logit.reshape(T * B, VOCAB), dim=-1)
    ce = -logp.gather(1, y.reshape(-1, 1)).squeeze(1)                             # S:368
    ce = (ce * y_m.reshape(-1)).reshape(T, B)                                     # S:371-372
    if taps is not None:
        taps.update(h_t=h_t, c_t=c_t, alphas=torch.stack(alphas), logit=logit, h0=hs[0] if T else None)
    return ce.sum(0).mean()                                                       # S:373-374


def l2_penalty(P):
    """S:594-598."""
    return L2_COEF * sum((P[n] ** 2).sum() for n in l2_names(P))


def hidden_state_0(P, anno):
    """S:590 — unmasked mean (validation path)."""
    return torch.tanh(anno.mean(dim=(1, 2)) @ P["Wa2h"] + P["ba2h"])


# --------------------------------------------------------------------------------------
# WAP.get_word  (S:379-439): one decode step, no masks, dropout off
# --------------------------------------------------------------------------------------
def get_word(P, y, h_pre, alpha_past_pre, annotation):
    """y int64 [k] (y[0]<0 ⇒ zero embedding, S:381-384) → probs [k,111], h [k,256], alpha_past [k,h,w]."""
    if int(y[0]) < 0:
        emb = torch.zeros((1, DIM_WORD), dtype=annotation.dtype)
    else:
        emb = P["embed"][y]
    h, ctx, _, alpha_past = one_time_step(P, h_pre, alpha_past_pre, annotation, None, emb, None)
    logit = _logits(P, h, ctx, emb, None)
    return torch.softmax(logit, dim=-1), h, alpha_past, logit


def greedy_decode(P, annotation, maxlen: int = 100):
    """Intended semantics of the (dead) `stochastic and argmax` branch S:475-483 for ONE image:
    start with y=-1, take argmax, stop after emitting 0 (<eol>) or maxlen steps.
    Returns list of token ids (including the terminating 0 if emitted)."""
    assert annotation.shape[0] == 1
    h = hidden_state_0(P, annotation)
    ap = torch.zeros(annotation.shape[:3], dtype=annotation.dtype)
    y = torch.tensor([-1], dtype=torch.int64)
    out = []
    for _ in range(maxlen):
        p, h, ap, _ = get_word(P, y, h, ap, annotation)
        nw = int(p[0].argmax())
        out.append(nw)
        if nw == 0:
            break
        y = torch.tensor([nw], dtype=torch.int64)
    return out


def beam_decode(P, annotation, k: int = 10, maxlen: int = 100):
    """WAP.get_sample with stochastic=False (S:442-542) for ONE image + the pick at S:671-672.
    Returns (best_sequence, all_samples, all_scores)."""
    assert annotation.shape[0] == 1
    sample, sample_score = [], []
    live_k, dead_k = 1, 0
    hyp_samples = [[]]
    hyp_scores = np.zeros(1, np.float32)
    next_state = hidden_state_0(P, annotation)
    next_ap = torch.zeros(annotation.shape[:3], dtype=annotation.dtype)
    next_w = torch.tensor([-1], dtype=torch.int64)
    for _ in range(maxlen):
        ctx = annotation.expand(live_k, -1, -1, -1)                               # np.tile S:463
        p, st, ap, _ = get_word(P, next_w, next_state, next_ap, ctx)
        next_p = p.detach().to(torch.float32).numpy()                             # session.run → fp32 numpy
        cand = hyp_scores[:, None] - np.log(next_p)                               # S:485
        flat = cand.flatten()
        ranks = flat.argsort()[: (k - dead_k)]                                    # S:487
        ti, wi = ranks // VOCAB, ranks % VOCAB
        costs = flat[ranks]
        new_s = [hyp_samples[t] + [int(w)] for t, w in zip(ti, wi)]
        new_sc = costs.astype(np.float32)
        hyp_samples, hyp_scores_l, keep = [], [], []
        for i, s in enumerate(new_s):
            if s[-1] == 0:
                sample.append(s); sample_score.append(new_sc[i]); dead_k += 1     # S:513-516
            else:
                hyp_samples.append(s); hyp_scores_l.append(new_sc[i]); keep.append(int(ti[i]))
        hyp_scores = np.array(hyp_scores_l, np.float32)
        live_k = len(hyp_samples)
        if live_k < 1 or dead_k >= k:
            break
        next_w = torch.tensor([s[-1] for s in hyp_samples], dtype=torch.int64)
        next_state = st[keep]
        next_ap = ap[keep]
    if live_k > 0:                                                                # S:535-540
        for i in range(live_k):
            sample.append(hyp_samples[i]); sample_score.append(hyp_scores[i])
    score = np.array(sample_score) / np.array([len(s) for s in sample])           # S:671
    return sample[int(score.argmin())], sample, sample_score


# --------------------------------------------------------------------------------------
# one training step: cost (+L2), gradients, Adadelta, BN moving-stat update  (S:592-607, S:640)
# --------------------------------------------------------------------------------------
def train_step(P_np, acc, acc_delta, x, x_mask, y, y_mask, lr: float, drop: DropFn = None,
               dtype=torch.float64, inv_world: float = 1.0, extra_grads=None, gates=None):
    """Returns cost (incl. L2), grads (numpy, incl. L2 term), new params, new acc, new acc_delta.  gates: see dense_net."""
    P = to_torch(P_np, dtype, requires_grad=True)
    xt, xm = torch.as_tensor(x, dtype=dtype), torch.as_tensor(x_mask, dtype=dtype)
    yt, ym = torch.as_tensor(y, dtype=torch.int64), torch.as_tensor(y_mask, dtype=dtype)
    ann, am, new_moving = dense_net(P, xt, xm, True, drop, gates=gates)
    cost = get_cost(P, ann, yt, am, ym, drop) + l2_penalty(P)
    names = trainable_names(P)
    grads = torch.autograd.grad(cost, [P[n] for n in names])
    G = {n: g.detach().numpy() for n, g in zip(names, grads)}
    if extra_grads is not None:          # data-parallel: mean of replica gradients (D:704-717)
        for n in names:
            G[n] = (G[n] + extra_grads[n]) * inv_world
    newP, newA, newD = dict(P_np), dict(acc), dict(acc_delta)
    for n in names:
        g = G[n].astype(np.float64)
        a = ADADELTA_RHO * acc[n] + (1 - ADADELTA_RHO) * g * g
        d = np.sqrt(acc_delta[n] + ADADELTA_EPS) / np.sqrt(a + ADADELTA_EPS) * g
        newD[n] = ADADELTA_RHO * acc_delta[n] + (1 - ADADELTA_RHO) * d * d
        newA[n] = a
        newP[n] = (P_np[n].astype(np.float64) - lr * d).astype(P_np[n].dtype)
    for n, v in new_moving.items():
        newP[n] = v.detach().numpy().astype(P_np[n].dtype)
    return float(cost.detach()), G, newP, newA, newD


# --------------------------------------------------------------------------------------
# data.py prepare_data  (DATA:100-122) — restated for the loader parity test
# --------------------------------------------------------------------------------------
def prepare_data(images, seqs):
    hs = [im.shape[1] for im in images]; ws = [im.shape[2] for im in images]
    ls = [len(s) for s in seqs]
    B, H, W, T = len(images), max(hs), max(ws), max(ls) + 1
    x = np.zeros((B, H, W, 1), np.float32); xm = np.zeros((B, H, W), np.float32)
    y = np.zeros((T, B), np.int64); ym = np.zeros((T, B), np.float32)
    for i, (im, s) in enumerate(zip(images, seqs)):
        x[i, :hs[i], :ws[i], :] = np.moveaxis(im, 0, -1) / 255.0
        xm[i, :hs[i], :ws[i]] = 1.0
        y[:ls[i], i] = s
        ym[:ls[i] + 1, i] = 1.0
    return x, xm, y, ym


# --------------------------------------------------------------------------------------
# WER / ExpRate  (D:30-78)
# --------------------------------------------------------------------------------------
def edit_distance(label, rec) -> int:
    d = np.zeros((len(label) + 1, len(rec) + 1), np.int32)
    d[0, :] = np.arange(len(rec) + 1); d[:, 0] = np.arange(len(label) + 1)
    for i in range(1, len(label) + 1):
        for j in range(1, len(rec) + 1):
            d[i, j] = min(d[i - 1, j - 1] + (label[i - 1] != rec[j - 1]), d[i, j - 1] + 1, d[i - 1, j] + 1)
    return int(d[len(label), len(rec)])


def wer_exprate(rec: Dict[str, list], label: Dict[str, list]) -> Tuple[float, float]:
    td = tl = n = ok = 0
    for key, r in rec.items():
        dist = edit_distance(label[key], r)
        td += dist; tl += len(label[key]); n += 1; ok += (dist == 0)
    return td / tl, ok / n


# --------------------------------------------------------------------------------------
# deterministic synthetic inputs
# --------------------------------------------------------------------------------------
def synth_image(rs: np.random.Generator, H: int, W: int) -> np.ndarray:
    """uint8 [1,H,W], ~10 % ink."""
    return ((rs.random((H, W)) < 0.10) * 255).astype(np.uint8)[None]


def synth_batch(seed: int, B: int, H: int, W: int, caption_len: int, ragged: bool = False):
    rs = np.random.default_rng(seed)
    imgs, caps = [], []
    for i in range(B):
        h = H - (rs.integers(0, H // 3) if ragged and i else 0)
        w = W - (rs.integers(0, W // 3) if ragged and i else 0)
        L = caption_len - (rs.integers(0, max(caption_len // 2, 1)) if ragged and i else 0)
        imgs.append(synth_image(rs, int(h), int(w)))
        caps.append([int(t) for t in rs.integers(1, VOCAB, size=int(L))])
    return imgs, caps


# --------------------------------------------------------------------------------------
# test parameter sets
# --------------------------------------------------------------------------------------
def lively_params(seed: int = 0) -> Dict[str, np.ndarray]:
    """Reference-initialised weights with non-degenerate small terms: |gamma| in [0.5,1.5] with random
    sign, beta/biases/moving_mean non-zero, moving_variance in [0.5,1.5].  (The reference initialises
    all of those to 0 / 1 / tiny values, which hides indexing and sign bugs.)"""
    P = randomize_small_terms(init_params(seed), seed + 1)
    rs = np.random.RandomState(seed + 2)
    for name in P:
        if name.endswith("/gamma"):
            P[name] = (rs.uniform(0.5, 1.5, size=P[name].shape) * rs.choice([-1.0, 1.0], size=P[name].shape)).astype(np.float32)
    return P


def calibrate_moving(P_np: Dict[str, np.ndarray], x: np.ndarray, x_mask: np.ndarray) -> Dict[str, np.ndarray]:
    """Set every moving mean/variance to the batch statistics of `x` (so that inference-mode batch
    norm actually normalises, as it would after training)."""
    P = to_torch(P_np, torch.float64)
    _, _, nm = dense_net(P, torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(x_mask, dtype=torch.float64), True)
    Q = dict(P_np)
    for name, v in nm.items():
        old = P_np[name].astype(np.float64)
        Q[name] = ((v.numpy() - BN_MOMENTUM * old) / (1 - BN_MOMENTUM)).astype(np.float32)
    return Q


def drop_hash_keep(seed: int, site: int, n: int) -> np.ndarray:
    """numpy twin of csrc/common.cuh drop_keep(): keep-mask (0/1 float32) for elements [0,n) of `site`."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint32)
        h = idx ^ np.uint32((seed + site * 0x632BE5AB) & 0xFFFFFFFF)
        h = (h * np.uint32(0x9E3779B1)).astype(np.uint32); h ^= h >> np.uint32(16)
        h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32); h ^= h >> np.uint32(13)
        h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32); h ^= h >> np.uint32(16)
    return ((h >> np.uint32(8)) < np.uint32(13421772)).astype(np.float32)


def hash_drop_fn(seed: int) -> Callable:
    """DropFn reproducing the engine's counter-based dropout: site = conv index k for 'drop{k}',
    99 for 'drop_logit'; element index = flat index of the (pre-pool) activation tensor."""
    def fn(key: str, shape):
        site = 99 if key == "drop_logit" else int(key[4:])
        n = int(np.prod(shape))
        return torch.as_tensor(drop_hash_keep(seed & 0xFFFFFFFF, site, n).reshape(shape))
    return fn


def relu_kink_margin(P_np, x, x_mask, drop: DropFn = None) -> float:
    """Smallest |pre-activation| over every BN->ReLU site of a training-mode forward.  A gradient
    comparison across precisions is only well posed when no element sits on a ReLU kink (with ~2e6
    activations and values O(1) an element within 1e-6 of zero shows up in about every other
    random batch; fp32 then legitimately gates it differently from fp64)."""
    P = to_torch(P_np, torch.float64)
    taps: dict = {}
    dense_net(P, torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(x_mask, dtype=torch.float64), True, drop, taps=taps)
    return min(v for k, v in taps.items() if k.startswith("minpre"))
