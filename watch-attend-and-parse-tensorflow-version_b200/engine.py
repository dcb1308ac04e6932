"""Host-side mirror of the reference model classes over the C ABI (include/wap_abi.h).

Reference interface being mirrored (file:line in /root/reference/model-single-GPU.py):
    Watcher_train.dense_net(input_x, mask_x)                      :73-132
    Attender / Parser / WAP constructors (hyper-parameters)       :136-163, :209-242, :311-334
    WAP.get_cost(annotation, y, a_mask, y_mask)                   :337-375
    WAP.get_word(y, h_pre, alpha_past_pre, annotation)            :379-439
    WAP.get_sample(..., k, maxlen, stochastic, ...)               :442-542
    hidden_state_0                                                :590
    one optimiser step  sess.run([cost, trainer], feed_dict)      :640-641 (L2 :594-598, Adadelta :602-607)

torch is used for device memory, streams and torch.distributed only; every FLOP of the model runs
in libwap_b200.so.  There is no CPU / eager fallback.
"""
from __future__ import annotations

import collections
import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

_p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None   # noqa: E731

ADADELTA_RHO, ADADELTA_EPS, L2_COEF = 0.95, 1e-8, 1e-4   # TF defaults (S:602) / S:598
VOCAB, DIM_DEC, DIM_WORD, DIM_ATT, ANNOT_C = 111, 256, 256, 512, 684


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class ParamStore:
    """Flat fp32 device buffers + named views (TF variable names of the reference graph)."""

    def __init__(self, device: torch.device):
        self.device = device
        L = _lib.lib()
        self.table = _lib.param_table()
        self.n = int(L.wap_param_floats())
        self.n_decay = int(L.wap_decay_floats())
        self.n_moving = int(L.wap_moving_floats())
        self.flat = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.moving = torch.zeros(self.n_moving, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        for name, off, shape, flags in self.table:
            src = self.moving if flags & 2 else self.flat
            self.views[name] = src[off: off + int(np.prod(shape))].view(*shape)
        for name, _, _, flags in self.table:
            if name.endswith("moving_variance"):
                self.views[name].fill_(1.0)

    def names(self, trainable_only: bool = False):
        return [n for n, _, _, f in self.table if not (trainable_only and f & 2)]

    def load(self, values: Dict[str, np.ndarray]) -> None:
        missing = [n for n in self.names() if n not in values]
        if missing:
            raise KeyError(f"missing variables: {missing[:5]} ...")
        for name in self.names():
            v = torch.as_tensor(np.asarray(values[name], dtype=np.float32))
            self.views[name].copy_(v.reshape(self.views[name].shape))

    def pack_into(self, flat: torch.Tensor, values: Dict[str, np.ndarray]) -> None:
        """Write per-variable arrays (trainable variables only) into a flat buffer with this layout
        (used for optimiser state: accum / accum_update)."""
        for name, off, shape, flags in self.table:
            if flags & 2 or name not in values:
                continue
            v = torch.as_tensor(np.asarray(values[name], dtype=np.float32)).reshape(-1)
            flat[off: off + v.numel()].copy_(v)

    def unpack(self, flat: torch.Tensor) -> Dict[str, np.ndarray]:
        host = flat.detach().cpu().numpy()
        return {name: host[off: off + int(np.prod(shape))].reshape(shape).copy()
                for name, off, shape, flags in self.table if not flags & 2}

    def state_dict(self) -> Dict[str, np.ndarray]:
        return {n: self.views[n].detach().cpu().numpy().copy() for n in self.names()}

    def init_reference(self, seed: int = 0) -> None:
        """Reference initialisers (S:25-42 norm_weight / conv_norm_weight / ortho_weight; conv limits
        from Watcher_train.bound S:67-70; gamma ~ U(+-1/sqrt(C)) S:81; biases 0)."""
        rs = np.random.RandomState(seed)
        vals: Dict[str, np.ndarray] = {}
        for name, _, shape, flags in self.table:
            if name.endswith("kernel"):
                kh, kw, cin, cout = shape
                fin = cout if (kh == 1 and cout not in (96,)) else cin     # transition quirk (S:120-121)
                lim = math.sqrt(6.0 / (fin * kh * kw + cout * kh * kw))
                vals[name] = rs.uniform(-lim, lim, size=shape)
            elif name.endswith("gamma"):
                g = 1.0 / math.sqrt(shape[0])
                vals[name] = rs.uniform(-g, g, size=shape)
            elif name.endswith("moving_variance"):
                vals[name] = np.ones(shape)
            elif name.endswith(("beta", "moving_mean")) or (len(shape) == 1):
                vals[name] = np.zeros(shape)
            elif name == "alpha_past_filter":
                lim = math.sqrt(6.0 / (121 * 1 + 121 * shape[3]))
                vals[name] = rs.uniform(-lim, lim, size=shape)
            elif name in ("U_hz_hr", "U_hz_hr_nl"):
                vals[name] = np.concatenate([np.linalg.svd(rs.randn(DIM_DEC, DIM_DEC))[0] for _ in range(2)], 1)
            elif name in ("U_rh", "U_rh_nl"):
                vals[name] = np.linalg.svd(rs.randn(DIM_DEC, DIM_DEC))[0]
            elif name == "W_yz_yr":
                lim = math.sqrt(6.0 / (DIM_WORD + DIM_DEC))
                vals[name] = rs.uniform(-lim, lim, size=shape)
            else:
                lim = math.sqrt(6.0 / (shape[0] + shape[1]))
                vals[name] = rs.uniform(-lim, lim, size=shape)
        self.load(vals)


class WAPEngine:
    """Owns parameters, optimiser state and workspaces of one GPU replica."""

    PRECISIONS = {"fp32": 0, "tf32": 1, "tf32x3": 2, "bf16": 3}

    def __init__(self, device: Optional[torch.device] = None, seed: int = 0, init: bool = True, precision: str = "tf32x3",
                 use_graphs: bool = True):
        """precision / use_graphs default to the benchmarked path: error-compensated 3xTF32 tcgen05 convolutions (the
        tensor-core parity mode) and CUDA-graph replay of forward+backward for batch shapes seen at least twice."""
        if not torch.cuda.is_available():
            raise _lib.WapError("WAPEngine needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.wap_create(C.byref(h)), "wap_create")
        self._h = h                      # per-engine handle: precision mode, launch switches, side streams, dropout-seed source
        self.P = ParamStore(self.device)
        if init:
            self.P.init_reference(seed)
        self.grads = torch.zeros_like(self.P.flat)
        self.accum = torch.zeros_like(self.P.flat)
        self.accum_update = torch.zeros_like(self.P.flat)
        self._ws: Dict[tuple, torch.Tensor] = {}
        self._cost = torch.zeros(2, dtype=torch.float64, device=self.device)   # [CE cost, L2 cost]
        self.step_count = 0
        self.dropout_seed = 1234
        # CUDA-graph replay of forward+backward per (shapes, dropout, precision): removes the launch gaps between the ~1700 kernels
        # of a step.  A shape is captured the SECOND time it is seen (the reference's area-bucketed batches rarely repeat a shape:
        # capturing each one once would only add capture cost), LRU-evicted beyond max_graphs.  All graphs share the engine's
        # role-keyed workspaces; growing a workspace invalidates every captured graph (they hold raw pointers).
        self.use_graphs = bool(use_graphs)
        self.max_graphs = 8
        self._graphs: "collections.OrderedDict[tuple, list]" = collections.OrderedDict()
        self._graph_seen: Dict[tuple, int] = {}
        self._seed_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.replayed_launches = 0
        self.world_size = 1
        self.rank = 0
        self.launches_last_step = 0
        self.set_conv_precision(precision)

    @property
    def L(self):
        """The C ABI with this engine's handle bound to the calling thread (wap_make_current)."""
        self._lib.wap_make_current(self._h)
        return self._lib

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._lib.wap_destroy(self._h)
                self._h = None
        except Exception:       # interpreter shutdown
            pass

    # ------------------------------------------------------------------ precision
    def set_conv_precision(self, mode: str) -> None:
        """'fp32' (FFMA), 'tf32' (tcgen05 single pass, 10-bit mantissa products), 'tf32x3' (error-compensated 3xTF32 on
        tcgen05: fp32-grade accuracy, the default) or 'bf16' (bf16 storage + tcgen05 kind::f16, the speed mode of BASELINE
        configs[2]) for the dense-block / transition convolutions.  Per engine (handle)."""
        _lib.check(self.L.wap_set_conv_precision(self.PRECISIONS[mode]), "wap_set_conv_precision")

    def conv_precision(self) -> str:
        return ["fp32", "tf32", "tf32x3", "bf16"][self.L.wap_get_conv_precision()]

    # ------------------------------------------------------------------ helpers
    def _buf(self, key: tuple, n: int, dtype=torch.float32) -> torch.Tensor:
        """Workspace by ROLE ('enc', 'dec', 'ann', ...), grown to the largest size seen (10 % head-room): the area-bucketed
        batches of the reference give almost every step its own (B,H,W,T), so per-shape buffers would never be reused."""
        t = self._ws.get(key)
        if t is None or t.numel() < n:
            if t is not None:
                self._graphs.clear()          # captured graphs hold raw pointers into the buffer being replaced
                del self._ws[key], t
            t = torch.empty(max(int(n * 1.1), 4), dtype=dtype, device=self.device)
            self._ws[key] = t
        return t

    @staticmethod
    def _check_tokens(y) -> None:
        """Token ids index the embedding / logit tables on the device (S:340-341, S:368): reject ids outside [0, 111) while
        they are still on the host (the reference fails in tf.one_hot / the dictionary assert S:490).  -1 (zero embedding,
        S:381-384) is only legal for get_word."""
        if isinstance(y, torch.Tensor):
            if y.is_cuda or y.numel() == 0:
                return
            lo, hi = int(y.min()), int(y.max())
        else:
            a = np.asarray(y)
            if a.size == 0:
                return
            lo, hi = int(a.min()), int(a.max())
        if lo < 0 or hi >= VOCAB:
            raise ValueError(f"token id out of range [0, {VOCAB}): min {lo}, max {hi}")

    def annotation_hw(self, H: int, W: int):
        h, w = C.c_int(), C.c_int()
        _lib.check(self.L.wap_annotation_hw(H, W, C.byref(h), C.byref(w)), "wap_annotation_hw")
        return h.value, w.value

    def _dev(self, a, dtype) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            return a.to(self.device, dtype=dtype, non_blocking=True).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a)).to(self.device, dtype=dtype, non_blocking=True)

    # ------------------------------------------------------------------ Watcher
    def mask_subsample(self, x_mask: torch.Tensor) -> torch.Tensor:
        B, H, W = x_mask.shape
        h, w = self.annotation_hw(H, W)
        a_mask = torch.empty((B, h, w), dtype=torch.float32, device=self.device)
        _lib.check(self.L.wap_mask_subsample(_stream(), _p(x_mask), B, H, W, _p(a_mask)), "wap_mask_subsample")
        return a_mask

    def dense_net(self, x, x_mask=None, training: bool = False, dropout: bool = True, seed: Optional[int] = None):
        """Watcher_train.dense_net (S:73-132): x [B,H,W,1] -> annotation [B,h,w,684], mask [B,h,w]."""
        x = self._dev(x, torch.float32)
        B, H, W = x.shape[0], x.shape[1], x.shape[2]
        h, w = self.annotation_hw(H, W)
        ws = self._buf(("enc",), self.L.wap_encoder_workspace_floats(B, H, W, int(training)))
        ann = torch.empty((B, h, w, ANNOT_C), dtype=torch.float32, device=self.device)
        sd = self.dropout_seed if seed is None else seed
        self.L.wap_set_dropout_seed_source(None)
        _lib.check(self.L.wap_encoder_forward(_stream(), _p(self.P.flat), _p(self.P.moving), _p(x), B, H, W, int(training),
                                              int(bool(dropout) and bool(training)), C.c_uint32(sd & 0xFFFFFFFF), _p(ann), _p(ws)),
                   "wap_encoder_forward")
        a_mask = None
        if x_mask is not None:
            a_mask = self.mask_subsample(self._dev(x_mask, torch.float32))
        return ann, a_mask

    # ------------------------------------------------------------------ decode
    def precompute(self, ann: torch.Tensor, a_mask: Optional[torch.Tensor] = None):
        B, h, w, _ = ann.shape
        watch = torch.empty((B, h, w, DIM_ATT), dtype=torch.float32, device=self.device)
        h0 = torch.empty((B, DIM_DEC), dtype=torch.float32, device=self.device)
        qf = torch.empty((121, DIM_ATT), dtype=torch.float32, device=self.device)
        scratch = self._buf(("mean",), B * ANNOT_C)
        _lib.check(self.L.wap_attn_precompute(_stream(), _p(self.P.flat), _p(ann), _p(a_mask), B, h, w, _p(watch), _p(h0), _p(qf),
                                              _p(scratch)), "wap_attn_precompute")
        return watch, h0, qf

    def hidden_state_0(self, anno: torch.Tensor) -> torch.Tensor:
        """S:590 (unmasked mean)."""
        B, h, w, _ = anno.shape
        h0 = torch.empty((B, DIM_DEC), dtype=torch.float32, device=self.device)
        scratch = self._buf(("mean",), B * ANNOT_C)
        _lib.check(self.L.wap_attn_precompute(_stream(), _p(self.P.flat), _p(anno), None, B, h, w, None, _p(h0), None, _p(scratch)),
                   "wap_attn_precompute")
        return h0

    def get_word(self, y, h_pre, alpha_past_pre, annotation, watch=None, qf=None):
        """WAP.get_word (S:379-439).  annotation may be [1,h,w,684] shared by all k rows (beam)."""
        y = self._dev(y, torch.int32)
        k = y.shape[0]
        _, h, w, _ = annotation.shape
        shared = int(annotation.shape[0] == 1 and k > 1)
        if watch is None or qf is None:
            watch, _, qf = self.precompute(annotation)
        h_pre = self._dev(h_pre, torch.float32); alpha_past_pre = self._dev(alpha_past_pre, torch.float32)
        probs = torch.empty((k, VOCAB), dtype=torch.float32, device=self.device)
        word = torch.empty((k,), dtype=torch.int32, device=self.device)
        h_new = torch.empty((k, DIM_DEC), dtype=torch.float32, device=self.device)
        ap_new = torch.empty((k, h, w), dtype=torch.float32, device=self.device)
        ws = self._buf(("step",), self.L.wap_decoder_step_workspace_floats(k, h, w))
        _lib.check(self.L.wap_decoder_step(_stream(), _p(self.P.flat), _p(annotation), _p(watch), _p(qf), shared, k, h, w, _p(y),
                                           _p(h_pre), _p(alpha_past_pre), _p(probs), _p(word), _p(h_new), _p(ap_new), _p(ws)),
                   "wap_decoder_step")
        return probs, word, h_new, ap_new

    def get_context(self, annotation4ctx, h_t_1, alpha_past4ctx, a_mask=None, watch=None, qf=None):
        """Attender.get_context (S:166-206) -> (context [B,684], alpha [B,h,w], alpha_past' [B,h,w])."""
        B, h, w, _ = annotation4ctx.shape
        if watch is None or qf is None:
            watch, _, qf = self.precompute(annotation4ctx)
        h_t_1 = self._dev(h_t_1, torch.float32); ap = self._dev(alpha_past4ctx, torch.float32)
        am = self._dev(a_mask, torch.float32) if a_mask is not None else None
        ctx = torch.empty((B, ANNOT_C), dtype=torch.float32, device=self.device)
        alpha = torch.empty((B, h, w), dtype=torch.float32, device=self.device); ap_new = torch.empty_like(alpha)
        ws = self._buf(("step",), self.L.wap_decoder_step_workspace_floats(B, h, w))
        _lib.check(self.L.wap_get_context(_stream(), _p(self.P.flat), _p(annotation4ctx), _p(watch), _p(qf), _p(am), B, h, w, _p(h_t_1), _p(ap),
                                          _p(ctx), _p(alpha), _p(ap_new), _p(ws)), "wap_get_context")
        return ctx, alpha, ap_new

    def one_time_step(self, carry, emb_mask, watch=None, qf=None):
        """Parser.one_time_step (S:260-308): carry = (h, ctx, alpha, alpha_past, annotation, a_mask), emb_mask = (emb_y, y_mask)
        -> the same 6-tuple for the next step (tf.scan body of get_ht_ctx, S:246-255)."""
        h_prev, _, _, ap, ann, a_mask = carry
        emb_y, y_mask = emb_mask
        B, h, w, _ = ann.shape
        if watch is None or qf is None:
            watch, _, qf = self.precompute(ann)
        emb_y = self._dev(emb_y, torch.float32); h_prev = self._dev(h_prev, torch.float32); ap = self._dev(ap, torch.float32)
        am = self._dev(a_mask, torch.float32) if a_mask is not None else None
        ym = self._dev(y_mask, torch.float32) if y_mask is not None else None
        h_new = torch.empty((B, DIM_DEC), dtype=torch.float32, device=self.device)
        ctx = torch.empty((B, ANNOT_C), dtype=torch.float32, device=self.device)
        alpha = torch.empty((B, h, w), dtype=torch.float32, device=self.device); ap_new = torch.empty_like(alpha)
        ws = self._buf(("step",), self.L.wap_decoder_step_workspace_floats(B, h, w))
        _lib.check(self.L.wap_one_time_step(_stream(), _p(self.P.flat), _p(ann), _p(watch), _p(qf), _p(am), _p(ym), B, h, w, _p(emb_y), _p(h_prev),
                                            _p(ap), _p(h_new), _p(ctx), _p(alpha), _p(ap_new), _p(ws)), "wap_one_time_step")
        return h_new, ctx, alpha, ap_new, ann, a_mask

    # ------------------------------------------------------------------ checkpoint (the reference has none that fires: SURVEY 5)
    def save(self, path: str, schedule: Optional[dict] = None) -> None:
        """.npz keyed by the TF variable names (+ '__opt__/accum/<name>', '__opt__/accum_update/<name>', '__step__',
        and '__sched__/<key>' for the training-loop schedule: lrate, halfLrFlag, bad_counter, history_errs, epoch, uidx)."""
        out = dict(self.P.state_dict())
        for k, v in (schedule or {}).items():
            out["__sched__/" + k] = np.asarray(v)
        for k, v in self.P.unpack(self.accum).items():
            out["__opt__/accum/" + k] = v
        for k, v in self.P.unpack(self.accum_update).items():
            out["__opt__/accum_update/" + k] = v
        out["__step__"] = np.array(self.step_count)
        np.savez(path, **out)

    def load(self, path: str) -> dict:
        """Restores parameters, moving statistics, Adadelta slots and the step counter; returns the saved schedule dict."""
        z = np.load(path)
        self.P.load({k: z[k] for k in self.P.names()})
        acc = {k[len("__opt__/accum/"):]: z[k] for k in z.files if k.startswith("__opt__/accum/")}
        accu = {k[len("__opt__/accum_update/"):]: z[k] for k in z.files if k.startswith("__opt__/accum_update/")}
        if acc:
            self.P.pack_into(self.accum, acc); self.P.pack_into(self.accum_update, accu)
        if "__step__" in z.files:
            self.step_count = int(z["__step__"])
        return {k[len("__sched__/"):]: z[k] for k in z.files if k.startswith("__sched__/")}

    def decoding(self):
        """Context manager for decoding many images with FROZEN weights (the validation loop, S:661-683): the fused coverage filter
        Q' = alpha_past_filter * U_f is computed once and handed to the decode kernels (wap_set_fused_coverage_filter) instead of being
        recomputed per image.  Do not update / load weights inside the block."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            qf = torch.empty((121, DIM_ATT), dtype=torch.float32, device=self.device)
            dummy = torch.zeros((1, 1, 1, ANNOT_C), dtype=torch.float32, device=self.device)
            _lib.check(self.L.wap_attn_precompute(_stream(), _p(self.P.flat), _p(dummy), None, 1, 1, 1, None, None, _p(qf), None), "wap_attn_precompute")
            self.L.wap_set_fused_coverage_filter(_p(qf))
            try:
                yield self
            finally:
                self.L.wap_set_fused_coverage_filter(None)
        return cm()

    def greedy_decode_images(self, images, maxlen=100, from_host: bool = False):
        """Greedy decode of a sequence of single images the way the reference validates (one unpadded image at a time, S:661-683), as a TWO-STREAM
        PIPELINE: the encoder of image i+1 runs on one stream while the token loop of image i (16 SMs in the cluster kernel) runs on another -
        the two halves use disjoint workspaces and the annotation handed over is a fresh tensor per image.  `images`: device tensors [1,H,W,1] in
        [0,1], or (from_host) pinned uint8 [H,W] host tensors that are uploaded and scaled on the encoder stream.  `maxlen`: int or per-image
        list.  Returns [(tokens [1,maxlen_i] int32, lengths [1] int32)] on the device; call inside `with eng.decoding():` for frozen weights."""
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_pipe_streams"):
            self._pipe_streams = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
        s_enc, s_dec = self._pipe_streams
        s_enc.wait_stream(cur); s_dec.wait_stream(cur)
        out = []
        for i, im in enumerate(images):
            ml = maxlen[i] if isinstance(maxlen, (list, tuple)) else maxlen
            with torch.cuda.stream(s_enc):
                x = (im.to(self.device, non_blocking=True).to(torch.float32) / 255.)[None, :, :, None] if from_host else im
                ann, _ = self.dense_net(x, None, training=False)
                ev = torch.cuda.Event(); ev.record(s_enc)
            with torch.cuda.stream(s_dec):
                s_dec.wait_event(ev)
                ann.record_stream(s_dec)
                toks, lens, _ = self.greedy_decode(ann, ml)
            out.append((toks, lens))
        cur.wait_stream(s_enc); cur.wait_stream(s_dec)
        return out

    def get_sample_images(self, images, k: int = 10, maxlen: int = 100):
        """Beam search (get_sample + pick, S:669-672) over a sequence of single images [1,H,W,1] with the same two-stream pipeline as
        greedy_decode_images; the per-image results stay on the device until the end (one synchronisation for the whole list).  Images whose
        annotation grid exceeds the persistent kernel's limits take the host-bookkeeping path in place.  Returns [picked token list]."""
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_pipe_streams"):
            self._pipe_streams = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
        s_enc, s_dec = self._pipe_streams
        s_enc.wait_stream(cur); s_dec.wait_stream(cur)
        pend = []
        for im in images:
            with torch.cuda.stream(s_enc):
                ann, _ = self.dense_net(im, None, training=False)
                ev = torch.cuda.Event(); ev.record(s_enc)
            _, h, w, _ = ann.shape
            with torch.cuda.stream(s_dec):
                s_dec.wait_event(ev)
                ann.record_stream(s_dec)
                if not self.beam_supported(h, w, k, maxlen):
                    pend.append(self._get_sample_host(ann, k, maxlen)[0]); continue
                tokens = torch.zeros((k, maxlen), dtype=torch.int32, device=self.device)
                lengths = torch.zeros((k,), dtype=torch.int32, device=self.device)
                scores = torch.zeros((k,), dtype=torch.float32, device=self.device)
                meta = torch.zeros((4,), dtype=torch.int32, device=self.device)
                ws = self._buf(("greedy_p",), self.L.wap_greedy_persistent_workspace_floats(k, h, w))
                _lib.check(self.L.wap_beam_decode_persistent(_stream(), _p(self.P.flat), _p(ann), h, w, k, maxlen, _p(tokens), _p(lengths), _p(scores),
                                                             _p(meta), _p(ws)), "wap_beam_decode_persistent")
                pend.append((tokens, lengths, meta))
        cur.wait_stream(s_enc); cur.wait_stream(s_dec)
        torch.cuda.current_stream().synchronize()
        out = []
        for p in pend:
            if isinstance(p, tuple):
                tokens, lengths, meta = p
                best = int(meta[1]); out.append(tokens[best, : int(lengths[best])].tolist())
            else:
                out.append(list(p))
        return out

    def greedy_decode(self, annotation: torch.Tensor, maxlen: int = 100, check_every: int = 8, persistent: Optional[bool] = None,
                      ignore_eol: bool = False):
        """Greedy decode of every image in the batch; returns (tokens [B,maxlen] int32, lengths [B] int32, steps).
        persistent (default: whenever the shape allows): one cooperative kernel per <= 16 images runs the whole
        token loop on the device; otherwise the step-by-step kernels with `check_every` host syncs."""
        B, h, w, _ = annotation.shape
        tokens = torch.zeros((B, maxlen), dtype=torch.int32, device=self.device)
        lengths = torch.zeros((B,), dtype=torch.int32, device=self.device)
        if persistent is None:
            persistent = h * w <= 2048 and (h + 10) * (((w + 7) // 8) * 8 + 10) <= 6144
        if persistent:
            steps_dev = torch.zeros((1,), dtype=torch.int32, device=self.device)
            for b0 in range(0, B, 16):
                nb = min(16, B - b0)
                ws = self._buf(("greedy_p",), self.L.wap_greedy_persistent_workspace_floats(nb, h, w))
                _lib.check(self.L.wap_greedy_decode_persistent(_stream(), _p(self.P.flat), _p(annotation[b0:b0 + nb]), nb, h, w, maxlen,
                                                               int(ignore_eol), _p(tokens[b0:b0 + nb]), _p(lengths[b0:b0 + nb]), _p(ws), _p(steps_dev)),
                           "wap_greedy_decode_persistent")
            return tokens, lengths, steps_dev
        ws = self._buf(("greedy",), self.L.wap_greedy_workspace_floats(B, h, w))
        steps = C.c_int(0)
        _lib.check(self.L.wap_greedy_decode(_stream(), _p(self.P.flat), _p(annotation), B, h, w, maxlen, check_every, _p(tokens),
                                            _p(lengths), _p(ws), C.byref(steps)), "wap_greedy_decode")
        return tokens, lengths, steps.value

    def beam_supported(self, h: int, w: int, k: int, maxlen: int) -> bool:
        return h * w <= 2048 and (h + 10) * (((w + 7) // 8) * 8 + 10) <= 6144 and 1 <= k <= 16 and 1 <= maxlen <= 128

    def get_sample(self, annotation: torch.Tensor, k: int = 10, maxlen: int = 100, on_device: Optional[bool] = None):
        """Beam search of WAP.get_sample(stochastic=False) (S:442-542) + the length-normalised pick (S:671-672) for ONE image
        [1,h,w,684]: returns (picked token list, all samples, their scores).  Default: ONE persistent cooperative kernel runs the
        whole search on the device (wap_beam_decode_persistent: live hypotheses = rows of the step, candidate ranking and hypothesis
        bookkeeping in the kernel); the host reads the finished samples once.  Shapes beyond the kernel's shared-memory carve-up
        (h*w > 2048) take the step-by-step kernels with the bookkeeping on the host."""
        assert annotation.shape[0] == 1
        _, h, w, _ = annotation.shape
        if on_device is None:
            on_device = self.beam_supported(h, w, k, maxlen)
        if not on_device:
            return self._get_sample_host(annotation, k, maxlen)
        tokens = torch.zeros((k, maxlen), dtype=torch.int32, device=self.device)
        lengths = torch.zeros((k,), dtype=torch.int32, device=self.device)
        scores = torch.zeros((k,), dtype=torch.float32, device=self.device)
        meta = torch.zeros((4,), dtype=torch.int32, device=self.device)
        ws = self._buf(("greedy_p",), self.L.wap_greedy_persistent_workspace_floats(k, h, w))
        _lib.check(self.L.wap_beam_decode_persistent(_stream(), _p(self.P.flat), _p(annotation), h, w, k, maxlen, _p(tokens), _p(lengths),
                                                     _p(scores), _p(meta), _p(ws)), "wap_beam_decode_persistent")
        m = meta.cpu().numpy()                                   # the ONE device->host synchronisation of the search
        ns, best = int(m[0]), int(m[1])
        tk, ln, sc = tokens[:ns].cpu().numpy(), lengths[:ns].cpu().numpy(), scores[:ns].cpu().numpy()
        sample = [tk[i, : ln[i]].tolist() for i in range(ns)]
        return sample[best], sample, [np.float32(v) for v in sc]

    def _get_sample_host(self, annotation: torch.Tensor, k: int = 10, maxlen: int = 100):
        """The same search with the k x 111 candidate bookkeeping on the host (numpy, as the reference does it) and one
        wap_decoder_step + D2H of the probabilities per token: kept for annotation grids beyond the persistent kernel's limits
        and as the cross-check of the device search."""
        assert annotation.shape[0] == 1
        watch, h0, qf = self.precompute(annotation)
        _, h, w, _ = annotation.shape
        sample, sample_score = [], []
        live_k, dead_k = 1, 0
        hyp_samples = [[]]
        hyp_scores = np.zeros(1, np.float32)
        next_w = torch.full((1,), -1, dtype=torch.int32, device=self.device)
        next_state = h0
        next_ap = torch.zeros((1, h, w), dtype=torch.float32, device=self.device)
        for _ in range(maxlen):
            p, _, st, ap = self.get_word(next_w, next_state, next_ap, annotation, watch, qf)
            next_p = p.cpu().numpy()
            cand = hyp_scores[:, None] - np.log(next_p)
            flat = cand.flatten()
            ranks = flat.argsort()[: (k - dead_k)]
            ti, wi = ranks // VOCAB, ranks % VOCAB
            costs = flat[ranks]
            new_samples = [hyp_samples[t] + [int(v)] for t, v in zip(ti, wi)]
            hyp_samples, scores, keep = [], [], []
            for i, s in enumerate(new_samples):
                if s[-1] == 0:
                    sample.append(s); sample_score.append(np.float32(costs[i])); dead_k += 1
                else:
                    hyp_samples.append(s); scores.append(np.float32(costs[i])); keep.append(int(ti[i]))
            hyp_scores = np.array(scores, np.float32)
            live_k = len(hyp_samples)
            if live_k < 1 or dead_k >= k:
                break
            idx = torch.as_tensor(keep, device=self.device, dtype=torch.long)
            next_w = torch.as_tensor([s[-1] for s in hyp_samples], dtype=torch.int32, device=self.device)
            next_state = st.index_select(0, idx).contiguous()
            next_ap = ap.index_select(0, idx).contiguous()
        if live_k > 0:
            for i in range(live_k):
                sample.append(hyp_samples[i]); sample_score.append(hyp_scores[i])
        score = np.array(sample_score) / np.array([len(s) for s in sample])
        return sample[int(score.argmin())], sample, sample_score

    # ------------------------------------------------------------------ training
    def get_cost(self, annotation, y, a_mask, y_mask, dropout: bool = False, seed: Optional[int] = None) -> torch.Tensor:
        """WAP.get_cost (S:337-375): returns the CE cost (device float64 scalar, no L2 term)."""
        B, h, w, _ = annotation.shape
        self._check_tokens(y)
        y = self._dev(y, torch.int32); y_mask = self._dev(y_mask, torch.float32)
        T = y.shape[0]
        ws = self._buf(("dec",), self.L.wap_decoder_train_workspace_floats(B, h, w, T))
        sd = self.dropout_seed if seed is None else seed
        self.L.wap_set_dropout_seed_source(None)
        _lib.check(self.L.wap_decoder_train_forward(_stream(), _p(self.P.flat), _p(annotation), _p(a_mask), _p(y), _p(y_mask), B, h, w, T,
                                                    int(dropout), C.c_uint32(sd & 0xFFFFFFFF), _p(self._cost), _p(ws)),
                   "wap_decoder_train_forward")
        return self._cost[0]

    def forward_backward(self, x, y, x_mask, y_mask, dropout: bool = True, seed: Optional[int] = None):
        """Forward + backward of cost (no L2 term, no update): fills self.grads; returns device cost.
        With self.use_graphs the launch sequence is captured once per (shapes, dropout, precision) into a CUDA graph and
        replayed afterwards; the dropout seed then travels through device memory (wap_set_dropout_seed_source)."""
        sd = (self.dropout_seed + 7919 * self.step_count + 104729 * self.rank) if seed is None else seed
        self._check_tokens(y)
        key = (tuple(x.shape), tuple(y.shape), bool(dropout), self.L.wap_get_conv_precision())
        ent = self._graphs.get(key) if self.use_graphs else None
        if ent is None:
            seen = self._graph_seen.get(key, 0) + 1
            if self.use_graphs:
                if len(self._graph_seen) > 4096:
                    self._graph_seen.clear()
                self._graph_seen[key] = seen
            if not self.use_graphs or seen < 2:
                self.L.wap_set_dropout_seed_source(None)
                return self._forward_backward(self._dev(x, torch.float32), self._dev(y, torch.int32), self._dev(x_mask, torch.float32),
                                              self._dev(y_mask, torch.float32), dropout, sd)
        fresh = ent is None
        if fresh:
            mk = lambda a, dt: torch.empty(tuple(a.shape), dtype=dt, device=self.device)     # noqa: E731
            ent = [None, mk(x, torch.float32), mk(y, torch.int32), mk(x_mask, torch.float32), mk(y_mask, torch.float32), 0]
        _, sx, sy, sxm, sym, _ = ent
        for dst, src in ((sx, x), (sy, y), (sxm, x_mask), (sym, y_mask)):
            dst.copy_(src if isinstance(src, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(src)), non_blocking=True)
        self._seed_dev.fill_(int(np.array(sd & 0xFFFFFFFF, dtype=np.uint32).astype(np.int32)))
        self.L.wap_set_dropout_seed_source(_p(self._seed_dev))
        if not fresh:
            self._graphs.move_to_end(key)
            ent[0].replay()
            self.replayed_launches += ent[5]
            return self._cost[0]
        self._forward_backward(sx, sy, sxm, sym, dropout, sd)          # this step, eagerly (also sizes every workspace)
        torch.cuda.current_stream().synchronize()
        n0 = self._lib.wap_launch_count()
        g = torch.cuda.CUDAGraph()
        # thread_local: a prefetch thread's pin_memory() / event queries must not invalidate the capture in progress
        with torch.cuda.graph(g, capture_error_mode="thread_local"):    # capture only: nothing runs, self.grads/_cost keep this step's values
            self._forward_backward(sx, sy, sxm, sym, dropout, sd)
        ent[0], ent[5] = g, int(self._lib.wap_launch_count() - n0)
        while len(self._graphs) >= self.max_graphs:
            self._graphs.popitem(last=False)                            # least recently used
        self._graphs[key] = ent
        return self._cost[0]

    def reserve(self, B: int, H: int, W: int, T: int) -> None:
        """Pre-size every training workspace for the largest batch the caller will feed (B images padded to H x W, T caption rows incl. <eol>).
        The role buffers otherwise grow on demand (`_buf`), and with the reference's area-bucketed batches a growth step (a multi-GB cudaMalloc)
        lands inside training steps until the largest bucket has been seen; a trainer knows the maximum from its data set (the reference filters
        by maxImagesize / maxlen, DATA:15-17), so model-single-GPU.py calls this once before the first epoch."""
        h, w = self.annotation_hw(H, W)
        self._buf(("amask",), B * h * w)
        self._buf(("enc",), self.L.wap_encoder_workspace_floats(B, H, W, 1))
        self._buf(("dec",), self.L.wap_decoder_train_workspace_floats(B, h, w, T))
        self._buf(("ann",), B * h * w * ANNOT_C)
        self._buf(("dann",), B * h * w * ANNOT_C)
        self._buf(("x",), B * H * W); self._buf(("xm",), B * H * W)

    def _forward_backward(self, x, y, x_mask, y_mask, dropout, sd):
        B, H, W = x.shape[0], x.shape[1], x.shape[2]
        T = y.shape[0]
        sd_c = C.c_uint32(sd & 0xFFFFFFFF)
        h, w = self.annotation_hw(H, W)
        st = _stream()
        a_mask = self._buf(("amask",), B * h * w)
        _lib.check(self.L.wap_mask_subsample(st, _p(x_mask), B, H, W, _p(a_mask)), "wap_mask_subsample")
        ews = self._buf(("enc",), self.L.wap_encoder_workspace_floats(B, H, W, 1))
        dws = self._buf(("dec",), self.L.wap_decoder_train_workspace_floats(B, h, w, T))
        ann = self._buf(("ann",), B * h * w * ANNOT_C)
        d_ann = self._buf(("dann",), B * h * w * ANNOT_C)
        self.grads.zero_()
        _lib.check(self.L.wap_encoder_forward(st, _p(self.P.flat), _p(self.P.moving), _p(x), B, H, W, 1, int(dropout), sd_c, _p(ann), _p(ews)),
                   "wap_encoder_forward")
        _lib.check(self.L.wap_decoder_train_forward(st, _p(self.P.flat), _p(ann), _p(a_mask), _p(y), _p(y_mask), B, h, w, T, int(dropout), sd_c,
                                                    _p(self._cost), _p(dws)), "wap_decoder_train_forward")
        _lib.check(self.L.wap_train_backward(st, _p(self.P.flat), _p(self.grads), _p(x), B, H, W, _p(ann), _p(a_mask), _p(y), _p(y_mask), T,
                                             int(dropout), sd_c, _p(d_ann), _p(ews), _p(dws)), "wap_train_backward")
        return self._cost[0]

    def upload(self, pb):
        """GPU half of data.py prepare_data (DATA:100-122) for a `data.PackedBatch`: ONE H2D copy of the packed host buffer,
        then wap_prepare_batch writes x [B,H,W,1] (/255, zero padding) and x_mask; y / y_mask are views of the uploaded
        buffer.  Returns device (x, x_mask, y int32 [T,B], y_mask [T,B]) valid until the next upload()."""
        src = pb.buf if isinstance(pb.buf, torch.Tensor) else torch.as_tensor(pb.buf)
        dev = self._buf(("packed",), (pb.nbytes + 3) // 4 + 4).view(torch.uint8)[: pb.nbytes]
        dev.copy_(src, non_blocking=True)
        B, H, W, T = pb.B, pb.H, pb.W, pb.T
        x = self._buf(("x",), B * H * W)[: B * H * W].view(B, H, W, 1)
        xm = self._buf(("xm",), B * H * W)[: B * H * W].view(B, H, W)
        base = dev.data_ptr()
        _lib.check(self.L.wap_prepare_batch(_stream(), C.c_void_p(base + pb.off_pix), C.c_void_p(base), C.c_void_p(base + pb.off_hw), B, H, W,
                                            _p(x), _p(xm)), "wap_prepare_batch")
        y = dev[pb.off_y: pb.off_ym].view(torch.int32).view(T, B)
        ym = dev[pb.off_ym: pb.off_ym + 4 * T * B].view(torch.float32).view(T, B)
        return x, xm, y, ym

    def train_step_packed(self, pb, lr: float = 1.0, dropout: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        """train_step on a `data.PackedBatch` (raw uint8 images + captions in one pinned host buffer)."""
        x, xm, y, ym = self.upload(pb)
        return self.train_step(x, y, xm, ym, lr=lr, dropout=dropout, seed=seed)

    def launch_count(self) -> int:
        """Kernels launched by the library so far, including those replayed from captured graphs."""
        return int(self._lib.wap_launch_count()) + self.replayed_launches

    def apply_update(self, lr: float) -> torch.Tensor:
        """L2 (S:594-598) + Adadelta (S:602-607) on the flat buffers; returns device L2 cost."""
        n = self.P.n
        l2_ptr = C.c_void_p(self._cost.data_ptr() + 8)
        _lib.check(self.L.wap_adadelta_l2_step(_stream(), _p(self.P.flat), _p(self.grads), _p(self.accum), _p(self.accum_update), n,
                                               self.P.n_decay, C.c_float(lr), C.c_float(ADADELTA_RHO), C.c_float(ADADELTA_EPS),
                                               C.c_float(L2_COEF), C.c_float(1.0 / self.world_size), l2_ptr), "wap_adadelta_l2_step")
        return self._cost[1]

    def train_step(self, x, y, x_mask, y_mask, lr: float = 1.0, dropout: bool = True, seed: Optional[int] = None) -> torch.Tensor:
        """One `sess.run([cost, trainer])` (S:640-641): returns cost incl. the L2 term (device float64)."""
        self.forward_backward(x, y, x_mask, y_mask, dropout, seed)
        if self.world_size > 1:
            torch.distributed.all_reduce(self.grads)      # NCCL sum over NVLink; 1/N folded into the optimiser kernel
        self.apply_update(lr)
        self.step_count += 1
        return self._cost.sum()


# --------------------------------------------------------------------------------------
# reference-named facades (constructor signatures of model-single-GPU.py)
# --------------------------------------------------------------------------------------
class Watcher_train:
    def __init__(self, blocks, level, growth_rate, training, dropout_rate=0.2, dense_channels=0, transition=0.5,
                 input_conv_filters=48, input_conv_stride=2, input_conv_kernel=(7, 7), engine: Optional[WAPEngine] = None):
        if (blocks, level, growth_rate, float(dropout_rate), float(transition), input_conv_filters, input_conv_stride,
                tuple(input_conv_kernel)) != (3, 16, 24, 0.2, 0.5, 48, 2, (7, 7)):
            raise ValueError("the sm_100a kernels are specialised for the reference configuration "
                             "(blocks=3, level=16, growth_rate=24, S:574)")
        self.training = training
        self.engine = engine or WAPEngine()
        self.dense_channels = ANNOT_C

    def dense_net(self, input_x, mask_x):
        training = self.training() if callable(self.training) else bool(self.training)
        return self.engine.dense_net(input_x, mask_x, training=training)


class Attender:
    def __init__(self, channels, dim_decoder, dim_attend, engine: Optional[WAPEngine] = None):
        if (channels, dim_decoder, dim_attend) != (ANNOT_C, DIM_DEC, DIM_ATT):
            raise ValueError("Attender is specialised for (684, 256, 512) (S:584)")
        self.channels, self.dim_decoder, self.dim_attend = channels, dim_decoder, dim_attend
        self.coverage_kernel = [11, 11]
        self.engine = engine

    def get_context(self, annotation4ctx, h_t_1, alpha_past4ctx, a_mask):
        return self.engine.get_context(annotation4ctx, h_t_1, alpha_past4ctx, a_mask)


class Parser:
    def __init__(self, hidden_dim, word_dim, attender, context_dim, engine: Optional[WAPEngine] = None):
        if (hidden_dim, word_dim, context_dim) != (DIM_DEC, DIM_WORD, ANNOT_C):
            raise ValueError("Parser is specialised for (256, 256, 684) (S:586)")
        self.attender, self.hidden_dim, self.word_dim, self.context_dim = attender, hidden_dim, word_dim, context_dim
        self.engine = engine or attender.engine

    def one_time_step(self, tuple_h0_ctx_alpha_alpha_past_annotation, tuple_emb_mask):
        return self.engine.one_time_step(tuple_h0_ctx_alpha_alpha_past_annotation, tuple_emb_mask)

    def get_ht_ctx(self, emb_y, target_hidden_state_0, annotations, a_m, y_m):
        """tf.scan over one_time_step (S:246-255): returns the stacked (h, ctx, alpha, alpha_past) of every step."""
        B, h, w, _ = annotations.shape
        watch, _, qf = self.engine.precompute(annotations)
        dev = self.engine.device
        carry = (target_hidden_state_0, None, None, torch.zeros((B, h, w), device=dev), annotations, a_m)
        hs, cs, als, aps = [], [], [], []
        for t in range(emb_y.shape[0]):
            carry = self.engine.one_time_step(carry, (emb_y[t], y_m[t] if y_m is not None else None), watch, qf)
            hs.append(carry[0]); cs.append(carry[1]); als.append(carry[2]); aps.append(carry[3])
        return torch.stack(hs), torch.stack(cs), torch.stack(als), torch.stack(aps), annotations, a_m


class WAP:
    def __init__(self, watcher, attender, parser, hidden_dim, word_dim, context_dim, target_dim, training):
        if target_dim != VOCAB:
            raise ValueError("target_dim must be 111 (S:490)")
        self.watcher, self.attender, self.parser, self.training = watcher, attender, parser, training
        self.engine: WAPEngine = watcher.engine
        attender.engine = attender.engine or self.engine
        parser.engine = parser.engine or self.engine

    def get_cost(self, cost_annotation, cost_y, a_m, y_m):
        training = self.training() if callable(self.training) else bool(self.training)
        return self.engine.get_cost(cost_annotation, cost_y, a_m, y_m, dropout=training)

    def get_word(self, sample_y, sample_h_pre, alpha_past_pre, sample_annotation):
        p, w, h, ap = self.engine.get_word(sample_y, sample_h_pre, alpha_past_pre, sample_annotation)
        return p, w, h, ap

    def get_sample(self, ctx0, h_0=None, k=10, maxlen=100, stochastic=False):
        if stochastic:
            toks, lens, _ = self.engine.greedy_decode(ctx0, maxlen)
            n = int(lens[0])
            return [toks[0, :n].tolist()], [0.0]
        best, sample, score = self.engine.get_sample(ctx0, k, maxlen)
        return sample, score
