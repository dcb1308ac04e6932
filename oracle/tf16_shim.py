"""TEST INFRASTRUCTURE ONLY.  A minimal EAGER stand-in for the slice of the TensorFlow-1.6 API that the reference's model file
uses (`tf.layers.conv2d`, `tf.layers.batch_normalization`, `tf.layers.dropout`, the pooling layers, `tf.nn.conv2d`, `tf.tensordot`,
`tf.scan`, `tf.Variable`, element-wise ops, reductions, `softmax_cross_entropy_with_logits_v2` ...), backed by torch fp64.

Why: TensorFlow 1.6 cannot be installed here, so the reference's graph code cannot run as shipped.  With this module registered
as `tensorflow`, tests/test_reference_graph_cpu.py executes the reference's OWN classes (Watcher_train.dense_net,
Attender.get_context, Parser.one_time_step / get_ht_ctx, WAP.get_cost) straight from /root/reference and compares them with
oracle/wap_oracle.py.  That pins the oracle's STRUCTURE (op order, slices, gate order, bias placement, maxout pairing, masks, the
shift of the embedding, the scan carry ...) to the reference source; what remains a restatement is only the semantics of the
individual TF ops implemented below, each one documented (asymmetric SAME padding, ceil-mode pools, valid-count
average pool, fused batch norm with biased batch variance / Bessel-corrected moving variance, dropout keep-scale).

Variables: `tf.Variable(value, name=...)` and the variables the layers create (TF default names `conv2d[_k]/kernel`,
`batch_normalization[_k]/{gamma,beta,moving_mean,moving_variance}` in creation order) take their value from `STATE.preset[name]`
when present, so a test can run the reference code on the oracle's parameter set.  Dropout masks come from `STATE.drop(key, shape)`
with the oracle's site keys (the i-th dropout call of a forward pass is site i+1: conv index 1..98, then the logit dropout = 99).
Nothing here is used by the product path."""
from __future__ import annotations

import math
import types
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64
float32 = "float32"; int32 = "int32"; int64 = "int64"
bool = "bool"      # noqa: A001  (tf.bool)


class _State:
    def __init__(self):
        self.reset()

    def reset(self, preset: Optional[Dict[str, np.ndarray]] = None, drop: Optional[Callable] = None, requires_grad: bool = False):
        self.preset = dict(preset or {})
        self.drop = drop
        self.requires_grad = requires_grad                    # trainable variables become autograd leaves (tf.gradients stand-in)
        self.variables: Dict[str, torch.Tensor] = {}          # every variable created, by TF name
        self.moving_updates: Dict[str, torch.Tensor] = {}     # UPDATE_OPS results: new moving statistics
        self.n_conv = 0; self.n_bn = 0; self.n_drop = 0
        self.rng = np.random.RandomState(0)


STATE = _State()


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x))


def _f(x):
    x = _t(x)
    return x.to(DT) if x.is_floating_point() else x


def _var(name, make):
    if name in STATE.preset:
        v = torch.as_tensor(np.asarray(STATE.preset[name])).to(DT).clone()
    else:
        v = _f(make())
    if STATE.requires_grad and not name.endswith(("moving_mean", "moving_variance")):
        v.requires_grad_(True)
    STATE.variables[name] = v
    return v


# ----------------------------------------------------------------------------------------------- variables / initialisers
def Variable(initial_value, name=None, **_):
    assert name is not None
    return _var(name, lambda: initial_value)


def random_uniform_initializer(minval, maxval, dtype=None, **_):
    return lambda shape: STATE.rng.uniform(minval, maxval, size=shape)


def placeholder(*_, **__):
    raise RuntimeError("the shim is eager: construct the reference classes with python values instead of placeholders")


# ----------------------------------------------------------------------------------------------- SAME padding helpers
def _same_pad(n_in, k, s):
    """TF SAME: out = ceil(in/s); total = max((out-1)s + k - in, 0); before = total // 2, after = the rest."""
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    return total // 2, total - total // 2


def _conv_nhwc(x, w_hwio, stride, padding):
    """x [B,H,W,C], w [kh,kw,Cin,Cout] (cross-correlation, like TF)."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    xc = x.permute(0, 3, 1, 2)
    if padding.upper() == "SAME":
        pt, pb = _same_pad(x.shape[1], kh, stride); pl, pr = _same_pad(x.shape[2], kw, stride)
        xc = F.pad(xc, (pl, pr, pt, pb))
    y = F.conv2d(xc, w_hwio.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


# ----------------------------------------------------------------------------------------------- tf.layers
def _conv2d_layer(inputs, filters, kernel_size, strides=1, padding="valid", data_format="channels_last", use_bias=True,
                  kernel_initializer=None, **_):
    assert data_format == "channels_last" and not use_bias
    x = _f(inputs)
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    s = strides if isinstance(strides, int) else strides[0]
    name = "conv2d" + (f"_{STATE.n_conv}" if STATE.n_conv else "") + "/kernel"
    STATE.n_conv += 1
    w = _var(name, lambda: kernel_initializer((kh, kw, x.shape[3], filters)))
    return _conv_nhwc(x, w, s, padding)


def _batch_normalization(inputs, training=False, momentum=0.99, scale=True, gamma_initializer=None, epsilon=1e-3, **_):
    """tf.layers.batch_normalization on a 4-D NHWC tensor (fused kernel, TF >= 1.4): normalises with the batch mean and the BIASED
    batch variance over (N,H,W) when training; moving_mean <- m*mm + (1-m)*mean, moving_variance <- m*mv + (1-m)*UNBIASED variance."""
    x = _f(inputs)
    C = x.shape[-1]
    base = "batch_normalization" + (f"_{STATE.n_bn}" if STATE.n_bn else "")
    STATE.n_bn += 1
    gamma = _var(base + "/gamma", lambda: gamma_initializer((C,)) if gamma_initializer is not None else np.ones(C))
    beta = _var(base + "/beta", lambda: np.zeros(C))
    mm = _var(base + "/moving_mean", lambda: np.zeros(C))
    mv = _var(base + "/moving_variance", lambda: np.ones(C))
    train = training() if callable(training) else training
    if train:
        n = x.numel() // C
        mean = x.mean(dim=(0, 1, 2))
        var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
        STATE.moving_updates[base + "/moving_mean"] = momentum * mm + (1 - momentum) * mean
        STATE.moving_updates[base + "/moving_variance"] = momentum * mv + (1 - momentum) * var * (n / max(n - 1, 1))
    else:
        mean, var = mm, mv
    y = (x - mean) / torch.sqrt(var + epsilon)
    return y * gamma + beta if scale else y + beta


def _dropout_layer(inputs, rate=0.5, training=False, **_):
    """tf.layers.dropout: training only; keep with probability 1-rate and scale the kept values by 1/(1-rate)."""
    x = _f(inputs)
    key = "drop_logit" if STATE.n_drop == 98 else f"drop{STATE.n_drop + 1}"
    STATE.n_drop += 1
    train = training() if callable(training) else training
    if not train or STATE.drop is None:
        return x
    return x * STATE.drop(key, tuple(x.shape)).to(DT) / (1.0 - rate)


def _max_pooling2d(inputs, pool_size, strides, padding="valid", **_):
    """2x2/2 SAME: odd sizes get one padded row/col at the bottom/right; padded cells never win."""
    assert list(pool_size) == [2, 2] and strides == 2 and padding.upper() == "SAME"
    return F.max_pool2d(_f(inputs).permute(0, 3, 1, 2), 2, 2, ceil_mode=True).permute(0, 2, 3, 1)


def _average_pooling2d(inputs, pool_size, strides, padding="valid", **_):
    """2x2/2 SAME: the divisor is the number of VALID cells in the window."""
    assert list(pool_size) == [2, 2] and strides == 2 and padding.upper() == "SAME"
    return F.avg_pool2d(_f(inputs).permute(0, 3, 1, 2), 2, 2, ceil_mode=True, count_include_pad=False).permute(0, 2, 3, 1)


layers = types.SimpleNamespace(conv2d=_conv2d_layer, batch_normalization=_batch_normalization, dropout=_dropout_layer,
                               max_pooling2d=_max_pooling2d, average_pooling2d=_average_pooling2d)


# ----------------------------------------------------------------------------------------------- tf.nn
def _nn_conv2d(input, filter, strides, padding, **_):       # noqa: A002
    assert list(strides) == [1, 1, 1, 1]
    return _conv_nhwc(_f(input), _f(filter), 1, padding)


def _softmax_ce_v2(logits=None, labels=None, **_):
    return -(labels * torch.log_softmax(_f(logits), dim=-1)).sum(-1)


nn = types.SimpleNamespace(
    relu=lambda x: torch.relu(_f(x)), conv2d=_nn_conv2d, softmax=lambda logits=None, **_: torch.softmax(_f(logits), dim=-1),
    embedding_lookup=lambda params, ids: _f(params)[_t(ids).long()],
    softmax_cross_entropy_with_logits_v2=_softmax_ce_v2)


# ----------------------------------------------------------------------------------------------- ops
def tensordot(a, b, axes):
    assert axes == 1
    return torch.tensordot(_f(a), _f(b), dims=1)


def shape(x):
    return list(_t(x).shape)


def reshape(x, shp):
    return _t(x).reshape([int(s) for s in shp])


def _axes(axis):
    return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def reduce_sum(x, axis=None, keepdims=False, **_):
    return _f(x).sum() if axis is None else _f(x).sum(dim=_axes(axis), keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False, **_):
    return _f(x).mean() if axis is None else _f(x).mean(dim=_axes(axis), keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False, **_):
    return _f(x).amax(dim=_axes(axis), keepdim=keepdims)


tanh = lambda x: torch.tanh(_f(x))            # noqa: E731
sigmoid = lambda x: torch.sigmoid(_f(x))      # noqa: E731
exp = lambda x: torch.exp(_f(x))              # noqa: E731
multiply = lambda a, b: _f(a) * _f(b)         # noqa: E731
pow = lambda a, b: _f(a) ** b                 # noqa: E731,A001
squeeze = lambda x, axis=None: _t(x).squeeze(axis)        # noqa: E731


def zeros(shp, dtype=None, **_):
    return torch.zeros([int(s) for s in shp], dtype=DT)


def fill(dims, value):
    return torch.full([int(s) for s in dims], float(value), dtype=DT)


def concat(values, axis):
    return torch.cat([_f(v) for v in values], dim=axis)


def strided_slice(x, begin, end, strides):
    assert all(s == 1 for s in strides)
    sl = tuple(slice(int(b), int(e)) for b, e in zip(begin, end))
    return _t(x)[sl]


def one_hot(indices, depth, **_):
    return F.one_hot(_t(indices).long(), int(depth)).to(DT)


def scan(fn, elems, initializer):
    """tf.scan over the leading axis of every tensor in `elems`; returns every carried value stacked over time."""
    T = _t(elems[0]).shape[0]
    carry, outs = initializer, []
    for t in range(T):
        carry = fn(carry, tuple(e[t] for e in elems))
        # TF tensors are immutable (`x += y` rebinds); torch's `+=` mutates, and the reference accumulates alpha_past that way,
        # so the values handed to the stack are snapshots
        outs.append(tuple(o.clone() if isinstance(o, torch.Tensor) else o for o in carry))
    return tuple(torch.stack([_f(o[i]) for o in outs]) for i in range(len(initializer)))


def multinomial(probs, num_samples, **_):
    """tf.multinomial draws a sample (and the reference feeds it probabilities where logits are expected, S:438); the beam search
    never uses the drawn word, so the stand-in returns the arg-max."""
    return _f(probs).argmax(dim=-1, keepdim=True)


def cond(pred, true_fn, false_fn):
    return true_fn() if pred else false_fn()


# contrib imports at the top of the reference file
contrib = types.SimpleNamespace(layers=types.SimpleNamespace(batch_norm=None), framework=types.SimpleNamespace(arg_scope=None))
GraphKeys = types.SimpleNamespace(UPDATE_OPS="update_ops", TRAINABLE_VARIABLES="trainable_variables")
train = types.SimpleNamespace()


def install():
    """Registers this module as `tensorflow` (+ the two contrib sub-modules the reference imports)."""
    import sys
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    for name, attrs in (("tensorflow.contrib", {}), ("tensorflow.contrib.layers", {"batch_norm": None}),
                        ("tensorflow.contrib.framework", {"arg_scope": None})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m


def uninstall():
    import sys
    for name in ("tensorflow", "tensorflow.contrib", "tensorflow.contrib.layers", "tensorflow.contrib.framework"):
        sys.modules.pop(name, None)
