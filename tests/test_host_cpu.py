"""Host-side logic that runs without a GPU: the C-ABI library loads and exports every symbol of
include/wap_abi.h, parameter layout vs the reference variable set, geometry, loader parity,
fail-loud behaviour, and the N>1 data-parallel plumbing over gloo (world_size 2)."""
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import O, ROOT

import wap_b200  # noqa: E402
from wap_b200 import _lib  # noqa: E402


def _header_functions():
    src = open(os.path.join(ROOT, "include", "wap_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wap_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    fns = _header_functions()
    assert len(fns) >= 25
    L = _lib.lib()
    for f in fns:
        assert hasattr(L, f), f"libwap_b200.so does not export {f}"
        assert f in _lib.SIGNATURES, f"_lib.SIGNATURES lacks {f}"
    assert sorted(_lib.SIGNATURES) == fns
    assert L.wap_abi_version() == 1


def test_param_layout_matches_reference_variable_set():
    table = _lib.param_table()
    P = O.init_params(0)
    assert {n for n, _, _, _ in table} == set(P)
    flat_end = 0
    for name, off, shape, flags in table:
        assert tuple(P[name].shape) == tuple(shape), name
        assert off % 4 == 0, name
        if not flags & 2:
            flat_end = max(flat_end, off + int(np.prod(shape)))
    n, nd = _lib.lib().wap_param_floats(), _lib.lib().wap_decay_floats()
    assert n >= flat_end and n - 5428520 < 16 and n % 4 == 0
    l2 = set(O.l2_names(P))
    for name, off, shape, flags in table:
        if flags & 2:
            assert name.endswith(("moving_mean", "moving_variance"))
            continue
        assert bool(flags & 1) == (name in l2), name
        assert (off < nd) == (name in l2), name          # decayed tensors first, BN gamma/beta after
    # TF creation order of the conv kernels (checkpoint-compatible names)
    convs = [n for n, _, _, _ in table if n.endswith("/kernel")]
    assert convs == [O.conv_name(k) for k in range(99)]


@pytest.mark.parametrize("H,W", [(64, 256), (128, 512), (37, 75), (1, 1), (17, 16), (320, 1250)])
def test_annotation_geometry(H, W):
    import ctypes as C
    h, w = C.c_int(), C.c_int()
    assert _lib.lib().wap_annotation_hw(H, W, C.byref(h), C.byref(w)) == 0
    c = lambda v: -(-v // 2)  # noqa: E731
    assert (h.value, w.value) == (c(c(c(c(H)))), c(c(c(c(W)))))
    assert _lib.lib().wap_encoder_workspace_floats(2, H, W, 1) > _lib.lib().wap_encoder_workspace_floats(2, H, W, 0) > 0
    assert _lib.lib().wap_decoder_train_workspace_floats(2, h.value, w.value, 5) > 0
    assert _lib.lib().wap_greedy_workspace_floats(2, h.value, w.value) > _lib.lib().wap_decoder_step_workspace_floats(2, h.value, w.value) > 0


def test_bad_arguments_report_errors():
    L = _lib.lib()
    assert L.wap_annotation_hw(0, 5, None, None) != 0
    assert b"bad size" in L.wap_last_error()
    assert L.wap_param_info(10 ** 6, None, 0, None, None, None, None) != 0
    with pytest.raises(_lib.WapError):
        _lib.check(L.wap_encoder_forward(None, None, None, None, 1, 8, 8, 0, 0, 0, None, None), "wap_encoder_forward")


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_engine_fails_loudly_without_gpu():
    with pytest.raises(_lib.WapError):
        wap_b200.WAPEngine()


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.WapError, match="no CPU fallback"):
        _lib.lib()


# ------------------------------------------------------------------ loader (data.py)
def _fake_corpus(tmp_path, n=23, seed=0):
    rs = np.random.RandomState(seed)
    vocab = {f"t{i}": i for i in range(1, 12)}; vocab["<eol>"] = 0
    feats, lines = {}, []
    for i in range(n):
        H, W = int(rs.randint(8, 40)), int(rs.randint(8, 90))
        feats[f"u{i}"] = rs.randint(0, 256, size=(1, H, W)).astype(np.uint8)
        L = int(rs.randint(1, 9)) if i != 5 else 30           # one over-long caption
        lines.append(f"u{i}\t" + " ".join(f"t{int(t)}" for t in rs.randint(1, 12, size=L)))
    feats["big"] = np.zeros((1, 80, 80), np.uint8); lines.append("big\tt1 t2")    # one over-large image
    (tmp_path / "f.pkl").write_bytes(pickle.dumps(feats))
    (tmp_path / "c.txt").write_text("\n".join(lines) + "\n")
    (tmp_path / "d.txt").write_text("\n".join(f"{k}\t{v}" for k, v in vocab.items()) + "\n")
    return tmp_path


def test_prepare_data_matches_oracle():
    imgs, caps = O.synth_batch(3, 4, 30, 50, 6, ragged=True)
    a = wap_b200.prepare_data(imgs, caps); b = O.prepare_data(imgs, caps)
    for u, v in zip(a, b):
        assert u.dtype == v.dtype and np.array_equal(u, v)
    assert a[2].dtype == np.int64 and a[2].shape == (7, 4) and a[3][:, 0].sum() == 7


def test_data_iterator_bucketing_invariants(tmp_path, capsys):
    d = _fake_corpus(tmp_path)
    lex = wap_b200.load_dict(str(d / "d.txt"))
    assert lex["<eol>"] == 0 and len(lex) == 12
    batches, uids = wap_b200.dataIterator(str(d / "f.pkl"), str(d / "c.txt"), lex, batch_size=4, batch_Imagesize=6000, maxlen=10, maxImagesize=5000)
    assert "u5" not in uids and "big" not in uids and len(uids) == 22
    areas = []
    for xs, ys in batches:
        assert len(xs) == len(ys) <= 4
        if xs:
            assert max(a.shape[1] * a.shape[2] for a in xs) * len(xs) <= 6000
            areas += [a.shape[1] * a.shape[2] for a in xs]
    assert areas == sorted(areas) and len(areas) == 22


@pytest.mark.skipif(not os.path.exists("/root/reference/data.py"), reason="reference checkout not present")
def test_data_iterator_matches_reference_implementation(tmp_path):
    """Runs the reference's own data.py (imported from /root/reference, read-only) on a synthetic
    corpus and compares batch composition, uid order and prepare_data output."""
    d = _fake_corpus(tmp_path, n=41, seed=3)
    code = (
        "import sys, pickle; sys.path.insert(0, '/root/reference'); import data as R\n"
        f"lex = R.load_dict('{d}/d.txt')\n"
        f"b, u = R.dataIterator('{d}/f.pkl', '{d}/c.txt', lex, 4, 6000, 10, 5000)\n"
        "out = [(len(x), [a.shape for a in x], y) for x, y in b]\n"
        "pd = R.prepare_data(b[2][0], b[2][1])\n"
        f"pickle.dump((out, u, pd), open('{d}/ref.pkl', 'wb'))\n")
    subprocess.run([sys.executable, "-c", code], check=True, capture_output=True)
    ref_out, ref_u, ref_pd = pickle.load(open(d / "ref.pkl", "rb"))
    lex = wap_b200.load_dict(str(d / "d.txt"))
    b, u = wap_b200.dataIterator(str(d / "f.pkl"), str(d / "c.txt"), lex, 4, 6000, 10, 5000)
    assert u == ref_u
    assert [(len(x), [a.shape for a in x], y) for x, y in b] == ref_out
    for got, ref in zip(wap_b200.prepare_data(b[2][0], b[2][1]), ref_pd):
        assert got.dtype == ref.dtype and np.array_equal(got, ref)


def test_synthetic_generators_agree():
    from wap_b200 import synth
    for rag in (False, True):
        a = synth.synth_batch(5, 3, 40, 72, 6, rag); b = O.synth_batch(5, 3, 40, 72, 6, rag)
        assert all(np.array_equal(x, y) for x, y in zip(a[0], b[0])) and a[1] == b[1]


# ------------------------------------------------------------------ data-parallel plumbing over gloo
_DP_WORKER = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.environ["WAP_ROOT"])
import wap_b200
from wap_b200 import dp
rank, world, local = dp.init_distributed("gloo")
assert world == 2
n = wap_b200._lib.lib().wap_param_floats()
rs = np.random.RandomState(100 + rank)
g = torch.as_tensor(rs.randn(n).astype(np.float32))            # this replica's flat gradient
mine = g.clone()
dp.allreduce_sum_(g)                                             # what WAPEngine.train_step does for N > 1
other = torch.as_tensor(np.random.RandomState(100 + (1 - rank)).randn(n).astype(np.float32))
assert torch.allclose(g, mine + other, atol=1e-6)
mean = g * (1.0 / world)                                         # grad_scale folded into wap_adadelta_l2_step
assert torch.allclose(mean, (mine + other) / 2, atol=1e-6)
p = torch.full((8,), float(rank)); dp.broadcast_params(p, 0); assert float(p.sum()) == 0.0
m = torch.full((4,), float(rank + 1)); dp.sync_moving_stats(m); assert torch.allclose(m, torch.full((4,), 1.5))
shards = dp.shard_batches(list(range(11)), rank, world)
assert len(shards) == 5 and all(s % world == rank for s in shards)
print(f"DP_OK {rank}\n", end="", flush=True)
'''


def test_dp_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "w.py"; script.write_text(_DP_WORKER)
    env = dict(os.environ, WAP_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("DP_OK") == 2


def test_dp_mean_of_replica_gradients_equals_oracle_semantics():
    """SyncReplicasOptimizer semantics (model-distributed.py:704-717): one update with the MEAN of the
    replica gradients == the oracle's train_step fed the other replica's gradient with inv_world = 1/2."""
    P = O.lively_params(1)
    names = O.trainable_names(P)
    z = {n: np.zeros(P[n].shape) for n in names}
    xa = O.prepare_data(*O.synth_batch(1, 1, 20, 28, 2)); xb = O.prepare_data(*O.synth_batch(2, 1, 20, 28, 2))
    _, Ga, _, _, _ = O.train_step(P, z, dict(z), *xa, 1.0)
    _, Gb, _, _, _ = O.train_step(P, z, dict(z), *xb, 1.0)
    _, G, P2, _, _ = O.train_step(P, z, dict(z), *xa, 1.0, extra_grads=Gb, inv_world=0.5)
    for n in names[:40]:
        np.testing.assert_allclose(G[n], 0.5 * (Ga[n] + Gb[n]), rtol=1e-12, atol=1e-15)


def test_scorer_matches_reference_definition(tmp_path):
    from wap_b200 import scorer
    rs = np.random.RandomState(0)
    rec, lab = {}, {}
    for i in range(30):
        lab[f"k{i}"] = [str(t) for t in rs.randint(0, 6, size=rs.randint(1, 9))]
        rec[f"k{i}"] = lab[f"k{i}"] if i % 3 == 0 else [str(t) for t in rs.randint(0, 6, size=rs.randint(0, 9))]
    assert scorer.score(rec, lab) == pytest.approx(O.wer_exprate(rec, lab))
    (tmp_path / "r.txt").write_text("".join(k + " " + " ".join(v) + "\n" for k, v in rec.items()))
    (tmp_path / "l.txt").write_text("".join(k + "\t" + " ".join(v) + "\n" for k, v in lab.items()))
    wer, sacc = scorer.process(str(tmp_path / "r.txt"), str(tmp_path / "l.txt"), str(tmp_path / "o.wer"))
    lines = (tmp_path / "o.wer").read_text().splitlines()
    assert lines[0] == f"WER {wer}" and lines[1] == f"ExpRate {sacc}"     # format parsed at model-single-GPU.py:702-705


def test_reference_initialisers():
    """ParamStore.init_reference reproduces the reference's initialiser families (model-single-GPU.py:25-42, 67-70, 81):
    Glorot-uniform limits, orthogonal GRU recurrences, gamma ~ U(+-1/sqrt(C)), zero biases / beta, unit moving variance.
    (ParamStore only needs the library's parameter table, so this runs without a GPU.)"""
    import math
    import torch
    import wap_b200
    from wap_b200.engine import ParamStore
    ps = ParamStore(torch.device("cpu"))
    ps.init_reference(seed=3)
    sd = ps.state_dict()
    # Glorot limits: dense matrix [in, out], conv kernels with k^2 fans
    for name, fin, fout in (("W_a", 256, 512), ("U_a", 684, 512), ("Wc", 684, 256), ("embed", 111, 256)):
        lim = math.sqrt(6.0 / (fin + fout))
        w = sd[name]
        assert w.shape == (fin, fout) and np.abs(w).max() <= lim * (1 + 1e-6) and np.abs(w).max() > 0.97 * lim, name
        assert abs(w.std() - lim / math.sqrt(3.0)) < 0.03 * lim, name                  # uniform: std = lim / sqrt(3)
    k3 = sd["conv2d_2/kernel"]                                                        # 3x3, 96 -> 24
    assert k3.shape == (3, 3, 96, 24) and np.abs(k3).max() <= math.sqrt(6.0 / (9 * 96 + 9 * 24)) * (1 + 1e-6)
    k0 = sd["conv2d/kernel"] if "conv2d/kernel" in sd else sd["conv2d_0/kernel"]
    assert k0.shape == (7, 7, 1, 48)
    # orthogonal recurrences (ortho_weight): U U^T = I, and the 512-wide ones are two orthogonal blocks side by side
    for name in ("U_rh", "U_rh_nl"):
        u = sd[name].astype(np.float64)
        assert np.abs(u @ u.T - np.eye(256)).max() < 1e-5, name
    for name in ("U_hz_hr", "U_hz_hr_nl"):
        u = sd[name].astype(np.float64)
        for half in (u[:, :256], u[:, 256:]):
            assert np.abs(half @ half.T - np.eye(256)).max() < 1e-5, name
    # batch norm and biases
    g = sd["batch_normalization_50/gamma"]
    assert np.abs(g).max() <= 1.0 / math.sqrt(g.shape[0]) + 1e-7
    assert not sd["batch_normalization_50/beta"].any() and not sd["batch_normalization_50/moving_mean"].any()
    assert (sd["batch_normalization_50/moving_variance"] == 1.0).all()
    for name in ("U_f_b", "W_a_b", "b_yz_yr", "bo", "ba2h"):
        assert not sd[name].any(), name
    # a different seed gives different weights, the same seed the same
    ps2 = ParamStore(torch.device("cpu")); ps2.init_reference(seed=3)
    assert np.array_equal(ps2.state_dict()["W_a"], sd["W_a"])
    ps2.init_reference(seed=4)
    assert not np.array_equal(ps2.state_dict()["W_a"], sd["W_a"])


def test_param_store_roundtrip_and_checkpoint_names():
    """ParamStore.load / state_dict / pack_into / unpack are exact inverses, keyed by the TF variable names (the checkpoint
    format of WAPEngine.save is state_dict + '__opt__/accum/<name>' + '__opt__/accum_update/<name>')."""
    import torch
    from helpers import O
    from wap_b200.engine import ParamStore
    P = O.lively_params(4)
    ps = ParamStore(torch.device("cpu"))
    ps.load(P)
    sd = ps.state_dict()
    assert set(sd) == set(P)
    for n in P:
        assert np.array_equal(sd[n], np.asarray(P[n], np.float32)), n
    flat = torch.zeros_like(ps.flat)
    tr = {n: np.asarray(P[n], np.float32) * 2 for n in ps.names(trainable_only=True)}
    ps.pack_into(flat, tr)
    back = ps.unpack(flat)
    assert set(back) == set(tr) and all(np.array_equal(back[n], tr[n]) for n in tr)
    # L2-decayed tensors come first: everything before n_decay is a non-batch-norm variable
    for name, off, shape, flags in ps.table:
        if not flags & 2:
            assert (off < ps.n_decay) == (not name.startswith("batch_normalization")), name


def test_pack_batch_layout_and_packed_prefetcher():
    """data.pack_batch (host half of the GPU-side prepare_data): offsets, sizes, captions and raw pixels are where
    wap_prepare_batch / WAPEngine.upload read them; Prefetcher(packed=True) yields the same buffers in order."""
    import wap_b200
    rng = np.random.default_rng(5)
    sizes = [(9, 14), (12, 7), (3, 20)]
    imgs = [rng.integers(0, 256, size=(1, h, w), dtype=np.uint8) for h, w in sizes]
    caps = [[5, 6, 7], [1], [9, 8, 7, 6]]
    pb = wap_b200.pack_batch(imgs, caps)
    assert (pb.B, pb.H, pb.W, pb.T) == (3, 12, 20, 5) and pb.off_pix % 16 == 0
    buf = pb.buf
    offs = buf[:pb.off_hw].view(np.int64); hw = buf[pb.off_hw:pb.off_y].view(np.int32).reshape(3, 2)
    assert hw.tolist() == [list(s) for s in sizes] and offs.tolist() == [0, 126, 210]
    _, _, y, ym = wap_b200.prepare_data(imgs, caps)
    assert np.array_equal(buf[pb.off_y:pb.off_ym].view(np.int32).reshape(5, 3), y.astype(np.int32))
    assert np.array_equal(buf[pb.off_ym:pb.off_ym + 60].view(np.float32).reshape(5, 3), ym)
    for b, im in enumerate(imgs):
        n = im.shape[1] * im.shape[2]
        assert np.array_equal(buf[pb.off_pix + offs[b]: pb.off_pix + offs[b] + n], im.reshape(-1))
    with pytest.raises(ValueError):
        wap_b200.pack_batch(imgs, [[111], [1], [2]])
    got = list(wap_b200.Prefetcher([(imgs, caps), ([], []), (imgs[:1], caps[:1])], packed=True, depth=1))
    assert len(got) == 2 and np.array_equal(got[0].buf, buf) and got[1].B == 1


def test_handles_create_destroy_and_defaults():
    """wap_create / wap_make_current / wap_destroy (host-only: no device needed): per-handle precision and switches, default
    mode 3xTF32, the default handle is untouched by a private one."""
    import ctypes as C
    import wap_b200
    L = wap_b200._lib.lib()
    L.wap_make_current(None)
    d0 = L.wap_get_conv_precision(); p0 = L.wap_get_pdl()
    h1, h2 = C.c_void_p(), C.c_void_p()
    assert L.wap_create(C.byref(h1)) == 0 and L.wap_create(C.byref(h2)) == 0 and h1.value != h2.value
    L.wap_make_current(h1)
    assert L.wap_get_conv_precision() == 2 and L.wap_get_pdl() == 1 and L.wap_get_overlap() == 1
    L.wap_set_pdl(0); L.wap_set_overlap(0)
    L.wap_make_current(h2)
    assert L.wap_get_pdl() == 1 and L.wap_get_overlap() == 1 and L.wap_current() == h2.value
    L.wap_make_current(h1)
    assert L.wap_get_pdl() == 0 and L.wap_get_overlap() == 0
    assert L.wap_set_conv_precision(7) != 0 and b"mode" in L.wap_last_error()
    L.wap_make_current(None)
    assert L.wap_get_conv_precision() == d0 and L.wap_get_pdl() == p0 and not L.wap_current()
    assert L.wap_destroy(h1) == 0 and L.wap_destroy(h2) == 0


def test_prefetcher_order_errors_and_early_close():
    """data.Prefetcher: same batches, same order and same arrays as calling prepare_data inline; empty batches skipped; a worker
    exception reaches the consumer; close() lets the worker exit."""
    import time
    import wap_b200
    from wap_b200 import synth
    batches = []
    for i in range(7):
        imgs, caps = synth.synth_batch(10 + i, 1 + i % 3, 16 + 4 * i, 24 + 8 * (i % 4), 2 + i % 5, ragged=True)
        batches.append((imgs, caps))
    batches.insert(3, ([], []))                                   # an empty shard slot (dp.shard_batches pads with these)
    got = list(wap_b200.Prefetcher(batches, pinned=False, depth=2))
    ref = [wap_b200.prepare_data(bx, by) for bx, by in batches if bx]
    assert len(got) == len(ref) == 7
    for g, r in zip(got, ref):
        assert all(np.array_equal(a, b) for a, b in zip(g, r))

    def bad():
        yield batches[0]
        raise ValueError("corrupt record")
    it = iter(wap_b200.Prefetcher(bad(), pinned=False, depth=1))
    next(it)
    with pytest.raises(ValueError):
        next(it)

    def endless():
        while True:
            yield batches[1]
    pf = wap_b200.Prefetcher(endless(), pinned=False, depth=2)
    it = iter(pf); next(it); next(it)
    pf.close()
    time.sleep(0.5)
    assert not pf._thread.is_alive()


def _reference_functions(path, names, extra_globals=None):
    """Compiles the named top-level functions of a reference source file WITHOUT importing it (the module imports TensorFlow 1.6,
    which cannot be installed here).  The source is read from /root/reference at test time; nothing is copied into the repo."""
    import ast
    src = open(path).read().replace("\t", "    ")
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in keep} == set(names), [n.name for n in keep]
    ns = {"np": np}
    ns.update(extra_globals or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def test_scorer_matches_the_references_own_code(tmp_path):
    """WER / ExpRate: wap_b200.scorer and the oracle's wer_exprate against model-distributed.py:30-78 itself (cmp_result, process),
    executed from /root/reference, on random recognition results incl. empty hypotheses and exact matches."""
    from wap_b200 import scorer
    cmp_result, process = _reference_functions("/root/reference/model-distributed.py", ["cmp_result", "process"])
    rs = np.random.RandomState(5)
    rec, lab = {}, {}
    for i in range(60):
        lab[f"u{i}"] = [str(t) for t in rs.randint(0, 7, size=rs.randint(1, 12))]
        rec[f"u{i}"] = lab[f"u{i}"] if i % 4 == 0 else [str(t) for t in rs.randint(0, 7, size=rs.randint(0, 12))]
    for k in rec:                                                   # edit distance of every pair
        d_ref, n_ref = cmp_result(lab[k], rec[k])
        assert (scorer.edit_distance(lab[k], rec[k]) if hasattr(scorer, "edit_distance") else None) in (None, int(d_ref))
        assert n_ref == len(lab[k])
    (tmp_path / "r.txt").write_text("".join(k + " " + " ".join(v) + "\n" for k, v in rec.items()))
    (tmp_path / "l.txt").write_text("".join(k + "\t" + " ".join(v) + "\n" for k, v in lab.items()))
    process(str(tmp_path / "r.txt"), str(tmp_path / "l.txt"), str(tmp_path / "ref.wer"))
    wer, sacc = scorer.process(str(tmp_path / "r.txt"), str(tmp_path / "l.txt"), str(tmp_path / "ours.wer"))
    assert (tmp_path / "ref.wer").read_text() == (tmp_path / "ours.wer").read_text()           # byte-identical result file
    ow, oe = O.wer_exprate(rec, lab)
    assert ow == pytest.approx(wer) and oe == pytest.approx(sacc)


def test_initialiser_families_match_the_references_own_code():
    """norm_weight / conv_norm_weight / ortho_weight of model-single-GPU.py:25-42, executed from /root/reference: same shapes,
    the same uniform limits and orthogonality as ParamStore.init_reference / oracle.init_params produce."""
    import math
    norm_weight, conv_norm_weight, ortho_weight = _reference_functions(
        "/root/reference/model-single-GPU.py", ["norm_weight", "conv_norm_weight", "ortho_weight"], {"rng": np.random.RandomState(1)})
    P = O.init_params(0)
    for name, fin, fout in (("W_a", 256, 512), ("U_a", 684, 512), ("Wo", 128, 111)):
        r = norm_weight(fin, fout)
        assert r.shape == P[name].shape and r.dtype == np.float32
        lim = math.sqrt(6.0 / (fin + fout))
        assert abs(np.abs(r).max() - lim) < 0.02 * lim and np.abs(P[name]).max() <= lim * (1 + 1e-6) and abs(np.abs(P[name]).max() - lim) < 0.02 * lim
        assert abs(r.std() - P[name].std()) < 0.03 * lim
    rk = conv_norm_weight(96, 24, [3, 3])
    assert rk.shape == P["conv2d_2/kernel"].shape == (3, 3, 96, 24)
    assert abs(np.abs(rk).max() - np.abs(P["conv2d_2/kernel"]).max()) < 0.02 * np.abs(rk).max()
    u = ortho_weight(256).astype(np.float64)
    assert np.abs(u @ u.T - np.eye(256)).max() < 1e-4
    for name in ("U_rh", "U_rh_nl"):
        v = P[name].astype(np.float64)
        assert v.shape == (256, 256) and np.abs(v @ v.T - np.eye(256)).max() < 1e-4


def test_beam_search_restatement_against_the_references_own_get_sample():
    """Row a14: the reference's OWN beam-search bookkeeping (WAP.get_sample, model-single-GPU.py:442-542, compiled from
    /root/reference without importing TensorFlow) is run with a fake `session` whose `run` is the oracle's get_word; the hypotheses
    and scores it returns must equal the oracle's restatement (oracle.beam_decode), which is what the GPU tests compare
    WAPEngine.get_sample with.  Also pins the -1 => zero-embedding start and the length-normalised pick (S:671-672)."""
    import ast
    import copy
    import torch
    path = "/root/reference/model-single-GPU.py"
    tree = ast.parse(open(path).read().replace("\t", "    "))
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "WAP")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "get_sample")
    sent = {k: k for k in ("anno", "infer_y", "alpha_past", "h_pre", "if_trainning")}       # feed_dict keys: module globals of the reference
    ns = {"np": np, "copy": copy, "argmax": False, **sent}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    ref_get_sample = ns["get_sample"]

    P = O.lively_params(6)
    Pt = O.to_torch(P)

    class Session:                                     # sess.run([p, w, h, alpha], feed_dict) of S:473 served by the oracle step
        def run(self, fetches, feed_dict):
            assert fetches == ["p", "w", "h", "alpha"] and feed_dict["if_trainning"] is False
            p, h, ap, _ = O.get_word(Pt, torch.as_tensor(feed_dict["infer_y"]), torch.as_tensor(feed_dict["h_pre"], dtype=torch.float64),
                                     torch.as_tensor(feed_dict["alpha_past"], dtype=torch.float64), torch.as_tensor(feed_dict["anno"], dtype=torch.float64))
            p32 = p.detach().to(torch.float32).numpy()
            return p32, p32.argmax(1), h.detach().numpy(), ap.detach().numpy()

    rs = np.random.RandomState(11)
    for case, (h, w, k, maxlen) in enumerate([(4, 8, 10, 12), (3, 5, 3, 8), (5, 9, 10, 6)]):
        ann = np.maximum(rs.randn(1, h, w, O.ANNOT_C), 0).astype(np.float32) * 0.5
        h0 = O.hidden_state_0(Pt, torch.as_tensor(ann, dtype=torch.float64)).numpy()
        samples, scores = ref_get_sample(None, "p", "w", "h", "alpha", ann, h0, k, maxlen, False, Session(), False)
        best, o_samples, o_scores = O.beam_decode(Pt, torch.as_tensor(ann, dtype=torch.float64), k, maxlen)
        assert [list(map(int, s)) for s in samples] == [list(map(int, s)) for s in o_samples], case
        assert np.allclose(np.array(scores, np.float64), np.array(o_scores, np.float64), rtol=1e-6, atol=1e-6), case
        norm = np.array(scores) / np.array([len(s) for s in samples])                     # S:671-672
        assert list(map(int, samples[int(norm.argmin())])) == list(map(int, best)), case


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first): one JSON line with the contract keys, the CPU oracle port as the thing
    timed (`kind: port`), no host<->device bytes, exit code 0 - runnable without a GPU."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["higher_is_better"] is True and d["value"] > 0 and "workload" in d["config"]
