"""Concurrent kernel timeline of one segment of the training step (CUPTI through torch.profiler - kernels launched through the C ABI are
recorded like any other).  Unlike the ncu launch list (serialised, cold caches) this shows what overlaps what and the gaps in between.
usage: python tools/timeline.py [dec_bwd|dec_fwd|enc_fwd|enc_bwd] [precision] [fuse_mask] -> gpurun_out/timeline_<seg>.txt"""
import sys, os, json, ctypes as C, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wap_b200
from wap_b200 import synth
from wap_b200.engine import _p, _stream, ANNOT_C
seg = sys.argv[1] if len(sys.argv) > 1 else "dec_bwd"
eng = wap_b200.WAPEngine(seed=0); eng.set_conv_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32x3")
L = eng.L
if len(sys.argv) > 3: L.wap_debug_set_fuse_gates(int(sys.argv[3]))
B, H, W, T = 8, 128, 512, 40
imgs, caps = synth.synth_batch(0, B, H, W, T)
x, xm, y, ym = wap_b200.prepare_data(imgs, caps)
x, xm = torch.as_tensor(x).cuda(), torch.as_tensor(xm).cuda(); y, ym = torch.as_tensor(y).cuda().int(), torch.as_tensor(ym).cuda()
T = y.shape[0]
h, w = eng.annotation_hw(H, W)
a_mask = eng.mask_subsample(xm)
ews = eng._buf(("enc",), L.wap_encoder_workspace_floats(B, H, W, 1)); dws = eng._buf(("dec",), L.wap_decoder_train_workspace_floats(B, h, w, T))
ann = eng._buf(("ann",), B * h * w * ANNOT_C); d_ann = eng._buf(("dann",), B * h * w * ANNOT_C)
sd = C.c_uint32(5)
segs = collections.OrderedDict([
    ("enc_fwd", lambda st: L.wap_encoder_forward(st, _p(eng.P.flat), _p(eng.P.moving), _p(x), B, H, W, 1, 1, sd, _p(ann), _p(ews))),
    ("dec_fwd", lambda st: L.wap_decoder_train_forward(st, _p(eng.P.flat), _p(ann), _p(a_mask), _p(y), _p(ym), B, h, w, T, 1, sd, _p(eng._cost), _p(dws))),
    ("dec_bwd", lambda st: L.wap_decoder_train_backward(st, _p(eng.P.flat), _p(eng.grads), _p(ann), _p(a_mask), _p(y), _p(ym), B, h, w, T, 1, sd, _p(d_ann), _p(dws))),
    ("enc_bwd", lambda st: L.wap_encoder_backward(st, _p(eng.P.flat), _p(eng.grads), _p(x), B, H, W, 1, sd, _p(ann), _p(d_ann), _p(ews)))])


def run_all(only=None):
    for n, f in segs.items():
        if only is None or n == only: f(_stream())


for _ in range(3): eng.grads.zero_(); run_all()
torch.cuda.synchronize()
# capture the segment into a graph so that host launch cost does not shape the timeline
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
        segs[seg](_stream())
torch.cuda.synchronize()
for _ in range(2): g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay(); torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/timeline_{seg}.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
out = open(f"gpurun_out/timeline_{seg}.txt", "w")
print(f"# {seg}: {len(ev)} device activities, span {(ev[-1]['ts'] + ev[-1]['dur'] - t0):.1f} us", file=out)
print("# start_us  dur_us  stream  name", file=out)
for e in ev:
    print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f}  {e['args'].get('stream', '?'):>3}  {e['name'][:90]}", file=out)
out.close()
os.remove(path)
# per-kernel totals
tot = collections.defaultdict(lambda: [0.0, 0])
for e in ev: tot[e["name"][:70]][0] += e["dur"]; tot[e["name"][:70]][1] += 1
print(f"{seg}: span {(ev[-1]['ts'] + ev[-1]['dur'] - t0):.1f} us, {len(ev)} activities")
for n, (d, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:14]: print(f"{d:9.1f} us {c:5d}  {n}")
