// Flat parameter layout + encoder geometry (host only).
#include "model.h"
#include <algorithm>

namespace wap {

static size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

static std::string conv_name(int k) { return k == 0 ? "conv2d/kernel" : "conv2d_" + std::to_string(k) + "/kernel"; }
static std::string bn_name(int k, const char* what) {
  return (k == 0 ? std::string("batch_normalization/") : "batch_normalization_" + std::to_string(k) + "/") + what;
}

static ParamLayout build() {
  ParamLayout L;
  size_t off = 0;
  auto add = [&](const std::string& name, std::initializer_list<int64_t> dims, int flags) -> size_t {
    TensorInfo t; t.name = name; t.offset = off; t.ndim = (int)dims.size(); t.flags = flags;
    int i = 0; for (auto d : dims) t.dims[i++] = d;
    for (; i < 4; ++i) t.dims[i] = 1;
    L.tensors.push_back(t);
    size_t o = off; off += pad4(t.numel());
    return o;
  };
  // ---- encoder conv kernels (HWIO), TF creation order (model-single-GPU.py:76-130)
  int k = 0, c = kStemC;
  auto conv = [&](int kind, int cin, int cout, int ks, int block) {
    ConvLayer& cl = L.conv[k];
    cl.k = k; cl.kind = kind; cl.cin = cin; cl.cout = cout; cl.ks = ks; cl.block = block;
    cl.w = add(conv_name(k), {ks, ks, cin, cout}, 1);
    ++k;
  };
  conv(kStem, 1, kStemC, 7, 0);
  for (int b = 0; b < kBlocks; ++b) {
    for (int j = 0; j < kLevels; ++j) {
      conv(kC1, c, kBottle, 1, b);
      conv(kC3, kBottle, kGrowth, 3, b);
      c += kGrowth;
    }
    if (b < kBlocks - 1) { conv(kTr, c, c / 2, 1, b); c = c / 2; }
  }
  // ---- decoder tensors
  const int64_t D = kD, n = kN, m = kMw, na = kNA, K = kVocab;
  L.U_f = add("U_f", {na, na}, 1);            L.U_f_b = add("U_f_b", {na}, 1);
  L.U_a = add("U_a", {D, na}, 1);             L.U_a_b = add("U_a_b", {na}, 1);
  L.W_a = add("W_a", {n, na}, 1);             L.W_a_b = add("W_a_b", {na}, 1);
  L.V_a = add("V_a", {na, 1}, 1);             L.V_a_b = add("V_a_b", {1}, 1);
  L.Q = add("alpha_past_filter", {kCovK, kCovK, 1, na}, 1);
  L.W_yz_yr = add("W_yz_yr", {m, 2 * n}, 1);  L.b_yz_yr = add("b_yz_yr", {2 * n}, 1);
  L.U_hz_hr = add("U_hz_hr", {n, 2 * n}, 1);
  L.W_yh = add("W_yh", {m, n}, 1);            L.b_yh = add("b_yh", {n}, 1);
  L.U_rh = add("U_rh", {n, n}, 1);
  L.U_hz_hr_nl = add("U_hz_hr_nl", {n, 2 * n}, 1);  L.b_hz_hr_nl = add("b_hz_hr_nl", {2 * n}, 1);
  L.W_c_z_r = add("W_c_z_r", {D, 2 * n}, 1);
  L.U_rh_nl = add("U_rh_nl", {n, n}, 1);      L.b_rh_nl = add("b_rh_nl", {n}, 1);
  L.W_c_h_nl = add("W_c_h_nl", {D, n}, 1);
  L.embed = add("embed", {K, m}, 1);
  L.Wa2h = add("Wa2h", {D, n}, 1);            L.ba2h = add("ba2h", {n}, 1);
  L.Wc = add("Wc", {D, m}, 1);                L.bc = add("bc", {m}, 1);
  L.Wh = add("Wh", {n, m}, 1);                L.bh = add("bh", {m}, 1);
  L.Wy = add("Wy", {m, m}, 1);                L.by = add("by", {m}, 1);
  L.Wo = add("Wo", {m / 2, K}, 1);            L.bo = add("bo", {K}, 1);
  L.n_decay = off;
  // ---- batch-norm gamma/beta (no L2: names start with 'batch_normalization')
  for (int i = 0; i < kNumConv; ++i) {
    L.conv[i].gamma = add(bn_name(i, "gamma"), {L.conv[i].cout}, 0);
    L.conv[i].beta = add(bn_name(i, "beta"), {L.conv[i].cout}, 0);
  }
  L.n_total = off;
  // ---- moving statistics: separate buffer
  off = 0;
  for (int i = 0; i < kNumConv; ++i) {
    L.conv[i].mmean = add(bn_name(i, "moving_mean"), {L.conv[i].cout}, 2);
    L.conv[i].mvar = add(bn_name(i, "moving_variance"), {L.conv[i].cout}, 2);
  }
  L.n_moving = off;
  return L;
}

const ParamLayout& layout() {
  static const ParamLayout L = build();
  return L;
}

EncGeom make_geom(int B, int H, int W) {
  EncGeom g;
  g.B = B; g.H = H; g.W = W;
  g.Hs = cdiv(H, 2); g.Ws = cdiv(W, 2);
  int h = cdiv(g.Hs, 2), w = cdiv(g.Ws, 2), c = kStemC;
  for (int b = 0; b < kBlocks; ++b) {
    g.h[b] = h; g.w[b] = w; g.M[b] = B * h * w; g.Cin0[b] = c;
    c += kLevels * kGrowth; g.Ctot[b] = c;
    c /= 2; h = cdiv(h, 2); w = cdiv(w, 2);
  }
  return g;
}

EncWorkspace make_enc_workspace(const EncGeom& g, bool training) {
  EncWorkspace ws;
  memset(&ws, 0, sizeof(ws));
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += pad4(n); return o; };
  const ParamLayout& L = layout();
  // tcgen05 weight copies FIRST and always complete (forward + dgrad copies): their offsets then do not depend on the geometry
  // or on `training`, so the device-side layer table (encoder.cu: c_wprep) is a constant of the model (~13 M floats).
  for (int i = 0; i < kNumConv; ++i) {       // tcgen05 weight copies: fwd [N(pad)][K], bwd (dgrad) [Cin][K']
    const ConvLayer& c = L.conv[i];
    ws.wt_bwd_ld[i] = c.cout;
    if (c.kind == kC1 || c.kind == kTr) {
      ws.wt_fwd[i] = take((size_t)c.cout * c.cin); ws.wt_fwd_lo[i] = take((size_t)c.cout * c.cin);
      if (c.kind == kTr) { ws.wt_bwd[i] = take((size_t)c.cin * c.cout); ws.wt_bwd_lo[i] = take((size_t)c.cin * c.cout); }
      if (c.kind == kC1) {
        const int level = (i - (1 + c.block * (2 * kLevels + 1))) / 2;     // 1x1 of level j has conv index base + 2j
        if ((level & 1) == 0) {          // even level: allocate the pair buffer (rows = Cin of the odd partner = cin + growth)
          const size_t n = (size_t)(c.cin + kGrowth) * 2 * kBottle;
          const size_t hi = take(n), lo = take(n);
          ws.wt_bwd[i] = hi + kBottle; ws.wt_bwd_lo[i] = lo + kBottle;           // even level: columns 96..191
          ws.wt_bwd[i + 2] = hi; ws.wt_bwd_lo[i + 2] = lo;                       // odd level : columns 0..95
        }
        ws.wt_bwd_ld[i] = 2 * kBottle;
      }
    } else if (c.kind == kC3) {
      ws.wt_fwd[i] = take((size_t)32 * 9 * kBottle); ws.wt_fwd_lo[i] = take((size_t)32 * 9 * kBottle);
      { ws.wt_bwd[i] = take((size_t)kBottle * 9 * kDy2Ld); ws.wt_bwd_lo[i] = take((size_t)kBottle * 9 * kDy2Ld); }
    }
  }
  ws.stats = take((size_t)kNumConv * 2 * kMaxBnC * 2);   // doubles (offset is a multiple of 4 floats = 16 B)
  ws.red = take(2 * kMaxBnC * 2 * 2);     // two [2][kMaxBnC] double buffers: BN backward n uses buffer n&1 and zeroes the other for n+1
  ws.saved = take((size_t)kNumConv * 2 * kMaxBnC);
  ws.bnbar = take(256);
  ws.stem_raw = take((size_t)g.B * g.Hs * g.Ws * kStemC);
  for (int b = 0; b < kBlocks; ++b) ws.buf[b] = take((size_t)g.M[b] * g.Ctot[b]);
  size_t sy1 = 0, sa1 = 0, sy2 = 0;   // shared scratch for inference
  if (!training) {
    sy1 = take((size_t)g.M[0] * kBottle); sa1 = take((size_t)g.M[0] * kBottle); sy2 = take((size_t)g.M[0] * kGrowth);
  }
  for (int i = 0; i < kNumConv; ++i) {
    const ConvLayer& c = L.conv[i];
    const size_t M = g.M[c.block];
    if (c.kind == kC1) {
      ws.y1[i] = training ? take(M * kBottle) : sy1;
      ws.a1[i] = training ? take(M * kBottle) : sa1;
    } else if (c.kind == kC3) {
      ws.y2[i] = training ? take(M * kGrowth) : sy2;
    } else if (c.kind == kTr) {
      ws.ytr[c.block] = take(M * c.cout);
    }
  }
  if (training) {
    for (int b = 0; b < kBlocks - 1; ++b) ws.gbuf[b] = take((size_t)g.M[b] * g.Ctot[b]);
    ws.dy1 = take((size_t)g.M[0] * 2 * kBottle);      // [M, 192]: odd levels in columns 0..95, even levels in 96..191 (ping-pong + pair GEMM)
    ws.da1 = take((size_t)g.M[0] * kBottle);
    ws.dy2 = take((size_t)g.M[0] * kDy2Ld);
    ws.dy1b = ws.dy1 + kBottle;
    ws.dy2b = take((size_t)g.M[0] * kDy2Ld);
    size_t tr = 0;
    for (int b = 0; b < kBlocks - 1; ++b) tr = std::max(tr, (size_t)g.M[b] * (g.Ctot[b] / 2));
    ws.dytr = take(tr);
    ws.gstem = take((size_t)g.B * g.Hs * g.Ws * kStemC);
  }
  ws.total = off;
  return ws;
}

}  // namespace wap
