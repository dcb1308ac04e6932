// DenseNet "Watcher" encoder in the bf16 speed mode (wap_set_conv_precision(3); BASELINE configs[2]: "batch 16 mixed HxW bf16").
// Same graph as encoder.cu (model-single-GPU.py:73-132); what changes is storage and arithmetic of the 98 dense-block /
// transition convolutions:
//   * block buffers, raw conv outputs (y1, y2, ytr), activations (a1) and the conv-output gradients (dy1, dy2, dytr) are bf16 in
//     HBM: half the bytes of every pass over them;
//   * the convolutions are tcgen05 kind::f16 MMAs on bf16 operands (one MMA per 16 K-elements: no 3xTF32 splitting, no TMEM
//     operand ring), fp32 accumulation in TMEM; weights are re-packed to bf16 K-major copies from the live fp32 weights each call;
//   * batch statistics come from the fp32 accumulators in the conv epilogue, batch-norm arithmetic, the gradient buffers of the
//     dense blocks (G: sums over 16 levels), every weight gradient and the stem stay fp32.
// The block-2 buffer has a row stride of 688 channels (TMA needs 16-byte multiples: 684 bf16 = 1368 B does not qualify) and is
// converted to the fp32 annotation [B,h,w,684] the decoder consumes.  Workspace: the fp32 layout of make_enc_workspace is reused
// (each region holds its bf16 tensor in its first half), so both modes share one allocation.
#include <cuda_bf16.h>

#include <algorithm>

#include "bn_kernels.cuh"
#include "model.h"
#include "tc_gemm.cuh"

namespace wap {

using bf16 = __nv_bfloat16;
int run_stem_conv(cudaStream_t st, const float* x, const float* w, float* out, double* stats, int B, int H, int W, const EncGeom& g);
int run_stem_wgrad(cudaStream_t st, const float* x, const float* gstem, float* dw, int B, int H, int W, const EncGeom& g);
int upload_wprep_table();
int bn_fold(cudaStream_t st, const float* params, const float* moving, float* out);
extern int g_tc_debug_mask;

constexpr int kLd2 = 688;                 // row stride (elements) of the bf16 block-2 buffer
constexpr int kDy2LdB = 32;               // bf16 dy2 rows: 24 channels + 8 zeros (64 B)
static inline int pad8(int n) { return (n + 7) & ~7; }

// ------------------------------------------------------------------------------------ weight packing
// bf16 K-major copies, one launch for all 98 layers (same static layer table as the fp32 modes: offsets in floats into the
// workspace; each bf16 copy lives at the start of the corresponding fp32 region).
//   1x1 / transition fwd : [Cout][pad8(Cin)] (block 2 has Cin = 4 mod 8)  bwd (dgrad): 1x1 pair buffer [Cin][192] (see encoder.cu), transition [Cin][pad8(Cout)]
//   3x3 fwd row-tile     : [96][3*128]  row 24 j + co, column 128 dy + ci    9-tap: [32][9*128]  row co, column 128 tap + ci
//   3x3 bwd (dgrad)      : [96][9*64]   row ci, column 64 tap' + co  (tap' mirrored, co >= 24 zero)
struct WPrepLayerB { int kind, cin, cout, bld, block; long long w, f, b; };
__constant__ WPrepLayerB c_wprep_b[kNumConv];

__global__ void __launch_bounds__(256) prep_conv_weights_bf16_kernel(const float* __restrict__ params, float* __restrict__ wsp, int with_bwd, int rowtile_blocks) {
  pdl_sync();
  const WPrepLayerB L = c_wprep_b[blockIdx.y];
  if (L.kind == kStem) return;
  const float* w = params + L.w;
  bf16* f = reinterpret_cast<bf16*>(wsp + L.f);
  bf16* b = reinterpret_cast<bf16*>(wsp + L.b);
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (L.kind == kC3) {
    if ((rowtile_blocks >> L.block) & 1) {
      for (int i = t0; i < 96 * 384; i += stride) {
        const int n = i / 384, k = i - n * 384, dy = k >> 7, ci = k & 127, j = n / L.cout, co = n - j * L.cout;
        f[i] = __float2bfloat16_rn((n < 3 * L.cout && ci < L.cin) ? __ldg(w + ((size_t)(dy * 3 + j) * L.cin + ci) * L.cout + co) : 0.f);
      }
    } else {
      for (int i = t0; i < 32 * 1152; i += stride) {
        const int n = i / 1152, k = i - n * 1152, tap = k >> 7, ci = k & 127;
        f[i] = __float2bfloat16_rn((n < L.cout && ci < L.cin) ? __ldg(w + ((size_t)tap * L.cin + ci) * L.cout + n) : 0.f);
      }
    }
    if (with_bwd) {
      for (int i = t0; i < L.cin * 576; i += stride) {
        const int ci = i / 576, k = i - ci * 576, tap = k >> 6, co = k & 63;
        b[i] = __float2bfloat16_rn(co < L.cout ? __ldg(w + ((size_t)(8 - tap) * L.cin + ci) * L.cout + co) : 0.f);
      }
    }
  } else {                                                    // 1x1: w[cin][cout]
    const int n_el = L.cin * L.cout;
    const int bld = L.kind == kTr ? ((L.cout + 7) & ~7) : L.bld;
    const int fld = (L.cin + 7) & ~7;
    for (int i = t0; i < n_el; i += stride) {
      const int n = i / L.cin, k = i - n * L.cin;
      f[(size_t)n * fld + k] = __float2bfloat16_rn(__ldg(w + (size_t)k * L.cout + n));
      if (with_bwd) { const int r = i / L.cout, cc = i - r * L.cout; b[(size_t)r * bld + cc] = __float2bfloat16_rn(__ldg(w + i)); }
    }
  }
}

static int prep_conv_weights_bf16(cudaStream_t st, const float* params, float* wsp, bool with_bwd, int rowtile_blocks) {
  static bool uploaded = false;
  if (!uploaded) {
    const ParamLayout& L = layout();
    const EncWorkspace ws = make_enc_workspace(make_geom(1, 16, 16), true);
    WPrepLayerB h[kNumConv];
    for (int i = 0; i < kNumConv; ++i) {
      h[i].kind = L.conv[i].kind; h[i].cin = L.conv[i].cin; h[i].cout = L.conv[i].cout; h[i].bld = ws.wt_bwd_ld[i]; h[i].block = L.conv[i].block;
      h[i].w = (long long)L.conv[i].w; h[i].f = (long long)ws.wt_fwd[i];
      // the fp32 pair buffer interleaves odd / even levels at a 96-FLOAT column offset; in bf16 the same [Cin][192] buffer starts at the
      // hi region of the EVEN level and the odd level's columns are 0..95, the even level's 96..191 (element offsets)
      h[i].b = (long long)ws.wt_bwd[i];
    }
    // 1x1 pairs: both levels write into the buffer that starts at the odd level's fp32 offset (= pair base); the even level's
    // bf16 columns start 96 ELEMENTS (48 floats) into each row
    for (int i = 0; i < kNumConv; ++i)
      if (L.conv[i].kind == kC1) {
        const int level = (i - (1 + L.conv[i].block * (2 * kLevels + 1))) / 2;
        if ((level & 1) == 0) h[i].b = (long long)ws.wt_bwd[i + 2] + kBottle / 2;      // pair base (floats) + 96 bf16
      }
    WAP_CHECK(cudaMemcpyToSymbol(c_wprep_b, h, sizeof(h), 0, cudaMemcpyHostToDevice));
    uploaded = true;
  }
  WAP_CHECK(launch_pdl(prep_conv_weights_bf16_kernel, dim3(24, kNumConv), 256, 0, st, params, wsp, with_bwd ? 1 : 0, rowtile_blocks));
  WAP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------ annotation conversion
__global__ void __launch_bounds__(256) bf16_rows_to_f32_kernel(const bf16* __restrict__ src, int lds, float* __restrict__ dst, int ldd, size_t rows, int cols4) {
  pdl_sync();
  const size_t n = rows * cols4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols4; const int c = (int)(i - r * cols4) * 4;
    float v[4];
    ld4<bf16>(src, r * lds + c, v);
    st4<float>(dst, r * ldd + c, v);
  }
}

static inline bf16* B16(float* base, size_t off) { return reinterpret_cast<bf16*>(base + off); }

// ------------------------------------------------------------------------------------ forward
int encoder_forward_bf16(cudaStream_t st, const float* params, float* moving, const float* x, int B, int H, int W,
                         int training, int dropout, uint32_t seed, float* annotation, float* wsp) {
  const ParamLayout& L = layout();
  const EncGeom g = make_geom(B, H, W);
  const EncWorkspace ws = make_enc_workspace(g, training != 0);
  double* stats = reinterpret_cast<double*>(wsp + ws.stats);
  float* saved = wsp + ws.saved;
  WAP_REQUIRE(tc_available(), "bf16 mode needs cuTensorMapEncodeTiled from the driver");
  WAP_REQUIRE(annotation != nullptr, "bf16 mode: the fp32 annotation output is required");
  if (training) WAP_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * kNumConv * 2 * kMaxBnC, st));
  int rowtile_blocks = 0;
  for (int b = 0; b < kBlocks; ++b) if (g.w[b] <= 128 && (g_tc_debug_mask & 256)) rowtile_blocks |= 1 << b;
  WAP_TRY(upload_wprep_table());
  WAP_TRY(prep_conv_weights_bf16(st, params, wsp, training != 0, rowtile_blocks));
  bf16* buf[kBlocks] = {B16(wsp, ws.buf[0]), B16(wsp, ws.buf[1]), B16(wsp, ws.buf[2])};
  const int ldb[kBlocks] = {g.Ctot[0], g.Ctot[1], kLd2};
  // inference: batch norm + ReLU of the dense levels folded into the conv epilogues (see encoder.cu bn_fold_kernel)
  const bool fold = !training && (g_tc_debug_mask & 512);
  if (fold) WAP_TRY(bn_fold(st, params, moving, saved));
  auto fsc = [&](int kk) -> const float* { return saved + (size_t)kk * 2 * kMaxBnC; };
  auto bn_args = [&](int k, const void* raw, int Hi, int Wi, void* out, int Ho, int Wo, int ld_out, bool drop) {
    const ConvLayer& c = L.conv[k];
    BnActArgs a;
    a.raw = raw; a.B = B; a.Hi = Hi; a.Wi = Wi; a.C = c.cout;
    a.out = out; a.Ho = Ho; a.Wo = Wo; a.ld_out = ld_out;
    a.stats = training ? stats + (size_t)k * 2 * kMaxBnC : nullptr;
    a.mmean = moving + c.mmean; a.mvar = moving + c.mvar;
    a.gamma = params + c.gamma; a.beta = params + c.beta;
    a.saved = training ? saved + (size_t)k * 2 * kMaxBnC : nullptr;
    a.training = training; a.dropout = (training && dropout && drop) ? 1 : 0; a.seed = seed; a.seed_src = seed_src(); a.site = (uint32_t)k;
    a.round_tf32 = 0;
    return a;
  };
  // ---- stem (fp32 CUDA-core conv; pooled activation stored as bf16 into block 0)
  WAP_TRY(run_stem_conv(st, x, params + L.conv[0].w, wsp + ws.stem_raw, training ? stats : nullptr, B, H, W, g));
  WAP_TRY((run_bn_act<float, bf16>(st, 1, bn_args(0, wsp + ws.stem_raw, g.Hs, g.Ws, buf[0], g.h[0], g.w[0], ldb[0], false))));
  int k = 1;
  for (int b = 0; b < kBlocks; ++b) {
    const int M = g.M[b], h = g.h[b], w = g.w[b], ld = ldb[b];
    int C = g.Cin0[b];
    for (int j = 0; j < kLevels; ++j, k += 2) {
      bf16* y1 = B16(wsp, ws.y1[k]); bf16* a1 = B16(wsp, ws.a1[k]); bf16* y2 = B16(wsp, ws.y2[k + 1]);
      double* s1 = training ? stats + (size_t)k * 2 * kMaxBnC : nullptr;
      double* s3 = training ? stats + (size_t)(k + 1) * 2 * kMaxBnC : nullptr;
      if (fold) {
        WAP_TRY(tc_gemm_bf16(st, buf[b], ld, B16(wsp, ws.wt_fwd[k]), pad8(C), M, kBottle, C,
                             TcEpi{reinterpret_cast<float*>(a1), kBottle, 0, nullptr, nullptr, 0, 1, fsc(k), fsc(k) + kMaxBnC}, kProfConv1Fwd));
        const TcEpi e3{reinterpret_cast<float*>(buf[b] + C), ld, 0, nullptr, nullptr, 0, 1, fsc(k + 1), fsc(k + 1) + kMaxBnC};
        if ((rowtile_blocks >> b) & 1) WAP_TRY(tc_conv3x3_rowtile_bf16(st, a1, kBottle, B, h, w, B16(wsp, ws.wt_fwd[k + 1]), e3, kProfConv3Fwd));
        else WAP_TRY(tc_conv3x3_bf16(st, a1, kBottle, B, h, w, kBottle, B16(wsp, ws.wt_fwd[k + 1]), 32, kGrowth, e3, kProfConv3Fwd));
        C += kGrowth;
        continue;
      }
      WAP_TRY(tc_gemm_bf16(st, buf[b], ld, B16(wsp, ws.wt_fwd[k]), pad8(C), M, kBottle, C, TcEpi{reinterpret_cast<float*>(y1), kBottle, 0, s1, nullptr, 0, 1}, kProfConv1Fwd));
      WAP_TRY((run_bn_act<bf16, bf16>(st, 0, bn_args(k, y1, h, w, a1, h, w, kBottle, true))));
      if ((rowtile_blocks >> b) & 1)
        WAP_TRY(tc_conv3x3_rowtile_bf16(st, a1, kBottle, B, h, w, B16(wsp, ws.wt_fwd[k + 1]), TcEpi{reinterpret_cast<float*>(y2), kGrowth, 0, s3, nullptr, 0, 1}, kProfConv3Fwd));
      else
        WAP_TRY(tc_conv3x3_bf16(st, a1, kBottle, B, h, w, kBottle, B16(wsp, ws.wt_fwd[k + 1]), 32, kGrowth, TcEpi{reinterpret_cast<float*>(y2), kGrowth, 0, s3, nullptr, 0, 1}, kProfConv3Fwd));
      WAP_TRY((run_bn_act<bf16, bf16>(st, 0, bn_args(k + 1, y2, h, w, buf[b] + C, h, w, ld, true))));
      C += kGrowth;
    }
    if (b < kBlocks - 1) {
      const ConvLayer& ct = L.conv[k];
      bf16* yt = B16(wsp, ws.ytr[b]);
      double* sk = training ? stats + (size_t)k * 2 * kMaxBnC : nullptr;
      WAP_TRY(tc_gemm_bf16(st, buf[b], ld, B16(wsp, ws.wt_fwd[k]), pad8(C), M, ct.cout, C, TcEpi{reinterpret_cast<float*>(yt), ct.cout, 0, sk, nullptr, 0, 1}, kProfTransFwd));
      WAP_TRY((run_bn_act<bf16, bf16>(st, 2, bn_args(k, yt, h, w, buf[b + 1], g.h[b + 1], g.w[b + 1], ldb[b + 1], true))));
      ++k;
    }
  }
  {   // annotation [M2, 684] fp32  <-  block-2 buffer [M2, 688] bf16
    const size_t rows = (size_t)g.M[2];
    WAP_CHECK(launch_pdl(bf16_rows_to_f32_kernel, (unsigned)std::min<size_t>(cdivz(rows * (kD / 4), 256), 148 * 8), 256, 0, st,
                         (const bf16*)buf[2], kLd2, annotation, kD, rows, kD / 4));
    WAP_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------ backward
// d_annotation [B,h2,w2,684] fp32 is the gradient buffer of block 2 (clobbered); grads (flat fp32) is accumulated (+=).
int encoder_backward_bf16(cudaStream_t st, const float* params, float* grads, const float* x, int B, int H, int W,
                          int dropout, uint32_t seed, const float* annotation, float* d_annotation, float* wsp) {
  (void)annotation;
  const ParamLayout& L = layout();
  const EncGeom g = make_geom(B, H, W);
  const EncWorkspace ws = make_enc_workspace(g, true);
  float* saved = wsp + ws.saved;
  double* red2[2] = {reinterpret_cast<double*>(wsp + ws.red), reinterpret_cast<double*>(wsp + ws.red) + 2 * kMaxBnC};
  int bn_no = 0;
  WAP_CHECK(cudaMemsetAsync(red2[0], 0, sizeof(double) * 4 * kMaxBnC, st));
  const bf16* buf[kBlocks] = {B16(wsp, ws.buf[0]), B16(wsp, ws.buf[1]), B16(wsp, ws.buf[2])};
  const int ldb[kBlocks] = {g.Ctot[0], g.Ctot[1], kLd2};
  float* G[kBlocks] = {wsp + ws.gbuf[0], wsp + ws.gbuf[1], d_annotation};
  constexpr int kDy1Ld = 2 * kBottle;
  // dy1 [M,192] bf16: odd level -> columns 0..95, even level -> 96..191; dy2 ping-pong [M,32] bf16
  bf16* dy1base = B16(wsp, ws.dy1);
  bf16* dy1v[2] = {dy1base + kBottle, dy1base};
  bf16* dy2v[2] = {B16(wsp, ws.dy2), B16(wsp, ws.dy2b)};
  float* da1 = wsp + ws.da1; bf16* dytr = B16(wsp, ws.dytr);
  SideStreams* ss = overlap_enabled() ? side_streams() : nullptr;
  cudaStream_t sw = ss != nullptr ? ss->s[0] : st;
  bool pend[5] = {false, false, false, false, false};
  auto fork = [&]() -> int {
    if (ss != nullptr) { WAP_CHECK(cudaEventRecord(ss->fork[0], st)); WAP_CHECK(cudaStreamWaitEvent(sw, ss->fork[0], 0)); }
    return 0;
  };
  auto mark = [&](int e) -> int { if (ss != nullptr) { WAP_CHECK(cudaEventRecord(ss->done[e], sw)); pend[e] = true; } return 0; };
  auto join = [&](int e) -> int { if (ss != nullptr && pend[e]) { WAP_CHECK(cudaStreamWaitEvent(st, ss->done[e], 0)); pend[e] = false; } return 0; };
  for (int i = 0; i < 2; ++i) WAP_CHECK(cudaMemsetAsync(dy2v[i], 0, sizeof(bf16) * (size_t)g.M[0] * kDy2LdB, st));
  auto bwd_args = [&](int k, const void* raw, int Hi, int Wi, const float* dout, int Ho, int Wo, int ld_dout, void* gbuf, int ld_g, bool drop) {
    const ConvLayer& c = L.conv[k];
    BnBwdArgs a;
    a.raw = raw; a.B = B; a.Hi = Hi; a.Wi = Wi; a.C = c.cout;
    a.dout = dout; a.Ho = Ho; a.Wo = Wo; a.ld_dout = ld_dout;
    a.saved = saved + (size_t)k * 2 * kMaxBnC; a.gamma = params + c.gamma; a.beta = params + c.beta;
    a.g = gbuf; a.ld_g = ld_g; a.red = red2[bn_no & 1]; a.red_next = red2[(bn_no + 1) & 1]; a.bar = nullptr; ++bn_no; a.dgamma = grads + c.gamma; a.dbeta = grads + c.beta;
    a.dropout = (dropout && drop) ? 1 : 0; a.seed = seed; a.seed_src = seed_src(); a.site = (uint32_t)k;
    a.round_tf32 = 0;
    return a;
  };
  int k = kNumConv - 1;
  for (int b = kBlocks - 1; b >= 0; --b) {
    const int M = g.M[b], h = g.h[b], w = g.w[b], ld = ldb[b], ldg = g.Ctot[b];
    if (b < kBlocks - 1) {   // transition b -> b+1 (conv index k)
      const ConvLayer& ct = L.conv[k];
      const int C = g.Ctot[b], ldt = pad8(ct.cout);
      WAP_TRY(join(4));
      WAP_TRY((run_bn_bwd<bf16, bf16>(st, 2, bwd_args(k, B16(wsp, ws.ytr[b]), h, w, G[b + 1], g.h[b + 1], g.w[b + 1], g.Ctot[b + 1], dytr, ldt, true))));
      WAP_TRY(fork());
      for (int n0 = 0; n0 < ct.cout; n0 += 128)      // wgrad  dW[C, C/2] += buf^T * dy   (column chunks of <= 128)
        WAP_TRY(tc_wgrad_1x1_bf16(sw, buf[b], ld, dytr + n0, ldt, M, C, std::min(128, ct.cout - n0), grads + ct.w + n0, ct.cout, kProfTransBwd));
      WAP_TRY(mark(4));
      // dgrad  G[b] = dy * W^T   (initialises the whole gradient buffer of block b)
      WAP_TRY(tc_gemm_bf16(st, dytr, ldt, B16(wsp, ws.wt_bwd[k]), ldt, M, C, ct.cout, TcEpi{G[b], ldg, 0, nullptr, nullptr, 0, 0}, kProfTransBwd));
      --k;
    }
    int C = g.Ctot[b] - kGrowth;
    for (int j = kLevels - 1; j >= 0; --j, k -= 2, C -= kGrowth) {
      const int k3 = k, k1 = k - 1;
      const ConvLayer& c3 = L.conv[k3];
      const ConvLayer& c1 = L.conv[k1];
      const bf16* y1 = B16(wsp, ws.y1[k1]); const bf16* a1 = B16(wsp, ws.a1[k1]); const bf16* y2 = B16(wsp, ws.y2[k3]);
      const int par = j & 1;
      bf16* dy1 = dy1v[par]; bf16* dy2 = dy2v[par];
      WAP_TRY(join(par));
      WAP_TRY((run_bn_bwd<bf16, bf16>(st, 0, bwd_args(k3, y2, h, w, G[b] + C, h, w, ldg, dy2, kDy2LdB, true))));
      WAP_TRY(fork());
      WAP_TRY(tc_wgrad_3x3_bf16(sw, a1, kBottle, dy2, kDy2LdB, B, h, w, kBottle, kGrowth, grads + c3.w, kGrowth, kProfConv3Wgrad));
      WAP_TRY(mark(par));
      // dgrad 3x3: d a1[M,96] fp32 = conv3x3(dy2, mirrored weights)
      WAP_TRY(tc_conv3x3_bf16(st, dy2, kDy2LdB, B, h, w, kDy2LdB, B16(wsp, ws.wt_bwd[k3]), kBottle, kBottle, TcEpi{da1, kBottle, 0, nullptr, nullptr, 0, 0}, kProfConv3Dgrad));
      WAP_TRY(join(2 + par));
      WAP_TRY((run_bn_bwd<bf16, bf16>(st, 0, bwd_args(k1, y1, h, w, da1, h, w, kBottle, dy1, kDy1Ld, true))));
      WAP_TRY(fork());
      WAP_TRY(tc_wgrad_1x1_bf16(sw, buf[b], ld, dy1, kDy1Ld, M, C, kBottle, grads + c1.w, kBottle, kProfConv1Wgrad));
      WAP_TRY(mark(2 + par));
      // dgrad 1x1, two levels per pass over the block gradient buffer (see encoder.cu): pair buffer [C_odd][192] bf16
      const bf16* pair = B16(wsp, ws.wt_bwd[(j & 1) ? k1 : k1 + 2]);       // fp32 offset of the odd level = pair base
      if (j & 1) {
        WAP_TRY(tc_gemm_bf16(st, dy1, kDy1Ld, pair + (size_t)(C - kGrowth) * kDy1Ld, kDy1Ld, M, kGrowth, kBottle,
                             TcEpi{G[b] + (C - kGrowth), ldg, 1, nullptr, nullptr, 0, 0}, kProfConv1Dgrad));
      } else {
        WAP_TRY(tc_gemm_bf16(st, dy1base, kDy1Ld, pair, kDy1Ld, M, C, 2 * kBottle, TcEpi{G[b], ldg, 1, nullptr, nullptr, 0, 0}, kProfConv1Dgrad));
      }
    }
  }
  {   // stem: maxpool + ReLU + BN backward (fp32), then the 7x7 weight gradient
    float* gstem = wsp + ws.gstem;
    WAP_TRY((run_bn_bwd<float, float>(st, 1, bwd_args(0, wsp + ws.stem_raw, g.Hs, g.Ws, G[0], g.h[0], g.w[0], g.Ctot[0], gstem, kStemC, false))));
    WAP_TRY(run_stem_wgrad(st, x, gstem, grads + L.conv[0].w, B, H, W, g));
  }
  for (int e = 0; e < 5; ++e) WAP_TRY(join(e));
  return 0;
}

}  // namespace wap
