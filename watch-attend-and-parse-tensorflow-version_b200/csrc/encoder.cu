// DenseNet "Watcher" encoder: forward (training / inference) and backward.
// Reference semantics: model-single-GPU.py:73-132 (Watcher_train.dense_net) with TF-1.6 'SAME'
// padding, fused batch-norm (batch statistics incl. padded pixels, Bessel-corrected moving
// variance), ReLU, dropout(0.2) and 2x2 SAME max/avg pooling.
//
// Data layout: NHWC fp32.  Each dense block owns ONE buffer [B*h*w, Ctot] (Ctot = 432/600/684);
// level j reads channels [0,C_j) in place and its 24 new channels are written at [C_j, C_j+24):
// the reference's tf.concat (S:111) costs nothing here.
#include <algorithm>

#include "gemm.cuh"
#include "model.h"
#include "tc_gemm.cuh"
#include "bn_kernels.cuh"

namespace wap {

using Cfg128x32 = TileCfg<128, 32, 1, 1>;

// =====================================================================================
// stem: 7x7 stride-2 SAME conv, Cin = 1, Cout = 48 (S:78-79) + per-channel batch statistics
// =====================================================================================
constexpr int kStemTile = 16, kStemPatch = kStemTile * 2 + 5;   // 37

__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                 double* __restrict__ stats, int H, int W, int Hs, int Ws, int pt, int pl) {
  pdl_sync();
  __shared__ float patch[kStemPatch][kStemPatch + 1];
  __shared__ __align__(16) float ws[49 * kStemC];
  __shared__ float s_stat[2 * kStemC];
  const int b = blockIdx.z, oy0 = blockIdx.y * kStemTile, ox0 = blockIdx.x * kStemTile;
  const int tid = threadIdx.x, tx = tid % kStemTile, ty = tid / kStemTile;
  for (int i = tid; i < 49 * kStemC; i += 256) ws[i] = __ldg(w + i);
  if (tid < 2 * kStemC) s_stat[tid] = 0.f;
  const int iy0 = oy0 * 2 - pt, ix0 = ox0 * 2 - pl;
  for (int i = tid; i < kStemPatch * kStemPatch; i += 256) {
    const int r = i / kStemPatch, c = i % kStemPatch;
    const int iy = iy0 + r, ix = ix0 + c;
    float v = 0.f;
    if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) v = __ldg(x + ((size_t)b * H + iy) * W + ix);
    patch[r][c] = v;
  }
  __syncthreads();
  float acc[kStemC];
#pragma unroll
  for (int c = 0; c < kStemC; ++c) acc[c] = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky)
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const float v = patch[ty * 2 + ky][tx * 2 + kx];
      const float4* wr = reinterpret_cast<const float4*>(&ws[(ky * 7 + kx) * kStemC]);
#pragma unroll
      for (int q = 0; q < kStemC / 4; ++q) {
        const float4 ww = wr[q];
        acc[q * 4 + 0] = fmaf(v, ww.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(v, ww.y, acc[q * 4 + 1]);
        acc[q * 4 + 2] = fmaf(v, ww.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(v, ww.w, acc[q * 4 + 3]);
      }
    }
  const int oy = oy0 + ty, ox = ox0 + tx;
  const bool valid = oy < Hs && ox < Ws;
  if (valid) {
    float4* o = reinterpret_cast<float4*>(out + (((size_t)b * Hs + oy) * Ws + ox) * kStemC);
#pragma unroll
    for (int q = 0; q < kStemC / 4; ++q) o[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
  }
  if (stats != nullptr) {
#pragma unroll
    for (int c = 0; c < kStemC; ++c) {
      const float v = valid ? acc[c] : 0.f;
      const float s = warp_sum(v), q = warp_sum(v * v);
      if ((tid & 31) == 0) { atomicAdd(&s_stat[c], s); atomicAdd(&s_stat[kStemC + c], q); }
    }
    __syncthreads();
    if (tid < 2 * kStemC) atomicAdd(stats + tid, (double)s_stat[tid]);
  }
}

// dW[49][48] += sum over output pixels of x_patch[tap] * dy[pix][c]
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                  int H, int W, int Hs, int Ws, int pt, int pl) {
  pdl_sync();
  constexpr int kHalf = kStemTile * kStemTile / 2;         // 128 output pixels per pass
  constexpr int kOutPerThread = (49 * kStemC + 255) / 256;  // 10
  __shared__ float patch[kStemPatch][kStemPatch + 1];
  __shared__ float dys[kHalf][kStemC + 1];
  const int b = blockIdx.z, oy0 = blockIdx.y * kStemTile, ox0 = blockIdx.x * kStemTile;
  const int tid = threadIdx.x;
  const int iy0 = oy0 * 2 - pt, ix0 = ox0 * 2 - pl;
  for (int i = tid; i < kStemPatch * kStemPatch; i += 256) {
    const int r = i / kStemPatch, c = i % kStemPatch;
    const int iy = iy0 + r, ix = ix0 + c;
    float v = 0.f;
    if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) v = __ldg(x + ((size_t)b * H + iy) * W + ix);
    patch[r][c] = v;
  }
  float s[kOutPerThread];
#pragma unroll
  for (int i = 0; i < kOutPerThread; ++i) s[i] = 0.f;
  for (int half = 0; half < 2; ++half) {
    __syncthreads();
    for (int i = tid; i < kHalf * kStemC; i += 256) {
      const int p = i / kStemC + half * kHalf, c = i % kStemC;
      const int oy = oy0 + p / kStemTile, ox = ox0 + p % kStemTile;
      float v = 0.f;
      if (oy < Hs && ox < Ws) v = __ldg(dy + (((size_t)b * Hs + oy) * Ws + ox) * kStemC + c);
      dys[p - half * kHalf][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kOutPerThread; ++i) {
      const int o = tid + i * 256;
      if (o < 49 * kStemC) {
        const int tap = o / kStemC, c = o % kStemC, ky = tap / 7, kx = tap % 7;
        float acc = s[i];
#pragma unroll 4
        for (int pp = 0; pp < kHalf; ++pp) {
          const int p = pp + half * kHalf;
          acc = fmaf(patch[(p / kStemTile) * 2 + ky][(p % kStemTile) * 2 + kx], dys[pp][c], acc);
        }
        s[i] = acc;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kOutPerThread; ++i) {
    const int o = tid + i * 256;
    if (o < 49 * kStemC) atomicAdd(dw + o, s[i]);
  }
}

// =====================================================================================
// tcgen05 path: TF32-rounded (RNA), K-major copies of the conv weights, rebuilt from the live fp32
// weights at every call (one launch for all 98 layers).
//   fwd copy : [Cout(pad 32)][K]   with k = (tap*Cin + ci)          (B operand of the forward GEMM)
//   bwd copy : 1x1 / transition: the HWIO matrix itself [Cin][Cout] (B operand of dgrad, K = Cout)
//              3x3: [Cin=96][tap'*32 + co], tap' mirrored, co >= 24 zero  (dy2 rows are 32 floats)
// =====================================================================================
struct WPrepLayer { int kind, cin, cout, bld, block; long long w, f, b, fl, bl; };
__constant__ WPrepLayer c_wprep[kNumConv];

// rowtile_blocks: bit b set = the 3x3 convs of dense block b use the row-tile weight layout (image rows fit one 128-pixel tile)
__global__ void __launch_bounds__(256) prep_conv_weights_kernel(const float* __restrict__ params, float* __restrict__ wsp, int with_bwd, int rowtile_blocks) {
  pdl_sync();
  const WPrepLayer L = c_wprep[blockIdx.y];
  if (L.kind == kStem) return;
  const float* w = params + L.w;
  float* f = wsp + L.f; float* fl = wsp + L.fl;
  float* b = wsp + L.b; float* bl = wsp + L.bl;
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  auto split = [](float v, float* hi, float* lo) { const float h = tf32_rna(v); *hi = h; *lo = tf32_rna(v - h); };
  if (L.kind == kC3) {
    const int K = 9 * L.cin;                                  // 864
    if ((rowtile_blocks >> L.block) & 1) {                    // [96][288]: row 24 j + co, column 96 dy + ci  <-  w[dy][j][ci][co]
      const int K3 = 3 * L.cin;
      for (int i = t0; i < 96 * K3; i += stride) {
        const int n = i / K3, k = i - n * K3, dy = k / L.cin, ci = k - dy * L.cin, j = n / L.cout, co = n - j * L.cout;
        split(n < 3 * L.cout ? __ldg(w + ((size_t)(dy * 3 + j) * L.cin + ci) * L.cout + co) : 0.f, f + i, fl + i);
      }
    } else
    for (int i = t0; i < 32 * K; i += stride) {
      const int n = i / K, k = i - n * K;
      split(n < L.cout ? __ldg(w + (size_t)k * L.cout + n) : 0.f, f + i, fl + i);
    }
    if (with_bwd) {
      const int Kb = 9 * kDy2Ld;                              // 288
      for (int i = t0; i < L.cin * Kb; i += stride) {
        const int ci = i / Kb, k = i - ci * Kb, tap = k / kDy2Ld, co = k - tap * kDy2Ld;
        split(co < L.cout ? __ldg(w + ((size_t)(8 - tap) * L.cin + ci) * L.cout + co) : 0.f, b + i, bl + i);
      }
    }
  } else {                                                    // 1x1: w[cin][cout]
    const int n_el = L.cin * L.cout;
    for (int i = t0; i < n_el; i += stride) {
      const int n = i / L.cin, k = i - n * L.cin;
      split(__ldg(w + (size_t)k * L.cout + n), f + i, fl + i);
      if (with_bwd) { const int r = i / L.cout, cc = i - r * L.cout; split(__ldg(w + i), b + (size_t)r * L.bld + cc, bl + (size_t)r * L.bld + cc); }
    }
  }
}

// Inference: batch norm with moving statistics is an affine map per channel, folded into the conv epilogue of the tensor-core paths:
// scale = gamma / sqrt(moving_var + eps), shift = beta - moving_mean * scale (same expressions and rounding as bn_prologue, so the fused
// result equals conv -> bn_act bit for bit).  One launch for all 99 layers into the (unused in inference) `saved` region: [k][2][kMaxBnC].
struct BnFoldLayer { int cout; long long gamma, beta, mmean, mvar; };
__constant__ BnFoldLayer c_bnfold[kNumConv];
__global__ void __launch_bounds__(128) bn_fold_kernel(const float* __restrict__ params, const float* __restrict__ moving, float* __restrict__ out) {
  pdl_sync();
  const BnFoldLayer L = c_bnfold[blockIdx.x];
  float* o = out + (size_t)blockIdx.x * 2 * kMaxBnC;
  for (int c = threadIdx.x; c < L.cout; c += blockDim.x) {
    const double mean = moving[L.mmean + c], var = moving[L.mvar + c];
    const float invstd = (float)(1.0 / sqrt(var + (double)kBnEps));
    const float sc = __ldg(params + L.gamma + c) * invstd;
    o[c] = sc;
    o[kMaxBnC + c] = __ldg(params + L.beta + c) - (float)mean * sc;
  }
}

int g_dbg_layer = -1; float* g_dbg_buf = nullptr;   // debug capture of backward intermediates of one dense level (1x1 index)
int g_tc_debug_mask = 0x3FF;              // debug: bit0 transition dgrad, bit1 3x3 dgrad, bit2 1x1 dgrad, bit3 1x1 fwd, bit4 3x3 fwd, bit5 transition fwd, bit8 row-tile 3x3, bit9 inference BN folded into the conv epilogues, bit10 (off by default: measured slower) BN-backward reduce of the 1x1 output fused into the 3x3 dgrad epilogue, bit11 single-kernel BN backward (grid barrier) for small maps
void set_conv_precision(int mode) { cur()->precision = mode; }
int get_conv_precision() { return cur()->precision; }

// The weight-copy region sits at the START of every encoder workspace (make_enc_workspace), so the layer table is the same
// for every geometry: it is uploaded ONCE (synchronously, outside any stream capture - wap_create / the first eager call)
// and never touched again.  (Round 1 re-uploaded a geometry-dependent host table at every call: a captured graph re-read
// that table at replay time, after a different geometry had overwritten it.)
int upload_wprep_table() {
  static bool uploaded = false;
  if (uploaded) return 0;
  const ParamLayout& L = layout();
  const EncWorkspace ws = make_enc_workspace(make_geom(1, 16, 16), true);
  WPrepLayer h[kNumConv];
  for (int i = 0; i < kNumConv; ++i) {
    h[i].kind = L.conv[i].kind; h[i].cin = L.conv[i].cin; h[i].cout = L.conv[i].cout; h[i].bld = ws.wt_bwd_ld[i]; h[i].block = L.conv[i].block;
    h[i].w = (long long)L.conv[i].w; h[i].f = (long long)ws.wt_fwd[i]; h[i].b = (long long)ws.wt_bwd[i];
    h[i].fl = (long long)ws.wt_fwd_lo[i]; h[i].bl = (long long)ws.wt_bwd_lo[i];
  }
  WAP_CHECK(cudaMemcpyToSymbol(c_wprep, h, sizeof(h), 0, cudaMemcpyHostToDevice));
  BnFoldLayer f[kNumConv];
  for (int i = 0; i < kNumConv; ++i) {
    f[i].cout = L.conv[i].cout; f[i].gamma = (long long)L.conv[i].gamma; f[i].beta = (long long)L.conv[i].beta;
    f[i].mmean = (long long)L.conv[i].mmean; f[i].mvar = (long long)L.conv[i].mvar;
  }
  WAP_CHECK(cudaMemcpyToSymbol(c_bnfold, f, sizeof(f), 0, cudaMemcpyHostToDevice));
  uploaded = true;
  return 0;
}
int bn_fold(cudaStream_t st, const float* params, const float* moving, float* out) {
  WAP_TRY(upload_wprep_table());
  WAP_CHECK(launch_pdl(bn_fold_kernel, kNumConv, 128, 0, st, params, moving, out));
  WAP_LAUNCH_CHECK();
  return 0;
}

static int prep_conv_weights(cudaStream_t st, const float* params, float* wsp, bool with_bwd, int rowtile_blocks) {
  WAP_TRY(upload_wprep_table());
  WAP_CHECK(launch_pdl(prep_conv_weights_kernel, dim3(24, kNumConv), 256, 0, st, params, wsp, with_bwd ? 1 : 0, rowtile_blocks));
  WAP_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// test support: the ReLU gate (BN output > 0) of every element of conv layer k, evaluated with the SAME expression as the
// forward / backward kernels (fmaf(raw, gamma*invstd, beta - mean*gamma*invstd)) from the raw outputs and statistics a
// training-mode forward left in the workspace.  The gradient tests hand these gates to the fp64 oracle: an element within
// fp32 rounding of the kink is then gated identically on both sides and the comparison is well posed for every element.
// =====================================================================================
template <typename TRaw>
__global__ void __launch_bounds__(256) bn_gate_kernel(const TRaw* __restrict__ raw, const float* __restrict__ saved, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, int C, size_t n, uint8_t* __restrict__ out) {
  pdl_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float sc = __ldg(gamma + c) * saved[kMaxBnC + c];
    const float sh = __ldg(beta + c) - saved[c] * sc;
    out[i] = fmaf((float)raw[i], sc, sh) > 0.f ? 1 : 0;
  }
}
int debug_bn_gates(cudaStream_t st, const float* params, const float* wsp, int B, int H, int W, int k, uint8_t* out) {
  WAP_REQUIRE(k >= 0 && k < kNumConv, "wap_debug_bn_gates: layer %d out of range", k);
  const bool b16 = get_conv_precision() == 3 && k > 0;      // bf16 mode keeps the raw conv outputs of the dense blocks in bf16 (same offsets)
  const ParamLayout& L = layout();
  const EncGeom g = make_geom(B, H, W);
  const EncWorkspace ws = make_enc_workspace(g, true);
  const ConvLayer& c = L.conv[k];
  const float* raw; size_t npix;
  if (c.kind == kStem) { raw = wsp + ws.stem_raw; npix = (size_t)B * g.Hs * g.Ws; }
  else if (c.kind == kC1) { raw = wsp + ws.y1[k]; npix = g.M[c.block]; }
  else if (c.kind == kC3) { raw = wsp + ws.y2[k]; npix = g.M[c.block]; }
  else { raw = wsp + ws.ytr[c.block]; npix = g.M[c.block]; }
  const size_t n = npix * c.cout;
  const unsigned grid = (unsigned)std::min<size_t>(cdivz(n, 256), 148 * 8);
  const float* sv = wsp + ws.saved + (size_t)k * 2 * kMaxBnC;
  if (b16) WAP_CHECK(launch_pdl(bn_gate_kernel<__nv_bfloat16>, grid, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(raw), sv, params + c.gamma, params + c.beta, c.cout, n, out));
  else WAP_CHECK(launch_pdl(bn_gate_kernel<float>, grid, 256, 0, st, raw, sv, params + c.gamma, params + c.beta, c.cout, n, out));
  WAP_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// host orchestration
// =====================================================================================
static void stem_pad(int n_in, int n_out, int* before) {
  const int total = std::max((n_out - 1) * 2 + 7 - n_in, 0);
  *before = total / 2;
}
// stem conv (+statistics) / stem weight gradient launchers, shared with encoder_bf16.cu
int run_stem_conv(cudaStream_t st, const float* x, const float* w, float* out, double* stats, int B, int H, int W, const EncGeom& g) {
  int pt, pl; stem_pad(H, g.Hs, &pt); stem_pad(W, g.Ws, &pl);
  dim3 grid(cdiv(g.Ws, kStemTile), cdiv(g.Hs, kStemTile), B);
  WAP_CHECK(launch_pdl(stem_conv_kernel, grid, 256, 0, st, x, w, out, stats, H, W, g.Hs, g.Ws, pt, pl));
  WAP_LAUNCH_CHECK();
  return 0;
}
int run_stem_wgrad(cudaStream_t st, const float* x, const float* gstem, float* dw, int B, int H, int W, const EncGeom& g) {
  int pt, pl; stem_pad(H, g.Hs, &pt); stem_pad(W, g.Ws, &pl);
  dim3 grid(cdiv(g.Ws, kStemTile), cdiv(g.Hs, kStemTile), B);
  WAP_CHECK(launch_pdl(stem_wgrad_kernel, grid, 256, 0, st, x, gstem, dw, H, W, g.Hs, g.Ws, pt, pl));
  WAP_LAUNCH_CHECK();
  return 0;
}
int encoder_forward_bf16(cudaStream_t st, const float* params, float* moving, const float* x, int B, int H, int W,
                         int training, int dropout, uint32_t seed, float* annotation, float* wsp);
int encoder_backward_bf16(cudaStream_t st, const float* params, float* grads, const float* x, int B, int H, int W,
                          int dropout, uint32_t seed, const float* annotation, float* d_annotation, float* wsp);

int encoder_forward(cudaStream_t st, const float* params, float* moving, const float* x, int B, int H, int W,
                    int training, int dropout, uint32_t seed, float* annotation, float* wsp) {
  if (get_conv_precision() == 3) return encoder_forward_bf16(st, params, moving, x, B, H, W, training, dropout, seed, annotation, wsp);
  const ParamLayout& L = layout();
  const EncGeom g = make_geom(B, H, W);
  const EncWorkspace ws = make_enc_workspace(g, training != 0);
  double* stats = reinterpret_cast<double*>(wsp + ws.stats);
  float* saved = wsp + ws.saved;
  if (training) WAP_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * kNumConv * 2 * kMaxBnC, st));
  const int prec = get_conv_precision();
  const bool tc = prec != 0, x3 = prec == 2;
  int rowtile_blocks = 0;
  if (tc) {
    WAP_REQUIRE(tc_available(), "TF32 tensor-core mode requested but cuTensorMapEncodeTiled is unavailable");
    // 3x3 convs of a block whose rows fit one 128-pixel tile use the row-tile form (one third of the MMAs): bit b of the mask
    for (int b = 0; b < kBlocks; ++b) if (g.w[b] <= 128 && (g_tc_debug_mask & 256)) rowtile_blocks |= 1 << b;
    WAP_TRY(prep_conv_weights(st, params, wsp, training != 0, rowtile_blocks));
  }
  auto lo = [&](size_t off) -> const float* { return x3 ? wsp + off : nullptr; };
  float* buf[kBlocks] = {wsp + ws.buf[0], wsp + ws.buf[1], annotation != nullptr ? annotation : wsp + ws.buf[2]};
  // inference in the 3xTF32 mode: batch norm + ReLU of the dense levels folded into the conv epilogues (no bn_act launches, no y1 / y2 round trip)
  const bool fold = !training && x3 && (g_tc_debug_mask & 512);
  if (fold) WAP_TRY(bn_fold(st, params, moving, saved));
  auto fsc = [&](int kk) -> const float* { return saved + (size_t)kk * 2 * kMaxBnC; };
  auto bn_args = [&](int k, const float* raw, int Hi, int Wi, float* out, int Ho, int Wo, int ld_out, bool drop) {
    const ConvLayer& c = L.conv[k];
    BnActArgs a;
    a.raw = raw; a.B = B; a.Hi = Hi; a.Wi = Wi; a.C = c.cout;
    a.out = out; a.Ho = Ho; a.Wo = Wo; a.ld_out = ld_out;
    a.stats = training ? stats + (size_t)k * 2 * kMaxBnC : nullptr;
    a.mmean = moving + c.mmean; a.mvar = moving + c.mvar;
    a.gamma = params + c.gamma; a.beta = params + c.beta;
    a.saved = training ? saved + (size_t)k * 2 * kMaxBnC : nullptr;
    a.training = training; a.dropout = (training && dropout && drop) ? 1 : 0; a.seed = seed; a.seed_src = seed_src(); a.site = (uint32_t)k;
    a.round_tf32 = (tc && !x3) ? 1 : 0;
    return a;
  };
  // ---- stem (S:76-88)
  {
    int pt, pl; stem_pad(H, g.Hs, &pt); stem_pad(W, g.Ws, &pl);
    dim3 grid(cdiv(g.Ws, kStemTile), cdiv(g.Hs, kStemTile), B);
    WAP_CHECK(launch_pdl(stem_conv_kernel, grid, 256, 0, st, x, params + L.conv[0].w, wsp + ws.stem_raw, training ? stats : nullptr,
                                           H, W, g.Hs, g.Ws, pt, pl));
    WAP_LAUNCH_CHECK();
    WAP_TRY(run_bn_act(st, 1, bn_args(0, wsp + ws.stem_raw, g.Hs, g.Ws, buf[0], g.h[0], g.w[0], g.Ctot[0], false)));
  }
  // ---- dense blocks + transitions (S:91-130)
  int k = 1;
  for (int b = 0; b < kBlocks; ++b) {
    const int M = g.M[b], h = g.h[b], w = g.w[b], ld = g.Ctot[b];
    int C = g.Cin0[b];
    for (int j = 0; j < kLevels; ++j, k += 2) {
      const ConvLayer& c1 = L.conv[k];
      const ConvLayer& c3 = L.conv[k + 1];
      float* y1 = wsp + ws.y1[k]; float* a1 = wsp + ws.a1[k]; float* y2 = wsp + ws.y2[k + 1];
      {   // 1x1 bottleneck  C -> 96
        double* sk = training ? stats + (size_t)k * 2 * kMaxBnC : nullptr;
        if (fold) {
          WAP_TRY(tc_gemm(st, buf[b], ld, wsp + ws.wt_fwd[k], lo(ws.wt_fwd_lo[k]), C, M, kBottle, C,
                          TcEpi{a1, kBottle, 0, nullptr, nullptr, 0, 0, fsc(k), fsc(k) + kMaxBnC}, kProfConv1Fwd));
        } else if (tc) {
          WAP_TRY(tc_gemm(st, buf[b], ld, wsp + ws.wt_fwd[k], lo(ws.wt_fwd_lo[k]), C, M, kBottle, C, TcEpi{y1, kBottle, 0, sk, nullptr, 0}, kProfConv1Fwd));
        } else {
          ARowMajor A{buf[b], ld, M, C};
          BRowMajor Bw{params + c1.w, kBottle, C, kBottle};
          EpiStore e{y1, kBottle, nullptr, false, sk};
          WAP_TRY((launch_gemm<Cfg128x96>(st, A, Bw, e, M, kBottle, C, 1, kProfConv1Fwd)));
        }
        if (!fold) WAP_TRY(run_bn_act(st, 0, bn_args(k, y1, h, w, a1, h, w, kBottle, true)));
      }
      {   // 3x3 growth  96 -> 24, written in place at channel offset C
        double* sk = training ? stats + (size_t)(k + 1) * 2 * kMaxBnC : nullptr;
        if (fold) {
          const TcEpi e3{buf[b] + C, ld, 0, nullptr, nullptr, 0, 0, fsc(k + 1), fsc(k + 1) + kMaxBnC};     // activated output straight into the block buffer
          if ((rowtile_blocks >> b) & 1)
            WAP_TRY(tc_conv3x3_rowtile(st, a1, kBottle, B, h, w, wsp + ws.wt_fwd[k + 1], lo(ws.wt_fwd_lo[k + 1]), e3, kProfConv3Fwd));
          else
            WAP_TRY(tc_conv3x3(st, a1, kBottle, B, h, w, kBottle, wsp + ws.wt_fwd[k + 1], lo(ws.wt_fwd_lo[k + 1]), 32, kGrowth, e3, kProfConv3Fwd));
        } else if (tc) {
          if ((rowtile_blocks >> b) & 1)
            WAP_TRY(tc_conv3x3_rowtile(st, a1, kBottle, B, h, w, wsp + ws.wt_fwd[k + 1], lo(ws.wt_fwd_lo[k + 1]), TcEpi{y2, kGrowth, 0, sk, nullptr, 0}, kProfConv3Fwd));
          else
            WAP_TRY(tc_conv3x3(st, a1, kBottle, B, h, w, kBottle, wsp + ws.wt_fwd[k + 1], lo(ws.wt_fwd_lo[k + 1]), 32, kGrowth, TcEpi{y2, kGrowth, 0, sk, nullptr, 0}, kProfConv3Fwd));
        } else {
          AIm2col3 A{a1, kBottle, B, h, w, kBottle, false};
          BRowMajor Bw{params + c3.w, kGrowth, 9 * kBottle, kGrowth};
          EpiStore e{y2, kGrowth, nullptr, false, sk};
          WAP_TRY((launch_gemm<Cfg256x24>(st, A, Bw, e, M, kGrowth, 9 * kBottle, 1, kProfConv3Fwd)));
        }
        if (!fold) WAP_TRY(run_bn_act(st, 0, bn_args(k + 1, y2, h, w, buf[b] + C, h, w, ld, true)));
      }
      C += kGrowth;
    }
    if (b < kBlocks - 1) {   // transition: 1x1 C -> C/2, BN, ReLU, dropout, avgpool (S:116-130)
      const ConvLayer& ct = L.conv[k];
      float* yt = wsp + ws.ytr[b];
      double* sk = training ? stats + (size_t)k * 2 * kMaxBnC : nullptr;
      if (tc) {
        WAP_TRY(tc_gemm(st, buf[b], ld, wsp + ws.wt_fwd[k], lo(ws.wt_fwd_lo[k]), C, M, ct.cout, C, TcEpi{yt, ct.cout, 0, sk, nullptr, 0}, kProfTransFwd));
      } else {
        ARowMajor A{buf[b], ld, M, C};
        BRowMajor Bw{params + ct.w, ct.cout, C, ct.cout};
        EpiStore e{yt, ct.cout, nullptr, false, sk};
        WAP_TRY((launch_gemm<Cfg128x128>(st, A, Bw, e, M, ct.cout, C, 1, kProfTransFwd)));
      }
      WAP_TRY(run_bn_act(st, 2, bn_args(k, yt, h, w, buf[b + 1], g.h[b + 1], g.w[b + 1], g.Ctot[b + 1], true)));
      ++k;
    }
  }
  return 0;
}

// d_annotation [B,h2,w2,684] is the gradient wrt the encoder output; it is used as the gradient
// buffer of block 2 and is clobbered.  grads (flat, same layout as params) is accumulated (+=).
int encoder_backward(cudaStream_t st, const float* params, float* grads, const float* x, int B, int H, int W,
                     int dropout, uint32_t seed, const float* annotation, float* d_annotation, float* wsp) {
  if (get_conv_precision() == 3) return encoder_backward_bf16(st, params, grads, x, B, H, W, dropout, seed, annotation, d_annotation, wsp);
  const ParamLayout& L = layout();
  const EncGeom g = make_geom(B, H, W);
  const EncWorkspace ws = make_enc_workspace(g, true);
  float* saved = wsp + ws.saved;
  double* red2[2] = {reinterpret_cast<double*>(wsp + ws.red), reinterpret_cast<double*>(wsp + ws.red) + 2 * kMaxBnC};
  int bn_no = 0;
  WAP_CHECK(cudaMemsetAsync(red2[0], 0, sizeof(double) * 4 * kMaxBnC, st));
  unsigned* bnbar = reinterpret_cast<unsigned*>(wsp + ws.bnbar);
  const bool bn_fused = (g_tc_debug_mask & 2048) != 0;     // single-kernel BN backward for the small feature maps (bn_bwd_fused_kernel)
  if (bn_fused) WAP_CHECK(cudaMemsetAsync(bnbar, 0, sizeof(unsigned) * 256, st));
  const float* buf[kBlocks] = {wsp + ws.buf[0], wsp + ws.buf[1], annotation != nullptr ? annotation : wsp + ws.buf[2]};
  float* G[kBlocks] = {wsp + ws.gbuf[0], wsp + ws.gbuf[1], d_annotation};
  constexpr int kDy1Ld = 2 * kBottle;
  float* dy1v[2] = {wsp + ws.dy1b, wsp + ws.dy1}; float* dy2v[2] = {wsp + ws.dy2, wsp + ws.dy2b};     // dy1: odd level -> columns 0..95
  float* da1 = wsp + ws.da1; float* dytr = wsp + ws.dytr;
  // Weight gradients only feed `grads`, so they run on a side stream next to the dgrad -> BN-backward chain of the
  // following levels (which leaves most SMs idle in blocks 1 and 2: 64 and 16 row tiles).  dy1 / dy2 ping-pong between two
  // buffers; the main stream waits for the weight gradient that read a buffer before overwriting it two levels later.
  SideStreams* ss = overlap_enabled() ? side_streams() : nullptr;
  cudaStream_t sw = ss != nullptr ? ss->s[0] : st;
  bool pend[5] = {false, false, false, false, false};     // done[] events outstanding: wgrad3 (parity 0/1), wgrad1 (parity 0/1), transition
  auto fork = [&]() -> int {
    if (ss != nullptr) { WAP_CHECK(cudaEventRecord(ss->fork[0], st)); WAP_CHECK(cudaStreamWaitEvent(sw, ss->fork[0], 0)); }
    return 0;
  };
  auto mark = [&](int e) -> int { if (ss != nullptr) { WAP_CHECK(cudaEventRecord(ss->done[e], sw)); pend[e] = true; } return 0; };
  auto join = [&](int e) -> int { if (ss != nullptr && pend[e]) { WAP_CHECK(cudaStreamWaitEvent(st, ss->done[e], 0)); pend[e] = false; } return 0; };
  const int prec = get_conv_precision();
  const bool tc = prec != 0, x3 = prec == 2;
  auto lo = [&](size_t off) -> const float* { return x3 ? wsp + off : nullptr; };
  // dy2 rows are 32 floats (24 channels + zero pad) so that one row is one 128-byte TMA/UMMA row
  for (int i = 0; i < 2; ++i) WAP_CHECK(cudaMemsetAsync(dy2v[i], 0, sizeof(float) * (size_t)g.M[0] * kDy2Ld, st));
  auto bwd_args = [&](int k, const float* raw, int Hi, int Wi, const float* dout, int Ho, int Wo, int ld_dout, float* gbuf, int ld_g, bool drop) {
    const ConvLayer& c = L.conv[k];
    BnBwdArgs a;
    a.raw = raw; a.B = B; a.Hi = Hi; a.Wi = Wi; a.C = c.cout;
    a.dout = dout; a.Ho = Ho; a.Wo = Wo; a.ld_dout = ld_dout;
    a.saved = saved + (size_t)k * 2 * kMaxBnC; a.gamma = params + c.gamma; a.beta = params + c.beta;
    a.g = gbuf; a.ld_g = ld_g; a.red = red2[bn_no & 1]; a.red_next = red2[(bn_no + 1) & 1]; a.bar = bn_fused ? bnbar + bn_no : nullptr; ++bn_no; a.dgamma = grads + c.gamma; a.dbeta = grads + c.beta;
    a.dropout = (dropout && drop) ? 1 : 0; a.seed = seed; a.seed_src = seed_src(); a.site = (uint32_t)k;
    a.round_tf32 = (tc && !x3) ? 1 : 0;
    return a;
  };
  int k = kNumConv - 1;     // 98: last 3x3
  for (int b = kBlocks - 1; b >= 0; --b) {
    const int M = g.M[b], h = g.h[b], w = g.w[b], ld = g.Ctot[b];
    if (b < kBlocks - 1) {   // transition b -> b+1 (conv index k)
      const ConvLayer& ct = L.conv[k];
      const int C = ld;
      WAP_TRY(join(4));
      WAP_TRY(run_bn_bwd(st, 2, bwd_args(k, wsp + ws.ytr[b], h, w, G[b + 1], g.h[b + 1], g.w[b + 1], g.Ctot[b + 1], dytr, ct.cout, true)));
      WAP_TRY(fork());
      if (tc && (g_tc_debug_mask & 64)) {   // wgrad  dW[C, C/2] += buf^T * dy   (column chunks of <= 96)
        for (int n0 = 0; n0 < ct.cout; n0 += 96)
          WAP_TRY(tc_wgrad_1x1(sw, buf[b], ld, dytr + n0, ct.cout, M, C, std::min(96, ct.cout - n0), grads + ct.w + n0, ct.cout, x3, kProfTransBwd));
      } else {
        AColMajor A{buf[b], ld, C, M};
        BRowMajor Bd{dytr, ct.cout, M, ct.cout};
        EpiAtomic e{grads + ct.w, ct.cout};
        const int tiles = cdiv(C, 128) * cdiv(ct.cout, 128);
        WAP_TRY((launch_gemm<Cfg128x128>(sw, A, Bd, e, C, ct.cout, M, std::max(1, std::min(cdiv(M, 64), cdiv(592, tiles))), kProfTransBwd)));
      }
      WAP_TRY(mark(4));
      if (tc && (g_tc_debug_mask & 1)) {   // dgrad  G[b] = dy * W^T   (initialises the whole gradient buffer of block b)
        WAP_TRY(tc_gemm(st, dytr, ct.cout, wsp + ws.wt_bwd[k], lo(ws.wt_bwd_lo[k]), ct.cout, M, C, ct.cout, TcEpi{G[b], ld, 0, nullptr, nullptr, 0}, kProfTransBwd));
      } else {
        ARowMajor A{dytr, ct.cout, M, ct.cout};
        BColMajor Bw{params + ct.w, ct.cout, ct.cout, C};
        EpiStore e{G[b], ld, nullptr, false, nullptr};
        WAP_TRY((launch_gemm<Cfg128x128>(st, A, Bw, e, M, C, ct.cout, 1, kProfTransBwd)));
      }
      --k;
    }
    int C = ld - kGrowth;    // input channels of the last level's 1x1
    for (int j = kLevels - 1; j >= 0; --j, k -= 2, C -= kGrowth) {
      const int k3 = k, k1 = k - 1;
      const ConvLayer& c3 = L.conv[k3];
      const ConvLayer& c1 = L.conv[k1];
      const float* y1 = wsp + ws.y1[k1]; const float* a1 = wsp + ws.a1[k1]; const float* y2 = wsp + ws.y2[k3];
      const int par = j & 1;
      float* dy1 = dy1v[par]; float* dy2 = dy2v[par];
      // BN/ReLU/dropout backward of the 3x3 output
      WAP_TRY(join(par));
      WAP_TRY(run_bn_bwd(st, 0, bwd_args(k3, y2, h, w, G[b] + C, h, w, ld, dy2, kDy2Ld, true)));
      WAP_TRY(fork());
      if (tc && (g_tc_debug_mask & 64)) {   // wgrad 3x3: dW[864,24] += im2col(a1)^T * dy2
        WAP_TRY(tc_wgrad_3x3(sw, a1, kBottle, dy2, kDy2Ld, B, h, w, kBottle, kGrowth, grads + c3.w, kGrowth, x3, kProfConv3Wgrad));
      } else {
        AIm2col3T A{a1, kBottle, B, h, w, kBottle};
        BRowMajor Bd{dy2, kDy2Ld, M, kGrowth};
        EpiAtomic e{grads + c3.w, kGrowth};
        const int tiles = cdiv(9 * kBottle, 128);
        WAP_TRY((launch_gemm<Cfg128x32>(sw, A, Bd, e, 9 * kBottle, kGrowth, M, std::max(1, std::min(cdiv(M, 64), cdiv(592, tiles))), kProfConv3Wgrad)));
      }
      WAP_TRY(mark(par));
      const bool fuse_bnb = tc && (g_tc_debug_mask & 2) && (g_tc_debug_mask & 1024) && !(k1 == g_dbg_layer && g_dbg_buf != nullptr);
      if (fuse_bnb) {
        // dgrad 3x3 with the FIRST pass of the 1x1 output's BN backward in its epilogue: the accumulator (d a1) is gated by the ReLU / dropout of
        // that BN, stored as g into dy1 and column-summed (sum g, sum g*xhat) - d a1 is never written, bn_bwd_reduce is not launched
        WAP_TRY(join(2 + par));                 // dy1[par] is free once the 1x1 weight gradient of two levels ago has read it
        const BnBwdArgs ab = bwd_args(k1, y1, h, w, nullptr, h, w, kBottle, dy1, kDy1Ld, true);
        TcEpi e{dy1, kDy1Ld, 0, ab.red, nullptr, 0, 0, nullptr, nullptr, y1, kBottle, ab.saved, ab.gamma, ab.beta, ab.dropout, ab.seed, ab.site, ab.seed_src};
        WAP_TRY(tc_conv3x3(st, dy2, kDy2Ld, B, h, w, kDy2Ld, wsp + ws.wt_bwd[k3], lo(ws.wt_bwd_lo[k3]), kBottle, kBottle, e, kProfConv3Dgrad));
        WAP_TRY(run_bn_bwd_apply(st, ab));
      } else {
      if (tc && (g_tc_debug_mask & 2)) {   // dgrad 3x3: d a1[M,96]
        WAP_TRY(tc_conv3x3(st, dy2, kDy2Ld, B, h, w, kDy2Ld, wsp + ws.wt_bwd[k3], lo(ws.wt_bwd_lo[k3]), kBottle, kBottle, TcEpi{da1, kBottle, 0, nullptr, nullptr, 0}, kProfConv3Dgrad));
      } else {
        AIm2col3 A{dy2, kDy2Ld, B, h, w, kGrowth, false};
        BDgrad3 Bw{params + c3.w, kBottle, kGrowth};
        EpiStore e{da1, kBottle, nullptr, false, nullptr};
        WAP_TRY((launch_gemm<Cfg128x96>(st, A, Bw, e, M, kBottle, 9 * kGrowth, 1, kProfConv3Dgrad)));
      }
      if (k1 == g_dbg_layer && g_dbg_buf != nullptr) {   // [da1 (M*96) | dy2 (M*32)]
        WAP_CHECK(cudaMemcpyAsync(g_dbg_buf, da1, sizeof(float) * (size_t)M * kBottle, cudaMemcpyDeviceToDevice, st));
        WAP_CHECK(cudaMemcpyAsync(g_dbg_buf + (size_t)M * kBottle, dy2, sizeof(float) * (size_t)M * kDy2Ld, cudaMemcpyDeviceToDevice, st));
      }
      WAP_TRY(join(2 + par));
      WAP_TRY(run_bn_bwd(st, 0, bwd_args(k1, y1, h, w, da1, h, w, kBottle, dy1, kDy1Ld, true)));
      }
      WAP_TRY(fork());
      if (k1 == g_dbg_layer && g_dbg_buf != nullptr)     // [.. | dy1 (M*96)]
        WAP_CHECK(cudaMemcpy2DAsync(g_dbg_buf + (size_t)M * (kBottle + kDy2Ld), sizeof(float) * kBottle, dy1, sizeof(float) * kDy1Ld, sizeof(float) * kBottle, M,
                                    cudaMemcpyDeviceToDevice, st));
      if (tc && (g_tc_debug_mask & 64)) {   // wgrad 1x1: dW[C,96] += buf[:, :C]^T * dy1
        WAP_TRY(tc_wgrad_1x1(sw, buf[b], ld, dy1, kDy1Ld, M, C, kBottle, grads + c1.w, kBottle, x3, kProfConv1Wgrad));
      } else {
        AColMajor A{buf[b], ld, C, M};
        BRowMajor Bd{dy1, kDy1Ld, M, kBottle};
        EpiAtomic e{grads + c1.w, kBottle};
        const int tiles = cdiv(C, 128);
        WAP_TRY((launch_gemm<Cfg128x96>(sw, A, Bd, e, C, kBottle, M, std::max(1, std::min(cdiv(M, 64), cdiv(592, tiles))), kProfConv1Wgrad)));
      }
      WAP_TRY(mark(2 + par));
      if (tc && (g_tc_debug_mask & 4) && (g_tc_debug_mask & 128)) {
        // dgrad 1x1, two levels per pass over the block gradient buffer.  The read-modify-write of G[b][:, :C] is what this
        // K = 96 GEMM costs, so the odd level of a pair only updates the 24 channels the next level's backward reads right
        // away (its own input slice [C-24, C)) and the even level applies BOTH contributions to [0, C_even) in one K = 192
        // GEMM: A = [dy1_odd | dy1_even] (adjacent columns of one buffer), B = the pair weight buffer [C][192].
        const int bld = ws.wt_bwd_ld[k1];
        if (j & 1) {
          WAP_TRY(tc_gemm(st, dy1, kDy1Ld, wsp + ws.wt_bwd[k1] + (size_t)(C - kGrowth) * bld, x3 ? wsp + ws.wt_bwd_lo[k1] + (size_t)(C - kGrowth) * bld : nullptr,
                          bld, M, kGrowth, kBottle, TcEpi{G[b] + (C - kGrowth), ld, 1, nullptr, nullptr, 0}, kProfConv1Dgrad));
        } else {
          WAP_TRY(tc_gemm(st, dy1v[1], kDy1Ld, wsp + ws.wt_bwd[k1 + 2], lo(ws.wt_bwd_lo[k1 + 2]), bld, M, C, 2 * kBottle,
                          TcEpi{G[b], ld, 1, nullptr, nullptr, 0}, kProfConv1Dgrad));
        }
      } else if (tc && (g_tc_debug_mask & 4)) {   // dgrad 1x1: G[b][:, :C] += dy1 * W1^T
        WAP_TRY(tc_gemm(st, dy1, kDy1Ld, wsp + ws.wt_bwd[k1], lo(ws.wt_bwd_lo[k1]), ws.wt_bwd_ld[k1], M, C, kBottle, TcEpi{G[b], ld, 1, nullptr, nullptr, 0}, kProfConv1Dgrad));
      } else {
        ARowMajor A{dy1, kDy1Ld, M, kBottle};
        BColMajor Bw{params + c1.w, kBottle, kBottle, C};
        EpiStore e{G[b], ld, nullptr, true, nullptr};
        WAP_TRY((launch_gemm<Cfg128x128>(st, A, Bw, e, M, C, kBottle, 1, kProfConv1Dgrad)));
      }
    }
  }
  // ---- stem: maxpool + ReLU + BN backward, then the 7x7 weight gradient (no dgrad: input is the image)
  {
    float* gstem = wsp + ws.gstem;
    WAP_TRY(run_bn_bwd(st, 1, bwd_args(0, wsp + ws.stem_raw, g.Hs, g.Ws, G[0], g.h[0], g.w[0], g.Ctot[0], gstem, kStemC, false)));
    int pt, pl; stem_pad(H, g.Hs, &pt); stem_pad(W, g.Ws, &pl);
    dim3 grid(cdiv(g.Ws, kStemTile), cdiv(g.Hs, kStemTile), B);
    WAP_CHECK(launch_pdl(stem_wgrad_kernel, grid, 256, 0, st, x, gstem, grads + L.conv[0].w, H, W, g.Hs, g.Ws, pt, pl));
    WAP_LAUNCH_CHECK();
  }
  for (int e = 0; e < 5; ++e) WAP_TRY(join(e));      // every side-stream weight gradient rejoins the caller's stream
  return 0;
}

}  // namespace wap
