// Inference-side decoder: attention precompute (hoisted U_a*a, fused coverage filter, h0),
// WAP.get_word step (model-single-GPU.py:379-439), greedy decode loop (:475-483 intended semantics).
#include <algorithm>
#include <utility>

#include "decoder.cuh"

namespace wap {

// mean over positions (masked, S:345; unmasked, S:590)
__global__ void __launch_bounds__(256)
ann_mean_kernel(const float* __restrict__ ann, const float* __restrict__ a_mask, float* __restrict__ mean, int L) {
  pdl_sync();
  extern __shared__ float sm[];          // w[L] | red[64] | part[16][64]
  float* s_w = sm; float* red = sm + ((L + 3) & ~3); float* part = red + 64;
  const int b = blockIdx.y, tid = threadIdx.x;
  float sum = 0.f;
  for (int l = tid; l < L; l += 256) { const float m = a_mask != nullptr ? a_mask[(size_t)b * L + l] : 1.f; s_w[l] = m; sum += m; }
  sum = block_sum(sum, red);
  const int c0 = blockIdx.x * 64, cq = tid & 15, lg = tid >> 4, c = c0 + cq * 4;
  float4 acc = make4(0.f);
  if (c < kD) {
    const float* base = ann + (size_t)b * L * kD + c;
    for (int l = lg; l < L; l += 16) {
      const float4 v = ldg4(base + (size_t)l * kD);
      const float m = s_w[l];
      acc.x = fmaf(m, v.x, acc.x); acc.y = fmaf(m, v.y, acc.y); acc.z = fmaf(m, v.z, acc.z); acc.w = fmaf(m, v.w, acc.w);
    }
  }
  *reinterpret_cast<float4*>(&part[lg * 64 + cq * 4]) = acc;
  __syncthreads();
  if (tid < 64 && c0 + tid < kD) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) s += part[g * 64 + tid];
    mean[(size_t)b * kD + c0 + tid] = s / sum;
  }
}

__global__ void tanh_inplace_kernel(float* x, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = tanhf(x[i]);
}

// watch = ann*U_a + U_a_b ; qf = Q*U_f ; h0 = tanh(mean(ann)*Wa2h + ba2h).  mean_scratch: [B,684]
int attn_precompute(cudaStream_t st, const float* params, const float* ann, const float* a_mask, int B, int h, int w,
                    float* watch, float* h0, float* qf, float* mean_scratch) {
  const ParamLayout& P = layout();
  const int L = h * w;
  if (watch != nullptr)
    WAP_TRY(gemm_nn(st, ann, kD, params + P.U_a, kNA, watch, kNA, B * L, kNA, kD, params + P.U_a_b, false));
  if (qf != nullptr)
    WAP_TRY(gemm_nn(st, params + P.Q, kNA, params + P.U_f, kNA, qf, kNA, kCovTaps, kNA, kNA, nullptr, false));
  if (h0 != nullptr) {
    const size_t smem = sizeof(float) * (((L + 3) & ~3) + 64 + 16 * 64);
    WAP_REQUIRE(smem <= 48 * 1024, "annotation grid too large: L=%d", L);
    WAP_CHECK(launch_pdl(ann_mean_kernel, dim3(cdiv(kD, 64), B), 256, smem, st, ann, a_mask, mean_scratch, L));
    WAP_LAUNCH_CHECK();
    FillArgs f; f.rows = B; f.nseg = 1; f.seg[0] = {h0, kN, kN, params + P.ba2h};
    WAP_TRY(fill_rows(st, f));
    SkinnyArgs a; a.X = mean_scratch; a.ldx = kD; a.ids = nullptr; a.Bn = B; a.K = kD; a.nseg = 1;
    a.seg[0] = {params + P.Wa2h, kN, kN, h0, kN};
    WAP_TRY(skinny_gemm(st, a));
    WAP_CHECK(launch_pdl(tanh_inplace_kernel, cdiv(B * kN, 256), 256, 0, st, h0, B * kN));
    WAP_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------ head
__global__ void lg_init_kernel(float* LG, int Bn, const float* by, const float* bc, const float* bh) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Bn * kMw) { const int j = i % kMw; LG[i] = by[j] + bc[j] + bh[j]; }
}
__global__ void maxout_kernel(const float* __restrict__ LG, float* __restrict__ MO, int rows) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * (kMw / 2)) {
    const float2 v = *reinterpret_cast<const float2*>(LG + 2 * (size_t)i);
    MO[i] = fmaxf(v.x, v.y);
  }
}
// softmax over 111 logits (one warp per row) + first-index argmax
__global__ void __launch_bounds__(256)
softmax_argmax_kernel(const float* __restrict__ LO, float* __restrict__ probs, int* __restrict__ word, int rows) {
  pdl_sync();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* lo = LO + (size_t)r * kVocab;
  float v[4]; int idx[4];
  float mx = -INFINITY; int am = 0x7fffffff;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    idx[q] = q * 32 + lane;
    v[q] = idx[q] < kVocab ? lo[idx[q]] : -INFINITY;
    if (v[q] > mx) { mx = v[q]; am = idx[q]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
  }
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) { v[q] = idx[q] < kVocab ? expf(v[q] - mx) : 0.f; s += v[q]; }
  s = warp_sum(s);
  const float inv = 1.f / s;
  if (probs != nullptr)
#pragma unroll
    for (int q = 0; q < 4; ++q) if (idx[q] < kVocab) probs[(size_t)r * kVocab + idx[q]] = v[q] * inv;
  if (word != nullptr && lane == 0) word[r] = am;
}

// step workspace carve-up (floats)
struct StepWs {
  float *G1, *UH, *prop1, *pre_h, *SP, *G2, *PH2, *CH, *prop2, *COV, *E, *alpha, *ctx, *LG, *MO, *LO;
};
static StepWs carve_step(float* w, int Bn, int L, size_t* total) {
  StepWs s; size_t o = 0;
  auto take = [&](size_t n) { float* p = w + o; o += (n + 3) & ~(size_t)3; return p; };
  s.G1 = take((size_t)Bn * 3 * kN); s.UH = take((size_t)Bn * kN); s.prop1 = take((size_t)Bn * kN); s.pre_h = take((size_t)Bn * kN);
  s.SP = take((size_t)Bn * kNA); s.G2 = take((size_t)Bn * 2 * kN); s.PH2 = take((size_t)Bn * kN); s.CH = take((size_t)Bn * kN);
  s.prop2 = take((size_t)Bn * kN); s.COV = take((size_t)Bn * L * kNA); s.E = take((size_t)Bn * L); s.alpha = take((size_t)Bn * L);
  s.ctx = take((size_t)Bn * kD); s.LG = take((size_t)Bn * kMw); s.MO = take((size_t)Bn * kMw / 2); s.LO = take((size_t)Bn * 112);
  if (total != nullptr) *total = o;
  return s;
}
size_t step_ws_floats(int Bn, int L) { size_t t; carve_step(nullptr, Bn, L, &t); return t; }

// WAP.get_word: y[Bn] (id < 0 => zero embedding), h_prev, alpha_past -> probs, argmax word, h_new, alpha_past'
int decoder_word_step(cudaStream_t st, const float* params, const float* ann, const float* watch, const float* qf,
                      const float* a_mask, int shared, int Bn, int h, int w, const int* y, const float* h_prev,
                      const float* ap_in, float* probs, int* word, float* h_new, float* ap_out, float* wsp) {
  const ParamLayout& P = layout();
  const int L = h * w;
  StepWs s = carve_step(wsp, Bn, L, nullptr);
  {   // emb-dependent pre-activations: G1 = emb*[W_yz_yr | W_yh] + b ; LG = emb*Wy + (by+bc+bh)
    FillArgs f; f.rows = Bn; f.nseg = 3;
    f.seg[0] = {s.G1, 3 * kN, 2 * kN, params + P.b_yz_yr};
    f.seg[1] = {s.G1 + 2 * kN, 3 * kN, kN, params + P.b_yh};
    f.seg[2] = {s.LO, kVocab, kVocab, params + P.bo};
    WAP_TRY(fill_rows(st, f));
    WAP_CHECK(launch_pdl(lg_init_kernel, cdiv(Bn * kMw, 256), 256, 0, st, s.LG, Bn, params + P.by, params + P.bc, params + P.bh));
    WAP_LAUNCH_CHECK();
    SkinnyArgs a; a.X = params + P.embed; a.ldx = kMw; a.ids = y; a.Bn = Bn; a.K = kMw; a.nseg = 3;
    a.seg[0] = {params + P.W_yz_yr, 2 * kN, 2 * kN, s.G1, 3 * kN};
    a.seg[1] = {params + P.W_yh, kN, kN, s.G1 + 2 * kN, 3 * kN};
    a.seg[2] = {params + P.Wy, kMw, kMw, s.LG, kMw};
    WAP_TRY(skinny_gemm(st, a));
  }
  StepCtx c;
  c.Bn = Bn; c.h = h; c.w = w; c.L = L; c.params = params; c.ann = ann; c.watch = watch; c.shared = shared; c.qf = qf;
  c.a_mask = a_mask; c.y_mask = nullptr; c.h_prev = h_prev; c.ap_in = ap_in;
  c.G1 = s.G1; c.UH = s.UH; c.prop1 = s.prop1; c.pre_h = s.pre_h; c.SP = s.SP; c.G2 = s.G2; c.PH2 = s.PH2; c.CH = s.CH;
  c.prop2 = s.prop2; c.COV = s.COV; c.TANH = s.COV; c.E = s.E; c.alpha = s.alpha; c.ap_out = ap_out; c.ctx = s.ctx; c.h_new = h_new;
  c.extra_ctx = {params + P.Wc, kMw, kMw, s.LG, kMw}; c.has_extra_ctx = 1;
  c.tc_mode = 0; c.Ppat = nullptr; c.qfT_hi = nullptr; c.qfT_lo = nullptr;
  WAP_TRY(decoder_step(st, c));
  {   // logits: + h*Wh ; maxout ; Wo ; softmax (dropout is off at decode time, S:433 with training=False)
    SkinnyArgs a; a.X = h_new; a.ldx = kN; a.ids = nullptr; a.Bn = Bn; a.K = kN; a.nseg = 1;
    a.seg[0] = {params + P.Wh, kMw, kMw, s.LG, kMw};
    WAP_TRY(skinny_gemm(st, a));
    WAP_CHECK(launch_pdl(maxout_kernel, cdiv(Bn * kMw / 2, 256), 256, 0, st, s.LG, s.MO, Bn));
    WAP_LAUNCH_CHECK();
    SkinnyArgs o; o.X = s.MO; o.ldx = kMw / 2; o.ids = nullptr; o.Bn = Bn; o.K = kMw / 2; o.nseg = 1;
    o.seg[0] = {params + P.Wo, kVocab, kVocab, s.LO, kVocab};
    WAP_TRY(skinny_gemm(st, o));
    WAP_CHECK(launch_pdl(softmax_argmax_kernel, cdiv(Bn * 32, 256), 256, 0, st, s.LO, probs, word, Bn));
    WAP_LAUNCH_CHECK();
  }
  return 0;
}


// Parser.one_time_step (model-single-GPU.py:260-308) for a GIVEN embedding emb_y[Bn,256] (the scan body of get_ht_ctx):
// -> h_new, context, alpha, alpha_past'.  a_mask [Bn,L] / y_mask [Bn] may be null.
int decoder_one_time_step(cudaStream_t st, const float* params, const float* ann, const float* watch, const float* qf,
                          const float* a_mask, const float* y_mask, int Bn, int h, int w, const float* emb_y, const float* h_prev,
                          const float* ap_in, float* h_new, float* ctx, float* alpha, float* ap_out, float* wsp) {
  const ParamLayout& P = layout();
  const int L = h * w;
  StepWs s = carve_step(wsp, Bn, L, nullptr);
  {
    FillArgs f; f.rows = Bn; f.nseg = 2;
    f.seg[0] = {s.G1, 3 * kN, 2 * kN, params + P.b_yz_yr};
    f.seg[1] = {s.G1 + 2 * kN, 3 * kN, kN, params + P.b_yh};
    WAP_TRY(fill_rows(st, f));
    SkinnyArgs a; a.X = emb_y; a.ldx = kMw; a.ids = nullptr; a.Bn = Bn; a.K = kMw; a.nseg = 2;
    a.seg[0] = {params + P.W_yz_yr, 2 * kN, 2 * kN, s.G1, 3 * kN};
    a.seg[1] = {params + P.W_yh, kN, kN, s.G1 + 2 * kN, 3 * kN};
    WAP_TRY(skinny_gemm(st, a));
  }
  StepCtx c;
  c.Bn = Bn; c.h = h; c.w = w; c.L = L; c.params = params; c.ann = ann; c.watch = watch; c.shared = 0; c.qf = qf;
  c.a_mask = a_mask; c.y_mask = y_mask; c.h_prev = h_prev; c.ap_in = ap_in;
  c.G1 = s.G1; c.UH = s.UH; c.prop1 = s.prop1; c.pre_h = s.pre_h; c.SP = s.SP; c.G2 = s.G2; c.PH2 = s.PH2; c.CH = s.CH;
  c.prop2 = s.prop2; c.COV = s.COV; c.TANH = s.COV; c.E = s.E; c.alpha = alpha; c.ap_out = ap_out; c.ctx = ctx; c.h_new = h_new;
  c.has_extra_ctx = 0; c.tc_mode = 0; c.Ppat = nullptr; c.qfT_hi = nullptr; c.qfT_lo = nullptr;
  return decoder_step(st, c);
}

// Attender.get_context (model-single-GPU.py:166-206): annotation, h_{t-1}, alpha_past, a_mask -> context, alpha, alpha_past'
int attender_get_context(cudaStream_t st, const float* params, const float* ann, const float* watch, const float* qf,
                         const float* a_mask, int Bn, int h, int w, const float* h_t_1, const float* ap_in,
                         float* ctx, float* alpha, float* ap_out, float* wsp) {
  const ParamLayout& P = layout();
  const int L = h * w;
  StepWs s = carve_step(wsp, Bn, L, nullptr);
  FillArgs f; f.rows = Bn; f.nseg = 1; f.seg[0] = {s.SP, kNA, kNA, params + P.W_a_b};
  WAP_TRY(fill_rows(st, f));
  SkinnyArgs a; a.X = h_t_1; a.ldx = kN; a.ids = nullptr; a.Bn = Bn; a.K = kN; a.nseg = 1;
  a.seg[0] = {params + P.W_a, kNA, kNA, s.SP, kNA};
  WAP_TRY(skinny_gemm(st, a));
  WAP_TRY(attention_core(st, params, ann, watch, qf, a_mask, 0, Bn, h, w, s.SP, ap_in, s.COV, s.COV, s.E, alpha, ap_out, ctx));
  return 0;
}

// ------------------------------------------------------------------------------------ greedy loop
// tokens[b, t] receives the argmax word while sequence b is alive; lengths[b] = number of tokens
// written (including the terminating <eol>=0 if it was emitted).  state: y (next input id).
__global__ void greedy_update_kernel(const int* __restrict__ word, int* __restrict__ y, int* __restrict__ tokens,
                                     int* __restrict__ lengths, int* __restrict__ alive, int Bn, int t, int maxlen) {
  pdl_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bn) return;
  if (lengths[b] >= 0) return;                  // finished earlier (length fixed)
  const int wd = word[b];
  tokens[(size_t)b * maxlen + t] = wd;
  y[b] = wd;
  if (wd == 0 || t == maxlen - 1) { lengths[b] = t + 1; atomicSub(alive, 1); }
}
__global__ void greedy_init_kernel(int* y, int* lengths, int* alive, int Bn) {
  pdl_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < Bn) { y[b] = -1; lengths[b] = -1; }
  if (b == 0) *alive = Bn;
}

struct GreedyWs { float *stepws, *hA, *hB, *apA, *apB, *mean, *watch, *qf; int *y, *word, *alive; };
static GreedyWs carve_greedy(float* wsp, int Bn, int L, size_t* total) {
  GreedyWs g; size_t o = 0;
  auto take = [&](size_t n) { float* p = wsp + o; o += (n + 3) & ~(size_t)3; return p; };
  g.stepws = take(step_ws_floats(Bn, L));
  g.hA = take((size_t)Bn * kN); g.hB = take((size_t)Bn * kN);
  g.apA = take((size_t)Bn * L); g.apB = take((size_t)Bn * L);
  g.mean = take((size_t)Bn * kD);
  g.watch = take((size_t)Bn * L * kNA);
  g.qf = take((size_t)kCovTaps * kNA);
  g.y = reinterpret_cast<int*>(take(Bn)); g.word = reinterpret_cast<int*>(take(Bn));
  g.alive = reinterpret_cast<int*>(take(4));
  if (total != nullptr) *total = o;
  return g;
}
size_t greedy_ws_floats(int Bn, int L) { size_t t; carve_greedy(nullptr, Bn, L, &t); return t; }

// Greedy decode of Bn independent images.  Host-synchronous every `check_every` steps (early exit).
int greedy_decode(cudaStream_t st, const float* params, const float* ann, int Bn, int h, int w, int maxlen,
                  int check_every, int* tokens, int* lengths, float* wsp, int* steps_run) {
  const int L = h * w;
  GreedyWs g = carve_greedy(wsp, Bn, L, nullptr);
  float *stepws = g.stepws, *hA = g.hA, *hB = g.hB, *apA = g.apA, *apB = g.apB, *mean = g.mean, *watch = g.watch, *qf = g.qf;
  int *y = g.y, *word = g.word, *alive = g.alive;
  WAP_TRY(attn_precompute(st, params, ann, nullptr, Bn, h, w, watch, hA, qf, mean));   // hidden_state_0: unmasked (S:590)
  WAP_CHECK(cudaMemsetAsync(apA, 0, sizeof(float) * Bn * L, st));
  WAP_CHECK(launch_pdl(greedy_init_kernel, cdiv(Bn, 128), 128, 0, st, y, lengths, alive, Bn));
  WAP_LAUNCH_CHECK();
  int t = 0;
  for (; t < maxlen; ++t) {
    WAP_TRY(decoder_word_step(st, params, ann, watch, qf, nullptr, 0, Bn, h, w, y, hA, apA, nullptr, word, hB, apB, stepws));
    WAP_CHECK(launch_pdl(greedy_update_kernel, cdiv(Bn, 128), 128, 0, st, word, y, tokens, lengths, alive, Bn, t, maxlen));
    WAP_LAUNCH_CHECK();
    std::swap(hA, hB); std::swap(apA, apB);
    if (check_every > 0 && ((t + 1) % check_every == 0)) {
      int host_alive = 0;
      WAP_CHECK(cudaMemcpyAsync(&host_alive, alive, sizeof(int), cudaMemcpyDeviceToHost, st));
      WAP_CHECK(cudaStreamSynchronize(st));
      if (host_alive <= 0) { ++t; break; }
    }
  }
  if (steps_run != nullptr) *steps_run = t;
  return 0;
}

}  // namespace wap
