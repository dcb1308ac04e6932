// Decoder time-step kernels (forward).  See decoder.cuh for the data flow.
#include <algorithm>

#include "decoder.cuh"
#include "tc_gemm.cuh"

namespace wap {

// =====================================================================================
// skinny GEMM
// =====================================================================================
// GRU-1 (S:273-289) / GRU-2 (S:292-306) gate of one element; shared by the gate kernels and the X generators
__device__ __forceinline__ void gru1_eval(float gr, float gz, float eh, float uh, float h, float m, bool masked,
                                          float& r, float& z, float& p, float& ph) {
  r = sigmoidf_(gr); z = sigmoidf_(gz);
  p = tanhf(fmaf(uh, r, eh));
  ph = z * h + (1.f - z) * p;
  if (masked) ph = m * ph + (1.f - m) * h;
}
__device__ __forceinline__ void gru2_eval(float gr, float gz, float ph2, float ch, float preh, float m, bool masked,
                                          float& r, float& z, float& p, float& h) {
  r = sigmoidf_(gr); z = sigmoidf_(gz);
  p = tanhf(fmaf(ph2, r, ch));
  h = z * preh + (1.f - z) * p;
  if (masked) h = m * h + (1.f - m) * preh;
}
__global__ void __launch_bounds__(64) skinny_gemm_kernel(const SkinnyArgs a) {
  pdl_sync();
  __shared__ float xs[kSkRows][kSkKC];
  int s = 0;
  while (s + 1 < a.nseg && (int)blockIdx.x >= a.tile0[s + 1]) ++s;
  const SkinnySeg sg = a.seg[s];
  const bool first_tile = (int)blockIdx.x == a.tile0[s];
  const int n = ((int)blockIdx.x - a.tile0[s]) * kSkCols + threadIdx.x * 4;
  const int k0 = blockIdx.y * kSkKC;
  const int Ks = sg.kcount > 0 ? sg.kcount : a.K;
  if (k0 >= Ks) return;
  const bool vec = ((sg.ldw & 3) == 0) && (n + 3 < sg.N);
  float4 wv[kSkKC];
#pragma unroll
  for (int kk = 0; kk < kSkKC; ++kk) {
    const int k = k0 + kk;
    float4 v = make4(0.f);
    if (k < Ks && n < sg.N) {
      const float* q = sg.W + (size_t)k * sg.ldw + n;
      if (vec) v = ldg4(q);
      else {
        v.x = __ldg(q);
        if (n + 1 < sg.N) v.y = __ldg(q + 1);
        if (n + 2 < sg.N) v.z = __ldg(q + 2);
        if (n + 3 < sg.N) v.w = __ldg(q + 3);
      }
    }
    wv[kk] = v;
  }
  for (int b0 = 0; b0 < a.Bn; b0 += kSkRows) {
    __syncthreads();
    for (int i = threadIdx.x; i < kSkRows * kSkKC; i += 64) {
      const int bb = i / kSkKC, kk = i % kSkKC;
      const int b = b0 + bb, k = k0 + kk + sg.koff;       // k indexes X
      float v = 0.f;
      if (b < a.Bn && k0 + kk < Ks) {
        if (a.gen == 1 || a.gen == 2) {
          const bool masked = a.gmask != nullptr;
          const float m = masked ? a.gmask[b] : 1.f;
          const size_t e = (size_t)b * kN + k;
          float r, z, p;
          if (a.gen == 1) {
            const float* g = a.gG + (size_t)b * 3 * kN;
            gru1_eval(g[k], g[kN + k], g[2 * kN + k], a.gA[e], a.gB[e], m, masked, r, z, p, v);
          } else {
            const float* g = a.gG + (size_t)b * 2 * kN;
            gru2_eval(g[k], g[kN + k], a.gA[e], a.gB[e], a.gC[e], m, masked, r, z, p, v);
          }
          if (blockIdx.x == 0) {
            a.oRZ[(size_t)b * 2 * kN + k] = r; a.oRZ[(size_t)b * 2 * kN + kN + k] = z;
            a.oProp[e] = p; a.oX[e] = v;
          }
        } else if (a.gen == 3) {       // GRU-2 backward (S:292-306 differentiated)
          const int part = k / kN, j = k - part * kN;
          const size_t e = (size_t)b * kN + j;
          const float m = a.gmask != nullptr ? a.gmask[b] : 1.f;
          const float dh = a.gA[e] + a.gB[e];
          const float r = a.gG[(size_t)b * 2 * kN + j], z = a.gG[(size_t)b * 2 * kN + kN + j];
          const float p = a.gC[e], ph = a.gD[e];
          const float dhn = m * dh;
          const float dpp = dhn * (1.f - z) * (1.f - p * p);
          v = part == 0 ? dpp * r : part == 1 ? dpp * a.gE[e] * r * (1.f - r) : part == 2 ? dhn * (ph - p) * z * (1.f - z) : dpp;
          if (first_tile && (s == 0 || part == 0)) a.oX[(size_t)b * 4 * kN + k] = v;
          if (first_tile && s == 1 && part == 0) atomicAdd(a.oAcc + e, (1.f - m) * dh + dhn * z);
        } else if (a.gen == 4) {       // GRU-1 backward (S:273-289 differentiated)
          const int part = k / kN, j = k - part * kN;
          const size_t e = (size_t)b * kN + j;
          const float m = a.gmask != nullptr ? a.gmask[b] : 1.f;
          const float d = a.gA[e];
          const float r = a.gG[(size_t)b * 2 * kN + j], z = a.gG[(size_t)b * 2 * kN + kN + j];
          const float p = a.gC[e];
          const float dp_ = m * d;
          const float dpp = dp_ * (1.f - z) * (1.f - p * p);
          v = part == 0 ? dpp * a.gD[e] * r * (1.f - r) : part == 1 ? dp_ * (a.gE[e] - p) * z * (1.f - z) : dpp * r;
          if (first_tile) {
            a.oX[(size_t)b * 3 * kN + k] = v;
            if (part == 2) { a.oProp[e] = dpp; atomicAdd(a.oAcc + e, (1.f - m) * d + dp_ * z); }
          }
        } else if (a.ids != nullptr) { const int id = a.ids[b]; if (id >= 0) v = __ldg(a.X + (size_t)id * a.ldx + k); }
        else v = __ldg(a.X + (size_t)b * a.ldx + k);
      }
      xs[bb][kk] = v;
    }
    __syncthreads();
    if (n < sg.N) {
#pragma unroll 4
      for (int bb = 0; bb < kSkRows; ++bb) {
        if (b0 + bb >= a.Bn) break;
        float4 acc = make4(0.f);
#pragma unroll
        for (int kk = 0; kk < kSkKC; ++kk) {
          const float x = xs[bb][kk];
          acc.x = fmaf(x, wv[kk].x, acc.x); acc.y = fmaf(x, wv[kk].y, acc.y);
          acc.z = fmaf(x, wv[kk].z, acc.z); acc.w = fmaf(x, wv[kk].w, acc.w);
        }
        float* o = sg.out + (size_t)(b0 + bb) * sg.ldo + n;
        atomicAdd(o, acc.x);
        if (n + 1 < sg.N) atomicAdd(o + 1, acc.y);
        if (n + 2 < sg.N) atomicAdd(o + 2, acc.z);
        if (n + 3 < sg.N) atomicAdd(o + 3, acc.w);
      }
    }
  }
}

int skinny_gemm(cudaStream_t st, SkinnyArgs a) {
  WAP_REQUIRE(a.nseg >= 1 && a.nseg <= 4, "skinny_gemm: nseg=%d", a.nseg);
  a.tile0[0] = 0;
  int kmax = 0;
  for (int s = 0; s < a.nseg; ++s) {
    a.tile0[s + 1] = a.tile0[s] + cdiv(a.seg[s].N, kSkCols);
    const int ks = a.seg[s].kcount > 0 ? a.seg[s].kcount : a.K;
    WAP_REQUIRE(a.seg[s].koff >= 0 && a.seg[s].koff + ks <= a.K, "skinny_gemm: segment %d reads X[%d:%d] of %d", s, a.seg[s].koff, a.seg[s].koff + ks, a.K);
    kmax = std::max(kmax, ks);
  }
  dim3 grid(a.tile0[a.nseg], cdiv(kmax, kSkKC));
  WAP_CHECK(launch_pdl(skinny_gemm_kernel, grid, 64, 0, st, a));
  WAP_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// row-broadcast fill (bias initialisation of the atomically-accumulated outputs)
// =====================================================================================
__global__ void __launch_bounds__(256) fill_rows_kernel(const FillArgs a) {
  pdl_sync();
  const FillSeg sg = a.seg[blockIdx.y];
  const int total = a.rows * sg.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / sg.n, j = i - b * sg.n;
    sg.dst[(size_t)b * sg.ld + j] = sg.bias != nullptr ? __ldg(sg.bias + j) : 0.f;
  }
}
int fill_rows(cudaStream_t st, const FillArgs& a) {
  int mx = 1;
  for (int s = 0; s < a.nseg; ++s) mx = max(mx, a.rows * a.seg[s].n);
  dim3 grid(min(cdiv(mx, 256), 64), a.nseg);
  WAP_CHECK(launch_pdl(fill_rows_kernel, grid, 256, 0, st, a));
  WAP_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// GRU-1 gates (S:273-289): r1|z1 = sigmoid(G1[:, :512]); prop = tanh(UH*r1 + E_h);
// pre_h = z1*h + (1-z1)*prop, blended with the caption mask.
// =====================================================================================
__global__ void __launch_bounds__(256)
gru1_gate_kernel(float* __restrict__ G1, const float* __restrict__ UH, const float* __restrict__ h_prev,
                 const float* __restrict__ y_mask, float* __restrict__ prop1, float* __restrict__ pre_h, int Bn, float* __restrict__ rz_out) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * kN) return;
  const int b = i / kN, j = i - b * kN;
  float* g = G1 + (size_t)b * 3 * kN;
  float r, z, p, ph;
  gru1_eval(g[j], g[kN + j], g[2 * kN + j], UH[i], h_prev[i], y_mask != nullptr ? y_mask[b] : 1.f, y_mask != nullptr, r, z, p, ph);
  if (rz_out != nullptr) { rz_out[(size_t)b * 2 * kN + j] = r; rz_out[(size_t)b * 2 * kN + kN + j] = z; }
  else { g[j] = r; g[kN + j] = z; }
  prop1[i] = p; pre_h[i] = ph;
}

// GRU-2 gates (S:292-306)
__global__ void __launch_bounds__(256)
gru2_gate_kernel(float* __restrict__ G2, const float* __restrict__ PH2, const float* __restrict__ CH,
                 const float* __restrict__ pre_h, const float* __restrict__ y_mask,
                 float* __restrict__ prop2, float* __restrict__ h_new, int Bn, float* __restrict__ rz_out) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * kN) return;
  const int b = i / kN, j = i - b * kN;
  float* g = G2 + (size_t)b * 2 * kN;
  float r, z, p, h;
  gru2_eval(g[j], g[kN + j], PH2[i], CH[i], pre_h[i], y_mask != nullptr ? y_mask[b] : 1.f, y_mask != nullptr, r, z, p, h);
  if (rz_out != nullptr) { rz_out[(size_t)b * 2 * kN + j] = r; rz_out[(size_t)b * 2 * kN + kN + j] = z; }
  else { g[j] = r; g[kN + j] = z; }
  prop2[i] = p; h_new[i] = h;
}
int gru2_gate(cudaStream_t st, float* G2, const float* PH2, const float* CH, const float* pre_h, const float* y_mask, float* prop2, float* h_new,
              int Bn, float* rz_out) {
  WAP_CHECK(launch_pdl(gru2_gate_kernel, cdiv(Bn * kN, 256), 256, 0, st, G2, PH2, CH, pre_h, y_mask, prop2, h_new, Bn, rz_out));
  WAP_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// attention energies (S:185-188): e[b,l] = V_a . tanh(cov[b,l,:] + watch[b,l,:] + speller[b,:]) + V_a_b
// one warp per annotation position; 128-bit loads; warp-shuffle reduction.
// =====================================================================================
__global__ void __launch_bounds__(256)
energy_kernel(const float* COV, const float* __restrict__ watch, const float* __restrict__ SP,
              const float* __restrict__ Va, const float* __restrict__ Vab, float* TANH,
              float* __restrict__ E, int Bn, int L, int shared) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Bn * L) return;
  const int b = warp / L, l = warp - b * L;
  const float* cv = COV + (size_t)warp * kNA;
  const float* wt = watch + (size_t)(shared ? l : warp) * kNA;
  const float* sp = SP + (size_t)b * kNA;
  float* th = TANH + (size_t)warp * kNA;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < kNA / 128; ++q) {
    const int c = q * 128 + lane * 4;
    const float4 a = *reinterpret_cast<const float4*>(cv + c);
    const float4 w4 = ldg4(wt + c), p4 = ldg4(sp + c), v4 = ldg4(Va + c);
    float4 t;
    t.x = tanhf(a.x + w4.x + p4.x); t.y = tanhf(a.y + w4.y + p4.y);
    t.z = tanhf(a.z + w4.z + p4.z); t.w = tanhf(a.w + w4.w + p4.w);
    *reinterpret_cast<float4*>(th + c) = t;
    s = fmaf(t.x, v4.x, s); s = fmaf(t.y, v4.y, s); s = fmaf(t.z, v4.z, s); s = fmaf(t.w, v4.w, s);
  }
  s = warp_sum(s);
  if (lane == 0) E[warp] = s + __ldg(Vab);
}

// =====================================================================================
// masked softmax over the L positions of one image + alpha_past update + context (S:190-203)
// grid = (ceil(684/64), Bn); every CTA recomputes the (cheap) softmax and owns 64 context channels.
// =====================================================================================
constexpr int kCtxChunk = 32;          // context channels per CTA: 22 x Bn CTAs (176 at B = 8) instead of 88
constexpr int kCtxLanes = kCtxChunk / 4, kCtxGroups = 256 / kCtxLanes;
__global__ void __launch_bounds__(256)
softmax_ctx_kernel(const float* __restrict__ E, const float* __restrict__ a_mask, const float* __restrict__ ap_in,
                   const float* __restrict__ ann, float* __restrict__ alpha, float* __restrict__ ap_out,
                   float* __restrict__ ctx, int L, int shared) {
  pdl_sync();
  extern __shared__ float sm[];          // alpha[L] | red[64] | part[kCtxGroups][kCtxChunk]
  float* s_alpha = sm;
  float* red = sm + ((L + 3) & ~3);      // keep `part` 16-byte aligned for the float4 stores
  float* part = red + 64;
  const int b = blockIdx.y, tid = threadIdx.x;
  const float* e = E + (size_t)b * L;
  const float* mk = a_mask != nullptr ? a_mask + (size_t)b * L : nullptr;
  float mx = -INFINITY;
  for (int l = tid; l < L; l += 256) if (mk == nullptr || mk[l] != 0.f) mx = fmaxf(mx, e[l]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int l = tid; l < L; l += 256) {
    float u = 0.f;
    if (mk == nullptr) u = expf(e[l] - mx);
    else if (mk[l] != 0.f) u = expf(e[l] - mx) * mk[l];
    s_alpha[l] = u; sum += u;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int l = tid; l < L; l += 256) {
    const float al = s_alpha[l] * inv;
    s_alpha[l] = al;
    if (blockIdx.x == 0) { alpha[(size_t)b * L + l] = al; ap_out[(size_t)b * L + l] = ap_in[(size_t)b * L + l] + al; }
  }
  __syncthreads();
  const int c0 = blockIdx.x * kCtxChunk;
  const int cq = tid % kCtxLanes, lg = tid / kCtxLanes;     // float4 lanes x position groups
  const int c = c0 + cq * 4;
  float4 acc = make4(0.f);
  if (c < kD) {
    const float* base = ann + (size_t)(shared ? 0 : b) * L * kD + c;
#pragma unroll 8
    for (int l = lg; l < L; l += kCtxGroups) {
      const float4 v = ldg4(base + (size_t)l * kD);
      const float al = s_alpha[l];
      acc.x = fmaf(al, v.x, acc.x); acc.y = fmaf(al, v.y, acc.y); acc.z = fmaf(al, v.z, acc.z); acc.w = fmaf(al, v.w, acc.w);
    }
  }
  *reinterpret_cast<float4*>(&part[lg * kCtxChunk + cq * 4]) = acc;
  __syncthreads();
  if (tid < kCtxChunk && c0 + tid < kD) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < kCtxGroups; ++g) s += part[g * kCtxChunk + tid];
    ctx[(size_t)b * kD + c0 + tid] = s;
  }
}

// Ppat[pos][tap (128, taps >= 121 zero)] = alpha_past[b, y + dy, x + dx]   (11x11 SAME patches, float4 per thread)
__global__ void __launch_bounds__(256) cov_im2col_kernel(const float* __restrict__ ap, float* __restrict__ Ppat, int Bn, int H, int W) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Bn * H * W * 32) return;
  const int pos = (int)(i >> 5), t4 = (int)(i & 31) * 4;
  const int x = pos % W, t = pos / W, y = t % H;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int tap = t4 + j, dy = tap / kCovK - kCovK / 2, dx = tap % kCovK - kCovK / 2;
    v[j] = (tap < kCovTaps && (unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W) ? __ldg(ap + pos + dy * W + dx) : 0.f;
  }
  *reinterpret_cast<float4*>(Ppat + (size_t)pos * 128 + t4) = make_float4(v[0], v[1], v[2], v[3]);
}
int cov_im2col(cudaStream_t st, const float* ap, float* Ppat, int Bn, int H, int W) {
  WAP_CHECK(launch_pdl(cov_im2col_kernel, (unsigned)cdivz((size_t)Bn * H * W * 32, 256), 256, 0, st, ap, Ppat, Bn, H, W));
  WAP_LAUNCH_CHECK();
  return 0;
}
// Q'^T hi/lo [512][128] (taps >= 121 zero) for the tensor-core coverage GEMM
__global__ void prep_qfT_kernel(const float* __restrict__ qf, float* __restrict__ hi, float* __restrict__ lo) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kNA * 128) return;
  const int c = i >> 7, tap = i & 127;
  const float v = tap < kCovTaps ? __ldg(qf + (size_t)tap * kNA + c) : 0.f;
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  const float h = __uint_as_float(u);
  hi[i] = h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v - h));
  lo[i] = __uint_as_float(u);
}
int prep_qfT(cudaStream_t st, const float* qf, float* hi, float* lo) {
  WAP_CHECK(launch_pdl(prep_qfT_kernel, cdiv(kNA * 128, 256), 256, 0, st, qf, hi, lo));
  WAP_LAUNCH_CHECK();
  return 0;
}

// energies -> masked softmax -> alpha_past update -> context, given the coverage tensor COV (S:185-203)
int attention_tail(cudaStream_t st, const float* p, const float* ann, const float* watch, const float* a_mask, int shared, int Bn, int L,
                   const float* SP, const float* ap_in, float* COV, float* TANH, float* E, float* alpha, float* ap_out, float* ctx) {
  const ParamLayout& P = layout();
  WAP_CHECK(launch_pdl(energy_kernel, cdiv(Bn * L * 32, 256), 256, 0, st, COV, watch, SP, p + P.V_a, p + P.V_a_b, TANH, E, Bn, L, shared));
  WAP_LAUNCH_CHECK();
  const size_t smem = sizeof(float) * (((L + 3) & ~3) + 64 + kCtxGroups * kCtxChunk);
  WAP_REQUIRE(smem <= 48 * 1024, "annotation grid too large for softmax_ctx_kernel: L=%d", L);
  dim3 grid(cdiv(kD, kCtxChunk), Bn);
  WAP_CHECK(launch_pdl(softmax_ctx_kernel, grid, 256, smem, st, E, a_mask, ap_in, ann, alpha, ap_out, ctx, L, shared));
  WAP_LAUNCH_CHECK();
  return 0;
}
// Attender.get_context after the speller projection: coverage GEMM (fp32 FFMA) + attention_tail
int attention_core(cudaStream_t st, const float* p, const float* ann, const float* watch, const float* qf, const float* a_mask, int shared,
                   int Bn, int h, int w, const float* SP, const float* ap_in, float* COV, float* TANH, float* E, float* alpha,
                   float* ap_out, float* ctx) {
  const ParamLayout& P = layout();
  ACovPatch A{ap_in, Bn, h, w};
  BRowMajor Bq{qf, kNA, kCovTaps, kNA};
  EpiStore e{COV, kNA, p + P.U_f_b, false, nullptr};
  WAP_TRY((launch_gemm<Cfg128x128>(st, A, Bq, e, Bn * h * w, kNA, 124, 1, kProfCovGemm)));
  return attention_tail(st, p, ann, watch, a_mask, shared, Bn, h * w, SP, ap_in, COV, TANH, E, alpha, ap_out, ctx);
}

// =====================================================================================
// one full decoder step
// =====================================================================================
int decoder_step(cudaStream_t st, const StepCtx& c) {
  const ParamLayout& P = layout();
  const float* p = c.params;
  const int Bn = c.Bn, L = c.L;
  if (!c.skip_fill) {   // bias-initialise the atomically accumulated buffers of this step
    FillArgs f; f.rows = Bn; f.nseg = 5;
    f.seg[0] = {c.UH, kN, kN, nullptr};
    f.seg[1] = {c.SP, kNA, kNA, p + P.W_a_b};
    f.seg[2] = {c.G2, 2 * kN, 2 * kN, p + P.b_hz_hr_nl};
    f.seg[3] = {c.PH2, kN, kN, p + P.b_rh_nl};
    f.seg[4] = {c.CH, kN, kN, nullptr};
    WAP_TRY(fill_rows(st, f));
  }
  const bool side_cov = c.tc_mode != 0 && c.side != nullptr;
  auto fork_cov = [&](const float* ap) -> int {     // coverage chain of one step on the side stream: im2col + GEMM -> COV
    WAP_CHECK(cudaEventRecord(c.ev_fork, st)); WAP_CHECK(cudaStreamWaitEvent(c.side, c.ev_fork, 0));
    WAP_CHECK(launch_pdl(cov_im2col_kernel, cdiv(Bn * L * 32, 256), 256, 0, c.side, ap, c.Ppat, Bn, c.h, c.w));
    WAP_LAUNCH_CHECK();
    WAP_TRY(tc_gemm(c.side, c.Ppat, 128, c.qfT_hi, c.tc_mode >= 2 ? c.qfT_lo : nullptr, 128, Bn * L, kNA, 128,
                    TcEpi{c.COV, kNA, 0, nullptr, p + P.U_f_b, 0}, kProfCovGemm));
    WAP_CHECK(cudaEventRecord(c.ev_join, c.side));
    return 0;
  };
  if (side_cov && !c.cov_ready) WAP_TRY(fork_cov(c.ap_in));   // ap_in is final (previous step's softmax): next to the GRU-1 half
  {   // h_prev * [U_hz_hr | U_rh]     (fused scan: h_prev itself is the GRU-2 gate of the previous step, evaluated here)
    SkinnyArgs a; a.X = c.h_prev; a.ldx = kN; a.ids = nullptr; a.Bn = Bn; a.K = kN; a.nseg = 2;
    a.seg[0] = {p + P.U_hz_hr, 2 * kN, 2 * kN, c.G1, 3 * kN};
    a.seg[1] = {p + P.U_rh, kN, kN, c.UH, kN};
    if (c.fuse_gates && c.prev_G2 != nullptr) {
      a.gen = 2; a.gG = c.prev_G2; a.gA = c.prev_PH2; a.gB = c.prev_CH; a.gC = c.prev_pre_h; a.gmask = c.prev_mask;
      a.oRZ = c.prev_RZ2; a.oProp = c.prev_prop2; a.oX = const_cast<float*>(c.h_prev);
    }
    WAP_TRY(skinny_gemm(st, a));
  }
  if (!c.fuse_gates) {
    WAP_CHECK(launch_pdl(gru1_gate_kernel, cdiv(Bn * kN, 256), 256, 0, st, c.G1, c.UH, c.h_prev, c.y_mask, c.prop1, c.pre_h, Bn, c.RZ1));
    WAP_LAUNCH_CHECK();
  }
  {   // pre_h * [W_a | U_hz_hr_nl | U_rh_nl]
    SkinnyArgs a; a.X = c.pre_h; a.ldx = kN; a.ids = nullptr; a.Bn = Bn; a.K = kN; a.nseg = 3;
    if (c.fuse_gates) {
      a.gen = 1; a.gG = c.G1; a.gA = c.UH; a.gB = c.h_prev; a.gmask = c.y_mask;
      a.oRZ = c.RZ1; a.oProp = c.prop1; a.oX = c.pre_h;
    }
    a.seg[0] = {p + P.W_a, kNA, kNA, c.SP, kNA};
    a.seg[1] = {p + P.U_hz_hr_nl, 2 * kN, 2 * kN, c.G2, 2 * kN};
    a.seg[2] = {p + P.U_rh_nl, kN, kN, c.PH2, kN};
    WAP_TRY(skinny_gemm(st, a));
  }
  if (side_cov) {
    WAP_CHECK(cudaStreamWaitEvent(st, c.ev_join, 0));
    WAP_TRY(attention_tail(st, p, c.ann, c.watch, c.a_mask, c.shared, Bn, L, c.SP, c.ap_in, c.COV, c.TANH, c.E, c.alpha, c.ap_out, c.ctx));
    if (c.cov_next) WAP_TRY(fork_cov(c.ap_out));      // COV / Ppat of this step are consumed: the next step's chain starts beside the ctx GEMM
  } else if (c.tc_mode != 0) {   // coverage on tensor cores: im2col patches [Bn*L,128] x Q'^T[512][128]  (+ U_f_b)
    WAP_CHECK(launch_pdl(cov_im2col_kernel, cdiv(Bn * L * 32, 256), 256, 0, st, c.ap_in, c.Ppat, Bn, c.h, c.w));
    WAP_LAUNCH_CHECK();
    WAP_TRY(tc_gemm(st, c.Ppat, 128, c.qfT_hi, c.tc_mode >= 2 ? c.qfT_lo : nullptr, 128, Bn * L, kNA, 128,
                    TcEpi{c.COV, kNA, 0, nullptr, p + P.U_f_b, 0}, kProfCovGemm));
    WAP_TRY(attention_tail(st, p, c.ann, c.watch, c.a_mask, c.shared, Bn, L, c.SP, c.ap_in, c.COV, c.TANH, c.E, c.alpha, c.ap_out, c.ctx));
  } else {
    WAP_TRY(attention_core(st, p, c.ann, c.watch, c.qf, c.a_mask, c.shared, Bn, c.h, c.w, c.SP, c.ap_in, c.COV, c.TANH, c.E, c.alpha, c.ap_out, c.ctx));
  }
  {   // ctx * [W_c_z_r | W_c_h_nl (| Wc)]
    SkinnyArgs a; a.X = c.ctx; a.ldx = kD; a.ids = nullptr; a.Bn = Bn; a.K = kD; a.nseg = 2;
    a.seg[0] = {p + P.W_c_z_r, 2 * kN, 2 * kN, c.G2, 2 * kN};
    a.seg[1] = {p + P.W_c_h_nl, kN, kN, c.CH, kN};
    if (c.has_extra_ctx) { a.seg[2] = c.extra_ctx; a.nseg = 3; }
    WAP_TRY(skinny_gemm(st, a));
  }
  if (!c.fuse_gates) {      // fused scan: the next step's first GEMM (or the caller after the last step) evaluates this gate
    WAP_CHECK(launch_pdl(gru2_gate_kernel, cdiv(Bn * kN, 256), 256, 0, st, c.G2, c.PH2, c.CH, c.pre_h, c.y_mask, c.prop2, c.h_new, Bn, c.RZ2));
    WAP_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace wap
