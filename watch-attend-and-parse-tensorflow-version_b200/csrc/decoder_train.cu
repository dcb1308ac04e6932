// Training-side decoder: WAP.get_cost forward (model-single-GPU.py:337-375: embedding shift, h0,
// tf.scan over Parser.one_time_step, logits/maxout/dropout/softmax-CE) and its backward (BPTT
// through the coverage attention).
//
// Formulation differences from the reference graph that do not change the mathematics:
//   * U_a*a (+b) is hoisted out of the scan (the reference recomputes it every step, S:177);
//   * coverage conv (S:171) and U_f projection (S:173) are fused into one 121-tap filter
//     Q' = Q*U_f, recomputed from the current weights at every call; gradients are mapped back
//     with dQ = dQ'*U_f^T, dU_f = Q^T*dQ';
//   * every weight gradient is ONE GEMM over all T*B rows after the time loop instead of T
//     rank-B updates inside it; only the data gradients stay in the sequential loop.
#include <algorithm>

#include "decoder.cuh"
#include "tc_gemm.cuh"

namespace wap {

int get_conv_precision();
int prep_qfT(cudaStream_t st, const float* qf, float* hi, float* lo);
int cov_im2col(cudaStream_t st, const float* ap, float* Ppat, int Bn, int H, int W);
size_t step_ws_floats(int Bn, int L);
int attn_precompute(cudaStream_t st, const float* params, const float* ann, const float* a_mask, int B, int h, int w,
                    float* watch, float* h0, float* qf, float* mean_scratch);

// ------------------------------------------------------------------------------------ workspace
constexpr int kLdA = 4 * kN, kLdB = 3 * kN;
int g_fuse_gates = 3;       // wap_debug_set_fuse_gates(): bit 0 forward GRU gates inside the consuming skinny GEMMs, bit 1 the same for the BPTT + fused attention backward
struct DecTrainWs {
  // forward (saved)
  float *EMB, *G1, *H, *UH, *PROP1, *PREH, *SP, *G2, *PH2, *CH, *PROP2, *TANH, *ALPHA, *AP, *CTX;
  float *RZ1, *RZ2;          // sigmoid gates r|z of GRU-1 / GRU-2 per step [TB, 512] (G1 / G2 keep the raw pre-activations)
  float *watch, *qf, *mean, *COV, *E, *LG, *MO, *LO;
  double* cost;
  // backward
  float *dLG, *dMO, *dHlog, *dCTX, *dEMB, *dG1, *dEh, *dUH, *dSP, *dG2, *dCH, *dPH2, *DZ, *dWatch, *DF;
  float *dAP0, *dAP1, *dalpha, *de, *dh0, *dh1, *dpre_h, *dQf, *dmean;
  float *dVaP, *dVabP;       // fused BPTT: per-(step, image) partial sums of dV_a [TB, 512] / dV_a_b [TB] (16 CTAs per address instead of all 128), column-summed after the loop
  float *DH, *DPREH;         // fused BPTT: per-step carried gradients [(T+1)B, 256] / [TB, 256], zero-filled once (every contribution is an atomic add)
  float *T_Wczr, *T_Wchnl, *T_Uhzhrnl, *T_Urhnl, *T_Wa, *T_Uhzhr, *T_Urh;
  float *qf_hi, *qf_lo, *Ua_hi, *Ua_lo, *UaT_hi, *UaT_lo;    // tensor-core operand copies (3xTF32 hi/lo parts)
  float *Ppat, *qfT_hi, *qfT_lo;                              // coverage im2col patches [B*L,128], Q'^T [512][128]
  // hi/lo operand copies of the head's weights (tensor-core form of the small dense GEMMs): forward uses W^T ([N][K] of a [K][N] matrix)
  float *UfT[2], *WyzrT[2], *WyhT[2], *WhT[2], *WcT[2], *WyT[2], *WhN[2], *WcN[2], *WyN[2];
  float *PpatAll;                                             // patches of all T steps [T*B*L, 128] for the tensor-core dQ' weight gradient
};
static DecTrainWs carve_train(float* base, int B, int L, int T, size_t* total) {
  DecTrainWs w; size_t o = 0;
  auto take = [&](size_t n) { float* p = base + o; o += (n + 3) & ~(size_t)3; return p; };
  const size_t TB = (size_t)T * B, BL = (size_t)B * L;
  w.cost = reinterpret_cast<double*>(take(4));
  w.EMB = take(TB * kMw); w.G1 = take(TB * 3 * kN); w.H = take((TB + B) * kN); w.UH = take(TB * kN);
  w.PROP1 = take(TB * kN); w.PREH = take(TB * kN); w.SP = take(TB * kNA); w.G2 = take(TB * 2 * kN);
  w.PH2 = take(TB * kN); w.CH = take(TB * kN); w.PROP2 = take(TB * kN); w.TANH = take(TB * L * kNA);
  w.RZ1 = take(TB * 2 * kN); w.RZ2 = take(TB * 2 * kN);
  w.ALPHA = take(TB * L); w.AP = take((TB + B) * L); w.CTX = take(TB * kD);
  w.watch = take(BL * kNA); w.qf = take((size_t)kCovTaps * kNA); w.mean = take((size_t)B * kD);
  w.COV = take(BL * kNA); w.E = take(BL); w.LG = take(TB * kMw); w.MO = take(TB * kMw / 2); w.LO = take(TB * kVocab);
  w.dLG = take(TB * kMw); w.dMO = take(TB * kMw / 2); w.dHlog = take(TB * kN); w.dCTX = take(TB * kD); w.dEMB = take(TB * kMw);
  // per-step gradient rows laid out so that the BPTT data gradients are single skinny GEMMs over concatenated operands:
  //   DA row (kLdA = 1024): [dPH2 | dG2 (r|z) | dCH]     DB row (kLdB = 768): [dG1 (r|z) | dUH]
  w.dPH2 = take(TB * kLdA); w.dG2 = w.dPH2 + kN; w.dCH = w.dPH2 + 3 * kN;
  w.dG1 = take(TB * kLdB); w.dUH = w.dG1 + 2 * kN;
  w.dEh = take(TB * kN); w.dSP = take(TB * kNA); w.DZ = take(TB * L * kNA); w.dWatch = take(BL * kNA); w.DF = take(BL * 124 * 2);     // two DF buffers (split-K ping-pong of the fused BPTT)
  w.dAP0 = take(BL); w.dAP1 = take(BL); w.dalpha = take(BL); w.de = take(BL);
  w.dh0 = take((size_t)B * kN); w.dh1 = take((size_t)B * kN); w.dpre_h = take((size_t)B * kN);
  w.dQf = take((size_t)kCovTaps * kNA); w.dmean = take((size_t)B * kD);
  w.DH = take((TB + B) * kN); w.DPREH = take(TB * kN); w.dVaP = take(TB * kNA); w.dVabP = take(TB);
  // stacked transposes matching the concatenated operands: [T_Wczr; T_Wchnl] (768x684), [T_Urhnl; T_Uhzhrnl] (768x256), [T_Uhzhr; T_Urh] (768x256)
  w.T_Wczr = take((size_t)2 * kN * kD); w.T_Wchnl = take((size_t)kN * kD); w.T_Urhnl = take((size_t)kN * kN);
  w.T_Uhzhrnl = take((size_t)2 * kN * kN); w.T_Wa = take((size_t)kNA * kN); w.T_Uhzhr = take((size_t)2 * kN * kN); w.T_Urh = take((size_t)kN * kN);
  w.qf_hi = take((size_t)kCovTaps * kNA); w.qf_lo = take((size_t)kCovTaps * kNA);
  w.Ua_hi = take((size_t)kD * kNA); w.Ua_lo = take((size_t)kD * kNA); w.UaT_hi = take((size_t)kD * kNA); w.UaT_lo = take((size_t)kD * kNA);
  w.Ppat = take(BL * 128); w.qfT_hi = take((size_t)kNA * 128); w.qfT_lo = take((size_t)kNA * 128);
  w.PpatAll = take(TB * L * 128);
  for (int i = 0; i < 2; ++i) {
    w.UfT[i] = take((size_t)kNA * kNA); w.WyzrT[i] = take((size_t)2 * kN * kMw); w.WyhT[i] = take((size_t)kN * kMw); w.WhT[i] = take((size_t)kMw * kN);
    w.WcT[i] = take((size_t)kMw * kD); w.WyT[i] = take((size_t)kMw * kMw); w.WhN[i] = take((size_t)kN * kMw); w.WcN[i] = take((size_t)kD * kMw); w.WyN[i] = take((size_t)kMw * kMw);
  }
  if (total != nullptr) *total = o;
  return w;
}
size_t decoder_train_ws_floats(int B, int L, int T) { size_t t; carve_train(nullptr, B, L, T, &t); return t; }

// ------------------------------------------------------------------------------------ small kernels
// EMB[t,b,:] = t == 0 ? 0 : embed[y[t-1,b]]   (S:340-343)
__global__ void emb_shift_kernel(const float* __restrict__ embed, const int* __restrict__ y, float* __restrict__ EMB, int T, int B) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * B * kMw) return;
  const int j = i % kMw, r = i / kMw, t = r / B;
  EMB[i] = (t == 0) ? 0.f : __ldg(embed + (size_t)y[r - B] * kMw + j);
}
__global__ void emb_scatter_kernel(const float* __restrict__ dEMB, const int* __restrict__ y, float* __restrict__ dembed, int T, int B) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * B * kMw) return;
  const int j = i % kMw, r = i / kMw, t = r / B;
  if (t > 0) atomicAdd(dembed + (size_t)y[r - B] * kMw + j, dEMB[i]);
}
__global__ void bias3_rows_kernel(float* LG, int rows, const float* by, const float* bc, const float* bh) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * kMw) { const int j = i % kMw; LG[i] = by[j] + bc[j] + bh[j]; }
}
// maxout over adjacent pairs + dropout (site kNumConv) (S:359-362)
__global__ void maxout_drop_kernel(const float* __restrict__ LG, float* __restrict__ MO, int rows, int dropout, uint32_t seed,
                                   const uint32_t* __restrict__ seed_src) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (kMw / 2)) return;
  if (dropout && seed_src != nullptr) seed = __ldg(seed_src);
  const float2 v = *reinterpret_cast<const float2*>(LG + 2 * (size_t)i);
  float m = fmaxf(v.x, v.y);
  if (dropout) m = drop_keep(seed, kNumConv, (uint32_t)i) ? m * (1.f / (1.f - kDropRate)) : 0.f;
  MO[i] = m;
}
__global__ void maxout_drop_bwd_kernel(const float* __restrict__ LG, const float* __restrict__ dMO, float* __restrict__ dLG,
                                       int rows, int dropout, uint32_t seed,
                                   const uint32_t* __restrict__ seed_src) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (kMw / 2)) return;
  if (dropout && seed_src != nullptr) seed = __ldg(seed_src);
  const float2 v = *reinterpret_cast<const float2*>(LG + 2 * (size_t)i);
  float g = dMO[i];
  if (dropout) g = drop_keep(seed, kNumConv, (uint32_t)i) ? g * (1.f / (1.f - kDropRate)) : 0.f;
  float2 o = (v.x >= v.y) ? make_float2(g, 0.f) : make_float2(0.f, g);
  *reinterpret_cast<float2*>(dLG + 2 * (size_t)i) = o;
}
// LO[r,j] = bo[j] + sum_i MO[r,i] Wo[i,j]
__global__ void wo_fwd_kernel(const float* __restrict__ MO, const float* __restrict__ Wo, const float* __restrict__ bo,
                              float* __restrict__ LO, int rows) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kVocab) return;
  const int r = i / kVocab, j = i - r * kVocab;
  float s = __ldg(bo + j);
  const float* m = MO + (size_t)r * (kMw / 2);
#pragma unroll 8
  for (int k = 0; k < kMw / 2; ++k) s = fmaf(m[k], __ldg(Wo + k * kVocab + j), s);
  LO[i] = s;
}
__global__ void wo_bwd_data_kernel(const float* __restrict__ dLO, const float* __restrict__ Wo, float* __restrict__ dMO, int rows) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (kMw / 2)) return;
  const int r = i / (kMw / 2), k = i - r * (kMw / 2);
  const float* d = dLO + (size_t)r * kVocab; const float* wr = Wo + (size_t)k * kVocab;
  float s = 0.f;
  for (int j = 0; j < kVocab; ++j) s = fmaf(d[j], __ldg(wr + j), s);
  dMO[i] = s;
}
__global__ void wo_bwd_weight_kernel(const float* __restrict__ MO, const float* __restrict__ dLO, float* __restrict__ dWo,
                                     float* __restrict__ dbo, int rows) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (kMw / 2 + 1) * kVocab) return;
  const int k = i / kVocab, j = i - k * kVocab;
  float s = 0.f;
  if (k < kMw / 2) { for (int r = 0; r < rows; ++r) s = fmaf(MO[(size_t)r * (kMw / 2) + k], dLO[(size_t)r * kVocab + j], s); dWo[i] += s; }
  else { for (int r = 0; r < rows; ++r) s += dLO[(size_t)r * kVocab + j]; dbo[j] += s; }
}
// softmax cross-entropy (S:368-374): cost += mask*CE/B ; dLO = mask/B*(softmax - onehot) in place.  warp per row.
__global__ void __launch_bounds__(256)
ce_kernel(float* __restrict__ LO, const int* __restrict__ y, const float* __restrict__ y_mask, double* __restrict__ cost,
          int rows, int B, float grad_scale) {
  pdl_sync();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  float* lo = LO + (size_t)r * kVocab;
  float v[4]; float mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < 4; ++q) { const int j = q * 32 + lane; v[q] = j < kVocab ? lo[j] : -INFINITY; mx = fmaxf(mx, v[q]); }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) { const int j = q * 32 + lane; if (j < kVocab) s += expf(v[q] - mx); }
  s = warp_sum(s);
  const float lse = mx + logf(s);
  const int tgt = y[r];
  const float m = y_mask[r];
  if (lane == 0) atomicAdd(cost, (double)(m * (lse - lo[tgt]) / (float)B));
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = q * 32 + lane;
    if (j < kVocab) lo[j] = grad_scale * m / (float)B * (expf(v[q] - lse) - (j == tgt ? 1.f : 0.f));
  }
}
// column sums: out[j] += sum_r X[r*ld + j]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int ld, int rows, int n, float* __restrict__ out) {
  pdl_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int r0 = blockIdx.y * 64, r1 = min(rows, r0 + 64);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += X[(size_t)r * ld + j];
  atomicAdd(out + j, s);
}
static int colsum(cudaStream_t st, const float* X, int ld, int rows, int n, float* out) {
  WAP_CHECK(launch_pdl(colsum_kernel, dim3(cdiv(n, 256), cdiv(rows, 64)), 256, 0, st, X, ld, rows, n, out));
  WAP_LAUNCH_CHECK();
  return 0;
}
// up to 8 transposes in one launch (blockIdx.z = matrix)
struct TransposeBatch { const float* src[8]; float* dst[8]; int R[8], C[8]; int n; };
__global__ void transpose_multi_kernel(const TransposeBatch b) {
  pdl_sync();
  __shared__ float t[32][33];
  const int m = blockIdx.z, R = b.R[m], C = b.C[m];
  if ((int)blockIdx.x * 32 >= C || (int)blockIdx.y * 32 >= R) return;
  const float* src = b.src[m]; float* dst = b.dst[m];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) if (r0 + i < R && c < C) t[i][threadIdx.x] = src[(size_t)(r0 + i) * C + c];
  __syncthreads();
  const int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) if (c0 + i < C && r < R) dst[(size_t)(c0 + i) * R + r] = t[threadIdx.x][i];
}
static int transpose_multi(cudaStream_t st, const TransposeBatch& b) {
  int mc = 1, mr = 1;
  for (int i = 0; i < b.n; ++i) { mc = std::max(mc, cdiv(b.C[i], 32)); mr = std::max(mr, cdiv(b.R[i], 32)); }
  WAP_CHECK(launch_pdl(transpose_multi_kernel, dim3(mc, mr, b.n), dim3(32, 8), 0, st, b));
  WAP_LAUNCH_CHECK();
  return 0;
}
__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
__global__ void add_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// ------------------------------------------------------------------------------------ forward
int decoder_train_forward(cudaStream_t st, const float* params, const float* ann, const float* a_mask, const int* y,
                          const float* y_mask, int B, int h, int w, int T, int dropout, uint32_t seed,
                          double* cost_out, float* wsp) {
  const ParamLayout& P = layout();
  const int L = h * w, TB = T * B;
  DecTrainWs W = carve_train(wsp, B, L, T, nullptr);
  const int prec = get_conv_precision();
  const bool tc = prec != 0;
  WAP_REQUIRE(!tc || tc_available(), "tensor-core conv precision requested but cuTensorMapEncodeTiled is unavailable (no silent FFMA fallback)");
  auto lo = [&](float* const* pair) -> const float* { return prec >= 2 ? pair[1] : nullptr; };
  if (tc) {   // hoisted watch GEMM [B*L,684]x[684,512] on tensor cores, and every small dense GEMM of the head as well: the FFMA tiles took
              // 30-110 us each at M = T*B rows (0.3 ms of serial work before and after the scan); one launch splits all their weights into hi/lo
    TcPrepBatch pb; pb.n = 7;
    pb.e[0] = {params + P.U_a, kD, kNA, 1, W.UaT_hi, W.UaT_lo};
    pb.e[1] = {params + P.U_f, kNA, kNA, 1, W.UfT[0], W.UfT[1]};
    pb.e[2] = {params + P.W_yz_yr, kMw, 2 * kN, 1, W.WyzrT[0], W.WyzrT[1]};
    pb.e[3] = {params + P.W_yh, kMw, kN, 1, W.WyhT[0], W.WyhT[1]};
    pb.e[4] = {params + P.Wh, kN, kMw, 1, W.WhT[0], W.WhT[1]};
    pb.e[5] = {params + P.Wc, kD, kMw, 1, W.WcT[0], W.WcT[1]};
    pb.e[6] = {params + P.Wy, kMw, kMw, 1, W.WyT[0], W.WyT[1]};
    WAP_TRY(tc_prep_multi(st, pb));
    WAP_TRY(attn_precompute(st, params, ann, a_mask, B, h, w, nullptr, W.H, nullptr, W.mean));
    WAP_TRY(tc_gemm(st, params + P.Q, kNA, W.UfT[0], lo(W.UfT), kNA, kCovTaps, kNA, kNA, TcEpi{W.qf, kNA, 0, nullptr, nullptr, 0}, kProfOtherGemm));   // Q' = Q U_f
    WAP_TRY(tc_gemm(st, ann, kD, W.UaT_hi, prec >= 2 ? W.UaT_lo : nullptr, kD, B * L, kNA, kD, TcEpi{W.watch, kNA, 0, nullptr, params + P.U_a_b, 0}, kProfOtherGemm));
    WAP_TRY(prep_qfT(st, W.qf, W.qfT_hi, W.qfT_lo));
  } else {
    WAP_TRY(attn_precompute(st, params, ann, a_mask, B, h, w, W.watch, W.H, W.qf, W.mean));   // H[0] = h0 (masked mean, S:345-347)
  }
  WAP_CHECK(launch_pdl(emb_shift_kernel, cdiv(TB * kMw, 256), 256, 0, st, params + P.embed, y, W.EMB, T, B));
  WAP_LAUNCH_CHECK();
  // emb-dependent GRU-1 pre-activations for all steps at once
  if (tc) {
    WAP_TRY(tc_gemm(st, W.EMB, kMw, W.WyzrT[0], lo(W.WyzrT), kMw, TB, 2 * kN, kMw, TcEpi{W.G1, 3 * kN, 0, nullptr, params + P.b_yz_yr, 0}, kProfOtherGemm));
    WAP_TRY(tc_gemm(st, W.EMB, kMw, W.WyhT[0], lo(W.WyhT), kMw, TB, kN, kMw, TcEpi{W.G1 + 2 * kN, 3 * kN, 0, nullptr, params + P.b_yh, 0}, kProfOtherGemm));
  } else {
    WAP_TRY(gemm_nn(st, W.EMB, kMw, params + P.W_yz_yr, 2 * kN, W.G1, 3 * kN, TB, 2 * kN, kMw, params + P.b_yz_yr, false));
    WAP_TRY(gemm_nn(st, W.EMB, kMw, params + P.W_yh, kN, W.G1 + 2 * kN, 3 * kN, TB, kN, kMw, params + P.b_yh, false));
  }
  WAP_CHECK(cudaMemsetAsync(W.AP, 0, sizeof(float) * B * L, st));
  {   // bias rows of the atomically accumulated per-step buffers, all T steps at once
    FillArgs f; f.rows = TB; f.nseg = 5;
    f.seg[0] = {W.UH, kN, kN, nullptr};
    f.seg[1] = {W.SP, kNA, kNA, params + P.W_a_b};
    f.seg[2] = {W.G2, 2 * kN, 2 * kN, params + P.b_hz_hr_nl};
    f.seg[3] = {W.PH2, kN, kN, params + P.b_rh_nl};
    f.seg[4] = {W.CH, kN, kN, nullptr};
    WAP_TRY(fill_rows(st, f));
  }
  SideStreams* sl = (tc && overlap_enabled()) ? side_streams() : nullptr;
  const bool fuse = (g_fuse_gates & 1) != 0;     // GRU gates evaluated inside the consuming skinny GEMMs (2 kernels less per time step)
  for (int t = 0; t < T; ++t) {
    const size_t r = (size_t)t * B;
    StepCtx c;
    c.Bn = B; c.h = h; c.w = w; c.L = L; c.params = params; c.ann = ann; c.watch = W.watch; c.shared = 0; c.qf = W.qf;
    c.a_mask = a_mask; c.y_mask = y_mask + r; c.h_prev = W.H + r * kN; c.ap_in = W.AP + r * L;
    c.G1 = W.G1 + r * 3 * kN; c.UH = W.UH + r * kN; c.prop1 = W.PROP1 + r * kN; c.pre_h = W.PREH + r * kN;
    c.SP = W.SP + r * kNA; c.G2 = W.G2 + r * 2 * kN; c.PH2 = W.PH2 + r * kN; c.CH = W.CH + r * kN; c.prop2 = W.PROP2 + r * kN;
    c.COV = W.COV; c.TANH = W.TANH + r * L * kNA; c.E = W.E; c.alpha = W.ALPHA + r * L; c.ap_out = W.AP + (r + B) * L;
    c.ctx = W.CTX + r * kD; c.h_new = W.H + (r + B) * kN; c.has_extra_ctx = 0;
    c.tc_mode = tc ? prec : 0; c.Ppat = W.Ppat; c.qfT_hi = W.qfT_hi; c.qfT_lo = W.qfT_lo;
    c.skip_fill = 1;
    c.fuse_gates = fuse ? 1 : 0; c.RZ1 = W.RZ1 + r * 2 * kN; c.RZ2 = W.RZ2 + r * 2 * kN;
    if (fuse && t > 0) {
      const size_t q = r - B;
      c.prev_G2 = W.G2 + q * 2 * kN; c.prev_PH2 = W.PH2 + q * kN; c.prev_CH = W.CH + q * kN; c.prev_pre_h = W.PREH + q * kN; c.prev_mask = y_mask + q;
      c.prev_RZ2 = W.RZ2 + q * 2 * kN; c.prev_prop2 = W.PROP2 + q * kN;
    }
    if (sl != nullptr) { c.side = sl->s[2]; c.ev_fork = sl->fork[2]; c.ev_join = sl->done[6]; c.cov_ready = t > 0; c.cov_next = t + 1 < T; }
    WAP_TRY(decoder_step(st, c));
  }
  if (fuse) {   // the GRU-2 gate of the last step has no following step to evaluate it
    const size_t q = (size_t)(T - 1) * B;
    WAP_TRY(gru2_gate(st, W.G2 + q * 2 * kN, W.PH2 + q * kN, W.CH + q * kN, W.PREH + q * kN, y_mask + q, W.PROP2 + q * kN, W.H + (q + B) * kN, B,
                      W.RZ2 + q * 2 * kN));
  }
  // logits for all steps (S:354-364)
  WAP_CHECK(launch_pdl(bias3_rows_kernel, cdiv(TB * kMw, 256), 256, 0, st, W.LG, TB, params + P.by, params + P.bc, params + P.bh));
  WAP_LAUNCH_CHECK();
  if (tc) {
    // few output tiles (T*B rows x 256 columns), long reductions: split-K, the partial sums added atomically onto the bias rows
    const int sk = TB <= 1024 ? 4 : 1;
    WAP_TRY(tc_gemm(st, W.H + (size_t)B * kN, kN, W.WhT[0], lo(W.WhT), kN, TB, kMw, kN, TcEpi{W.LG, kMw, sk > 1 ? 0 : 1, nullptr, nullptr, 0}, kProfOtherGemm, sk));
    WAP_TRY(tc_gemm(st, W.CTX, kD, W.WcT[0], lo(W.WcT), kD, TB, kMw, kD, TcEpi{W.LG, kMw, sk > 1 ? 0 : 1, nullptr, nullptr, 0}, kProfOtherGemm, sk > 1 ? 2 * sk : 1));
    WAP_TRY(tc_gemm(st, W.EMB, kMw, W.WyT[0], lo(W.WyT), kMw, TB, kMw, kMw, TcEpi{W.LG, kMw, sk > 1 ? 0 : 1, nullptr, nullptr, 0}, kProfOtherGemm, sk));
  } else {
    WAP_TRY(gemm_nn(st, W.H + (size_t)B * kN, kN, params + P.Wh, kMw, W.LG, kMw, TB, kMw, kN, nullptr, true));
    WAP_TRY(gemm_nn(st, W.CTX, kD, params + P.Wc, kMw, W.LG, kMw, TB, kMw, kD, nullptr, true));
    WAP_TRY(gemm_nn(st, W.EMB, kMw, params + P.Wy, kMw, W.LG, kMw, TB, kMw, kMw, nullptr, true));
  }
  WAP_CHECK(launch_pdl(maxout_drop_kernel, cdiv(TB * kMw / 2, 256), 256, 0, st, W.LG, W.MO, TB, dropout, seed, seed_src()));
  WAP_LAUNCH_CHECK();
  WAP_CHECK(launch_pdl(wo_fwd_kernel, cdiv(TB * kVocab, 256), 256, 0, st, W.MO, params + P.Wo, params + P.bo, W.LO, TB));
  WAP_LAUNCH_CHECK();
  WAP_CHECK(cudaMemsetAsync(W.cost, 0, sizeof(double), st));
  WAP_CHECK(launch_pdl(ce_kernel, cdiv(TB * 32, 256), 256, 0, st, W.LO, y, y_mask, W.cost, TB, B, 1.f));
  WAP_LAUNCH_CHECK();
  if (cost_out != nullptr) WAP_CHECK(cudaMemcpyAsync(cost_out, W.cost, sizeof(double), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ------------------------------------------------------------------------------------ backward kernels
// GRU-2 backward (elementwise).  dh_in: carry; dHlog: logit-path gradient of h_t.
__global__ void __launch_bounds__(256)
gru2_bwd_kernel(const float* __restrict__ dh_in, const float* __restrict__ dHlog, const float* __restrict__ G2,
                const float* __restrict__ prop2, const float* __restrict__ PH2, const float* __restrict__ pre_h,
                const float* __restrict__ y_mask, float* __restrict__ dG2, float* __restrict__ dCH, float* __restrict__ dPH2,
                float* __restrict__ dpre_h, int Bn) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * kN) return;
  const int b = i / kN, j = i - b * kN;
  const float m = y_mask != nullptr ? y_mask[b] : 1.f;
  const float dh = dh_in[i] + dHlog[i];
  const float r = G2[(size_t)b * 2 * kN + j], z = G2[(size_t)b * 2 * kN + kN + j];       // G2 = the saved gates RZ2
  const float p = prop2[i], ph = pre_h[i];
  const float dhn = m * dh;
  const float dz = dhn * (ph - p), dp = dhn * (1.f - z);
  const float dpp = dp * (1.f - p * p);
  dCH[(size_t)b * kLdA + j] = dpp;
  dPH2[(size_t)b * kLdA + j] = dpp * r;
  const float dr = dpp * PH2[i];
  dG2[(size_t)b * kLdA + j] = dr * r * (1.f - r);
  dG2[(size_t)b * kLdA + kN + j] = dz * z * (1.f - z);
  dpre_h[i] = (1.f - m) * dh + dhn * z;
}
// GRU-1 backward (elementwise)
__global__ void __launch_bounds__(256)
gru1_bwd_kernel(const float* __restrict__ dpre_h, const float* __restrict__ G1, const float* __restrict__ prop1,
                const float* __restrict__ UH, const float* __restrict__ h_prev, const float* __restrict__ y_mask,
                float* __restrict__ dG1, float* __restrict__ dEh, float* __restrict__ dUH, float* __restrict__ dh_out, int Bn) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * kN) return;
  const int b = i / kN, j = i - b * kN;
  const float m = y_mask != nullptr ? y_mask[b] : 1.f;
  const float d = dpre_h[i];
  const float r = G1[(size_t)b * 2 * kN + j], z = G1[(size_t)b * 2 * kN + kN + j];       // G1 = the saved gates RZ1 (ld 2kN)
  const float p = prop1[i], h = h_prev[i];
  const float dp_ = m * d;
  const float dz = dp_ * (h - p), dprop = dp_ * (1.f - z);
  const float dpp = dprop * (1.f - p * p);
  dEh[i] = dpp;
  dUH[(size_t)b * kLdB + j] = dpp * r;
  const float dr = dpp * UH[i];
  dG1[(size_t)b * kLdB + j] = dr * r * (1.f - r);
  dG1[(size_t)b * kLdB + kN + j] = dz * z * (1.f - z);
  dh_out[i] = (1.f - m) * d + dp_ * z;
}
// d alpha[b,l] = a[b,l,:] . dctx[b,:] + dAP[b,l]    (warp per position)
__global__ void __launch_bounds__(256)
dalpha_kernel(const float* __restrict__ ann, const float* __restrict__ dctx, const float* __restrict__ dAP,
              float* __restrict__ dalpha, int Bn, int L) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Bn * L) return;
  const int b = warp / L;
  const float* a = ann + (size_t)warp * kD; const float* d = dctx + (size_t)b * kD;
  float s = 0.f;
  for (int c = lane * 4; c < kD; c += 128) {
    const float4 av = ldg4(a + c), dv = ldg4(d + c);
    s = fmaf(av.x, dv.x, s); s = fmaf(av.y, dv.y, s); s = fmaf(av.z, dv.z, s); s = fmaf(av.w, dv.w, s);
  }
  s = warp_sum(s);
  if (lane == 0) dalpha[warp] = s + dAP[warp];
}
// softmax backward per image: de = alpha * (dalpha - sum(alpha*dalpha))
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ dalpha, float* __restrict__ de, int L) {
  pdl_sync();
  __shared__ float red[33];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int l = threadIdx.x; l < L; l += 256) s = fmaf(alpha[(size_t)b * L + l], dalpha[(size_t)b * L + l], s);
  s = block_sum(s, red);
  for (int l = threadIdx.x; l < L; l += 256) { const float a = alpha[(size_t)b * L + l]; de[(size_t)b * L + l] = a * (dalpha[(size_t)b * L + l] - s); }
}
// dz[b,l,c] = de[b,l] * V_a[c] * (1 - tanh^2)  -> DZ, dWatch += dz, dSP[b,c] += sum_l dz, dV_a[c] += sum de*tanh, dV_a_b += sum de
constexpr int kDzPix = 16;        // 16 rows per CTA: half the atomics on dSP / dV_a of 8 (the contended part), still 128 CTAs x 16 loads in flight
__global__ void __launch_bounds__(128)
dz_kernel(const float* __restrict__ TANH, const float* __restrict__ de, const float* __restrict__ Va,
          float* __restrict__ DZ, float* __restrict__ dSP, float* __restrict__ dVa, float* __restrict__ dVab, int L) {
  pdl_sync();
  const int b = blockIdx.y, l0 = blockIdx.x * kDzPix, c = threadIdx.x * 4;
  const float4 va = ldg4(Va + c);
  float4 sp = make4(0.f), dv = make4(0.f);
  float sde = 0.f;
  float4 t[kDzPix]; float d[kDzPix];
#pragma unroll
  for (int j = 0; j < kDzPix; ++j) {          // all loads in flight before the first use
    const bool ok = l0 + j < L;
    const size_t row = (size_t)b * L + (ok ? l0 + j : l0);
    d[j] = ok ? de[row] : 0.f;
    t[j] = ldg4(TANH + row * kNA + c);
  }
#pragma unroll
  for (int j = 0; j < kDzPix; ++j) {
    if (l0 + j >= L) break;
    const size_t row = (size_t)b * L + l0 + j;
    float4 z;
    z.x = d[j] * va.x * (1.f - t[j].x * t[j].x); z.y = d[j] * va.y * (1.f - t[j].y * t[j].y);
    z.z = d[j] * va.z * (1.f - t[j].z * t[j].z); z.w = d[j] * va.w * (1.f - t[j].w * t[j].w);
    *reinterpret_cast<float4*>(DZ + row * kNA + c) = z;
    sp.x += z.x; sp.y += z.y; sp.z += z.z; sp.w += z.w;
    dv.x = fmaf(d[j], t[j].x, dv.x); dv.y = fmaf(d[j], t[j].y, dv.y); dv.z = fmaf(d[j], t[j].z, dv.z); dv.w = fmaf(d[j], t[j].w, dv.w);
    sde += d[j];
  }
  float* s = dSP + (size_t)b * kNA + c;
  atomicAdd(s, sp.x); atomicAdd(s + 1, sp.y); atomicAdd(s + 2, sp.z); atomicAdd(s + 3, sp.w);
  atomicAdd(dVa + c, dv.x); atomicAdd(dVa + c + 1, dv.y); atomicAdd(dVa + c + 2, dv.z); atomicAdd(dVa + c + 3, dv.w);
  if (threadIdx.x == 0) atomicAdd(dVab, sde);
}
// Attention backward of one time step in ONE kernel (dalpha_kernel + softmax_bwd_kernel + dz_kernel).  The softmax backward needs
// s_b = sum_l alpha_l * dalpha_l over the whole image, but dalpha_l = a_l . dctx + dAP_l and sum_l alpha_l a_l is the context vector the
// forward saved, so s_b = ctx_b . dctx_b + sum_l alpha_l dAP_l: a 684 + L term dot product every CTA recomputes - no grid-wide reduction.
// One CTA = 16 positions of one image: warp w owns the annotation dot products of positions 4w..4w+3 (4 independent accumulators, one
// interleaved butterfly), then the dz body as above with the tanh rows already in flight.
template <int PIX> __global__ void __launch_bounds__(128)
attn_bwd_kernel(const float* __restrict__ ann, const float* __restrict__ dctx, const float* __restrict__ ctx,
                const float* __restrict__ alpha, const float* __restrict__ dAP, const float* __restrict__ TANH,
                const float* __restrict__ Va, float* __restrict__ DZ, float* __restrict__ dSP, float* __restrict__ dVa,
                float* __restrict__ dVab, int L, float4* __restrict__ zero, int zero_n4) {
  pdl_sync();
  constexpr int PW = PIX / 4;           // positions per warp
  __shared__ __align__(16) float s_d[kD];
  __shared__ float red[33];
  __shared__ float s_de[PIX];
  const int b = blockIdx.y, l0 = blockIdx.x * PIX, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = tid * 4;
  float4 t[PIX];
#pragma unroll
  for (int j = 0; j < PIX; ++j) t[j] = ldg4(TANH + ((size_t)b * L + min(l0 + j, L - 1)) * kNA + c);
  float s = 0.f;
  for (int i = tid; i < kD; i += 128) { const float d = dctx[(size_t)b * kD + i]; s_d[i] = d; s = fmaf(ctx[(size_t)b * kD + i], d, s); }
  for (int l = tid; l < L; l += 128) s = fmaf(alpha[(size_t)b * L + l], dAP[(size_t)b * L + l], s);
  s = block_sum(s, red);
  // split-K coverage-backward GEMM: this step's kernel clears the DF buffer the step after next accumulates into (see the BPTT loop)
  for (int i = (blockIdx.y * gridDim.x + blockIdx.x) * 128 + tid; i < zero_n4; i += gridDim.x * gridDim.y * 128) zero[i] = make4(0.f);
  float acc[PW];
  const float* arow[PW];
#pragma unroll
  for (int q = 0; q < PW; ++q) { acc[q] = 0.f; arow[q] = ann + ((size_t)b * L + min(l0 + warp * PW + q, L - 1)) * kD; }
  for (int i = lane * 4; i < kD; i += 128) {
    const float4 dv = *reinterpret_cast<const float4*>(s_d + i);
#pragma unroll
    for (int q = 0; q < PW; ++q) {
      const float4 av = ldg4(arow[q] + i);
      acc[q] = fmaf(av.x, dv.x, acc[q]); acc[q] = fmaf(av.y, dv.y, acc[q]); acc[q] = fmaf(av.z, dv.z, acc[q]); acc[q] = fmaf(av.w, dv.w, acc[q]);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int q = 0; q < PW; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], off);
  if (lane < PW) {
    const int l = l0 + warp * PW + lane;
    float da = acc[0];
#pragma unroll
    for (int q = 1; q < PW; ++q) if (lane == q) da = acc[q];
    s_de[warp * PW + lane] = l < L ? alpha[(size_t)b * L + l] * (da + dAP[(size_t)b * L + l] - s) : 0.f;
  }
  __syncthreads();
  const float4 va = ldg4(Va + c);
  float4 sp = make4(0.f), dv = make4(0.f);
  float sde = 0.f;
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    if (l0 + j >= L) break;
    const size_t row = (size_t)b * L + l0 + j;
    const float d = s_de[j];
    float4 z;
    z.x = d * va.x * (1.f - t[j].x * t[j].x); z.y = d * va.y * (1.f - t[j].y * t[j].y);
    z.z = d * va.z * (1.f - t[j].z * t[j].z); z.w = d * va.w * (1.f - t[j].w * t[j].w);
    *reinterpret_cast<float4*>(DZ + row * kNA + c) = z;
    sp.x += z.x; sp.y += z.y; sp.z += z.z; sp.w += z.w;
    dv.x = fmaf(d, t[j].x, dv.x); dv.y = fmaf(d, t[j].y, dv.y); dv.z = fmaf(d, t[j].z, dv.z); dv.w = fmaf(d, t[j].w, dv.w);
    sde += d;
  }
  float* o = dSP + (size_t)b * kNA + c;
  atomicAdd(o, sp.x); atomicAdd(o + 1, sp.y); atomicAdd(o + 2, sp.z); atomicAdd(o + 3, sp.w);
  float* dva = dVa + (size_t)b * kNA + c;          // per-image partial rows (column-summed after the loop)
  atomicAdd(dva, dv.x); atomicAdd(dva + 1, dv.y); atomicAdd(dva + 2, dv.z); atomicAdd(dva + 3, dv.w);
  if (tid == 0) atomicAdd(dVab + b, sde);
}
// dWatch[i] = sum_t DZ[t][i]   (one pass after the scan instead of a read-modify-write of dWatch every step)
__global__ void __launch_bounds__(256)
sum_steps_kernel(const float* __restrict__ DZ, float* __restrict__ out, int T, size_t n4) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 acc = make4(0.f);
  int t = 0;
  for (; t + 8 <= T; t += 8) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ldg4(DZ + ((size_t)(t + j) * n4 + i) * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  for (; t < T; ++t) { const float4 v = ldg4(DZ + ((size_t)t * n4 + i) * 4); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
  *reinterpret_cast<float4*>(out + i * 4) = acc;
}
// dAP_out[b,l'] = dAP_in[b,l'] + sum_tap DF[b, l'-off(tap), tap]   (transposed 11x11 stencil).  One CTA = 32 consecutive
// positions x 16 warps; warp w sums taps w, w+16, ... (8 independent loads per thread), then a 16-way smem reduction.
constexpr int kGatherWarps = 16;
__global__ void __launch_bounds__(32 * kGatherWarps)
cov_bwd_gather_kernel(const float* __restrict__ DF, const float* __restrict__ dAP_in, float* __restrict__ dAP_out,
                      int Bn, int H, int W, int tap_major) {
  pdl_sync();
  __shared__ float part[kGatherWarps][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane, n = Bn * H * W;
  float s = 0.f;
  if (i < n) {
    const int x = i % W, t = i / W, y = t % H, b = t / H;
    constexpr int kIt = (kCovTaps + kGatherWarps - 1) / kGatherWarps;
    float v[kIt];
#pragma unroll
    for (int j = 0; j < kIt; ++j) {
      const int tap = warp + j * kGatherWarps;
      const int dy = tap / kCovK - kCovK / 2, dx = tap % kCovK - kCovK / 2;
      const int ys = y - dy, xs = x - dx;     // source position l with l + off(tap) = l'
      const bool ok = tap < kCovTaps && (unsigned)ys < (unsigned)H && (unsigned)xs < (unsigned)W;
      const size_t src = (size_t)(b * H + (ok ? ys : 0)) * W + (ok ? xs : 0);
      const float val = tap_major ? DF[(size_t)(ok ? tap : 0) * n + src] : DF[src * 124 + (ok ? tap : 0)];   // tap-major: coalesced across x
      v[j] = ok ? val : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kIt; ++j) s += v[j];
  }
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && i < n) {
    float a = dAP_in[i];
#pragma unroll
    for (int k = 0; k < kGatherWarps; ++k) a += part[k][lane];
    dAP_out[i] = a;
  }
}
// d ann[b,l,c] += sum_t alpha[t,b,l] * dctx[t,b,c]
__global__ void __launch_bounds__(256)
dann_ctx_kernel(const float* __restrict__ ALPHA, const float* __restrict__ DCTX, float* __restrict__ dann, int T, int B, int L) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int C4 = kD / 4;
  if (i >= (size_t)B * L * C4) return;
  const int c = (int)(i % C4) * 4; const size_t row = i / C4; const int b = (int)(row / L);
  float4 acc = *reinterpret_cast<float4*>(dann + row * kD + c);
  for (int t = 0; t < T; ++t) {
    const float al = ALPHA[(size_t)t * B * L + row];
    const float4 d = ldg4(DCTX + ((size_t)t * B + b) * kD + c);
    acc.x = fmaf(al, d.x, acc.x); acc.y = fmaf(al, d.y, acc.y); acc.z = fmaf(al, d.z, acc.z); acc.w = fmaf(al, d.w, acc.w);
  }
  *reinterpret_cast<float4*>(dann + row * kD + c) = acc;
}
// h0 = tanh(mean*Wa2h + b): dpre = dh * (1 - h0^2)
__global__ void tanh_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ h, float* __restrict__ dpre, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dpre[i] = dh[i] * (1.f - h[i] * h[i]);
}
// d ann[b,l,:] += mask[b,l]/sum(mask[b]) * dmean[b,:]
__global__ void __launch_bounds__(256)
dann_mean_kernel(const float* __restrict__ dmean, const float* __restrict__ a_mask, float* __restrict__ dann, int B, int L) {
  pdl_sync();
  __shared__ float red[33];
  const int b = blockIdx.y;
  float s = 0.f;
  for (int l = threadIdx.x; l < L; l += 256) s += a_mask != nullptr ? a_mask[(size_t)b * L + l] : 1.f;
  s = block_sum(s, red);
  const int C4 = kD / 4;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < L * C4; i += gridDim.x * 256) {
    const int c = (i % C4) * 4, l = i / C4;
    const float m = (a_mask != nullptr ? a_mask[(size_t)b * L + l] : 1.f) / s;
    if (m == 0.f) continue;
    const float4 d = ldg4(dmean + (size_t)b * kD + c);
    float4* o = reinterpret_cast<float4*>(dann + ((size_t)b * L + l) * kD + c);
    float4 v = *o; v.x = fmaf(m, d.x, v.x); v.y = fmaf(m, d.y, v.y); v.z = fmaf(m, d.z, v.z); v.w = fmaf(m, d.w, v.w); *o = v;
  }
}

// ------------------------------------------------------------------------------------ backward
// grads (flat) accumulated (+=); d_ann [B,L,684] is written (=).
// the caller's stream waits for the decoder weight gradients still running on the side stream
int decoder_join(cudaStream_t st) {
  if (cur()->dec_join_pending) {
    SideStreams* ss = side_streams();
    if (ss != nullptr) WAP_CHECK(cudaStreamWaitEvent(st, ss->done[5], 0));
    cur()->dec_join_pending = false;
  }
  return 0;
}
int decoder_train_backward(cudaStream_t st, const float* params, float* grads, const float* ann, const float* a_mask,
                           const int* y, const float* y_mask, int B, int h, int w, int T, int dropout, uint32_t seed,
                           float* d_ann, float* wsp, bool defer_join) {
  const ParamLayout& P = layout();
  const int L = h * w, TB = T * B;
  DecTrainWs W = carve_train(wsp, B, L, T, nullptr);
  const float* p = params; float* g = grads;
  // transposed copies for the in-loop data gradients
  {
    TransposeBatch tb; tb.n = 7;
    const float* src[7] = {p + P.W_c_z_r, p + P.W_c_h_nl, p + P.U_hz_hr_nl, p + P.U_rh_nl, p + P.W_a, p + P.U_hz_hr, p + P.U_rh};
    float* dst[7] = {W.T_Wczr, W.T_Wchnl, W.T_Uhzhrnl, W.T_Urhnl, W.T_Wa, W.T_Uhzhr, W.T_Urh};
    const int R[7] = {kD, kD, kN, kN, kN, kN, kN}, C[7] = {2 * kN, kN, 2 * kN, kN, kNA, 2 * kN, kN};
    for (int i = 0; i < 7; ++i) { tb.src[i] = src[i]; tb.dst[i] = dst[i]; tb.R[i] = R[i]; tb.C[i] = C[i]; }
    WAP_TRY(transpose_multi(st, tb));
  }
  const int prec = get_conv_precision();
  const bool tc = prec != 0;
  WAP_REQUIRE(!tc || tc_available(), "tensor-core conv precision requested but cuTensorMapEncodeTiled is unavailable (no silent FFMA fallback)");
  if (tc) {
    TcPrepBatch pb; pb.n = 5;
    pb.e[0] = {W.qf, kNA, kCovTaps, 0, W.qf_hi, W.qf_lo};
    pb.e[1] = {p + P.U_a, kNA, kD, 0, W.Ua_hi, W.Ua_lo};
    pb.e[2] = {p + P.Wh, kMw, kN, 0, W.WhN[0], W.WhN[1]};
    pb.e[3] = {p + P.Wc, kMw, kD, 0, W.WcN[0], W.WcN[1]};
    pb.e[4] = {p + P.Wy, kMw, kMw, 0, W.WyN[0], W.WyN[1]};
    WAP_TRY(tc_prep_multi(st, pb));
  }
  // ---- head (LO holds dLO after the forward CE kernel)
  WAP_CHECK(launch_pdl(wo_bwd_data_kernel, cdiv(TB * kMw / 2, 256), 256, 0, st, W.LO, p + P.Wo, W.dMO, TB));
  WAP_LAUNCH_CHECK();
  WAP_CHECK(launch_pdl(maxout_drop_bwd_kernel, cdiv(TB * kMw / 2, 256), 256, 0, st, W.LG, W.dMO, W.dLG, TB, dropout, seed, seed_src()));
  WAP_LAUNCH_CHECK();
  {   // weight gradients of the logit layer: side stream (joined together with the tail, see below)
    SideStreams* sh = overlap_enabled() ? side_streams() : nullptr;
    cudaStream_t swh = sh != nullptr ? sh->s[1] : st;
    if (sh != nullptr) { WAP_CHECK(cudaEventRecord(sh->fork[1], st)); WAP_CHECK(cudaStreamWaitEvent(swh, sh->fork[1], 0)); }
    WAP_CHECK(launch_pdl(wo_bwd_weight_kernel, cdiv((kMw / 2 + 1) * kVocab, 256), 256, 0, swh, W.MO, W.LO, g + P.Wo, g + P.bo, TB));
    WAP_LAUNCH_CHECK();
    if (tc)      // d EMB (logit path) is only read after the BPTT (embedding scatter on this stream): keep it off the main chain
      WAP_TRY(tc_gemm(swh, W.dLG, kMw, W.WyN[0], prec >= 2 ? W.WyN[1] : nullptr, kMw, TB, kMw, kMw, TcEpi{W.dEMB, kMw, 0, nullptr, nullptr, 0}, kProfOtherGemm));
    WAP_TRY(gemm_tn_atomic(swh, W.H + (size_t)B * kN, kN, W.dLG, kMw, g + P.Wh, kMw, kN, kMw, TB));
    WAP_TRY(gemm_tn_atomic(swh, W.CTX, kD, W.dLG, kMw, g + P.Wc, kMw, kD, kMw, TB));
    WAP_TRY(gemm_tn_atomic(swh, W.EMB, kMw, W.dLG, kMw, g + P.Wy, kMw, kMw, kMw, TB));
    WAP_TRY(colsum(swh, W.dLG, kMw, TB, kMw, g + P.bh));
    WAP_TRY(colsum(swh, W.dLG, kMw, TB, kMw, g + P.bc));
    WAP_TRY(colsum(swh, W.dLG, kMw, TB, kMw, g + P.by));
  }
  if (tc) {     // data gradients of the logit layer on tensor cores (the FFMA tiles took 30-60 us each at M = T*B rows, in front of the BPTT)
    WAP_TRY(tc_gemm(st, W.dLG, kMw, W.WhN[0], prec >= 2 ? W.WhN[1] : nullptr, kMw, TB, kN, kMw, TcEpi{W.dHlog, kN, 0, nullptr, nullptr, 0}, kProfOtherGemm));
    WAP_TRY(tc_gemm(st, W.dLG, kMw, W.WcN[0], prec >= 2 ? W.WcN[1] : nullptr, kMw, TB, kD, kMw, TcEpi{W.dCTX, kD, 0, nullptr, nullptr, 0}, kProfOtherGemm));
  } else {
    WAP_TRY(gemm_nt(st, W.dLG, kMw, p + P.Wh, kMw, W.dHlog, kN, TB, kN, kMw, nullptr, false));
    WAP_TRY(gemm_nt(st, W.dLG, kMw, p + P.Wc, kMw, W.dCTX, kD, TB, kD, kMw, nullptr, false));
    WAP_TRY(gemm_nt(st, W.dLG, kMw, p + P.Wy, kMw, W.dEMB, kMw, TB, kMw, kMw, nullptr, false));
  }
  // ---- BPTT
  WAP_CHECK(cudaMemsetAsync(W.dh0, 0, sizeof(float) * B * kN, st));
  WAP_CHECK(cudaMemsetAsync(W.dAP0, 0, sizeof(float) * B * L, st));
  WAP_CHECK(cudaMemsetAsync(W.dSP, 0, sizeof(float) * TB * kNA, st));
  float* dh_in = W.dh0; float* dh_out = W.dh1; float* dAP_in = W.dAP0; float* dAP_out = W.dAP1;
  SideStreams* sl = overlap_enabled() ? side_streams() : nullptr;
  cudaStream_t s2 = sl != nullptr ? sl->s[2] : st;
  auto fork2 = [&](int e) -> int {
    if (sl != nullptr) { WAP_CHECK(cudaEventRecord(sl->fork[e], st)); WAP_CHECK(cudaStreamWaitEvent(s2, sl->fork[e], 0)); }
    return 0;
  };
  bool pend_ap = false;
  const bool fuse_bptt = (g_fuse_gates & 2) != 0;
  if (fuse_bptt) {
    // Four dependent kernels per step (eight before): the elementwise GRU backward passes are evaluated inside the skinny GEMMs that
    // consume them, dalpha / softmax backward / dz are one kernel.  Every contribution to the carried gradients is an atomic add into
    // per-step rows zero-filled here; DH[T] = 0 is the initial carry.  Side stream: the coverage backward DZ -> DF -> dAP only.
    WAP_CHECK(cudaMemsetAsync(W.DH, 0, sizeof(float) * ((size_t)(TB + B) * kN + (size_t)TB * kN + (size_t)TB * kNA + (((size_t)TB + 3) & ~(size_t)3)), st));     // DH, DPREH, dVaP, dVabP are adjacent
    const int mtiles = cdiv(B * L, 128);
    const int splitk = std::max(1, std::min(8, 148 / mtiles));      // (4 and 16 measured at 16 row tiles: 8 is the fastest; 8 positions per CTA of the attention kernel: slower)
    const size_t BL = (size_t)B * L;
    if (tc && splitk > 1) WAP_CHECK(cudaMemsetAsync(W.DF, 0, sizeof(float) * BL * 124 * 2, st));
    for (int t = T - 1; t >= 0; --t) {
      const size_t r = (size_t)t * B;
      const float* ym = y_mask + r;
      float* dctx = W.dCTX + r * kD;                 // holds the logit-path part already
      float* dpre = W.DPREH + r * kN;
      {   // [dG2|dCH] -> dctx and [dPH2|dG2] -> dpre_h in one launch; X = the GRU-2 backward of (DH[t+1] + dHlog[t])
        SkinnyArgs a; a.X = nullptr; a.ldx = 0; a.ids = nullptr; a.Bn = B; a.K = 4 * kN; a.nseg = 2;
        a.seg[0] = {W.T_Wczr, kD, kD, dctx, kD, kN, 3 * kN};
        a.seg[1] = {W.T_Urhnl, kN, kN, dpre, kN, 0, 3 * kN};
        a.gen = 3; a.gA = W.DH + (r + B) * kN; a.gB = W.dHlog + r * kN; a.gG = W.RZ2 + r * 2 * kN; a.gC = W.PROP2 + r * kN;
        a.gD = W.PREH + r * kN; a.gE = W.PH2 + r * kN; a.gmask = ym; a.oX = W.dPH2 + r * kLdA; a.oAcc = dpre;
        WAP_TRY(skinny_gemm(st, a));
      }
      if (sl != nullptr && pend_ap) WAP_CHECK(cudaStreamWaitEvent(st, sl->done[7], 0));     // dAP_in of this step (side-stream gather of step t+1)
      float* DZt = W.DZ + r * L * kNA;
      // coverage backward GEMM DF^T[121][B*L] = (DZ_t * Q'^T)^T: few output tiles, long reduction and on the recurrence cycle
      // (attn_bwd -> DF -> gather -> attn_bwd), so it runs split-K with atomics into one of two DF buffers; the attention kernel of this
      // step clears the buffer that step t-1 will use (its last reader, the gather of step t+1, has finished: this kernel waited for it).
      float* DFt = W.DF + (size_t)(t & 1) * BL * 124;
      float* DFnext = W.DF + (size_t)((t + 1) & 1) * BL * 124;
      const int zero_n4 = (tc && splitk > 1) ? (int)(BL * 124 / 4) : 0;
      WAP_CHECK(launch_pdl(attn_bwd_kernel<16>, dim3(cdiv(L, 16), B), 128, 0, st, ann, dctx, W.CTX + r * kD, W.ALPHA + r * L, dAP_in,
                           W.TANH + r * L * kNA, p + P.V_a, DZt, W.dSP + r * kNA, W.dVaP + r * kNA, W.dVabP + r, L, reinterpret_cast<float4*>(DFnext), zero_n4));
      WAP_LAUNCH_CHECK();
      WAP_TRY(fork2(3));
      if (tc) WAP_TRY(tc_gemm(s2, DZt, kNA, W.qf_hi, prec >= 2 ? W.qf_lo : nullptr, kNA, B * L, kCovTaps, kNA, TcEpi{DFt, B * L, 0, nullptr, nullptr, 1}, kProfOtherGemm, splitk));
      else WAP_TRY(gemm_nt(s2, DZt, kNA, W.qf, kNA, DFt, 124, B * L, kCovTaps, kNA, nullptr, false));
      WAP_CHECK(launch_pdl(cov_bwd_gather_kernel, cdiv(B * L, 32), 32 * kGatherWarps, 0, s2, DFt, dAP_in, dAP_out, B, h, w, tc ? 1 : 0));
      WAP_LAUNCH_CHECK();
      if (sl != nullptr) { WAP_CHECK(cudaEventRecord(sl->done[7], s2)); pend_ap = true; }
      {   // dpre_h += dSP*W_a^T
        SkinnyArgs a; a.ids = nullptr; a.Bn = B; a.X = W.dSP + r * kNA; a.ldx = kNA; a.K = kNA; a.nseg = 1;
        a.seg[0] = {W.T_Wa, kN, kN, dpre, kN};
        WAP_TRY(skinny_gemm(st, a));
      }
      {   // DH[t] = direct term + [dG1|dUH] * [U_hz_hr^T; U_rh^T]; X = the GRU-1 backward of dpre_h
        SkinnyArgs a; a.X = nullptr; a.ldx = 0; a.ids = nullptr; a.Bn = B; a.K = 3 * kN; a.nseg = 1;
        a.seg[0] = {W.T_Uhzhr, kN, kN, W.DH + r * kN, kN};
        a.gen = 4; a.gA = dpre; a.gG = W.RZ1 + r * 2 * kN; a.gC = W.PROP1 + r * kN; a.gD = W.UH + r * kN; a.gE = W.H + r * kN; a.gmask = ym;
        a.oX = W.dG1 + r * kLdB; a.oProp = W.dEh + r * kN; a.oAcc = W.DH + r * kN;
        WAP_TRY(skinny_gemm(st, a));
      }
      std::swap(dAP_in, dAP_out);
    }
    dh_in = W.DH;          // d h0
  } else {
  // Two chains per step.  Main stream: gru2_bwd -> [dG2|dCH] -> dctx -> dalpha -> softmax_bwd -> dz -> dSP -> dpre_h -> gru1_bwd ->
  // [dG1|dUH] -> dh.  Side stream: [dPH2|dG2] -> dpre_h (joins before gru1_bwd) and the coverage backward DZ -> DF -> dAP (joins before
  // the next step's dalpha).
  for (int t = T - 1; t >= 0; --t) {
    const size_t r = (size_t)t * B;
    const float* ym = y_mask + r;
    float* dctx = W.dCTX + r * kD;                 // holds the logit-path part already
    WAP_CHECK(launch_pdl(gru2_bwd_kernel, cdiv(B * kN, 256), 256, 0, st, dh_in, W.dHlog + r * kN, W.RZ2 + r * 2 * kN, W.PROP2 + r * kN, W.PH2 + r * kN,
                         W.PREH + r * kN, ym, W.dG2 + r * kLdA, W.dCH + r * kLdA, W.dPH2 + r * kLdA, W.dpre_h, B));
    WAP_LAUNCH_CHECK();
    WAP_TRY(fork2(2));
    {   // side: dpre_h += [dPH2|dG2] * [U_rh_nl^T; U_hz_hr_nl^T]
      SkinnyArgs a; a.ids = nullptr; a.Bn = B; a.X = W.dPH2 + r * kLdA; a.ldx = kLdA; a.K = 3 * kN; a.nseg = 1;
      a.seg[0] = {W.T_Urhnl, kN, kN, W.dpre_h, kN};
      WAP_TRY(skinny_gemm(s2, a));
      if (sl != nullptr) WAP_CHECK(cudaEventRecord(sl->done[6], s2));
    }
    {   // dctx += [dG2|dCH] * [W_c_z_r^T; W_c_h_nl^T]
      SkinnyArgs a; a.ids = nullptr; a.Bn = B; a.X = W.dG2 + r * kLdA; a.ldx = kLdA; a.K = 3 * kN; a.nseg = 1;
      a.seg[0] = {W.T_Wczr, kD, kD, dctx, kD};
      WAP_TRY(skinny_gemm(st, a));
    }
    if (sl != nullptr && pend_ap) WAP_CHECK(cudaStreamWaitEvent(st, sl->done[7], 0));     // dAP_in of this step (side-stream gather of step t+1)
    WAP_CHECK(launch_pdl(dalpha_kernel, cdiv(B * L * 32, 256), 256, 0, st, ann, dctx, dAP_in, W.dalpha, B, L));
    WAP_LAUNCH_CHECK();
    WAP_CHECK(launch_pdl(softmax_bwd_kernel, B, 256, 0, st, W.ALPHA + r * L, W.dalpha, W.de, L));
    WAP_LAUNCH_CHECK();
    float* DZt = W.DZ + r * L * kNA;
    WAP_CHECK(launch_pdl(dz_kernel, dim3(cdiv(L, kDzPix), B), 128, 0, st, W.TANH + r * L * kNA, W.de, p + P.V_a, DZt, W.dSP + r * kNA,
                         g + P.V_a, g + P.V_a_b, L));
    WAP_LAUNCH_CHECK();
    WAP_TRY(fork2(3));
    // side: coverage backward: DF[B*L,121] = DZ * Q'^T ; dAP_out = dAP_in + transposed stencil
    if (tc) WAP_TRY(tc_gemm(s2, DZt, kNA, W.qf_hi, prec >= 2 ? W.qf_lo : nullptr, kNA, B * L, kCovTaps, kNA, TcEpi{W.DF, B * L, 0, nullptr, nullptr, 1}, kProfOtherGemm));   // tap-major DF^T [121][B*L]
    else WAP_TRY(gemm_nt(s2, DZt, kNA, W.qf, kNA, W.DF, 124, B * L, kCovTaps, kNA, nullptr, false));
    WAP_CHECK(launch_pdl(cov_bwd_gather_kernel, cdiv(B * L, 32), 32 * kGatherWarps, 0, s2, W.DF, dAP_in, dAP_out, B, h, w, tc ? 1 : 0));
    WAP_LAUNCH_CHECK();
    if (sl != nullptr) { WAP_CHECK(cudaEventRecord(sl->done[7], s2)); pend_ap = true; }
    {   // dpre_h += dSP*W_a^T
      SkinnyArgs a; a.ids = nullptr; a.Bn = B; a.X = W.dSP + r * kNA; a.ldx = kNA; a.K = kNA; a.nseg = 1;
      a.seg[0] = {W.T_Wa, kN, kN, W.dpre_h, kN};
      WAP_TRY(skinny_gemm(st, a));
    }
    if (sl != nullptr) WAP_CHECK(cudaStreamWaitEvent(st, sl->done[6], 0));
    WAP_CHECK(launch_pdl(gru1_bwd_kernel, cdiv(B * kN, 256), 256, 0, st, W.dpre_h, W.RZ1 + r * 2 * kN, W.PROP1 + r * kN, W.UH + r * kN, W.H + r * kN, ym,
                         W.dG1 + r * kLdB, W.dEh + r * kN, W.dUH + r * kLdB, dh_out, B));
    WAP_LAUNCH_CHECK();
    {   // dh_prev += [dG1|dUH] * [U_hz_hr^T; U_rh^T]
      SkinnyArgs a; a.ids = nullptr; a.Bn = B; a.X = W.dG1 + r * kLdB; a.ldx = kLdB; a.K = 3 * kN; a.nseg = 1;
      a.seg[0] = {W.T_Uhzhr, kN, kN, dh_out, kN};
      WAP_TRY(skinny_gemm(st, a));
    }
    std::swap(dh_in, dh_out); std::swap(dAP_in, dAP_out);
  }
  }
  if (sl != nullptr && pend_ap) WAP_CHECK(cudaStreamWaitEvent(st, sl->done[7], 0));
  // watch path: dWatch = sum_t DZ_t
  WAP_CHECK(launch_pdl(sum_steps_kernel, cdiv((size_t)B * L * kNA / 4, 256), 256, 0, st, W.DZ, W.dWatch, T, (size_t)B * L * kNA / 4));
  WAP_LAUNCH_CHECK();
  // Everything below that only feeds `grads` (the batched weight gradients) runs on a side stream next to the d_annotation
  // path and - when the caller defers the join (wap_train_backward) - next to the encoder backward.
  SideStreams* ss = overlap_enabled() ? side_streams() : nullptr;
  cudaStream_t sw = ss != nullptr ? ss->s[1] : st;
  if (ss != nullptr) { WAP_CHECK(cudaEventRecord(ss->fork[1], st)); WAP_CHECK(cudaStreamWaitEvent(sw, ss->fork[1], 0)); }
  // ---- batched weight gradients over all T*B rows
  const float* Hprev = W.H;   // H[0..T-1]
  WAP_TRY(gemm_tn_atomic(sw, W.EMB, kMw, W.dG1, kLdB, g + P.W_yz_yr, 2 * kN, kMw, 2 * kN, TB));
  WAP_TRY(colsum(sw, W.dG1, kLdB, TB, 2 * kN, g + P.b_yz_yr));
  WAP_TRY(gemm_tn_atomic(sw, W.EMB, kMw, W.dEh, kN, g + P.W_yh, kN, kMw, kN, TB));
  WAP_TRY(colsum(sw, W.dEh, kN, TB, kN, g + P.b_yh));
  WAP_TRY(gemm_tn_atomic(sw, Hprev, kN, W.dG1, kLdB, g + P.U_hz_hr, 2 * kN, kN, 2 * kN, TB));
  WAP_TRY(gemm_tn_atomic(sw, Hprev, kN, W.dUH, kLdB, g + P.U_rh, kN, kN, kN, TB));
  WAP_TRY(gemm_tn_atomic(sw, W.PREH, kN, W.dSP, kNA, g + P.W_a, kNA, kN, kNA, TB));
  WAP_TRY(colsum(sw, W.dSP, kNA, TB, kNA, g + P.W_a_b));
  WAP_TRY(colsum(sw, W.dSP, kNA, TB, kNA, g + P.U_f_b));
  WAP_TRY(colsum(sw, W.dSP, kNA, TB, kNA, g + P.U_a_b));
  if (fuse_bptt) { WAP_TRY(colsum(sw, W.dVaP, kNA, TB, kNA, g + P.V_a)); WAP_TRY(colsum(sw, W.dVabP, 1, TB, 1, g + P.V_a_b)); }
  WAP_TRY(gemm_tn_atomic(sw, W.PREH, kN, W.dG2, kLdA, g + P.U_hz_hr_nl, 2 * kN, kN, 2 * kN, TB));
  WAP_TRY(colsum(sw, W.dG2, kLdA, TB, 2 * kN, g + P.b_hz_hr_nl));
  WAP_TRY(gemm_tn_atomic(sw, W.PREH, kN, W.dPH2, kLdA, g + P.U_rh_nl, kN, kN, kN, TB));
  WAP_TRY(colsum(sw, W.dPH2, kLdA, TB, kN, g + P.b_rh_nl));
  WAP_TRY(gemm_tn_atomic(sw, W.CTX, kD, W.dG2, kLdA, g + P.W_c_z_r, 2 * kN, kD, 2 * kN, TB));
  WAP_TRY(gemm_tn_atomic(sw, W.CTX, kD, W.dCH, kLdA, g + P.W_c_h_nl, kN, kD, kN, TB));
  // embedding: dEMB = dLG*Wy^T (already) + dG1*W_yz_yr^T + dEh*W_yh^T ; scatter-add into the table
  WAP_TRY(gemm_nt(sw, W.dG1, kLdB, p + P.W_yz_yr, 2 * kN, W.dEMB, kMw, TB, kMw, 2 * kN, nullptr, true));
  WAP_TRY(gemm_nt(sw, W.dEh, kN, p + P.W_yh, kN, W.dEMB, kMw, TB, kMw, kN, nullptr, true));
  WAP_CHECK(launch_pdl(emb_scatter_kernel, cdiv(TB * kMw, 256), 256, 0, sw, W.dEMB, y, g + P.embed, T, B));
  WAP_LAUNCH_CHECK();
  // fused coverage filter: dQ' = sum_t patches(AP_t)^T * DZ_t  (one GEMM over T*B images)
  {
    WAP_CHECK(cudaMemsetAsync(W.dQf, 0, sizeof(float) * kCovTaps * kNA, sw));
    if (tc) {
      // tensor cores: materialise the patches of every step once (im2col of all T*B alpha_past images), then dQ'[121,512] += P^T * DZ as the
      // weight-gradient GEMM the encoder uses (both operands MN-major, reduction over T*B*L rows split over grid.z), 96 columns per launch.
      // The FFMA form of this GEMM held all 148 SMs for ~0.5 ms next to the latency-bound encoder backward.
      WAP_TRY(cov_im2col(sw, W.AP, W.PpatAll, T * B, h, w));
      for (int n0 = 0; n0 < kNA; n0 += 96)
        WAP_TRY(tc_wgrad_1x1(sw, W.PpatAll, 128, W.DZ + n0, kNA, TB * L, kCovTaps, std::min(96, kNA - n0), W.dQf + n0, kNA, prec >= 2, kProfOtherGemm));
    } else {
    ACovPatchT A{W.AP, T * B, h, w};
    BRowMajor Bd{W.DZ, kNA, TB * L, kNA};
    EpiAtomic e{W.dQf, kNA};
    const int K = ((TB * L + 3) / 4) * 4;
    WAP_TRY((launch_gemm<Cfg128x128>(sw, A, Bd, e, kCovTaps, kNA, K, std::max(1, std::min(cdiv(K, 64), 148)))));
    }
    // dQ += dQ' * U_f^T ; dU_f += Q^T * dQ'
    WAP_TRY(gemm_nt(sw, W.dQf, kNA, p + P.U_f, kNA, g + P.Q, kNA, kCovTaps, kNA, kNA, nullptr, true));
    WAP_TRY(gemm_tn_atomic(sw, p + P.Q, kNA, W.dQf, kNA, g + P.U_f, kNA, kNA, kNA, kCovTaps));
  }
  WAP_TRY(gemm_tn_atomic(sw, ann, kD, W.dWatch, kNA, g + P.U_a, kNA, kD, kNA, B * L));     // dU_a += ann^T * dWatch
  if (ss != nullptr) { WAP_CHECK(cudaEventRecord(ss->done[5], sw)); cur()->dec_join_pending = true; }
  // d ann = dWatch * U_a^T
  if (tc) WAP_TRY(tc_gemm(st, W.dWatch, kNA, W.Ua_hi, prec >= 2 ? W.Ua_lo : nullptr, kNA, B * L, kD, kNA, TcEpi{d_ann, kD, 0, nullptr, nullptr, 0}, kProfOtherGemm));
  else WAP_TRY(gemm_nt(st, W.dWatch, kNA, p + P.U_a, kNA, d_ann, kD, B * L, kD, kNA, nullptr, false));
  // context path
  WAP_CHECK(launch_pdl(dann_ctx_kernel, (unsigned)cdivz((size_t)B * L * (kD / 4), 256), 256, 0, st, W.ALPHA, W.dCTX, d_ann, T, B, L));
  WAP_LAUNCH_CHECK();
  // h0 path: dh_in now holds d h0
  WAP_CHECK(launch_pdl(tanh_bwd_kernel, cdiv(B * kN, 256), 256, 0, st, dh_in, W.H, W.dpre_h, B * kN));
  WAP_LAUNCH_CHECK();
  WAP_TRY(gemm_tn_atomic(st, W.mean, kD, W.dpre_h, kN, g + P.Wa2h, kN, kD, kN, B));
  WAP_TRY(colsum(st, W.dpre_h, kN, B, kN, g + P.ba2h));
  WAP_TRY(gemm_nt(st, W.dpre_h, kN, p + P.Wa2h, kN, W.dmean, kD, B, kD, kN, nullptr, false));
  WAP_CHECK(launch_pdl(dann_mean_kernel, dim3(std::min(cdiv(L * (kD / 4), 256), 64), B), 256, 0, st, W.dmean, a_mask, d_ann, B, L));
  WAP_LAUNCH_CHECK();
  if (!defer_join) WAP_TRY(decoder_join(st));
  return 0;
}

}  // namespace wap
