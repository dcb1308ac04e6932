// Decoder ("Attender" + "Parser" + WAP head) building blocks shared by the inference step,
// the greedy loop and the training scan.  Reference: model-single-GPU.py:166-206 (get_context),
// :260-308 (one_time_step), :337-375 (get_cost), :379-439 (get_word).
#pragma once
#include "gemm.cuh"
#include "model.h"

namespace wap {

// ---------------------------------------------------------------- skinny GEMM (few rows, streamed weights)
//   out_s[b, n] += sum_k X[b,k] * W_s[k,n]      for up to 4 weight "segments" sharing the same X.
// Grid = (column tiles of 256 over all segments) x (K chunks of 16): every CTA streams a distinct
// 16x256 weight block exactly once with 16 independent 128-bit loads per thread in flight, and
// adds its partial products with fp32 atomics into outputs pre-initialised with the bias.
constexpr int kSkRows = 16, kSkKC = 16, kSkCols = 256;
struct SkinnySeg { const float* W; int ldw, N; float* out; int ldo; int koff = 0, kcount = 0; };   // this segment multiplies X[koff : koff + kcount] (kcount 0 = K)
struct SkinnyArgs {
  const float* X; int ldx;      // X[b*ldx + k]   (or gathered rows: table + ids[b]*ldx, zero row if ids[b] < 0)
  const int* ids;
  int Bn, K, nseg;
  SkinnySeg seg[4];
  int tile0[5];
  // X generator (training scan): instead of reading X, every CTA evaluates the GRU gate that produces it for its own 16-wide K
  // chunk (cheap, redundant across column tiles) and the CTAs of column tile 0 also store what the backward pass needs.  Removes
  // the gate kernel from the dependent chain of the time step.
  //   gen 1: X = pre_h  from (G1 raw [ld 3kN], UH, h_prev, mask)          stores r|z -> oRZ [ld 2kN], prop -> oProp, pre_h -> oX
  //   gen 2: X = h_new  from (G2 raw [ld 2kN], PH2, CH, pre_h, mask)       stores r|z -> oRZ,        prop -> oProp, h    -> oX
  // BPTT: the elementwise GRU backward evaluated the same way (X = a row of gate gradients); the direct (non-GEMM) term of the carried
  // gradient is ADDED atomically into oAcc, which the caller zero-fills (its other contributions arrive by atomics as well).
  //   gen 3: X = [dPH2 | dG2 r | dG2 z | dCH] (1024) from (dh_in = gA, dHlog = gB, RZ2 = gG [ld 2kN], prop2 = gC, pre_h = gD, PH2 = gE, mask);
  //          seg[0] must be the dctx segment (koff 256), seg[1] the dpre_h segment (koff 0); X -> oX [ld 4kN]; direct term -> oAcc (dpre_h)
  //   gen 4: X = [dG1 r | dG1 z | dUH] (768) from (dpre_h = gA, RZ1 = gG, prop1 = gC, UH = gD, h_prev = gE, mask);
  //          X -> oX [ld 3kN]; dE_h -> oProp; direct term -> oAcc (dh_prev)
  int gen = 0;
  const float *gG = nullptr, *gA = nullptr, *gB = nullptr, *gC = nullptr, *gD = nullptr, *gE = nullptr, *gmask = nullptr;
  float *oRZ = nullptr, *oProp = nullptr, *oX = nullptr, *oAcc = nullptr;
};
int skinny_gemm(cudaStream_t st, SkinnyArgs a);

// ---------------------------------------------------------------- one decoder time step
struct StepCtx {
  int Bn, h, w, L;
  const float* params;
  const float* ann;     // [Bn,L,684]  (or [1,L,684] when shared)
  const float* watch;   // [Bn,L,512]  (or [1,L,512] when shared): U_a a + U_a_b, hoisted out of the scan
  int shared;           // beam search: all rows share one image
  const float* qf;      // fused coverage filter Q' = Q[121,512] * U_f[512,512]
  const float* a_mask;  // [Bn,L] or null
  const float* y_mask;  // [Bn] or null
  const float* h_prev;  // [Bn,256]
  const float* ap_in;   // [Bn,L]   alpha_past before this step
  // activations (saved per step in training)
  float* G1;            // [Bn,768] in: emb*W_yz_yr+b | emb*W_yh+b ; out: r1 | z1 | (unchanged E_h)
  float* UH;            // [Bn,256] h*U_rh
  float* prop1;         // [Bn,256]
  float* pre_h;         // [Bn,256]
  float* SP;            // [Bn,512] speller
  float* G2;            // [Bn,512] r2 | z2
  float* PH2;           // [Bn,256] pre_h*U_rh_nl + b
  float* CH;            // [Bn,256] ctx*W_c_h_nl
  float* prop2;         // [Bn,256]
  float* COV;           // [Bn*L,512] coverage scratch
  float* TANH;          // [Bn*L,512] tanh(cov+watch+speller) (may alias COV when not training)
  float* E;             // [Bn,L]
  float* alpha;         // [Bn,L]
  float* ap_out;        // [Bn,L]
  float* ctx;           // [Bn,684]
  float* h_new;         // [Bn,256]
  // optional extra segment on the ctx GEMM (inference logits: ctx*Wc -> LG)
  SkinnySeg extra_ctx; int has_extra_ctx;
  // tensor-core coverage GEMM (training scan): patches materialised in Ppat [Bn*L,128], Q'^T hi/lo [512][128]; tc_mode 0 = FFMA
  int tc_mode; float* Ppat; const float* qfT_hi; const float* qfT_lo;
  // training scan: the bias rows of all T steps are filled by one launch before the loop
  int skip_fill = 0;
  // training scan: GRU gates evaluated inside the consuming skinny GEMMs (see SkinnyArgs::gen); r|z go to RZ1/RZ2 (ld 2kN).
  // prev_* describe step t-1 (whose GRU-2 gate produces this step's h_prev = c.h_prev, stored by the generator); null at t = 0.
  int fuse_gates = 0;
  float *RZ1 = nullptr, *RZ2 = nullptr;
  const float *prev_G2 = nullptr, *prev_PH2 = nullptr, *prev_CH = nullptr, *prev_pre_h = nullptr, *prev_mask = nullptr;
  float *prev_RZ2 = nullptr, *prev_prop2 = nullptr;
  // training scan: coverage chain (im2col + GEMM, depends only on ap_in) on a side stream, next to the GRU-1 half of the step
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // ... and one step ahead: alpha_past of step t+1 is final after softmax_ctx of step t, so step t forks the coverage chain of step t+1
  // right there (cov_next) instead of step t+1 forking it after the ctx GEMM of step t; cov_ready = the previous step already did
  int cov_next = 0, cov_ready = 0;
};
int decoder_step(cudaStream_t st, const StepCtx& c);
int gru2_gate(cudaStream_t st, float* G2, const float* PH2, const float* CH, const float* pre_h, const float* y_mask, float* prop2, float* h_new,
              int Bn, float* rz_out);
int attention_tail(cudaStream_t st, const float* p, const float* ann, const float* watch, const float* a_mask, int shared, int Bn, int L,
                   const float* SP, const float* ap_in, float* COV, float* TANH, float* E, float* alpha, float* ap_out, float* ctx);
int attention_core(cudaStream_t st, const float* p, const float* ann, const float* watch, const float* qf, const float* a_mask, int shared,
                   int Bn, int h, int w, const float* SP, const float* ap_in, float* COV, float* TANH, float* E, float* alpha,
                   float* ap_out, float* ctx);

// fills up to 8 row-broadcast segments: dst[b*ld + j] = bias ? bias[j] : 0
struct FillSeg { float* dst; int ld, n; const float* bias; };
struct FillArgs { int rows, nseg; FillSeg seg[8]; };
int fill_rows(cudaStream_t st, const FillArgs& a);

}  // namespace wap
