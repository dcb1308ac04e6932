// Persistent greedy decoder: the WHOLE token loop of WAP.get_word / get_sample(argmax)
// (model-single-GPU.py:379-439, :461-483) in ONE cooperative kernel launch.
//
// One CTA per SM.  Every decoder weight (6.5 MB fp32) is sliced across the CTAs and stays RESIDENT in
// shared memory for the whole sequence: each CTA owns <= 2 GRU hidden units (their r/z/candidate
// columns of all six GRU matrices), <= 4 speller columns of W_a, one maxout pair of the logit layer,
// and a [121 taps x 128 channels] slice of the fused coverage filter Q' = Q*U_f.  A time step is six
// phases separated by a grid-wide barrier (atomic counter + acquire spin; co-residency is guaranteed
// by the cooperative launch):
//   A  GRU-1 gates for the owned units              (needs y_{t-1}, h_{t-1})      -> pre_h
//   B  speller columns, GRU-2 pre_h-side products   (needs pre_h)                 -> SP (global), smem
//   C  coverage stencil + tanh + V_a dot            (CTA = position group x 128-channel group) -> e += partial
//   D  masked softmax, alpha_past update, context   (CTA = image x channel chunk) -> ctx
//   E  GRU-2 for the owned units; emb/ctx part of the owned logit pair           -> h_t
//   F  h part of the logit pair, maxout                                          -> MO
//   G  (no barrier) every CTA redundantly evaluates Wo, argmax and the stop rule, so the next token and
//      the termination decision are known CTA-locally: no host round trip, no extra barrier.
// Per step the only HBM/L2 traffic is the annotation (context), the hoisted watch tensor, alpha_past and
// a few KB of vectors: the algorithmic bytes of SURVEY 8d minus the (now on-chip) weights.
//
// BEAM MODE (beam_k > 0) = WAP.get_sample(stochastic=False) (model-single-GPU.py:442-542) + the length-normalised pick
// (:671-672) for ONE image: the rows are the live hypotheses, all sharing the image's annotation / watch tensor (the reference
// tiles ctx0, S:463).  Phase G then does the reference's numpy bookkeeping on the device, redundantly in every CTA so that all of
// them agree on the next rows without another barrier: cand = hyp_scores - log softmax (S:485-487), the (k - dead_k) smallest of the
// live_k x 111 candidates in rank order (S:488-494), finished hypotheses (word 0) moved to the sample list, the others become
// the new rows (S:500-519).  States are not copied: row i reads h / alpha_past of its parent through `src` from ping-pong buffers.
// Token lists are back-pointers in shared memory; CTA 0 unrolls them once at the end.  The host sees nothing until then.
#include <algorithm>

#include "decoder.cuh"

namespace wap {

int attn_precompute(cudaStream_t st, const float* params, const float* ann, const float* a_mask, int B, int h, int w,
                    float* watch, float* h0, float* qf, float* mean_scratch);

namespace {

constexpr int kPT = 256;                  // threads per CTA
constexpr int kMaxB = 16;                 // sequences per launch
constexpr int kMaxL = 2048;               // annotation positions per image supported by the smem carve-up
constexpr int kPadMax = 6144;             // floats of the zero-bordered alpha_past image (H+10)*(W+10) staged in smem
constexpr int kCG = 4, kCGW = kNA / kCG;  // coverage channel groups: 4 x 128
constexpr int kMinCtas = 128;             // ownership scheme needs >= 128 CTAs (<= 2 hidden units per CTA)
constexpr int kBeamMaxLen = 128;          // beam mode: token positions kept as back-pointers in shared memory
// beam bookkeeping in shared memory (floats): src[16] | hs[16] | cand cost [16][16] | cand idx [16][16] | sel cost[16] | sel idx[16] |
// dead {t, parent, score}[16] x 3 | counters[8] | back-pointers + words (2 x kBeamMaxLen x 16 bytes)
constexpr int kBeamSmemFloats = 16 + 16 + 256 + 256 + 16 + 16 + 48 + 8 + 2 * kBeamMaxLen * 16 / 4;

struct DecOff {   // offsets (floats) into the flat parameter buffer
  int U_f_b, W_a, W_a_b, V_a, V_a_b, W_yz_yr, b_yz_yr, U_hz_hr, W_yh, b_yh, U_rh, U_hz_hr_nl, b_hz_hr_nl, W_c_z_r, U_rh_nl,
      b_rh_nl, W_c_h_nl, embed, Wc, bc, Wh, bh, Wy, by, Wo, bo;
};

struct PersistArgs {
  const float* params; const float* ann; const float* watch; const float* qf; const float* h0;
  DecOff o;
  int Bn, H, W, L, maxlen, ignore_eol;
  int beam_k;       // 0: greedy over Bn images; > 0: beam search over ONE image with beam_k (= Bn) hypothesis rows
  float* bscore; int* bmeta;   // beam outputs: scores [beam_k]; {n_samples, best, steps_run, live_at_end}
  // global scratch (cross-CTA); read with ld.global.cg after a barrier
  float* hbuf;      // [2][Bn][256]  ping-pong hidden state
  float* pre_h;     // [Bn][256]
  float* SP;        // [Bn][512]
  float* E;         // [2][Bn][L]    energies (ping-pong: the buffer of step t+1 is zeroed during step t)
  float* ap;        // [2][Bn][L]    alpha_past (ping-pong: step t reads buffer (t-1)&1 through `src`, writes buffer t&1)
  float* ctx;       // [Bn][684]
  float* MO;        // [Bn][128]
  float* LOg;       // [2][Bn][112]  output logits accumulated by the pair owners (ping-pong, pre-set to bo)
  unsigned* barrier;
  int* tokens; int* lengths;   // [Bn][maxlen], [Bn]
  int* steps_run;
  long long* prof;   // optional [8]: accumulated SM cycles of CTA 0 per phase (A,B,C,D,E,F,G) + setup
};

// Grid-wide barrier: one monotonically increasing counter (RED.ADD by thread 0 of every CTA, acquire spin).
// Co-residency of all CTAs is guaranteed by the cooperative launch.  (A flag-per-CTA variant polled by a
// warp measured ~2 us SLOWER per barrier on B200: 148 acquire loads per waiter.)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned nblocks, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned target = epoch * nblocks;
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_dot(const float* x, const float* w, int n, int lane) {
  float s = 0.f;
  for (int k = lane; k < n; k += 32) s = fmaf(x[k], w[k], s);
  return warp_sum(s);
}
// N independent partial sums reduced across the warp with ONE interleaved butterfly (N sequential warp_sum chains are latency-bound)
template <int N> __device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
}
// column `col` of a row-major [K][N] matrix -> contiguous smem
__device__ __forceinline__ void load_col(float* dst, const float* W, int K, int N, int col) {
  for (int k = threadIdx.x; k < K; k += kPT) dst[k] = __ldg(W + (size_t)k * N + col);
}

// cross-CTA vector -> smem with independent 128-bit L2 loads (n4 float4's, both pointers 16-byte aligned)
__device__ __forceinline__ void stage_cg(float* dst, const float* src, int n4) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
  for (int i = threadIdx.x; i < n4; i += kPT) d4[i] = __ldcg(s4 + i);
}

constexpr int kSmemFloats = kCovTaps * kCGW + 2 * 6 * kN + 2 * 3 * kN + 2 * 3 * kD + 4 * kN + 2 * kMw + 2 * kD + 2 * kN +
                            kMaxB * kD + 2 * kMaxB * 8 + kPadMax + 64 + 16 * 64 + kMaxB * 112 + 2 * kMaxB + 2 * 112 +
                            kBeamSmemFloats;

__global__ void __launch_bounds__(kPT, 1) greedy_persistent_kernel(const PersistArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int cta = blockIdx.x, nC = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Bn = a.Bn, L = a.L;
  const float* P = a.params; const DecOff& o = a.o;
  // ---- shared memory carve-up
  float* Qs = sm;                                   // [121][128]   coverage filter slice
  float* WA = Qs + kCovTaps * kCGW;                 // [2][6][256]  GRU-1 columns of the owned units
  float* WB = WA + 2 * 6 * kN;                      // [2][3][256]  GRU-2 pre_h-side columns
  float* WE = WB + 2 * 3 * kN;                      // [2][3][684]  GRU-2 context-side columns
  float* WS = WE + 2 * 3 * kD;                      // [4][256]     speller columns
  float* WPy = WS + 4 * kN;                         // [2][256]     logit pair: Wy
  float* WPc = WPy + 2 * kMw;                       // [2][684]                 Wc
  float* WPh = WPc + 2 * kD;                        // [2][256]                 Wh
  float* X = WPh + 2 * kN;                          // [kMaxB][684] staging of the phase input vectors
  float* own = X + kMaxB * kD;                      // [2][kMaxB][8]: pre_h, g2r, g2z, ph2   (per owned unit)
  float* sal = own + 2 * kMaxB * 8;                 // [kPadMax] zero-bordered alpha_past image (phase C) / alpha row (phase D)
  float* red = sal + kPadMax;                       // [64]
  float* part = red + 64;                           // [16][64]
  float* LO = part + 16 * 64;                       // [kMaxB][112]
  int* sy = reinterpret_cast<int*>(LO + kMaxB * 112);   // [kMaxB] current token
  int* sfin = sy + kMaxB;                           // [kMaxB] finished flag
  float* WoR = reinterpret_cast<float*>(sfin + kMaxB);  // [112] row `cta` of Wo (the maxout unit this CTA owns)
  int* src = reinterpret_cast<int*>(WoR + 2 * 112);     // [16] row whose h / alpha_past of the previous step this row continues
  float* hs = reinterpret_cast<float*>(src + 16);       // [16] hyp_scores (S:454)
  float* cc = hs + 16;                                  // [16][16] per-row candidate costs (ascending)
  int* ci = reinterpret_cast<int*>(cc + 256);           // [16][16] their flat indices row * 111 + word
  float* selc = reinterpret_cast<float*>(ci + 256);     // [16] merged candidates in rank order
  int* seli = reinterpret_cast<int*>(selc + 16);
  int* dead_t = seli + 16; int* dead_p = dead_t + 16; float* dead_s = reinterpret_cast<float*>(dead_p + 16);   // finished hypotheses (S:505-508)
  int* cnt = reinterpret_cast<int*>(dead_s + 16);       // [0] live_k, [1] dead_k
  unsigned char* bk = reinterpret_cast<unsigned char*>(cnt + 8);   // [kBeamMaxLen][16] parent row of live row i at step t
  unsigned char* wd = bk + kBeamMaxLen * 16;                         // [kBeamMaxLen][16] its word
  const bool beam = a.beam_k > 0;
  const int shared_img = beam ? 1 : 0;                  // every row reads image 0's annotation / watch / h0

  // ---- ownership
  const int nu = (cta < kN) ? 1 + ((cta + nC) < kN ? 1 : 0) : 0;       // hidden units cta, cta+nC
  const int unit[2] = {cta, cta + nC};
  int nsp = 0; for (int c = cta; c < kNA; c += nC) ++nsp;               // speller columns cta + i*nC
  const bool has_pair = cta < kMw / 2;                                  // logit columns 2*cta, 2*cta+1
  const int cg = cta % kCG, pg = cta / kCG, npg = nC / kCG;             // coverage: channel group, position group
  const int npos = Bn * L;
  // position groups never span two images: gpi groups per image (one alpha_past staging, one im2col pass per <= 64 positions)
  const int gpi = max(1, npg / Bn), cimg = pg / gpi, cgrp = pg % gpi;
  const bool c_active = cimg < Bn;
  const int cpb = max(1, nC / Bn);                                      // context: CTAs per image
  const int db = cta / cpb, dchunk = cta % cpb;                         // image / chunk (only if db < Bn)
  const int c4_per = (kD / 4 + cpb - 1) / cpb;                          // float4 columns per chunk

  // ---- one-off: slice the weights into shared memory
  for (int i = tid; i < kCovTaps * kCGW; i += kPT) Qs[i] = __ldg(a.qf + (size_t)(i / kCGW) * kNA + cg * kCGW + (i % kCGW));
  for (int s = 0; s < nu; ++s) {
    const int j = unit[s];
    load_col(WA + (s * 6 + 0) * kN, P + o.W_yz_yr, kMw, 2 * kN, j);
    load_col(WA + (s * 6 + 1) * kN, P + o.W_yz_yr, kMw, 2 * kN, kN + j);
    load_col(WA + (s * 6 + 2) * kN, P + o.W_yh, kMw, kN, j);
    load_col(WA + (s * 6 + 3) * kN, P + o.U_hz_hr, kN, 2 * kN, j);
    load_col(WA + (s * 6 + 4) * kN, P + o.U_hz_hr, kN, 2 * kN, kN + j);
    load_col(WA + (s * 6 + 5) * kN, P + o.U_rh, kN, kN, j);
    load_col(WB + (s * 3 + 0) * kN, P + o.U_hz_hr_nl, kN, 2 * kN, j);
    load_col(WB + (s * 3 + 1) * kN, P + o.U_hz_hr_nl, kN, 2 * kN, kN + j);
    load_col(WB + (s * 3 + 2) * kN, P + o.U_rh_nl, kN, kN, j);
    load_col(WE + (s * 3 + 0) * kD, P + o.W_c_z_r, kD, 2 * kN, j);
    load_col(WE + (s * 3 + 1) * kD, P + o.W_c_z_r, kD, 2 * kN, kN + j);
    load_col(WE + (s * 3 + 2) * kD, P + o.W_c_h_nl, kD, kN, j);
  }
  for (int i = 0; i < nsp; ++i) load_col(WS + i * kN, P + o.W_a, kN, kNA, cta + i * nC);
  if (has_pair)
    for (int c = 0; c < 2; ++c) {
      load_col(WPy + c * kMw, P + o.Wy, kMw, kMw, 2 * cta + c);
      load_col(WPc + c * kD, P + o.Wc, kD, kMw, 2 * cta + c);
      load_col(WPh + c * kN, P + o.Wh, kN, kMw, 2 * cta + c);
    }
  if (has_pair) for (int j = tid; j < kVocab; j += kPT) WoR[j] = __ldg(P + o.Wo + (size_t)cta * kVocab + j);
  if (tid < kMaxB) { sy[tid] = -1; sfin[tid] = 0; src[tid] = beam ? 0 : tid; hs[tid] = 0.f; }
  if (tid == 0) { cnt[0] = beam ? 1 : Bn; cnt[1] = 0; }
  // zero the step-0 energy buffer and both alpha_past buffers (ownerless: strided over the grid)
  for (int i = cta * kPT + tid; i < npos; i += nC * kPT) { a.E[i] = 0.f; a.ap[i] = 0.f; a.ap[npos + i] = 0.f; }
  for (int i = cta * kPT + tid; i < Bn * 112; i += nC * kPT) a.LOg[i] = (i % 112) < kVocab ? __ldg(P + o.bo + (i % 112)) : 0.f;
  unsigned epoch = 0;
  long long tprev = clock64();
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto mark = [&](int ph) { const long long now = clock64(); pacc[ph] += now - tprev; tprev = now; };
  grid_barrier(a.barrier, nC, epoch);
  mark(7);

  int t = 0;
  for (; t < a.maxlen; ++t) {
    const float* hprev = (t == 0) ? a.h0 : a.hbuf + (size_t)((t - 1) & 1) * Bn * kN;
    float* hnew = a.hbuf + (size_t)(t & 1) * Bn * kN;
    float* Ecur = a.E + (size_t)(t & 1) * npos;
    float* Enext = a.E + (size_t)((t + 1) & 1) * npos;
    float* LOcur = a.LOg + (size_t)(t & 1) * Bn * 112;
    float* LOnext = a.LOg + (size_t)((t + 1) & 1) * Bn * 112;
    const float* ap_old = a.ap + (size_t)((t + 1) & 1) * npos;     // written by step t-1 (zero for t = 0)
    float* ap_new = a.ap + (size_t)(t & 1) * npos;
    const int rows = cnt[0];                                       // live rows of this step (greedy: Bn)
    // ====================================================== A: GRU-1 for the owned units
    {   // X[b] = [emb(y_b) | h_prev[b]]   (independent 128-bit loads)
      float4* X4 = reinterpret_cast<float4*>(X);
#pragma unroll 4
      for (int i = tid; i < rows * (2 * kN / 4); i += kPT) {
        const int b = i / (2 * kN / 4), k4 = i - b * (2 * kN / 4);
        float4 v = make4(0.f);
        if (k4 < kMw / 4) { const int y = sy[b]; if (y >= 0) v = ldg4(P + o.embed + (size_t)y * kMw + k4 * 4); }
        else v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)((t == 0 && shared_img) ? 0 : src[b]) * kN) + (k4 - kMw / 4));
        X4[i] = v;
      }
    }
    __syncthreads();
    for (int p = warp; p < nu * rows; p += kPT / 32) {
      const int s = p / rows, b = p - s * rows, j = unit[s];
      const float* xe = X + b * 2 * kN; const float* xh = xe + kMw;
      const float* w = WA + s * 6 * kN;
      float d6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};           // er ez eh | hr hz hu
#pragma unroll
      for (int k8 = 0; k8 < kN / 32; ++k8) {
        const int k = lane + 32 * k8;
        const float ve = xe[k], vh = xh[k];
        d6[0] = fmaf(ve, w[k], d6[0]); d6[1] = fmaf(ve, w[kN + k], d6[1]); d6[2] = fmaf(ve, w[2 * kN + k], d6[2]);
        d6[3] = fmaf(vh, w[3 * kN + k], d6[3]); d6[4] = fmaf(vh, w[4 * kN + k], d6[4]); d6[5] = fmaf(vh, w[5 * kN + k], d6[5]);
      }
      warp_sum_n(d6);
      if (lane == 0) {
        const float r = sigmoidf_(d6[0] + d6[3] + __ldg(P + o.b_yz_yr + j)), z = sigmoidf_(d6[1] + d6[4] + __ldg(P + o.b_yz_yr + kN + j));
        const float prop = tanhf(fmaf(d6[5], r, d6[2] + __ldg(P + o.b_yh + j)));
        const float hp = xh[j];
        const float ph = z * hp + (1.f - z) * prop;
        own[(s * kMaxB + b) * 8 + 0] = ph;
        a.pre_h[(size_t)b * kN + j] = ph;
      }
    }
    for (int i = cta * kPT + tid; i < npos; i += nC * kPT) Enext[i] = 0.f;     // energies of the NEXT step
    for (int i = cta * kPT + tid; i < Bn * 112; i += nC * kPT) LOnext[i] = (i % 112) < kVocab ? __ldg(P + o.bo + (i % 112)) : 0.f;
    grid_barrier(a.barrier, nC, epoch);
    mark(0);
    // ====================================================== B: pre_h-side products
    stage_cg(X, a.pre_h, rows * kN / 4);
    __syncthreads();
    for (int p = warp; p < nu * rows; p += kPT / 32) {            // GRU-2 pre_h-side: three dot products per (unit, row), one butterfly
      const int s = p / rows, b = p - s * rows, j = unit[s];
      const float* x = X + b * kN; const float* w = WB + s * 3 * kN;
      float d3[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k8 = 0; k8 < kN / 32; ++k8) {
        const int k = lane + 32 * k8;
        const float v = x[k];
        d3[0] = fmaf(v, w[k], d3[0]); d3[1] = fmaf(v, w[kN + k], d3[1]); d3[2] = fmaf(v, w[2 * kN + k], d3[2]);
      }
      warp_sum_n(d3);
      if (lane == 0) {
        float* ow = own + (s * kMaxB + b) * 8;
        ow[1] = d3[0] + __ldg(P + o.b_hz_hr_nl + j); ow[2] = d3[1] + __ldg(P + o.b_hz_hr_nl + kN + j); ow[3] = d3[2] + __ldg(P + o.b_rh_nl + j);
      }
    }
    for (int p = 2 * warp; p < nsp * rows; p += 2 * (kPT / 32)) {   // speller columns: two (column, row) items per pass
      float d2[2] = {0.f, 0.f};
      int ii[2], bb[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) { const int q = min(p + e, nsp * rows - 1); ii[e] = q / rows; bb[e] = q - ii[e] * rows; }
#pragma unroll
      for (int k8 = 0; k8 < kN / 32; ++k8) {
        const int k = lane + 32 * k8;
        d2[0] = fmaf(X[bb[0] * kN + k], WS[ii[0] * kN + k], d2[0]); d2[1] = fmaf(X[bb[1] * kN + k], WS[ii[1] * kN + k], d2[1]);
      }
      warp_sum_n(d2);
      if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (p + e < nsp * rows) { const int col = cta + ii[e] * nC; a.SP[(size_t)bb[e] * kNA + col] = d2[e] + __ldg(P + o.W_a_b + col); }
      }
    }
    grid_barrier(a.barrier, nC, epoch);
    mark(1);
    // ====================================================== C: coverage stencil + tanh + V_a dot (partial over 128 channels)
    // Register-tiled stencil (round 2; the round-1 form materialised a [121][64] im2col tile per pass): a warp takes a tile of 8 consecutive
    // positions of one image row, a lane 4 of the CTA's 128 channels; per filter row the 18-wide alpha_past window of the tile is read once
    // (broadcast loads) and reused by the 11 horizontal taps - 29 shared-memory loads per 352 FMAs, the loop runs at the FFMA issue rate.
    {
      const int W8 = (a.W + 7) & ~7, PW = W8 + 2 * (kCovK / 2), PH = a.H + 2 * (kCovK / 2);
      const int tiles_x = W8 >> 3, ntile = a.H * tiles_x;
      const int c0 = cg * kCGW + lane * 4;
      const float4 ub4 = ldg4(P + o.U_f_b + c0), va4 = ldg4(P + o.V_a + c0);
      if (c_active && cimg < rows) {
        const int b = cimg;
        const int t0 = (int)((long long)ntile * cgrp / gpi), t1 = (int)((long long)ntile * (cgrp + 1) / gpi);
        __syncthreads();
        for (int i = tid; i < PH * PW; i += kPT) {            // alpha_past of image b with a zero border of 5 (and zero columns up to W8)
          const int yy = i / PW - kCovK / 2, xx = i % PW - kCovK / 2;
          sal[i] = ((unsigned)yy < (unsigned)a.H && (unsigned)xx < (unsigned)a.W) ? __ldcg(ap_old + (size_t)src[b] * L + yy * a.W + xx) : 0.f;
        }
        __syncthreads();
        const float4 sp4 = __ldcg(reinterpret_cast<const float4*>(a.SP + (size_t)b * kNA + c0));
        for (int tb = t0; tb < t1; tb += kPT / 32) {
          const int tile = tb + warp;
          const bool act = tile < t1;
          const int ty0 = act ? tile / tiles_x : 0, x0 = act ? (tile - ty0 * tiles_x) * 8 : 0;
          float4 w4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {                       // all watch rows in flight before the stencil
            const int l = ty0 * a.W + min(x0 + i, a.W - 1);
            w4[i] = ldg4(a.watch + (size_t)(shared_img ? l : b * L + l) * kNA + c0);
          }
          float acc[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
          if (act) {
#pragma unroll 1
            for (int fy = 0; fy < kCovK; ++fy) {
              const float* prow = sal + (ty0 + fy) * PW + x0;
              float pv[18];
#pragma unroll
              for (int i = 0; i < 18; ++i) pv[i] = prow[i];
#pragma unroll
              for (int fx = 0; fx < kCovK; ++fx) {
                const float4 q = *reinterpret_cast<const float4*>(Qs + (fy * kCovK + fx) * kCGW + lane * 4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  acc[i][0] = fmaf(pv[i + fx], q.x, acc[i][0]); acc[i][1] = fmaf(pv[i + fx], q.y, acc[i][1]);
                  acc[i][2] = fmaf(pv[i + fx], q.z, acc[i][2]); acc[i][3] = fmaf(pv[i + fx], q.w, acc[i][3]);
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float e = va4.x * tanhf(acc[i][0] + ub4.x + w4[i].x + sp4.x) + va4.y * tanhf(acc[i][1] + ub4.y + w4[i].y + sp4.y) +
                      va4.z * tanhf(acc[i][2] + ub4.z + w4[i].z + sp4.z) + va4.w * tanhf(acc[i][3] + ub4.w + w4[i].w + sp4.w);
            e = warp_sum(e);
            if (lane == 0 && act && x0 + i < a.W) atomicAdd(Ecur + (size_t)b * L + ty0 * a.W + x0 + i, e);
          }
        }
      }
      __syncthreads();
    }
    grid_barrier(a.barrier, nC, epoch);
    mark(2);
    // ====================================================== D: softmax over L, alpha_past update, context chunk
    if (db < rows) {
      const float vab = __ldg(P + o.V_a_b);
      float mx = -INFINITY;
      for (int l = tid; l < L; l += kPT) { const float e = __ldcg(Ecur + (size_t)db * L + l) + vab; sal[l] = e; mx = fmaxf(mx, e); }
      mx = block_max(mx, red);
      float sum = 0.f;
      for (int l = tid; l < L; l += kPT) { const float u = expf(sal[l] - mx); sal[l] = u; sum += u; }
      sum = block_sum(sum, red);
      const float inv = 1.f / sum;
      for (int l = tid; l < L; l += kPT) {
        const float al = sal[l] * inv;
        sal[l] = al;
        if (dchunk == 0) ap_new[(size_t)db * L + l] = __ldcg(ap_old + (size_t)src[db] * L + l) + al;   // next read: phase C of step t+1
      }
      __syncthreads();
      const int cq = tid & 15, lg = tid >> 4;                  // 16 float4 columns x 16 position groups
      for (int q0 = 0; q0 < c4_per; q0 += 16) {
        const int c4 = dchunk * c4_per + q0 + cq;
        float4 acc = make4(0.f);
        if (q0 + cq < c4_per && c4 < kD / 4) {
          const float* base = a.ann + (size_t)(shared_img ? 0 : db) * L * kD + c4 * 4;
          for (int l = lg; l < L; l += 16) {
            const float4 v = ldg4(base + (size_t)l * kD);
            const float al = sal[l];
            acc.x = fmaf(al, v.x, acc.x); acc.y = fmaf(al, v.y, acc.y); acc.z = fmaf(al, v.z, acc.z); acc.w = fmaf(al, v.w, acc.w);
          }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&part[lg * 64 + cq * 4]) = acc;
        __syncthreads();
        if (tid < 64) {
          const int cc = (dchunk * c4_per + q0) * 4 + tid;
          if (q0 * 4 + tid < c4_per * 4 && cc < kD) {
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < 16; ++g) s += part[g * 64 + tid];
            a.ctx[(size_t)db * kD + cc] = s;
          }
        }
      }
    }
    grid_barrier(a.barrier, nC, epoch);
    mark(3);
    // ====================================================== E: GRU-2 for the owned units + emb/ctx part of the logit pair
    stage_cg(X, a.ctx, rows * kD / 4);
    __syncthreads();
    for (int p = warp; p < (nu + (has_pair ? 2 : 0)) * rows; p += kPT / 32) {
      const int q = p / rows, b = p - q * rows;
      const float* xc = X + b * kD;
      if (q < nu) {
        const int s = q, j = unit[s];
        const float* w = WE + s * 3 * kD;
        float d3[3] = {0.f, 0.f, 0.f};
        for (int k = lane; k < kD; k += 32) {
          const float v = xc[k];
          d3[0] = fmaf(v, w[k], d3[0]); d3[1] = fmaf(v, w[kD + k], d3[1]); d3[2] = fmaf(v, w[2 * kD + k], d3[2]);
        }
        warp_sum_n(d3);
        if (lane == 0) {
          const float* ow = own + (s * kMaxB + b) * 8;
          const float r = sigmoidf_(ow[1] + d3[0]), z = sigmoidf_(ow[2] + d3[1]);
          const float prop = tanhf(fmaf(ow[3], r, d3[2]));
          hnew[(size_t)b * kN + j] = z * ow[0] + (1.f - z) * prop;
        }
      } else {
        const int c = q - nu, n = 2 * cta + c;
        float d2[2] = {0.f, 0.f};                               // ctx part | embedding part (gathered from the table)
        const int y = sy[b];
        for (int k = lane; k < kD; k += 32) d2[0] = fmaf(xc[k], WPc[c * kD + k], d2[0]);
        if (y >= 0) for (int k = lane; k < kMw; k += 32) d2[1] = fmaf(__ldg(P + o.embed + (size_t)y * kMw + k), WPy[c * kMw + k], d2[1]);
        warp_sum_n(d2);
        if (lane == 0) own[(c * kMaxB + b) * 8 + 4] = d2[0] + d2[1] + __ldg(P + o.by + n) + __ldg(P + o.bc + n) + __ldg(P + o.bh + n);
      }
    }
    grid_barrier(a.barrier, nC, epoch);
    mark(4);
    // ====================================================== F: h part of the logit pair, maxout
    if (has_pair) {
      stage_cg(X, hnew, rows * kN / 4);
      __syncthreads();
      for (int p = warp; p < 2 * rows; p += kPT / 32) {
        const int c = p / rows, b = p - c * rows;
        const float v = warp_dot(X + b * kN, WPh + c * kN, kN, lane);
        if (lane == 0) own[(c * kMaxB + b) * 8 + 5] = v + own[(c * kMaxB + b) * 8 + 4];
      }
      __syncthreads();
      // maxout of the owned pair, then its rank-1 contribution MO[b,cta] * Wo[cta,:] to the output logits (RED.ADD)
      for (int i = tid; i < rows * kVocab; i += kPT) {
        const int b = i / kVocab, j = i - b * kVocab;
        const float m = fmaxf(own[(0 * kMaxB + b) * 8 + 5], own[(1 * kMaxB + b) * 8 + 5]);
        atomicAdd(LOcur + b * 112 + j, m * WoR[j]);
      }
    }
    grid_barrier(a.barrier, nC, epoch);
    mark(5);
    // ====================================================== G: argmax + stop rule (redundant in every CTA; logits were accumulated in F)
    stage_cg(LO, LOcur, Bn * 112 / 4);
    __syncthreads();
    if (!beam) {
      for (int b = warp; b < Bn; b += kPT / 32) {
        float mx = -INFINITY; int am = 0x7fffffff;
        for (int j = lane; j < kVocab; j += 32) { const float v = LO[b * 112 + j]; if (v > mx) { mx = v; am = j; } }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const float om = __shfl_xor_sync(0xffffffffu, mx, off);
          const int oa = __shfl_xor_sync(0xffffffffu, am, off);
          if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
        }
        if (lane == 0) {
          if (!sfin[b]) {
            if (cta == 0) a.tokens[(size_t)b * a.maxlen + t] = am;
            if ((am == 0 && !a.ignore_eol) || t == a.maxlen - 1) { sfin[b] = 1; if (cta == 0) a.lengths[b] = t + 1; }
          }
          sy[b] = am;
        }
      }
      __syncthreads();
      mark(6);
      int alive = 0;
      for (int b = 0; b < Bn; ++b) alive += sfin[b] ? 0 : 1;
      if (alive == 0) { ++t; break; }
    } else {
      // ---- beam bookkeeping of get_sample (S:485-533).  nr = k - dead_k candidates survive this step.
      const int live = rows, dead = cnt[1], nr = a.beam_k - dead;
      // (1) per live row: cost[j] = hyp_score + (logsumexp - logit[j]) = hyp_score - log softmax[j]; its nr smallest, ascending
      for (int b = warp; b < live; b += kPT / 32) {
        float v[4]; float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int j = lane + 32 * i; v[i] = j < kVocab ? LO[b * 112 + j] : -INFINITY; mx = fmaxf(mx, v[i]); }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) sum += (lane + 32 * i < kVocab) ? expf(v[i] - mx) : 0.f;
        sum = warp_sum(sum);
        const float base = hs[b] + mx + logf(sum);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (lane + 32 * i < kVocab) ? base - v[i] : INFINITY;
        for (int r = 0; r < nr; ++r) {
          float m = v[0]; int mi = lane;
#pragma unroll
          for (int i = 1; i < 4; ++i) if (v[i] < m) { m = v[i]; mi = lane + 32 * i; }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m, off);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
            if (om < m || (om == m && oi < mi)) { m = om; mi = oi; }
          }
          if ((mi & 31) == lane) {                                 // remove the winner from its owner's registers
#pragma unroll
            for (int i = 0; i < 4; ++i) if ((mi >> 5) == i) v[i] = INFINITY;
          }
          if (lane == 0) { cc[b * 16 + r] = m; ci[b * 16 + r] = b * kVocab + mi; }
        }
      }
      __syncthreads();
      // (2) merge the live x nr per-row winners: the nr smallest overall, in rank order (flat.argsort()[:k-dead_k], S:488-489)
      if (warp == 0) {
        float v[8]; int vi[8];
        const int n = live * nr;                                   // <= 16 * 16
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int q = lane + 32 * i;
          const bool ok = q < n;
          const int b = ok ? q / nr : 0, r = ok ? q - b * nr : 0;
          v[i] = ok ? cc[b * 16 + r] : INFINITY; vi[i] = ok ? ci[b * 16 + r] : 0x7fffffff;
        }
        for (int r = 0; r < nr; ++r) {
          float m = v[0]; int mi = vi[0]; int slot = 0;
#pragma unroll
          for (int i = 1; i < 8; ++i) if (v[i] < m || (v[i] == m && vi[i] < mi)) { m = v[i]; mi = vi[i]; slot = i; }
          float bm = m; int bi = mi;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, bm, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (om < bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
          }
          if (bi == mi && bm == m && mi != 0x7fffffff) {          // flat indices are unique: exactly one lane owns the winner
#pragma unroll
            for (int i = 0; i < 8; ++i) if (i == slot) { v[i] = INFINITY; vi[i] = 0x7fffffff; }
          }
          if (lane == 0) { selc[r] = bm; seli[r] = bi; }
        }
        __syncwarp();
        // (3) split into finished / continuing hypotheses (S:500-519), one thread: <= 16 entries
        if (lane == 0) {
          int nl = 0, nd = dead;
          int nsrc[16], nw[16]; float nsc[16];
          for (int r = 0; r < nr; ++r) {
            const int par = seli[r] / kVocab, word = seli[r] - par * kVocab;
            if (word == 0) { dead_t[nd] = t; dead_p[nd] = par; dead_s[nd] = selc[r]; ++nd; }
            else { nsrc[nl] = par; nw[nl] = word; nsc[nl] = selc[r]; ++nl; }
          }
          for (int i = 0; i < nl; ++i) {
            src[i] = nsrc[i]; sy[i] = nw[i]; hs[i] = nsc[i];
            if (t < kBeamMaxLen) { bk[t * 16 + i] = (unsigned char)nsrc[i]; wd[t * 16 + i] = (unsigned char)nw[i]; }
          }
          cnt[0] = nl; cnt[1] = nd;
        }
      }
      __syncthreads();
      mark(6);
      if (cnt[0] < 1 || cnt[1] >= a.beam_k) { ++t; break; }        // S:521-524
    }
  }
  if (beam && cta == 0) {
    // ---- samples = finished hypotheses in order of death, then the rows still alive (S:535-540); unroll the back-pointers;
    //      pick argmin(score / len) (S:671-672).  One thread per sample, then thread 0 picks.
    __syncthreads();
    const int nd = cnt[1], nl = cnt[0], ns = nd + nl, tl = t - 1;       // tl = last executed step
    if (tid < ns) {
      int len, row, tt;
      int* out = a.tokens + (size_t)tid * a.maxlen;
      if (tid < nd) { len = dead_t[tid] + 1; out[len - 1] = 0; row = dead_p[tid]; tt = dead_t[tid] - 1; }
      else { len = tl + 1; row = tid - nd; tt = tl; }
      for (; tt >= 0; --tt) { out[tt] = wd[tt * 16 + row]; row = bk[tt * 16 + row]; }
      a.lengths[tid] = len;
      a.bscore[tid] = tid < nd ? dead_s[tid] : hs[tid - nd];
      cc[tid] = a.bscore[tid] / (float)len;
    }
    __syncthreads();
    if (tid == 0) {
      int best = 0;
      for (int i = 1; i < ns; ++i) if (cc[i] < cc[best]) best = i;
      a.bmeta[0] = ns; a.bmeta[1] = best; a.bmeta[2] = t; a.bmeta[3] = nl;
    }
  }
  if (cta == 0 && tid == 0 && a.steps_run != nullptr) *a.steps_run = t;
  if (cta == 0 && tid == 0 && a.prof != nullptr)
    for (int i = 0; i < 8; ++i) a.prof[i] = pacc[i];
}

}  // namespace

size_t greedy_persist_ws_floats(int Bn, int L) {
  size_t o = 0;
  auto take = [&](size_t n) { o += (n + 3) & ~(size_t)3; };
  take((size_t)2 * Bn * kN); take((size_t)Bn * kN); take((size_t)Bn * kNA); take((size_t)2 * Bn * L); take((size_t)2 * Bn * L);
  take((size_t)Bn * kD); take((size_t)Bn * (kMw / 2)); take(8 * 160 + 24); take((size_t)2 * Bn * 112);
  take((size_t)Bn * L * kNA); take((size_t)kCovTaps * kNA); take((size_t)Bn * kN); take((size_t)Bn * kD); take(4);
  return o;
}

// Shared launcher.  beam_k == 0: greedy decode of Bn <= 16 images.  beam_k > 0: beam search over ONE image (ann [1,h,w,684]) with
// beam_k hypothesis rows.  Plus the hoisted precompute.  Never synchronises the host.
static int launch_persistent(cudaStream_t st, const float* params, const float* ann, int Bn, int h, int w, int maxlen, int ignore_eol, int beam_k,
                             int* tokens, int* lengths, float* bscore, int* bmeta, float* wsp, int* steps_run_dev) {
  const int L = h * w;
  WAP_REQUIRE(Bn >= 1 && Bn <= kMaxB && L <= kMaxL && (h + 10) * (((w + 7) & ~7) + 10) <= kPadMax,
              "persistent decoder: rows=%d (<=%d), L=%d (<=%d), padded grid %d (<=%d)", Bn, kMaxB, L, kMaxL, (h + 10) * (((w + 7) & ~7) + 10), kPadMax);
  static int num_ctas = 0;
  const size_t smem = sizeof(float) * kSmemFloats;
  if (num_ctas == 0) {
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    WAP_CHECK(cudaGetDevice(&dev));
    WAP_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WAP_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    WAP_REQUIRE(coop != 0, "device does not support cooperative launch");
    WAP_CHECK(cudaFuncSetAttribute(greedy_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WAP_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_persistent_kernel, kPT, smem));
    WAP_REQUIRE(per_sm >= 1, "persistent decoder does not fit on an SM (smem %zu B)", smem);
    num_ctas = std::min(160, (sms / kCG) * kCG);
    WAP_REQUIRE(num_ctas >= kMinCtas, "persistent decoder needs >= %d SMs (have %d)", kMinCtas, sms);
  }
  const int nimg = beam_k > 0 ? 1 : Bn;             // images behind the rows
  size_t o = 0;
  auto take = [&](size_t n) { float* p = wsp + o; o += (n + 3) & ~(size_t)3; return p; };
  PersistArgs a;
  a.params = params; a.ann = ann; a.Bn = Bn; a.H = h; a.W = w; a.L = L; a.maxlen = maxlen; a.ignore_eol = ignore_eol;
  a.beam_k = beam_k; a.bscore = bscore; a.bmeta = bmeta;
  a.hbuf = take((size_t)2 * Bn * kN); a.pre_h = take((size_t)Bn * kN); a.SP = take((size_t)Bn * kNA);
  a.E = take((size_t)2 * Bn * L); a.ap = take((size_t)2 * Bn * L); a.ctx = take((size_t)Bn * kD); a.MO = take((size_t)Bn * (kMw / 2));
  a.barrier = reinterpret_cast<unsigned*>(take(8 * 160 + 24));   // barrier counter (+ padding), then the profile words
  a.LOg = take((size_t)2 * Bn * 112);
  float* watch = take((size_t)Bn * L * kNA); float* qf = take((size_t)kCovTaps * kNA); float* h0 = take((size_t)Bn * kN);
  float* mean = take((size_t)Bn * kD);
  a.watch = watch; a.qf = cur()->qf_cached != nullptr ? cur()->qf_cached : qf; a.h0 = h0; a.tokens = tokens; a.lengths = lengths; a.steps_run = steps_run_dev;
  a.prof = reinterpret_cast<long long*>(a.barrier + 8 * 160);   // 8 x int64 after the barrier slots
  const ParamLayout& P = layout();
  DecOff& d = a.o;
  d.U_f_b = (int)P.U_f_b; d.W_a = (int)P.W_a; d.W_a_b = (int)P.W_a_b; d.V_a = (int)P.V_a; d.V_a_b = (int)P.V_a_b;
  d.W_yz_yr = (int)P.W_yz_yr; d.b_yz_yr = (int)P.b_yz_yr; d.U_hz_hr = (int)P.U_hz_hr; d.W_yh = (int)P.W_yh; d.b_yh = (int)P.b_yh;
  d.U_rh = (int)P.U_rh; d.U_hz_hr_nl = (int)P.U_hz_hr_nl; d.b_hz_hr_nl = (int)P.b_hz_hr_nl; d.W_c_z_r = (int)P.W_c_z_r;
  d.U_rh_nl = (int)P.U_rh_nl; d.b_rh_nl = (int)P.b_rh_nl; d.W_c_h_nl = (int)P.W_c_h_nl; d.embed = (int)P.embed; d.Wc = (int)P.Wc;
  d.bc = (int)P.bc; d.Wh = (int)P.Wh; d.bh = (int)P.bh; d.Wy = (int)P.Wy; d.by = (int)P.by; d.Wo = (int)P.Wo; d.bo = (int)P.bo;
  WAP_TRY(attn_precompute(st, params, ann, nullptr, nimg, h, w, watch, h0, cur()->qf_cached != nullptr ? nullptr : qf, mean));     // hidden_state_0 is unmasked (S:590)
  WAP_CHECK(cudaMemsetAsync(a.barrier, 0, (8 * 160 + 24) * sizeof(unsigned), st));
  WAP_CHECK(cudaMemsetAsync(lengths, 0, sizeof(int) * Bn, st));
  void* args[] = {(void*)&a};
  WAP_CHECK(cudaLaunchCooperativeKernel((const void*)greedy_persistent_kernel, dim3(num_ctas), dim3(kPT), args, smem, st));
  ++g_launches;
  return 0;
}

int greedy_decode_cluster(cudaStream_t st, const float* params, const float* ann, int Bn, int h, int w, int maxlen, int ignore_eol,
                          int* tokens, int* lengths, float* wsp, int* steps_run_dev);
int g_greedy_impl = 0;      // wap_debug_set_greedy_impl: 0 = cluster-per-image kernel when the shape allows (default), 1 = always the grid-wide persistent kernel

int greedy_decode_persistent(cudaStream_t st, const float* params, const float* ann, int Bn, int h, int w, int maxlen, int ignore_eol,
                             int* tokens, int* lengths, float* wsp, int* steps_run_dev) {
  // One 16-CTA cluster per image, no grid barrier: 25 us per token step however many images run side by side - but only 7 such
  // clusters are co-resident on a B200 (cudaOccupancyMaxActiveClusters), so a batch of 8+ images would take a second wave (2 x 25 us
  // > the 44.5 us per step of the grid-wide kernel below, whose step time does not grow up to 16 images).
  if (g_greedy_impl == 0 && Bn <= 7) {
    const int r = greedy_decode_cluster(st, params, ann, Bn, h, w, maxlen, ignore_eol, tokens, lengths, wsp, steps_run_dev);
    if (r <= 0) return r;       // 0 done, < 0 error; 1: shape / device not supported by the cluster kernel
  }
  return launch_persistent(st, params, ann, Bn, h, w, maxlen, ignore_eol, 0, tokens, lengths, nullptr, nullptr, wsp, steps_run_dev);
}

// WAP.get_sample(k, maxlen, stochastic=False) + the length-normalised pick, wholly on the device (see the file header).
// tokens [k][maxlen], lengths [k], scores [k]: the samples (finished first, in order of death; then the live rows); meta [4] =
// {number of samples, index of the picked one, steps executed, live rows at the end}.
int beam_decode_persistent(cudaStream_t st, const float* params, const float* ann, int h, int w, int k, int maxlen,
                           int* tokens, int* lengths, float* scores, int* meta, float* wsp) {
  WAP_REQUIRE(k >= 1 && k <= kMaxB && maxlen >= 1 && maxlen <= kBeamMaxLen, "beam_decode_persistent: k=%d (1..%d), maxlen=%d (1..%d)", k, kMaxB, maxlen,
              kBeamMaxLen);
  return launch_persistent(st, params, ann, k, h, w, maxlen, 0, k, tokens, lengths, scores, meta, wsp, nullptr);
}

}  // namespace wap
